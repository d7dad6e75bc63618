/* scan_logic.h - per-sequence post-processing of the lookup scanners (reference triobin.c,
 * trioeval.c, chkerr.c, sexchr.c), separated from the lookups themselves.
 *
 * Every scanner of the reference rolls the k-mers of a sequence, calls yak_ch_get on each, and then
 * works on the per-position results only.  Here the lookups of a whole batch of sequences are one
 * device call (yakb_scan_seqs, include/yak_b200.h) that fills `vals`; these functions do the rest on
 * the host.  vals[i] for the k-mer ENDING at base i: -2 no k-mer ends here, -1 absent, else the
 * stored count / flag bits.  A batch is what bseq_read returns (bseq.c:33-57); lines come out in the
 * order of the reference run with -t1 (worker lines of the batch first, then its summary lines).
 *
 * Plain C with no dependency on the library, so the CPU tests can drive it with the oracle's lookups.
 */
#ifndef YAKB_SCAN_LOGIC_H
#define YAKB_SCAN_LOGIC_H
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	int64_t n_seq;
	char *const *names;
	const int64_t *lens;
	const int16_t *vals; /* concatenated, sum(lens) entries */
} yakb_scan_batch_t;

/* triobin.c:41-146 */
typedef struct { int k, print_diff; double ratio_thres; } yakb_triobin_opt_t;
void yakb_triobin_batch(FILE *out, const yakb_triobin_opt_t *opt, const yakb_scan_batch_t *b);

/* trioeval.c:40-151, 190-208 */
typedef struct { int k, min_n, print_err, print_frag; } yakb_trioeval_opt_t;
typedef struct { int64_t n_pair, n_site, n_switch, n_err, n_par[2]; } yakb_trioeval_sum_t;
void yakb_trioeval_header(FILE *out);
void yakb_trioeval_batch(FILE *out, const yakb_trioeval_opt_t *opt, const yakb_scan_batch_t *b, yakb_trioeval_sum_t *sum);
void yakb_trioeval_footer(FILE *out, const yakb_trioeval_sum_t *sum);

/* chkerr.c:22-69 */
typedef struct { int k, min_cnt, min_streak; } yakb_chkerr_opt_t;
void yakb_chkerr_batch(FILE *out, const yakb_chkerr_opt_t *opt, const yakb_scan_batch_t *b);

/* sexchr.c:28-72, 121-122 */
void yakb_sexchr_header(FILE *out);
void yakb_sexchr_batch(FILE *out, int hap, const yakb_scan_batch_t *b);

#ifdef __cplusplus
}
#endif
#endif
