/* yak.h - the drop-in boundary: the C API of lh3/yak (reference yak.h:1-109) as exported by
 * libyakb200.so.  Names, argument meaning, struct layouts and error conventions are the
 * reference's, because this header is what main.c / inspect.c / qv.c and friends compile against
 * when count.o htab.o bbf.o (and the scan loop of qv.o) are swapped for the B200 library; the
 * implementation behind every symbol is new (yak_b200/csrc).  Each block cites the reference
 * declaration it replaces.  GPU-only extensions live in yak_b200.h.
 */
#ifndef YAK_H
#define YAK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YAKS_VERSION "0.1-r93-b200"

/* reference yak.h:8-14 - counter width and bloom block geometry (fixed by the file format) */
#define YAK_MAX_KMER     31
#define YAK_COUNTER_BITS 10
#define YAK_N_COUNTS     (1<<YAK_COUNTER_BITS)
#define YAK_MAX_COUNT    ((1<<YAK_COUNTER_BITS)-1)
#define YAK_BLK_SHIFT    9
#define YAK_BLK_MASK     ((1<<(YAK_BLK_SHIFT)) - 1)

/* reference yak.h:16-21 - yak_ch_restore_core() modes; all six are implemented (htab.c:396-476: the
 * trio / sex-chromosome modes remap counts to flag bits while loading, csrc/capi.cu yak_ch_restore_core) */
#define YAK_LOAD_ALL       1
#define YAK_LOAD_TRIOBIN1  2
#define YAK_LOAD_TRIOBIN2  3
#define YAK_LOAD_SEXCHR1   4
#define YAK_LOAD_SEXCHR2   5
#define YAK_LOAD_SEXCHR3   6

#define YAK_MAGIC "YAK\2"   /* reference yak.h:23 */

/* reference yak.h:25-31: `yak count` options (defaults from yak_copt_init, misc.c:23-32) */
typedef struct {
	int32_t bf_shift, bf_n_hash;
	int32_t k;
	int32_t pre;
	int32_t n_thread;
	int64_t chunk_size;
} yak_copt_t;

/* reference yak.h:33-40: `yak qv` options (defaults from yak_qopt_init, qv.c:137-144) */
typedef struct {
	int32_t print_each, print_err_kmer;
	int32_t min_len;
	int32_t n_threads;
	double min_frac;
	double fpr;
	int64_t chunk_size;
} yak_qopt_t;

/* reference yak.h:42-47 */
typedef struct {
	int64_t tot;
	double qv_raw, qv, cov, err;
	double fpr_lower, fpr_upper;
	double adj_cnt[1<<YAK_COUNTER_BITS];
} yak_qstat_t;

/* reference yak.h:49-52.  In this library `b` of a filter owned by a yak_ch_t is a DEVICE
 * pointer (the sub-filter inside the table's bloom arena); filters from yak_bf_init() are
 * device-resident too.  Host code must not dereference it (the reference never does outside
 * bbf.c). */
typedef struct {
	int n_shift, n_hashes;
	uint8_t *b;
} yak_bf_t;

struct yak_ht_t; /* opaque, reference yak.h:54; here: a handle into the device table */

/* reference yak.h:56-59 */
typedef struct {
	struct yak_ht_t *h;
	yak_bf_t *b;
} yak_ch1_t;

/* reference yak.h:61-65: callers read k, pre and tot directly (qv.c:40-43, main.c:58,197) */
typedef struct {
	int k, pre, n_hash, n_shift;
	uint64_t tot;
	yak_ch1_t *h;
} yak_ch_t;

/* reference yak.h:67-70 */
typedef struct {
	uint64_t x;
	int c;
} yak_knt_t;

extern int yak_verbose;                       /* reference yak.h:72, sys.c:5 */
extern unsigned char seq_nt4_table[256];      /* reference yak.h:73, misc.c:4-21 */

void yak_copt_init(yak_copt_t *opt);          /* reference yak.h:75, misc.c:23-32 */

/* reference yak.h:77-79, bbf.c */
yak_bf_t *yak_bf_init(int n_shift, int n_hashes);
void yak_bf_destroy(yak_bf_t *b);
int yak_bf_insert(yak_bf_t *b, uint64_t hash);

/* reference yak.h:81-87, htab.c:13-100,353-367 */
yak_ch_t *yak_ch_init(int k, int pre, int n_hash, int n_shift);
void yak_ch_destroy(yak_ch_t *h);
void yak_ch_destroy_bf(yak_ch_t *h);
int yak_ch_insert_list(yak_ch_t *h, int create_new, int n, const uint64_t *a);
int yak_ch_get(const yak_ch_t *h, uint64_t x);
int yak_ch_inc(yak_ch_t *h, uint64_t x);
yak_knt_t *yak_ch_getseq(const yak_ch_t *h, int w, uint32_t *n);

/* reference yak.h:89-96, htab.c:102-347 */
void yak_ch_tighten(yak_ch_t *h);
void yak_ch_clear(yak_ch_t *h, int n_thread);
void yak_ch_hist(const yak_ch_t *h, int64_t cnt[YAK_N_COUNTS], int n_thread);
void yak_ch_shrink(yak_ch_t *h, int min, int max, int n_thread);
void yak_ch_merge(yak_ch_t *h0, yak_ch_t *h1, int min, int max, int n_thread, int pre_resize);
void yak_ch_setcnt(yak_ch_t *h, int cnt, int n_thread);
void yak_ch_subtract(yak_ch_t *h0, const yak_ch_t *h1, int n_thread);
void yak_ch_isec(yak_ch_t *h0, const yak_ch_t *h1, int n_thread);

/* reference yak.h:98-100, htab.c:373-481 */
int yak_ch_dump(const yak_ch_t *h, const char *fn);
yak_ch_t *yak_ch_restore(const char *fn);
yak_ch_t *yak_ch_restore_core(yak_ch_t *ch0, const char *fn, int mode, ...);

/* reference yak.h:102-103, count.c:147-193 */
yak_ch_t *yak_count(const char *fn, const yak_copt_t *opt, yak_ch_t *h0);
void yak_recount(const char *fn, yak_ch_t *h);

/* reference yak.h:105-107, qv.c:116-244.  yak_qv_solve (host FP64, off the hot path) is NOT
 * exported by this library: link the reference's qv.c:146-244 + 6gjdn.c for it;
 * the CLI carries its own restatement (yak_b200/cli/qv_solve.c) */
void yak_qopt_init(yak_qopt_t *opt);
void yak_qv(const yak_qopt_t *opt, const char *fn, const yak_ch_t *ch, int64_t *cnt);

#ifdef __cplusplus
}
#endif
#endif
