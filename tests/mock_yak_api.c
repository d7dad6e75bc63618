/* mock_yak_api.c - TEST INFRASTRUCTURE ONLY.  The yak.h entry points the CLI's host code (yak_b200/cli/main.c) calls, answered
 * by the CPU oracle (oracle/yak_oracle.c), so that the command flows - option parsing, the two-pass protocol, cntasm's
 * merge/shrink schedule, two-file inspect - can be run here, without a GPU, against the reference binary
 * (tests/test_cli_flows_cpu.py).  It is linked into a throw-away executable under the test's temp directory and never into
 * the product: libyakb200.so has no CPU path. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "yak.h"
#include "yak_b200.h"
#include "../oracle/yak_oracle.h"

int yak_verbose = 3;
#define O(h) ((yo_ch_t*)(h))

void yak_copt_init(yak_copt_t *o) { o->bf_shift = 0; o->bf_n_hash = 4; o->k = 31; o->pre = 10; o->n_thread = 4; o->chunk_size = 10000000; } /* misc.c:23-32 */
void yak_qopt_init(yak_qopt_t *o) { memset(o, 0, sizeof(*o)); o->chunk_size = 1000000000; o->n_threads = 4; o->min_frac = 0.5; o->fpr = 0.00004; }

yak_ch_t *yak_count(const char *fn, const yak_copt_t *opt, yak_ch_t *h0)
{
	return (yak_ch_t*)yo_count_file_chunked(fn, opt->k, opt->pre, opt->bf_shift, opt->bf_n_hash, O(h0), 0, opt->chunk_size);
}
void yak_recount(const char *fn, yak_ch_t *h) { yo_ch_clear(O(h)); yo_count_file_chunked(fn, h->k, h->pre, 0, 4, O(h), 0, 1LL << 62); }
void yak_ch_destroy(yak_ch_t *h) { if (h) yo_ch_destroy(O(h)); }
void yak_ch_destroy_bf(yak_ch_t *h) { yo_ch_destroy_bf(O(h)); }
void yak_ch_clear(yak_ch_t *h, int n) { (void)n; yo_ch_clear(O(h)); }
void yak_ch_hist(const yak_ch_t *h, int64_t cnt[YAK_N_COUNTS], int n) { (void)n; yo_ch_hist(O(h), cnt); }
void yak_ch_shrink(yak_ch_t *h, int min, int max, int n) { (void)n; yo_ch_shrink(O(h), min, max); }
void yak_ch_setcnt(yak_ch_t *h, int cnt, int n) { (void)n; yo_ch_setcnt(O(h), cnt); }
void yak_ch_merge(yak_ch_t *h0, yak_ch_t *h1, int min, int max, int n, int pre_resize) { (void)n; yo_ch_merge(O(h0), O(h1), min, max, pre_resize); }
void yak_ch_tighten(yak_ch_t *h) { yo_ch_tighten(O(h)); }
void yak_ch_subtract(yak_ch_t *h0, const yak_ch_t *h1, int n) { (void)n; yo_ch_subtract(O(h0), O(h1)); }
void yak_ch_isec(yak_ch_t *h0, const yak_ch_t *h1, int n) { (void)n; yo_ch_isec(O(h0), O(h1)); }
int yak_ch_dump(const yak_ch_t *h, const char *fn) { return yo_ch_dump(O(h), fn); }
yak_ch_t *yak_ch_restore(const char *fn) { return (yak_ch_t*)yo_ch_restore(fn); }
int yakb_ch_get_batch(const yak_ch_t *h, uint64_t n, const uint64_t *x, int32_t *out)
{
	uint64_t i;
	for (i = 0; i < n; ++i) out[i] = yo_ch_get(O(h), x[i]);
	return 0;
}
/* not exercised by the flow tests */
yak_knt_t *yak_ch_getseq(const yak_ch_t *h, int w, uint32_t *n) { (void)h; (void)w; *n = 0; return 0; }
/* qv.c:88-135 without -p / -E (those lines are printed inside the library): every record of the file in one batch */
void yak_qv(const yak_qopt_t *opt, const char *fn, const yak_ch_t *ch, int64_t *cnt)
{
	yo_reader_t *r = yo_reader_open(fn);
	const char *seq, *name;
	char *cat = 0;
	int64_t *lens = 0, n = 0, m = 0, tot = 0, cap = 0, len;
	memset(cnt, 0, YAK_N_COUNTS * sizeof(int64_t));
	if (r == 0) return;
	while ((len = yo_reader_next(r, &seq, &name)) >= 0) {
		if (n == m) { m = m ? m * 2 : 256; lens = (int64_t*)realloc(lens, m * sizeof(int64_t)); }
		if (tot + len + 1 > cap) { cap = (tot + len + 1) * 2; cat = (char*)realloc(cat, cap); }
		memcpy(cat + tot, seq, len);
		tot += len; lens[n++] = len;
	}
	yo_reader_close(r);
	yo_qv_seqs(O(ch), n, lens, cat ? cat : "", opt->min_len, opt->min_frac, cnt, 0, 0);
	free(cat); free(lens);
}
#ifdef MOCK_STUB_SCANNERS
int yakb_cmd_triobin(int argc, char *argv[]) { (void)argc; (void)argv; return 1; }
int yakb_cmd_trioeval(int argc, char *argv[]) { (void)argc; (void)argv; return 1; }
int yakb_cmd_chkerr(int argc, char *argv[]) { (void)argc; (void)argv; return 1; }
int yakb_cmd_sexchr(int argc, char *argv[]) { (void)argc; (void)argv; return 1; }
#else
/* the scanners of cli/scan.c: table loads with flag modes, one lookup call per batch of sequences */
#include <stdarg.h>
yak_ch_t *yak_ch_restore_core(yak_ch_t *ch0, const char *fn, int mode, ...)
{
	int min_cnt = 0, mid_cnt = 0;
	va_list ap;
	va_start(ap, mode);
	if (mode == YAK_LOAD_TRIOBIN1 || mode == YAK_LOAD_TRIOBIN2) { min_cnt = va_arg(ap, int); mid_cnt = va_arg(ap, int); }
	va_end(ap);
	return (yak_ch_t*)yo_ch_restore_core(O(ch0), fn, mode, min_cnt, mid_cnt, 0);
}
int yakb_scan_seqs(const yak_ch_t *h, int64_t n_seq, const int64_t *lens, const char *cat, int16_t *out)
{
	int64_t i, off = 0;
	for (i = 0; i < n_seq; off += lens[i++]) yo_scan_seq(O(h), lens[i], cat + off, out + off);
	return 0;
}
#endif
