/* scan_logic.c - see scan_logic.h.  Restated from the reference's worker functions; the k-mer roll and
 * the yak_ch_get calls of those workers are replaced by the per-position values of one batched
 * device lookup. */
#include <stdlib.h>
#include <string.h>
#include "scan_logic.h"

/* parental class of a k-mer from the 4 flag bits of a TRIOBIN table (htab.c:448-452 put the class of
 * file 1 in bits 0-1 and of file 2 in bits 2-3; class 2 = seen at least mid_cnt times):
 * 1 = file-1 specific, 2 = file-2 specific, 0 = neither (triobin.c:79-82, trioeval.c:78-81) */
static inline int hap_type(int flag)
{
	const int c1 = flag & 3, c2 = flag >> 2 & 3;
	return c1 == 2 && c2 == 0 ? 1 : c2 == 2 && c1 == 0 ? 2 : 0;
}

/* maximal runs of equal values in t[0..len): calls f(ctx, start, end, value) left to right */
typedef void (*run_fn)(void *ctx, int64_t st, int64_t en, int type);
static void for_each_run(const uint8_t *t, int64_t len, run_fn f, void *ctx)
{
	int64_t st = 0, i;
	for (i = 1; i <= len; ++i)
		if (i == len || t[i] != t[st]) { f(ctx, st, i, t[st]); st = i; }
}

static uint8_t *type_track(const int16_t *v, int64_t len, uint8_t **buf, int64_t *cap)
{
	int64_t i;
	if (len > *cap) { *cap = len + (len >> 1) + 64; *buf = (uint8_t*)realloc(*buf, *cap); }
	for (i = 0; i < len; ++i) (*buf)[i] = v[i] < 0 ? 0 : (uint8_t)hap_type(v[i]);
	return *buf;
}

/* ------------------------------------------------------------------ triobin */

typedef struct { int c[16], sc[2], nk; } tb_cnt_t;
typedef struct { tb_cnt_t *cnt; int k; } tb_run_ctx_t;

static void tb_run(void *ctx_, int64_t st, int64_t en, int type)
{
	tb_run_ctx_t *ctx = (tb_run_ctx_t*)ctx_;
	if (type > 0 && en - st >= ctx->k - 4) ctx->cnt->sc[type - 1] += (int)(en - st); /* triobin.c:92-96 */
}

/* triobin.c:103-124 */
static char tb_call(const tb_cnt_t *t, int k, double ratio_thres)
{
	const int sp = t->sc[0], sm = t->sc[1];
	const int p2 = t->c[2], m2 = t->c[8]; /* k-mers of class 2 in one parent and absent from the other */
	if (sp == 0 && sm == 0) {
		if (p2 == m2) return '0';
		if (p2 >= k - 4 + m2 && (m2 <= 1 || p2 * 0.05 > m2)) return 'p';
		if (m2 >= k - 4 + p2 && (p2 <= 1 || m2 * 0.05 > p2)) return 'm';
		return '0';
	}
	if (sp > k && sm > k) return 'a';
	if (sp >= k - 4 + sm && sp * 0.05 >= sm && p2 * ratio_thres > m2) return 'p';
	if (sm >= k - 4 + sp && sm * 0.05 >= sp && m2 * ratio_thres > p2) return 'm';
	return 'a';
}

void yakb_triobin_batch(FILE *out, const yakb_triobin_opt_t *opt, const yakb_scan_batch_t *b)
{
	tb_cnt_t *cnt = (tb_cnt_t*)calloc(b->n_seq > 0 ? b->n_seq : 1, sizeof(tb_cnt_t));
	uint8_t *buf = 0;
	int64_t cap = 0, s, off = 0;
	for (s = 0; s < b->n_seq; off += b->lens[s], ++s) { /* tb_worker, triobin.c:41-99 */
		const int16_t *v = b->vals + off;
		const int64_t len = b->lens[s];
		int64_t i;
		tb_run_ctx_t ctx;
		for (i = 0; i < len; ++i) {
			int flag;
			if (v[i] == -2) continue;
			flag = v[i] < 0 ? 0 : v[i];
			++cnt[s].nk;
			if (flag < 16) ++cnt[s].c[flag];
			if (opt->print_diff && (flag >> 2 & 3) != (flag & 3))
				fprintf(out, "D\t%s\t%d\t%d\t%d\n", b->names[s], (int)i, flag & 3, flag >> 2 & 3);
		}
		ctx.cnt = &cnt[s]; ctx.k = opt->k;
		for_each_run(type_track(v, len, &buf, &cap), len, tb_run, &ctx);
	}
	for (s = 0; s < b->n_seq; ++s) { /* triobin.c:140-146 */
		const int *c = cnt[s].c;
		fprintf(out, "%s\t%c\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\n", b->names[s], tb_call(&cnt[s], opt->k, opt->ratio_thres),
		        cnt[s].sc[0], cnt[s].sc[1], c[2], c[8], c[1], c[4], cnt[s].nk, c[0]);
	}
	free(buf); free(cnt);
}

/* ------------------------------------------------------------------ trioeval */

typedef struct { int nk, c[4], d[2]; } te_cnt_t;
typedef struct {
	FILE *out;
	const yakb_trioeval_opt_t *opt;
	const char *name;
	te_cnt_t *cnt;
	int last;                        /* class of the previous accepted streak, 0 = none yet */
	int f_type, f_st, f_en, f_cnt;   /* the open fragment: consecutive accepted streaks of one class */
} te_run_ctx_t;

static void te_flush_frag(te_run_ctx_t *x)
{
	if (x->f_type > 0 && x->opt->print_frag)
		fprintf(x->out, "F\t%s\t%d\t%d\t%d\t%d\n", x->name, x->f_type, x->f_st, x->f_en, x->f_cnt);
}

static void te_run(void *ctx_, int64_t st, int64_t en, int type) /* trioeval.c:92-115 */
{
	te_run_ctx_t *x = (te_run_ctx_t*)ctx_;
	const int k = x->opt->k;
	int n, c;
	if (type == 0 || en - st < x->opt->min_n) return; /* unclassified stretch, or a streak too short to trust */
	n = (int)((en - st + k - 1) / k);                  /* independent markers the streak stands for */
	c = type - 1;
	x->cnt->c[c << 1 | c] += n - 1;
	x->cnt->d[c] += n;
	if (x->last > 0) {
		++x->cnt->c[(x->last - 1) << 1 | c];
		if (x->opt->print_err && x->last - 1 != c)
			fprintf(x->out, "E\t%s\t%d\t%d\t%d\n", x->name, (int)en, x->last, c + 1);
	}
	if (x->f_type != type) {
		te_flush_frag(x);
		x->f_type = type; x->f_st = (int)st + 1 - k; x->f_cnt = 0;
	}
	++x->f_cnt; x->f_en = (int)en + 1;
	x->last = type;
}

void yakb_trioeval_header(FILE *out) /* trioeval.c:190-195 */
{
	fprintf(out, "C\tS  seqName     #patKmer  #matKmer  #pat-pat  #pat-mat  #mat-pat  #mat-mat  seqLen\n");
	fprintf(out, "C\tF  seqName     type      startPos  endPos    count\n");
	fprintf(out, "C\tW  #switchErr  denominator  switchErrRate\n");
	fprintf(out, "C\tH  #hammingErr denominator  hammingErrRate\n");
	fprintf(out, "C\tN  #totPatKmer #totMatKmer  errRate\n");
	fprintf(out, "C\n");
}

void yakb_trioeval_batch(FILE *out, const yakb_trioeval_opt_t *opt, const yakb_scan_batch_t *b, yakb_trioeval_sum_t *sum)
{
	te_cnt_t *cnt = (te_cnt_t*)calloc(b->n_seq > 0 ? b->n_seq : 1, sizeof(te_cnt_t));
	uint8_t *buf = 0;
	int64_t cap = 0, s, off = 0;
	for (s = 0; s < b->n_seq; off += b->lens[s], ++s) { /* te_worker, trioeval.c:40-117 */
		const int16_t *v = b->vals + off;
		const int64_t len = b->lens[s];
		int64_t i;
		te_run_ctx_t x;
		memset(&x, 0, sizeof(x));
		x.out = out; x.opt = opt; x.name = b->names[s]; x.cnt = &cnt[s];
		for (i = 0; i < len; ++i) if (v[i] != -2) ++cnt[s].nk;
		for_each_run(type_track(v, len, &buf, &cap), len, te_run, &x);
		te_flush_frag(&x);
	}
	for (s = 0; s < b->n_seq; ++s) { /* trioeval.c:136-147 */
		const int *c = cnt[s].c, *d = cnt[s].d;
		sum->n_par[0] += d[0];
		sum->n_par[1] += d[1];
		if (d[0] + d[1] >= 2) {
			sum->n_pair += c[0] + c[1] + c[2] + c[3];
			sum->n_switch += c[1] + c[2];
			sum->n_site += d[0] + d[1];
			sum->n_err += d[0] < d[1] ? d[0] : d[1];
		}
		fprintf(out, "S\t%s\t%d\t%d\t%d\t%d\t%d\t%d\t%d\n", b->names[s], d[0], d[1], c[0], c[1], c[2], c[3], (int)b->lens[s]);
	}
	free(buf); free(cnt);
}

void yakb_trioeval_footer(FILE *out, const yakb_trioeval_sum_t *sum) /* trioeval.c:204-206 */
{
	const int64_t lo = sum->n_par[0] < sum->n_par[1] ? sum->n_par[0] : sum->n_par[1];
	fprintf(out, "W\t%ld\t%ld\t%.6f\n", (long)sum->n_switch, (long)sum->n_pair, (double)sum->n_switch / sum->n_pair);
	fprintf(out, "H\t%ld\t%ld\t%.6f\n", (long)sum->n_err, (long)sum->n_site, (double)sum->n_err / sum->n_site);
	fprintf(out, "N\t%ld\t%ld\t%.6f\n", (long)sum->n_par[0], (long)sum->n_par[1], (double)lo / (sum->n_par[0] + sum->n_par[1]));
}

/* ------------------------------------------------------------------ chkerr */

void yakb_chkerr_batch(FILE *out, const yakb_chkerr_opt_t *opt, const yakb_scan_batch_t *b)
{
	int64_t s, off = 0;
	for (s = 0; s < b->n_seq; off += b->lens[s], ++s) { /* te_worker of chkerr.c:22-69 */
		const int16_t *v = b->vals + off;
		const int64_t len = b->lens[s];
		int64_t i, last = -1;
		int streak = 0;
		for (i = 0; i < len; ++i) {
			if (v[i] == -2 || v[i] >= opt->min_cnt) continue; /* an absent k-mer counts as -1 (chkerr.c:55) */
			if (i != last + 1) { /* a new stretch of low-count k-mers starts: report the one that ended */
				if (streak > opt->min_streak)
					fprintf(out, "%s\t%d\t%d\t%d\n", b->names[s], (int)(last + 1 - opt->k - (streak - 1)), (int)(last + 1), streak);
				streak = 1;
			} else ++streak;
			last = i;
		}
		if (streak > opt->min_streak)
			fprintf(out, "%s\t%d\t%d\t%d\n", b->names[s], (int)(last + 1 - opt->k - (streak - 1)), (int)(last + 1), streak);
	}
}

/* ------------------------------------------------------------------ sexchr */

void yakb_sexchr_header(FILE *out) /* sexchr.c:121-122 */
{
	fprintf(out, "C\tS  seqName  originalHap  0  #k-mer  #sexchr  #sex1-specifc  #sex2-specific\n");
	fprintf(out, "C\n");
}

void yakb_sexchr_batch(FILE *out, int hap, const yakb_scan_batch_t *b)
{
	int64_t s, off = 0;
	for (s = 0; s < b->n_seq; off += b->lens[s], ++s) { /* sc_worker, sexchr.c:28-72 */
		const int16_t *v = b->vals + off;
		long n_k = 0, n_any = 0, n_1 = 0, n_2 = 0;
		int64_t i;
		for (i = 0; i < b->lens[s]; ++i) {
			if (v[i] == -2) continue;
			++n_k;
			if (v[i] > 0) { ++n_any; n_1 += v[i] == 1; n_2 += v[i] == 2; }
		}
		fprintf(out, "S\t%s\t%d\t0\t%ld\t%ld\t%ld\t%ld\n", b->names[s], hap, n_k, n_any, n_1, n_2);
	}
}
