/* qv_solve.c - DERIVED FROM the reference, NOT new work: yak_qv_solve (reference qv.c:146-244) and its 3x3 solver
 * (6gjdn.c) restated statement by statement with renamed variables.  SURVEY section 2 marks this component OUT of scope
 * ("keep the reference C unchanged"): it is O(1024) host FP64 arithmetic behind the scan, and whoever links libyakb200
 * into the reference's own `yak` keeps the reference's qv.c / 6gjdn.c for it.  The file exists only so
 * that the stand-alone yak-b200 command line can print the CT/FR/ER/CV/QV lines; the operation order is the reference's
 * because the printed FP64 values must agree digit for digit.  It is no part of the hot path and claims no credit.
 */
#include <math.h>
#include <stdio.h>
#include <string.h>
#include "yak.h"

#define QV_SCALE 4.3429448190325175 /* 10 / ln(10) */
#define FIT_DEG 2                   /* quadratic fit of adjacent-count ratios */
#define FIT_N (FIT_DEG + 1)

/* Gauss-Jordan elimination with full pivoting on an n x n system with one right-hand side
 * (6gjdn.c:5-88 with m = 1).  Column swaps are undone on the solution at the end. */
static int solve_full_pivot(double *a, double *b, int n)
{
	int col_of[FIT_N], k, i, j;
	for (k = 0; k < n; ++k) {
		double best = 0.0, piv;
		int pr = k, pc = k;
		for (i = k; i < n; ++i)
			for (j = k; j < n; ++j)
				if (fabs(a[i * n + j]) > best) best = fabs(a[i * n + j]), pr = i, pc = j;
		if (best + 1.0 == 1.0) { fprintf(stderr, "ERROR: fail\n"); return -1; }
		col_of[k] = pc;
		if (pc != k)
			for (i = 0; i < n; ++i) { double t = a[i * n + k]; a[i * n + k] = a[i * n + pc]; a[i * n + pc] = t; }
		if (pr != k) {
			double t;
			for (j = k; j < n; ++j) { t = a[k * n + j]; a[k * n + j] = a[pr * n + j]; a[pr * n + j] = t; }
			t = b[k]; b[k] = b[pr]; b[pr] = t;
		}
		piv = a[k * n + k];
		for (j = k + 1; j < n; ++j) a[k * n + j] = a[k * n + j] / piv;
		b[k] = b[k] / piv;
		for (j = k + 1; j < n; ++j)
			for (i = 0; i < n; ++i)
				if (i != k) a[i * n + j] = a[i * n + j] - a[i * n + k] * a[k * n + j];
		for (i = 0; i < n; ++i)
			if (i != k) b[i] = b[i] - a[i * n + k] * b[k];
	}
	for (k = n - 1; k >= 0; --k)
		if (col_of[k] != k) { double t = b[k]; b[k] = b[col_of[k]]; b[col_of[k]] = t; }
	return 0;
}

int yak_qv_solve(const int64_t *hist, const int64_t *cnt, int kmer, double fpr, yak_qstat_t *qs)
{
	int c, k, i, j, peak = -1, valley = -1, n_fit;
	int32_t peak_cnt = 0, valley_cnt;
	double x[8], y[8], pw[(2 * FIT_DEG + 1) * 8], A[FIT_N * FIT_N], B[FIT_N], adj_sum;

	memset(qs, 0, sizeof(*qs));
	qs->qv = -1.0, qs->err = cnt[0];
	for (c = 0; c < YAK_N_COUNTS; ++c) qs->tot += cnt[c], qs->adj_cnt[c] = cnt[c];
	qs->qv_raw = qs->tot > 0 && qs->tot > cnt[0] ? -QV_SCALE * log(log((double)qs->tot / (qs->tot - cnt[0])) / kmer) : -1.0;

	/* coverage peak of the query histogram (counts 2..1022) and the valley before it; the
	 * reference keeps these running extrema in 32-bit ints (qv.c:150,162-165) */
	for (c = 2; c < YAK_N_COUNTS - 1; ++c)
		if (peak_cnt < cnt[c]) peak_cnt = cnt[c], peak = c;
	for (c = 2, valley_cnt = peak_cnt; c < peak; ++c)
		if (valley_cnt > cnt[c]) valley_cnt = cnt[c], valley = c;
	/* no query k-mer with a count in 2..1022: the reference evaluates cnt[-1] / hist[-1] here (undefined; its binary prints
	 * -nan, i.e. 0/0, unless the table holds saturated k-mers).  Say 0/0 and be deterministic. */
	qs->cov = peak >= 0 ? (double)cnt[peak] / hist[peak] : 0.0 / (peak + 1.0);

	qs->fpr_upper = 1.0;
	for (c = 2; c < peak; ++c) {
		double e = cnt[c] / (qs->cov * hist[c]);
		if (qs->fpr_upper > e) qs->fpr_upper = e;
	}
	if (fpr > qs->fpr_upper) fpr = qs->fpr_upper * 0.5;
	qs->fpr_lower = 0.0;
	if (valley > 2 && hist[2] > hist[valley]) {
		double e = (cnt[2] - cnt[valley]) / (qs->cov * (hist[2] - hist[valley]));
		if (qs->fpr_lower < e) qs->fpr_lower = e;
	}
	if (fpr < qs->fpr_lower) fpr = qs->fpr_lower;
	if (qs->fpr_lower >= qs->fpr_upper)
		fprintf(stderr, "Warning: the FPR upper bound is smaller than the lower bound. Trust the lower bound.\n");

	if (peak <= 4) return -1; /* not high-coverage data: no adjusted QV */
	n_fit = peak - valley + 1 < 8 ? peak - valley + 1 : 8;
	if (n_fit < 3) return -1;

	for (c = peak - 1; c >= valley; --c) { /* remove the share explained by read errors */
		double e = (hist[c] - cnt[c] / qs->cov) / (1.0 - fpr);
		qs->adj_cnt[c] = cnt[c] - e * qs->cov * fpr;
		if (qs->adj_cnt[c] < 0.0) qs->adj_cnt[c] = 0.0;
	}
	for (k = 0; k < n_fit; ++k) {
		x[k] = valley + k;
		y[k] = qs->adj_cnt[valley + k + 1] / qs->adj_cnt[valley + k];
	}
	for (k = 0; k < n_fit; ++k) {
		double t = 1.0;
		for (i = 0; i <= 2 * FIT_DEG; ++i) pw[i * 8 + k] = t, t *= x[k];
	}
	for (i = 0; i <= FIT_DEG; ++i) { /* normal equations of the least-squares polynomial */
		double s;
		for (j = 0; j <= i; ++j) {
			for (k = 0, s = 0.0; k < n_fit; ++k) s += pw[(i + j) * 8 + k];
			A[i * FIT_N + j] = A[j * FIT_N + i] = s;
		}
		for (k = 0, s = 0.0; k < n_fit; ++k) s += pw[i * 8 + k] * y[k];
		B[i] = s;
	}
	solve_full_pivot(A, B, FIT_N);
	for (c = valley - 1; c >= 0; --c) { /* extrapolate the ratio below the valley */
		double r = 0.0, t = 1.0;
		for (i = 0; i <= FIT_DEG; ++i) r += B[i] * t, t *= c;
		if (r < 1.01) r = 1.01;
		qs->adj_cnt[c] = qs->adj_cnt[c + 1] / r;
	}
	for (c = 0, adj_sum = 0.0; c < YAK_N_COUNTS; ++c) adj_sum += qs->adj_cnt[c];
	if (adj_sum <= (double)qs->tot) {
		qs->err = qs->tot - adj_sum;
		qs->qv = -QV_SCALE * log(log(qs->tot / adj_sum) / kmer);
	} else {
		fprintf(stderr, "WARNING: failed to estimate the calibrated QV\n");
		qs->err = 0;
		qs->qv = qs->qv_raw;
	}
	return 0;
}
