/* inspect_logic.c - `yak inspect` (reference inspect.c:8-106) as plain host logic over a batched lookup.
 * One file: histogram of the counts in a .yak file, streamed.  Two files: for every key of in1.yak the count the second
 * table returns for it, as a 1024 x 1024 contingency table, printed as the reference's SN / QV lines.  The reference calls
 * yak_ch_get(ch, key) once per key with the STORED key (id << 10 | count) where a hash is expected (inspect.c:57, SURVEY
 * quirk Q7): sub-table = low `pre` bits of that word, id = word >> pre.  The same words go to `lookup` here, in batches -
 * the CLI passes yakb_ch_get_batch (one kernel per batch), the CPU test passes the oracle. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "yak.h"

typedef int (*yakb_lookup_f)(void *ctx, uint64_t n, const uint64_t *x, int32_t *out);
int yak_qv_solve(const int64_t *hist, const int64_t *cnt, int kmer, double fpr, yak_qstat_t *qs); /* qv_solve.c */

#define N YAK_N_COUNTS

/* hist2 == NULL: one-file mode.  Returns 0, or 1 with a message on stderr (the reference asserts instead). */
int yakb_inspect_run(FILE *out, const char *fn1, int max_cnt, const int64_t *hist2, yakb_lookup_f lookup, void *ctx, int64_t batch)
{
	FILE *fp;
	char magic[4];
	uint32_t t[3], u[2];
	int64_t tot[N], *cnt = 0, acc_tot = 0, n_buf = 0;
	uint64_t *keys = 0;
	int32_t *got = 0;
	int i, j, n_sub, kmer, rc = 0;
	const double fpr = 0.00004; /* inspect.c:15 */
	if ((fp = fopen(fn1, "rb")) == 0) { fprintf(stderr, "ERROR: failed to open '%s'\n", fn1); return 1; }
	if (fread(magic, 1, 4, fp) != 4 || memcmp(magic, YAK_MAGIC, 4) != 0 || fread(t, 4, 3, fp) != 3 || t[2] != YAK_COUNTER_BITS) {
		fprintf(stderr, "ERROR: not a .yak file\n");
		fclose(fp);
		return 1;
	}
	kmer = (int)t[0];
	n_sub = 1 << t[1];
	memset(tot, 0, sizeof(tot));
	if (batch < 1) batch = 1;
	if (hist2) {
		cnt = (int64_t*)calloc((size_t)N * N, sizeof(int64_t)); /* cnt[in1][in2] */
		keys = (uint64_t*)malloc(batch * sizeof(uint64_t));
		got = (int32_t*)malloc(batch * sizeof(int32_t));
	}
#define FLUSH() do { \
		if (n_buf > 0) { \
			int64_t q; \
			if (lookup(ctx, (uint64_t)n_buf, keys, got) != 0) { fprintf(stderr, "ERROR: lookup failed\n"); rc = 1; } \
			else for (q = 0; q < n_buf; ++q) ++cnt[(keys[q] & YAK_MAX_COUNT) * N + (got[q] < 0 ? 0 : got[q])]; /* inspect.c:58 */ \
			n_buf = 0; \
		} } while (0)
	for (i = 0; i < n_sub && rc == 0; ++i) {
		uint32_t left;
		if (fread(u, 4, 2, fp) != 2) break;
		for (left = u[1]; left > 0 && rc == 0;) {
			uint64_t tmp[4096];
			size_t m = left < 4096 ? left : 4096, r, q;
			r = fread(tmp, 8, m, fp);
			for (q = 0; q < r; ++q) {
				++tot[tmp[q] & YAK_MAX_COUNT];
				if (hist2) { keys[n_buf++] = tmp[q]; if (n_buf == batch) FLUSH(); }
			}
			if (r != m) { i = n_sub; break; } /* truncated file */
			left -= (uint32_t)m;
		}
	}
	fclose(fp);
	if (hist2 && rc == 0) FLUSH();
	if (rc == 0 && hist2) { /* inspect.c:66-95 */
		int64_t *acc = (int64_t*)malloc((size_t)N * N * sizeof(int64_t)), acc_cnt[N];
		if (max_cnt > N - 1) max_cnt = N - 1; /* the reference indexes acc_cnt[1..max_cnt] */
		memcpy(acc, cnt, (size_t)N * N * sizeof(int64_t));
		memset(acc_cnt, 0, sizeof(acc_cnt));
		for (j = N - 2; j >= 1; --j)
			for (i = 0; i < N; ++i) acc[i * N + j] += acc[i * N + (j + 1)];
		for (i = N - 1; i >= 0; --i) {
			acc_tot += tot[i];
			if (acc_tot == 0 || tot[i] == 0) continue;
			fprintf(out, "SN\t%d\t%ld\t%ld", i, (long)tot[i], (long)hist2[i]);
			for (j = 1; j <= max_cnt; ++j) {
				acc_cnt[j] += acc[i * N + j];
				fprintf(out, "\t%.4f", (double)acc_cnt[j] / acc_tot);
			}
			fprintf(out, "\n");
		}
		memcpy(acc, cnt, (size_t)N * N * sizeof(int64_t));
		for (i = N - 2; i >= 0; --i)
			for (j = 0; j < N; ++j) acc[i * N + j] += acc[(i + 1) * N + j];
		for (i = max_cnt; i >= 1; --i) {
			yak_qstat_t qs;
			if (tot[i] == 0) continue;
			yak_qv_solve(hist2, &acc[i * N], kmer, fpr, &qs);
			fprintf(out, "QV\t%d\t%ld\t%ld\t%.3f\t%.3f\n", i, (long)qs.tot, (long)acc[i * N], qs.qv_raw, qs.qv);
		}
		free(acc);
	} else if (rc == 0) { /* inspect.c:96-103; the hash-table column is 0 without a second file */
		for (i = N - 1; i >= 0; --i) {
			acc_tot += tot[i];
			if (acc_tot == 0) continue;
			fprintf(out, "HS\t%d\t%ld\t%ld\t%ld\n", i, 0L, (long)tot[i], (long)acc_tot);
		}
	}
	free(cnt); free(keys); free(got);
	return rc;
}
