/* yak-b200: the `yak` command line (count / recount / cntasm / subtract / isec / print / qv / inspect / version; scanners in scan.c) over libyakb200.so.
 * Plain host C calling the C ABI of include/yak.h; flags, defaults, messages and the two-pass
 * bloom protocol follow the reference CLI (main.c:13-64 count, 163-215 qv, 325-379 dispatch;
 * inspect.c:8-106).  Nothing here touches the GPU directly. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <assert.h>
#include <sys/time.h>
#include <sys/resource.h>
#include "yak.h"
#include "yak_b200.h"

int yak_qv_solve(const int64_t *hist, const int64_t *cnt, int kmer, double fpr, yak_qstat_t *qs); /* qv_solve.c */
int yakb_cmd_triobin(int argc, char *argv[]);  /* scan.c */
int yakb_cmd_trioeval(int argc, char *argv[]);
int yakb_cmd_chkerr(int argc, char *argv[]);
int yakb_cmd_sexchr(int argc, char *argv[]);
int64_t yakb_cli_parse_num(const char *s);

static double t_real0;
static double realtime(void) { struct timeval tv; gettimeofday(&tv, 0); return tv.tv_sec + tv.tv_usec * 1e-6; }
static double cputime(void) { struct rusage r; getrusage(RUSAGE_SELF, &r); return r.ru_utime.tv_sec + r.ru_stime.tv_sec + 1e-6 * (r.ru_utime.tv_usec + r.ru_stime.tv_usec); }
static long peakrss(void) { struct rusage r; getrusage(RUSAGE_SELF, &r); return r.ru_maxrss * 1024; }

/* "10k" / "1.5g" style numbers (reference yak-priv.h:75-84) */
int64_t yakb_cli_parse_num(const char *s)
{
	char *p;
	double x = strtod(s, &p);
	if (*p == 'G' || *p == 'g') x *= 1e9;
	else if (*p == 'M' || *p == 'm') x *= 1e6;
	else if (*p == 'K' || *p == 'k') x *= 1e3;
	return (int64_t)(x + .499);
}

static int cmd_count(int argc, char *argv[])
{
	yak_copt_t opt;
	yak_ch_t *h;
	char *fn_out = 0;
	int c;
	yak_copt_init(&opt);
	while ((c = getopt(argc, argv, "k:p:K:t:b:H:o:g:")) >= 0) {
		if (c == 'g') setenv("YAKB_GPUS", optarg, 1); /* ours: spread the table over this many GPUs (the library reads YAKB_GPUS) */
		else if (c == 'k') opt.k = atoi(optarg);
		else if (c == 'p') opt.pre = atoi(optarg);
		else if (c == 'K') opt.chunk_size = yakb_cli_parse_num(optarg);
		else if (c == 't') opt.n_thread = atoi(optarg);
		else if (c == 'b') opt.bf_shift = atoi(optarg);
		else if (c == 'H') opt.bf_n_hash = (int)yakb_cli_parse_num(optarg);
		else if (c == 'o') fn_out = optarg;
	}
	if (argc - optind < 1) {
		fprintf(stderr, "Usage: yak-b200 count [options] <in.fa> [in.fa]\n");
		fprintf(stderr, "Options:\n");
		fprintf(stderr, "  -k INT     k-mer size [%d]\n", opt.k);
		fprintf(stderr, "  -p INT     prefix length [%d]\n", opt.pre);
		fprintf(stderr, "  -b INT     set Bloom filter size to 2**INT bits; 0 to disable [%d]\n", opt.bf_shift);
		fprintf(stderr, "  -H INT     use INT hash functions for Bloom filter [%d]\n", opt.bf_n_hash);
		fprintf(stderr, "  -t INT     number of worker threads (accepted, unused on the GPU) [%d]\n", opt.n_thread);
		fprintf(stderr, "  -o FILE    dump the count hash table to FILE []\n");
		fprintf(stderr, "  -K INT     chunk size [100m]\n");
		fprintf(stderr, "  -g INT     number of GPUs to spread the sub-tables over (a power of two; also YAKB_GPUS) [1]\n");
		fprintf(stderr, "Note: -b37 is recommended for human reads\n");
		return 1;
	}
	if (opt.pre < YAK_COUNTER_BITS) { fprintf(stderr, "ERROR: -p should be at least %d\n", YAK_COUNTER_BITS); return 1; }
	if (opt.k >= 64) { fprintf(stderr, "ERROR: -k must be smaller than 64\n"); return 1; }
	else if (opt.k >= 32) fprintf(stderr, "WARNING: counts are inexact if -k is greater than 31\n");
	h = yak_count(argv[optind], &opt, 0);
	if (h == 0) { fprintf(stderr, "ERROR: failed to count '%s' (no such file, or no CUDA device)\n", argv[optind]); return 1; }
	if (opt.bf_shift > 0) { /* two-pass protocol, reference main.c:54-60 */
		yak_ch_destroy_bf(h);
		yak_ch_clear(h, opt.n_thread);
		h = yak_count(argc - optind >= 2 ? argv[optind + 1] : argv[optind], &opt, h);
		yak_ch_shrink(h, 2, YAK_MAX_COUNT, opt.n_thread);
		fprintf(stderr, "[M::%s] %ld distinct k-mers after shrinking\n", "main_count", (long)h->tot);
	}
	if (fn_out) yak_ch_dump(h, fn_out);
	yak_ch_destroy(h);
	return 0;
}

static int cmd_qv(int argc, char *argv[])
{
	yak_qopt_t opt;
	yak_ch_t *ch;
	int64_t cnt[YAK_N_COUNTS], hist[YAK_N_COUNTS];
	yak_qstat_t qs;
	int c, i, kmer;
	yak_qopt_init(&opt);
	while ((c = getopt(argc, argv, "K:t:l:f:pe:E")) >= 0) {
		if (c == 'K') opt.chunk_size = yakb_cli_parse_num(optarg);
		else if (c == 'l') opt.min_len = (int)yakb_cli_parse_num(optarg);
		else if (c == 'f') opt.min_frac = atof(optarg);
		else if (c == 't') opt.n_threads = atoi(optarg);
		else if (c == 'p') opt.print_each = 1;
		else if (c == 'E') opt.print_err_kmer = 1;
		else if (c == 'e') opt.fpr = atof(optarg);
	}
	if (argc - optind < 2) {
		fprintf(stderr, "Usage: yak-b200 qv [options] <kmer.hash> <seq.fa>\n");
		fprintf(stderr, "Options:\n");
		fprintf(stderr, "  -l NUM      min sequence length [%d]\n", opt.min_len);
		fprintf(stderr, "  -f FLOAT    min k-mer fraction [%g]\n", opt.min_frac);
		fprintf(stderr, "  -e FLOAT    false positive rate [%g]\n", opt.fpr);
		fprintf(stderr, "  -p          print QV for each sequence\n");
		fprintf(stderr, "  -E          print the positions of wrong k-mers\n");
		fprintf(stderr, "  -t INT      number of threads (accepted, unused on the GPU) [%d]\n", opt.n_threads);
		fprintf(stderr, "  -K NUM      batch size [1g]\n");
		return 1;
	}
	ch = yak_ch_restore(argv[optind]);
	if (ch == 0) { fprintf(stderr, "ERROR: failed to load '%s'\n", argv[optind]); return 1; }
	kmer = ch->k;
	yak_ch_hist(ch, hist, opt.n_threads);
	printf("CC\tCT  kmer_occurrence    short_read_kmer_count  raw_input_kmer_count  adjusted_input_kmer_count\n");
	printf("CC\tFR  fpr_lower_bound    fpr_upper_bound\n");
	printf("CC\tER  total_input_kmers  adjusted_error_kmers\n");
	printf("CC\tCV  coverage\n");
	printf("CC\tQV  raw_quality_value  adjusted_quality_value\n");
	printf("CC\n");
	yak_qv(&opt, argv[optind + 1], ch, cnt);
	yak_qv_solve(hist, cnt, kmer, opt.fpr, &qs);
	for (i = YAK_N_COUNTS - 1; i >= 0; --i)
		printf("CT\t%d\t%ld\t%ld\t%.3f\n", i, (long)hist[i], (long)cnt[i], qs.adj_cnt[i]);
	printf("FR\t%.3g\t%.3g\n", qs.fpr_lower, qs.fpr_upper);
	printf("ER\t%ld\t%.3f\n", (long)qs.tot, qs.err);
	printf("CV\t%.3f\n", qs.cov);
	printf("QV\t%.3f\t%.3f\n", qs.qv_raw, qs.qv);
	yak_ch_destroy(ch);
	return 0;
}

/* reference inspect.c:8-106; the logic is cli/inspect_logic.c, the lookups of the two-file mode one batched call each */
typedef int (*yakb_lookup_f)(void *ctx, uint64_t n, const uint64_t *x, int32_t *out);
int yakb_inspect_run(FILE *out, const char *fn1, int max_cnt, const int64_t *hist2, yakb_lookup_f lookup, void *ctx, int64_t batch);
static int inspect_lookup(void *ctx, uint64_t n, const uint64_t *x, int32_t *out) { return yakb_ch_get_batch((const yak_ch_t*)ctx, n, x, out); }

static int cmd_inspect(int argc, char *argv[])
{
	int c, max_cnt = 20, rc;
	int64_t hist[YAK_N_COUNTS];
	yak_ch_t *ch;
	while ((c = getopt(argc, argv, "m:")) >= 0)
		if (c == 'm') max_cnt = atoi(optarg);
	if (argc - optind < 1) {
		fprintf(stderr, "Usage: yak-b200 inspect [options] <in1.yak> [in2.yak]\n");
		fprintf(stderr, "Options:\n");
		fprintf(stderr, "  -m INT    max count (effective with in2.yak) [%d]\n", max_cnt);
		fprintf(stderr, "Notes: when in2.yak is present, inspect evaluates the k-mer QV of in1.yak and\n");
		fprintf(stderr, "  the k-mer sensitivity of in2.yak.\n");
		return 1;
	}
	if (argc - optind < 2) return yakb_inspect_run(stdout, argv[optind], max_cnt, 0, 0, 0, 1);
	if ((ch = yak_ch_restore(argv[optind + 1])) == 0) { fprintf(stderr, "ERROR: failed to load '%s'\n", argv[optind + 1]); return 1; }
	yak_ch_hist(ch, hist, 1);
	rc = yakb_inspect_run(stdout, argv[optind], max_cnt, hist, inspect_lookup, ch, 16 << 20);
	yak_ch_destroy(ch);
	return rc;
}

/* reference main.c:66-88: tighten, count the k-mers of a table again from other reads */
static int cmd_recount(int argc, char *argv[])
{
	yak_ch_t *h;
	char *fn_out = "-";
	int c;
	while ((c = getopt(argc, argv, "o:")) >= 0)
		if (c == 'o') fn_out = optarg;
	if (argc - optind < 2) { fprintf(stderr, "Usage: yak-b200 recount [-o out.yak] <in.yak> <in.fa>\n"); return 1; }
	if ((h = yak_ch_restore(argv[optind])) == 0) { fprintf(stderr, "ERROR: failed to load '%s'\n", argv[optind]); return 1; }
	yak_ch_tighten(h);
	yak_recount(argv[optind + 1], h);
	yak_ch_dump(h, fn_out);
	yak_ch_destroy(h);
	return 0;
}

/* reference main.c:90-162: one table over several assemblies - count each input, keep its k-mers with min <= count <= max as
 * "seen once", add the samples up (yak_ch_merge), drop k-mers absent from more than -e samples every -s samples and at the end */
static int cmd_cntasm(int argc, char *argv[])
{
	yak_ch_t *h = 0, *h1;
	char *fn_in = 0, *fn_out = 0;
	int c, i, min_cnt = 1, max_cnt = 1, max_out = 0, check_n = 10, pre_resize = 0;
	yak_copt_t opt;
	yak_copt_init(&opt);
	opt.chunk_size = yakb_cli_parse_num("1.9g");
	while ((c = getopt(argc, argv, "k:p:K:t:i:o:c:x:e:s:r")) >= 0) {
		if (c == 'k') opt.k = atoi(optarg);
		else if (c == 'c') min_cnt = atoi(optarg);
		else if (c == 'x') max_cnt = atoi(optarg);
		else if (c == 'e') max_out = atoi(optarg);
		else if (c == 's') check_n = atoi(optarg);
		else if (c == 'r') pre_resize = 1;
		else if (c == 'p') opt.pre = atoi(optarg);
		else if (c == 'K') opt.chunk_size = yakb_cli_parse_num(optarg);
		else if (c == 't') opt.n_thread = atoi(optarg);
		else if (c == 'i') fn_in = optarg;
		else if (c == 'o') fn_out = optarg;
	}
	if (argc - optind < 1) {
		fprintf(stderr, "Usage: yak-b200 cntasm [options] <in1.fa> [in2.fa [...]]\n");
		fprintf(stderr, "Options:\n");
		fprintf(stderr, "  -k INT     k-mer size [%d]\n", opt.k);
		fprintf(stderr, "  -c INT     min count [%d]\n", min_cnt);
		fprintf(stderr, "  -x INT     max count [%d]\n", max_cnt);
		fprintf(stderr, "  -p INT     prefix length [%d]\n", opt.pre);
		fprintf(stderr, "  -r         resize before merging (same result bytes as the reference's -r)\n");
		fprintf(stderr, "  -t INT     number of worker threads (accepted, unused on the GPU) [%d]\n", opt.n_thread);
		fprintf(stderr, "  -e INT     exclude a k-mer if absent from INT samples [%d]\n", max_out);
		fprintf(stderr, "  -s INT     shrink the hash table every INT samples [%d]\n", check_n);
		fprintf(stderr, "  -K INT     chunk size [1.9g]\n");
		fprintf(stderr, "  -i FILE    input k-mer dump []\n");
		fprintf(stderr, "  -o FILE    output k-mer dump []\n");
		fprintf(stderr, "Note: if input and output file names are identical, input is overwritten\n");
		return 1;
	}
	if (opt.pre < YAK_COUNTER_BITS) { fprintf(stderr, "ERROR: -p should be at least %d\n", YAK_COUNTER_BITS); return 1; }
	if (opt.k >= 32) { fprintf(stderr, "ERROR: -k must be <=31\n"); return 1; }
	if (check_n <= 0) check_n = 1; /* the reference divides by -s (main.c:154) */
	if (fn_in) {
		h = yak_ch_restore(fn_in);
		if (h == 0) fprintf(stderr, "WARNING: failed to read %s. Continue anyway\n", fn_in);
	}
	for (i = optind; i < argc; ++i) {
		int n = i - optind + 1; /* samples so far */
		h1 = yak_count(argv[i], &opt, 0);
		if (h1 == 0) { /* the reference goes on with a null table here (quirk Q8) and crashes */
			fprintf(stderr, "ERROR: failed to count '%s' (no such file, or no CUDA device)\n", argv[i]);
			yak_ch_destroy(h);
			return 1;
		}
		if (h == 0) {
			h = h1;
			yak_ch_shrink(h, min_cnt, max_cnt, opt.n_thread);
			yak_ch_setcnt(h, 1, opt.n_thread);
		} else yak_ch_merge(h, h1, min_cnt, max_cnt, opt.n_thread, pre_resize); /* frees h1 */
		if (i == argc - 1 || (n > max_out && n % check_n == 0))
			yak_ch_shrink(h, n - max_out, YAK_MAX_COUNT, opt.n_thread);
		fprintf(stderr, "[M::%s::%.3f*%.2f] processed file %s; %ld distinct k-mers in the hash table\n", "main_cntasm",
		        realtime() - t_real0, cputime() / (realtime() - t_real0), argv[i], (long)h->tot);
	}
	yak_ch_tighten(h);
	if (fn_out) yak_ch_dump(h, fn_out);
	yak_ch_destroy(h);
	return 0;
}

/* reference main.c:217-284: k-mers of the first table absent from (subtract) / present in (isec) the others */
static int cmd_setop(int argc, char *argv[], int isec)
{
	yak_ch_t *h0, *h1;
	char *fn_out = "-";
	int c, i, n_thread = 8;
	while ((c = getopt(argc, argv, "o:t:")) >= 0) {
		if (c == 'o') fn_out = optarg;
		else if (c == 't') n_thread = atoi(optarg);
	}
	if (argc - optind < 2) { fprintf(stderr, "Usage: yak-b200 %s [-o out.yak] <in0.yak> <in1.yak> [...]\n", isec ? "isec" : "subtract"); return 1; }
	if ((h0 = yak_ch_restore(argv[optind])) == 0) { fprintf(stderr, "ERROR: failed to load '%s'\n", argv[optind]); return 1; }
	for (i = optind + 1; i < (isec ? argc : optind + 2); ++i) { /* subtract takes exactly one subtrahend (main.c:234-235) */
		if ((h1 = yak_ch_restore(argv[i])) == 0) { fprintf(stderr, "ERROR: failed to load '%s'\n", argv[i]); return 1; }
		if (isec) yak_ch_isec(h0, h1, n_thread); else yak_ch_subtract(h0, h1, n_thread);
		yak_ch_destroy(h1);
	}
	yak_ch_tighten(h0); /* main.c:243, 279 */
	yak_ch_dump(h0, fn_out);
	yak_ch_destroy(h0);
	return 0;
}

/* reference main.c:286-323: the k-mers of a table as text (k <= 31), optionally with counts */
static int cmd_print(int argc, char *argv[])
{
	yak_ch_t *h;
	int w, j, c, out_cnt = 0;
	uint32_t i, n;
	char buf[65];
	while ((c = getopt(argc, argv, "c")) >= 0)
		if (c == 'c') out_cnt = 1;
	if (argc - optind < 1) { fprintf(stderr, "Usage: yak-b200 print [-c] <in.yak>\n"); return 1; }
	if ((h = yak_ch_restore(argv[optind])) == 0) return 1;
	yak_ch_tighten(h);
	for (w = 0; w < 1 << h->pre; ++w) {
		yak_knt_t *a = yak_ch_getseq(h, w, &n);
		for (i = 0; i < n; ++i) {
			for (j = 0; j < h->k; ++j) buf[h->k - j - 1] = "ACGT"[a[i].x >> j * 2 & 3];
			buf[h->k] = 0;
			if (out_cnt) printf("%s\t%d\n", buf, a[i].c); else puts(buf);
		}
		free(a);
	}
	yak_ch_destroy(h);
	return 0;
}

int main(int argc, char *argv[])
{
	int ret, i;
	t_real0 = realtime();
	if (argc == 1) {
		fprintf(stderr, "Usage: yak-b200 <command> <argument>\n");
		fprintf(stderr, "Command:\n");
		fprintf(stderr, "  count     count k-mers\n");
		fprintf(stderr, "  recount   count existing k-mers\n");
		fprintf(stderr, "  cntasm    collate counts per dataset\n");
		fprintf(stderr, "  subtract  subtract k-mer sets\n");
		fprintf(stderr, "  isec      intersect k-mer sets\n");
		fprintf(stderr, "  print     print k-mers for k<=31\n");
		fprintf(stderr, "  qv        evaluate quality values\n");
		fprintf(stderr, "  triobin   trio binning\n");
		fprintf(stderr, "  trioeval  evaluate phasing accuracy with trio\n");
		fprintf(stderr, "  chkerr    find streaks of low-count k-mers\n");
		fprintf(stderr, "  sexchr    sex-chromosome k-mer content of two haplotype assemblies\n");
		fprintf(stderr, "  inspect   k-mer hash tables\n");
		fprintf(stderr, "  version   print version number\n");
		return 1;
	}
	if (strcmp(argv[1], "count") == 0) ret = cmd_count(argc - 1, argv + 1);
	else if (strcmp(argv[1], "recount") == 0) ret = cmd_recount(argc - 1, argv + 1);
	else if (strcmp(argv[1], "cntasm") == 0) ret = cmd_cntasm(argc - 1, argv + 1);
	else if (strcmp(argv[1], "subtract") == 0) ret = cmd_setop(argc - 1, argv + 1, 0);
	else if (strcmp(argv[1], "isec") == 0) ret = cmd_setop(argc - 1, argv + 1, 1);
	else if (strcmp(argv[1], "print") == 0) ret = cmd_print(argc - 1, argv + 1);
	else if (strcmp(argv[1], "qv") == 0) ret = cmd_qv(argc - 1, argv + 1);
	else if (strcmp(argv[1], "triobin") == 0) ret = yakb_cmd_triobin(argc - 1, argv + 1);
	else if (strcmp(argv[1], "trioeval") == 0) ret = yakb_cmd_trioeval(argc - 1, argv + 1);
	else if (strcmp(argv[1], "chkerr") == 0) ret = yakb_cmd_chkerr(argc - 1, argv + 1);
	else if (strcmp(argv[1], "sexchr") == 0) ret = yakb_cmd_sexchr(argc - 1, argv + 1);
	else if (strcmp(argv[1], "inspect") == 0) ret = cmd_inspect(argc - 1, argv + 1);
	else if (strcmp(argv[1], "version") == 0) { puts(YAKS_VERSION); return 0; }
	else { fprintf(stderr, "[E::%s] unknown command\n", __func__); return 1; }
	if (ret == 0) {
		fprintf(stderr, "[M::%s] Version: %s\n", __func__, YAKS_VERSION);
		fprintf(stderr, "[M::%s] CMD:", __func__);
		for (i = 0; i < argc; ++i) fprintf(stderr, " %s", argv[i]);
		fprintf(stderr, "\n[M::%s] Real time: %.3f sec; CPU: %.3f sec; Peak RSS: %.3f GB\n", __func__,
		        realtime() - t_real0, cputime(), peakrss() / 1024.0 / 1024.0 / 1024.0);
	}
	return ret;
}
