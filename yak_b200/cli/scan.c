/* scan.c - the lookup scanners of the reference CLI over the batched device lookup:
 * triobin (triobin.c:148-197), trioeval (trioeval.c:153-212), chkerr (chkerr.c:99-133), sexchr
 * (sexchr.c:97-140).  Same options, defaults, output lines and chunking (bseq.c:33-57) as the reference;
 * per batch the k-mer lookups are ONE library call (yakb_scan_seqs), the per-sequence logic is in
 * scan_logic.c.  Lines come out in input order (the reference's order with -t1). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "yak_b200.h"
#include "scan_logic.h"

int64_t yakb_cli_parse_num(const char *s); /* main.c */

typedef struct {
	void *rd;
	int ref_workers;          /* pipeline workers the reference has left (2 at the start: kt_pipeline(2, ...)) */
	int64_t n_seq, m_seq;
	char **names;
	int64_t *lens;
	char *cat;
	int64_t l_cat, m_cat;
	int16_t *vals;
	int64_t m_vals;
} batch_reader_t;

static int br_open(batch_reader_t *br, const char *fn)
{
	memset(br, 0, sizeof(*br));
	br->rd = yakb_fastx_open(fn);
	br->ref_workers = 2;
	return br->rd != 0;
}

static void br_reset(batch_reader_t *br)
{
	int64_t i;
	for (i = 0; i < br->n_seq; ++i) free(br->names[i]);
	br->n_seq = 0; br->l_cat = 0;
}

/* bseq.c:33-57: records until the batch holds >= chunk_size bases; 0 at the end of the input */
static int br_next(batch_reader_t *br, int64_t chunk_size, const char *who)
{
	const char *seq, *name;
	int64_t len;
	br_reset(br);
	if (br->ref_workers == 0) return 0;
	for (;;) {
		len = yakb_fastx_next(br->rd, &seq, &name);
		if (len == -2) { /* a truncated FASTQ record ends bseq_read's batch (bseq.c:40); a call that read nothing retires one
		                    of the reference's two pipeline workers (kthread.c:119), the second ends the input */
			if (br->n_seq > 0) break;
			if (--br->ref_workers == 0) break;
			continue;
		}
		if (len < 0) { br->ref_workers = 0; break; }
		if (br->n_seq == br->m_seq) {
			br->m_seq = br->m_seq ? br->m_seq << 1 : 256;
			br->names = (char**)realloc(br->names, br->m_seq * sizeof(char*));
			br->lens = (int64_t*)realloc(br->lens, br->m_seq * sizeof(int64_t));
		}
		if (br->l_cat + len > br->m_cat) {
			br->m_cat = (br->l_cat + len) + ((br->l_cat + len) >> 1) + 4096;
			br->cat = (char*)realloc(br->cat, br->m_cat);
		}
		memcpy(br->cat + br->l_cat, seq, len);
		br->names[br->n_seq] = strdup(name);
		br->lens[br->n_seq++] = len;
		br->l_cat += len;
		if (br->l_cat >= chunk_size) break;
	}
	if (br->n_seq == 0) return 0;
	fprintf(stderr, "[M::%s] read %d sequences\n", who, (int)br->n_seq);
	if (br->l_cat > br->m_vals) {
		br->m_vals = br->l_cat + (br->l_cat >> 1) + 4096;
		br->vals = (int16_t*)realloc(br->vals, br->m_vals * sizeof(int16_t));
	}
	return 1;
}

static void br_close(batch_reader_t *br)
{
	br_reset(br);
	free(br->names); free(br->lens); free(br->cat); free(br->vals);
	if (br->rd) yakb_fastx_close(br->rd);
	memset(br, 0, sizeof(*br));
}

static int br_lookup(batch_reader_t *br, const yak_ch_t *ch, yakb_scan_batch_t *b)
{
	if (yakb_scan_seqs(ch, br->n_seq, br->lens, br->cat, br->vals) != 0) return -1;
	b->n_seq = br->n_seq; b->names = br->names; b->lens = br->lens; b->vals = br->vals;
	return 0;
}

static yak_ch_t *load_trio(const char *pat, const char *mat, int min_cnt, int mid_cnt)
{
	yak_ch_t *ch = yak_ch_restore_core(0, pat, YAK_LOAD_TRIOBIN1, min_cnt, mid_cnt);
	if (ch == 0) { fprintf(stderr, "ERROR: fail to load '%s'\n", pat); return 0; }
	if (yak_ch_restore_core(ch, mat, YAK_LOAD_TRIOBIN2, min_cnt, mid_cnt) == 0) {
		fprintf(stderr, "ERROR: fail to load '%s'\n", mat);
		yak_ch_destroy(ch);
		return 0;
	}
	return ch;
}

int yakb_cmd_triobin(int argc, char *argv[])
{
	int c, min_cnt = 2, mid_cnt = 5, n_threads = 8;
	yakb_triobin_opt_t opt;
	yakb_scan_batch_t b;
	batch_reader_t br;
	yak_ch_t *ch;
	opt.print_diff = 0; opt.ratio_thres = 0.33;
	while ((c = getopt(argc, argv, "c:d:t:pr:")) >= 0) {
		if (c == 'c') min_cnt = atoi(optarg);
		else if (c == 'd') mid_cnt = atoi(optarg);
		else if (c == 't') n_threads = atoi(optarg);
		else if (c == 'p') opt.print_diff = 1;
		else if (c == 'r') opt.ratio_thres = atof(optarg);
	}
	if (argc - optind < 2) {
		fprintf(stderr, "Usage: yak-b200 triobin [options] <pat.yak> <mat.yak> <seq.fa>\n");
		fprintf(stderr, "Options:\n");
		fprintf(stderr, "  -c INT     min occurrence [%d]\n", min_cnt);
		fprintf(stderr, "  -d INT     mid occurrence [%d]\n", mid_cnt);
		fprintf(stderr, "  -t INT     number of threads (accepted, unused on the GPU) [%d]\n", n_threads);
		return 1;
	}
	if ((ch = load_trio(argv[optind], argv[optind + 1], min_cnt, mid_cnt)) == 0) return 1;
	opt.k = ch->k;
	if (!br_open(&br, argv[optind + 2])) {
		fprintf(stderr, "ERROR: fail to open file '%s'\n", argv[optind + 2]);
		exit(1);
	}
	while (br_next(&br, 200000000, "tb_pipeline")) { /* triobin.c:13 CHUNK_SIZE */
		if (br_lookup(&br, ch, &b) != 0) return 1;
		yakb_triobin_batch(stdout, &opt, &b);
	}
	br_close(&br);
	yak_ch_destroy(ch);
	return 0;
}

int yakb_cmd_trioeval(int argc, char *argv[])
{
	int c, min_cnt = 2, mid_cnt = 5, n_threads = 8;
	int64_t cnt[YAK_N_COUNTS];
	yakb_trioeval_opt_t opt;
	yakb_trioeval_sum_t sum;
	yakb_scan_batch_t b;
	batch_reader_t br;
	yak_ch_t *ch;
	opt.min_n = 2; opt.print_err = 0; opt.print_frag = 1;
	memset(&sum, 0, sizeof(sum));
	while ((c = getopt(argc, argv, "c:d:t:n:eF")) >= 0) {
		if (c == 'c') min_cnt = atoi(optarg);
		else if (c == 'd') mid_cnt = atoi(optarg);
		else if (c == 't') n_threads = atoi(optarg);
		else if (c == 'n') opt.min_n = atoi(optarg);
		else if (c == 'e') opt.print_err = 1;
		else if (c == 'F') opt.print_frag = 0;
	}
	if (argc - optind < 2) {
		fprintf(stderr, "Usage: yak-b200 trioeval [options] <pat.yak> <mat.yak> <seq.fa>\n");
		fprintf(stderr, "Options:\n");
		fprintf(stderr, "  -c INT     min occurrence [%d]\n", min_cnt);
		fprintf(stderr, "  -d INT     mid occurrence [%d]\n", mid_cnt);
		fprintf(stderr, "  -n INT     min streak [%d]\n", opt.min_n);
		fprintf(stderr, "  -t INT     number of threads (accepted, unused on the GPU) [%d]\n", n_threads);
		fprintf(stderr, "  -e         print error positions\n");
		return 1;
	}
	if ((ch = load_trio(argv[optind], argv[optind + 1], min_cnt, mid_cnt)) == 0) return 1;
	yak_ch_hist(ch, cnt, n_threads);
	fprintf(stderr, "[M::%s] %ld file1-specific k-mers and %ld file2-specific k-mers\n", "main_trioeval",
	        (long)cnt[0 << 2 | 2], (long)cnt[2 << 2 | 0]);
	opt.k = ch->k;
	if (!br_open(&br, argv[optind + 2])) {
		fprintf(stderr, "ERROR: fail to open file '%s'\n", argv[optind + 2]);
		exit(1);
	}
	yakb_trioeval_header(stdout);
	while (br_next(&br, 1000000000, "te_pipeline")) { /* trioeval.c:13 CHUNK_SIZE */
		if (br_lookup(&br, ch, &b) != 0) return 1;
		yakb_trioeval_batch(stdout, &opt, &b, &sum);
	}
	br_close(&br);
	yak_ch_destroy(ch);
	yakb_trioeval_footer(stdout, &sum);
	return 0;
}

int yakb_cmd_chkerr(int argc, char *argv[])
{
	int c, n_threads = 8;
	yakb_chkerr_opt_t opt;
	yakb_scan_batch_t b;
	batch_reader_t br;
	yak_ch_t *ch;
	opt.min_cnt = 3; opt.min_streak = 5;
	while ((c = getopt(argc, argv, "t:c:s:")) >= 0) {
		if (c == 't') n_threads = atoi(optarg);
		else if (c == 'c') opt.min_cnt = atoi(optarg);
		else if (c == 's') opt.min_streak = atoi(optarg);
	}
	if (argc - optind < 2) {
		fprintf(stderr, "Usage: yak-b200 chkerr [options] <count.yak> <seq.fa>\n");
		fprintf(stderr, "Options:\n");
		fprintf(stderr, "  -t INT    number of threads (accepted, unused on the GPU) [%d]\n", n_threads);
		fprintf(stderr, "  -c INT    min k-mer count [%d]\n", opt.min_cnt);
		fprintf(stderr, "  -s INT    min k-mer streak [%d]\n", opt.min_streak);
		return 1;
	}
	if ((ch = yak_ch_restore(argv[optind])) == 0) { fprintf(stderr, "ERROR: fail to load '%s'\n", argv[optind]); return 1; }
	opt.k = ch->k;
	if (!br_open(&br, argv[optind + 1])) {
		fprintf(stderr, "ERROR: fail to open file '%s'\n", argv[optind + 1]);
		exit(1);
	}
	while (br_next(&br, 1000000000, "ce_pipeline")) { /* chkerr.c:107 chunk_size */
		if (br_lookup(&br, ch, &b) != 0) return 1;
		yakb_chkerr_batch(stdout, &opt, &b);
	}
	br_close(&br);
	yak_ch_destroy(ch);
	return 0;
}

int yakb_cmd_sexchr(int argc, char *argv[])
{
	int c, i, n_threads = 8;
	int64_t chunk_size = 1000000000; /* sexchr.c:13 */
	yakb_scan_batch_t b;
	batch_reader_t br;
	yak_ch_t *ch;
	while ((c = getopt(argc, argv, "t:K:")) >= 0) {
		if (c == 't') n_threads = atoi(optarg);
		else if (c == 'K') chunk_size = yakb_cli_parse_num(optarg);
	}
	if (argc - optind < 5) {
		fprintf(stderr, "Usage: yak-b200 sexchr [options] <chrY.yak> <chrX.yak> <PAR.yak> <hap1.fa> <hap2.fa>\n");
		fprintf(stderr, "Options:\n");
		fprintf(stderr, "  -t INT     number of threads (accepted, unused on the GPU) [%d]\n", n_threads);
		fprintf(stderr, "  -K NUM     chunk size [1g]\n");
		return 1;
	}
	ch = yak_ch_restore_core(0, argv[optind], YAK_LOAD_SEXCHR1);
	if (ch == 0 || yak_ch_restore_core(ch, argv[optind + 1], YAK_LOAD_SEXCHR2) == 0 ||
	    yak_ch_restore_core(ch, argv[optind + 2], YAK_LOAD_SEXCHR3) == 0) {
		fprintf(stderr, "ERROR: fail to load the k-mer tables\n");
		return 1;
	}
	yakb_sexchr_header(stdout);
	for (i = 1; i <= 2; ++i) {
		if (!br_open(&br, argv[optind + i + 2])) {
			fprintf(stderr, "ERROR: fail to open file '%s'\n", argv[optind + i + 2]);
			exit(1);
		}
		while (br_next(&br, chunk_size, "sc_pipeline")) {
			if (br_lookup(&br, ch, &b) != 0) return 1;
			yakb_sexchr_batch(stdout, i, &b);
		}
		br_close(&br);
	}
	yak_ch_destroy(ch);
	return 0;
}
