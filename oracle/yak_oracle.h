/* yak_oracle.h - CPU restatement of lh3/yak's k-mer count / lookup path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under yak_b200/ (the product) may include, link or call
 * this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker.  It is a plain sequential C restatement of the reference
 * algorithm (no threads, no GPU), each function citing the reference file:line it follows.
 *
 * Parity pin: the outputs of oracle/_ref/ (the unmodified reference compiled from
 * /root/reference by oracle/Makefile) on seeded inputs are committed as digests under
 * tests/golden/ and reproduced by this code in tests/test_oracle_cpu.py, together with the
 * known-answer vectors of SURVEY.md section 4 and the reference's own table set-ops called
 * through oracle/_ref/libyakref.so (tests/test_setops.py).
 */
#ifndef YAK_ORACLE_H
#define YAK_ORACLE_H
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YO_COUNTER_BITS 10
#define YO_MAX_COUNT 1023
#define YO_N_COUNTS 1024

/* ---- hashes: yak-priv.h:11-68 ---- */
uint64_t yo_hash64(uint64_t key, uint64_t mask);
uint64_t yo_hash64_64(uint64_t key);
uint64_t yo_hash_long(const uint64_t x[4]);
uint64_t yo_hash64_inv(uint64_t key, uint64_t mask);
/* khashl.h:98 + the uint32 truncation at khashl.h:262 (yak_ch_hash -> khint_t) */
uint32_t yo_slot_home(uint64_t stored_key, uint32_t bits);
/* misc.c:4-21 */
extern const unsigned char yo_nt4[256];

/* ---- open-addressing set with khashl's exact probe / resize behaviour: khashl.h:137-221 ---- */
typedef struct {
	uint32_t bits, count;
	uint32_t *used;   /* 1 bit per slot */
	uint64_t *keys;   /* NULL until the first resize */
} yo_set_t;
uint32_t yo_set_capacity(const yo_set_t *s);
int      yo_set_resize(yo_set_t *s, uint32_t request);
uint32_t yo_set_put(yo_set_t *s, uint64_t key, int *absent);
uint32_t yo_set_get(const yo_set_t *s, uint64_t key); /* == capacity if absent */
int      yo_set_used(const yo_set_t *s, uint32_t i);

/* ---- blocked bloom filter: bbf.c:5-42 ---- */
typedef struct {
	int n_shift, n_hashes;
	uint8_t *b;
} yo_bloom_t;
yo_bloom_t *yo_bloom_init(int n_shift, int n_hashes);
void yo_bloom_destroy(yo_bloom_t *b);
int  yo_bloom_insert(yo_bloom_t *b, uint64_t hash);

/* ---- count table: htab.c ---- */
typedef struct {
	int k, pre, n_hash, n_shift;
	uint64_t tot;
	yo_set_t *h;       /* 1<<pre sub-tables */
	yo_bloom_t **b;    /* 1<<pre filters or all NULL */
} yo_ch_t;
yo_ch_t *yo_ch_init(int k, int pre, int n_hash, int n_shift);
void yo_ch_destroy(yo_ch_t *h);
void yo_ch_destroy_bf(yo_ch_t *h);
int  yo_ch_insert_list(yo_ch_t *h, int create_new, int n, const uint64_t *a);
void yo_ch_insert_events(yo_ch_t *h, int create_new, int64_t n, const uint64_t *a);
int  yo_ch_get(const yo_ch_t *h, uint64_t x);
int  yo_ch_inc(yo_ch_t *h, uint64_t x);
void yo_ch_clear(yo_ch_t *h);
void yo_ch_hist(const yo_ch_t *h, int64_t cnt[YO_N_COUNTS]);
void yo_ch_shrink(yo_ch_t *h, int min, int max);
void yo_ch_tighten(yo_ch_t *h);
void yo_ch_setcnt(yo_ch_t *h, int cnt);
void yo_ch_merge(yo_ch_t *h0, yo_ch_t *h1, int min, int max, int pre_resize); /* consumes h1 */
void yo_ch_subtract(yo_ch_t *h0, const yo_ch_t *h1);
void yo_ch_isec(yo_ch_t *h0, const yo_ch_t *h1);
int  yo_ch_dump(const yo_ch_t *h, const char *fn);
yo_ch_t *yo_ch_restore(const char *fn);
/* htab.c:396-476: load into ch0 (or a new table) with the count -> flag remapping of `mode` */
yo_ch_t *yo_ch_restore_core(yo_ch_t *ch0, const char *fn, int mode, int min_cnt, int mid_cnt, int64_t n_io[2]);
/* per-position lookups of one sequence, the loop shared by triobin/trioeval/chkerr/sexchr */
void yo_scan_seq(const yo_ch_t *ch, int64_t len, const char *seq, int16_t *out);
/* serialise into a malloc'd buffer (same bytes as yo_ch_dump); returns length */
int64_t yo_ch_dump_mem(const yo_ch_t *h, uint8_t **out);

/* ---- event stream: count.c:28-60 ---- */
/* append the hashed canonical k-mers of one sequence to out[] (capacity >= len); returns #events */
int64_t yo_extract(int k, int64_t len, const char *seq, uint64_t *out);

/* ---- whole-file driver: count.c:85-166 + main.c:53-60, sequential ---- */
yo_ch_t *yo_count_file(const char *fn, int k, int pre, int bf_shift, int bf_n_hash, yo_ch_t *h0,
                       int64_t *n_events);            /* chunk_size = 10 M, yak_copt_init's default */
/* the -K chunk size decides what happens after a truncated FASTQ record (count.c:93,109) */
yo_ch_t *yo_count_file_chunked(const char *fn, int k, int pre, int bf_shift, int bf_n_hash, yo_ch_t *h0,
                               int64_t *n_events, int64_t chunk_size);
/* same over sequences already in memory (concatenated, lens[] gives each length) */
yo_ch_t *yo_count_seqs(int64_t n_seq, const int64_t *lens, const char *cat, int k, int pre,
                       int bf_shift, int bf_n_hash, yo_ch_t *h0, int64_t *n_events);

/* ---- qv scan: qv.c:34-86, 116-135 ---- */
/* per-sequence tot / non0 written to out_tot/out_non0 (may be NULL); cnt[1024] accumulated */
void yo_qv_seqs(const yo_ch_t *ch, int64_t n_seq, const int64_t *lens, const char *cat,
                int min_len, double min_frac, int64_t cnt[YO_N_COUNTS],
                int32_t *out_tot, int32_t *out_non0);

/* ---- FASTA/FASTQ reader with kseq.h:192-232 semantics (plain or gz) ---- */
typedef struct yo_reader_s yo_reader_t;
yo_reader_t *yo_reader_open(const char *fn);
/* returns length >= 0, -1 at EOF, -2 on truncated quality; *seq valid until next call */
int64_t yo_reader_next(yo_reader_t *r, const char **seq, const char **name);
void yo_reader_close(yo_reader_t *r);

#ifdef __cplusplus
}
#endif
#endif
