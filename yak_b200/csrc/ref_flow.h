/* ref_flow.h - how far the reference reads a file that holds FASTQ records with a truncated quality string.
 *
 * The reference's readers are step 0 of a kt_pipeline: `yak count` (count.c:88-110, 3 workers: count.c:162) collects records
 * of at least k bases until they hold chunk_size bases; bseq_read (bseq.c:33-57; qv.c:126, triobin.c:191, trioeval.c:203,
 * chkerr.c:129, sexchr.c:135: 2 workers) collects every record until chunk_size bases.  kseq_read's -2 (truncated quality)
 * ends the call that met it; the next call resumes at the next header character (kseq.h:192-199).  A call that collected
 * nothing returns NULL, and that retires the worker that made it (kthread.c:119): the input ends with the last worker.
 * Plain C, shared by the readers (fastx.cpp), yak_qv (capi.cu) and the tests (yakb_ref_flow_sim). */
#ifndef YAKB_REF_FLOW_H
#define YAKB_REF_FLOW_H
#include <stdint.h>

typedef struct {
	int workers, min_len;     /* pipeline workers left; records shorter than min_len do not count (count.c:95) */
	int64_t n, size, chunk;   /* records and bases the current call holds; a call is full at `chunk` bases */
} yakb_ref_flow_t;

static inline void yakb_ref_flow_init(yakb_ref_flow_t *f, int workers, int64_t chunk, int min_len)
{
	f->workers = workers; f->min_len = min_len; f->n = f->size = 0; f->chunk = chunk > 0 ? chunk : 1;
}
/* a record of `len` bases was read */
static inline void yakb_ref_flow_record(yakb_ref_flow_t *f, int64_t len)
{
	if (len < f->min_len) return;
	++f->n; f->size += len;
	if (f->size >= f->chunk) f->n = f->size = 0; /* count.c:106 / bseq.c:53: the call is full, the next one starts */
}
/* kseq_read returned -2: 1 = reading goes on behind the record, 0 = the input ends here */
static inline int yakb_ref_flow_bad(yakb_ref_flow_t *f)
{
	if (f->workers <= 0) return 0; /* initialised with 0 workers: a plain `while (kseq_read(ks) >= 0)` loop (yak_recount, count.c:176) */
	if (f->n == 0 && --f->workers <= 0) return 0;
	f->n = f->size = 0;
	return 1;
}
#endif
