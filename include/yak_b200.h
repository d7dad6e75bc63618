/* yak_b200.h - GPU-side extensions of the drop-in library (libyakb200.so), C ABI.
 *
 * yak.h is the reference's API; the entry points here expose the same hot path at chunk
 * granularity on raw host/device pointers, for callers that already hold data on the device
 * (bench.py, the multi-GPU harness, tests).  No torch types, plain pointers and sizes.
 * All functions return 0 on success and a negative value on error (message on stderr), unless
 * stated otherwise.
 */
#ifndef YAK_B200_H
#define YAK_B200_H

#include <stdint.h>
#include "yak.h"

#ifdef __cplusplus
extern "C" {
#endif

const char *yakb_version(void);
int yakb_device_count(void);                       /* 0 when no CUDA device is usable */

/* stats[4] = {k-mer events, pending (not in table before the chunk), put-events, new keys} */

/* One chunk of the count path (reference count.c:111-143 + htab.c:51-78) on a device-resident
 * ASCII base stream: any byte that is not A/C/G/T/U (either case) or 0..3 separates reads. */
int yakb_count_ascii_dev(yak_ch_t *h, const void *d_asc, uint64_t n, int create_new, uint64_t stats[4]);
/* same, bases in host memory (copied to the device inside the call) */
int yakb_count_ascii_host(yak_ch_t *h, const char *asc, uint64_t n, int create_new, uint64_t stats[4]);
/* One chunk given as already hashed k-mers in file order (what count.c:17-26 buffers hold, all
 * sub-tables mixed); the receive side of the multi-GPU exchange. */
int yakb_count_events_dev(yak_ch_t *h, const uint64_t *d_ev, uint64_t n, int create_new, uint64_t stats[4]);

/* Extraction only (reference count.c:28-60): hashed canonical k-mers of a device ASCII stream in
 * file order, stably partitioned by owner rank = (hash & (2^pre-1)) * world >> pre.
 * d_out must hold n entries; counts[world] (host) receives the per-rank counts. */
int yakb_extract_route_dev(const void *d_asc, uint64_t n, int k, int pre, int world,
                           uint64_t *d_out, uint64_t *counts, void *cuda_stream);
/* the same without waiting for the device: d_counts[world] is DEVICE memory (the caller exchanges the counts
 * of all ranks with one collective and reads them back once); world is a power of two <= 16 */
int yakb_extract_route_async(const void *d_asc, uint64_t n, int k, int pre, int world,
                             uint64_t *d_out, uint64_t *d_counts, void *cuda_stream);

/* Device-side text ingest (csrc/ingest.cu): FASTA/FASTQ text in the strict 2- / 4-line layout -> the dense "SEQ\nSEQ\n..."
 * base stream the calls above take, with every record's layout checked on the device.
 * yakb_record_start_before: the last offset in (lo, pos] of `text` (host memory, `size` bytes) where a record starts, or lo;
 * a guess that the device check confirms or refutes.
 * yakb_ingest_dev: d_raw[n] (n < 2^32) must start at a record and end behind a record's newline; d_out (capacity n)
 * receives the bases; d_res (DEVICE, 3 x u64): [0] = 0 if the layout holds for every record (else the caller must use the
 * host parser, kseq.h:192-232 semantics), [1] = bytes written to d_out, [2] = lines.  Does not wait for the device. */
uint64_t yakb_record_start_before(const void *text, uint64_t lo, uint64_t pos, uint64_t size, int lines_per_record);
int yakb_ingest_dev(const void *d_raw, uint64_t n, int lines_per_record, void *d_out, uint64_t *d_res, void *cuda_stream);

/* Multi-GPU: one shard of a table whose 2^pre sub-tables are split over `world` (power of two)
 * GPUs; shard `rank` owns sub-tables [rank*2^pre/world, (rank+1)*2^pre/world) and ignores events
 * of other sub-tables.  Feed it with yakb_count_events_dev after the all-to-all.  The shards'
 * yakb_ch_dump_shard_mem outputs concatenated in rank order (header from rank 0 only) are the
 * bytes yak_ch_dump would write for the whole table. */
yak_ch_t *yakb_ch_init_shard(int k, int pre, int n_hash, int n_shift, int rank, int world);
int64_t yakb_ch_dump_shard_mem(const yak_ch_t *h, int with_header, uint8_t **out);
/* the same image written at `offset` of the existing file `fn` (the ranks of a job write one .yak side by side);
 * _size tells beforehand how many bytes it takes.  Both return -1 on failure. */
int64_t yakb_ch_dump_shard_size(const yak_ch_t *h, int with_header);
int64_t yakb_ch_dump_shard_at(const yak_ch_t *h, int with_header, const char *fn, uint64_t offset);

/* batched yak_ch_get (reference htab.c:93-100): out[i] = count or -1 */
int yakb_ch_get_batch(const yak_ch_t *h, uint64_t n, const uint64_t *x, int32_t *out);
int yakb_ch_get_batch_dev(const yak_ch_t *h, uint64_t n, const uint64_t *d_x, int32_t *d_out);

/* qv scan (reference qv.c:34-86) over sequences already in host memory: `cat` holds the
 * sequences back to back, lens[i] their lengths.  cnt[1024] is accumulated like yak_qv's;
 * tot/non0 (may be NULL) receive the per-sequence numbers of the SQ line. */
int yakb_qv_seqs(const yak_ch_t *h, int64_t n_seq, const int64_t *lens, const char *cat,
                 int min_len, double min_frac, int64_t cnt[YAK_N_COUNTS], int32_t *tot, int32_t *non0);

/* The lookup loop shared by the reference's other scanners (triobin.c:62-86, trioeval.c:61-89,
 * chkerr.c:35-56, sexchr.c:42-66; any k < 64) as one batched device call over sequences in host memory
 * (`cat` back to back, lens[i] each): out[j] for the k-mer ENDING at base j of `cat` = yak_ch_get's
 * value (-1 absent), or -2 where no k-mer ends.  yak_b200/cli/scan_logic.c does the per-sequence part. */
int yakb_scan_seqs(const yak_ch_t *h, int64_t n_seq, const int64_t *lens, const char *cat, int16_t *out);

/* serialise exactly the bytes yak_ch_dump writes into a malloc'd buffer; returns the length */
int64_t yakb_ch_dump_mem(const yak_ch_t *h, uint8_t **out);
/* make room for this many distinct keys per sub-table up front (optional) */
int yakb_ch_reserve(yak_ch_t *h, uint64_t keys_per_subtable);
/* the CUDA stream (cudaStream_t) every kernel of this table is launched on, for event timing */
void *yakb_ch_stream(const yak_ch_t *h);
/* bytes of device memory held by the table, bloom and journal */
uint64_t yakb_ch_device_bytes(const yak_ch_t *h);
/* number of kernels this library launched so far in this process */
uint64_t yakb_kernel_launches(void);
/* GPUs a table lives on: 1, or the G of a table yak_count() spread over the GPUs of this process (environment YAKB_GPUS = a
 * power of two: shard r on device r owns sub-tables [r*2^pre/G, (r+1)*2^pre/G); per batch one host thread per GPU copies its
 * part of the reads, extracts and groups the k-mers by owner, and every GPU pulls its runs from its peers over NVLink).
 * Such a table supports the `yak count` flow: yak_count, yak_ch_destroy_bf / clear / shrink / hist / get / insert_list /
 * dump / destroy; the other entry points report an error. */
int yakb_ch_gpus(const yak_ch_t *h);
/* The library keeps device blocks of destroyed tables for the next table (csrc/dbuf.cuh; bounded by
 * YAKB_CACHE_GB, default 48): bytes held idle right now, and a call that hands them all back. */
uint64_t yakb_device_cache_bytes(void);
void yakb_device_cache_trim(void);

/* test hook, no GPU: the host side of yak_ch_restore_core (reference htab.c:419-472) - header checks, then the sub-table
 * capacities, offsets and keys (counts mapped to flag bits for the TRIOBIN / SEXCHR modes) as the device receives them; the
 * 2^pre key arrays are read by `threads` threads (0 = one per core, at most 16).  Arrays are malloc'd for the caller.
 * Returns 0, -1 unreadable, -2 wrong magic, -3 wrong counter bits. */
int yakb_yak_file_read(const char *fn, int mode, int min_cnt, int mid_cnt, int threads, uint32_t *k, uint32_t *pre,
                       uint32_t **caps, uint64_t **off, uint64_t **keys);

/* the host-side FASTA/FASTQ record reader yak_count/yak_qv use (reference kseq.h:192-232
 * semantics; plain or gzip; NULL or "-" = stdin).  next() returns the sequence length, -1 at EOF,
 * -2 on a truncated quality string; *seq / *name stay valid until the following call. */
void *yakb_fastx_open(const char *fn);
int64_t yakb_fastx_next(void *reader, const char **seq, const char **name);
void yakb_fastx_close(void *reader);
/* A regular file made of BGZF blocks (bgzip, samtools, most sequencer pipelines) is inflated by a pool of threads
 * (csrc/bgzf.h; same byte stream as zlib's gzread, reference count.c:150-151).  yakb_fastx_open uses one thread per
 * core; here bgzf_threads < 0 keeps zlib's sequential reader, job_bytes = inflated bytes per unit of work. */
void *yakb_fastx_open_bgzf(const char *fn, int bgzf_threads, uint64_t job_bytes);
int yakb_fastx_bgzf_threads(void *reader);   /* 0 when the input does not go through the pool */
/* `yak count -b` reads its input twice (reference main.c:53-60).  With YAKB_TEXT_CACHE_GB=<GB> (off by default) the first
 * pass over a compressed file keeps the inflated text in an anonymous memory file and the second pass parses that with the
 * parser pool instead of inflating again (csrc/textcache.h).  Test hooks for the two halves, no GPU involved:
 * open a reader that keeps the text, commit after the last fill (1 = kept), ask for the path a second pass would open. */
void *yakb_fastx_open_tee(const char *fn);
int yakb_fastx_tee_commit(void *reader);
int yakb_text_cache_path(const char *fn, char *buf, int len);
void yakb_text_cache_release(void);
/* bulk form used by yak_count: append whole records (length >= min_len) as "SEQ\n" until `target`
 * bytes; returns bytes appended; *done = input exhausted; *need != 0: grow buf to that size */
int64_t yakb_fastx_fill(void *reader, char *buf, uint64_t cap, uint64_t target, int min_len,
                        int64_t *n_seq, int *done, uint64_t *need);
/* multi-threaded variant for plain regular files (speculative block parse + exact stitching,
 * csrc/fastx_par.h); open returns NULL for gzip / stdin / unreadable files */
void *yakb_pfastx_open(const char *fn, uint64_t block_bytes, int threads);
int64_t yakb_pfastx_fill(void *reader, char *buf, uint64_t cap, uint64_t target, int min_len,
                         int64_t *n_seq, int *done, uint64_t *need);
uint64_t yakb_pfastx_redo(void *reader);   /* blocks that had to be re-parsed sequentially */
/* The reference's -K (count.c:106) for the bulk forms: whether `yak count` reads on behind a FASTQ record with a
 * truncated quality string depends on it (count.c:93,109,162; kthread.c:119).  Default 10 M (misc.c:31). */
void yakb_fastx_set_chunk(void *reader, int64_t chunk_size);
/* the rule itself (csrc/ref_flow.h), for tests: lens[i] = length of record i, or -2 for a truncated FASTQ record;
 * read_out[i] = 1 if the reference reads record i with `workers` pipeline threads (3 for count, 2 for the scanners) */
void yakb_ref_flow_sim(const int64_t *lens, int64_t n, int workers, int64_t chunk_size, int min_len, uint8_t *read_out);
void yakb_pfastx_set_chunk(void *reader, int64_t chunk_size);
/* the reference's pipeline threads for the caller being served (csrc/ref_flow.h): 3 = yak count (the default), 2 = qv and the
 * scanners (with flow_min_len 0: bseq_read keeps every record), 0 = yak_recount's plain loop, which ends at the first such record */
void yakb_fastx_set_workers(void *reader, int workers);
void yakb_pfastx_set_flow(void *reader, int64_t chunk_size, int workers, int flow_min_len);
void yakb_pfastx_close(void *reader);
/* skip n_skip records, then append up to n_take records (those of length >= min_len) to buf as
 * "SEQ\n"; returns the records consumed (-1: buf too small).  Lets each rank of a multi-GPU job
 * take its contiguous slice of every chunk of one shared input file. */
int64_t yakb_fastx_read_slice(void *reader, int64_t n_skip, int64_t n_take, int min_len,
                              char *buf, uint64_t cap, uint64_t *n_bytes, int64_t *n_seq);

/* per-kernel device time (CUDA events on the table's stream): enable, run chunks, read
 * {"kernel": [total_ms, launches], ...} as JSON text; returns its length or -1 if buf is too small */
void yakb_prof_enable(int on);
int yakb_prof_json(char *buf, uint64_t cap);

/* seeded synthetic data on the device (same stream as yak_b200/synth.py): genome as 2-bit codes
 * packed 32 per u64; reads as text records of fixed size: fmt 0 "SEQ\n" (L+1 bytes),
 * fmt 1 FASTA ">r\nSEQ\n" (L+4), fmt 2 FASTQ "@r\nSEQ\n+\nIII..\n" (2L+7) */
int yakb_synth_genome_dev(uint64_t seed_g, uint64_t G, uint64_t *d_genome2, void *cuda_stream);
int yakb_synth_reads_dev(const uint64_t *d_genome2, uint64_t G, uint64_t seed_r, uint64_t first, uint64_t n_reads,
                         int L, double err, int n_pct, int fmt, uint8_t *d_asc, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif
