// fastx_par.h - multi-threaded FASTA/FASTQ ingest for uncompressed files with EXACTLY the record
// semantics of the sequential reader (kseq.h:192-232).
//
// The file is cut into fixed-size blocks.  Workers parse their block speculatively from a guessed
// record start; a sequential stitcher then replays the (tiny) gap between the block start and the
// guess with the TRUE parser state carried over from the previous block and accepts the speculative
// result only if the true state arrives at the guess exactly where a record begins - otherwise it
// re-parses that block sequentially.  The parser is a deterministic state machine, so accepted
// blocks are byte-for-byte what the sequential reader would have produced, for any input.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include <memory>
#include <mutex>

namespace yakb {

// the record grammar as a resumable state machine over memory buffers
struct FastxCore {
	enum { S_FIND, S_NAME, S_SEQ, S_PLUS, S_QUAL };
	int st = S_FIND, last = 0, last_qual = 0;
	bool bol = true, stopped = false, qline_nonempty = false;
	// stopped by a FASTQ record with a truncated quality string (kseq's -2): index into the buffer of that feed() call
	// where kseq would go on (just behind the quality lines it read), or 0 after settle(); -1 otherwise.  To resume the
	// way a later kseq_read call does (kseq.h:192-199), clear `stopped` and feed from there: the state is S_FIND, last 0.
	int64_t bad_at = -1;
	int64_t qual_len = 0, qual_lines = 0, cur_len = 0; // cur_len: sequence length of the open record
	std::string rec;            // bytes of the open record carried between feed() calls
	// consume n bytes; completed records of length >= min_len are appended to out as "SEQ\n"
	void feed(const unsigned char *p, size_t n, int min_len, std::vector<uint8_t> &out, int64_t *n_seq);
	// end of input: close the open record the way kseq_read would
	void finish(int min_len, std::vector<uint8_t> &out, int64_t *n_seq);
	// a FASTQ record whose quality is complete is closed lazily by the next byte; do it now
	void settle(int min_len, std::vector<uint8_t> &out, int64_t *n_seq);
	bool at_record_boundary() const { return !stopped && ((st == S_FIND && last == 0) || (st == S_SEQ && bol)); }
};

struct BlockJob {
	uint64_t blk = ~0ull;         // block number this slot holds
	int state = 0;                // 0 free, 1 being parsed, 2 parsed, 3 stitched, its output still being copied out
	int min_len = 0;              // the record-length filter it was parsed with
	size_t n = 0, q = 0;
	FastxCore spec;               // state after the speculative parse
	std::vector<uint8_t> out;     // its output
	int64_t nseq = 0;
	// second speculation, for the bytes before the guess (all of the block when it holds no record start, as inside a
	// chromosome-sized sequence): "the block starts in the middle of a sequence".  mid_ok: [0, q) is nothing but sequence
	// lines; mid = their bases; mid_bol_in / mid_bol_out = at a line start before / after them.
	bool mid_ok = false, mid_bol_in = false, mid_bol_out = false;
	std::vector<uint8_t> mid;
};

// Streaming design: the file is memory-mapped; a pool of workers parses blocks ahead of the consumer
// (a ring of slots bounds the distance) for the whole life of the reader, and also carries out the big
// copies into the caller's buffer.  fill() is the in-order stitcher.
class ParallelFastx {
public:
	ParallelFastx();
	~ParallelFastx();
	ParallelFastx(const ParallelFastx&) = delete;
	ParallelFastx &operator=(const ParallelFastx&) = delete;
	// false if the file cannot be opened or is not a plain regular file (gzip, stdin): use FastxReader
	bool open(const char *fn, size_t block_bytes = 0, int threads = 0);
	void close();
	// same contract as FastxReader::fill
	size_t fill(uint8_t *dst, size_t cap, size_t target, int min_len, int64_t *n_seq, bool *done, size_t *need);
	uint64_t mis_speculations() const { return n_redo_; }
	// the reference's -K (count.c:106): it decides whether reading goes on behind a truncated FASTQ record, see ref_resumes()
	void set_ref_chunk(int64_t chunk_size) { ref_chunk_ = chunk_size > 0 ? chunk_size : 1; }
	// the general form: `workers` pipeline threads (3 in yak count, count.c:162; 2 in qv and the other scanners, qv.c:126),
	// and which records a step-0 call counts: those of at least flow_min_len bases (k in count.c:95; 0 for bseq_read,
	// bseq.c:33-57, which keeps every record), or fill()'s own min_len when flow_min_len < 0
	void set_ref_flow(int64_t chunk_size, int workers, int flow_min_len) { set_ref_chunk(chunk_size); ref_workers_ = ref_workers0_ = workers; flow_min_len_ = flow_min_len; }

private:
	struct Impl;
	bool work_one(std::unique_lock<std::mutex> &lk, bool may_parse);
	bool ref_resumes(uint64_t bad_off, int min_len);
	void feed_true(const unsigned char *buf, size_t len, uint64_t file_off, int min_len, std::vector<uint8_t> &out, int64_t *ns);
	int64_t ref_chunk_ = 10000000;   // yak_copt_init's default (misc.c:31)
	int ref_workers_ = 3, ref_workers0_ = 3, flow_min_len_ = -1; // count.c:162
	uint64_t anchor_off_ = 0;        // where the reference's current step-0 call (or an earlier one) began: file start, or behind the last bad record
	Impl *im_ = nullptr;
	int fd_ = -1;
	const unsigned char *map_ = nullptr;
	uint64_t size_ = 0, nblocks_ = 0, true_blocks_ = 0; // true_blocks_: blocks stitched so far
	size_t block_ = 0;
	int threads_ = 1;
	FastxCore true_;                 // exact parser state at the end of the last stitched block
	std::vector<uint8_t> spill_;     // parsed output that did not fit the caller's buffer yet
	std::vector<uint8_t> gap_;       // the stitcher's scratch
	size_t spill_pos_ = 0;
	int64_t spill_seq_ = 0;          // records inside spill_ not yet reported
	bool finished_ = false;
	uint64_t n_redo_ = 0;
};

} // namespace yakb
