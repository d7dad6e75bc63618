// fastx.h - host-side FASTA/FASTQ reader with the record semantics of the reference's parser
// (kseq.h:192-232), restated around a large block buffer; plain or gzip input through zlib.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include <zlib.h>
#include "ref_flow.h"
#include "bgzf.h"

namespace yakb {

class FastxReader {
public:
	FastxReader() {}
	~FastxReader() { close(); }
	// NULL or "-" = stdin (count.c:151).  A regular file made of BGZF blocks is inflated by a pool of threads (bgzf.h):
	// bgzf_threads 0 = one per core, < 0 = zlib's sequential reader as for any other gzip file (also YAKB_NO_PBGZF=1)
	bool open(const char *fn, int bgzf_threads = 0, size_t bgzf_job_bytes = 4u << 20);
	int bgzf_threads() const;   // 0 when the input is not read through the BGZF pool
	// keep a copy of every block of (inflated) text the source delivers in fd, up to budget bytes (csrc/textcache.h);
	// tee_bytes(): the size of the copy once the source has delivered its last block and everything fit, else UINT64_MAX
	void tee_to(int fd, uint64_t budget) { tee_fd_ = fd; tee_budget_ = budget; tee_bytes_ = 0; tee_ok_ = fd >= 0; }
	uint64_t tee_bytes() const { return tee_ok_ && src_end_seen_ ? tee_bytes_ : UINT64_MAX; }
	void close();
	// next record: sequence bytes (line ends removed) in seq(); returns length, -1 at EOF,
	// -2 on a truncated quality string (kseq.h:189-191)
	int64_t next();
	const std::string &seq() const { return seq_; }
	const std::string &name() const { return name_; }
	// Bulk path of yak_count: append whole records (those with length >= min_len) to dst as
	// "SEQ\n" until at least `target` bytes are in, dst is full, or the input ends.  Sequence lines
	// go from the block buffer straight into dst.  Returns bytes appended; *n_seq += records kept;
	// *done = input exhausted (or a malformed FASTQ record stopped the parse, like kseq's -2).
	// A record that does not fit is carried over to the next call; if it can never fit, *need is
	// set to the dst size required.
	size_t fill(uint8_t *dst, size_t cap, size_t target, int min_len, int64_t *n_seq, bool *done, size_t *need);
	// fill() goes on behind a FASTQ record with a truncated quality string exactly where the reference's `yak count`
	// does: its step-0 call ends at such a record; a call that had collected nothing (records >= min_len bases; a call
	// is also full at chunk_size bases, count.c:106) retires one of the pipeline's three workers, the third ends the
	// input (count.c:109,162; kthread.c:119).  chunk_size is the reference's -K.
	void set_ref_chunk(int64_t chunk_size) { ref_chunk_ = chunk_size > 0 ? chunk_size : 1; flow_.chunk = ref_chunk_; }
	// workers = 3 is `yak count`; workers = 0 a plain read loop that ends at the first truncated record (yak_recount, count.c:176)
	void set_ref_workers(int workers) { flow_.workers = workers; }

private:
	// gzip / stdin input is read (and inflated) by a helper thread a few blocks ahead of the parser, so that
	// inflating and parsing overlap; YAKB_NO_READAHEAD=1 reads in the parsing thread
	struct Ahead;
	Ahead *ahead_ = nullptr;
	BgzfPool *bgzf_ = nullptr;
	bool src_last_ = false;    // the block read_block_() just returned is the last one
	bool src_end_seen_ = false; // the source has delivered its last block (the parse may stop earlier: truncated-record rule)
	int tee_fd_ = -1;
	uint64_t tee_budget_ = 0, tee_bytes_ = 0;
	bool tee_ok_ = false;
	int64_t read_source_();    // read_block_ without the tee
	int64_t read_block_();     // next block into buf_; returns its length (0 possible at the end of the input), sets src_last_
	int getc_();
	// append the rest of the current line to s (without the '\n'); false if nothing was left
	bool line_(std::string &s, int64_t *count_only);
	gzFile fp_ = nullptr;
	std::vector<unsigned char> buf_;
	int64_t beg_ = 0, end_ = 0;
	bool eof_ = false;
	int last_ = 0, last_qual_ = 0;
	std::string seq_, name_;
	// fill() state machine
	bool refill_();
	enum { S_FIND, S_NAME, S_SEQ, S_PLUS, S_QUAL } st_ = S_FIND;
	bool bol_ = true;          // at the beginning of a line
	bool qline_nonempty_ = false;
	int64_t cur_len_ = 0, qual_len_ = 0, qual_lines_ = 0;
	std::string carry_;        // a record that did not fit the caller's buffer
	bool carry_ready_ = false; // carry_ holds a COMPLETE record waiting for the next fill()
	bool in_carry_ = false;    // the current record is being collected in carry_
	int64_t ref_chunk_ = 10000000;               // yak_copt_init's default (misc.c:31)
	yakb_ref_flow_t flow_ = {3, 0, 0, 0, 10000000}; // count.c:162; min_len 0: fill() only reports records it keeps
};

} // namespace yakb
