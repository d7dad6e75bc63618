// bgzf.h - blocked gzip (BGZF: bgzip, samtools, most sequencer pipelines) inflated by a pool of threads.
// A .gz file is one deflate stream per gzip member; zlib inflates it at ~0.3 GB/s of text on one core, which is
// what bounds `yak count reads.fq.gz` in the reference (kseq over gzread, count.c:150-151) and bounded it here.
// BGZF members are <= 64 KB each and carry their compressed size in the gzip header's extra field ('B','C'), so
// the member boundaries are known without inflating: workers take runs of members and inflate them side by side,
// the reader consumes the runs in file order.  The byte stream is exactly gzread's: a member that is not BGZF
// (plain gzip appended to the file, a truncated last block) or does not inflate to what its header announces
// is inflated sequentially from there on, member after member like zlib; broken data end the input after the bytes
// that inflate before the error.  (On a corrupt stream gzread - and with it the reference - also stops, but drops
// the output of the read call that met the error: up to 16 KB of text, kseq.h:82.)
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#include <zlib.h>

namespace yakb {

class BgzfPool {
public:
	BgzfPool() {}
	~BgzfPool() { close(); }
	// true if fn is a regular file whose first gzip member is a BGZF block; job_bytes = inflated bytes per unit of work
	bool open(const char *fn, int threads = 0, size_t job_bytes = 4u << 20);
	void close();
	// the next run of inflated bytes, swapped into `out` (its old storage is reused by the pool); returns the number of
	// valid bytes; *last is set when nothing follows.  The final call may return 0 with *last set.
	int64_t next(std::vector<unsigned char> &out, bool *last);
	int threads() const { return (int)workers_.size(); }
	// size in bytes of the BGZF block at p (n bytes available), 0 if p does not start with a BGZF member header
	static size_t block_size(const unsigned char *p, size_t n);

private:
	struct Job {
		size_t in_off = 0, in_len = 0, out_len = 0; // planned: bytes of the file, inflated bytes by the members' ISIZE fields
		size_t produced = 0;                        // inflated bytes actually in `out`
		size_t fail_off = 0;                        // fatal: file offset of the member that did not inflate as announced
		bool ready = false, fatal = false;
		std::vector<unsigned char> out;
	};
	bool plan_(Job &j);          // under mu_: cut the next run of members off the file; false when the BGZF part is over
	void work_();
	void inflate_(Job &j, z_stream &zs);
	int fd_ = -1;
	const unsigned char *map_ = nullptr;
	size_t size_ = 0, job_bytes_ = 0;
	size_t scan_off_ = 0;        // first byte no job covers yet
	bool plan_end_ = false;      // the member at scan_off_ is not BGZF (or the file ends there): no more jobs
	int64_t n_planned_ = 0, n_consumed_ = 0;
	bool stop_ = false, to_tail_ = false; // to_tail_: a member did not inflate as its header promised; the sequential path takes over there
	std::vector<Job> ring_;
	std::vector<std::thread> workers_;
	std::mutex mu_;
	std::condition_variable cv_;
	// sequential inflate of whatever follows the BGZF part (plain gzip members, a truncated block)
	z_stream tz_;
	bool tail_on_ = false, tail_init_ = false, tail_end_ = false, tail_member_start_ = true;
	size_t tail_pos_ = 0;
};

} // namespace yakb
