// fastx_par.cpp - see fastx_par.h
#include "fastx_par.h"
#include <string.h>
#include <stdlib.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/stat.h>
#include <sys/mman.h>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <deque>
#include <algorithm>

namespace yakb {

// ---------------------------------------------------------------- the grammar (kseq.h:192-232)
// Sequence bytes go straight into `out`; `rs` marks where this call's part of the open record starts there.
// Between calls the open record's bytes live in `rec`, and they stay there while later calls add to the
// record (a chromosome-sized sequence spans hundreds of blocks: moving it back and forth would be quadratic);
// `rec` is spliced in front of out[rs..] once, when the record is closed.

// close the record whose bytes are rec + out[rs..]: keep it (terminator appended) or drop it
static inline void close_in_out(FastxCore &c, size_t &rs, bool keep, int min_len, std::vector<uint8_t> &out, int64_t *n_seq)
{
	if (keep && c.cur_len >= min_len) {
		if (!c.rec.empty()) out.insert(out.begin() + rs, (const uint8_t*)c.rec.data(), (const uint8_t*)c.rec.data() + c.rec.size());
		out.push_back('\n');
		++*n_seq;
	} else out.resize(rs);
	c.rec.clear();
	rs = out.size();
	c.cur_len = 0;
}

// close the record whose bytes are still carried in c.rec
static inline void close_carried(FastxCore &c, bool keep, int min_len, std::vector<uint8_t> &out, int64_t *n_seq)
{
	if (keep && c.cur_len >= min_len) {
		out.insert(out.end(), (const uint8_t*)c.rec.data(), (const uint8_t*)c.rec.data() + c.rec.size()); // same element type: one memmove
		out.push_back('\n');
		++*n_seq;
	}
	c.rec.clear();
	c.cur_len = 0;
}

void FastxCore::feed(const unsigned char *p, size_t n, int min_len, std::vector<uint8_t> &out, int64_t *n_seq)
{
	size_t i = 0, rs = out.size();
	while (i < n && !stopped) {
		if (st == S_FIND) {
			if (last) { st = S_NAME; last = 0; rs = out.size(); cur_len = 0; continue; } // header character consumed already
			while (i < n && p[i] != '>' && p[i] != '@') ++i; // normally the very next byte
			if (i < n) { ++i; st = S_NAME; rs = out.size(); cur_len = 0; }
		} else if (st == S_NAME || st == S_PLUS) { // rest of the header / '+' line
			const unsigned char *nl = (const unsigned char*)memchr(p + i, '\n', n - i);
			if (!nl) { i = n; break; }
			i = nl - p + 1;
			if (st == S_NAME) { st = S_SEQ; bol = true; } else { st = S_QUAL; qual_len = 0; qual_lines = 0; bol = true; qline_nonempty = false; }
		} else if (st == S_SEQ) {
			if (bol) {
				const unsigned char c = p[i];
				if (c == '\n') { ++i; continue; }
				if (c == '>' || c == '@') { ++i; close_in_out(*this, rs, true, min_len, out, n_seq); last = c; st = S_FIND; continue; }
				if (c == '+') { ++i; st = S_PLUS; continue; }
				bol = false;
			}
			const unsigned char *nl = (const unsigned char*)memchr(p + i, '\n', n - i);
			const size_t len = nl ? (size_t)(nl - (p + i)) : n - i;
			out.insert(out.end(), p + i, p + i + len);
			cur_len += (int64_t)len;
			i += len + (nl ? 1 : 0);
			if (nl) { // kseq.h:146: a CR before the line end goes; the line's last byte may have arrived in an earlier call
				if (cur_len > 1) {
					if (out.size() > rs) { if (out.back() == '\r') { out.pop_back(); --cur_len; } }
					else if (!rec.empty() && rec.back() == '\r') { rec.pop_back(); --cur_len; }
				}
				bol = true;
			}
		} else { // S_QUAL: only lengths matter
			if (bol && qual_lines > 0 && qual_len >= cur_len) { // kseq.h:224
				const bool ok = qual_len == cur_len;
				close_in_out(*this, rs, ok, min_len, out, n_seq);
				st = S_FIND; last = 0;
				if (!ok) { stopped = true; bad_at = (int64_t)i; } // kseq's -2: the caller decides whether reading goes on
				continue;
			}
			const unsigned char *nl = (const unsigned char*)memchr(p + i, '\n', n - i);
			const size_t len = nl ? (size_t)(nl - (p + i)) : n - i;
			if (len > 0) { qual_len += len; last_qual = p[i + len - 1]; qline_nonempty = true; }
			bol = false;
			i += len + (nl ? 1 : 0);
			if (nl) { // the CR rule looks at the whole line, which may have arrived in pieces
				if (qline_nonempty && qual_len > 1 && last_qual == '\r') --qual_len;
				bol = true; ++qual_lines; qline_nonempty = false;
			}
		}
	}
	// park the open record's bytes until the next call
	const bool open = !stopped && (st == S_SEQ || st == S_PLUS || st == S_QUAL);
	if (open) { if (out.size() > rs) rec.append((const char*)out.data() + rs, out.size() - rs); }
	else rec.clear();
	out.resize(rs);
}

void FastxCore::settle(int min_len, std::vector<uint8_t> &out, int64_t *n_seq)
{
	if (!stopped && st == S_QUAL && bol && qual_lines > 0 && qual_len >= cur_len) {
		const bool ok = qual_len == cur_len;
		close_carried(*this, ok, min_len, out, n_seq);
		st = S_FIND; last = 0;
		if (!ok) { stopped = true; bad_at = 0; }
	}
}

void FastxCore::finish(int min_len, std::vector<uint8_t> &out, int64_t *n_seq)
{
	if (stopped) return;
	if (st == S_SEQ) { // FASTA record ended by EOF
		if (!bol && cur_len > 1 && !rec.empty() && rec.back() == '\r') { rec.pop_back(); --cur_len; }
		close_carried(*this, true, min_len, out, n_seq);
	} else if (st == S_QUAL) { // kseq.h:224-226: EOF ends the quality; lengths must agree
		if (!bol && qline_nonempty && qual_len > 1 && last_qual == '\r') --qual_len;
		close_carried(*this, qual_len == cur_len, min_len, out, n_seq);
	} else { rec.clear(); cur_len = 0; } // header only (empty sequence), or '+' line without quality (kseq.h:222)
	st = S_FIND; last = 0; stopped = true;
}

// ---------------------------------------------------------------- speculative block parse

// first position in [0,n) that looks like the start of a record: a line start holding '>' or '@'
// with a plausible continuation.  Returns n when there is none.  A wrong guess costs time only.
static size_t guess_record_start(const unsigned char *p, size_t n, bool first_block)
{
	size_t i = 0;
	if (!first_block) { // move to the first line start inside the block
		const unsigned char *nl = (const unsigned char*)memchr(p, '\n', n);
		if (!nl) return n;
		i = nl - p + 1;
	}
	while (i < n) {
		const unsigned char c = p[i];
		const unsigned char *e1 = (const unsigned char*)memchr(p + i, '\n', n - i);
		if (c == '>' || c == '@') {
			if (!e1 || (size_t)(e1 - p) + 1 >= n) return i; // cannot look further: let the stitcher judge
			const unsigned char *l2 = e1 + 1;
			if (c == '>') { if (*l2 != '@' && *l2 != '+' && *l2 != '>') return i; } // not a quality line that starts with '>'
			else { // FASTQ: the line after the sequence line starts with '+'
				const unsigned char *e2 = (const unsigned char*)memchr(l2, '\n', n - (l2 - p));
				if (!e2 || (size_t)(e2 - p) + 1 >= n) return i;
				if (e2[1] == '+' || *l2 == '>' || *l2 == '@') return i; // ('@' header then another header: empty record)
			}
		}
		if (!e1) return n;
		i = e1 - p + 1;
	}
	return n;
}

// The bytes of p[0..n) read as the inside of a sequence (state S_SEQ of the grammar, kseq.h:206-213): line ends dropped,
// a CR before a line end dropped too (kseq.h:146; the caller checks that the record already held two bases, the rule's
// other condition).  False when anything else shows up - a line starting with '>', '@' or '+' - or when the first line
// end's CR would sit in the previous block.
static bool mid_sequence_parse(const unsigned char *p, size_t n, bool bol, BlockJob &j)
{
	j.mid.clear();
	j.mid_bol_in = bol;
	size_t i = 0;
	while (i < n) {
		if (bol) {
			const unsigned char c = p[i];
			if (c == '>' || c == '@' || c == '+') return false;
			if (c == '\n') { ++i; continue; }
			bol = false;
		}
		const unsigned char *nl = (const unsigned char*)memchr(p + i, '\n', n - i);
		const size_t len = nl ? (size_t)(nl - (p + i)) : n - i;
		if (len == 0 && i == 0) return false; // the line ended in the previous block: its CR, if any, is not ours to see
		j.mid.insert(j.mid.end(), p + i, p + i + len);
		i += len + (nl ? 1 : 0);
		if (nl) { if (len > 0 && j.mid.back() == '\r') j.mid.pop_back(); bol = true; }
	}
	j.mid_bol_out = bol;
	return true;
}

// ---------------------------------------------------------------- the pool

struct CopyTask { const uint8_t *src; uint8_t *dst; size_t len; BlockJob *slot; };

struct ParallelFastx::Impl {
	std::mutex mu;
	std::condition_variable cv;          // one condition for everything: the events are rare (one per block)
	std::vector<BlockJob> ring;
	std::deque<CopyTask> copyq;
	std::vector<std::thread> workers;
	uint64_t next_parse = 0;             // next block a worker may take
	int copies_open = 0;                 // copy tasks queued or running
	int min_len = 0;
	bool started = false, stop = false;
};

ParallelFastx::ParallelFastx() {}
ParallelFastx::~ParallelFastx() { close(); }

bool ParallelFastx::open(const char *fn, size_t block_bytes, int threads)
{
	close();
	if (fn == nullptr || strcmp(fn, "-") == 0) return false;
	fd_ = ::open(fn, O_RDONLY);
	if (fd_ < 0) return false;
	struct stat sb;
	unsigned char magic[2] = {0, 0};
	if (fstat(fd_, &sb) != 0 || !S_ISREG(sb.st_mode) || (pread(fd_, magic, 2, 0) == 2 && magic[0] == 0x1f && magic[1] == 0x8b)) {
		close();
		return false;
	}
	size_ = (uint64_t)sb.st_size;
	if (size_ > 0) {
		void *m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
		if (m == MAP_FAILED) { close(); return false; }
		map_ = (const unsigned char*)m;
		madvise(m, size_, MADV_SEQUENTIAL);
	}
	block_ = block_bytes ? block_bytes : (2u << 20);
	nblocks_ = (size_ + block_ - 1) / block_;
	if (threads <= 0) {
		const char *e = getenv("YAKB_PARSE_THREADS");
		threads = e && atoi(e) > 0 ? atoi(e) : (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
	}
	threads_ = threads;
	finished_ = false; n_redo_ = 0; true_blocks_ = 0;
	ref_workers_ = ref_workers0_; anchor_off_ = 0;
	true_ = FastxCore();
	spill_.clear(); spill_pos_ = 0; spill_seq_ = 0;
	im_ = new Impl;
	im_->ring.resize((size_t)std::max(4 * threads_, 16));
	return true;
}

void ParallelFastx::close()
{
	if (im_) {
		{ std::lock_guard<std::mutex> lk(im_->mu); im_->stop = true; }
		im_->cv.notify_all();
		for (auto &th : im_->workers) th.join();
		delete im_;
		im_ = nullptr;
	}
	if (map_) munmap((void*)map_, size_);
	map_ = nullptr;
	if (fd_ >= 0) ::close(fd_);
	fd_ = -1;
}

// One unit of pool work, lock held on entry and on exit; false if there was nothing to do.
// Copies come first: they free ring slots, which is what parsing waits for.
bool ParallelFastx::work_one(std::unique_lock<std::mutex> &lk, bool may_parse)
{
	Impl &im = *im_;
	if (!im.copyq.empty()) {
		CopyTask t = im.copyq.front(); im.copyq.pop_front();
		lk.unlock();
		memcpy(t.dst, t.src, t.len);
		lk.lock();
		t.slot->state = 0; --im.copies_open;
		im.cv.notify_all();
		return true;
	}
	if (may_parse && im.next_parse < nblocks_) {
		BlockJob &j = im.ring[im.next_parse % im.ring.size()];
		if (j.state == 0) {
			const uint64_t b = im.next_parse++;
			const int ml = im.min_len;
			j.state = 1; j.blk = b;
			lk.unlock();
			const uint64_t off = b * (uint64_t)block_;
			const unsigned char *raw = map_ + off;
			j.n = (size_t)std::min<uint64_t>(block_, size_ - off);
			j.q = guess_record_start(raw, j.n, off == 0);
			j.spec = FastxCore();
			j.out.clear();
			j.nseq = 0; j.min_len = ml;
			if (j.q < j.n) j.spec.feed(raw + j.q, j.n - j.q, ml, j.out, &j.nseq);
			j.mid_ok = off > 0 && j.q >= 4096 && mid_sequence_parse(raw, j.q, raw[-1] == '\n', j);
			lk.lock();
			j.state = 2;
			im.cv.notify_all();
			return true;
		}
	}
	return false;
}

// What the reference does behind a truncated FASTQ record (kseq's -2) depends on its pipeline (count.c:88-110, 162;
// kthread.c:119): the step-0 call that met the record ends there; if it had collected records (>= k bases each) they are
// processed and the next call resumes at the next header character; a call that collected nothing retires one of the three
// pipeline workers, and the third such call (the end of the file makes three at once) ends the input.  "Collected nothing"
// needs the lengths of the records since the call began (a call also ends once it holds chunk_size bases), which the block
// outputs do not keep - so on a bad record, and only then, the stretch since the last decision is parsed again, sequentially.
bool ParallelFastx::ref_resumes(uint64_t bad_off, int min_len)
{
	if (ref_workers0_ <= 0) return false; // a plain read loop (yak_recount, count.c:176): the first truncated record ends it
	FastxCore e;
	std::vector<uint8_t> scratch;
	int64_t sum = 0, cnt = 0; // bases and records the reference's current call holds
	const int fml = flow_min_len_ >= 0 ? flow_min_len_ : min_len;
	for (uint64_t off = anchor_off_; off < bad_off; ) {
		const size_t len = (size_t)std::min<uint64_t>(1u << 20, bad_off - off);
		int64_t ns = 0;
		scratch.clear();
		e.feed(map_ + off, len, fml, scratch, &ns);
		for (const uint8_t *p = scratch.data(), *end = p + scratch.size(); p < end; ) { // "SEQ\n" of every record the call counts
			const uint8_t *nl = (const uint8_t*)memchr(p, '\n', end - p);
			sum += nl - p; ++cnt;
			if (sum >= ref_chunk_) sum = cnt = 0; // count.c:106 / bseq.c:53: that call is full, the next one starts
			p = nl + 1;
		}
		off += len;
	}
	anchor_off_ = bad_off;
	if (cnt == 0 && --ref_workers_ <= 0) return false;
	return true;
}

// true_.feed() over a stretch of the file, going on behind truncated records where the reference does
void ParallelFastx::feed_true(const unsigned char *buf, size_t len, uint64_t file_off, int min_len, std::vector<uint8_t> &out, int64_t *ns)
{
	size_t pos = 0;
	for (;;) {
		true_.feed(buf + pos, len - pos, min_len, out, ns);
		if (!(true_.stopped && true_.bad_at >= 0)) return;
		const size_t at = pos + (size_t)true_.bad_at;
		true_.bad_at = -1;
		if (!ref_resumes(file_off + at, min_len)) return; // the input ends here for the reference too
		true_.stopped = false;
		pos = at;
	}
}

size_t ParallelFastx::fill(uint8_t *dst, size_t cap, size_t target, int min_len, int64_t *n_seq, bool *done, size_t *need)
{
	Impl &im = *im_;
	size_t n = 0;
	*done = false;
	if (need) *need = 0;
	if (target > cap) target = cap;
	const size_t W = im.ring.size();

	if (!im.started) { // the record-length filter is known from here on: start the pool
		std::lock_guard<std::mutex> lk(im.mu);
		im.started = true; im.min_len = min_len;
		for (int t = 0; t < threads_; ++t)
			im.workers.emplace_back([this]() {
				std::unique_lock<std::mutex> lk(im_->mu);
				while (!im_->stop) if (!work_one(lk, true)) im_->cv.wait(lk);
			});
	} else if (im.min_len != min_len) { std::lock_guard<std::mutex> lk(im.mu); im.min_len = min_len; } // blocks parsed with the old value are redone below

	// parsed bytes go straight to the caller's buffer while they fit, else to spill_
	auto emit = [&](const uint8_t *p, size_t len, int64_t recs) {
		if (len == 0) return;
		if (spill_.empty() && n + len <= cap) { memcpy(dst + n, p, len); n += len; *n_seq += recs; }
		else { spill_.insert(spill_.end(), p, p + len); spill_seq_ += recs; }
	};
	auto wait_copies = [&]() { // the caller may read dst only after every copy into it has landed
		std::unique_lock<std::mutex> lk(im.mu);
		while (im.copies_open > 0) if (!work_one(lk, false)) im.cv.wait(lk);
	};
	uint64_t consumed = true_blocks_;
	std::vector<uint8_t> &gap = gap_; // a member: its capacity (a whole chromosome, at times) survives the call
	for (;;) {
		// hand over what a previous call could not place, whole records only
		if (spill_pos_ < spill_.size()) {
			size_t room = cap - n, avail = spill_.size() - spill_pos_, take = std::min(room, avail);
			if (take < avail) { // cut at a record boundary
				const uint8_t *b = spill_.data() + spill_pos_;
				size_t k = take;
				while (k > 0 && b[k - 1] != '\n') --k;
				take = k;
			}
			if (take == 0 && n == 0) { // not even one record fits
				const uint8_t *b = spill_.data() + spill_pos_;
				const uint8_t *nl = (const uint8_t*)memchr(b, '\n', avail);
				if (need) *need = (nl ? (size_t)(nl - b) : avail) + 1;
				true_blocks_ = consumed;
				return 0;
			}
			memcpy(dst + n, spill_.data() + spill_pos_, take);
			if (take == avail) { *n_seq += spill_seq_; spill_seq_ = 0; }
			else { // records in the delivered part = its newlines
				int64_t c = 0;
				for (const uint8_t *b = dst + n, *e = b + take; b < e; ++c) { b = (const uint8_t*)memchr(b, '\n', e - b); if (!b) break; ++b; }
				*n_seq += c; spill_seq_ -= c;
			}
			n += take; spill_pos_ += take;
			if (spill_pos_ < spill_.size()) break; // caller's buffer is full
			spill_.clear(); spill_pos_ = 0;
		}
		if (n >= target) break;
		if (finished_) { *done = true; break; }
		if (consumed == nblocks_) { // end of input: close the open record
			int64_t gs = 0;
			gap.clear();
			true_.finish(min_len, gap, &gs);
			emit(gap.data(), gap.size(), gs);
			finished_ = true;
			continue;
		}
		// the next block in file order, parsed by the pool (the caller helps while it waits)
		BlockJob &j = im.ring[consumed % W];
		{
			std::unique_lock<std::mutex> lk(im.mu);
			while (!(j.state == 2 && j.blk == consumed)) if (!work_one(lk, true)) im.cv.wait(lk);
		}
		const uint64_t blk_off = consumed * (uint64_t)block_;
		const unsigned char *raw = map_ + blk_off;
		const size_t jq = j.q, jn = j.n; // the slot may be handed back (and refilled) as soon as its copy is queued
		int64_t gs = 0;
		bool queued = false;
		gap.clear();
		if (j.mid_ok && !true_.stopped && true_.st == FastxCore::S_SEQ && true_.bol == j.mid_bol_in && true_.cur_len >= 2) {
			// the bytes before the guess continue an open sequence, as the worker assumed: take its bases
			true_.rec.append((const char*)j.mid.data(), j.mid.size());
			true_.cur_len += (int64_t)j.mid.size();
			true_.bol = j.mid_bol_out;
		} else feed_true(raw, jq, blk_off, min_len, gap, &gs); // the gap before the guess, with the true state
		if (jq < jn) {
			true_.settle(min_len, gap, &gs);
			if (true_.stopped && true_.bad_at >= 0) { // the record that ends right at the guess has a truncated quality
				true_.bad_at = -1;
				if (ref_resumes(blk_off + jq, min_len)) true_.stopped = false;
			}
			if (j.min_len == min_len && true_.at_record_boundary()) { // the guess was a real record start: adopt the speculative result
				if (true_.st == FastxCore::S_SEQ && true_.rec.size() >= (1u << 16) && true_.cur_len >= min_len) {
					// a long carried record ends here: straight from the carry buffer to where it goes, not through `gap`
					emit(gap.data(), gap.size(), gs);
					gap.clear(); gs = 0;
					const uint8_t *r = (const uint8_t*)true_.rec.data();
					const size_t len = true_.rec.size();
					if (spill_.empty() && n + len + 1 <= cap) { memcpy(dst + n, r, len); n += len; dst[n++] = '\n'; ++*n_seq; }
					else { spill_.insert(spill_.end(), r, r + len); spill_.push_back('\n'); ++spill_seq_; }
					true_.rec.clear(); true_.cur_len = 0;
				} else if (true_.st == FastxCore::S_SEQ) close_carried(true_, true, min_len, gap, &gs);
				emit(gap.data(), gap.size(), gs);
				{ // take over the speculative end state, but keep the carry buffer's capacity (it is freed otherwise, and the
				  // next chromosome grows it again page fault by page fault)
					std::string keep = std::move(true_.rec);
					keep.assign(j.spec.rec);
					true_ = std::move(j.spec); // before the slot can be handed back (a finished copy frees it)
					true_.rec = std::move(keep);
				}
				const int64_t spec_bad = true_.stopped ? true_.bad_at : -1; // the worker stopped at a truncated record
				if (j.out.size() >= (1u << 16) && spill_.empty() && n + j.out.size() <= cap) { // big: copied by the pool
					std::lock_guard<std::mutex> lk(im.mu);
					im.copyq.push_back({j.out.data(), dst + n, j.out.size(), &j});
					++im.copies_open; j.state = 3; queued = true;
					n += j.out.size(); *n_seq += j.nseq;
				} else emit(j.out.data(), j.out.size(), j.nseq);
				if (spec_bad >= 0) { // the rest of the block, with the true state, if the reference reads on
					true_.bad_at = -1;
					const size_t at = jq + (size_t)spec_bad;
					if (ref_resumes(blk_off + at, min_len)) {
						true_.stopped = false;
						int64_t rs = 0;
						gap.clear();
						feed_true(raw + at, jn - at, blk_off + at, min_len, gap, &rs);
						emit(gap.data(), gap.size(), rs);
					}
				}
			} else { // wrong guess: this block again, sequentially, from the true state
				feed_true(raw + jq, jn - jq, blk_off + jq, min_len, gap, &gs);
				emit(gap.data(), gap.size(), gs);
				++n_redo_;
			}
		} else emit(gap.data(), gap.size(), gs);
		{
			std::lock_guard<std::mutex> lk(im.mu);
			if (!queued) j.state = 0;
		}
		im.cv.notify_all();
		++consumed;
	}
	true_blocks_ = consumed;
	wait_copies();
	return n;
}

} // namespace yakb
