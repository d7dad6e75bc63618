// fastx.cpp - see fastx.h.  Record grammar follows kseq.h:192-232:
//   * a record starts at the next '>' or '@' (searched anywhere when the previous record was FASTQ,
//     otherwise it is the header character that ended the previous sequence);
//   * the name ends at the first whitespace, the rest of that line is a comment;
//   * sequence lines follow until a line that begins with '>', '@' or '+'; empty lines are
//     skipped; one trailing '\r' is dropped when the accumulated length exceeds 1 (kseq.h:146);
//   * after '+': skip that line, then read quality lines until at least as long as the sequence;
//     a length mismatch is the -2 error.
#include "fastx.h"
#include <string.h>
#include <ctype.h>
#include <stdlib.h>
#include <unistd.h>
#include <thread>
#include <mutex>
#include <condition_variable>

namespace yakb {

static const size_t kBuf = 4u << 20;

struct FastxReader::Ahead {
	static const int N = 3;
	std::vector<unsigned char> blk[N];
	int64_t len[N];
	int head = 0, count = 0;
	bool done = false, stop = false; // done: the producer has delivered its last block
	std::mutex mu;
	std::condition_variable cv;
	std::thread th;
};

int FastxReader::bgzf_threads() const { return bgzf_ ? bgzf_->threads() : 0; }

bool FastxReader::open(const char *fn, int bgzf_threads, size_t bgzf_job_bytes)
{
	close();
	buf_.resize(kBuf);
	beg_ = end_ = 0; eof_ = src_last_ = src_end_seen_ = false; last_ = 0;
	yakb_ref_flow_init(&flow_, 3, ref_chunk_, 0);
	if (fn != nullptr && strcmp(fn, "-") != 0 && bgzf_threads >= 0 && !getenv("YAKB_NO_PBGZF")) {
		bgzf_ = new BgzfPool;
		if (bgzf_->open(fn, bgzf_threads, bgzf_job_bytes)) return true; // the pool runs ahead of the parser by itself
		delete bgzf_;
		bgzf_ = nullptr;
	}
	fp_ = (fn == nullptr || strcmp(fn, "-") == 0) ? gzdopen(0, "r") : gzopen(fn, "r");
	if (!fp_) return false;
	gzbuffer(fp_, 1u << 20);
	if (!getenv("YAKB_NO_READAHEAD")) {
		ahead_ = new Ahead;
		Ahead *a = ahead_;
		gzFile fp = fp_;
		a->th = std::thread([a, fp]() {
			for (int slot = 0;; slot = (slot + 1) % Ahead::N) {
				{
					std::unique_lock<std::mutex> lk(a->mu);
					a->cv.wait(lk, [a] { return a->stop || a->count < Ahead::N; });
					if (a->stop) return;
				}
				std::vector<unsigned char> &b = a->blk[slot];
				b.resize(kBuf);
				const int got = gzread(fp, b.data(), (unsigned)kBuf);
				std::lock_guard<std::mutex> lk(a->mu);
				a->len[slot] = got > 0 ? got : 0;
				++a->count;
				if (got < (int)kBuf) a->done = true; // a short read is the end (the same rule as the direct reads)
				a->cv.notify_all();
				if (a->done) return;
			}
		});
	}
	return true;
}

void FastxReader::close()
{
	if (ahead_) {
		{ std::lock_guard<std::mutex> lk(ahead_->mu); ahead_->stop = true; }
		ahead_->cv.notify_all();
		if (ahead_->th.joinable()) ahead_->th.join();
		delete ahead_;
		ahead_ = nullptr;
	}
	if (fp_) gzclose(fp_);
	fp_ = nullptr;
	delete bgzf_;
	bgzf_ = nullptr;
}

int64_t FastxReader::read_block_()
{
	const int64_t n = read_source_();
	if (src_last_) src_end_seen_ = true;
	if (tee_ok_ && n > 0) { // the second pass of a compressed file reads this copy instead of inflating again
		if (tee_bytes_ + (uint64_t)n > tee_budget_) tee_ok_ = false;
		else {
			int64_t done = 0;
			while (done < n) {
				const ssize_t w = write(tee_fd_, buf_.data() + done, (size_t)(n - done));
				if (w <= 0) { tee_ok_ = false; break; }
				done += w;
			}
			tee_bytes_ += (uint64_t)n;
		}
	}
	return n;
}

int64_t FastxReader::read_source_()
{
	if (bgzf_) return bgzf_->next(buf_, &src_last_);
	if (!ahead_) {
		const int got = gzread(fp_, buf_.data(), (unsigned)kBuf);
		src_last_ = got < (int)kBuf; // a short read is the end
		return got > 0 ? got : 0;
	}
	Ahead *a = ahead_;
	std::unique_lock<std::mutex> lk(a->mu);
	a->cv.wait(lk, [a] { return a->count > 0 || a->done; });
	if (a->count == 0) { src_last_ = true; return 0; }
	buf_.swap(a->blk[a->head]);      // the parser owns the block now; the producer refills the other vector
	const int64_t n = a->len[a->head];
	a->head = (a->head + 1) % Ahead::N;
	--a->count;
	src_last_ = a->done && a->count == 0; // the producer flags its last block in the same critical section that delivers it
	a->cv.notify_all();
	return n;
}

int FastxReader::getc_()
{
	if (beg_ >= end_) {
		if (eof_) return -1;
		beg_ = 0;
		end_ = read_block_();
		if (src_last_) eof_ = true;
		if (end_ == 0) return -1;
	}
	return buf_[beg_++];
}

bool FastxReader::line_(std::string &s, int64_t *count_only)
{
	bool got = false;
	for (;;) {
		if (beg_ >= end_) {
			if (eof_) break;
			beg_ = 0;
			end_ = read_block_();
			if (src_last_) eof_ = true;
			if (end_ == 0) break;
		}
		got = true;
		const unsigned char *p = buf_.data() + beg_;
		const unsigned char *nl = (const unsigned char*)memchr(p, '\n', end_ - beg_);
		const int64_t len = nl ? nl - p : end_ - beg_;
		if (count_only) {
			*count_only += len;
			if (len > 0) last_qual_ = p[len - 1];
		} else s.append((const char*)p, len);
		beg_ += len + (nl ? 1 : 0);
		if (nl) break;
	}
	return got;
}

bool FastxReader::refill_()
{
	if (beg_ < end_) return true;
	if (eof_) return false;
	beg_ = 0;
	end_ = read_block_();
	if (src_last_) eof_ = true;
	return end_ > 0;
}

size_t FastxReader::fill(uint8_t *dst, size_t cap, size_t target, int min_len, int64_t *n_seq, bool *done, size_t *need)
{
	size_t n = 0, rec_start = 0;
	*done = false;
	if (need) *need = 0;
	if (target > cap) target = cap;
	if (carry_ready_) { // the record that overflowed the previous buffer
		if (carry_.size() + 1 > cap) { if (need) *need = carry_.size() + 1; return 0; }
		memcpy(dst, carry_.data(), carry_.size());
		n = carry_.size(); dst[n++] = '\n'; ++*n_seq;
		carry_.clear(); carry_ready_ = false;
	}
	// close the record being parsed: keep it (terminator appended) or drop it
	auto finish_record = [&](bool keep) {
		if (keep && cur_len_ >= min_len) yakb_ref_flow_record(&flow_, cur_len_); // count.c:95,105-106
		if (in_carry_) {
			if (keep && cur_len_ >= min_len) carry_ready_ = true; else carry_.clear();
			in_carry_ = false;
		} else if (keep && cur_len_ >= min_len) { dst[n++] = '\n'; ++*n_seq; }
		else n = rec_start;
	};
	auto put = [&](const unsigned char *p, int64_t len) { // append sequence bytes of the current record
		if (!in_carry_ && n + (size_t)len + 1 > cap) { // would not fit (1 byte kept for the terminator): move to carry_
			carry_.assign((const char*)dst + rec_start, n - rec_start);
			n = rec_start; in_carry_ = true;
		}
		if (in_carry_) carry_.append((const char*)p, len); else { memcpy(dst + n, p, len); n += len; }
		cur_len_ += len;
	};
	auto drop_cr = [&]() { // kseq.h:146: one trailing CR once the accumulated length exceeds 1
		if (cur_len_ > 1) {
			if (in_carry_) { if (carry_.back() == '\r') { carry_.pop_back(); --cur_len_; } }
			else if (dst[n - 1] == '\r') { --n; --cur_len_; }
		}
	};
	for (;;) {
		if (carry_ready_) break;                       // buffer is full: hand over what we have
		if (st_ == S_FIND && n >= target) break;       // enough for this batch, stop between records
		if (!refill_()) {                              // end of input
			if (st_ == S_SEQ) { if (!bol_) drop_cr(); finish_record(true); } // FASTA record ended by EOF
			else if (st_ == S_QUAL) {                    // kseq.h:224-226: EOF ends the quality; lengths must agree
				if (!bol_ && qual_len_ > 1 && last_qual_ == '\r') --qual_len_;
				finish_record(qual_len_ == cur_len_);
			} else if (st_ == S_PLUS) finish_record(false); // kseq.h:222: no quality string at all
			st_ = S_FIND; last_ = 0;
			*done = !carry_ready_; // a last record that did not fit dst waits in carry_: the caller must come back for it
			break;
		}
		const unsigned char *p = buf_.data() + beg_;
		const int64_t avail = end_ - beg_;
		if (st_ == S_FIND) {
			if (last_) { st_ = S_NAME; bol_ = false; rec_start = n; cur_len_ = 0; in_carry_ = false; last_ = 0; continue; }
			// Fast path for what sequencers write: a four-line FASTQ record that lies whole in the buffer - header, one
			// sequence line, '+' line, one quality line of the same length, no CR.  Anything else (multi-line, FASTA,
			// empty or over-long lines, a record cut by the buffer's end or too big for dst) goes through the state
			// machine below, which this shortcut reproduces step for step.
			if (p[0] == '@') {
				const unsigned char *e = p + avail;
				const unsigned char *nl1 = (const unsigned char*)memchr(p, '\n', avail);
				const unsigned char *sq = nl1 ? nl1 + 1 : e;
				if (sq < e && *sq != '\n' && *sq != '>' && *sq != '@' && *sq != '+') {
					const unsigned char *nl2 = (const unsigned char*)memchr(sq, '\n', e - sq);
					if (nl2 && nl2 + 1 < e && nl2[1] == '+' && nl2[-1] != '\r') {
						const unsigned char *nl3 = (const unsigned char*)memchr(nl2 + 1, '\n', e - (nl2 + 1));
						const unsigned char *ql = nl3 ? nl3 + 1 : e;
						const int64_t len = nl2 - sq;
						if (ql + len < e && ql[len] == '\n' && ql[len - 1] != '\r' && memchr(ql, '\n', len) == nullptr) {
							if (len < min_len) { beg_ = ql + len + 1 - buf_.data(); bol_ = true; continue; } // dropped (count.c:95)
							if (n + (size_t)len + 1 <= cap) {
								memcpy(dst + n, sq, len);
								n += len; dst[n++] = '\n'; ++*n_seq;
								yakb_ref_flow_record(&flow_, len);
								beg_ = ql + len + 1 - buf_.data(); bol_ = true;
								continue;
							}
						}
					}
				}
			}
			int64_t i = 0;
			while (i < avail && p[i] != '>' && p[i] != '@') ++i;
			beg_ += i;
			if (i < avail) { ++beg_; st_ = S_NAME; rec_start = n; cur_len_ = 0; in_carry_ = false; }
		} else if (st_ == S_NAME || st_ == S_PLUS) {   // skip the rest of the header / '+' line
			const unsigned char *nl = (const unsigned char*)memchr(p, '\n', avail);
			if (!nl) { beg_ = end_; continue; }
			beg_ += nl - p + 1;
			if (st_ == S_NAME) { st_ = S_SEQ; bol_ = true; } else { st_ = S_QUAL; qual_len_ = 0; qual_lines_ = 0; bol_ = true; qline_nonempty_ = false; }
		} else if (st_ == S_SEQ) {
			if (bol_) {
				const unsigned char c = p[0];
				if (c == '\n') { ++beg_; continue; }
				if (c == '>' || c == '@') { ++beg_; finish_record(true); last_ = c; st_ = S_FIND; continue; }
				if (c == '+') { ++beg_; st_ = S_PLUS; continue; }
				bol_ = false;
			}
			const unsigned char *nl = (const unsigned char*)memchr(p, '\n', avail);
			const int64_t len = nl ? nl - p : avail;
			put(p, len);
			beg_ += len + (nl ? 1 : 0);
			if (nl) { drop_cr(); bol_ = true; }
		} else { // S_QUAL: only lengths matter
			if (bol_ && qual_lines_ > 0 && qual_len_ >= cur_len_) { // kseq.h:224: at least one line, stop once long enough
				const bool ok = qual_len_ == cur_len_;
				finish_record(ok);
				st_ = S_FIND; last_ = 0;
				if (!ok) { // kseq's -2 ends the reference's step-0 call; see set_ref_chunk()
					if (!yakb_ref_flow_bad(&flow_)) { eof_ = true; beg_ = end_; *done = true; break; }
					// else the next call resumes at the next header character (state S_FIND, last_ 0)
				}
				continue;
			}
			const unsigned char *nl = (const unsigned char*)memchr(p, '\n', avail);
			const int64_t len = nl ? nl - p : avail;
			if (len > 0) { qual_len_ += len; last_qual_ = p[len - 1]; qline_nonempty_ = true; }
			bol_ = false;
			beg_ += len + (nl ? 1 : 0);
			if (nl) { // the CR rule looks at the whole line, which may have arrived in pieces
				if (qline_nonempty_ && qual_len_ > 1 && last_qual_ == '\r') --qual_len_;
				bol_ = true; ++qual_lines_; qline_nonempty_ = false;
			}
		}
	}
	return n;
}

int64_t FastxReader::next()
{
	int c;
	if (last_ == 0) {
		while ((c = getc_()) != -1 && c != '>' && c != '@') {}
		if (c == -1) return -1;
		last_ = c;
	}
	seq_.clear(); name_.clear();
	bool any = false;
	while ((c = getc_()) != -1) { any = true; if (isspace(c)) break; name_.push_back((char)c); }
	if (!any) return -1;
	if (c != -1 && c != '\n') { std::string dummy; int64_t n = 0; line_(dummy, &n); }
	while ((c = getc_()) != -1 && c != '>' && c != '+' && c != '@') {
		if (c == '\n') continue;
		seq_.push_back((char)c);
		line_(seq_, nullptr);
		if (seq_.size() > 1 && seq_.back() == '\r') seq_.pop_back();
	}
	if (c == '>' || c == '@') last_ = c;
	if (c != '+') return (int64_t)seq_.size();
	while ((c = getc_()) != -1 && c != '\n') {}
	if (c == -1) return -2;
	int64_t ql = 0;
	for (;;) {
		std::string dummy;
		last_qual_ = 0;
		int64_t before = ql;
		if (!line_(dummy, &ql)) break;
		// the '\r' rule applies to the accumulated quality string (kseq.h:146)
		if (ql > 1 && ql > before && last_qual_ == '\r') --ql;
		if (ql >= (int64_t)seq_.size()) break;
	}
	last_ = 0;
	if (ql != (int64_t)seq_.size()) return -2;
	return (int64_t)seq_.size();
}

} // namespace yakb
