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

namespace yakb {

static const size_t kBuf = 4u << 20;

bool FastxReader::open(const char *fn)
{
	close();
	fp_ = (fn == nullptr || strcmp(fn, "-") == 0) ? gzdopen(0, "r") : gzopen(fn, "r");
	if (!fp_) return false;
	gzbuffer(fp_, 1u << 20);
	buf_.resize(kBuf);
	beg_ = end_ = 0; eof_ = false; last_ = 0;
	return true;
}

void FastxReader::close()
{
	if (fp_) gzclose(fp_);
	fp_ = nullptr;
}

int FastxReader::getc_()
{
	if (beg_ >= end_) {
		if (eof_) return -1;
		beg_ = 0;
		int got = gzread(fp_, buf_.data(), (unsigned)buf_.size());
		end_ = got > 0 ? got : 0;
		if (end_ < (int64_t)buf_.size()) eof_ = true;
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
			int r = gzread(fp_, buf_.data(), (unsigned)buf_.size());
			end_ = r > 0 ? r : 0;
			if (end_ < (int64_t)buf_.size()) eof_ = true;
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
