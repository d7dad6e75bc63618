// bgzf.cpp - see bgzf.h.  Member layout (RFC 1952 + the SAM specification's BGZF section): 10 fixed header bytes with
// FLG.FEXTRA set, XLEN, extra subfields among which {'B','C', SLEN = 2, BSIZE = member size - 1}, deflate data, CRC32, ISIZE.
#include "bgzf.h"
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>

namespace yakb {

static inline uint32_t le16(const unsigned char *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8; }
static inline uint32_t le32(const unsigned char *p) { return le16(p) | le16(p + 2) << 16; }

size_t BgzfPool::block_size(const unsigned char *p, size_t n)
{
	if (n < 28 || p[0] != 31 || p[1] != 139 || p[2] != 8 || !(p[3] & 4)) return 0;
	const size_t xend = 12 + (size_t)le16(p + 10);
	if (xend > n) return 0;
	for (size_t i = 12; i + 4 <= xend; i += 4 + le16(p + i + 2)) {
		if (p[i] == 'B' && p[i + 1] == 'C' && le16(p + i + 2) == 2 && i + 6 <= xend) {
			const size_t bs = (size_t)le16(p + i + 4) + 1;
			return bs >= xend + 2 + 8 && bs <= n ? bs : 0;
		}
	}
	return 0;
}

bool BgzfPool::open(const char *fn, int threads, size_t job_bytes)
{
	close();
	struct stat st;
	fd_ = ::open(fn, O_RDONLY);
	if (fd_ < 0) return false;
	if (fstat(fd_, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size < 28) { close(); return false; }
	size_ = (size_t)st.st_size;
	void *m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
	if (m == MAP_FAILED) { close(); return false; }
	map_ = (const unsigned char*)m;
	if (block_size(map_, size_) == 0) { close(); return false; }
	madvise(m, size_, MADV_SEQUENTIAL);
	if (threads <= 0) {
		const char *e = getenv("YAKB_PARSE_THREADS");
		threads = e && atoi(e) > 0 ? atoi(e) : (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
	}
	job_bytes_ = std::max<size_t>(job_bytes, 1);
	scan_off_ = 0; plan_end_ = stop_ = to_tail_ = false; n_planned_ = n_consumed_ = 0;
	ring_.assign((size_t)threads + 3, Job());
	for (int i = 0; i < threads; ++i) workers_.emplace_back([this] { work_(); });
	return true;
}

void BgzfPool::close()
{
	{ std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
	cv_.notify_all();
	for (auto &t : workers_) if (t.joinable()) t.join();
	workers_.clear();
	ring_.clear();
	if (tail_init_) inflateEnd(&tz_);
	tail_init_ = tail_on_ = tail_end_ = false;
	if (map_) munmap((void*)map_, size_);
	map_ = nullptr;
	if (fd_ >= 0) ::close(fd_);
	fd_ = -1;
	stop_ = false;
}

bool BgzfPool::plan_(Job &j)
{
	size_t off = scan_off_, out = 0;
	int nb = 0;
	while (off < size_) {
		const size_t bs = block_size(map_ + off, size_ - off);
		if (bs == 0) break;
		const uint32_t isize = le32(map_ + off + bs - 4);
		if (isize > (1u << 26)) break; // not something a BGZF writer produces: leave it to zlib
		off += bs; out += isize; ++nb;
		if (out >= job_bytes_ || nb >= 4096) break;
	}
	if (nb == 0) return false;
	j.in_off = scan_off_; j.in_len = off - scan_off_; j.out_len = out;
	j.produced = 0; j.ready = j.fatal = false;
	scan_off_ = off;
	return true;
}

void BgzfPool::inflate_(Job &j, z_stream &zs)
{
	if (j.out.size() < j.out_len + 8) j.out.resize(j.out_len + 8); // never an empty buffer: inflate() rejects a null next_out
	size_t off = j.in_off, pos = 0;
	const size_t end = j.in_off + j.in_len;
	while (off < end) {
		const size_t bs = block_size(map_ + off, end - off);
		if (bs == 0 || inflateReset(&zs) != Z_OK) { j.fatal = true; break; }
		zs.next_in = (Bytef*)(map_ + off); zs.avail_in = (uInt)bs;
		zs.next_out = j.out.data() + pos; zs.avail_out = (uInt)(j.out_len - pos);
		const int ret = inflate(&zs, Z_FINISH); // gzip wrapper: zlib checks CRC32 and ISIZE of the member itself
		// anything but "exactly this member, exactly these bytes" (broken data, but also a BSIZE or ISIZE field that lies,
		// which gzread never looks at): the member is read again by the sequential path, which behaves like zlib
		if (ret != Z_STREAM_END || zs.avail_in != 0) { j.fatal = true; break; }
		pos = j.out_len - zs.avail_out;
		off += bs;
	}
	j.produced = pos;
	j.fail_off = off;
}

void BgzfPool::work_()
{
	z_stream zs;
	memset(&zs, 0, sizeof(zs));
	if (inflateInit2(&zs, 15 + 16) != Z_OK) return;
	for (;;) {
		Job *j;
		{
			std::unique_lock<std::mutex> lk(mu_);
			// job n_planned_ goes to slot n_planned_ % N, which is free once job n_planned_ - N has been consumed
			cv_.wait(lk, [this] { return stop_ || plan_end_ || n_planned_ - n_consumed_ < (int64_t)ring_.size(); });
			if (stop_ || plan_end_) break;
			j = &ring_[(size_t)(n_planned_ % (int64_t)ring_.size())];
			if (!plan_(*j)) { plan_end_ = true; cv_.notify_all(); break; }
			++n_planned_;
		}
		inflate_(*j, zs);
		{
			std::lock_guard<std::mutex> lk(mu_);
			j->ready = true;
			if (j->fatal) plan_end_ = true; // what follows a member the pool could not read is not the pool's to read
		}
		cv_.notify_all();
	}
	inflateEnd(&zs);
}

int64_t BgzfPool::next(std::vector<unsigned char> &out, bool *last)
{
	*last = false;
	for (;;) {
		std::unique_lock<std::mutex> lk(mu_);
		if (to_tail_) break;
		cv_.wait(lk, [this] {
			return (n_consumed_ < n_planned_ && ring_[(size_t)(n_consumed_ % (int64_t)ring_.size())].ready) || (plan_end_ && n_consumed_ == n_planned_);
		});
		if (n_consumed_ == n_planned_) break; // the BGZF part of the file has been delivered
		Job &j = ring_[(size_t)(n_consumed_ % (int64_t)ring_.size())];
		out.swap(j.out);
		const int64_t n = (int64_t)j.produced;
		if (j.fatal) { to_tail_ = true; scan_off_ = j.fail_off; } // (plan_end_ is set: scan_off_ is not moved any more)
		j.ready = false;
		++n_consumed_;
		cv_.notify_all();
		if (n > 0) return n;
	}
	// The member at scan_off_ is not a BGZF block: the rest of the file is inflated here, one member after the other.
	// Like gzread, whatever follows a member is read if it is another gzip member and ignored otherwise (zlib gz_look);
	// everything that inflates before an error or the end of a truncated file is delivered.
	if (!tail_on_) {
		tail_on_ = true;
		memset(&tz_, 0, sizeof(tz_));
		if (inflateInit2(&tz_, 15 + 16) != Z_OK) tail_end_ = true;
		else tail_init_ = true;
		tail_pos_ = scan_off_;
		tail_member_start_ = true;
	}
	const size_t want = std::max<size_t>(job_bytes_, 1u << 16);
	if (out.size() < want) out.resize(want);
	size_t n = 0;
	while (!tail_end_ && n < want) {
		if (tail_member_start_) {
			if (tail_pos_ + 1 >= size_ || map_[tail_pos_] != 31 || map_[tail_pos_ + 1] != 139 || inflateReset(&tz_) != Z_OK) { tail_end_ = true; break; }
			tail_member_start_ = false;
		}
		const size_t in = std::min<size_t>(size_ - tail_pos_, 1u << 30);
		tz_.next_in = (Bytef*)(map_ + tail_pos_); tz_.avail_in = (uInt)in;
		tz_.next_out = out.data() + n; tz_.avail_out = (uInt)(want - n);
		const int ret = inflate(&tz_, Z_NO_FLUSH);
		tail_pos_ += in - tz_.avail_in;
		n = want - tz_.avail_out;
		if (ret == Z_STREAM_END) tail_member_start_ = true;
		else if (ret != Z_OK || (tz_.avail_in == 0 && tail_pos_ >= size_ && tz_.avail_out != 0)) tail_end_ = true; // broken, or cut short
	}
	*last = tail_end_;
	return (int64_t)n;
}

} // namespace yakb
