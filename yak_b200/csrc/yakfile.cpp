// yakfile.cpp - see yakfile.h
// The host side of yak_ch_restore_core (htab.c:419-472): header checks, then per sub-table {capacity, size, keys}.  The
// reference reads key by key; a table of human reads is 25 GB, so here the 2^pre headers are walked first (each tells where
// the next one lies) and the key arrays are then read by several threads straight to their place in one dense array.
// Short files read like the reference's unchecked freads leave things: a sub-table whose header is missing is empty, one
// whose keys are cut short has the whole keys that are there.  Flag modes map the counts (htab.c:448-469) afterwards.
// Returns 0, -1 (cannot open / shorter than the header), -2 (magic), -3 (counter bits).
#include "yakfile.h"
#include "../../include/yak.h"
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <unistd.h>
#include <algorithm>
#include <atomic>
#include <thread>

namespace yakb {

static double wall_now() { struct timeval tv; gettimeofday(&tv, nullptr); return tv.tv_sec + tv.tv_usec * 1e-6; }
static bool timing_on() { static int v = -1; if (v < 0) { const char *e = getenv("YAKB_TIMING"); v = e && atoi(e) > 0; } return v != 0; }

KeyBuf::~KeyBuf() { if (p) munmap(p, bytes); }
bool KeyBuf::alloc(size_t count)
{
	if (p) munmap(p, bytes);
	p = nullptr; n = count;
	bytes = (std::max<size_t>(count, 1) * 8 + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
	void *m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
	if (m == MAP_FAILED) { n = 0; return false; }
	madvise(m, bytes, MADV_HUGEPAGE);
	p = (uint64_t*)m;
	return true;
}

static bool pread_all(int fd, void *dst, size_t n, uint64_t at)
{
	char *p = (char*)dst;
	while (n) {
		const ssize_t r = pread(fd, p, std::min<size_t>(n, 1u << 30), (off_t)at);
		if (r <= 0) return false;
		p += r; at += (uint64_t)r; n -= (size_t)r;
	}
	return true;
}
int read_yak_file(const char *fn, int mode, int min_cnt, int mid_cnt, YakFile &yf, int threads, bool header_only)
{
	const int fd = open(fn, O_RDONLY);
	if (fd < 0) return -1;
	struct Close { int fd; ~Close() { close(fd); } } closer{fd};
	struct stat st;
	if (fstat(fd, &st) != 0) return -1;
	const bool regular = S_ISREG(st.st_mode);
	char head[16];
	if (!regular) { // a pipe (`yak qv <(...)`): no offsets to jump to - read it front to back
		FILE *fp = fdopen(dup(fd), "rb");
		if (!fp) return -1;
		struct FClose { FILE *f; ~FClose() { fclose(f); } } fcloser{fp};
		if (fread(head, 1, 4, fp) != 4) return -1;
		if (strncmp(head, YAK_MAGIC, 4) != 0) return -2;
		uint32_t t[3];
		if (fread(t, 4, 3, fp) != 3) return -1;
		yf.k = t[0]; yf.pre = t[1]; yf.counter_bits = t[2];
		if (t[2] != YAK_COUNTER_BITS) return -3;
		if (yf.pre < YAK_COUNTER_BITS || yf.pre > 30 || yf.k < 1 || yf.k > 63) return -1; // not a table yak_ch_init could have made (htab.c:17, main.c:46)
		const int P = 1 << yf.pre; // (a pipe cannot be opened twice: header_only is not honoured here, the body comes along)
		yf.caps.assign(P, 0); yf.off.assign(P + 1, 0);
		std::vector<uint64_t> tmp;
		for (int s = 0; s < P; ++s) {
			uint32_t u[2] = {0, 0};
			if (fread(u, 4, 2, fp) != 2) u[0] = u[1] = 0;
			yf.caps[s] = u[0];
			const size_t base = tmp.size();
			tmp.resize(base + u[1]);
			const size_t got = u[1] ? fread(tmp.data() + base, 8, u[1], fp) : 0;
			tmp.resize(base + got);
			yf.off[s + 1] = tmp.size();
		}
		if (!yf.keys.alloc(tmp.size())) return -1;
		if (!tmp.empty()) memcpy(yf.keys.data(), tmp.data(), tmp.size() * 8);
	} else {
		const uint64_t fsize = (uint64_t)st.st_size;
		if (fsize < 4 || !pread_all(fd, head, 4, 0)) return -1;
		if (strncmp(head, YAK_MAGIC, 4) != 0) return -2;
		if (fsize < 16 || !pread_all(fd, head + 4, 12, 4)) return -1;
		uint32_t t[3];
		memcpy(t, head + 4, 12);
		yf.k = t[0]; yf.pre = t[1]; yf.counter_bits = t[2];
		if (t[2] != YAK_COUNTER_BITS) return -3;
		// before anything is sized from the header of an untrusted file: what yak_ch_init could have made (htab.c:17, main.c:46)
		if (yf.pre < YAK_COUNTER_BITS || yf.pre > 30 || yf.k < 1 || yf.k > 63) return -1;
		if (header_only) return 0;
		const int P = 1 << yf.pre;
		yf.caps.assign(P, 0); yf.off.assign(P + 1, 0);
		std::vector<uint64_t> at(P, 0); // file offset of each sub-table's keys
		uint64_t o = 16;
		for (int s = 0; s < P; ++s) {
			uint32_t u[2] = {0, 0};
			if (o + 8 <= fsize && pread_all(fd, u, 8, o)) o += 8; else { u[0] = u[1] = 0; o = fsize; }
			const uint64_t got = std::min<uint64_t>(u[1], (fsize - o) / 8);
			yf.caps[s] = u[0];
			at[s] = o;
			yf.off[s + 1] = yf.off[s] + got;
			o = got == u[1] ? o + got * 8 : fsize; // a short array is the end of the file
		}
		const uint64_t n = yf.off[P];
		const double t_hdr = wall_now();
		if (!yf.keys.alloc(n)) return -1;
		const double t_alloc = wall_now();
		if (threads <= 0) threads = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
		if (n < (1u << 20)) threads = 1;
		std::atomic<int> next{0}, failed{0};
		auto work = [&]() {
			for (;;) {
				const int s0 = next.fetch_add(16);
				if (s0 >= P) break;
				for (int s = s0; s < std::min(P, s0 + 16); ++s) {
					const uint64_t m = yf.off[s + 1] - yf.off[s];
					if (m && !pread_all(fd, yf.keys.data() + yf.off[s], m * 8, at[s])) failed = 1;
				}
			}
		};
		std::vector<std::thread> th;
		for (int i = 1; i < threads; ++i) th.emplace_back(work);
		work();
		for (auto &x : th) x.join();
		if (timing_on()) fprintf(stderr, "[T::read_yak_file] %.2f GB: allocation %.3f s, key arrays %.3f s on %d threads\n", n * 8e-9, t_alloc - t_hdr, wall_now() - t_alloc, threads);
		if (failed) return -1;
	}
	if (mode != YAK_LOAD_ALL) { // counts -> flag bits, in place; TRIOBIN drops the k-mers below min_cnt (htab.c:448-469)
		const uint64_t cmask = YAK_MAX_COUNT;
		const int P = 1 << yf.pre;
		uint64_t w = 0, r = 0;
		for (int s = 0; s < P; ++s) {
			const uint64_t end = yf.off[s + 1];
			for (; r < end; ++r) {
				uint64_t key = yf.keys[r];
				if (mode == YAK_LOAD_TRIOBIN1 || mode == YAK_LOAD_TRIOBIN2) {
					const int cnt = (int)(key & cmask), shift = mode == YAK_LOAD_TRIOBIN1 ? 0 : 2;
					if (cnt >= mid_cnt) key = (key & ~cmask) | (uint64_t)(2 << shift);
					else if (cnt >= min_cnt) key = (key & ~cmask) | (uint64_t)(1 << shift);
					else continue;
				} else key = (key & ~cmask) | (uint64_t)(1 << (mode - YAK_LOAD_SEXCHR1));
				yf.keys[w++] = key;
			}
			yf.off[s + 1] = w;
		}
		yf.keys.shrink(w);
	}
	return 0;
}

} // namespace yakb
