// yakfile.h - the host side of reading a .yak table file (reference htab.c:419-472), see yakfile.cpp
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <vector>

namespace yakb {

// the keys of a table file: anonymous memory that is never zeroed by us and is backed by huge pages where the kernel gives
// them (a std::vector of 3 G keys spends seconds on first-touch page faults in one thread before a byte is read; here the
// reading threads touch their own parts)
struct KeyBuf {
	uint64_t *p = nullptr;
	size_t n = 0, bytes = 0;
	KeyBuf() {}
	KeyBuf(const KeyBuf&) = delete;
	KeyBuf &operator=(const KeyBuf&) = delete;
	~KeyBuf();
	bool alloc(size_t count);
	uint64_t *data() { return p; }
	size_t size() const { return n; }
	uint64_t &operator[](size_t i) { return p[i]; }
	void shrink(size_t count) { n = count; }
};

struct YakFile {
	uint32_t k = 0, pre = 0, counter_bits = 0;
	std::vector<uint32_t> caps;   // capacity of each sub-table as saved (khashl's n_buckets)
	std::vector<uint64_t> off;    // keys of sub-table s: keys[off[s] .. off[s+1])
	KeyBuf keys;
};

// Returns 0, -1 (cannot open / shorter than the header), -2 (magic), -3 (counter bits; yf.counter_bits holds the file's).
// mode / min_cnt / mid_cnt: yak_ch_restore_core's (YAK_LOAD_*: the flag modes map counts to class bits and TRIOBIN drops
// k-mers below min_cnt, htab.c:448-469).  threads 0 = one per core, at most 16.  header_only: stop after the checks
// (regular files; a pipe cannot be opened twice, so there the body comes along).
int read_yak_file(const char *fn, int mode, int min_cnt, int mid_cnt, YakFile &yf, int threads, bool header_only = false);

} // namespace yakb
