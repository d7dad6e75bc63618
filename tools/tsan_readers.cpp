// tsan_readers.cpp - the host readers (sequential + read-ahead thread, BGZF pool, parser pool) under ThreadSanitizer / AddressSanitizer.
// Inputs t.fq, t.fq.gz (BGZF), t_mix.fq.gz (BGZF + plain gzip), t_bad.fq.gz (one byte flipped) in the directory given as argv[1];
// tests/test_bgzf_cpu.py writes them, compiles this file with -fsanitize=thread and runs it.
#include <string>
#include "fastx.h"
#include "fastx_par.h"
#include <stdio.h>
#include <vector>
using namespace yakb;
static size_t run(const char *fn, int th, size_t job, size_t cap) {
	FastxReader r; if (!r.open(fn, th, job)) return 0;
	std::vector<uint8_t> buf(cap); size_t tot = 0; int64_t ns = 0; bool done = false; size_t need = 0;
	while (!done) { size_t n = r.fill(buf.data(), cap, cap, 0, &ns, &done, &need); if (need) { cap = need + 8; buf.resize(cap); continue; } tot += n; }
	return tot;
}
static size_t runp(const char *fn, int th, size_t blk, size_t cap) {
	ParallelFastx r; if (!r.open(fn, blk, th)) return 0;
	std::vector<uint8_t> buf(cap); size_t tot = 0; int64_t ns = 0; bool done = false; size_t need = 0;
	while (!done) { size_t n = r.fill(buf.data(), cap, cap, 0, &ns, &done, &need); if (need) { cap = need + 8; buf.resize(cap); continue; } tot += n; }
	return tot;
}
int main(int argc, char **argv) {
	const std::string dir = argc > 1 ? argv[1] : ".";
	size_t want = run((dir + "/t.fq").c_str(), 0, 1 << 20, 1 << 20);
	for (int th = 1; th <= 6; ++th) for (size_t job : {(size_t)1, (size_t)20000, (size_t)1 << 20}) {
		size_t a = run((dir + "/t.fq.gz").c_str(), th, job, 100000), b = run((dir + "/t_mix.fq.gz").c_str(), th, job, 1 << 20);
		run((dir + "/t_bad.fq.gz").c_str(), th, job, 1 << 20);
		if (a != want || b != want) { printf("MISMATCH %d %zu: %zu %zu %zu\n", th, job, a, b, want); return 1; }
		{ FastxReader r; r.open((dir + "/t.fq.gz").c_str(), th, job); std::vector<uint8_t> buf(5000); int64_t ns = 0; bool d = false; size_t need = 0; r.fill(buf.data(), 5000, 5000, 0, &ns, &d, &need); } // closed while the pool is busy
	}
	for (int th = 1; th <= 6; ++th) if (runp((dir + "/t.fq").c_str(), th, 5000, 100000) != want) { printf("MISMATCH pool %d\n", th); return 1; }
	printf("ok %zu\n", want);
	return 0;
}
