// textcache.h - `yak count -b37 reads.fq.gz` reads its input twice (main.c:53-60), and inflating is the slow part of reading
// (zlib: ~0.3 GB/s of text per stream).  While the first pass reads a compressed file, the sequential reader can keep a copy
// of the inflated text in an anonymous memory file (memfd); the second pass of the same file then maps that copy and parses
// it with the parser pool like any plain file (csrc/fastx_par.h), skipping zlib altogether.
// Opt-in: YAKB_TEXT_CACHE_GB=<GB of host memory it may use> (0 / unset = off); one file at a time; the copy is dropped as soon
// as the second pass has opened it, or when another file takes its place.  A copy is only kept when the reader has seen the
// end of the input (not when the truncated-record rule stopped it early) and the text fit the budget.
#pragma once
#include <stdint.h>
#include <string>

namespace yakb {

// budget in bytes from YAKB_TEXT_CACHE_GB (0 = off)
uint64_t text_cache_budget();
// a new memfd for the text of `fn` (-1 if the cache is off or fn cannot be stat'ed); key receives fn's identity
int text_cache_begin(const char *fn, std::string *key);
// the reader wrote the whole text (bytes) to fd: keep it (takes ownership of fd) - or, with bytes == UINT64_MAX, drop it
void text_cache_end(const std::string &key, int fd, uint64_t bytes);
// pass 2: a path that maps the cached text of `fn` ("/proc/self/fd/N"), or "" when there is none / fn changed on disk.
// The caller opens it and then calls text_cache_release(), which closes the cache's descriptor (the mapping lives on).
std::string text_cache_lookup(const char *fn);
void text_cache_release();

} // namespace yakb
