// yakb_dev.cuh - device-side primitives of the k-mer count/lookup path (sm_100a).
//
// Semantics restated from the reference (cited per function); the data layout is ours:
//   * one count table for all 2^pre sub-tables: `slots[s*cap + i]`, cap = uniform per-sub-table
//     capacity (any size, fast-range indexed), slot = (v>>pre)<<10 | count like the reference's
//     stored key (htab.c:9-11), EMPTY = ~0.
//   * the khashl slot order the .yak format exposes is NOT kept here; it is rebuilt on demand
//     from the per-sub-table insertion journal (layout.cu).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#define YAKB_COUNTER_BITS 10
#define YAKB_MAX_COUNT 1023u
#define YAKB_EMPTY 0xFFFFFFFFFFFFFFFFull

namespace yakb {

// yak-priv.h:11-21
__host__ __device__ __forceinline__ uint64_t hash64(uint64_t key, uint64_t mask)
{
	key = (~key + (key << 21)) & mask;
	key = key ^ key >> 24;
	key = (key * 265) & mask;
	key = key ^ key >> 14;
	key = (key * 21) & mask;
	key = key ^ key >> 28;
	key = (key + (key << 31)) & mask;
	return key;
}

// yak-priv.h:23-33
__host__ __device__ __forceinline__ uint64_t hash64_64(uint64_t key)
{
	key = ~key + (key << 21);
	key = key ^ key >> 24;
	key = key * 265;
	key = key ^ key >> 14;
	key = key * 21;
	key = key ^ key >> 28;
	key = key + (key << 31);
	return key;
}

// yak-priv.h:41-68
__host__ __device__ __forceinline__ uint64_t hash64_inv(uint64_t key, uint64_t mask)
{
	uint64_t t;
	t = key - (key << 31);
	key = (key - (t << 31)) & mask;
	t = key ^ key >> 28;
	key = key ^ t >> 28;
	key = (key * 14933078535860113213ull) & mask;
	t = key ^ key >> 14;
	t = key ^ t >> 14;
	t = key ^ t >> 14;
	key = key ^ t >> 14;
	key = (key * 15244667743933553977ull) & mask;
	t = key ^ key >> 24;
	key = key ^ t >> 24;
	t = ~key;
	t = ~(key - (t << 21));
	t = ~(key - (t << 21));
	key = ~(key - (t << 21)) & mask;
	return key;
}

// khashl.h:98 + uint32 truncation (khashl.h:262): home slot of a stored key in a 2^bits table
__host__ __device__ __forceinline__ uint32_t kh_home(uint64_t stored, uint32_t bits)
{
	return ((uint32_t)(stored >> YAKB_COUNTER_BITS) * 2654435769u) >> (32 - bits);
}

// misc.c:4-21 as arithmetic (no table): A/a C/c G/g T/t U/u -> 0..3, bytes 0..3 -> themselves, else 4
__host__ __device__ __forceinline__ uint32_t nt4(uint32_t c)
{
	if (c < 4) return c;
	uint32_t u = c & 0xDFu; // fold case for letters
	if ((c | 0x20u) < 'a' || (c | 0x20u) > 'z') return 4;
	return u == 'A' ? 0 : u == 'C' ? 1 : u == 'G' ? 2 : (u == 'T' || u == 'U') ? 3 : 4;
}

// ---- our own table addressing (not the reference's) ----------------------------------------
// A sub-table region is an array of 32-byte buckets of 4 slots (one DRAM sector).  A key lives in
// the first bucket of its probe sequence (home bucket, then +1 ...) that had a free slot when it
// was inserted; slots are never freed, so a lookup may stop at the first bucket that still has an
// EMPTY slot.  Most lookups therefore cost exactly one sector.
#define YAKB_BUCKET 4

// EMPTY = ~0 is itself a possible stored value: the key whose id bits are all ones at count 1023 - which needs a 54-bit id,
// i.e. k >= 32 together with pre = 10.  That value is never written: the key's counter stops at 1022 in its slot
// (YAKB_ALMOST_EMPTY) and the last step is kept as one byte per sub-table in front of the table, YAKB_SAT_BYTES before
// slots[0] (pre = 10 means at most 1024 sub-tables).  Every test is on the VALUE, so configurations in which the value cannot
// occur pay one compare on the paths that write or read a counter and nothing else.
#define YAKB_SAT_BYTES 1024
#define YAKB_ALMOST_EMPTY 0xFFFFFFFFFFFFFFFEull
__device__ __forceinline__ uint8_t *sat_of(const uint64_t *slots) { return (uint8_t*)slots - YAKB_SAT_BYTES; }
// the 10-bit count of stored value v of sub-table s
__device__ __forceinline__ uint32_t slot_count(const uint64_t *slots, uint32_t s, uint64_t v)
{
	return (v == YAKB_ALMOST_EMPTY && sat_of(slots)[s]) ? YAKB_MAX_COUNT : (uint32_t)(v & YAKB_MAX_COUNT);
}

struct Bucket { uint64_t k[4]; };

__device__ __forceinline__ uint32_t tab_home(uint64_t x, uint32_t nbk) // home bucket of x among nbk buckets
{
	uint32_t h = (uint32_t)x * 2654435769u ^ (uint32_t)(x >> 32) * 0x85EBCA6Bu;
	return (uint32_t)(((uint64_t)h * nbk) >> 32);
}

// one 256-bit load per bucket (sm_100: LDG.E.256; two 128-bit loads were two L2 requests for the same sector)
__device__ __forceinline__ Bucket load_bucket(const uint64_t *p)
{
	Bucket r;
	asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(r.k[0]), "=l"(r.k[1]), "=l"(r.k[2]), "=l"(r.k[3]) : "l"(p));
	return r;
}

// slot (0..3) of the bucket holding x, or -1; *free_slot = first EMPTY slot or -1
__device__ __forceinline__ int bucket_match(const Bucket &b, uint64_t x, int *free_slot)
{
	int m = -1, f = -1;
#pragma unroll
	for (int i = 3; i >= 0; --i) {
		if (b.k[i] == YAKB_EMPTY) f = i;
		else if ((b.k[i] >> YAKB_COUNTER_BITS) == x) m = i;
	}
	*free_slot = f;
	return m;
}

// b.k[m] without a run-time register index
__device__ __forceinline__ uint64_t bucket_get(const Bucket &b, int m)
{
	return m == 0 ? b.k[0] : m == 1 ? b.k[1] : m == 2 ? b.k[2] : b.k[3];
}

// Find x (= v>>pre) in region `reg` (cap slots, cap % 4 == 0). Returns the slot index or -1.
__device__ __forceinline__ int64_t tab_find(const uint64_t *reg, uint32_t cap, uint64_t x)
{
	const uint32_t nbk = cap / YAKB_BUCKET;
	uint32_t bi = tab_home(x, nbk);
	for (uint32_t n = 0; n < nbk; ++n) {
		const Bucket b = load_bucket(reg + (uint64_t)bi * YAKB_BUCKET);
		int f, m = bucket_match(b, x, &f);
		if (m >= 0) return (int64_t)bi * YAKB_BUCKET + m;
		if (f >= 0) return -1;
		if (++bi == nbk) bi = 0;
	}
	return -1;
}

// Saturating ++ of the 10-bit counter with other threads possibly incrementing the same slot
// (htab.c:69-70 / 73-74: "if (count < 1023) ++count").
__device__ __forceinline__ void slot_inc(uint64_t *p, uint64_t cur, uint32_t by, uint8_t *sat_s)
{
	for (;;) {
		uint32_t c = (uint32_t)(cur & YAKB_MAX_COUNT);
		if (c >= YAKB_MAX_COUNT) return;
		uint32_t nc = c + by > YAKB_MAX_COUNT ? YAKB_MAX_COUNT : c + by;
		uint64_t want = (cur & ~(uint64_t)YAKB_MAX_COUNT) | nc;
		if (want == YAKB_EMPTY) { *sat_s = 1; want = YAKB_ALMOST_EMPTY; if (want == cur) return; } // see YAKB_SAT_BYTES
		uint64_t prev = atomicCAS((unsigned long long*)p, (unsigned long long)cur, (unsigned long long)want);
		if (prev == cur) return;
		cur = prev;
	}
}

// Insert-or-find `val` (stored form, key = val>>10).  Returns 1 if inserted, 0 if the key was
// already there (*slot_out = its slot).  Safe against concurrent inserts of OTHER keys.
__device__ __forceinline__ int tab_insert(uint64_t *reg, uint32_t cap, uint64_t val, uint64_t **slot_out, uint64_t *cur_out)
{
	const uint64_t x = val >> YAKB_COUNTER_BITS;
	const uint32_t nbk = cap / YAKB_BUCKET;
	uint32_t bi = tab_home(x, nbk);
	for (;;) {
		uint64_t *bp = reg + (uint64_t)bi * YAKB_BUCKET;
		const Bucket b = load_bucket(bp);
		int f, m = bucket_match(b, x, &f);
		if (m >= 0) { *slot_out = bp + m; *cur_out = bucket_get(b, m); return 0; }
		if (f >= 0) {
			uint64_t prev = atomicCAS((unsigned long long*)(bp + f), (unsigned long long)YAKB_EMPTY, (unsigned long long)val);
			if (prev == YAKB_EMPTY) { *slot_out = bp + f; *cur_out = val; return 1; }
			continue; // somebody took that slot: look at the same bucket again
		}
		if (++bi == nbk) bi = 0;
	}
}

// bbf.c:25-42 on one 64-byte block held as 16 words: test-and-set n_hashes bits, return #already set
__device__ __forceinline__ int bloom_block_insert(uint32_t *blk, uint32_t h1, uint32_t h2, int n_hashes)
{
	int cnt = 0;
	if ((h2 & 31) == 0) h2 = (h2 + 1) & 511;
	uint32_t z = h1;
	for (int i = 0; i < n_hashes; ++i, z = (z + h2) & 511) {
		uint32_t m = 1u << (z & 31);
		cnt += (blk[z >> 5] & m) != 0;
		blk[z >> 5] |= m;
	}
	return cnt;
}

} // namespace yakb
