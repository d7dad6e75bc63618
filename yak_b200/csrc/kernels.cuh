// kernels.cuh - device code shared by engine.cu and extras.cu (packing, the rolling canonical
// k-mer, ordered compaction of flagged positions).
#pragma once
#include "yakb_dev.cuh"

namespace yakb {

// ---- ASCII -> packed.  Word W of w2 holds bases 32W..32W+31, first base in the top 2 bits;
//      wm[W] bit (31-r) set = base 32W+r is not A/C/G/T/U (separator, N, padding).
//      Both arrays carry YAKB_PADW words in front of word 0 and are padded behind to a whole
//      256-word tile; every padding word reads "32 invalid bases", so window code never needs a
//      bounds test and whole tiles can be bulk-copied.  The kernel writes words -PADW .. npad-1.
#define YAKB_PADW 4
static __global__ void __launch_bounds__(256) pack_ascii_kernel(const uint8_t *__restrict__ asc, uint64_t n,
                                                         uint64_t *__restrict__ w2, uint32_t *__restrict__ wm, uint64_t nwords, uint64_t npad)
{
	const int64_t W = (int64_t)(blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) - YAKB_PADW;
	if (W >= (int64_t)npad) return;
	if (W < 0 || W >= (int64_t)nwords) { w2[W] = 0; wm[W] = 0xFFFFFFFFu; return; }
	uint64_t base = (uint64_t)W * 32, w = 0;
	uint32_t m = 0;
	if (base + 32 <= n && ((uintptr_t)(asc + base) & 15) == 0) {
		const uint4 *q = (const uint4*)(asc + base);
		uint4 a = __ldg(q), b = __ldg(q + 1);
		uint32_t u[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
		for (int r = 0; r < 32; ++r) {
			uint32_t c = nt4((u[r >> 2] >> (8 * (r & 3))) & 255);
			w |= (uint64_t)(c & 3) << (62 - 2 * r);
			m |= (c >> 2) << (31 - r);
		}
	} else {
		for (int r = 0; r < 32; ++r) {
			uint32_t c = base + r < n ? nt4(asc[base + r]) : 4;
			w |= (uint64_t)(c & 3) << (62 - 2 * r);
			m |= (c >> 2) << (31 - r);
		}
	}
	w2[W] = w; wm[W] = m;
}
// words a packed buffer needs for nwords data words (front pad + whole tiles + slack for bulk copies)
static inline uint64_t packed_words(uint64_t nwords) { return YAKB_PADW + (nwords + 255) / 256 * 256 + 8; }
static inline uint64_t packed_npad(uint64_t nwords) { return (nwords + 255) / 256 * 256 + 8; }

// ---- TMA-style bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra DONE_%=;\n"
		"bra WAIT_%=;\n"
		"DONE_%=:\n"
		"}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- rolling canonical k-mer over the 32 positions of word W.  count.c:28-43 (k<32) /
//      count.c:45-60 (k>=32).  init() builds the state after the k-1 bases before the word;
//      step(r, h) consumes position r and returns true (with the hash) when a k-mer ends there.
template<bool LONGK>
struct Roller {
	uint64_t x0, x1, x2, x3, mask, cw;
	uint32_t cm;
	int l, k, shift;

	// w2 / wm may be global arrays (index = word number) or a staged tile (index local to the tile):
	// indices down to W-2 (w2) and W-2 (wm) are read, which the padding / the tile halo provide
	__device__ __forceinline__ void init(const uint64_t *w2, const uint32_t *wm, int64_t W, int k_)
	{
		k = k_;
		mask = LONGK ? (1ULL << k) - 1 : (1ULL << 2 * k) - 1;
		shift = LONGK ? k - 1 : 2 * (k - 1);
		x0 = x1 = x2 = x3 = 0; l = 0;
		if (!LONGK) {
			// the last k-1 (<= 30) bases of the previous word are its low 2(k-1) bits.  A window never
			// crosses an invalid base, so whatever sits under invalid positions is shifted out before
			// the next emission.
			{
				const int n = k - 1;
				const uint64_t pw = w2[W - 1];
				const uint32_t pm = n ? (wm[W - 1] & (uint32_t)((1ull << n) - 1)) : 0;
				x0 = n ? pw & ((1ull << 2 * n) - 1) : 0;
				// reverse complement of those n bases, newest at the top: reverse the 2-bit groups
				uint64_t c = ~x0 & (n ? (1ull << 2 * n) - 1 : 0), r = __brevll(c);
				r = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
				x1 = n ? (r >> (64 - 2 * n)) << 2 : 0;
				l = pm ? __ffs(pm) - 1 : n;
			}
		} else {
			for (int64_t p = W * 32 - (k - 1); p < W * 32; ++p) { // p may be negative: >>5 floors, &31 wraps
				uint32_t c = (uint32_t)(w2[p >> 5] >> (62 - 2 * (p & 31))) & 3;
				uint32_t inv = (wm[p >> 5] >> (31 - (p & 31))) & 1;
				x0 = (x0 << 1 | (c & 1)) & mask;
				x1 = (x1 << 1 | (c >> 1)) & mask;
				x2 = x2 >> 1 | (uint64_t)(1 - (c & 1)) << shift;
				x3 = x3 >> 1 | (uint64_t)(1 - (c >> 1)) << shift;
				l = inv ? 0 : l + 1;
			}
		}
		cw = w2[W]; cm = wm[W];
	}

	__device__ __forceinline__ bool step(int r, uint64_t &h)
	{
		const uint32_t c = (uint32_t)(cw >> (62 - 2 * r)) & 3, inv = (cm >> (31 - r)) & 1;
		if (LONGK) {
			x0 = (x0 << 1 | (c & 1)) & mask;
			x1 = (x1 << 1 | (c >> 1)) & mask;
			x2 = x2 >> 1 | (uint64_t)(1 - (c & 1)) << shift;
			x3 = x3 >> 1 | (uint64_t)(1 - (c >> 1)) << shift;
		} else {
			x0 = (x0 << 2 | c) & mask;
			x1 = x1 >> 2 | (uint64_t)(3 - c) << shift;
		}
		l = inv ? 0 : (l < 64 ? l + 1 : l);
		if (l < k) return false;
		if (LONGK) h = x1 < x3 ? hash64_64(x0) + hash64_64(x1) : hash64_64(x2) + hash64_64(x3); // yak-priv.h:35-39
		else h = hash64(x0 < x1 ? x0 : x1, mask);
		return true;
	}
};

template<bool LONGK, class F>
__device__ __forceinline__ void roll_word(const uint64_t *__restrict__ w2, const uint32_t *__restrict__ wm,
                                          uint64_t W, int k, F &&emit)
{
	Roller<LONGK> ro;
	ro.init(w2, wm, (int64_t)W, k);
#pragma unroll
	for (int r = 0; r < 32; ++r) { uint64_t h; if (ro.step(r, h)) emit(r, h); }
}

__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t *total)
{
	__shared__ uint32_t s_w[8];
	__shared__ uint32_t s_tot;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	uint32_t inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
	if (lane == 31) s_w[w] = inc;
	__syncthreads();
	if (threadIdx.x == 0) { uint32_t a = 0; for (int i = 0; i < 8; ++i) { uint32_t t = s_w[i]; s_w[i] = a; a += t; } s_tot = a; }
	__syncthreads();
	uint32_t r = s_w[w] + inc - v;
	if (total) *total = s_tot;
	__syncthreads();
	return r;
}

// ---- pending events in file order: pv[j] = hash, ppos[j] = position
template<bool LONGK>
__global__ void __launch_bounds__(256) compact_fused(const uint64_t *__restrict__ w2, const uint32_t *__restrict__ wm, uint64_t nwords, int k,
                                                     const uint32_t *__restrict__ flags, const uint32_t *__restrict__ tileoff,
                                                     uint64_t tile0, uint32_t off0, uint64_t *__restrict__ pv, uint32_t *__restrict__ ppos)
{
	// tiles [tile0, tile0 + gridDim.x) of the chunk; off0 = tileoff[tile0], so the list starts at index 0
	const uint64_t W = (tile0 + blockIdx.x) * 256ull + threadIdx.x;
	uint32_t f = W < nwords ? flags[W] : 0;
	uint32_t o = tileoff[tile0 + blockIdx.x] - off0 + block_excl_scan_256(__popc(f), nullptr);
	if (f) roll_word<LONGK>(w2, wm, W, k, [&](int r, uint64_t h) {
		if (f >> r & 1) { pv[o] = h; ppos[o] = (uint32_t)(W * 32 + r); ++o; }
	});
}


} // namespace yakb
