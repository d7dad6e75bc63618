// radix.cu - see radix.cuh.  Hand-written replacements for the library scan / radix sort / select
// the first version of the path used between its kernels.
#include "radix.cuh"
#include <stdio.h>
#include <atomic>

namespace yakb {

static inline uint32_t cdiv(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

static std::atomic<uint64_t> g_radix_launches{0};
void radix_note_launch(int n) { g_radix_launches += n; }
uint64_t radix_launches() { return g_radix_launches.load(); }

// ------------------------------------------------------------------ exclusive scan

// block of 256 threads x 4 items: out = exclusive prefix inside the block, sums[block] = block total
__global__ void __launch_bounds__(256) scan_block_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint64_t n,
                                                         uint32_t *__restrict__ sums)
{
	__shared__ uint32_t s_w[8];
	const uint64_t base = (blockIdx.x * 256ull + threadIdx.x) * 4;
	uint32_t a[4];
#pragma unroll
	for (int i = 0; i < 4; ++i) a[i] = base + i < n ? in[base + i] : 0;
	const uint32_t tsum = a[0] + a[1] + a[2] + a[3];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	uint32_t inc = tsum;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
	if (lane == 31) s_w[w] = inc;
	__syncthreads();
	if (threadIdx.x == 0) { uint32_t acc = 0; for (int i = 0; i < 8; ++i) { uint32_t t = s_w[i]; s_w[i] = acc; acc += t; } if (sums) sums[blockIdx.x] = acc; }
	__syncthreads();
	uint32_t run = s_w[w] + inc - tsum;
#pragma unroll
	for (int i = 0; i < 4; ++i) { if (base + i < n) out[base + i] = run; run += a[i]; }
}

__global__ void __launch_bounds__(256) scan_add_kernel(uint32_t *__restrict__ out, uint64_t n, const uint32_t *__restrict__ block_off)
{
	const uint64_t base = (blockIdx.x * 256ull + threadIdx.x) * 4;
	const uint32_t add = block_off[blockIdx.x];
#pragma unroll
	for (int i = 0; i < 4; ++i) if (base + i < n) out[base + i] += add;
}

static void scan_level(const uint32_t *in, uint32_t *out, uint64_t n, cudaStream_t st, RadixScratch &rs, int level)
{
	const uint64_t nblk = (n + 1023) / 1024;
	if (nblk <= 1) {
		scan_block_kernel<<<1, 256, 0, st>>>(in, out, n, nullptr);
		return;
	}
	if (level >= 4) throw CudaError("[yakb] scan: input too large");
	uint32_t *sums = rs.lvl[level].as<uint32_t>(nblk);
	scan_block_kernel<<<(uint32_t)nblk, 256, 0, st>>>(in, out, n, sums);
	scan_level(sums, sums, nblk, st, rs, level + 1);
	scan_add_kernel<<<(uint32_t)nblk, 256, 0, st>>>(out, n, sums);
}

void exclusive_scan_u32(const uint32_t *in, uint32_t *out, uint64_t n, cudaStream_t st, RadixScratch &rs)
{
	if (n == 0) return;
	scan_level(in, out, n, st, rs, 0);
	YAKB_CUDA(cudaGetLastError());
	radix_note_launch(3);
}

// ------------------------------------------------------------------ radix sort

#define RTILE 4096 // 256 threads x 16 records

// lanes of the warp holding the same digit as this lane (invalid lanes match nobody): built from
// one ballot per digit bit - MATCH.ANY runs on the slow ADU pipe (ncu: 91 % busy with it)
__device__ __forceinline__ uint32_t same_digit_lanes(uint32_t d, bool ok, int nbits)
{
	uint32_t peers = __ballot_sync(0xffffffffu, ok);
#pragma unroll
	for (int b = 0; b < 8; ++b)
		if (b < nbits) {
			const uint32_t bit = d >> b & 1, m = __ballot_sync(0xffffffffu, bit);
			peers &= bit ? m : ~m;
		}
	return ok ? peers : 0;
}

__global__ void __launch_bounds__(256) radix_hist_kernel(const uint64_t *__restrict__ keys, uint64_t n, int shift, uint32_t dmask,
                                                         uint32_t *__restrict__ hist, uint32_t ntiles)
{
	__shared__ uint32_t s_h[256];
	s_h[threadIdx.x] = 0;
	__syncthreads();
	const uint64_t base = blockIdx.x * (uint64_t)RTILE;
#pragma unroll 4
	for (int i = 0; i < 16; ++i) {
		const uint64_t e = base + i * 256 + threadIdx.x;
		if (e < n) atomicAdd(&s_h[(uint32_t)(keys[e] >> shift) & dmask], 1u);
	}
	__syncthreads();
	if (threadIdx.x <= dmask) hist[(uint64_t)threadIdx.x * ntiles + blockIdx.x] = s_h[threadIdx.x];
}

// hist holds the exclusive prefix over (digit-major, tile) = the first output index of each (digit, tile).
// Records are ranked stably inside the tile ((warp, round, lane) order), parked in shared memory at
// their tile-local sorted position, and written out run by run so that stores coalesce.
template<bool IOTA, bool HASVAL>
__global__ void __launch_bounds__(256) radix_scatter_kernel(const uint64_t *__restrict__ k_in, const uint32_t *__restrict__ v_in,
                                                            uint64_t *__restrict__ k_out, uint32_t *__restrict__ v_out, uint64_t n,
                                                            int shift, uint32_t dmask, const uint32_t *__restrict__ hist, uint32_t ntiles)
{
	extern __shared__ unsigned char s_raw[];
	uint64_t *s_key = (uint64_t*)s_raw;                       // RTILE keys
	uint32_t *s_val = (uint32_t*)(s_raw + RTILE * 8);         // RTILE payloads (when HASVAL)
	__shared__ uint32_t s_c[8][256];   // per warp: count, then tile-local start, then running position
	__shared__ uint32_t s_dstart[256]; // tile-local start of each digit's run
	__shared__ uint32_t s_gbase[256];  // global index of the first record of each digit's run
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int nbits = 32 - __clz(dmask);
	for (int i = threadIdx.x; i < 8 * 256; i += 256) (&s_c[0][0])[i] = 0;
	__syncthreads();
	const uint64_t tbase = blockIdx.x * (uint64_t)RTILE, wbase = tbase + (uint64_t)w * 512;
	const uint32_t tcount = (uint32_t)(n - tbase < RTILE ? n - tbase : RTILE);
	uint64_t key[16];
	uint32_t val[16];
	// pass A: this warp's count per digit, elements taken in (round, lane) order
#pragma unroll
	for (int r = 0; r < 16; ++r) {
		const uint64_t e = wbase + r * 32 + lane;
		const bool ok = e < n;
		key[r] = ok ? k_in[e] : 0;
		if (HASVAL) val[r] = IOTA ? (uint32_t)e : (ok ? v_in[e] : 0);
		const uint32_t d = (uint32_t)(key[r] >> shift) & dmask;
		const uint32_t peers = same_digit_lanes(d, ok, nbits);
		if (ok && lane == __ffs(peers) - 1) s_c[w][d] += __popc(peers);
		__syncwarp();
	}
	__syncthreads();
	// digit totals of the tile -> tile-local run starts (exclusive scan over digits, 256 threads)
	{
		uint32_t tot = 0;
		if (threadIdx.x <= dmask) for (int i = 0; i < 8; ++i) tot += s_c[i][threadIdx.x];
		uint32_t inc = tot;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
		__shared__ uint32_t s_ws[8];
		if (lane == 31) s_ws[w] = inc;
		__syncthreads();
		uint32_t woff = 0;
		for (int i = 0; i < w; ++i) woff += s_ws[i];
		const uint32_t start = woff + inc - tot;
		s_dstart[threadIdx.x] = start;
		if (threadIdx.x <= dmask) {
			s_gbase[threadIdx.x] = hist[(uint64_t)threadIdx.x * ntiles + blockIdx.x];
			uint32_t acc = start;
			for (int i = 0; i < 8; ++i) { const uint32_t t = s_c[i][threadIdx.x]; s_c[i][threadIdx.x] = acc; acc += t; }
		}
	}
	__syncthreads();
	// pass B: same order again, now with running tile-local positions; park in shared memory
#pragma unroll
	for (int r = 0; r < 16; ++r) {
		const uint64_t e = wbase + r * 32 + lane;
		const bool ok = e < n;
		const uint32_t d = (uint32_t)(key[r] >> shift) & dmask;
		const uint32_t peers = same_digit_lanes(d, ok, nbits);
		uint32_t pos = 0;
		if (ok) pos = s_c[w][d];
		__syncwarp();
		if (ok && lane == __ffs(peers) - 1) s_c[w][d] = pos + __popc(peers);
		__syncwarp();
		if (ok) {
			pos += __popc(peers & ((1u << lane) - 1));
			s_key[pos] = key[r];
			if (HASVAL) s_val[pos] = val[r];
		}
	}
	__syncthreads();
	// write-out: consecutive threads take consecutive sorted records; a digit's run is contiguous globally
	for (uint32_t j = threadIdx.x; j < tcount; j += 256) {
		const uint64_t kk = s_key[j];
		const uint32_t d = (uint32_t)(kk >> shift) & dmask;
		const uint32_t dst = s_gbase[d] + (j - s_dstart[d]);
		k_out[dst] = kk;
		if (HASVAL) v_out[dst] = s_val[j];
	}
}

__global__ void copy_pairs_kernel(const uint64_t *k_in, const uint32_t *v_in, uint64_t *k_out, uint32_t *v_out, uint64_t n)
{
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (i >= n) return;
	k_out[i] = k_in[i];
	if (v_out) v_out[i] = v_in ? v_in[i] : (uint32_t)i;
}

int radix_sort_pairs(const uint64_t *k_src, const uint32_t *v_src, uint64_t *k_a, uint32_t *v_a, uint64_t *k_b, uint32_t *v_b,
                     uint64_t n, int begin_bit, int end_bit, cudaStream_t st, RadixScratch &rs)
{
	if (n == 0) return 0;
	if (n >= 0xFFFFFFF0ull) throw CudaError("[yakb] radix sort: too many records");
	const bool hasval = v_a != nullptr;
	if (end_bit <= begin_bit) {
		copy_pairs_kernel<<<cdiv(n, 256), 256, 0, st>>>(k_src, v_src, k_a, v_a, n);
		return 0;
	}
	const uint32_t ntiles = cdiv(n, RTILE);
	const int nbits = end_bit - begin_bit, npass = (nbits + 7) / 8;
	const uint64_t *kin = k_src;
	const uint32_t *vin = v_src;
	int where = -1;
	for (int p = 0; p < npass; ++p) {
		// spread the bits evenly over the passes (e.g. 28 bits = 7+7+7+7)
		const int lo = begin_bit + (int)((int64_t)nbits * p / npass), hi = begin_bit + (int)((int64_t)nbits * (p + 1) / npass);
		const uint32_t dmask = (1u << (hi - lo)) - 1;
		uint32_t *hist = rs.hist.as<uint32_t>((uint64_t)(dmask + 1) * ntiles);
		radix_hist_kernel<<<ntiles, 256, 0, st>>>(kin, n, lo, dmask, hist, ntiles);
		exclusive_scan_u32(hist, hist, (uint64_t)(dmask + 1) * ntiles, st, rs);
		uint64_t *kout = where == 0 ? k_b : k_a;
		uint32_t *vout = where == 0 ? v_b : v_a;
		const bool iota = hasval && vin == nullptr;
		const size_t sm = hasval ? RTILE * 12 : RTILE * 8;
		if (!hasval) {
			cudaFuncSetAttribute(radix_scatter_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
			radix_scatter_kernel<false, false><<<ntiles, 256, sm, st>>>(kin, nullptr, kout, nullptr, n, lo, dmask, hist, ntiles);
		} else if (iota) {
			cudaFuncSetAttribute(radix_scatter_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
			radix_scatter_kernel<true, true><<<ntiles, 256, sm, st>>>(kin, nullptr, kout, vout, n, lo, dmask, hist, ntiles);
		} else {
			cudaFuncSetAttribute(radix_scatter_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
			radix_scatter_kernel<false, true><<<ntiles, 256, sm, st>>>(kin, vin, kout, vout, n, lo, dmask, hist, ntiles);
		}
		YAKB_CUDA(cudaGetLastError());
		radix_note_launch(2);
		where = where == 0 ? 1 : 0;
		kin = kout; vin = vout;
	}
	return where;
}

// ------------------------------------------------------------------ ordered compaction

__global__ void __launch_bounds__(256) flag_count_kernel(const uint8_t *__restrict__ flag, uint64_t n, uint32_t *__restrict__ tilecnt)
{
	__shared__ uint32_t s_c;
	if (threadIdx.x == 0) s_c = 0;
	__syncthreads();
	const uint64_t base = blockIdx.x * (uint64_t)RTILE + threadIdx.x * 16ull;
	uint32_t c = 0;
#pragma unroll
	for (int i = 0; i < 16; ++i) c += base + i < n && flag[base + i] != 0;
	if (c) atomicAdd(&s_c, c);
	__syncthreads();
	if (threadIdx.x == 0) tilecnt[blockIdx.x] = s_c;
}

__global__ void __launch_bounds__(256) flag_scatter_kernel(const uint64_t *__restrict__ in, const uint8_t *__restrict__ flag, uint64_t n,
                                                           const uint32_t *__restrict__ tileoff, uint64_t *__restrict__ out)
{
	__shared__ uint32_t s_w[8];
	const uint64_t base = blockIdx.x * (uint64_t)RTILE + threadIdx.x * 16ull;
	uint32_t m = 0;
#pragma unroll
	for (int i = 0; i < 16; ++i) if (base + i < n && flag[base + i] != 0) m |= 1u << i;
	const uint32_t c = __popc(m);
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	uint32_t inc = c;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
	if (lane == 31) s_w[w] = inc;
	__syncthreads();
	if (threadIdx.x == 0) { uint32_t acc = 0; for (int i = 0; i < 8; ++i) { uint32_t t = s_w[i]; s_w[i] = acc; acc += t; } }
	__syncthreads();
	uint32_t o = tileoff[blockIdx.x] + s_w[w] + inc - c;
	while (m) { const int i = __ffs(m) - 1; m &= m - 1; out[o++] = in[base + i]; }
}

void compact_flagged_u64(const uint64_t *in, const uint8_t *flag, uint64_t n, uint64_t *out, uint32_t *d_count,
                         cudaStream_t st, RadixScratch &rs)
{
	const uint32_t ntiles = cdiv(n, RTILE);
	uint32_t *tc = rs.hist.as<uint32_t>((uint64_t)ntiles + 1);
	YAKB_CUDA(cudaMemsetAsync(tc + ntiles, 0, 4, st));
	if (n) flag_count_kernel<<<ntiles, 256, 0, st>>>(flag, n, tc);
	exclusive_scan_u32(tc, tc, (uint64_t)ntiles + 1, st, rs);
	if (n) flag_scatter_kernel<<<ntiles, 256, 0, st>>>(in, flag, n, tc, out);
	YAKB_CUDA(cudaMemcpyAsync(d_count, tc + ntiles, 4, cudaMemcpyDeviceToDevice, st));
	YAKB_CUDA(cudaGetLastError());
	radix_note_launch(2);
}

} // namespace yakb
