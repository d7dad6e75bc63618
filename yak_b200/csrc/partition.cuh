// partition.cuh - the radix partition in front of the table probe (reference: count.c:17-26 `ch_insert_buf`,
// the per-sub-table buffers a chunk's hashed k-mers are appended to before `yak_ch_insert_list` runs on each).
//
// Here a partition ("zone") is a group of neighbouring sub-tables whose slice of the count table is a few tens
// of megabytes, so that the probes of one zone stay inside L2 / a few DRAM rows (tools/zone_sweep.cu: random
// 32-byte read-modify-writes run at 12-15 G/s over the whole table and at 30-45 G/s when confined like this).
//
// part_scatter: ONE pass over the packed reads.  A CTA takes a tile of 256 words (8192 positions), one word per
// thread; the thread rolls its word and keeps the up to 32 hashes in registers.  Then, per tile:
//   count   shared-memory histogram of the tile's events per zone
//   reserve exclusive scan over the zones (place of every zone's run in the staging area) and one global
//           atomicAdd per non-empty (tile, zone) that reserves the run's place in the zone list
//   place   every event goes to its run in the staging area (shared memory, zone order)
//   copy    the staging area leaves the SM front to back: consecutive threads write consecutive entries, so a
//           run reaches its zone list as whole 32-byte sectors (8-byte hashes) instead of one store per event
// The order of events inside a zone list is arbitrary (counter updates commute; what must stay ordered - the
// pending events - is recovered from the position that travels with each hash).
// A zone list has a fixed capacity; what does not fit goes to a small spill list, and if that overflows too
// (one k-mer repeated millions of times) the host falls back to the unpartitioned probe for this chunk.
#pragma once
#include "engine.cuh"
#include "kernels.cuh"

namespace yakb {

#define YAKB_PT_POS 8192     // positions per tile

// shared memory of part_scatter for Z zones
static inline size_t part_smem_bytes(uint32_t Z) { return (size_t)YAKB_PT_POS * (8 + 2) + ((size_t)3 * Z + 1) * 4; }

template<bool LONGK, bool ARRAY>
__global__ void __launch_bounds__(256, 2) part_scatter(const uint64_t *__restrict__ w2, const uint32_t *__restrict__ wm, uint64_t nwords, int k,
                                                      const uint64_t *__restrict__ ev_in, uint64_t n_in, int only_s,
                                                      uint32_t Pmask, Own own, int zshift, uint32_t Z, uint32_t zcap, unsigned int *zfill,
                                                      uint64_t *__restrict__ zev, uint32_t *__restrict__ zpos,
                                                      uint64_t *__restrict__ spill_ev, uint32_t *__restrict__ spill_pos,
                                                      unsigned int *n_spill, uint32_t spill_cap, unsigned long long *stats)
{
	extern __shared__ __align__(16) unsigned char s_raw[];
	uint64_t *s_ev = (uint64_t*)s_raw;                       // [8192] hashes in zone order
	uint32_t *s_cnt = (uint32_t*)(s_ev + YAKB_PT_POS);       // [Z]   events per zone, then the place cursor
	uint32_t *s_toff = s_cnt + Z;                            // [Z+1] start of the zone's run in the staging area
	uint32_t *s_goff = s_toff + Z + 1;                       // [Z]   start of the run in the zone list
	uint16_t *s_pos = (uint16_t*)(s_goff + Z);               // [8192] position inside the tile
	__shared__ uint32_t s_wsum[8];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t zpt = (Z + 255) / 256;                    // zones per thread in the reserve step (Z <= 2048)
	const uint64_t ntiles = (nwords + 255) / 256;
	uint32_t my_ev = 0;
	for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		for (uint32_t i = threadIdx.x; i < Z; i += 256) s_cnt[i] = 0;
		__syncthreads();
		// ---- roll (or fetch) the thread's 32 events; local position of event r: packed = 32*tid + r, array = 256*r + tid
		uint64_t h[32];
		uint32_t vm = 0;
		if (ARRAY) {
			const uint64_t base = tile * YAKB_PT_POS + threadIdx.x;
#pragma unroll
			for (int r = 0; r < 32; ++r) {
				const uint64_t i = base + (uint64_t)r * 256;
				h[r] = i < n_in ? ev_in[i] : 0;
				if (i < n_in && (only_s < 0 || ((uint32_t)h[r] & Pmask) == (uint32_t)only_s) &&
				    ((uint32_t)(h[r] >> own.shift) & own.mask) == own.rank) vm |= 1u << r;
			}
		} else {
			const uint64_t W = tile * 256 + threadIdx.x;
			if (W < nwords) {
				Roller<LONGK> ro;
				ro.init(w2, wm, (int64_t)W, k);
#pragma unroll
				for (int r = 0; r < 32; ++r) {
					h[r] = 0;
					if (ro.step(r, h[r]) && ((uint32_t)(h[r] >> own.shift) & own.mask) == own.rank) vm |= 1u << r;
				}
			} else {
#pragma unroll
				for (int r = 0; r < 32; ++r) h[r] = 0;
			}
		}
		my_ev += __popc(vm);
		// ---- count
#pragma unroll
		for (int r = 0; r < 32; ++r)
			if (vm >> r & 1) atomicAdd(&s_cnt[((uint32_t)h[r] & Pmask) >> zshift], 1u);
		__syncthreads();
		// ---- reserve: thread t owns zones [t*zpt, (t+1)*zpt)
		uint32_t tot[8], sum = 0;
#pragma unroll
		for (uint32_t q = 0; q < 8; ++q) {
			const uint32_t z = threadIdx.x * zpt + q;
			tot[q] = (q < zpt && z < Z) ? s_cnt[z] : 0;
			sum += tot[q];
		}
		uint32_t incl = sum;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
		if (lane == 31) s_wsum[warp] = incl;
		__syncthreads();
		uint32_t run = incl - sum;
		for (int w = 0; w < warp; ++w) run += s_wsum[w];
#pragma unroll
		for (uint32_t q = 0; q < 8; ++q) {
			const uint32_t z = threadIdx.x * zpt + q;
			if (q < zpt && z < Z) {
				s_toff[z] = run;
				s_cnt[z] = run;
				s_goff[z] = tot[q] ? atomicAdd(&zfill[z], tot[q]) : 0;
				run += tot[q];
			}
		}
		if (threadIdx.x == 255) s_toff[Z] = run;  // thread 255 owns the last zones (or none): its running sum is the tile's total
		__syncthreads();
		// ---- place
#pragma unroll
		for (int r = 0; r < 32; ++r)
			if (vm >> r & 1) {
				const uint32_t i = atomicAdd(&s_cnt[((uint32_t)h[r] & Pmask) >> zshift], 1u);
				s_ev[i] = h[r];
				s_pos[i] = (uint16_t)(ARRAY ? r * 256 + threadIdx.x : threadIdx.x * 32 + r);
			}
		__syncthreads();
		// ---- copy out
		const uint32_t T = s_toff[Z];
		const uint64_t pos0 = tile * YAKB_PT_POS;
		for (uint32_t i = threadIdx.x; i < T; i += 256) {
			const uint64_t v = s_ev[i];
			const uint32_t z = ((uint32_t)v & Pmask) >> zshift, pos = (uint32_t)(pos0 + s_pos[i]);
			const uint32_t li = s_goff[z] + (i - s_toff[z]);
			if (li < zcap) { zev[(uint64_t)z * zcap + li] = v; zpos[(uint64_t)z * zcap + li] = pos; }
			else {
				const uint32_t q = atomicAdd(n_spill, 1u);
				if (q < spill_cap) { spill_ev[q] = v; spill_pos[q] = pos; }
			}
		}
		__syncthreads();
	}
#pragma unroll
	for (int d = 16; d; d >>= 1) my_ev += __shfl_xor_sync(0xffffffffu, my_ev, d);
	if (lane == 0 && my_ev) atomicAdd(&stats[0], (unsigned long long)my_ev);
}

} // namespace yakb
