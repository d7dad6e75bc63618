// extras.cu - kernels around the count path: event extraction to an array (multi-GPU routing),
// qv per-sequence statistics, single-key ops, the seeded synthetic generator.
#include "engine.cuh"
#include "yakb_dev.cuh"
#include "kernels.cuh"
#include "extras.cuh"
#include <stdio.h>
#include <algorithm>

namespace yakb {

static inline uint32_t cdiv(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }


// ---- multi-GPU routing: the hashed k-mers of a rank's reads, in file order, stably grouped by the rank
//      that owns their sub-table (owner = top lw bits of the sub-table index = hash bits [pre-lw, pre)).
//
//      route_tile_kernel: ONE roll.  A CTA takes a tile of 256 words (8192 positions), a thread one word; the thread keeps
//      its up to 32 hashes in registers, counts them per owner (four 16-bit counters to a u64), a block scan of those packed
//      counters gives every thread its stable place inside the tile's owner runs, and the tile leaves the SM sorted by
//      owner (file order inside an owner) as one coalesced copy.  cnt[owner][tile] = events of that owner in the tile.
//      An exclusive scan over cnt (owner-major) is the final place of every (owner, tile) run;
//      route_gather_kernel copies the runs there (coalesced reads, runs of ~8192/world events written contiguously).
#define YAKB_RT_POS 8192
#define YAKB_RT_MAXW 16

template<bool LONGK>
__global__ void __launch_bounds__(256, 2) route_tile_kernel(const uint64_t *__restrict__ w2, const uint32_t *__restrict__ wm, uint64_t nwords, int k,
                                                           int oshift, uint32_t omask, int world, uint64_t *__restrict__ tile_ev,
                                                           uint32_t *__restrict__ cnt, uint32_t ntiles)
{
	extern __shared__ __align__(16) uint64_t s_ev[];  // [8192]
	__shared__ unsigned long long s_w[8][4];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint64_t tile = blockIdx.x;
	const uint64_t W = tile * 256 + threadIdx.x;
	uint64_t h[32];
	uint32_t vm = 0;
	if (W < nwords) {
		Roller<LONGK> ro;
		ro.init(w2, wm, (int64_t)W, k);
#pragma unroll
		for (int r = 0; r < 32; ++r) { h[r] = 0; if (ro.step(r, h[r])) vm |= 1u << r; }
	} else {
#pragma unroll
		for (int r = 0; r < 32; ++r) h[r] = 0;
	}
	// this thread's events per owner, packed
	unsigned long long a[4] = {0, 0, 0, 0};
#pragma unroll
	for (int r = 0; r < 32; ++r)
		if (vm >> r & 1) {
			const uint32_t o = (uint32_t)(h[r] >> oshift) & omask;
			const unsigned long long one = 1ull << (16 * (o & 3));
#pragma unroll
			for (int g = 0; g < 4; ++g) if ((o >> 2) == (uint32_t)g) a[g] += one;
		}
	// block-wide exclusive scan of the packed counters (fields never carry: a tile holds at most 8192 events)
	unsigned long long inc[4];
#pragma unroll
	for (int g = 0; g < 4; ++g) {
		inc[g] = a[g];
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, inc[g], d); if (lane >= d) inc[g] += t; }
		if (lane == 31) s_w[warp][g] = inc[g];
	}
	__syncthreads();
	unsigned long long cur[4], tot[4];
#pragma unroll
	for (int g = 0; g < 4; ++g) {
		unsigned long long before = 0, all = 0;
#pragma unroll
		for (int w = 0; w < 8; ++w) { const unsigned long long t = s_w[w][g]; if (w < warp) before += t; all += t; }
		cur[g] = before + inc[g] - a[g];
		tot[g] = all;
	}
	// start of every owner's run inside the tile, added to the thread's place (packed again)
	uint32_t acc = 0;
#pragma unroll
	for (int o = 0; o < YAKB_RT_MAXW; ++o) {
		cur[o >> 2] += (unsigned long long)acc << (16 * (o & 3));
		const uint32_t c = (uint32_t)(tot[o >> 2] >> (16 * (o & 3))) & 0xFFFFu;
		if (threadIdx.x == 0 && o < world) cnt[(uint64_t)o * ntiles + tile] = c;
		acc += c;
	}
	// place, in position order
#pragma unroll
	for (int r = 0; r < 32; ++r)
		if (vm >> r & 1) {
			const uint32_t o = (uint32_t)(h[r] >> oshift) & omask;
			const int sh = 16 * (o & 3);
			uint32_t pos = 0;
#pragma unroll
			for (int g = 0; g < 4; ++g) if ((o >> 2) == (uint32_t)g) { pos = (uint32_t)(cur[g] >> sh) & 0xFFFFu; cur[g] += 1ull << sh; }
			s_ev[pos] = h[r];
		}
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < acc; i += 256) tile_ev[tile * YAKB_RT_POS + i] = s_ev[i];
}

// goff = exclusive scan of cnt (owner-major, world*ntiles + 1 entries)
__global__ void __launch_bounds__(256) route_gather_kernel(const uint64_t *__restrict__ tile_ev, const uint32_t *__restrict__ goff, uint32_t ntiles,
                                                           int world, uint64_t *__restrict__ out)
{
	__shared__ uint32_t s_toff[YAKB_RT_MAXW + 1], s_g[YAKB_RT_MAXW];
	const uint64_t tile = blockIdx.x;
	if (threadIdx.x == 0) {
		uint32_t acc = 0;
		for (int o = 0; o < world; ++o) {
			const uint64_t idx = (uint64_t)o * ntiles + tile;
			s_toff[o] = acc; s_g[o] = goff[idx];
			acc += goff[idx + 1] - goff[idx];
		}
		for (int o = world; o <= YAKB_RT_MAXW; ++o) s_toff[o] = acc;
	}
	__syncthreads();
	const uint32_t T = s_toff[world];
	for (uint32_t i = threadIdx.x; i < T; i += 256) {
		int o = 0;
#pragma unroll
		for (int q = 1; q < YAKB_RT_MAXW; ++q) if (q < world && s_toff[q] <= i) o = q;
		out[(uint64_t)s_g[o] + (i - s_toff[o])] = tile_ev[tile * YAKB_RT_POS + i];
	}
}

__global__ void route_counts_kernel(const uint32_t *__restrict__ goff, uint32_t ntiles, int world, uint64_t *__restrict__ counts)
{
	const int o = threadIdx.x;
	if (o < world) counts[o] = (uint64_t)goff[(uint64_t)(o + 1) * ntiles] - (uint64_t)goff[(uint64_t)o * ntiles];
}

// d_counts (device, world entries) receives the events per owner; nothing here waits for the device
int extract_events_async(const uint8_t *d_asc, uint64_t n, int k, int pre, int world, uint64_t *d_out, uint64_t *d_counts,
                         cudaStream_t stream, RouteScratch &sc)
{
	int lw = 0;
	while ((1 << lw) < world) ++lw;
	if ((1 << lw) != world || lw > pre || world > YAKB_RT_MAXW) {
		fprintf(stderr, "[yakb] ERROR: world size must be a power of two <= %d\n", YAKB_RT_MAXW);
		return -1;
	}
	YAKB_CUDA(cudaMemsetAsync(d_counts, 0, (size_t)world * 8, stream));
	if (n == 0) return 0;
	if (n >= 0xFFFFFF00ull) { fprintf(stderr, "[yakb] ERROR: chunk too large (positions are 32-bit)\n"); return -1; }
	const uint64_t nwords = (n + 31) / 32, ntiles = (nwords + 255) / 256;
	uint64_t *w2 = sc.b[0].as<uint64_t>(packed_words(nwords)) + YAKB_PADW;
	uint32_t *wm = sc.b[1].as<uint32_t>(packed_words(nwords)) + YAKB_PADW;
	uint64_t *tile_ev = sc.b[2].as<uint64_t>(ntiles * YAKB_RT_POS);
	uint32_t *cnt = sc.b[3].as<uint32_t>((uint64_t)world * ntiles + 1);
	ProfScope ps("extract_route", stream);
	pack_ascii_kernel<<<cdiv(packed_npad(nwords) + YAKB_PADW, 256), 256, 0, stream>>>(d_asc, n, w2, wm, nwords, packed_npad(nwords));
	YAKB_CUDA(cudaMemsetAsync(cnt + (uint64_t)world * ntiles, 0, 4, stream));
	const size_t sm = (size_t)YAKB_RT_POS * 8;
	if (k >= 32) {
		YAKB_CUDA(cudaFuncSetAttribute(route_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
		route_tile_kernel<true><<<(uint32_t)ntiles, 256, sm, stream>>>(w2, wm, nwords, k, pre - lw, (uint32_t)world - 1, world, tile_ev, cnt, (uint32_t)ntiles);
	} else {
		YAKB_CUDA(cudaFuncSetAttribute(route_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
		route_tile_kernel<false><<<(uint32_t)ntiles, 256, sm, stream>>>(w2, wm, nwords, k, pre - lw, (uint32_t)world - 1, world, tile_ev, cnt, (uint32_t)ntiles);
	}
	exclusive_scan_u32(cnt, cnt, (uint64_t)world * ntiles + 1, stream, sc.rs);
	route_gather_kernel<<<(uint32_t)ntiles, 256, 0, stream>>>(tile_ev, cnt, (uint32_t)ntiles, world, d_out);
	route_counts_kernel<<<1, 32, 0, stream>>>(cnt, (uint32_t)ntiles, world, d_counts);
	YAKB_CUDA(cudaGetLastError());
	Engine::note_launch(4);
	Prof::units("extract_route", n);
	return 0;
}

int extract_events(const uint8_t *d_asc, uint64_t n, int k, int pre, int world, uint64_t *d_out, uint64_t *counts,
                   cudaStream_t stream, RouteScratch &sc)
{
	for (int r = 0; r < world; ++r) counts[r] = 0;
	uint64_t *d_counts = (uint64_t*)sc.b[4].need((size_t)std::max(world, 1) * 8);
	const int rc = extract_events_async(d_asc, n, k, pre, world, d_out, d_counts, stream, sc);
	if (rc != 0 || n == 0) return rc;
	YAKB_CUDA(cudaMemcpyAsync(counts, d_counts, (size_t)world * 8, cudaMemcpyDeviceToHost, stream));
	YAKB_CUDA(cudaStreamSynchronize(stream));
	Prof::resolve();
	return 0;
}

// ---- qv.c:44-85 per sequence: tot / non0 and the min_frac gate.  One warp per sequence.
//      seq_off[s] = offset of sequence s in the position array; its last position is a separator.
// Both kernels walk POSITIONS, not sequences (a warp per sequence leaves 50 warps busy on an assembly of 50 contigs):
// a warp takes 1024 consecutive positions, every lane tracks the sequence its position lies in, and lanes of the
// same sequence combine their flags with votes before one atomic per (warp step, sequence).
__device__ __forceinline__ uint64_t seq_of(const uint64_t *__restrict__ seq_off, uint64_t n_seq, uint64_t p)
{
	uint64_t lo = 0, hi = n_seq; // sequence s with seq_off[s] <= p < seq_off[s+1]
	while (hi - lo > 1) { const uint64_t mid = (lo + hi) >> 1; if (seq_off[mid] <= p) lo = mid; else hi = mid; }
	return lo;
}

__global__ void __launch_bounds__(256) qv_pos_stats_kernel(const int16_t *__restrict__ cnt, const uint64_t *__restrict__ seq_off, uint64_t n_seq,
                                                           uint64_t n, int min_len, int32_t *tot, int32_t *non0)
{
	const int lane = threadIdx.x & 31;
	const uint64_t base = ((blockIdx.x * 256ull + threadIdx.x) >> 5) * 1024;
	if (base >= n) return;
	uint64_t s = seq_of(seq_off, n_seq, min(base + lane, n - 1));
	for (int it = 0; it < 32; ++it) {
		const uint64_t p = base + it * 32 + lane;
		const bool in = p < n;
		if (in) while (p >= seq_off[s + 1]) ++s;
		int c = -1;
		if (in && p + 1 < seq_off[s + 1] && (int64_t)(seq_off[s + 1] - 1 - seq_off[s]) >= min_len) c = cnt[p]; // not the separator; qv.c:44
		const uint32_t peers = __match_any_sync(0xffffffffu, in ? s : ~0ull);
		const uint32_t tb = __ballot_sync(0xffffffffu, c >= 0) & peers, zb = __ballot_sync(0xffffffffu, c > 0) & peers;
		if (in && lane == __ffs(peers) - 1) {
			if (tb) atomicAdd(&tot[s], __popc(tb));
			if (zb) atomicAdd(&non0[s], __popc(zb));
		}
	}
}

__global__ void qv_pass_kernel(const uint64_t *__restrict__ seq_off, uint64_t n_seq, int min_len, double min_frac,
                               const int32_t *__restrict__ tot, const int32_t *__restrict__ non0, uint8_t *__restrict__ pass)
{
	const uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (s >= n_seq) return;
	const int64_t len = (int64_t)(seq_off[s + 1] - 1 - seq_off[s]);
	pass[s] = len >= min_len && !((double)non0[s] < (double)tot[s] * min_frac); // qv.c:83
}

__global__ void __launch_bounds__(256) qv_pos_hist_kernel(const int16_t *__restrict__ cnt, const uint64_t *__restrict__ seq_off, uint64_t n_seq,
                                                          uint64_t n, const uint8_t *__restrict__ pass, unsigned long long *hist)
{
	__shared__ uint32_t s_h[1024];
	for (int i = threadIdx.x; i < 1024; i += 256) s_h[i] = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31;
	for (uint64_t base = ((blockIdx.x * 256ull + threadIdx.x) >> 5) * 1024; base < n; base += ((gridDim.x * 256ull) >> 5) * 1024) {
		uint64_t s = seq_of(seq_off, n_seq, min(base + lane, n - 1));
		for (int it = 0; it < 32; ++it) {
			const uint64_t p = base + it * 32 + lane;
			if (p >= n) break;
			while (p >= seq_off[s + 1]) ++s;
			if (p + 1 < seq_off[s + 1] && pass[s]) { const int c = cnt[p]; if (c >= 0) atomicAdd(&s_h[c], 1u); }
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < 1024; i += 256) if (s_h[i]) atomicAdd(&hist[i], (unsigned long long)s_h[i]);
}

void qv_stats(const int16_t *d_cnt, const uint64_t *d_seq_off, uint64_t n_seq, uint64_t n, int min_len, double min_frac,
              int32_t *d_tot, int32_t *d_non0, uint8_t *d_pass, unsigned long long *d_hist, cudaStream_t stream)
{
	if (n_seq == 0 || n == 0) return;
	YAKB_CUDA(cudaMemsetAsync(d_tot, 0, n_seq * 4, stream));
	YAKB_CUDA(cudaMemsetAsync(d_non0, 0, n_seq * 4, stream));
	const uint64_t n_warps = (n + 1023) / 1024;
	qv_pos_stats_kernel<<<cdiv(n_warps * 32, 256), 256, 0, stream>>>(d_cnt, d_seq_off, n_seq, n, min_len, d_tot, d_non0);
	qv_pass_kernel<<<cdiv(n_seq, 256), 256, 0, stream>>>(d_seq_off, n_seq, min_len, min_frac, d_tot, d_non0, d_pass);
	qv_pos_hist_kernel<<<std::min<uint32_t>(cdiv(n_warps * 32, 256), 148 * 8), 256, 0, stream>>>(d_cnt, d_seq_off, n_seq, n, d_pass, d_hist);
	YAKB_CUDA(cudaGetLastError());
}

// ---- htab.c:80-91 for one key
__global__ void inc_one_kernel(uint64_t *slots, uint32_t cap, int pre, uint32_t Pmask, Own own, uint64_t v, int32_t *out)
{
	int32_t r = -1;
	if (cap && ((uint32_t)(v >> own.shift) & own.mask) == own.rank) {
		uint64_t *reg = slots + (uint64_t)((uint32_t)v & Pmask) * cap;
		int64_t q = tab_find(reg, cap, v >> pre);
		if (q >= 0) {
			uint64_t cur = reg[q];
			if (cur == YAKB_ALMOST_EMPTY) { sat_of(slots)[(uint32_t)v & Pmask & (YAKB_SAT_BYTES - 1)] = 1; r = YAKB_MAX_COUNT; } // YAKB_SAT_BYTES
			else {
				if ((cur & YAKB_MAX_COUNT) < YAKB_MAX_COUNT) reg[q] = ++cur;
				r = (int32_t)(cur & YAKB_MAX_COUNT);
			}
		}
	}
	*out = r;
}
int inc_one(Engine *e, uint64_t v)
{
	int32_t *d = (int32_t*)e->b_misc.need(16), h = -1;
	inc_one_kernel<<<1, 1, 0, e->stream>>>(e->slots, e->cap, e->pre, e->P - 1, e->own(), v, d);
	YAKB_CUDA(cudaMemcpyAsync(&h, d, 4, cudaMemcpyDeviceToHost, e->stream));
	YAKB_CUDA(cudaStreamSynchronize(e->stream));
	return h;
}

// ---- htab.c:219-235
__global__ void setcnt_kernel(uint64_t *slots, uint64_t total, uint32_t cap, uint32_t c)
{
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (i >= total) return;
	uint64_t v = slots[i];
	if (v == YAKB_EMPTY) return;
	uint64_t w = (v & ~(uint64_t)YAKB_MAX_COUNT) | c;
	if ((v | YAKB_MAX_COUNT) == YAKB_EMPTY) { // the key whose count 1023 is kept as a flag (YAKB_SAT_BYTES)
		sat_of(slots)[(uint32_t)(i / cap) & (YAKB_SAT_BYTES - 1)] = c == YAKB_MAX_COUNT;
		if (w == YAKB_EMPTY) w = YAKB_ALMOST_EMPTY;
	}
	slots[i] = w;
}
void setcnt(Engine *e, int c)
{
	const uint64_t total = (uint64_t)e->P * e->cap;
	if (total) setcnt_kernel<<<cdiv(total, 256), 256, 0, e->stream>>>(e->slots, total, e->cap, (uint32_t)c);
	YAKB_CUDA(cudaGetLastError());
	YAKB_CUDA(cudaStreamSynchronize(e->stream));
}

// ---- bbf.c:25-42 on a stand-alone device filter (yak_bf_insert of the public API)
__global__ void bf_insert_one_kernel(uint32_t *bits, int n_shift, int n_hashes, uint64_t hash, int *out)
{
	const int sh = n_shift - 9;
	uint32_t *blk = bits + (hash & ((1ull << sh) - 1)) * 16;
	*out = bloom_block_insert(blk, (uint32_t)(hash >> sh) & 511, (uint32_t)(hash >> n_shift) & 511, n_hashes);
}
int bf_insert_one(uint8_t *d_bits, int n_shift, int n_hashes, uint64_t hash)
{
	int *d = nullptr, h = 0;
	d = (int*)dev_alloc(4);
	bf_insert_one_kernel<<<1, 1>>>((uint32_t*)d_bits, n_shift, n_hashes, hash, d);
	YAKB_CUDA(cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost));
	dev_free(d);
	return h;
}

// ---- synthetic data (yak_b200/synth.py is the specification)
__host__ __device__ __forceinline__ uint64_t smix(uint64_t z)
{
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
#define SYNTH_GMUL 0xD1342543DE82EF95ull

__global__ void synth_genome_kernel(uint64_t seed_g, uint64_t G, uint64_t *g2, uint64_t nwords)
{
	uint64_t W = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (W >= nwords) return;
	uint64_t w = 0;
	for (int r = 0; r < 32; ++r) {
		uint64_t j = W * 32 + r;
		uint64_t c = j < G ? smix(seed_g * SYNTH_GMUL + j) >> 62 : 0;
		w |= c << (62 - 2 * r);
	}
	g2[W] = w;
}

// fmt 0: SEQ\n   fmt 1 (FASTA): >r\nSEQ\n   fmt 2 (FASTQ): @r\nSEQ\n+\nQUAL\n  (fixed record size)
__host__ __device__ inline uint64_t synth_rec_bytes(int L, int fmt) { return fmt == 0 ? L + 1 : fmt == 1 ? L + 4 : 2 * (uint64_t)L + 7; }

__global__ void synth_reads_kernel(const uint64_t *__restrict__ g2, uint64_t G, uint64_t seed_r, uint64_t first, uint64_t n_reads,
                                   int L, uint64_t thr, int n_pct, int fmt, uint8_t *__restrict__ asc)
{
	const uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	const uint64_t rec = synth_rec_bytes(L, fmt);
	if (idx >= n_reads * rec) return;
	const uint64_t i = idx / rec;
	int j = (int)(idx - i * rec);
	if (fmt) { // header "Xr\n"
		if (j < 3) { asc[idx] = j == 0 ? (fmt == 1 ? '>' : '@') : j == 1 ? 'r' : '\n'; return; }
		j -= 3;
	}
	if (j == L) { asc[idx] = '\n'; return; }
	if (j > L) { j -= L + 1; asc[idx] = j == 0 ? '+' : (j == 1 || j == L + 2) ? '\n' : 'I'; return; }
	const uint64_t a = smix(seed_r * SYNTH_GMUL + (first + i));
	const uint64_t start = smix(a + 1) % (G - L + 1);
	const uint64_t f = smix(a + 2);
	const bool rev = f & 1;
	const uint64_t gi = rev ? start + (L - 1 - j) : start + j;
	uint32_t b = (uint32_t)(g2[gi >> 5] >> (62 - 2 * (gi & 31))) & 3;
	if (rev) b = 3 - b;
	const uint64_t e = smix(a + 16 + (uint64_t)j);
	if ((e & 0xFFFFFF) < thr) b = (uint32_t)(e >> 24) & 3;
	uint8_t ch = "ACGT"[b];
	if (((f >> 8) % 100) < (uint64_t)n_pct && (int)((f >> 32) % (uint64_t)L) == j) ch = 'N';
	asc[idx] = ch;
}

void synth_genome(uint64_t seed_g, uint64_t G, uint64_t *d_g2, cudaStream_t stream)
{
	const uint64_t nwords = (G + 31) / 32;
	synth_genome_kernel<<<cdiv(nwords, 256), 256, 0, stream>>>(seed_g, G, d_g2, nwords);
	YAKB_CUDA(cudaGetLastError());
}
void synth_reads(const uint64_t *d_g2, uint64_t G, uint64_t seed_r, uint64_t first, uint64_t n_reads, int L, double err, int n_pct,
                 int fmt, uint8_t *d_asc, cudaStream_t stream)
{
	const uint64_t total = n_reads * synth_rec_bytes(L, fmt);
	if (total == 0) return;
	synth_reads_kernel<<<cdiv(total, 256), 256, 0, stream>>>(d_g2, G, seed_r, first, n_reads, L, (uint64_t)(err * 16777216.0), n_pct, fmt, d_asc);
	YAKB_CUDA(cudaGetLastError());
}

} // namespace yakb
