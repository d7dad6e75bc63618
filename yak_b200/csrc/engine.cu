// engine.cu - the device-resident count table and the per-chunk hot path (sm_100a).
//
// Per chunk (pass 1, create_new=1; reference count.c:111-143 + htab.c:51-78):
//   pack_ascii      ASCII -> 2-bit words + invalid mask                     (misc.c:4-21)
//   stage 1         every k-mer event meets the table: a key already present is a put-event whatever
//                   the bloom says (all its bits are set) -> counter++; otherwise its position is flagged
//                   "pending".  Large tables and chunks: part_scatter (partition.cuh: roll canonical k-mers,
//                   yak_hash64, radix partition into per-zone lists; count.c:17-60) + zone_probe (the table
//                   probed zone by zone).  Small ones: k1_fused (roll + probe in one kernel) / k1_array.
//   compact_*       pending events, in file order, in ranges of at most 256 M
//   sort by group   group = bloom block (= low n_shift-9 bits of the hash) or a hash-bit bucket
//   group_insert    one thread walks a group in file order: exact sequential bloom semantics
//                   (bbf.c:25-42), first-put detection, insert + counter++    (htab.c:62-71)
//   journal         new keys ordered by (sub-table, first-put time) appended as a segment
// Pass 2 / lookups (create_new=0) are stage 1 alone.
// The scans, the stable radix sorts and the ordered compactions between these kernels are ours too
// (radix.cu); no library kernels run on this path.
#include "engine.cuh"
#include "yakb_dev.cuh"
#include "kernels.cuh"
#include "extras.cuh"
#include "partition.cuh"
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <string>
#include <atomic>
#include <mutex>
#include <time.h>
#include <stdlib.h>

namespace yakb {

void cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
	char buf[512];
	snprintf(buf, sizeof(buf), "[yakb] CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
	fprintf(stderr, "%s\n", buf);
	throw CudaError(buf);
}

// ---- cached device memory (see dbuf.cuh)
namespace {
struct DevCache {
	struct Key { int dev; size_t bytes; bool operator<(const Key &o) const { return dev != o.dev ? dev < o.dev : bytes < o.bytes; } };
	std::mutex mu;
	std::map<void*, Key> live;                       // device and size of every block handed out
	std::multimap<Key, void*> idle;                  // freed blocks by (device, size)
	size_t idle_bytes = 0;
	size_t limit() { static size_t v = 0; if (!v) { const char *e = getenv("YAKB_CACHE_GB"); v = (size_t)(e ? atof(e) : 48.0) << 30; if (!v) v = 1; } return v; }
	bool enabled() { static int v = -1; if (v < 0) v = getenv("YAKB_NO_POOL") ? 0 : 1; return v != 0; }
	void drop(std::multimap<Key, void*>::iterator it) // cudaFree works on any device's pointer
	{
		cudaFree(it->second);
		idle_bytes -= it->first.bytes;
		idle.erase(it);
	}
	void drop_all() { while (!idle.empty()) drop(idle.begin()); }
};
DevCache g_dc;
}
void *dev_alloc(size_t bytes)
{
	void *p = nullptr;
	bytes = (std::max<size_t>(bytes, 1) + 255) & ~(size_t)255;
	if (!g_dc.enabled()) { YAKB_CUDA(cudaMalloc(&p, bytes)); return p; }
	int dev = 0;
	cudaGetDevice(&dev);
	std::lock_guard<std::mutex> lk(g_dc.mu);
	auto it = g_dc.idle.lower_bound({dev, bytes}); // smallest idle block of this device that is large enough, if it is not wastefully large
	if (it != g_dc.idle.end() && it->first.dev == dev && it->first.bytes <= bytes + (bytes >> 2) + (1u << 20)) {
		p = it->second;
		g_dc.live[p] = it->first;
		g_dc.idle_bytes -= it->first.bytes;
		g_dc.idle.erase(it);
		return p;
	}
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	cudaError_t e = cudaMalloc(&p, bytes);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	Prof::host("host:cudaMalloc", (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6); // 1-30 ms per GiB depending on the box
	if (e == cudaErrorMemoryAllocation && !g_dc.idle.empty()) { // idle blocks of the wrong sizes are in the way
		(void)cudaGetLastError();
		cudaDeviceSynchronize();
		g_dc.drop_all();
		e = cudaMalloc(&p, bytes);
	}
	YAKB_CUDA(e);
	g_dc.live[p] = {dev, bytes};
	return p;
}
void dev_free(void *p)
{
	if (p == nullptr) return;
	if (!g_dc.enabled()) { cudaFree(p); return; }
	cudaDeviceSynchronize(); // like cudaFree: nothing on the device uses the block any more
	std::lock_guard<std::mutex> lk(g_dc.mu);
	auto it = g_dc.live.find(p);
	if (it == g_dc.live.end()) { cudaFree(p); return; }
	const DevCache::Key key = it->second;
	g_dc.live.erase(it);
	g_dc.idle.insert({key, p});
	g_dc.idle_bytes += key.bytes;
	while (g_dc.idle_bytes > g_dc.limit() && !g_dc.idle.empty()) { // keep the hoard bounded: the largest block goes first
		auto big = g_dc.idle.begin();
		for (auto i = g_dc.idle.begin(); i != g_dc.idle.end(); ++i) if (i->first.bytes > big->first.bytes) big = i;
		g_dc.drop(big);
	}
}
size_t dev_pool_idle()
{
	std::lock_guard<std::mutex> lk(g_dc.mu);
	return g_dc.idle_bytes;
}
void dev_trim()
{
	cudaDeviceSynchronize();
	std::lock_guard<std::mutex> lk(g_dc.mu);
	g_dc.drop_all();
}

void *DBuf::need(size_t bytes)
{
	if (bytes > cap) {
		if (p) dev_free(p);
		p = nullptr;
		size_t want = bytes + (bytes >> 3) + 256;
		p = dev_alloc(want);
		cap = want;
	}
	return p;
}
void DBuf::release() { if (p) dev_free(p); p = nullptr; cap = 0; }

static inline uint32_t cdiv(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

// number of OUR kernels launched (cub's are not counted); reported by bench.py as gpu_launches
static std::atomic<uint64_t> g_launches{0};
uint64_t Engine::launches() { return g_launches.load() + radix_launches(); }
void Engine::note_launch(int n) { g_launches += n; }

// ---- per-kernel timing (off unless enabled)
static bool g_prof_on = false;
struct ProfPair { const char *name; cudaEvent_t a, b; };
static std::vector<ProfPair> g_prof_open, g_prof_done;
static std::map<std::string, std::pair<double, uint64_t>> g_prof_acc;
static std::map<std::string, uint64_t> g_prof_units; // work items (events, pending events, positions) a kernel name processed
static std::mutex g_prof_mu; // engines of a multi-GPU table run on one host thread each
void Prof::enable(bool on) { g_prof_on = on; }
bool Prof::on() { return g_prof_on; }
void Prof::reset() { std::lock_guard<std::mutex> lk(g_prof_mu); g_prof_acc.clear(); g_prof_units.clear(); }
void Prof::units(const char *name, uint64_t n) { if (g_prof_on) { std::lock_guard<std::mutex> lk(g_prof_mu); g_prof_units[name] += n; } }
void Prof::host(const char *name, double ms)
{
	if (!g_prof_on) return;
	std::lock_guard<std::mutex> lk(g_prof_mu);
	auto &e = g_prof_acc[name]; e.first += ms; e.second += 1;
}
void Prof::begin(const char *name, cudaStream_t s)
{
	std::lock_guard<std::mutex> lk(g_prof_mu);
	ProfPair p; p.name = name;
	cudaEventCreate(&p.a); cudaEventCreate(&p.b);
	cudaEventRecord(p.a, s);
	g_prof_open.push_back(p);
}
void Prof::end(cudaStream_t s)
{
	std::lock_guard<std::mutex> lk(g_prof_mu);
	ProfPair p = g_prof_open.back(); g_prof_open.pop_back();
	cudaEventRecord(p.b, s);
	g_prof_done.push_back(p);
}
void Prof::resolve()
{
	std::lock_guard<std::mutex> lk(g_prof_mu);
	for (auto &p : g_prof_done) {
		float ms = 0;
		if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
			auto &e = g_prof_acc[p.name]; e.first += ms; e.second += 1;
		}
		cudaEventDestroy(p.a); cudaEventDestroy(p.b);
	}
	g_prof_done.clear();
}
std::string Prof::json()
{
	std::lock_guard<std::mutex> lk(g_prof_mu);
	std::string o = "{";
	char buf[256];
	bool first = true;
	for (auto &kv : g_prof_acc) {
		auto u = g_prof_units.find(kv.first);
		snprintf(buf, sizeof(buf), "%s\"%s\": [%.6f, %llu, %llu]", first ? "" : ", ", kv.first.c_str(), kv.second.first, (unsigned long long)kv.second.second,
		         (unsigned long long)(u == g_prof_units.end() ? 0 : u->second));
		o += buf; first = false;
	}
	return o + "}";
}

// ============================================================ kernels

// Probe one event per lane, all 32 lanes in lock step (explicit convergence: data-dependent
// probe lengths otherwise leave the warp serialised).  `first` is the already loaded home bucket.
// Returns 1 if the key is in the table (its counter bumped by one, saturating), 0 if absent.
__device__ __forceinline__ int probe_inc_warp(uint64_t *reg, uint32_t nbk, uint64_t x, uint32_t bi, Bucket b, bool valid, uint8_t *sat_s)
{
	bool done = !valid;
	int found = 0;
	while (__any_sync(0xffffffffu, !done)) {
		if (!done) {
			int f, m = bucket_match(b, x, &f);
			if (m >= 0) {
				const uint64_t c = bucket_get(b, m);
				if ((c & YAKB_MAX_COUNT) == YAKB_MAX_COUNT) { found = 1; done = true; }
				else if (c == YAKB_ALMOST_EMPTY) { *sat_s = 1; found = 1; done = true; } // c + 1 would be EMPTY (YAKB_SAT_BYTES)
				else {
					uint64_t *p = reg + (uint64_t)bi * YAKB_BUCKET + m;
					const uint64_t prev = atomicCAS((unsigned long long*)p, (unsigned long long)c, (unsigned long long)(c + 1));
					if (prev == c) { found = 1; done = true; }
					else b = load_bucket(reg + (uint64_t)bi * YAKB_BUCKET); // lost a race on this counter: look again
				}
			} else if (f >= 0) done = true;
			else {
				if (++bi == nbk) bi = 0;
				b = load_bucket(reg + (uint64_t)bi * YAKB_BUCKET);
			}
		}
	}
	return found;
}

// Probe up to four events of one thread at once (vm = which of v[0..3] are live; bi / bk = their home buckets, already
// loaded by the caller so that the loads overlap whatever it did in between).  Returns the mask of events found in the
// table (counter bumped by one, saturating).  All 32 lanes call it together.  Every round handles ALL unfinished events
// of the thread: match, issue the counter CAS of every match, then look at the results; events whose bucket was full
// without the key move on to the next bucket, events that lost a CAS race look at the same bucket again - the follow-up
// loads of a round are issued together, so a warp pays one extra round trip per additional bucket, not one per event.
__device__ __forceinline__ uint32_t probe_inc4(uint64_t *slots, uint32_t cap, uint32_t nbk, int pre, uint32_t Pmask,
                                               const uint64_t (&v)[4], uint32_t vm, uint32_t (&bi)[4], Bucket (&bk)[4])
{
	uint32_t todo = vm, hit = 0;
	for (;;) {
		uint64_t expect[4], prev[4];
		uint32_t cas = 0, adv = 0;
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			expect[j] = prev[j] = 0;
			if (todo >> j & 1) {
				int f, m = bucket_match(bk[j], v[j] >> pre, &f);
				if (m >= 0) {
					const uint64_t c = bucket_get(bk[j], m);
					if ((c & YAKB_MAX_COUNT) == YAKB_MAX_COUNT) { hit |= 1u << j; todo &= ~(1u << j); } // htab.c:69: stays at the cap
					else if (c == YAKB_ALMOST_EMPTY) { // c + 1 would be EMPTY: the last step of this one key is a flag (YAKB_SAT_BYTES)
						sat_of(slots)[(uint32_t)v[j] & Pmask & (YAKB_SAT_BYTES - 1)] = 1;
						hit |= 1u << j; todo &= ~(1u << j);
					} else {
						uint64_t *q = slots + (uint64_t)((uint32_t)v[j] & Pmask) * cap + (uint64_t)bi[j] * YAKB_BUCKET + m;
						expect[j] = c;
						prev[j] = atomicCAS((unsigned long long*)q, (unsigned long long)c, (unsigned long long)(c + 1));
						cas |= 1u << j;
					}
				} else if (f >= 0) todo &= ~(1u << j);  // a free slot and no key: absent
				else adv |= 1u << j;                    // full without the key: the next bucket
			}
		}
#pragma unroll
		for (int j = 0; j < 4; ++j)
			if ((cas >> j & 1) && prev[j] == expect[j]) { hit |= 1u << j; todo &= ~(1u << j); }
		if (!__any_sync(0xffffffffu, todo != 0)) break;
#pragma unroll
		for (int j = 0; j < 4; ++j)
			if (todo >> j & 1) {
				if (adv >> j & 1) { if (++bi[j] == nbk) bi[j] = 0; }
				bk[j] = load_bucket(slots + (uint64_t)((uint32_t)v[j] & Pmask) * cap + (uint64_t)bi[j] * YAKB_BUCKET);
			}
	}
	return hit;
}

// ---- K1, fused front end: extraction + table probe.  One thread per 32-position word, persistent
//      CTAs over 256-word tiles.  flags[W] bit r = position 32W+r is a pending event.
//      stats[0] += events, glob_lput[s] = max(pos+1) over found (= put) events of sub-table s.
template<bool LONGK>
__global__ void __launch_bounds__(256, 3) k1_fused(const uint64_t *__restrict__ w2, const uint32_t *__restrict__ wm, uint64_t nwords,
                                                int k, int pre, uint32_t Pmask, Own own, uint64_t *slots, uint32_t cap, int create_new,
                                                uint32_t *__restrict__ flags, uint32_t *__restrict__ tilecnt,
                                                uint32_t *glob_lput, int smem_lp, unsigned long long *stats)
{
	extern __shared__ uint32_t s_lp[];
	// the packed read batch is staged tile by tile in shared memory with bulk async copies
	// (cp.async.bulk -> UBLKCP) two tiles deep: 2 halo + 256 words of bases, 4 halo + 256 mask words
	__shared__ alignas(128) uint64_t s_w2[2][264];
	__shared__ alignas(128) uint32_t s_wm[2][264];
	__shared__ alignas(8) uint64_t s_full[2], s_empty[2];
	const uint32_t P = Pmask + 1;
	if (smem_lp) for (uint32_t i = threadIdx.x; i < P; i += 256) s_lp[i] = 0;
	if (threadIdx.x == 0) {
		mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1);
		mbar_init(&s_empty[0], 8); mbar_init(&s_empty[1], 8);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	const uint64_t ntiles = (nwords + 255) / 256;
	const uint32_t nbk = cap / YAKB_BUCKET;
	const int lane = threadIdx.x & 31;
	uint32_t my_ev = 0;
	auto stage_tile = [&](uint32_t it, uint64_t tile) { // thread 0 only
		const int sidx = it & 1;
		if (it >= 2) mbar_wait(&s_empty[sidx], ((it >> 1) - 1) & 1); // all 8 warps are done with the previous use
		mbar_expect_tx(&s_full[sidx], 258 * 8 + 260 * 4);
		bulk_g2s(&s_w2[sidx][0], w2 + (int64_t)tile * 256 - 2, 258 * 8, &s_full[sidx]);
		bulk_g2s(&s_wm[sidx][0], wm + (int64_t)tile * 256 - 4, 260 * 4, &s_full[sidx]);
	};
	if (threadIdx.x == 0 && blockIdx.x < ntiles) stage_tile(0, blockIdx.x);
	uint32_t it = 0;
	for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) { // no block barrier inside: warps run free
		if (threadIdx.x == 0 && tile + gridDim.x < ntiles) stage_tile(it + 1, tile + gridDim.x);
		__syncwarp();
		const int sidx = it & 1;
		mbar_wait(&s_full[sidx], (it >> 1) & 1);
		const uint64_t W = tile * 256 + threadIdx.x;
		const bool live = W < nwords; // dead lanes walk along with no events so the warp stays whole
		uint32_t pend = 0;
		Roller<LONGK> ro;
		if (live) ro.init(&s_w2[sidx][2], &s_wm[sidx][4], (int64_t)threadIdx.x, k);
		__syncwarp();
		if (lane == 0) mbar_arrive(&s_empty[sidx]); // the roller keeps what it needs in registers
#pragma unroll 1
		for (int b = 0; b < 32; b += 4) { // 4 positions at a time: 4 independent home-bucket loads in flight
			uint64_t v[4];
			Bucket bk[4];
			uint32_t bi[4], vm = 0;
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				bi[j] = 0; v[j] = 0;
				if (live && ro.step(b + j, v[j]) && ((uint32_t)(v[j] >> own.shift) & own.mask) == own.rank) {
					vm |= 1u << j;
					if (cap) {
						bi[j] = tab_home(v[j] >> pre, nbk);
						bk[j] = load_bucket(slots + (uint64_t)((uint32_t)v[j] & Pmask) * cap + (uint64_t)bi[j] * YAKB_BUCKET);
					}
				}
			}
			__syncwarp();
			my_ev += __popc(vm);
			if (cap == 0) { pend |= vm << b; continue; }
			const uint32_t hit = probe_inc4(slots, cap, nbk, pre, Pmask, v, vm, bi, bk);
			if (create_new) {
				pend |= (vm & ~hit) << b;
#pragma unroll
				for (int j = 0; j < 4; ++j)
					if (hit >> j & 1) {
						const uint32_t s = (uint32_t)v[j] & Pmask, t = (uint32_t)(W * 32 + b + j) + 1;
						if (smem_lp) atomicMax(&s_lp[s], t); else atomicMax(&glob_lput[s], t);
					}
			}
		}
		if (create_new) {
			if (live) flags[W] = pend;
			uint32_t c = __popc(pend);
#pragma unroll
			for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
			if (lane == 0 && c) atomicAdd(&tilecnt[tile], c); // tilecnt zeroed by the host before the launch
		}
	}
#pragma unroll
	for (int d = 16; d; d >>= 1) my_ev += __shfl_xor_sync(0xffffffffu, my_ev, d);
	if (lane == 0 && my_ev) atomicAdd(&stats[0], (unsigned long long)my_ev);
	__syncthreads();
	if (smem_lp && create_new)
		for (uint32_t i = threadIdx.x; i < P; i += 256) if (s_lp[i]) atomicMax(&glob_lput[i], s_lp[i]);
}

// ---- K1, partitioned front end for large tables.  Random 32-byte read-modify-writes spread over tens of
//      gigabytes run at ~12-15 G/s on this part; the same operations confined to tens of megabytes of the table
//      at a time run 2-4 times faster (DRAM row locality + L2 merging; tools/zone_sweep.cu).  So the events of a
//      chunk are first partitioned into per-zone lists (zone = a group of neighbouring sub-tables; part_scatter in
//      partition.cuh - the reference's ch_insert_buf, count.c:17-26), then the table is probed zone by zone.
//      Counter updates commute, and the pending flag / last-put bookkeeping are per position, so the order in
//      which events reach the table is free.
//
//      zone_probe: persistent CTAs take slices of YAKB_ZSLICE events from a global work counter, in zone order, so
//      at any moment the whole grid works on one or two zones.  Per event the same bucket probe + counter
//      CAS as k1_fused; a miss sets the position's bit in flags[] (pass 1).  n_list = Z zone lists of zcap
//      entries each (fill counts in zfill[]), or one list (the spill).
//      The kernel is bound by memory latency (ncu: long-scoreboard stalls, 20 % issue slots used), so it runs one group
//      of 1024 events ahead of itself: while a group is probed, the next group's events are already in registers and
//      their home buckets on the way into L2 (prefetch.global.L2), across slice boundaries too - thread 0 fetches the
//      CTA's next work item a slice early.
//      zsub = sub-tables per zone when that is at most 8 (0 otherwise, and for the spill list): the last-put times of a
//      slice's hits are then reduced per sub-table inside the warp - one shared-memory atomic per (warp, slice, sub-table)
//      instead of one per hit, which in a zone list all fall on the same few counters and serialise.
#define YAKB_ZSLICE 4096
#define YAKB_ZGROUPS (YAKB_ZSLICE / 1024)
struct ZGroup { uint64_t v[4]; uint32_t pos[4]; uint32_t vm; };

__device__ __forceinline__ void zone_fetch(ZGroup &q, uint64_t item, uint32_t g, uint64_t n_items, uint32_t spz, uint32_t zcap,
                                           const uint64_t *__restrict__ zev, const uint32_t *__restrict__ zpos, const unsigned int *__restrict__ zfill,
                                           int pre, uint32_t Pmask, const uint64_t *slots, uint32_t cap, uint32_t nbk, bool prefetch)
{
	q.vm = 0;
	if (item >= n_items) return;
	const uint32_t z = (uint32_t)(item / spz), sl = (uint32_t)(item % spz);
	const uint32_t nz = min(zfill[z], zcap), start = sl * YAKB_ZSLICE + g * 1024;
	const uint64_t *ev = zev + (uint64_t)z * zcap;
	const uint32_t *ep = zpos + (uint64_t)z * zcap;
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		const uint32_t i = start + j * 256 + threadIdx.x;
		q.v[j] = 0; q.pos[j] = 0;
		if (i < nz) {
			q.v[j] = ev[i]; q.pos[j] = ep[i];
			q.vm |= 1u << j;
			if (prefetch) { // off by default: ncu shows 92 instead of 69 B of DRAM reads per event with it (whole lines come in) and no gain
				const uint64_t *bp = slots + (uint64_t)((uint32_t)q.v[j] & Pmask) * cap + (uint64_t)tab_home(q.v[j] >> pre, nbk) * YAKB_BUCKET;
				asm volatile("prefetch.global.L2 [%0];" :: "l"(bp));
			}
		}
	}
}

template<int MINB>
__global__ void __launch_bounds__(256, MINB) zone_probe(const uint64_t *__restrict__ zev, const uint32_t *__restrict__ zpos, uint32_t n_list, uint32_t zcap,
                                                    const unsigned int *__restrict__ zfill, int pre, uint32_t Pmask,
                                                    uint64_t *slots, uint32_t cap, int create_new, uint32_t *flags,
                                                    uint32_t *glob_lput, int smem_lp, unsigned int *work, uint32_t zsub, bool prefetch)
{
	extern __shared__ uint32_t s_lp[];
	__shared__ uint32_t s_item[2];
	const uint32_t P = Pmask + 1, nbk = cap / YAKB_BUCKET;
	if (smem_lp) for (uint32_t i = threadIdx.x; i < P; i += 256) s_lp[i] = 0;
	// slices per list; a single list (the spill) is cut by its actual fill, which is usually zero
	const uint32_t spz = ((n_list == 1 ? min(zfill[0], zcap) : zcap) + YAKB_ZSLICE - 1) / YAKB_ZSLICE;
	const uint64_t n_items = (uint64_t)n_list * spz;
	if (threadIdx.x == 0) s_item[0] = atomicAdd(work, 1u);
	__syncthreads();
	uint64_t item = s_item[0];
	ZGroup nx;
	zone_fetch(nx, item, 0, n_items, spz, zcap, zev, zpos, zfill, pre, Pmask, slots, cap, nbk, prefetch);
	for (uint32_t it = 0; item < n_items; ++it) {
		if (threadIdx.x == 0) s_item[(it + 1) & 1] = atomicAdd(work, 1u); // needed one slice from now
		const uint32_t s0 = (uint32_t)(item / spz) * zsub; // first sub-table of the slice's zone
		uint32_t tm[8];
#pragma unroll
		for (int i = 0; i < 8; ++i) tm[i] = 0;
		uint64_t next_item = item;
#pragma unroll 1
		for (uint32_t g = 0; g < YAKB_ZGROUPS; ++g) {
			ZGroup cur = nx;
			if (g + 1 < YAKB_ZGROUPS) zone_fetch(nx, item, g + 1, n_items, spz, zcap, zev, zpos, zfill, pre, Pmask, slots, cap, nbk, prefetch);
			else {
				__syncthreads(); // thread 0's fetch of the next item, issued a slice ago, is visible; one barrier per slice
				next_item = s_item[(it + 1) & 1];
				zone_fetch(nx, next_item, 0, n_items, spz, zcap, zev, zpos, zfill, pre, Pmask, slots, cap, nbk, prefetch);
			}
			Bucket bk[4];
			uint32_t bi[4];
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				bi[j] = 0;
				if (cur.vm >> j & 1) {
					bi[j] = tab_home(cur.v[j] >> pre, nbk);
					bk[j] = load_bucket(slots + (uint64_t)((uint32_t)cur.v[j] & Pmask) * cap + (uint64_t)bi[j] * YAKB_BUCKET);
				}
			}
			const uint32_t hit = probe_inc4(slots, cap, nbk, pre, Pmask, cur.v, cur.vm, bi, bk);
			if (create_new) {
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					if (!(cur.vm >> j & 1)) continue;
					if (hit >> j & 1) {
						const uint32_t s = (uint32_t)cur.v[j] & Pmask, t = cur.pos[j] + 1;
						if (zsub) {
							const uint32_t sl = s - s0;
#pragma unroll
							for (int i = 0; i < 8; ++i) if (sl == (uint32_t)i) tm[i] = max(tm[i], t);
						} else if (smem_lp) atomicMax(&s_lp[s], t);
						else atomicMax(&glob_lput[s], t);
					} else atomicOr(&flags[cur.pos[j] >> 5], 1u << (cur.pos[j] & 31));
				}
			}
		}
		if (create_new && zsub) {
#pragma unroll
			for (int i = 0; i < 8; ++i)
				if ((uint32_t)i < zsub) { // warp-uniform
					const uint32_t t = __reduce_max_sync(0xffffffffu, tm[i]);
					if ((threadIdx.x & 31) == 0 && t) { if (smem_lp) atomicMax(&s_lp[s0 + i], t); else atomicMax(&glob_lput[s0 + i], t); }
				}
		}
		item = next_item;
	}
	__syncthreads();
	if (smem_lp && create_new)
		for (uint32_t i = threadIdx.x; i < P; i += 256) if (s_lp[i]) atomicMax(&glob_lput[i], s_lp[i]);
}

// pending events per 256-word tile from the flag words (what k1_fused accumulates as it goes)
__global__ void __launch_bounds__(256) flag_tilecnt_kernel(const uint32_t *__restrict__ flags, uint64_t nwords, uint32_t *__restrict__ tilecnt)
{
	const uint64_t W = blockIdx.x * 256ull + threadIdx.x;
	uint32_t c = W < nwords ? __popc(flags[W]) : 0;
	uint32_t tot;
	block_excl_scan_256(c, &tot);
	if (threadIdx.x == 0) tilecnt[blockIdx.x] = tot;
}

// ---- K1, array front end (events already hashed: yak_ch_insert_list, multi-GPU receive side).
//      word W = events 32W..32W+31, lane = bit.
__global__ void __launch_bounds__(256) k1_array(const uint64_t *__restrict__ ev, uint64_t n, int pre, uint32_t Pmask, Own own,
                                                uint64_t *slots, uint32_t cap, int create_new, int only_s,
                                                uint32_t *__restrict__ flags, uint32_t *__restrict__ tilecnt,
                                                uint32_t *glob_lput, int smem_lp, unsigned long long *stats)
{
	extern __shared__ uint32_t s_lp[];
	__shared__ uint32_t s_cnt;
	__shared__ unsigned long long s_ev;
	const uint32_t P = Pmask + 1;
	if (smem_lp) for (uint32_t i = threadIdx.x; i < P; i += 256) s_lp[i] = 0;
	if (threadIdx.x == 0) s_ev = 0;
	__syncthreads();
	const uint64_t nwords = (n + 31) / 32, ntiles = (nwords + 255) / 256;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t my_ev = 0;
	for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		if (threadIdx.x == 0) s_cnt = 0;
		__syncthreads();
		for (int it = 0; it < 32; ++it) {
			const uint64_t W = tile * 256 + warp * 32 + it;
			if (W >= nwords) break; // warp-uniform
			const uint64_t i = W * 32 + lane;
			int valid = i < n, found = 0;
			uint64_t v = valid ? ev[i] : 0;
			if (valid && only_s >= 0 && ((uint32_t)v & Pmask) != (uint32_t)only_s) valid = 0;
			if (valid && ((uint32_t)(v >> own.shift) & own.mask) != own.rank) valid = 0; // not this shard's sub-table
			if (valid) ++my_ev;
			if (cap) {
				const uint32_t nbk = cap / YAKB_BUCKET, s = (uint32_t)v & Pmask;
				uint64_t *reg = slots + (uint64_t)s * cap;
				uint32_t bi = 0;
				Bucket bk;
				if (valid) { bi = tab_home(v >> pre, nbk); bk = load_bucket(reg + (uint64_t)bi * YAKB_BUCKET); }
				found = probe_inc_warp(reg, nbk, v >> pre, bi, bk, valid, sat_of(slots) + (s & (YAKB_SAT_BYTES - 1)));
			}
			if (valid) {
				if (found && create_new) {
					uint32_t t = (uint32_t)i + 1;
					if (smem_lp) atomicMax(&s_lp[(uint32_t)v & Pmask], t);
					else atomicMax(&glob_lput[(uint32_t)v & Pmask], t);
				}
			}
			if (create_new) {
				uint32_t pend = __ballot_sync(0xffffffffu, valid && !found);
				if (lane == 0) { flags[W] = pend; if (pend) atomicAdd(&s_cnt, __popc(pend)); }
			}
		}
		__syncthreads();
		if (threadIdx.x == 0 && create_new) tilecnt[tile] = s_cnt;
	}
	if (my_ev) atomicAdd(&s_ev, (unsigned long long)my_ev);
	__syncthreads();
	if (threadIdx.x == 0 && s_ev) atomicAdd(&stats[0], s_ev);
	if (smem_lp && create_new)
		for (uint32_t i = threadIdx.x; i < P; i += 256) if (s_lp[i]) atomicMax(&glob_lput[i], s_lp[i]);
}

__global__ void __launch_bounds__(256) compact_array(const uint64_t *__restrict__ ev, uint64_t nwords,
                                                     const uint32_t *__restrict__ flags, const uint32_t *__restrict__ tileoff,
                                                     uint64_t tile0, uint32_t off0, uint64_t *__restrict__ pv, uint32_t *__restrict__ ppos)
{
	const uint64_t W = (tile0 + blockIdx.x) * 256ull + threadIdx.x;
	uint32_t f = W < nwords ? flags[W] : 0;
	uint32_t o = tileoff[tile0 + blockIdx.x] - off0 + block_excl_scan_256(__popc(f), nullptr);
	while (f) {
		int r = __ffs(f) - 1;
		f &= f - 1;
		pv[o] = ev[W * 32 + r]; ppos[o] = (uint32_t)(W * 32 + r); ++o;
	}
}

// pending events per sub-table (upper bound of the new keys a chunk can add): shared-memory bins
__global__ void __launch_bounds__(256) pend_hist_kernel(const uint64_t *__restrict__ pv, uint64_t n, uint32_t Pmask, uint32_t *pend, int smem_ok)
{
	extern __shared__ uint32_t s_bins[];
	const uint32_t P = Pmask + 1;
	if (smem_ok) { for (uint32_t i = threadIdx.x; i < P; i += 256) s_bins[i] = 0; __syncthreads(); }
	for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += gridDim.x * 256ull) {
		const uint32_t s = (uint32_t)pv[i] & Pmask;
		if (smem_ok) atomicAdd(&s_bins[s], 1u); else atomicAdd(&pend[s], 1u);
	}
	if (smem_ok) {
		__syncthreads();
		for (uint32_t i = threadIdx.x; i < P; i += 256) if (s_bins[i]) atomicAdd(&pend[i], s_bins[i]);
	}
}

__global__ void max_need_kernel(const uint32_t *nkeys, const uint32_t *pend, uint32_t P, uint32_t *out)
{
	__shared__ uint32_t s_m;
	if (threadIdx.x == 0) s_m = 0;
	__syncthreads();
	uint32_t m = 0;
	for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) m = max(m, nkeys[i] + (pend ? pend[i] : 0));
	atomicMax(&s_m, m);
	__syncthreads();
	if (threadIdx.x == 0) *out = s_m;
}

// ---- the ordered part: one thread per group, events of a group walked in file order.
//      pflag: 2 bits per pending event j (file-order index), bit0 = put-event, bit1 = this put inserted a new key, as a
//      bitmap of u32 words set with atomicOr: 64 MB for a range of 256 M events, which stays in L2 - one byte per event
//      written at a random j was a partial-sector write to DRAM each (ncu: 238 B of DRAM traffic per pending event).
//      Groups are short (1-4 events) and only their first event starts a walk: a thread per event leaves 6 of 32 lanes busy
//      (ncu).  So a warp takes 128 consecutive events, finds the group heads among them (each lane looks at 4 events), lists
//      them in shared memory, and its lanes then take the heads 32 at a time - every lane walks a group.
__global__ void __launch_bounds__(256) group_insert(const uint64_t *__restrict__ sv, const uint32_t *__restrict__ sj, uint64_t n,
                                                    int G, int pre, uint32_t Pmask, int lw, uint64_t *slots, uint32_t cap,
                                                    uint32_t *bloom32, int nb, int sub_shift, int n_hash,
                                                    uint32_t *__restrict__ pflag)
{
	__shared__ uint8_t s_heads[8][128];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint64_t gmask = G >= 64 ? ~0ull : (1ull << G) - 1;
	const uint64_t ntile = (n + 127) / 128;
	for (uint64_t tile = blockIdx.x * 8ull + warp; tile < ntile; tile += gridDim.x * 8ull) {
	const uint64_t base = tile * 128 + (uint64_t)lane * 4;
	uint32_t hm = 0;
	{
		uint64_t prevk = base > 0 && base - 1 < n ? sv[base - 1] & gmask : 0;
#pragma unroll
		for (int q = 0; q < 4; ++q)
			if (base + q < n) {
				const uint64_t kq = sv[base + q] & gmask;
				if (base + q == 0 || kq != prevk) hm |= 1u << q;
				prevk = kq;
			}
	}
	uint32_t incl = __popc(hm);
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
	const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
	{
		uint32_t o = incl - __popc(hm);
#pragma unroll
		for (int q = 0; q < 4; ++q) if (hm >> q & 1) s_heads[warp][o++] = (uint8_t)(lane * 4 + q);
	}
	__syncwarp();
	for (uint32_t hr = lane; hr < total; hr += 32) {
	const uint64_t i = tile * 128 + s_heads[warp][hr];
	const uint64_t gk = sv[i] & gmask;
	// block address: the group key with the (constant) owner bits of a shard squeezed out
	const uint64_t baddr = lw ? ((gk >> pre) << (pre - lw)) | (gk & Pmask) : gk;
	for (uint64_t e = i; e < n; ++e) {
		const uint64_t v = sv[e];
		if ((v & gmask) != gk) break;
		const uint32_t j = sj[e];
		const uint32_t s = (uint32_t)v & Pmask;
		const uint64_t x = v >> pre;
		int put = 1;
		if (bloom32) { // bbf.c:25-42 on the 64-byte block this group owns
			uint32_t *blk = bloom32 + baddr * 16;
			uint32_t h1 = (uint32_t)(x >> nb) & 511, h2 = (uint32_t)(x >> sub_shift) & 511;
			if ((h2 & 31) == 0) h2 = (h2 + 1) & 511;
			int c = 0;
			if (n_hash <= 8) {
				// the n_hash bit positions are distinct (h2 is not a multiple of 32), so the tests are
				// independent: fetch the words first, then decide, then write back the changed words
				uint32_t w[8], m[8], old[8];
#pragma unroll
				for (int t = 0; t < 8; ++t)
					if (t < n_hash) {
						const uint32_t z = (h1 + t * h2) & 511;
						w[t] = z >> 5; m[t] = 1u << (z & 31);
						old[t] = __ldcg(&blk[w[t]]);
					}
#pragma unroll
				for (int t = 0; t < 8; ++t) if (t < n_hash) c += (old[t] & m[t]) != 0;
				if (c != n_hash) {
#pragma unroll
					for (int t = 0; t < 8; ++t)
						if (t < n_hash) {
							uint32_t nw = old[t];
#pragma unroll
							for (int u = 0; u < 8; ++u) if (u < n_hash && w[u] == w[t]) nw |= m[u];
							if (nw != old[t]) __stcg(&blk[w[t]], nw);
						}
				}
			} else {
				uint32_t z = h1;
				for (int t = 0; t < n_hash; ++t, z = (z + h2) & 511) {
					const uint32_t ww = __ldcg(&blk[z >> 5]), mm = 1u << (z & 31);
					c += (ww & mm) != 0;
					if (!(ww & mm)) __stcg(&blk[z >> 5], ww | mm);
				}
			}
			put = c == n_hash;
		}
		uint8_t flag = 0;
		if (put) { // htab.c:66-70
			uint64_t *slot, cur;
			if (tab_insert(slots + (uint64_t)s * cap, cap, x << YAKB_COUNTER_BITS | 1, &slot, &cur)) flag = 3;
			else { slot_inc(slot, cur, 1, sat_of(slots) + (s & (YAKB_SAT_BYTES - 1))); flag = 1; }
		}
		if (flag) atomicOr(&pflag[j >> 4], (uint32_t)flag << ((j & 15) * 2));
	}
	}
	__syncwarp(); // the next tile's head list overwrites this one
	}
}

// ---- per sub-table max position (+1) of pending put-events / new-key puts of this chunk
__global__ void __launch_bounds__(256) post_pending(const uint64_t *__restrict__ pv, const uint32_t *__restrict__ ppos,
                                                    const uint32_t *__restrict__ pflag, uint8_t *__restrict__ isnew, uint64_t n, uint32_t Pmask,
                                                    uint32_t *glob_lput, uint32_t *glob_lnew, int smem_ok, unsigned long long *stats)
{
	extern __shared__ uint32_t s_arr[];
	const uint32_t P = Pmask + 1;
	uint32_t *s_lp = s_arr, *s_ln = s_arr + P;
	if (smem_ok) for (uint32_t i = threadIdx.x; i < 2 * P; i += 256) s_arr[i] = 0;
	__syncthreads();
	uint32_t nput = 0;
	for (uint64_t j = blockIdx.x * 256ull + threadIdx.x; j < n; j += gridDim.x * 256ull) {
		const uint32_t f = pflag[j >> 4] >> ((j & 15) * 2) & 3;
		isnew[j] = f >> 1 & 1;
		if (!(f & 1)) continue;
		++nput;
		uint32_t s = (uint32_t)pv[j] & Pmask, t = ppos[j] + 1;
		if (smem_ok) { atomicMax(&s_lp[s], t); if (f & 2) atomicMax(&s_ln[s], t); }
		else { atomicMax(&glob_lput[s], t); if (f & 2) atomicMax(&glob_lnew[s], t); }
	}
	if (nput) atomicAdd(&stats[1], (unsigned long long)nput);
	__syncthreads();
	if (smem_ok)
		for (uint32_t i = threadIdx.x; i < P; i += 256) {
			if (s_lp[i]) atomicMax(&glob_lput[i], s_lp[i]);
			if (s_ln[i]) atomicMax(&glob_lnew[i], s_ln[i]);
		}
}

__global__ void merge_times_kernel(uint32_t P, uint32_t seq, const uint32_t *lput, const uint32_t *lnew, uint64_t *last_put, uint64_t *last_new)
{
	uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= P) return;
	if (lput[s]) last_put[s] = (uint64_t)(seq + 1) << 32 | lput[s];
	if (lnew[s]) last_new[s] = (uint64_t)(seq + 1) << 32 | lnew[s];
}

// off[s] = first index whose sub-table >= s (sorted by the low `pre` bits), s in [0, P]
__global__ void seg_offsets_kernel(const uint64_t *__restrict__ sorted, uint64_t n, uint32_t Pmask, uint64_t *off, uint32_t *nkeys)
{
	uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s > Pmask + 1) return;
	uint64_t lo = 0, hi = n;
	while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (((uint32_t)sorted[mid] & Pmask) < s) lo = mid + 1; else hi = mid; }
	off[s] = lo;
}
__global__ void seg_addkeys_kernel(const uint64_t *off, uint32_t P, uint32_t *nkeys)
{
	uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s < P) nkeys[s] += (uint32_t)(off[s + 1] - off[s]);
}
__global__ void seg_keys_kernel(const uint64_t *__restrict__ sorted, uint64_t n, int pre, uint64_t *__restrict__ keys)
{
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (i < n) keys[i] = sorted[i] >> pre << YAKB_COUNTER_BITS;
}

// ---- re-place every used slot into a table with a different per-sub-table capacity
__global__ void rehash_kernel(const uint64_t *__restrict__ old_slots, uint32_t old_cap, uint64_t total, uint64_t *new_slots, uint32_t new_cap)
{
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (i >= total) return;
	uint64_t val = old_slots[i];
	if (val == YAKB_EMPTY) return;
	uint64_t *slot, cur;
	tab_insert(new_slots + (i / old_cap) * new_cap, new_cap, val, &slot, &cur);
}

// stored keys (with counts) of sub-tables given by off[] into the table (restore / shrink rebuild)
__global__ void bulk_insert_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ off, uint32_t P, uint64_t n,
                                   uint64_t *slots, uint32_t cap)
{
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t lo = 0, hi = P; // sub-table s with off[s] <= i < off[s+1]
	while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (off[mid] <= i) lo = mid; else hi = mid; }
	uint64_t *slot, cur, val = keys[i];
	if (val == YAKB_EMPTY) { val = YAKB_ALMOST_EMPTY; sat_of(slots)[lo & (YAKB_SAT_BYTES - 1)] = 1; } // YAKB_SAT_BYTES
	tab_insert(slots + (uint64_t)lo * cap, cap, val, &slot, &cur); // a duplicate key in the file: first claim wins
}

// journal entries are keys without counts; unlike a table slot, a journal entry may legitimately be all ones (YAKB_SAT_BYTES)
__global__ void strip_counts_kernel(uint64_t *keys, uint64_t n)
{
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (i < n) keys[i] &= ~(uint64_t)YAKB_MAX_COUNT;
}

__global__ void clear_kernel(uint64_t *slots, uint64_t total)
{
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (i < total) { uint64_t v = slots[i]; if (v != YAKB_EMPTY) slots[i] = v & ~(uint64_t)YAKB_MAX_COUNT; }
}

__global__ void __launch_bounds__(256) hist_kernel(const uint64_t *__restrict__ slots, uint64_t total, uint32_t cap, unsigned long long *hist)
{
	__shared__ uint32_t s_h[1024];
	for (int i = threadIdx.x; i < 1024; i += 256) s_h[i] = 0;
	__syncthreads();
	for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < total; i += gridDim.x * 256ull) {
		uint64_t v = slots[i];
		if (v != YAKB_EMPTY) atomicAdd(&s_h[v == YAKB_ALMOST_EMPTY ? slot_count(slots, (uint32_t)(i / cap), v) : (uint32_t)(v & YAKB_MAX_COUNT)], 1u);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < 1024; i += 256) if (s_h[i]) atomicAdd(&hist[i], (unsigned long long)s_h[i]);
}

// htab.c:93-100
__global__ void get_batch_kernel(const uint64_t *__restrict__ xs, uint64_t n, int pre, uint32_t Pmask, Own own,
                                 const uint64_t *__restrict__ slots, uint32_t cap, int32_t *__restrict__ out)
{
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint64_t v = xs[i];
	int32_t r = -1;
	if (cap && ((uint32_t)(v >> own.shift) & own.mask) == own.rank) {
		int64_t q = tab_find(slots + (uint64_t)((uint32_t)v & Pmask) * cap, cap, v >> pre);
		if (q >= 0) r = (int32_t)slot_count(slots, (uint32_t)v & Pmask, slots[(uint64_t)((uint32_t)v & Pmask) * cap + q]);
	}
	out[i] = r;
}

// qv.c:48-66: per position the count of its k-mer (absent -> 0), or -1 where no k-mer ends
// raw = 1 (the other scanners: triobin.c:76, trioeval.c:75, chkerr.c:55, sexchr.c:61): yak_ch_get's own value,
// -1 for an absent k-mer, and -2 where no k-mer ends
template<bool LONGK>
__global__ void __launch_bounds__(256) qv_scan_kernel(const uint64_t *__restrict__ w2, const uint32_t *__restrict__ wm, uint64_t nwords, uint64_t n,
                                                      int k, int pre, uint32_t Pmask, Own own, const uint64_t *__restrict__ slots, uint32_t cap,
                                                      int16_t *__restrict__ out, int raw)
{
	const uint64_t W = blockIdx.x * 256ull + threadIdx.x;
	if (W >= nwords) return;
	int16_t res[32];
#pragma unroll
	for (int r = 0; r < 32; ++r) res[r] = raw ? -2 : -1;
	roll_word<LONGK>(w2, wm, W, k, [&](int r, uint64_t v) {
		int16_t c = raw ? -1 : 0;
		if (cap && ((uint32_t)(v >> own.shift) & own.mask) == own.rank) {
			const uint64_t *reg = slots + (uint64_t)((uint32_t)v & Pmask) * cap;
			int64_t q = tab_find(reg, cap, v >> pre);
			if (q >= 0) c = (int16_t)slot_count(slots, (uint32_t)v & Pmask, reg[q]);
		}
		res[r] = c;
	});
#pragma unroll
	for (int r = 0; r < 32; ++r) if (W * 32 + r < n) out[W * 32 + r] = res[r];
}

// ---- rebuild the reference's khashl layout (khashl.h:137-221) from the insertion journal.
//      The kick-out rehash is a sequential procedure per table, so one thread replays one
//      sub-table; small tables are replayed in shared memory (one warp per sub-table, lane 0 walks
//      the journal the other lanes stage), large ones in global scratch.
struct KhReplay {
	uint64_t *K;
	uint32_t *used, *occ;
	uint32_t bits = 0, n = 0, count = 0;
	static __device__ __forceinline__ uint32_t fw(uint32_t nb) { return nb < 32 ? 1u : nb >> 5; }
	__device__ __forceinline__ bool is_used(uint32_t i) const { return used[i >> 5] >> (i & 31) & 1; }
	__device__ void resize(uint32_t request) // khashl.h:152-195
	{
		uint32_t lg = 0, q = request;
		while ((q >>= 1) != 0) ++lg;
		if (request & (request - 1)) ++lg;
		const uint32_t new_bits = lg > 2 ? lg : 2, new_n = 1u << new_bits, new_mask = new_n - 1;
		if (count > (new_n >> 1) + (new_n >> 2)) return;
		for (uint32_t w = 0; w < fw(new_n); ++w) occ[w] = 0;
		for (uint32_t j = 0; j != n; ++j) {
			if (!is_used(j)) continue;
			uint64_t key = K[j];
			used[j >> 5] &= ~(1u << (j & 31));
			for (;;) {
				uint32_t i = kh_home(key, new_bits);
				while (occ[i >> 5] >> (i & 31) & 1) i = (i + 1) & new_mask;
				occ[i >> 5] |= 1u << (i & 31);
				if (i < n && is_used(i)) {
					uint64_t ev = K[i]; K[i] = key; key = ev;
					used[i >> 5] &= ~(1u << (i & 31));
				} else { K[i] = key; break; }
			}
		}
		uint32_t *tmp = used; used = occ; occ = tmp;
		bits = new_bits; n = new_n;
	}
	__device__ __forceinline__ void check() { if (count >= (n >> 1) + (n >> 2)) resize(n + 1); } // khashl.h:202-205
	__device__ void entry(uint64_t key) // one journal entry: a put (khashl.h:197-221) or an operation (Engine::OP_*)
	{
		const uint32_t op = (uint32_t)key & YAKB_MAX_COUNT;
		if (op) {
			const uint32_t r = (uint32_t)(key >> YAKB_COUNTER_BITS);
			if (op == 1) resize(r);
			else if (op == 2) check();
			else if (op == 3) { if ((uint64_t)count * 3 < n) resize(count * 3); }
			else if (op == 4) { if (r > n) resize(r); }
			return;
		}
		check();
		const uint32_t mask = n - 1;
		uint32_t i = kh_home(key, bits), start = i;
		while (is_used(i) && (K[i] >> YAKB_COUNTER_BITS) != (key >> YAKB_COUNTER_BITS)) {
			i = (i + 1) & mask;
			if (i == start) break;
		}
		if (!is_used(i)) { K[i] = key; used[i >> 5] |= 1u << (i & 31); ++count; }
	}
};

// the journal of one sub-table = its runs in every segment, in segment order
struct SegView { const uint64_t *const *keys; const uint64_t *const *off; int n; };

// global-scratch variant: one thread per listed sub-table.  The sub-table's output region (reg) is
// also its working key array; the slot-ordered keys end up compacted at its front.
__global__ void build_layout_kernel(const int *__restrict__ list, int nlist, SegView sv, int s_first,
                                    const uint8_t *__restrict__ pre_flag, const uint32_t *__restrict__ pre_val,
                                    const uint8_t *__restrict__ trailing,
                                    uint64_t *out_all, const uint64_t *__restrict__ ooff,
                                    uint32_t *bm_all, const uint64_t *__restrict__ boff,
                                    uint32_t *out_cap, uint32_t *out_size)
{
	const int li = blockIdx.x * blockDim.x + threadIdx.x;
	if (li >= nlist) return;
	const int t = list[li];
	KhReplay R;
	R.K = out_all + ooff[t];
	const uint64_t bwords = (boff[t + 1] - boff[t]) / 2;
	R.used = bm_all + boff[t]; R.occ = R.used + bwords;
	R.used[0] = 0;
	if (pre_flag[t]) R.resize(pre_val[t]);
	for (int c = 0; c < sv.n; ++c) {
		const uint64_t *k = sv.keys[c];
		for (uint64_t e = sv.off[c][s_first + t], e1 = sv.off[c][s_first + t + 1]; e < e1; ++e) R.entry(k[e]);
	}
	if (trailing[t]) R.check(); // a later put of an existing key (quirk Q3)
	out_cap[t] = R.n; out_size[t] = R.count;
	uint64_t r = 0;
	for (uint32_t i = 0; i < R.n; ++i) if (R.is_used(i)) R.K[r++] = R.K[i]; // r <= i: in-place compaction
}

// shared-memory variant: one warp per listed sub-table whose table never exceeds `slots` entries.
// dynamic smem: slots*8 (keys) + 2*fw(slots)*4 (bitmaps) + 512*8 (journal staging)
__global__ void __launch_bounds__(32) build_layout_smem_kernel(const int *__restrict__ list, uint32_t slots, SegView sv, int s_first,
                                    const uint8_t *__restrict__ pre_flag, const uint32_t *__restrict__ pre_val,
                                    const uint8_t *__restrict__ trailing,
                                    uint64_t *out_all, const uint64_t *__restrict__ ooff,
                                    uint32_t *out_cap, uint32_t *out_size)
{
	extern __shared__ unsigned char s_dyn[];
	const int t = list[blockIdx.x], lane = threadIdx.x;
	uint64_t *sK = (uint64_t*)s_dyn;
	uint32_t *sbm = (uint32_t*)(s_dyn + (size_t)slots * 8);
	const uint32_t bw = KhReplay::fw(slots);
	uint64_t *stage = (uint64_t*)(s_dyn + (size_t)slots * 8 + (size_t)bw * 8);
	KhReplay R;
	R.K = sK; R.used = sbm; R.occ = sbm + bw;
	if (lane == 0) { R.used[0] = 0; if (pre_flag[t]) R.resize(pre_val[t]); }
	for (int c = 0; c < sv.n; ++c) {
		const uint64_t *src = sv.keys[c] + sv.off[c][s_first + t];
		const uint64_t m = sv.off[c][s_first + t + 1] - sv.off[c][s_first + t];
		for (uint64_t e0 = 0; e0 < m; e0 += 512) {
			const uint32_t cc = (uint32_t)(m - e0 < 512 ? m - e0 : 512);
			__syncwarp();
			for (uint32_t i = lane; i < cc; i += 32) stage[i] = src[e0 + i]; // coalesced staging of the next 512 entries
			__syncwarp();
			if (lane == 0) for (uint32_t i = 0; i < cc; ++i) R.entry(stage[i]);
		}
	}
	if (lane == 0) {
		if (trailing[t]) R.check();
		out_cap[t] = R.n; out_size[t] = R.count;
	}
	// export in slot order with all lanes: every lane needs the final state of lane 0
	const uint32_t n = __shfl_sync(0xffffffffu, R.n, 0);
	const int swapped = __shfl_sync(0xffffffffu, (int)(R.used != sbm), 0);
	const uint32_t *used = swapped ? sbm + bw : sbm;
	__syncwarp();
	uint64_t *dst = out_all + ooff[t];
	uint32_t r = 0;
	for (uint32_t base = 0; base < n; base += 32) {
		const uint32_t i = base + lane;
		const bool u = i < n && (used[i >> 5] >> (i & 31) & 1);
		const uint32_t mask = __ballot_sync(0xffffffffu, u);
		if (u) dst[r + __popc(mask & ((1u << lane) - 1))] = sK[i];
		r += __popc(mask);
	}
}

// ---- the kick-out rehash (khashl.h:152-195) is one sequential walk per table, and on a table of megabytes every
//      step of it is a round trip to L2 or DRAM (~1 us per key when one lane does it alone).  Its accesses are two
//      streams, though - old slots upward, new homes at about twice that - so a few cache lines per array in shared
//      memory catch nearly all of them.  All 32 lanes run the walk in lock step on identical values (free under
//      SIMT); a missing line is fetched by the whole warp as one 128-byte load.  Write-through, and a write updates
//      the cached copy, so the cache never holds anything the table does not.
struct WarpCache {
	uint64_t k[64][16];  // key lines (16 slots)
	uint32_t u[16][32];  // lines of the old occupancy bitmap (1024 slots)
	uint32_t o[16][32];  // lines of the new one
	uint32_t kt[64], ut[16], ot[16];
};

__device__ __forceinline__ uint64_t wc_key(WarpCache &c, const uint64_t *K, uint32_t i, int lane)
{
	const uint32_t tag = i >> 4, s = tag & 63;
	if (c.kt[s] != tag) {
		__syncwarp();
		if (lane < 16) c.k[s][lane] = __ldcg(K + (uint64_t)tag * 16 + lane);
		c.kt[s] = tag;
		__syncwarp();
	}
	return c.k[s][i & 15];
}
__device__ __forceinline__ void wc_key_put(WarpCache &c, uint64_t *K, uint32_t i, uint64_t v)
{
	const uint32_t tag = i >> 4, s = tag & 63;
	K[i] = v;                                           // every lane, same address and value: one transaction
	if (c.kt[s] == tag) c.k[s][i & 15] = v;
}
// word w of a bitmap through its line cache (lines[16][32], tags[16])
__device__ __forceinline__ uint32_t wc_word(uint32_t (*lines)[32], uint32_t *tags, const uint32_t *bm, uint32_t w, int lane)
{
	const uint32_t tag = w >> 5, s = tag & 15;
	if (tags[s] != tag) {
		__syncwarp();
		lines[s][lane] = __ldcg(bm + (uint64_t)tag * 32 + lane);
		tags[s] = tag;
		__syncwarp();
	}
	return lines[s][w & 31];
}
__device__ __forceinline__ void wc_word_put(uint32_t (*lines)[32], uint32_t *tags, uint32_t *bm, uint32_t w, uint32_t v)
{
	const uint32_t tag = w >> 5, s = tag & 15;
	bm[w] = v;
	if (tags[s] == tag) lines[s][w & 31] = v;
}

// khashl.h:152-195 for a table that grows, run by the whole warp on identical state (R is the same in every lane)
__device__ void warp_resize(KhReplay &R, uint32_t request, WarpCache &c, int lane)
{
	uint32_t lg = 0, q = request;
	while ((q >>= 1) != 0) ++lg;
	if (request & (request - 1)) ++lg;
	const uint32_t new_bits = lg > 2 ? lg : 2, new_n = 1u << new_bits, new_mask = new_n - 1, n = R.n;
	if (R.count > (new_n >> 1) + (new_n >> 2)) return;
	if (new_n < n) { // shrinking is not on this kernel's path (puts only); keep the plain walk for it
		if (lane == 0) R.resize(request);
		__syncwarp();
		const int sw = __shfl_sync(0xffffffffu, (int)(R.bits != 0 && R.n != n), 0);
		if (sw && lane != 0) { uint32_t *tmp = R.used; R.used = R.occ; R.occ = tmp; }
		R.bits = __shfl_sync(0xffffffffu, R.bits, 0); R.n = __shfl_sync(0xffffffffu, R.n, 0);
		return;
	}
	for (uint32_t w = lane; w < KhReplay::fw(new_n); w += 32) R.occ[w] = 0;
	for (int i = lane; i < 64; i += 32) c.kt[i] = 0xFFFFFFFFu;
	if (lane < 16) c.ut[lane] = c.ot[lane] = 0xFFFFFFFFu;
	__syncwarp();
	for (uint32_t j = 0; j != n; ++j) {
		uint32_t uw = wc_word(c.u, c.ut, R.used, j >> 5, lane);
		if (!(uw >> (j & 31) & 1)) continue;
		uint64_t key = wc_key(c, R.K, j, lane);
		wc_word_put(c.u, c.ut, R.used, j >> 5, uw & ~(1u << (j & 31)));
		for (;;) {
			uint32_t i = kh_home(key, new_bits), ow;
			while ((ow = wc_word(c.o, c.ot, R.occ, i >> 5, lane)) >> (i & 31) & 1) i = (i + 1) & new_mask;
			wc_word_put(c.o, c.ot, R.occ, i >> 5, ow | 1u << (i & 31));
			if (i < n && ((uw = wc_word(c.u, c.ut, R.used, i >> 5, lane)) >> (i & 31) & 1)) { // an old key not moved yet: kick it out
				const uint64_t ev = wc_key(c, R.K, i, lane);
				wc_key_put(c, R.K, i, key);
				key = ev;
				wc_word_put(c.u, c.ut, R.used, i >> 5, uw & ~(1u << (i & 31)));
			} else { wc_key_put(c, R.K, i, key); break; }
		}
	}
	__syncwarp();
	uint32_t *tmp = R.used; R.used = R.occ; R.occ = tmp;
	R.bits = new_bits; R.n = new_n;
}

// warp variant for large sub-tables whose journal holds puts only (counting, shrink, restore):
// the FCFS put phases between two doublings run on all 32 lanes - a slot belongs to the lowest
// journal rank that probes it (atomicMin), a displaced rank moves on, which is exactly first-come
// first-served linear probing - while the kick-out rehash (khashl.h:152-195) stays with lane 0.
// cat = the sub-table's journal made contiguous, own = one u32 per slot (all ones = free).
__global__ void __launch_bounds__(32) build_layout_warp_kernel(const int *__restrict__ list,
                                    const uint64_t *__restrict__ cat_all, const uint64_t *__restrict__ catoff,
                                    const uint8_t *__restrict__ pre_flag, const uint32_t *__restrict__ pre_val,
                                    const uint8_t *__restrict__ trailing,
                                    uint64_t *out_all, const uint64_t *__restrict__ ooff,
                                    uint32_t *bm_all, const uint64_t *__restrict__ boff,
                                    uint32_t *own_all, uint32_t *out_cap, uint32_t *out_size, unsigned long long *phase_clk)
{
	const int t = list[blockIdx.x], lane = threadIdx.x;
	const uint64_t *cat = cat_all + catoff[t];
	const uint64_t m = catoff[t + 1] - catoff[t];
	uint32_t *own = own_all + ooff[t]; // same geometry as the key region
	KhReplay R;
	R.K = out_all + ooff[t];
	const uint64_t bwords = (boff[t + 1] - boff[t]) / 2;
	R.used = bm_all + boff[t]; R.occ = R.used + bwords;
	__shared__ WarpCache wcache;
	if (lane == 0) R.used[0] = 0;
	__syncwarp();
	if (pre_flag[t]) warp_resize(R, pre_val[t], wcache, lane); // R stays identical in all lanes throughout
	__syncwarp();
	uint64_t e0 = 0;
	long long t_rehash = 0, t_init = 0, t_put = 0, t_mat = 0, t0 = clock64(), t1;
	while (e0 < m) {
		if (R.count >= (R.n >> 1) + (R.n >> 2)) warp_resize(R, R.n + 1, wcache, lane); // doubling at the load limit (khashl.h:202-205)
		__syncwarp(); // the table as the rehash left it is visible to every lane from here
		t1 = clock64(); t_rehash += t1 - t0; t0 = t1;
		// broadcast the table state of lane 0
		const uint32_t n = __shfl_sync(0xffffffffu, R.n, 0), bits = __shfl_sync(0xffffffffu, R.bits, 0), count = __shfl_sync(0xffffffffu, R.count, 0);
		const int swapped = __shfl_sync(0xffffffffu, (int)(R.used != bm_all + boff[t]), 0);
		uint32_t *used = swapped ? bm_all + boff[t] + bwords : bm_all + boff[t];
		const uint32_t mask = n - 1, room = (n >> 1) + (n >> 2) - count; // puts until the next load check fires
		const uint64_t e1 = m - e0 < room ? m : e0 + room;
		for (uint32_t i = lane; i < n; i += 32) own[i] = 0xFFFFFFFFu;
		__syncwarp();
		t1 = clock64(); t_init += t1 - t0; t0 = t1;
		for (uint64_t eb = e0; eb < e1; eb += 32) { // 32 keys at a time, ranks = journal positions
			const uint64_t e = eb + lane;
			bool active = e < e1;
			uint32_t rank = (uint32_t)(e - e0), pos = active ? kh_home(cat[e], bits) : 0;
			while (active) {
				if (used[pos >> 5] >> (pos & 31) & 1) { pos = (pos + 1) & mask; continue; } // an older key
				const uint32_t old = atomicMin(&own[pos], rank);
				if (old == 0xFFFFFFFFu) active = false;              // free slot taken
				else { if (old > rank) rank = old; pos = (pos + 1) & mask; } // the later of the two moves on
			}
		}
		__syncwarp();
		t1 = clock64(); t_put += t1 - t0; t0 = t1;
		// materialise the phase: keys into their slots, occupancy bits; a lane owns whole bitmap words,
		// and reads the owners from L2 where the atomics put them
		for (uint32_t w = lane; w < (n < 32 ? 1u : n >> 5); w += 32) {
			uint32_t bitsw = used[w];
			for (uint32_t b = 0; b < 32 && w * 32 + b < n; ++b) {
				const uint32_t r = __ldcg(&own[w * 32 + b]);
				if (r != 0xFFFFFFFFu) { R.K[w * 32 + b] = cat[e0 + r]; bitsw |= 1u << b; }
			}
			used[w] = bitsw;
		}
		__syncwarp();
		R.count += (uint32_t)(e1 - e0);
		e0 = e1;
		t1 = clock64(); t_mat += t1 - t0; t0 = t1;
	}
	if (phase_clk && lane == 0) {
		atomicAdd(&phase_clk[0], (unsigned long long)t_rehash); atomicAdd(&phase_clk[1], (unsigned long long)t_init);
		atomicAdd(&phase_clk[2], (unsigned long long)t_put); atomicAdd(&phase_clk[3], (unsigned long long)t_mat);
	}
	if (trailing[t] && R.count >= (R.n >> 1) + (R.n >> 2)) warp_resize(R, R.n + 1, wcache, lane); // quirk Q3
	if (lane == 0) { out_cap[t] = R.n; out_size[t] = R.count; }
	__syncwarp();
	// in-place compaction to slot order (r <= i), cooperatively
	const uint32_t n = __shfl_sync(0xffffffffu, R.n, 0);
	const int swapped = __shfl_sync(0xffffffffu, (int)(R.used != bm_all + boff[t]), 0);
	const uint32_t *used = swapped ? bm_all + boff[t] + bwords : bm_all + boff[t];
	__syncwarp();
	uint32_t r = 0;
	for (uint32_t base = 0; base < n; base += 32) {
		const uint32_t i = base + lane;
		const bool u = i < n && (used[i >> 5] >> (i & 31) & 1);
		const uint64_t kv = u ? R.K[i] : 0;
		const uint32_t mk = __ballot_sync(0xffffffffu, u);
		__syncwarp();
		if (u) R.K[r + __popc(mk & ((1u << lane) - 1))] = kv; // r + rank <= i: never overtakes unread slots of later rounds
		r += __popc(mk);
		__syncwarp();
	}
}

// copy the runs of sub-tables given by `sub` (indices relative to s_first) of one journal segment to cat, behind earlier segments
__global__ void gather_seg_kernel(const uint64_t *__restrict__ seg_keys, const uint64_t *__restrict__ seg_off, int s_first,
                                  const int *__restrict__ sub, int nsub, const uint64_t *__restrict__ catoff, uint64_t *run, uint64_t *__restrict__ cat)
{
	// one block per listed sub-table
	const int t = sub[blockIdx.x];
	const uint64_t b = seg_off[s_first + t], n = seg_off[s_first + t + 1] - b;
	uint64_t *dst = cat + catoff[t] + run[t];
	for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = seg_keys[b + i];
	__syncthreads();
	if (threadIdx.x == 0) run[t] += n;
}

// dense[i] = the j-th key of sub-table t's run, for i = voff[t] + j, j < voff[t+1]-voff[t]
__global__ void densify_kernel(const uint64_t *__restrict__ out_all, const uint64_t *__restrict__ ooff, const uint64_t *__restrict__ voff,
                               int ns, uint64_t n, uint64_t *__restrict__ dense)
{
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t lo = 0, hi = ns;
	while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (voff[mid] <= i) lo = mid; else hi = mid; }
	dense[i] = out_all[ooff[lo] + (i - voff[lo])];
}

// attach the current counts to slot-ordered keys of sub-tables s0.. (off[] local to the range)
__global__ void fill_counts_kernel(uint64_t *keys, const uint64_t *__restrict__ off, int ns, int s0, uint64_t n,
                                   const uint64_t *__restrict__ slots, uint32_t cap)
{
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t lo = 0, hi = ns;
	while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (off[mid] <= i) lo = mid; else hi = mid; }
	const uint64_t *reg = slots + (uint64_t)(s0 + lo) * cap;
	uint64_t key = keys[i];
	int64_t q = tab_find(reg, cap, key >> YAKB_COUNTER_BITS);
	if (q >= 0) keys[i] = (key & ~(uint64_t)YAKB_MAX_COUNT) | slot_count(slots, (uint32_t)(s0 + lo), reg[q]);
}

// ============================================================ host side

Engine *Engine::create(int k, int pre, int n_hash, int n_shift, int rank, int world)
{
	if (pre < YAKB_COUNTER_BITS) return nullptr; // htab.c:17
	int lw = 0;
	while ((1 << lw) < world) ++lw;
	if ((1 << lw) != world || lw > pre || world > 16 || rank < 0 || rank >= world) { // 16: what the route kernels (extras.cu) partition by
		fprintf(stderr, "[yakb] ERROR: world size must be a power of two <= 16 and 0 <= rank < world\n");
		return nullptr;
	}
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0) {
		fprintf(stderr, "[yakb] ERROR: no CUDA device (%s); this library has no CPU path\n", cudaGetErrorString(e));
		return nullptr;
	}
	// the table is probed with independent 32-byte sector reads: ask L2 not to fetch wider lines
	if (!getenv("YAKB_L2_FETCH_DEFAULT")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
	Engine *g = new Engine;
	g->k = k, g->pre = pre, g->P = 1 << (pre - lw), g->lw = lw, g->rank = rank;
	YAKB_CUDA(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
	g->nkeys = (uint32_t*)dev_alloc(g->P * sizeof(uint32_t));
	g->last_put = (uint64_t*)dev_alloc(g->P * sizeof(uint64_t));
	g->last_new = (uint64_t*)dev_alloc(g->P * sizeof(uint64_t));
	YAKB_CUDA(cudaMemsetAsync(g->nkeys, 0, g->P * sizeof(uint32_t), g->stream));
	YAKB_CUDA(cudaMemsetAsync(g->last_put, 0, g->P * sizeof(uint64_t), g->stream));
	YAKB_CUDA(cudaMemsetAsync(g->last_new, 0, g->P * sizeof(uint64_t), g->stream));
	g->presize_flag.assign(g->P, 0);
	g->presize_val.assign(g->P, 0);
	g->nops.assign(g->P, 0);
	g->max_req.assign(g->P, 0);
	// htab.c:23-27 + bbf.c:9: a filter exists iff n_hash>0, n_shift>pre and 9 <= n_shift-pre <= 55
	if (n_hash > 0 && n_shift > pre) {
		g->n_hash = n_hash, g->n_shift = n_shift;
		int sub = n_shift - pre;
		if (sub >= 9 && sub + 9 <= 64) {
			g->nb = sub - 9;
			size_t bytes = (size_t)1 << (n_shift - 3 - lw);
			g->bloom = (uint8_t*)dev_alloc(bytes);
			YAKB_CUDA(cudaMemsetAsync(g->bloom, 0, bytes, g->stream));
		}
	}
	YAKB_CUDA(cudaStreamSynchronize(g->stream));
	return g;
}

Engine::~Engine()
{
	free_slots(); dev_free(nkeys); dev_free(bloom); dev_free(last_put); dev_free(last_new);
	journal_free_all();
	DBuf *all[] = {&b_w2, &b_wm, &b_flags, &b_tilecnt, &b_tileoff, &b_pv, &b_ppos, &b_sv, &b_sj, &b_sv2, &b_sj2, &b_pflag, &b_newv,
	               &b_newsorted, &b_tmp, &b_pend, &b_lput, &b_lnew, &b_stats, &b_misc, &b_zev, &b_zpos, &b_zsp, &b_zspp, &b_zfill,
	               &b_lay[0], &b_lay[1], &b_lay[2], &b_lay[3], &b_lay[4], &b_lay[5], &b_lay[6], &b_lay[7], &b_lay[8], &b_lay[9], &b_lay[10], &b_lay[11]};
	for (DBuf *b : all) b->release();
	rs.release();
	b_segp.release();
	if (stream) cudaStreamDestroy(stream);
}

void *Engine::journal_alloc(size_t bytes)
{
	bytes = (bytes + 255) & ~(size_t)255;
	if (slabs.empty() || slabs.back().used + bytes > slabs.back().cap) {
		Slab sl;
		sl.cap = std::max<size_t>(bytes, (size_t)2 << 30);
		// The first slab of a table with a filter is sized from the filter: -b tells the scale of the job (2^37 bits are meant for
		// ~3 G distinct k-mers = 25 GB of journal), and cudaMalloc costs 1-30 ms per GiB depending on the box - a 2 GiB slab every
		// other chunk was 40 ms per bench step on some boxes and none on others.  One allocation, in the first chunk.
		if (slabs.empty() && bloom) sl.cap = std::max<size_t>(sl.cap, ((size_t)3 << (n_shift - 3 - lw)) / 2);
		sl.used = 0;
		sl.p = (char*)dev_alloc(sl.cap);
		slabs.push_back(sl);
	}
	void *r = slabs.back().p + slabs.back().used;
	slabs.back().used += bytes;
	return r;
}

void Engine::journal_free_all()
{
	for (auto &sl : slabs) dev_free(sl.p);
	slabs.clear();
	journal.clear();
}

void Engine::destroy_bloom() { if (bloom) { dev_free(bloom); bloom = nullptr; } }
void Engine::free_slots() { if (slots) dev_free((uint8_t*)slots - YAKB_SAT_BYTES); slots = nullptr; }

uint64_t Engine::device_bytes() const
{
	uint64_t b = (uint64_t)P * cap * 8 + (bloom ? (uint64_t)1 << (n_shift - 3 - lw) : 0);
	for (auto &sl : slabs) b += sl.cap;
	return b;
}

void Engine::grow(uint32_t new_cap)
{
	uint64_t *ns = nullptr;
	const uint64_t total_new = (uint64_t)P * new_cap;
	if (total_new * 8 >= (4ull << 30)) { // old + new table live side by side: the partition lists (dead here) make room
		YAKB_CUDA(cudaStreamSynchronize(stream));
		b_zev.release(); b_zpos.release(); b_zsp.release(); b_zspp.release();
	}
	// YAKB_SAT_BYTES of flags in front of the table (yakb_dev.cuh): they move with the table
	uint8_t *base = (uint8_t*)dev_alloc(total_new * 8 + YAKB_SAT_BYTES);
	ns = (uint64_t*)(base + YAKB_SAT_BYTES);
	YAKB_CUDA(cudaMemsetAsync(base, 0, YAKB_SAT_BYTES, stream));
	YAKB_CUDA(cudaMemsetAsync(ns, 0xFF, total_new * 8, stream));
	if (slots) YAKB_CUDA(cudaMemcpyAsync(base, (uint8_t*)slots - YAKB_SAT_BYTES, YAKB_SAT_BYTES, cudaMemcpyDeviceToDevice, stream));
	if (slots && cap) {
		const uint64_t total_old = (uint64_t)P * cap;
		rehash_kernel<<<cdiv(total_old, 256), 256, 0, stream>>>(slots, cap, total_old, ns, new_cap);
		YAKB_CUDA(cudaGetLastError());
	}
	if (slots) {
		YAKB_CUDA(cudaStreamSynchronize(stream));
		free_slots();
	}
	slots = ns; cap = new_cap;
}

void Engine::reserve(uint64_t keys_per_subtable)
{
	uint64_t want = ((uint64_t)(keys_per_subtable / load_limit) + 16 + 3) & ~3ull;
	if (want > 0xFFFFFFF0ull) throw CudaError("[yakb] sub-table capacity overflow");
	if (want > cap) grow((uint32_t)want);
}

static int smem_lp_ok(int P, int arrays) { return (size_t)P * 4 * arrays <= 64 * 1024; }

template<class K> static void set_smem(K kern, size_t bytes)
{
	if (bytes > 48 * 1024) YAKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

ChunkStats Engine::count_ascii(const uint8_t *d_asc, uint64_t n, int create_new)
{
	ChunkStats st = {0, 0, 0, 0};
	if (n == 0) return st;
	if (n >= 0xFFFFFF00ull) throw CudaError("[yakb] chunk too large (positions are 32-bit)");
	const uint64_t nwords = (n + 31) / 32;
	uint64_t *w2 = b_w2.as<uint64_t>(packed_words(nwords)) + YAKB_PADW;
	uint32_t *wm = b_wm.as<uint32_t>(packed_words(nwords)) + YAKB_PADW;
	{ ProfScope ps("pack_ascii", stream);
	pack_ascii_kernel<<<cdiv(packed_npad(nwords) + YAKB_PADW, 256), 256, 0, stream>>>(d_asc, n, w2, wm, nwords, packed_npad(nwords)); }
	Prof::units("pack_ascii", n);
	YAKB_CUDA(cudaGetLastError());
	note_launch(1);
	return finish_chunk(nwords, create_new, w2, wm, nullptr, n, -1);
}

ChunkStats Engine::count_events(const uint64_t *d_ev, uint64_t n, int create_new, int only_s, bool ignore_bloom)
{
	ChunkStats st = {0, 0, 0, 0};
	if (n == 0) return st;
	if (n >= 0xFFFFFF00ull) throw CudaError("[yakb] chunk too large (positions are 32-bit)");
	return finish_chunk((n + 31) / 32, create_new, nullptr, nullptr, d_ev, n, only_s, ignore_bloom);
}

// ---- stage 1 of a chunk, partitioned (count.c:17-26): every event is written to the list of its zone (part_scatter),
//      then the table is probed zone by zone (zone_probe).  Returns false - having touched neither the table nor flags -
//      when the lists overflowed (heavily repeated k-mers): the caller then probes unpartitioned.
bool Engine::probe_partitioned(uint64_t nwords, int create_new, const uint64_t *w2, const uint32_t *wm, const uint64_t *d_ev, uint64_t n_units,
                               int only_s, uint32_t *flags, uint32_t *tilecnt, uint32_t *lput, unsigned long long *stats, int nsm)
{
	if (cap == 0) return false;
	// YAKB_ZONE: 0 = never, 1 = always (tests), unset / 2 = when it pays: a table far beyond L2 and a chunk large enough
	// that the scatter's fixed costs (Z reservations per tile) are amortised.  YAKB_ZONE_MB = size of a zone's table slice.
	static const int zmode = getenv("YAKB_ZONE") ? atoi(getenv("YAKB_ZONE")) : 2;
	static const double zmb = getenv("YAKB_ZONE_MB") ? atof(getenv("YAKB_ZONE_MB")) : 64.0;
	static const uint64_t zmin_pos = getenv("YAKB_ZONE_MIN_POS") ? (uint64_t)atoll(getenv("YAKB_ZONE_MIN_POS")) : (32ull << 20);
	const uint64_t n_pos = d_ev ? n_units : nwords * 32;
	const double table_bytes = (double)P * cap * 8;
	if (zmode == 0 || (zmode != 1 && (table_bytes < 2e9 || n_pos < zmin_pos))) return false;
	int zshift = 0;
	while ((1 << zshift) < P && (double)cap * 8 * (2 << zshift) <= zmb * 1048576.0) ++zshift;
	uint32_t Z = (uint32_t)P >> zshift;
	while (Z > 2048) { ++zshift; Z >>= 1; } // the reserve step of part_scatter gives a thread at most 8 zones
	if (Z < 2) return false;
	// list capacity: the events this chunk is expected to hold (positions x the events-per-position ratio of the chunks
	// before it; 1 until one was seen), spread evenly, plus 12.5 % and a constant; the rest spills
	static const uint64_t zslack = getenv("YAKB_ZONE_SLACK") ? (uint64_t)atoll(getenv("YAKB_ZONE_SLACK")) : 2048; // test knob: 0 forces spills
	const double ratio = d_ev ? 1.0 : ev_ratio;
	const uint64_t est = (uint64_t)((double)n_pos * ratio) + 1;
	const uint32_t zcap = (uint32_t)std::min<uint64_t>(n_pos, est / Z + est / Z / 8 + zslack);
	const uint32_t spill_cap = (uint32_t)std::min<uint64_t>(n_pos, est / 16 + 65536);
	uint64_t *zev = b_zev.as<uint64_t>((uint64_t)Z * zcap), *sp_ev = b_zsp.as<uint64_t>(spill_cap);
	uint32_t *zpos = b_zpos.as<uint32_t>((uint64_t)Z * zcap), *sp_pos = b_zspp.as<uint32_t>(spill_cap);
	unsigned int *zfill = b_zfill.as<unsigned int>(Z + 4); // [Z] fills, n_spill, work, work2
	YAKB_CUDA(cudaMemsetAsync(zfill, 0, (Z + 4) * 4, stream));
	const size_t sms = part_smem_bytes(Z);
	const uint64_t ntiles = (nwords + 255) / 256;
	const uint32_t grids = (uint32_t)std::min<uint64_t>(ntiles, (uint64_t)nsm * std::max<size_t>(1, std::min<size_t>(2, (227 * 1024) / (sms + 1024))));
	const bool longk = k >= 32;
	{ ProfScope ps("part_scatter", stream);
	if (d_ev) {
		set_smem(part_scatter<false, true>, sms);
		part_scatter<false, true><<<grids, 256, sms, stream>>>(nullptr, nullptr, nwords, k, d_ev, n_units, only_s, P - 1, own(), zshift, Z, zcap, zfill,
		                                                        zev, zpos, sp_ev, sp_pos, zfill + Z, spill_cap, stats);
	} else if (longk) {
		set_smem(part_scatter<true, false>, sms);
		part_scatter<true, false><<<grids, 256, sms, stream>>>(w2, wm, nwords, k, nullptr, 0, -1, P - 1, own(), zshift, Z, zcap, zfill,
		                                                        zev, zpos, sp_ev, sp_pos, zfill + Z, spill_cap, stats);
	} else {
		set_smem(part_scatter<false, false>, sms);
		part_scatter<false, false><<<grids, 256, sms, stream>>>(w2, wm, nwords, k, nullptr, 0, -1, P - 1, own(), zshift, Z, zcap, zfill,
		                                                         zev, zpos, sp_ev, sp_pos, zfill + Z, spill_cap, stats);
	} }
	YAKB_CUDA(cudaGetLastError());
	note_launch(1);
	unsigned int n_spill = 0;
	YAKB_CUDA(cudaMemcpyAsync(&n_spill, zfill + Z, 4, cudaMemcpyDeviceToHost, stream));
	YAKB_CUDA(cudaStreamSynchronize(stream));
	if (n_spill > spill_cap) { // events were dropped: start over without the partition (nothing but stats[0] was written)
		YAKB_CUDA(cudaMemsetAsync(stats, 0, 4 * sizeof(unsigned long long), stream));
		return false;
	}
	const int smem1 = create_new && smem_lp_ok(P, 1);
	const size_t sm1 = smem1 ? (size_t)P * 4 : 0;
	if (create_new) YAKB_CUDA(cudaMemsetAsync(flags, 0, nwords * 4, stream));
	{ ProfScope ps("zone_probe", stream);
	const uint32_t zsub = (1u << zshift) <= 8 ? (1u << zshift) : 0;
	static const int zpref = getenv("YAKB_ZPROBE_PREFETCH") ? atoi(getenv("YAKB_ZPROBE_PREFETCH")) : 0;
	static const int zocc = getenv("YAKB_ZPROBE_OCC") ? atoi(getenv("YAKB_ZPROBE_OCC")) : 2; // resident CTAs per SM (2: no register spills; measured 11.5 vs 10.2 G events/s for 3)
	if (zocc == 2) {
		set_smem(zone_probe<2>, sm1);
		zone_probe<2><<<nsm * 2, 256, sm1, stream>>>(zev, zpos, Z, zcap, zfill, pre, P - 1, slots, cap, create_new, flags, lput, smem1, zfill + Z + 1, zsub, zpref != 0);
	} else {
		set_smem(zone_probe<3>, sm1);
		zone_probe<3><<<nsm * 3, 256, sm1, stream>>>(zev, zpos, Z, zcap, zfill, pre, P - 1, slots, cap, create_new, flags, lput, smem1, zfill + Z + 1, zsub, zpref != 0);
	}
	// the spill list: one more list whose fill count is n_spill
	if (n_spill) set_smem(zone_probe<3>, sm1);
	if (n_spill) zone_probe<3><<<nsm * 3, 256, sm1, stream>>>(sp_ev, sp_pos, 1, spill_cap, zfill + Z, pre, P - 1, slots, cap, create_new, flags, lput, smem1, zfill + Z + 2, 0, false);
	if (create_new) flag_tilecnt_kernel<<<(uint32_t)ntiles, 256, 0, stream>>>(flags, nwords, tilecnt); }
	YAKB_CUDA(cudaGetLastError());
	note_launch(1 + (n_spill ? 1 : 0) + (create_new ? 1 : 0));
	return true;
}

// ---- stage 2 of a pass-1 chunk for the pending events of tiles [t0, t1) (n_pending of them, the first at index off0 of
//      the chunk's pending order): file-order list -> stable sort by group -> group_insert -> bookkeeping -> journal segment.
//      Returns the number of new keys.  Pending events see the table as the ranges before them left it, so cutting a chunk
//      into ranges changes nothing; it bounds the scratch memory of a large chunk in the filling phase.
uint64_t Engine::pending_range(uint64_t t0, uint64_t t1, uint32_t off0, uint32_t n_pending, uint64_t nwords, const uint64_t *w2, const uint32_t *wm,
                               const uint64_t *d_ev, const uint32_t *flags, const uint32_t *tileoff, uint32_t *lput, uint32_t *lnew,
                               unsigned long long *stats, uint8_t *bloom, int nsm)
{
	if (n_pending == 0) return 0;
	const uint32_t Pmask = P - 1;
	const bool longk = k >= 32;
	uint64_t *pv = b_pv.as<uint64_t>(n_pending);
	uint32_t *ppos = b_ppos.as<uint32_t>(n_pending);
	{ ProfScope ps("compact", stream);
	if (d_ev == nullptr) {
		if (longk) compact_fused<true><<<(uint32_t)(t1 - t0), 256, 0, stream>>>(w2, wm, nwords, k, flags, tileoff, t0, off0, pv, ppos);
		else compact_fused<false><<<(uint32_t)(t1 - t0), 256, 0, stream>>>(w2, wm, nwords, k, flags, tileoff, t0, off0, pv, ppos);
	} else compact_array<<<(uint32_t)(t1 - t0), 256, 0, stream>>>(d_ev, nwords, flags, tileoff, t0, off0, pv, ppos);
	}
	YAKB_CUDA(cudaGetLastError());
	// make room: every pending event may be a new key of its sub-table
	const int smem1h = smem_lp_ok(P, 1);
	const size_t sm1h = smem1h ? (size_t)P * 4 : 0;
	uint32_t *pend = b_pend.as<uint32_t>(P + 1);
	YAKB_CUDA(cudaMemsetAsync(pend, 0, (P + 1) * 4, stream));
	{ ProfScope ps("pend_hist", stream);
	set_smem(pend_hist_kernel, sm1h);
	pend_hist_kernel<<<std::min<uint32_t>(cdiv(n_pending, 256), nsm * 4), 256, sm1h, stream>>>(pv, n_pending, Pmask, pend, smem1h);
	max_need_kernel<<<1, 1024, 0, stream>>>(nkeys, pend, P, pend + P); }
	uint32_t need = 0;
	YAKB_CUDA(cudaMemcpyAsync(&need, pend + P, 4, cudaMemcpyDeviceToHost, stream));
	YAKB_CUDA(cudaStreamSynchronize(stream));
	if ((double)need > load_limit * cap) {
		// double; small tables (below 8 GB after the step) go up four times at once: half as many rehashes and allocations on the way up
		const uint64_t mult = (uint64_t)P * cap * 8 * 4 <= (8ull << 30) ? 4 : 2;
		uint64_t want = (std::max<uint64_t>((uint64_t)(need / load_limit) + 16, (uint64_t)cap * mult) + 3) & ~3ull;
		if (want > 0xFFFFFFF0ull) throw CudaError("[yakb] sub-table capacity overflow");
		ProfScope ps("grow", stream);
		grow((uint32_t)want);
	}
	// group key: the bloom block (bbf.c:27-28: low n_shift-9 bits of the hash, sub-table included)
	// or, without a filter, enough low hash bits to keep groups short
	int G;
	const int vbits = longk ? 64 : 2 * k;
	if (bloom) G = n_shift - 9;
	else { G = pre; while (G < vbits && (1ull << G) < (uint64_t)n_pending / 2) ++G; }
	if (G > vbits) G = vbits;
	uint64_t *sv = b_sv.as<uint64_t>(n_pending), *sv2 = b_sv2.as<uint64_t>(n_pending);
	uint32_t *sj = b_sj.as<uint32_t>(n_pending), *sj2 = b_sj2.as<uint32_t>(n_pending);
	{ ProfScope ps("group_sort", stream);
	if (radix_sort_pairs(pv, nullptr, sv, sj, sv2, sj2, n_pending, 0, G, stream, rs)) { sv = sv2; sj = sj2; } }
	const size_t pf_words = ((size_t)n_pending + 15) / 16;
	uint32_t *pflag = b_pflag.as<uint32_t>(pf_words + ((size_t)n_pending + 3) / 4 + 4); // 2 bits per event, then one new-key byte per event
	uint8_t *isnew_w = (uint8_t*)(pflag + pf_words);
	YAKB_CUDA(cudaMemsetAsync(pflag, 0, pf_words * 4, stream));
	{ ProfScope ps("group_insert", stream);
	group_insert<<<std::min<uint32_t>(cdiv(n_pending, 1024), nsm * 8), 256, 0, stream>>>(sv, sj, n_pending, G, pre, Pmask, lw, slots, cap,
	                                                                                  (uint32_t*)bloom, nb, n_shift - pre, n_hash, pflag); }
	YAKB_CUDA(cudaGetLastError());
	const int smem2 = smem_lp_ok(P, 2);
	const size_t sm2 = smem2 ? (size_t)P * 8 : 0;
	set_smem(post_pending, sm2);
	{ ProfScope ps("post_pending", stream);
	post_pending<<<std::min<uint32_t>(cdiv(n_pending, 256), nsm * 4), 256, sm2, stream>>>(pv, ppos, pflag, isnew_w, n_pending, Pmask, lput, lnew, smem2, stats); }
	YAKB_CUDA(cudaGetLastError());
	note_launch(5); // compact, pend_hist, max_need, group_insert, post_pending
	// new keys in file order, then stably by sub-table -> journal segment
	uint64_t *newv = b_newv.as<uint64_t>(n_pending);
	uint32_t *d_nsel = (uint32_t*)(stats + 2);
	const uint8_t *isnew = isnew_w; // written by post_pending
	{ ProfScope ps("journal(compact)", stream);
	compact_flagged_u64(pv, isnew, n_pending, newv, d_nsel, stream, rs); }
	uint32_t n_new = 0;
	YAKB_CUDA(cudaMemcpyAsync(&n_new, d_nsel, 4, cudaMemcpyDeviceToHost, stream));
	YAKB_CUDA(cudaStreamSynchronize(stream));
	if (n_new) {
		// stable by sub-table (the low pre-lw bits): newv is in file order, so each run is in first-put order
		uint64_t *sorted = b_newsorted.as<uint64_t>(n_new), *sorted2 = b_sv.as<uint64_t>(n_new);
		Segment seg;
		seg.n = n_new;
		seg.keys = (uint64_t*)journal_alloc((uint64_t)n_new * 8);
		seg.off = (uint64_t*)journal_alloc((uint64_t)(P + 1) * 8);
		ProfScope ps("journal(sort+seg)", stream);
		if (radix_sort_pairs(newv, nullptr, sorted, nullptr, sorted2, nullptr, n_new, 0, pre - lw, stream, rs)) sorted = sorted2;
		seg_offsets_kernel<<<cdiv(P + 1, 256), 256, 0, stream>>>(sorted, n_new, Pmask, seg.off, nkeys);
		seg_addkeys_kernel<<<cdiv(P, 256), 256, 0, stream>>>(seg.off, P, nkeys);
		seg_keys_kernel<<<cdiv(n_new, 256), 256, 0, stream>>>(sorted, n_new, pre, seg.keys);
		YAKB_CUDA(cudaGetLastError());
		journal.push_back(seg);
		note_launch(3);
	}
	return n_new;
}

// work items per kernel name of one chunk, for the roofline accounting of bench.py (SURVEY 8(d) bytes are per event)
static void note_units(bool parted, bool array, const ChunkStats &st)
{
	if (!Prof::on()) return;
	if (parted) { Prof::units("part_scatter", st.n_events); Prof::units("zone_probe", st.n_events); }
	else Prof::units(array ? "k1_array" : "k1_fused", st.n_events);
	for (const char *nm : {"compact", "pend_hist", "group_sort", "group_insert", "post_pending", "journal(compact)"}) Prof::units(nm, st.n_pending);
	Prof::units("journal(sort+seg)", st.n_new);
}

// common tail of both front ends.  n_units: #bases (ASCII) or #events (array front end)
ChunkStats Engine::finish_chunk(uint64_t nwords, int create_new, const uint64_t *w2, const uint32_t *wm,
                                const uint64_t *d_ev, uint64_t n_units, int only_s, bool ignore_bloom)
{
	uint8_t *const bloom = ignore_bloom ? nullptr : this->bloom; // yak_ch_merge puts straight into the set (htab.c:262)
	ChunkStats st = {0, 0, 0, 0};
	const uint32_t Pmask = P - 1;
	const bool longk = k >= 32;
	const uint64_t n_ev_in = n_units;
	const uint64_t ntiles = (nwords + 255) / 256;
	uint32_t *flags = create_new ? b_flags.as<uint32_t>(nwords) : nullptr;
	uint32_t *tilecnt = create_new ? b_tilecnt.as<uint32_t>(ntiles + 1) : nullptr;
	uint32_t *tileoff = create_new ? b_tileoff.as<uint32_t>(ntiles + 1) : nullptr;
	uint32_t *lput = b_lput.as<uint32_t>(P), *lnew = b_lnew.as<uint32_t>(P);
	unsigned long long *stats = b_stats.as<unsigned long long>(4);
	YAKB_CUDA(cudaMemsetAsync(stats, 0, 4 * sizeof(unsigned long long), stream));
	if (create_new) {
		YAKB_CUDA(cudaMemsetAsync(lput, 0, P * 4, stream));
		YAKB_CUDA(cudaMemsetAsync(lnew, 0, P * 4, stream));
		YAKB_CUDA(cudaMemsetAsync(tilecnt, 0, (ntiles + 1) * 4, stream));
	}
	int dev = 0, nsm = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
	// ---- stage 1: every event meets the table (hit: counter++; miss: the position is flagged pending)
	const bool parted = probe_partitioned(nwords, create_new, w2, wm, d_ev, n_units, only_s, flags, tilecnt, lput, stats, nsm);
	if (!parted) {
		const int smem1 = create_new && smem_lp_ok(P, 1);
		const size_t sm1 = smem1 ? (size_t)P * 4 : 0;
		const uint32_t grid1 = (uint32_t)std::min<uint64_t>(ntiles, (uint64_t)nsm * 4);
		ProfScope ps(d_ev ? "k1_array" : "k1_fused", stream);
		if (d_ev == nullptr) {
			if (longk) {
				set_smem(k1_fused<true>, sm1);
				k1_fused<true><<<grid1, 256, sm1, stream>>>(w2, wm, nwords, k, pre, Pmask, own(), slots, cap, create_new, flags, tilecnt, lput, smem1, stats);
			} else {
				set_smem(k1_fused<false>, sm1);
				k1_fused<false><<<grid1, 256, sm1, stream>>>(w2, wm, nwords, k, pre, Pmask, own(), slots, cap, create_new, flags, tilecnt, lput, smem1, stats);
			}
		} else {
			set_smem(k1_array, sm1);
			k1_array<<<grid1, 256, sm1, stream>>>(d_ev, n_ev_in, pre, Pmask, own(), slots, cap, create_new, only_s, flags, tilecnt, lput, smem1, stats);
		}
		YAKB_CUDA(cudaGetLastError());
		note_launch(1);
	}
	unsigned long long h_stats[4] = {0, 0, 0, 0};
	if (!create_new) {
		YAKB_CUDA(cudaMemcpyAsync(h_stats, stats, sizeof(h_stats), cudaMemcpyDeviceToHost, stream));
		YAKB_CUDA(cudaStreamSynchronize(stream));
		Prof::resolve();
		st.n_events = h_stats[0];
		note_ratio(st.n_events, d_ev ? 0 : nwords * 32);
		note_units(parted, d_ev != nullptr, st);
		++chunk_seq;
		return st;
	}
	// ---- stage 2: the pending events in file order, in ranges of tiles of about pend_max events
	exclusive_scan_u32(tilecnt, tileoff, ntiles + 1, stream, rs);
	uint32_t n_pending = 0;
	YAKB_CUDA(cudaMemcpyAsync(&n_pending, tileoff + ntiles, 4, cudaMemcpyDeviceToHost, stream));
	YAKB_CUDA(cudaStreamSynchronize(stream));
	st.n_pending = n_pending;
	uint64_t n_new = 0;
	if (n_pending) {
		static const uint64_t pend_max = getenv("YAKB_PEND_MAX") ? (uint64_t)atoll(getenv("YAKB_PEND_MAX")) : (256ull << 20);
		const uint64_t nsub = std::min<uint64_t>(ntiles, (n_pending + pend_max - 1) / pend_max);
		std::vector<uint64_t> tb(nsub + 1);
		std::vector<uint32_t> to(nsub + 1, 0);
		for (uint64_t i = 0; i <= nsub; ++i) tb[i] = ntiles * i / nsub;
		to[nsub] = n_pending;
		if (nsub > 1) {
			for (uint64_t i = 1; i < nsub; ++i) YAKB_CUDA(cudaMemcpyAsync(&to[i], tileoff + tb[i], 4, cudaMemcpyDeviceToHost, stream));
			YAKB_CUDA(cudaStreamSynchronize(stream));
		}
		for (uint64_t i = 0; i < nsub; ++i)
			n_new += pending_range(tb[i], tb[i + 1], to[i], to[i + 1] - to[i], nwords, w2, wm, d_ev, flags, tileoff, lput, lnew, stats, bloom, nsm);
	}
	YAKB_CUDA(cudaMemcpyAsync(h_stats, stats, sizeof(h_stats), cudaMemcpyDeviceToHost, stream));
	merge_times_kernel<<<cdiv(P, 256), 256, 0, stream>>>(P, chunk_seq, lput, lnew, last_put, last_new);
	YAKB_CUDA(cudaGetLastError());
	note_launch(1);
	YAKB_CUDA(cudaStreamSynchronize(stream));
	Prof::resolve();
	st.n_events = h_stats[0];
	st.n_put = h_stats[1] + (st.n_events - st.n_pending);
	st.n_new = n_new;
	tot += n_new;
	note_ratio(st.n_events, d_ev ? 0 : nwords * 32);
	note_units(parted, d_ev != nullptr, st);
	++chunk_seq;
	return st;
}

// ---- restore into an existing table (htab.c:436-472): key i (stored form, its low bits are the bits to
//      set) is put into its sub-table; an existing key gets the bits OR-ed in when `or_bits`, else stays as
//      it is.  isnew[i] = the put inserted the key; newcnt[s] / lnew[s] = new keys / last new position (+1).
__global__ void upsert_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ off, uint32_t P, uint64_t n,
                              uint64_t *slots, uint32_t cap, int or_bits, uint8_t *__restrict__ isnew, uint32_t *newcnt, uint32_t *lnew)
{
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t lo = 0, hi = P; // sub-table s with off[s] <= i < off[s+1]
	while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (off[mid] <= i) lo = mid; else hi = mid; }
	uint64_t val = keys[i];
	uint64_t *slot, cur;
	if (val == YAKB_EMPTY) { val = YAKB_ALMOST_EMPTY; sat_of(slots)[lo & (YAKB_SAT_BYTES - 1)] = 1; } // YAKB_SAT_BYTES
	if (tab_insert(slots + (uint64_t)lo * cap, cap, val, &slot, &cur)) {
		isnew[i] = 1;
		atomicAdd(&newcnt[lo], 1u);
		atomicMax(&lnew[lo], (uint32_t)(i - off[lo]) + 1);
	} else {
		isnew[i] = 0;
		if (or_bits && (val & YAKB_MAX_COUNT)) {
			// OR the bits in; the result EMPTY (all id bits and all ten low bits set) is stored as ALMOST_EMPTY + the flag
			uint64_t bits = keys[i] & YAKB_MAX_COUNT;
			for (;;) {
				uint64_t want = cur | bits;
				if (want == YAKB_EMPTY) { want = YAKB_ALMOST_EMPTY; sat_of(slots)[lo & (YAKB_SAT_BYTES - 1)] = 1; }
				if (want == cur) break;
				const uint64_t prev = atomicCAS((unsigned long long*)slot, (unsigned long long)cur, (unsigned long long)want);
				if (prev == cur) break;
				cur = prev;
			}
		}
	}
}

__global__ void upsert_times_kernel(uint32_t P, uint32_t seq, const uint64_t *__restrict__ off, const uint32_t *__restrict__ lnew,
                                    const uint32_t *__restrict__ newcnt, uint64_t *last_put, uint64_t *last_new, uint32_t *nkeys)
{
	uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= P) return;
	const uint64_t m = off[s + 1] - off[s];
	if (m) last_put[s] = (uint64_t)(seq + 1) << 32 | m;            // every key of the file is a put
	if (lnew[s]) last_new[s] = (uint64_t)(seq + 1) << 32 | lnew[s];
	nkeys[s] += newcnt[s];
}

uint64_t Engine::upsert(const std::vector<uint32_t> &caps, const std::vector<uint64_t> &off, const uint64_t *keys, bool or_bits)
{
	// yak_ht_resize(h, capacity in the file) on every sub-table first (htab.c:438), as journal operations
	std::vector<uint64_t> ops(P);
	for (int s = 0; s < P; ++s) ops[s] = (uint64_t)caps[s] << YAKB_COUNTER_BITS | OP_RESIZE;
	append_ops(ops);
	const uint64_t n = off[P];
	if (n == 0) return 0;
	std::vector<uint32_t> have;
	sizes(have);
	uint64_t mx = 8;
	for (int s = 0; s < P; ++s) mx = std::max<uint64_t>(mx, (uint64_t)have[s] + (off[s + 1] - off[s]));
	reserve(mx);
	uint64_t *d_keys = b_pv.as<uint64_t>(n), *d_off = b_tmp.as<uint64_t>(P + 1), *d_new = b_newv.as<uint64_t>(n);
	uint8_t *isnew = b_pflag.as<uint8_t>(n);
	uint32_t *newcnt = b_pend.as<uint32_t>(P + 1), *lnew = b_lnew.as<uint32_t>(P);
	unsigned long long *stats = b_stats.as<unsigned long long>(4);
	YAKB_CUDA(cudaMemcpyAsync(d_keys, keys, n * 8, cudaMemcpyHostToDevice, stream));
	YAKB_CUDA(cudaMemcpyAsync(d_off, off.data(), (uint64_t)(P + 1) * 8, cudaMemcpyHostToDevice, stream));
	YAKB_CUDA(cudaMemsetAsync(newcnt, 0, (P + 1) * 4, stream));
	YAKB_CUDA(cudaMemsetAsync(lnew, 0, P * 4, stream));
	YAKB_CUDA(cudaMemsetAsync(stats, 0, 4 * sizeof(unsigned long long), stream));
	upsert_kernel<<<cdiv(n, 256), 256, 0, stream>>>(d_keys, d_off, P, n, slots, cap, or_bits ? 1 : 0, isnew, newcnt, lnew);
	YAKB_CUDA(cudaGetLastError());
	compact_flagged_u64(d_keys, isnew, n, d_new, (uint32_t*)(stats + 2), stream, rs); // new keys, file order = sub-table order
	std::vector<uint32_t> h_newcnt(P);
	YAKB_CUDA(cudaMemcpyAsync(h_newcnt.data(), newcnt, P * 4, cudaMemcpyDeviceToHost, stream));
	YAKB_CUDA(cudaStreamSynchronize(stream));
	std::vector<uint64_t> noff(P + 1, 0);
	for (int s = 0; s < P; ++s) noff[s + 1] = noff[s] + h_newcnt[s];
	const uint64_t n_new = noff[P];
	if (n_new) {
		Segment seg;
		seg.n = n_new;
		seg.keys = (uint64_t*)journal_alloc(n_new * 8);
		seg.off = (uint64_t*)journal_alloc((uint64_t)(P + 1) * 8);
		YAKB_CUDA(cudaMemcpyAsync(seg.keys, d_new, n_new * 8, cudaMemcpyDeviceToDevice, stream));
		YAKB_CUDA(cudaMemcpyAsync(seg.off, noff.data(), (uint64_t)(P + 1) * 8, cudaMemcpyHostToDevice, stream));
		strip_counts_kernel<<<cdiv(n_new, 256), 256, 0, stream>>>(seg.keys, n_new); // journal entries are puts: no low bits
		YAKB_CUDA(cudaGetLastError());
		journal.push_back(seg);
	}
	upsert_times_kernel<<<cdiv(P, 256), 256, 0, stream>>>(P, chunk_seq, d_off, lnew, newcnt, last_put, last_new, nkeys);
	YAKB_CUDA(cudaGetLastError());
	YAKB_CUDA(cudaStreamSynchronize(stream)); // noff is read by the copy above until here
	note_launch(4);
	++chunk_seq;
	tot += n_new;
	return n_new;
}

void Engine::clear()
{
	const uint64_t total = (uint64_t)P * cap;
	if (total) clear_kernel<<<cdiv(total, 256), 256, 0, stream>>>(slots, total);
	if (slots) YAKB_CUDA(cudaMemsetAsync((uint8_t*)slots - YAKB_SAT_BYTES, 0, YAKB_SAT_BYTES, stream)); // the flags are counter bits too
	YAKB_CUDA(cudaGetLastError());
	YAKB_CUDA(cudaStreamSynchronize(stream));
}

void Engine::hist(int64_t cnt[1024])
{
	unsigned long long *d = (unsigned long long*)b_tmp.need(1024 * 8);
	YAKB_CUDA(cudaMemsetAsync(d, 0, 1024 * 8, stream));
	const uint64_t total = (uint64_t)P * cap;
	if (total) hist_kernel<<<std::min<uint32_t>(cdiv(total, 256), 148 * 8), 256, 0, stream>>>(slots, total, cap, d);
	YAKB_CUDA(cudaGetLastError());
	YAKB_CUDA(cudaMemcpyAsync(cnt, d, 1024 * 8, cudaMemcpyDeviceToHost, stream));
	YAKB_CUDA(cudaStreamSynchronize(stream));
}

void Engine::get_batch(const uint64_t *d_x, uint64_t n, int32_t *d_out)
{
	if (n == 0) return;
	get_batch_kernel<<<cdiv(n, 256), 256, 0, stream>>>(d_x, n, pre, P - 1, own(), slots, cap, d_out);
	YAKB_CUDA(cudaGetLastError());
	YAKB_CUDA(cudaStreamSynchronize(stream));
}

// capacity khashl ends with after `presize`, m distinct puts and the optional trailing put
static uint32_t final_capacity(bool pflag, uint32_t pval, uint64_t m, bool trailing)
{
	uint64_t n = 0, count = 0;
	if (pflag) {
		uint32_t lg = 0, q = pval;
		while ((q >>= 1) != 0) ++lg;
		if (pval & (pval - 1)) ++lg;
		n = 1ull << (lg > 2 ? lg : 2);
	}
	// puts: a doubling happens whenever count reaches 3/4 n before a put
	while (count < m) {
		uint64_t thr = (n >> 1) + (n >> 2);
		if (count >= thr) { n = n ? n * 2 : 4; continue; }
		count = std::min<uint64_t>(m, thr);
	}
	if (trailing && count >= (n >> 1) + (n >> 2)) n = n ? n * 2 : 4;
	return (uint32_t)n;
}

// Rebuild the khashl layout of sub-tables [s0, s1) batch by batch (scratch-bounded).  For every batch
// `fn(b0, ns, voff, d_voff, d_dense, cap, size)` sees: the first sub-table (relative to s0), their
// number, and a dense device array holding each sub-table's stored keys in slot order (run t =
// [voff[t], voff[t+1]), counts attached when asked), plus the khashl capacity / size of each.
template<class F> void Engine::layout_batches(int s0, int s1, bool with_counts, uint64_t reserve_bytes, F &&fn)
{
	const int nsub = s1 - s0;
	std::vector<uint32_t> h_nkeys(nsub);
	std::vector<uint64_t> h_lp(nsub), h_ln(nsub);
	YAKB_CUDA(cudaMemcpyAsync(h_nkeys.data(), nkeys + s0, nsub * 4, cudaMemcpyDeviceToHost, stream));
	YAKB_CUDA(cudaMemcpyAsync(h_lp.data(), last_put + s0, nsub * 8, cudaMemcpyDeviceToHost, stream));
	YAKB_CUDA(cudaMemcpyAsync(h_ln.data(), last_new + s0, nsub * 8, cudaMemcpyDeviceToHost, stream));
	// the journal as arrays of segment pointers
	std::vector<const uint64_t*> hk, ho;
	for (auto &seg : journal) { hk.push_back(seg.keys); ho.push_back(seg.off); }
	const uint64_t **d_segk = (const uint64_t**)b_segp.need(std::max<size_t>(hk.size(), 1) * 2 * sizeof(void*));
	const uint64_t **d_sego = d_segk + std::max<size_t>(hk.size(), 1);
	if (!hk.empty()) {
		YAKB_CUDA(cudaMemcpyAsync(d_segk, hk.data(), hk.size() * sizeof(void*), cudaMemcpyHostToDevice, stream));
		YAKB_CUDA(cudaMemcpyAsync(d_sego, ho.data(), ho.size() * sizeof(void*), cudaMemcpyHostToDevice, stream));
	}
	YAKB_CUDA(cudaStreamSynchronize(stream));
	SegView sv; sv.keys = d_segk; sv.off = d_sego; sv.n = (int)hk.size();
	// scratch budget: most of what is free now (large sub-tables replay one thread each, so the more
	// of them run at once the better), but never less than 2 GB
	size_t mem_free = 0, mem_total = 0;
	cudaMemGetInfo(&mem_free, &mem_total);
	mem_free += dev_pool_idle(); // idle pool blocks are ours to take
	mem_free += b_lay[0].cap + b_lay[1].cap + b_lay[2].cap + b_lay[3].cap + b_lay[7].cap; // our own grow-only scratch is reusable
	uint64_t budget = mem_free > reserve_bytes ? (uint64_t)((mem_free - reserve_bytes) * 0.8) : 0;
	budget = std::max<uint64_t>(budget, 2ull << 30);
	const char *env_sm = getenv("YAKB_LAYOUT_SMEM_MAX"), *env_wp = getenv("YAKB_LAYOUT_WARP"); // test knobs
	const uint32_t smem_max = env_sm ? (uint32_t)atoi(env_sm) : 16384;
	const bool use_warp = env_wp ? atoi(env_wp) != 0 : true;
	int b0 = 0;
	while (b0 < nsub) {
		std::vector<uint64_t> ooff(1, 0), boff(1, 0), catoff(1, 0);
		std::vector<uint8_t> trail, pf, kind;
		std::vector<uint32_t> pvv, capb;
		int b1 = b0;
		uint64_t bytes = 0;
		while (b1 < nsub) {
			const int s = s0 + b1;
			const bool tr = h_lp[b1] > h_ln[b1];
			uint32_t capf = final_capacity(presize_flag[s], presize_val[s], h_nkeys[b1], tr);
			if (nops[s]) { // scratch bound when operations sit in the journal: never below any explicit request
				uint64_t bound = std::max<uint64_t>(final_capacity(presize_flag[s], presize_val[s], h_nkeys[b1], true), 4);
				uint64_t want = std::max<uint64_t>((uint64_t)max_req[s], 3ull * h_nkeys[b1] + 4);
				while (bound < want) bound *= 2;
				capf = (uint32_t)std::min<uint64_t>(bound * 2, 0x80000000ull);
			}
			const bool small = capf <= smem_max; // replayed in shared memory: its region only receives the result
			const bool warp = !small && use_warp && nops[s] == 0; // large, puts only: 32 lanes per sub-table
			const uint64_t region = small ? std::max<uint64_t>(h_nkeys[b1], 1) : std::max<uint32_t>(capf, 4);
			const uint64_t bw = small ? 2 : 2 * (uint64_t)(capf < 32 ? 1 : capf >> 5);
			const uint64_t add = region * 8 + bw * 4 + (uint64_t)h_nkeys[b1] * 8 + (warp ? region * 4 + (uint64_t)h_nkeys[b1] * 8 : 0);
			if (b1 > b0 && bytes + add > budget) break;
			bytes += add;
			ooff.push_back(ooff.back() + region);
			boff.push_back(boff.back() + std::max<uint64_t>(bw, 2));
			trail.push_back(tr); pf.push_back(presize_flag[s]); pvv.push_back(presize_val[s]); capb.push_back(capf);
			kind.push_back(small ? 0 : warp ? 2 : 1);
			catoff.push_back(catoff.back() + (warp ? h_nkeys[b1] : 0));
			++b1;
		}
		const int ns = b1 - b0;
		uint64_t *d_out = b_lay[1].as<uint64_t>(std::max<uint64_t>(ooff.back(), 1));
		uint32_t *d_bm = b_lay[3].as<uint32_t>(std::max<uint64_t>(boff.back(), 1));
		uint64_t *d_ooff = b_lay[4].as<uint64_t>(ns + 1), *d_boff = b_lay[6].as<uint64_t>(ns + 1), *d_voff = b_lay[5].as<uint64_t>(ns + 1);
		uint32_t *d_pv = b_lay[8].as<uint32_t>(ns), *d_ocap = b_lay[9].as<uint32_t>(ns), *d_osize = b_lay[10].as<uint32_t>(ns);
		uint8_t *d_trail = b_lay[11].as<uint8_t>(2 * (size_t)ns), *d_pf = d_trail + ns;
		YAKB_CUDA(cudaMemcpyAsync(d_ooff, ooff.data(), (ns + 1) * 8, cudaMemcpyHostToDevice, stream));
		YAKB_CUDA(cudaMemcpyAsync(d_boff, boff.data(), (ns + 1) * 8, cudaMemcpyHostToDevice, stream));
		YAKB_CUDA(cudaMemcpyAsync(d_pv, pvv.data(), ns * 4, cudaMemcpyHostToDevice, stream));
		YAKB_CUDA(cudaMemcpyAsync(d_trail, trail.data(), ns, cudaMemcpyHostToDevice, stream));
		YAKB_CUDA(cudaMemcpyAsync(d_pf, pf.data(), ns, cudaMemcpyHostToDevice, stream));
		{ // small tables replay in shared memory (three size classes), the rest in global scratch
			const uint32_t cls[3] = {1024, 4096, 16384};
			std::vector<int> lists[5]; // 0-2 shared-memory classes, 3 one thread each, 4 one warp each
			for (int t = 0; t < ns; ++t) {
				int k = kind[t] == 2 ? 4 : 3;
				if (kind[t] == 0) for (int i = 2; i >= 0; --i) if (capb[t] <= cls[i]) k = i;
				lists[k].push_back(t);
			}
			std::vector<int> all;
			size_t start[6] = {0, 0, 0, 0, 0, 0};
			for (int k = 0; k < 5; ++k) { start[k] = all.size(); all.insert(all.end(), lists[k].begin(), lists[k].end()); }
			int *d_list = (int*)b_misc.need(std::max<size_t>(all.size(), 1) * sizeof(int));
			YAKB_CUDA(cudaMemcpyAsync(d_list, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
			for (int k = 0; k < 3; ++k) {
				if (lists[k].empty()) continue;
				ProfScope ps("layout(smem replay)", stream);
				const size_t sm = (size_t)cls[k] * 8 + (size_t)(cls[k] < 32 ? 1 : cls[k] >> 5) * 8 + 512 * 8;
				YAKB_CUDA(cudaFuncSetAttribute(build_layout_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
				build_layout_smem_kernel<<<(uint32_t)lists[k].size(), 32, sm, stream>>>(d_list + start[k], cls[k], sv, s0 + b0, d_pf, d_pv, d_trail,
				                                                                          d_out, d_ooff, d_ocap, d_osize);
			}
			if (!lists[3].empty()) {
				ProfScope ps("layout(thread replay)", stream);
				build_layout_kernel<<<cdiv(lists[3].size(), 32), 32, 0, stream>>>(d_list + start[3], (int)lists[3].size(), sv, s0 + b0, d_pf, d_pv, d_trail,
				                                                                  d_out, d_ooff, d_bm, d_boff, d_ocap, d_osize);
			}
			if (!lists[4].empty()) {
				uint64_t *d_cat = b_lay[0].as<uint64_t>(std::max<uint64_t>(catoff.back(), 1));
				uint32_t *d_own = b_lay[7].as<uint32_t>(std::max<uint64_t>(ooff.back(), 1));
				uint64_t *d_catoff = (uint64_t*)b_tmp.need((size_t)(2 * ns + 2) * 8), *d_run = d_catoff + ns + 1;
				YAKB_CUDA(cudaMemcpyAsync(d_catoff, catoff.data(), (ns + 1) * 8, cudaMemcpyHostToDevice, stream));
				YAKB_CUDA(cudaMemsetAsync(d_run, 0, (size_t)ns * 8, stream));
				{ ProfScope ps("layout(gather journal)", stream);
				for (auto &seg : journal)
					gather_seg_kernel<<<(uint32_t)lists[4].size(), 256, 0, stream>>>(seg.keys, seg.off, s0 + b0, d_list + start[4], (int)lists[4].size(), d_catoff, d_run, d_cat); }
				ProfScope ps("layout(warp replay)", stream);
				unsigned long long *d_clk = nullptr; // YAKB_LAYOUT_CLOCKS=1: cycles per phase of the warp replay, summed over sub-tables
				static const bool want_clk = getenv("YAKB_LAYOUT_CLOCKS") != nullptr;
				if (want_clk) { d_clk = (unsigned long long*)b_stats.need(8 * sizeof(unsigned long long)) + 4; YAKB_CUDA(cudaMemsetAsync(d_clk, 0, 32, stream)); }
				build_layout_warp_kernel<<<(uint32_t)lists[4].size(), 32, 0, stream>>>(d_list + start[4], d_cat, d_catoff, d_pf, d_pv, d_trail,
				                                                                       d_out, d_ooff, d_bm, d_boff, d_own, d_ocap, d_osize, d_clk);
				if (want_clk) {
					unsigned long long h_clk[4];
					YAKB_CUDA(cudaMemcpyAsync(h_clk, d_clk, 32, cudaMemcpyDeviceToHost, stream));
					YAKB_CUDA(cudaStreamSynchronize(stream));
					fprintf(stderr, "[T::layout warp replay] %zu sub-tables, cycles per sub-table: rehash %.0f, init %.0f, put %.0f, materialise %.0f\n", lists[4].size(),
					        (double)h_clk[0] / lists[4].size(), (double)h_clk[1] / lists[4].size(), (double)h_clk[2] / lists[4].size(), (double)h_clk[3] / lists[4].size());
				}
			}
			YAKB_CUDA(cudaGetLastError());
			YAKB_CUDA(cudaStreamSynchronize(stream)); // `all` must outlive the copy
		}
		std::vector<uint32_t> h_cap(ns), h_size(ns);
		YAKB_CUDA(cudaMemcpyAsync(h_cap.data(), d_ocap, ns * 4, cudaMemcpyDeviceToHost, stream));
		YAKB_CUDA(cudaMemcpyAsync(h_size.data(), d_osize, ns * 4, cudaMemcpyDeviceToHost, stream));
		YAKB_CUDA(cudaStreamSynchronize(stream));
		std::vector<uint64_t> voff(ns + 1, 0);
		for (int t = 0; t < ns; ++t) {
			if (h_size[t] > ooff[t + 1] - ooff[t]) throw CudaError("[yakb] layout: inconsistent journal");
			voff[t + 1] = voff[t] + h_size[t];
		}
		const uint64_t nv = voff.back();
		uint64_t *d_dense = b_lay[2].as<uint64_t>(std::max<uint64_t>(nv, 1));
		YAKB_CUDA(cudaMemcpyAsync(d_voff, voff.data(), (ns + 1) * 8, cudaMemcpyHostToDevice, stream));
		if (nv) {
			ProfScope ps("layout(densify+counts)", stream);
			densify_kernel<<<cdiv(nv, 256), 256, 0, stream>>>(d_out, d_ooff, d_voff, ns, nv, d_dense);
			if (with_counts && cap) fill_counts_kernel<<<cdiv(nv, 256), 256, 0, stream>>>(d_dense, d_voff, ns, s0 + b0, nv, slots, cap);
		}
		YAKB_CUDA(cudaGetLastError());
		fn(b0, ns, voff, d_voff, d_dense, h_cap, h_size);
		YAKB_CUDA(cudaStreamSynchronize(stream)); // voff is read by the copy above until here
		b0 = b1;
	}
}

void Engine::layout(int s0, int s1, LayoutOut &out, bool with_counts)
{
	const int nsub = s1 - s0;
	out.cap.assign(nsub, 0); out.size.assign(nsub, 0); out.off.assign(nsub + 1, 0); out.keys.clear();
	layout_batches(s0, s1, with_counts, 0, [&](int b0, int ns, const std::vector<uint64_t> &voff, const uint64_t *, const uint64_t *d_dense,
	                                             const std::vector<uint32_t> &h_cap, const std::vector<uint32_t> &h_size) {
		const uint64_t nv = voff.back(), base = out.keys.size();
		out.keys.resize(base + nv);
		if (nv) YAKB_CUDA(cudaMemcpyAsync(out.keys.data() + base, d_dense, nv * 8, cudaMemcpyDeviceToHost, stream));
		YAKB_CUDA(cudaStreamSynchronize(stream));
		for (int t = 0; t < ns; ++t) {
			out.cap[b0 + t] = h_cap[t]; out.size[b0 + t] = h_size[t];
			out.off[b0 + t + 1] = base + voff[t + 1];
		}
	});
}

void Engine::layout_device(int s0, int s1, bool with_counts, const LayoutFn &fn)
{
	layout_batches(s0, s1, with_counts, 0, [&](int b0, int ns, const std::vector<uint64_t> &voff, const uint64_t *, const uint64_t *d_dense,
	                                             const std::vector<uint32_t> &h_cap, const std::vector<uint32_t> &h_size) {
		YAKB_CUDA(cudaStreamSynchronize(stream)); // d_dense is complete
		fn(b0, ns, voff, d_dense, h_cap, h_size);
	});
}

void Engine::load_subtables(const std::vector<uint32_t> &caps, const std::vector<uint64_t> &off, const uint64_t *keys, bool keys_on_device)
{
	const uint64_t n = off[P];
	uint32_t mx = 0;
	std::vector<uint32_t> cnt(P);
	for (int s = 0; s < P; ++s) { cnt[s] = (uint32_t)(off[s + 1] - off[s]); mx = std::max(mx, cnt[s]); presize_flag[s] = 1; presize_val[s] = caps[s]; nops[s] = 0; max_req[s] = caps[s]; }
	reserve(std::max<uint32_t>(mx, 8));
	Segment seg;
	seg.n = n;
	seg.keys = (uint64_t*)journal_alloc(std::max<uint64_t>(n, 1) * 8);
	seg.off = (uint64_t*)journal_alloc((uint64_t)(P + 1) * 8);
	YAKB_CUDA(cudaMemcpyAsync(seg.off, off.data(), (uint64_t)(P + 1) * 8, cudaMemcpyHostToDevice, stream));
	if (n) {
		ProfScope ps("load(bulk insert)", stream);
		YAKB_CUDA(cudaMemcpyAsync(seg.keys, keys, n * 8, keys_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
		bulk_insert_kernel<<<cdiv(n, 256), 256, 0, stream>>>(seg.keys, seg.off, P, n, slots, cap);
		strip_counts_kernel<<<cdiv(n, 256), 256, 0, stream>>>(seg.keys, n); // journal keeps keys without counts
		YAKB_CUDA(cudaGetLastError());
	}
	YAKB_CUDA(cudaMemcpyAsync(nkeys, cnt.data(), P * 4, cudaMemcpyHostToDevice, stream));
	YAKB_CUDA(cudaStreamSynchronize(stream));
	journal.push_back(seg);
	tot = n;
}

// flag[i] = 1 if key i of a dense layout batch has its count in [lo, hi]; kept[t] counts them per sub-table
__global__ void shrink_flag_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ voff, int ns, uint64_t n,
                                   uint32_t lo_c, uint32_t hi_c, uint8_t *__restrict__ flag, uint32_t *kept)
{
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t lo = 0, hi = ns;
	while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (voff[mid] <= i) lo = mid; else hi = mid; }
	const uint32_t c = (uint32_t)keys[i] & YAKB_MAX_COUNT;
	const bool keep = c >= lo_c && c <= hi_c;
	flag[i] = keep;
	// one atomic per (warp, sub-table): neighbours almost always share the sub-table, and three billion atomics on a few
	// thousand addresses take seconds
	const uint32_t am = __activemask(), peers = __match_any_sync(am, lo);
	const uint32_t votes = __ballot_sync(am, keep) & peers;
	if ((threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1) && votes) atomicAdd(&kept[lo], (uint32_t)__popc(votes));
}

void Engine::shrink(int min, int max)
{
	if (!(max >= min && max <= 1023)) max = 1023;
	if (min < 0) min = 0;
	// htab.c:183-193: old slots upward, keep min<=count<=max, into a set pre-sized to the OLD size.
	// Layout, filter and compaction stay on the device; only per-sub-table counts come back.
	std::vector<uint64_t> off(P + 1, 0);
	std::vector<uint32_t> caps(P, 0);
	uint64_t total = 0;
	{ std::vector<uint32_t> z; sizes(z); for (uint32_t v : z) total += v; }
	uint64_t *d_kept = b_newv.as<uint64_t>(std::max<uint64_t>(total, 1)); // every batch appends here
	uint64_t n_kept = 0;
	layout_batches(0, P, true, 0, [&](int b0, int ns, const std::vector<uint64_t> &voff, const uint64_t *d_voff, const uint64_t *d_dense,
	                                   const std::vector<uint32_t> &, const std::vector<uint32_t> &h_size) {
		const uint64_t nv = voff.back();
		uint8_t *flag = b_pflag.as<uint8_t>(std::max<uint64_t>(nv, 1));
		uint32_t *d_cnt = b_pend.as<uint32_t>(ns + 1);
		YAKB_CUDA(cudaMemsetAsync(d_cnt, 0, (ns + 1) * 4, stream));
		std::vector<uint32_t> h_cnt(ns, 0);
		if (nv) {
			ProfScope ps("shrink(filter)", stream);
			shrink_flag_kernel<<<cdiv(nv, 256), 256, 0, stream>>>(d_dense, d_voff, ns, nv, (uint32_t)min, (uint32_t)max, flag, d_cnt);
			compact_flagged_u64(d_dense, flag, nv, d_kept + n_kept, d_cnt + ns, stream, rs);
			YAKB_CUDA(cudaGetLastError());
		}
		YAKB_CUDA(cudaMemcpyAsync(h_cnt.data(), d_cnt, ns * 4, cudaMemcpyDeviceToHost, stream));
		YAKB_CUDA(cudaStreamSynchronize(stream));
		for (int t = 0; t < ns; ++t) {
			n_kept += h_cnt[t];
			off[b0 + t + 1] = n_kept;
			caps[b0 + t] = h_size[t]; // yak_ht_resize(f, kh_size(g))
		}
	});
	rebuild_dev(caps, off, d_kept);
}

// empty the count table for a rebuild; the allocation is kept when it is large enough for the new content
// (freeing and re-allocating tens of gigabytes costs more than the rebuild itself)
void Engine::reset_table(const std::vector<uint64_t> &off)
{
	uint64_t mx = 8;
	for (int s = 0; s < P; ++s) mx = std::max<uint64_t>(mx, off[s + 1] - off[s]);
	const uint64_t want = ((uint64_t)(mx / load_limit) + 16 + 3) & ~3ull;
	if (slots && cap >= want) {
		YAKB_CUDA(cudaMemsetAsync(slots, 0xFF, (uint64_t)P * cap * 8, stream));
		YAKB_CUDA(cudaMemsetAsync((uint8_t*)slots - YAKB_SAT_BYTES, 0, YAKB_SAT_BYTES, stream));
	} else if (slots) { free_slots(); cap = 0; }
}

void Engine::rebuild(const std::vector<uint32_t> &caps, const std::vector<uint64_t> &off, const uint64_t *keys)
{
	// drop the old table and journal, start again from the given keys
	journal_free_all();
	reset_table(off);
	YAKB_CUDA(cudaMemsetAsync(last_put, 0, P * 8, stream));
	YAKB_CUDA(cudaMemsetAsync(last_new, 0, P * 8, stream));
	load_subtables(caps, off, keys);
}

// same as rebuild() with the keys already on the device (d_keys may live in a scratch buffer)
void Engine::rebuild_dev(const std::vector<uint32_t> &caps, const std::vector<uint64_t> &off, const uint64_t *d_keys)
{
	journal_free_all();
	reset_table(off);
	YAKB_CUDA(cudaMemsetAsync(last_put, 0, P * 8, stream));
	YAKB_CUDA(cudaMemsetAsync(last_new, 0, P * 8, stream));
	load_subtables(caps, off, d_keys, true);
}

void Engine::sizes(std::vector<uint32_t> &out)
{
	out.resize(P);
	YAKB_CUDA(cudaMemcpyAsync(out.data(), nkeys, P * 4, cudaMemcpyDeviceToHost, stream));
	YAKB_CUDA(cudaStreamSynchronize(stream));
}

void Engine::append_ops(const std::vector<uint64_t> &op)
{
	// a pending "put of an existing key after the last new key" must be replayed before the operation
	std::vector<uint64_t> h_lp(P), h_ln(P), ent, off(P + 1, 0);
	YAKB_CUDA(cudaMemcpyAsync(h_lp.data(), last_put, P * 8, cudaMemcpyDeviceToHost, stream));
	YAKB_CUDA(cudaMemcpyAsync(h_ln.data(), last_new, P * 8, cudaMemcpyDeviceToHost, stream));
	YAKB_CUDA(cudaStreamSynchronize(stream));
	for (int s = 0; s < P; ++s) {
		if (op[s]) {
			if (h_lp[s] > h_ln[s]) { ent.push_back(OP_CHECK); ++nops[s]; }
			ent.push_back(op[s]); ++nops[s];
			const uint32_t kind = (uint32_t)op[s] & 1023;
			if (kind == OP_RESIZE || kind == OP_RESIZE_IF_LARGER) max_req[s] = std::max<uint32_t>(max_req[s], (uint32_t)(op[s] >> 10));
		}
		off[s + 1] = ent.size();
	}
	if (ent.empty()) return;
	Segment seg;
	seg.n = ent.size();
	seg.keys = (uint64_t*)journal_alloc(seg.n * 8);
	seg.off = (uint64_t*)journal_alloc((uint64_t)(P + 1) * 8);
	YAKB_CUDA(cudaMemcpyAsync(seg.keys, ent.data(), seg.n * 8, cudaMemcpyHostToDevice, stream));
	YAKB_CUDA(cudaMemcpyAsync(seg.off, off.data(), (uint64_t)(P + 1) * 8, cudaMemcpyHostToDevice, stream));
	// the replayed check clears the trailing state of the sub-tables that got an operation
	for (int s = 0; s < P; ++s) if (op[s]) h_lp[s] = h_ln[s];
	YAKB_CUDA(cudaMemcpyAsync(last_put, h_lp.data(), P * 8, cudaMemcpyHostToDevice, stream));
	YAKB_CUDA(cudaStreamSynchronize(stream));
	journal.push_back(seg);
}

void qv_scan_ascii(Engine *e, const uint8_t *d_asc, uint64_t n, int16_t *d_cnt, int raw)
{
	if (n == 0) return;
	const uint64_t nwords = (n + 31) / 32;
	uint64_t *w2 = e->b_w2.as<uint64_t>(packed_words(nwords)) + YAKB_PADW;
	uint32_t *wm = e->b_wm.as<uint32_t>(packed_words(nwords)) + YAKB_PADW;
	pack_ascii_kernel<<<cdiv(packed_npad(nwords) + YAKB_PADW, 256), 256, 0, e->stream>>>(d_asc, n, w2, wm, nwords, packed_npad(nwords));
	if (e->k >= 32) qv_scan_kernel<true><<<cdiv(nwords, 256), 256, 0, e->stream>>>(w2, wm, nwords, n, e->k, e->pre, e->P - 1, e->own(), e->slots, e->cap, d_cnt, raw);
	else qv_scan_kernel<false><<<cdiv(nwords, 256), 256, 0, e->stream>>>(w2, wm, nwords, n, e->k, e->pre, e->P - 1, e->own(), e->slots, e->cap, d_cnt, raw);
	YAKB_CUDA(cudaGetLastError());
	YAKB_CUDA(cudaStreamSynchronize(e->stream));
}

} // namespace yakb
