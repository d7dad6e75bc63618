// capi.cu - the C ABI of libyakb200.so: the reference's yak.h entry points (include/yak.h) and
// the chunk-level extensions (include/yak_b200.h), over the device engine.
//
// A yak_ch_t handed out here is the head of a larger host object; `h[i].h` points at a small
// per-sub-table handle (the reference keeps a khashl set there) and `h[i].b`, when a bloom
// filter exists, at a yak_bf_t whose `b` is the DEVICE address of that sub-filter.
#include "../../include/yak.h"
#include "../../include/yak_b200.h"
#include "engine.cuh"
#include "extras.cuh"
#include "ingest.cuh"
#include "fastx.h"
#include "fastx_par.h"
#include "yakfile.h"
#include "textcache.h"
#include "yakb_dev.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <assert.h>
#include <math.h>
#include <algorithm>
#include <map>
#include <tuple>
#include <mutex>
#include <sys/stat.h>
#include <sys/mman.h>
#include <unistd.h>
#include <fcntl.h>
#include <thread>
#include <atomic>
#include <condition_variable>
#include "ref_flow.h"
#include <chrono>
#include <memory>
#include <sys/resource.h>
#include <sys/time.h>

using namespace yakb;

struct yak_ht_t { struct ChBox *owner; int idx; };

// One GPU's part of a table that yak_count spread over several GPUs of this process: shard r lives on
// device `dev` and owns sub-tables [r*P/G, (r+1)*P/G).  Its buffers belong to that device.
struct Shard {
	int dev = 0;
	Engine *eng = nullptr;
	RouteScratch route;
	DBuf d_in, d_ev, d_recv;
	cudaStream_t stream = nullptr;      // host->device copy + extraction of this shard's slice of a batch
};

struct ChBox {
	yak_ch_t pub;          // must stay first: yak_ch_t* <-> ChBox*
	uint32_t magic;
	Engine *eng;           // the table; shard 0 of a multi-GPU table
	std::vector<Shard*> shards; // empty for a table on one GPU
	int dev = 0;           // device of `eng`
	std::vector<yak_ht_t> handles;
	std::vector<yak_bf_t> filters;
	std::mutex mu;
	DBuf d_in, d_in2, d_aux, d_aux2;
	IngestScratch ing;                  // the device-side text ingest of yak_count (csrc/ingest.cu)
	cudaStream_t copy_stream = nullptr; // host->device copies of yak_count batches, overlapping the kernels
};
static const uint32_t kMagic = 0x59414B42; // "YAKB"

static ChBox *box_of(const yak_ch_t *h)
{
	ChBox *b = (ChBox*)h;
	if (b == nullptr || b->magic != kMagic) { fprintf(stderr, "[yakb] ERROR: yak_ch_t was not created by this library\n"); abort(); }
	return b;
}
// for the entry points that work on one-GPU tables only: a multi-GPU table (yak_count with YAKB_GPUS > 1) supports the
// `yak count` flow - yak_count, yak_ch_destroy_bf / clear / shrink / hist / get / insert_list / dump / destroy
static ChBox *box_single(const yak_ch_t *h, const char *func)
{
	ChBox *b = box_of(h);
	if (!b->shards.empty()) throw CudaError(std::string(func) + " is not available on a multi-GPU table (YAKB_GPUS > 1): dump it and restore it on one GPU");
	return b;
}
struct DevGuard {
	int prev = 0;
	explicit DevGuard(int d) { cudaGetDevice(&prev); if (d != prev) cudaSetDevice(d); }
	~DevGuard() { int cur = 0; cudaGetDevice(&cur); if (cur != prev) cudaSetDevice(prev); }
};
// fn(engine, rank) for the table's engine, or for every shard on its own device (side by side when `parallel`)
template<class F> static void each_shard(ChBox *b, bool parallel, F &&fn)
{
	if (b->shards.empty()) { fn(b->eng, 0); return; }
	const int G = (int)b->shards.size();
	if (!parallel) {
		for (int r = 0; r < G; ++r) { DevGuard g(b->shards[r]->dev); fn(b->shards[r]->eng, r); }
		return;
	}
	std::vector<std::thread> th;
	std::vector<std::string> err(G);
	for (int r = 0; r < G; ++r)
		th.emplace_back([&, r] {
			try { cudaSetDevice(b->shards[r]->dev); fn(b->shards[r]->eng, r); }
			catch (const std::exception &e) { err[r] = e.what(); }
		});
	for (auto &t : th) t.join();
	for (auto &e : err) if (!e.empty()) throw CudaError(e);
}

#define GUARD_BEGIN try {
#define GUARD_END(ret) } catch (const std::exception &e) { fprintf(stderr, "[yakb] ERROR: %s\n", e.what()); return ret; }
#define GUARD_END_VOID } catch (const std::exception &e) { fprintf(stderr, "[yakb] FATAL: %s\n", e.what()); abort(); }

// ------------------------------------------------------------------ misc.c / sys.c equivalents

extern "C" {

int yak_verbose = 3;

unsigned char seq_nt4_table[256];
static struct Nt4Init { Nt4Init() { for (int c = 0; c < 256; ++c) seq_nt4_table[c] = (unsigned char)nt4((uint32_t)c); } } g_nt4_init;

void yak_copt_init(yak_copt_t *o) // misc.c:23-32
{
	memset(o, 0, sizeof(*o));
	o->bf_shift = 0; o->bf_n_hash = 4; o->k = 31; o->pre = 10; o->n_thread = 4; o->chunk_size = 10000000;
}

void yak_qopt_init(yak_qopt_t *o) // qv.c:137-144
{
	memset(o, 0, sizeof(*o));
	o->chunk_size = 1000000000; o->n_threads = 4; o->min_frac = 0.5; o->fpr = 0.00004;
}

} // extern "C"

static double wall_now() { struct timeval tv; gettimeofday(&tv, nullptr); return tv.tv_sec + tv.tv_usec * 1e-6; }
static double g_t0 = wall_now();
static bool timing_on() { static int v = -1; if (v < 0) { const char *e = getenv("YAKB_TIMING"); v = e && atoi(e) > 0; } return v != 0; }
struct StageTimer {
	const char *name; double t0;
	StageTimer(const char *n) : name(n), t0(wall_now()) {}
	~StageTimer() { if (timing_on()) fprintf(stderr, "[T::%s] %.3f s\n", name, wall_now() - t0); }
};
static double cpu_now() { struct rusage r; getrusage(RUSAGE_SELF, &r); return r.ru_utime.tv_sec + r.ru_stime.tv_sec + 1e-6 * (r.ru_utime.tv_usec + r.ru_stime.tv_usec); }

// ------------------------------------------------------------------ yak_ch_*

static void attach_handles(ChBox *b)
{
	const int P = 1 << b->pub.pre;
	b->handles.resize(P);
	b->filters.clear();
	if (b->eng->bloom) b->filters.resize(P);
	b->pub.h = (yak_ch1_t*)calloc(P, sizeof(yak_ch1_t));
	for (int i = 0; i < P; ++i) {
		b->handles[i].owner = b; b->handles[i].idx = i;
		b->pub.h[i].h = &b->handles[i];
		if (b->eng->bloom) {
			const Engine *e = b->shards.empty() ? b->eng : b->shards[i / b->eng->P]->eng; // the engine that holds sub-table i
			b->filters[i].n_shift = e->n_shift - e->pre;
			b->filters[i].n_hashes = e->n_hash;
			b->filters[i].b = e->bloom + (size_t)(i & (e->P - 1)) * 64; // first block of sub-filter i (blocks interleave by sub-table)
			b->pub.h[i].b = &b->filters[i];
		}
	}
}

static yak_ch_t *ch_init_shard(int k, int pre, int n_hash, int n_shift, int rank, int world)
{
	GUARD_BEGIN
	StageTimer tm("yak_ch_init");
	if (pre < YAK_COUNTER_BITS) return 0;
	Engine *e = Engine::create(k, pre, n_hash, n_shift, rank, world);
	if (!e) return 0;
	ChBox *b = new ChBox;
	memset(&b->pub, 0, sizeof(b->pub));
	b->magic = kMagic; b->eng = e;
	cudaGetDevice(&b->dev);
	b->pub.k = k; b->pub.pre = pre; b->pub.n_hash = e->n_hash; b->pub.n_shift = e->n_shift; b->pub.tot = 0;
	attach_handles(b);
	return &b->pub;
	GUARD_END(0)
}

// how many GPUs yak_count spreads a NEW table over: YAKB_GPUS (default 1), a power of two, at most the devices present and 16
static int multi_gpus_wanted()
{
	const char *e = getenv("YAKB_GPUS");
	int g = e ? atoi(e) : 1, n = 0;
	if (g <= 1) return 1;
	if (cudaGetDeviceCount(&n) != cudaSuccess) return 1;
	while (g > n || (g & (g - 1))) --g;
	return g > 16 ? 16 : (g < 1 ? 1 : g);
}

// a table of G shards, shard r on device r (the devices CUDA_VISIBLE_DEVICES leaves visible)
static yak_ch_t *ch_init_multi(int k, int pre, int n_hash, int n_shift, int G)
{
	GUARD_BEGIN
	StageTimer tm("yak_ch_init(multi)");
	if (pre < YAK_COUNTER_BITS) return 0;
	ChBox *b = new ChBox;
	memset(&b->pub, 0, sizeof(b->pub));
	b->magic = kMagic; b->eng = nullptr;
	int dev0 = 0;
	cudaGetDevice(&dev0);
	std::atomic<bool> ok{true};
	for (int r = 0; r < G; ++r) { Shard *sh = new Shard; sh->dev = r; b->shards.push_back(sh); }
	{ // side by side: a 16 GiB filter takes the driver half a second to allocate, per GPU
		std::vector<std::thread> th;
		for (int r = 0; r < G; ++r)
			th.emplace_back([&, r] {
				Shard *sh = b->shards[r];
				try {
					cudaSetDevice(r);
					for (int q = 0; q < G; ++q) // peers read each other's routed events directly over NVLink; without peer access the copies are staged
						if (q != r) { int can = 0; cudaDeviceCanAccessPeer(&can, r, q); if (can && cudaDeviceEnablePeerAccess(q, 0) != cudaSuccess) (void)cudaGetLastError(); }
					sh->eng = Engine::create(k, pre, n_hash, n_shift, r, G);
					if (!sh->eng) { ok = false; return; }
					YAKB_CUDA(cudaStreamCreateWithFlags(&sh->stream, cudaStreamNonBlocking));
				} catch (const std::exception &e) { fprintf(stderr, "[yakb] ERROR: %s\n", e.what()); ok = false; }
			});
		for (auto &t : th) t.join();
	}
	cudaSetDevice(dev0);
	if (!ok.load()) {
		for (Shard *sh : b->shards) { DevGuard g(sh->dev); delete sh->eng; if (sh->stream) cudaStreamDestroy(sh->stream); delete sh; }
		delete b;
		return 0;
	}
	b->eng = b->shards[0]->eng; b->dev = b->shards[0]->dev;
	b->pub.k = k; b->pub.pre = pre; b->pub.n_hash = b->eng->n_hash; b->pub.n_shift = b->eng->n_shift; b->pub.tot = 0;
	attach_handles(b);
	fprintf(stderr, "[M::%s] table of 2^%d sub-tables on %d GPUs (%d sub-tables each)\n", __func__, pre, G, b->eng->P);
	return &b->pub;
	GUARD_END(0)
}

extern "C" yak_ch_t *yak_ch_init(int k, int pre, int n_hash, int n_shift) // htab.c:13-29
{
	return ch_init_shard(k, pre, n_hash, n_shift, 0, 1);
}

extern "C" yak_ch_t *yakb_ch_init_shard(int k, int pre, int n_hash, int n_shift, int rank, int world)
{
	return ch_init_shard(k, pre, n_hash, n_shift, rank, world);
}

extern "C" void yak_ch_destroy_bf(yak_ch_t *h) // htab.c:31-39
{
	ChBox *b = box_of(h);
	std::lock_guard<std::mutex> lk(b->mu);
	each_shard(b, false, [](Engine *e, int) { e->destroy_bloom(); });
	for (int i = 0; i < 1 << h->pre; ++i) h->h[i].b = 0;
	b->filters.clear();
}

extern "C" void yak_ch_destroy(yak_ch_t *h) // htab.c:41-49
{
	if (h == 0) return;
	StageTimer tm("yak_ch_destroy");
	ChBox *b = box_of(h);
	if (b->shards.empty()) delete b->eng;
	else
		for (Shard *sh : b->shards) {
			DevGuard g(sh->dev);
			delete sh->eng;
			sh->d_in.release(); sh->d_ev.release(); sh->d_recv.release();
			for (DBuf &d : sh->route.b) d.release();
			sh->route.rs.release();
			if (sh->stream) cudaStreamDestroy(sh->stream);
			delete sh;
		}
	free(b->pub.h);
	if (b->copy_stream) cudaStreamDestroy(b->copy_stream);
	b->d_in.release(); b->d_in2.release(); b->d_aux.release(); b->d_aux2.release();
	b->ing.release();
	b->magic = 0;
	delete b;
}

extern "C" int yak_ch_insert_list(yak_ch_t *h, int create_new, int n, const uint64_t *a) // htab.c:51-78
{
	GUARD_BEGIN
	if (n <= 0) return 0;
	ChBox *b = box_of(h);
	std::lock_guard<std::mutex> lk(b->mu);
	const int only = (int)(a[0] & (uint64_t)(b->eng->P - 1)); // region index; foreign shards are filtered by owner
	if (!b->shards.empty()) { // the shard that owns a[0]'s sub-table takes the list
		Shard *sh = b->shards[(a[0] & (((uint64_t)1 << h->pre) - 1)) / (uint64_t)b->eng->P];
		DevGuard g(sh->dev);
		uint64_t *d = sh->d_in.as<uint64_t>(n);
		YAKB_CUDA(cudaMemcpyAsync(d, a, (size_t)n * 8, cudaMemcpyHostToDevice, sh->eng->stream));
		return (int)sh->eng->count_events(d, n, create_new, only).n_new;
	}
	uint64_t *d = b->d_in.as<uint64_t>(n);
	YAKB_CUDA(cudaMemcpyAsync(d, a, (size_t)n * 8, cudaMemcpyHostToDevice, b->eng->stream));
	ChunkStats st = b->eng->count_events(d, n, create_new, only);
	return (int)st.n_new; // the caller adds this to h->tot (count.c:138)
	GUARD_END(0)
}

extern "C" int yakb_ch_get_batch(const yak_ch_t *h, uint64_t n, const uint64_t *x, int32_t *out)
{
	GUARD_BEGIN
	if (n == 0) return 0;
	ChBox *b = box_of(h);
	std::lock_guard<std::mutex> lk(b->mu);
	if (!b->shards.empty()) { // every shard answers for its own sub-tables (-1 elsewhere): the maximum is the table's answer
		std::vector<int32_t> part(n);
		for (uint64_t i = 0; i < n; ++i) out[i] = -1;
		for (Shard *sh : b->shards) {
			DevGuard g(sh->dev);
			uint64_t *dx = sh->d_in.as<uint64_t>(n);
			int32_t *dout = sh->d_recv.as<int32_t>(n);
			YAKB_CUDA(cudaMemcpyAsync(dx, x, n * 8, cudaMemcpyHostToDevice, sh->eng->stream));
			sh->eng->get_batch(dx, n, dout);
			YAKB_CUDA(cudaMemcpyAsync(part.data(), dout, n * 4, cudaMemcpyDeviceToHost, sh->eng->stream));
			YAKB_CUDA(cudaStreamSynchronize(sh->eng->stream));
			for (uint64_t i = 0; i < n; ++i) if (part[i] > out[i]) out[i] = part[i];
		}
		return 0;
	}
	uint64_t *dx = b->d_in.as<uint64_t>(n);
	int32_t *dout = b->d_aux.as<int32_t>(n);
	YAKB_CUDA(cudaMemcpyAsync(dx, x, n * 8, cudaMemcpyHostToDevice, b->eng->stream));
	b->eng->get_batch(dx, n, dout);
	YAKB_CUDA(cudaMemcpyAsync(out, dout, n * 4, cudaMemcpyDeviceToHost, b->eng->stream));
	YAKB_CUDA(cudaStreamSynchronize(b->eng->stream));
	return 0;
	GUARD_END(-1)
}

extern "C" int yakb_ch_get_batch_dev(const yak_ch_t *h, uint64_t n, const uint64_t *d_x, int32_t *d_out)
{
	GUARD_BEGIN
	ChBox *b = box_single(h, __func__);
	std::lock_guard<std::mutex> lk(b->mu);
	b->eng->get_batch(d_x, n, d_out);
	return 0;
	GUARD_END(-1)
}

extern "C" int yak_ch_get(const yak_ch_t *h, uint64_t x) // htab.c:93-100
{
	int32_t r = -1;
	if (yakb_ch_get_batch(h, 1, &x, &r) != 0) return -1;
	return r;
}

extern "C" int yak_ch_inc(yak_ch_t *h, uint64_t x) // htab.c:80-91
{
	GUARD_BEGIN
	ChBox *b = box_single(h, __func__);
	std::lock_guard<std::mutex> lk(b->mu);
	return inc_one(b->eng, x);
	GUARD_END(-1)
}

extern "C" void yak_ch_clear(yak_ch_t *h, int n_thread) // htab.c:116-130
{
	GUARD_BEGIN
	(void)n_thread;
	ChBox *b = box_of(h);
	std::lock_guard<std::mutex> lk(b->mu);
	each_shard(b, true, [](Engine *e, int) { e->clear(); });
	GUARD_END_VOID
}

extern "C" void yak_ch_hist(const yak_ch_t *h, int64_t cnt[YAK_N_COUNTS], int n_thread) // htab.c:156-169
{
	GUARD_BEGIN
	(void)n_thread;
	ChBox *b = box_of(h);
	std::lock_guard<std::mutex> lk(b->mu);
	if (b->shards.empty()) b->eng->hist(cnt);
	else {
		for (int i = 0; i < YAK_N_COUNTS; ++i) cnt[i] = 0;
		each_shard(b, false, [&](Engine *e, int) { int64_t part[YAK_N_COUNTS]; e->hist(part); for (int i = 0; i < YAK_N_COUNTS; ++i) cnt[i] += part[i]; });
	}
	GUARD_END_VOID
}

extern "C" void yak_ch_shrink(yak_ch_t *h, int min, int max, int n_thread) // htab.c:199-208
{
	GUARD_BEGIN
	(void)n_thread;
	StageTimer tm("yak_ch_shrink");
	ChBox *b = box_of(h);
	std::lock_guard<std::mutex> lk(b->mu);
	each_shard(b, true, [&](Engine *e, int) { e->shrink(min, max); });
	uint64_t tot = 0;
	each_shard(b, false, [&](Engine *e, int) { tot += e->tot; });
	h->tot = tot;
	GUARD_END_VOID
}

extern "C" void yak_ch_setcnt(yak_ch_t *h, int cnt, int n_thread) // htab.c:229-235
{
	GUARD_BEGIN
	(void)n_thread;
	assert(cnt >= 0 && cnt <= YAK_MAX_COUNT);
	ChBox *b = box_single(h, __func__);
	std::lock_guard<std::mutex> lk(b->mu);
	setcnt(b->eng, cnt);
	GUARD_END_VOID
}

extern "C" yak_knt_t *yak_ch_getseq(const yak_ch_t *h, int w, uint32_t *n) // htab.c:353-367
{
	GUARD_BEGIN
	assert(h->k < 32 && w < 1 << h->pre);
	ChBox *b = box_single(h, __func__);
	std::lock_guard<std::mutex> lk(b->mu);
	LayoutOut lo;
	b->eng->layout(w, w + 1, lo, true);
	*n = lo.size[0];
	yak_knt_t *a = (yak_knt_t*)calloc(*n ? *n : 1, sizeof(yak_knt_t));
	const uint64_t mask = (1ULL << h->k * 2) - 1;
	for (uint32_t i = 0; i < *n; ++i) {
		a[i].x = hash64_inv(lo.keys[i] >> YAK_COUNTER_BITS << h->pre | (uint64_t)w, mask);
		a[i].c = (int)(lo.keys[i] & YAK_MAX_COUNT);
	}
	return a;
	GUARD_END(0)
}

// htab.c:102-110: per sub-table `if (size*3 < capacity) resize(size*3)`; the condition needs the khashl
// capacity, so it is recorded as an operation and evaluated when the layout is replayed
extern "C" void yak_ch_tighten(yak_ch_t *h)
{
	GUARD_BEGIN
	ChBox *b = box_single(h, __func__);
	std::lock_guard<std::mutex> lk(b->mu);
	std::vector<uint64_t> op(b->eng->P, (uint64_t)Engine::OP_TIGHTEN);
	b->eng->append_ops(op);
	GUARD_END_VOID
}

// Sub-table ranges [s0, s1) whose key totals stay below `max_keys` (one sub-table at least): the slices in which
// a table is brought to the host.  A fixed 1024 sub-tables was the whole table at -p10, i.e. tens of gigabytes
// of host memory and more than 2^31 events in one count_events() call for human-size assemblies.
static std::vector<int> slice_bounds(Engine *e, uint64_t max_keys = 1ull << 28, int max_sub = 1024)
{
	std::vector<uint32_t> z;
	e->sizes(z);
	std::vector<int> b(1, 0);
	uint64_t acc = 0;
	for (int s = 0; s < e->P; ++s) {
		if (s > b.back() && (acc + z[s] > max_keys || s - b.back() >= max_sub)) { b.push_back(s); acc = 0; }
		acc += z[s];
	}
	if (e->P > b.back()) b.push_back(e->P);
	return b;
}

// all sub-tables of a table in slot order with counts, brought to the host in slices
template<class F> static void for_each_slice(Engine *e, F &&fn)
{
	const std::vector<int> b = slice_bounds(e);
	for (size_t i = 0; i + 1 < b.size(); ++i) {
		LayoutOut lo;
		e->layout(b[i], b[i + 1], lo, true);
		fn(b[i], b[i + 1], lo);
	}
}

// htab.c:241-285: every key of h1 (slot order, min <= count <= max) is put into h0 and bumps its
// counter by one; h1 is destroyed.  The puts run through the ordinary event path of h0 (file order =
// h1's slot order per sub-table), which also records first-put order for the layout.
extern "C" void yak_ch_merge(yak_ch_t *h0, yak_ch_t *h1, int min, int max, int n_thread, int pre_resize)
{
	GUARD_BEGIN
	(void)n_thread;
	ChBox *b0 = box_single(h0, __func__), *b1 = box_single(h1, __func__);
	assert(h0->k == h1->k && h0->pre == h1->pre && b0->eng->P == b1->eng->P);
	if (!(max >= min && max <= YAK_MAX_COUNT)) max = YAK_MAX_COUNT;
	{
		std::lock_guard<std::mutex> lk(b0->mu);
		Engine *e0 = b0->eng, *e1 = b1->eng;
		const int pre = h0->pre;
		const uint64_t own_hi = (uint64_t)e0->rank << (pre - e0->lw);
		if (pre_resize) { // htab.c:250-254
			std::vector<uint32_t> z0, z1;
			e0->sizes(z0); e1->sizes(z1);
			std::vector<uint64_t> op(e0->P);
			for (int s = 0; s < e0->P; ++s) op[s] = ((uint64_t)(((uint64_t)z0[s] + z1[s]) * 4 / 3 + 1) << 10) | Engine::OP_RESIZE_IF_LARGER;
			e0->append_ops(op);
		}
		for_each_slice(e1, [&](int s0, int s1, LayoutOut &lo) {
			std::vector<uint64_t> ev;
			ev.reserve(lo.keys.size());
			for (int s = s0; s < s1; ++s)
				for (uint64_t i = lo.off[s - s0]; i < lo.off[s - s0 + 1]; ++i) {
					const int c = (int)(lo.keys[i] & YAK_MAX_COUNT);
					if (c >= min && c <= max) ev.push_back((lo.keys[i] >> YAK_COUNTER_BITS) << pre | own_hi | (uint64_t)s);
				}
			if (ev.empty()) return;
			uint64_t *d = b0->d_in.as<uint64_t>(ev.size());
			YAKB_CUDA(cudaMemcpyAsync(d, ev.data(), ev.size() * 8, cudaMemcpyHostToDevice, e0->stream));
			e0->count_events(d, ev.size(), 1, -1, true);
		});
		std::vector<uint32_t> z;
		e0->sizes(z);
		uint64_t tot = 0;
		for (uint32_t v : z) tot += v;
		e0->tot = tot;
		h0->tot = tot; // htab.c:283-284
	}
	yak_ch_destroy(h1);
	GUARD_END_VOID
}

// htab.c:287-347: h0 keeps the keys (with their counts, in its own slot order) that are absent from
// (subtract) / present in (isec) h1, re-put into sets pre-sized to h0's old sizes
static void filter_by_membership(yak_ch_t *h0, const yak_ch_t *h1, bool keep_present)
{
	ChBox *b0 = box_single(h0, __func__), *b1 = box_single(h1, __func__);
	assert(h0->k == h1->k && h0->pre == h1->pre && b0->eng->P == b1->eng->P);
	std::lock_guard<std::mutex> lk(b0->mu);
	Engine *e0 = b0->eng, *e1 = b1->eng;
	const int pre = h0->pre, P = e0->P;
	const uint64_t own_hi = (uint64_t)e0->rank << (pre - e0->lw);
	std::vector<uint64_t> kept, off(P + 1, 0);
	std::vector<uint32_t> caps(P, 0);
	for_each_slice(e0, [&](int s0, int s1, LayoutOut &lo) {
		const uint64_t n = lo.keys.size();
		std::vector<uint64_t> q(n);
		std::vector<int32_t> r(n, -1);
		for (int s = s0; s < s1; ++s)
			for (uint64_t i = lo.off[s - s0]; i < lo.off[s - s0 + 1]; ++i) q[i] = (lo.keys[i] >> YAK_COUNTER_BITS) << pre | own_hi | (uint64_t)s;
		if (n) {
			uint64_t *dq = b1->d_in.as<uint64_t>(n);
			int32_t *dr = b1->d_aux.as<int32_t>(n);
			YAKB_CUDA(cudaMemcpyAsync(dq, q.data(), n * 8, cudaMemcpyHostToDevice, e1->stream));
			e1->get_batch(dq, n, dr);
			YAKB_CUDA(cudaMemcpyAsync(r.data(), dr, n * 4, cudaMemcpyDeviceToHost, e1->stream));
			YAKB_CUDA(cudaStreamSynchronize(e1->stream));
		}
		for (int s = s0; s < s1; ++s) {
			for (uint64_t i = lo.off[s - s0]; i < lo.off[s - s0 + 1]; ++i)
				if ((r[i] >= 0) == keep_present) kept.push_back(lo.keys[i]);
			off[s + 1] = kept.size();
			caps[s] = lo.size[s - s0]; // yak_ht_resize(f, kh_size(g0))
		}
	});
	e0->rebuild(caps, off, kept.data());
	h0->tot = e0->tot;
}
extern "C" void yak_ch_subtract(yak_ch_t *h0, const yak_ch_t *h1, int n_thread)
{
	GUARD_BEGIN
	(void)n_thread;
	filter_by_membership(h0, h1, false);
	GUARD_END_VOID
}
extern "C" void yak_ch_isec(yak_ch_t *h0, const yak_ch_t *h1, int n_thread)
{
	GUARD_BEGIN
	(void)n_thread;
	filter_by_membership(h0, h1, true);
	GUARD_END_VOID
}

// ------------------------------------------------------------------ dump / restore

// htab.c:373-394; sink(ptr,len) receives the bytes in order
template<class Sink> static void serialise_engine(ChBox *b, Engine *eng, Sink &&sink, bool header);
// a multi-GPU table is the rank-ordered concatenation of its shards (the header from the first)
template<class Sink> static void serialise(ChBox *b, Sink &&sink, bool header = true)
{
	if (b->shards.empty()) { serialise_engine(b, b->eng, sink, header); return; }
	for (size_t r = 0; r < b->shards.size(); ++r) { DevGuard g(b->shards[r]->dev); serialise_engine(b, b->shards[r]->eng, sink, header && r == 0); }
}
// Pinned staging buffers are process-wide and outlive a table: cudaMallocHost of two batch-sized
// buffers costs more than counting a small file, and yak_count is called once per pass.
namespace {
struct PinnedPool {
	std::mutex mu;
	std::vector<std::pair<uint8_t*, size_t>> idle;
	uint8_t *get(size_t bytes, size_t *cap)
	{
		std::lock_guard<std::mutex> lk(mu);
		int best = -1;
		for (int i = 0; i < (int)idle.size(); ++i)
			if (idle[i].second >= bytes && (best < 0 || idle[i].second < idle[best].second)) best = i;
		if (best >= 0) { uint8_t *p = idle[best].first; *cap = idle[best].second; idle.erase(idle.begin() + best); return p; }
		if (idle.size() >= 2) { cudaFreeHost(idle.back().first); idle.pop_back(); } // too small to be useful: do not hoard
		uint8_t *p = nullptr;
		YAKB_CUDA(cudaMallocHost(&p, bytes));
		*cap = bytes;
		return p;
	}
	void put(uint8_t *p, size_t cap) { if (p) { std::lock_guard<std::mutex> lk(mu); idle.push_back({p, cap}); } }
};
PinnedPool g_pinned;
}

// the keys of one layout batch, read front to back from the device through two page-locked pieces: while the sink (the file) takes
// piece i, piece i+1 is on its way - the image never exists as one host array (it was a pageable vector of gigabytes: page faults
// and a pageable copy cost more than the layout itself)
namespace {
struct KeyStream {
	static const uint64_t PIECE = 8ull << 20; // keys per piece (64 MB)
	const uint64_t *d; uint64_t n; cudaStream_t st;
	uint64_t *pin[2] = {nullptr, nullptr};
	size_t pcap[2] = {0, 0};
	cudaEvent_t ev[2];
	int64_t issued = -1;
	KeyStream(const uint64_t *d_, uint64_t n_, cudaStream_t st_) : d(d_), n(n_), st(st_)
	{
		for (int i = 0; i < 2; ++i) { pin[i] = (uint64_t*)g_pinned.get(PIECE * 8, &pcap[i]); YAKB_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming)); }
	}
	~KeyStream() { for (int i = 0; i < 2; ++i) { cudaEventSynchronize(ev[i]); cudaEventDestroy(ev[i]); g_pinned.put((uint8_t*)pin[i], pcap[i]); } }
	void issue(int64_t k)
	{
		const uint64_t a = (uint64_t)k * PIECE;
		if (a >= n) return;
		YAKB_CUDA(cudaMemcpyAsync(pin[k & 1], d + a, std::min(PIECE, n - a) * 8, cudaMemcpyDeviceToHost, st));
		YAKB_CUDA(cudaEventRecord(ev[k & 1], st));
	}
	// keys [a, a + *avail) on the host; accesses must go front to back
	const uint64_t *at(uint64_t a, uint64_t *avail)
	{
		const int64_t k = (int64_t)(a / PIECE);
		while (issued < k + 1) issue(++issued); // piece k and the one behind it (its buffer was consumed: we are past piece k-1)
		YAKB_CUDA(cudaEventSynchronize(ev[k & 1]));
		*avail = std::min(n, (uint64_t)(k + 1) * PIECE) - a;
		return pin[k & 1] + (a - (uint64_t)k * PIECE);
	}
};
}

template<class Sink> static void serialise_engine(ChBox *b, Engine *eng, Sink &&sink, bool header)
{
	const yak_ch_t *h = &b->pub;
	uint32_t t3[3] = {(uint32_t)h->k, (uint32_t)h->pre, YAK_COUNTER_BITS};
	if (header) {
		sink(YAK_MAGIC, 4);
		sink(t3, 12);
	}
	const double t0 = wall_now();
	double t_sink = 0;
	eng->layout_device(0, eng->P, true, [&](int, int ns, const std::vector<uint64_t> &voff, const uint64_t *d_dense,
	                                         const std::vector<uint32_t> &cap, const std::vector<uint32_t> &size) {
		const double tw = wall_now();
		KeyStream ks(d_dense, voff[ns], eng->stream);
		for (int t = 0; t < ns; ++t) { // htab.c:379-390: capacity, size, the keys in slot order
			uint32_t u[2] = {cap[t], size[t]};
			sink(u, 8);
			for (uint64_t a = voff[t]; a < voff[t + 1];) {
				uint64_t avail = 0;
				const uint64_t *p = ks.at(a, &avail);
				const uint64_t m = std::min(avail, voff[t + 1] - a);
				sink(p, (size_t)m * 8);
				a += m;
			}
		}
		t_sink += wall_now() - tw;
	});
	if (timing_on()) fprintf(stderr, "[T::serialise] %.3f s, of which streaming the image to the sink %.3f s\n", wall_now() - t0, t_sink);
}

// one engine's image written at `offset` of an open file, through a 16 MB buffer; returns the bytes written or -1
static int64_t dump_engine_at(ChBox *b, Engine *eng, bool header, int fd, uint64_t offset)
{
	std::vector<uint8_t> buf;
	buf.reserve(16u << 20);
	uint64_t at = offset;
	bool ok = true;
	auto flush = [&]() {
		size_t done = 0;
		while (ok && done < buf.size()) { const ssize_t w = pwrite(fd, buf.data() + done, buf.size() - done, (off_t)(at + done)); if (w <= 0) ok = false; else done += (size_t)w; }
		at += buf.size();
		buf.clear();
	};
	serialise_engine(b, eng, [&](const void *p, size_t n) {
		const uint8_t *q = (const uint8_t*)p;
		while (n) {
			const size_t m = std::min(n, (size_t)(16u << 20) - buf.size());
			buf.insert(buf.end(), q, q + m);
			q += m; n -= m;
			if (buf.size() >= (16u << 20)) flush();
		}
	}, header);
	flush();
	return ok ? (int64_t)(at - offset) : -1;
}
static int64_t engine_image_bytes(Engine *eng, bool header)
{
	std::vector<uint32_t> z;
	eng->sizes(z);
	int64_t n = header ? 16 : 0;
	for (uint32_t v : z) n += 8 + 8 * (int64_t)v;
	return n;
}

extern "C" int yak_ch_dump(const yak_ch_t *h, const char *fn)
{
	GUARD_BEGIN
	ChBox *b = box_of(h);
	std::lock_guard<std::mutex> lk(b->mu);
	StageTimer tm("yak_ch_dump");
	if (b->eng->lw && b->shards.empty()) { fprintf(stderr, "[yakb] ERROR: yak_ch_dump on one shard of a multi-GPU table; use yakb_ch_dump_shard_mem\n"); return -1; }
	if (!b->shards.empty() && strcmp(fn, "-")) { // a multi-GPU table into a file: every shard writes its own part, side by side
		const int fd = open(fn, O_WRONLY | O_CREAT | O_TRUNC, 0644);
		if (fd < 0) return -1;
		const int G = (int)b->shards.size();
		std::vector<uint64_t> off(G + 1, 0);
		for (int r = 0; r < G; ++r) { DevGuard g(b->shards[r]->dev); off[r + 1] = off[r] + (uint64_t)engine_image_bytes(b->shards[r]->eng, r == 0); }
		std::vector<int64_t> wrote(G, 0);
		each_shard(b, true, [&](Engine *e, int r) { wrote[r] = dump_engine_at(b, e, r == 0, fd, off[r]); });
		close(fd);
		for (int r = 0; r < G; ++r) if (wrote[r] != (int64_t)(off[r + 1] - off[r])) return -1;
		fprintf(stderr, "[M::%s] dumpped the hash table to file '%s'.\n", __func__, fn);
		return 0;
	}
	FILE *fp = strcmp(fn, "-") ? fopen(fn, "wb") : stdout;
	if (fp == 0) return -1;
	std::unique_ptr<char[]> iobuf(new char[1 << 20]); // per call: tables may be dumped from several threads at once
	setvbuf(fp, iobuf.get(), _IOFBF, 1 << 20);
	serialise(b, [&](const void *p, size_t n) { fwrite(p, 1, n, fp); });
	fprintf(stderr, "[M::%s] dumpped the hash table to file '%s'.\n", __func__, fn);
	if (fp != stdout) fclose(fp); else { fflush(fp); setvbuf(fp, nullptr, _IOLBF, 0); }
	return 0;
	GUARD_END(-1)
}

static int64_t dump_mem(const yak_ch_t *h, uint8_t **out, bool header)
{
	GUARD_BEGIN
	ChBox *b = box_of(h);
	std::lock_guard<std::mutex> lk(b->mu);
	std::vector<uint8_t> acc;
	serialise(b, [&](const void *p, size_t n) { acc.insert(acc.end(), (const uint8_t*)p, (const uint8_t*)p + n); }, header);
	*out = (uint8_t*)malloc(acc.size() ? acc.size() : 1);
	memcpy(*out, acc.data(), acc.size());
	return (int64_t)acc.size();
	GUARD_END(-1)
}
extern "C" int64_t yakb_ch_dump_mem(const yak_ch_t *h, uint8_t **out) { return dump_mem(h, out, true); }

// a shard's image straight into its place in a file the ranks write side by side (no copy through the caller: images of
// gigabytes).  size = what the image will take; at = write it at `offset` of the existing file `fn`.
extern "C" int64_t yakb_ch_dump_shard_mem(const yak_ch_t *h, int with_header, uint8_t **out) { return dump_mem(h, out, with_header != 0); }
extern "C" int64_t yakb_ch_dump_shard_size(const yak_ch_t *h, int with_header)
{
	GUARD_BEGIN
	ChBox *b = box_single(h, __func__);
	std::lock_guard<std::mutex> lk(b->mu);
	return engine_image_bytes(b->eng, with_header != 0);
	GUARD_END(-1)
}
extern "C" int64_t yakb_ch_dump_shard_at(const yak_ch_t *h, int with_header, const char *fn, uint64_t offset)
{
	GUARD_BEGIN
	ChBox *b = box_single(h, __func__);
	std::lock_guard<std::mutex> lk(b->mu);
	const int fd = open(fn, O_WRONLY);
	if (fd < 0) return -1;
	const int64_t n = dump_engine_at(b, b->eng, with_header != 0, fd, offset);
	close(fd);
	return n;
	GUARD_END(-1)
}

// test hook (no GPU): the arrays yak_ch_restore_core hands to the device, malloc'd for the caller
extern "C" int yakb_yak_file_read(const char *fn, int mode, int min_cnt, int mid_cnt, int threads, uint32_t *k, uint32_t *pre,
                                  uint32_t **caps, uint64_t **off, uint64_t **keys)
{
	YakFile yf;
	const int rc = read_yak_file(fn, mode, min_cnt, mid_cnt, yf, threads);
	if (rc != 0) return rc;
	const size_t P = yf.caps.size();
	*k = yf.k; *pre = yf.pre;
	*caps = (uint32_t*)malloc(P * 4 + 4); *off = (uint64_t*)malloc((P + 1) * 8); *keys = (uint64_t*)malloc(yf.keys.size() * 8 + 8);
	memcpy(*caps, yf.caps.data(), P * 4); memcpy(*off, yf.off.data(), (P + 1) * 8); memcpy(*keys, yf.keys.data(), yf.keys.size() * 8);
	return 0;
}

extern "C" yak_ch_t *yak_ch_restore_core(yak_ch_t *ch0, const char *fn, int mode, ...) // htab.c:396-476
{
	GUARD_BEGIN
	int min_cnt = 0, mid_cnt = 0, mode_err = 0;
	{
		va_list ap;
		va_start(ap, mode);
		if (mode == YAK_LOAD_ALL) {
		} else if (mode == YAK_LOAD_TRIOBIN1 || mode == YAK_LOAD_TRIOBIN2) {
			min_cnt = va_arg(ap, int);
			mid_cnt = va_arg(ap, int);
			if (ch0 == 0 && mode == YAK_LOAD_TRIOBIN2) mode_err = 1;
		} else if (mode == YAK_LOAD_SEXCHR1 || mode == YAK_LOAD_SEXCHR2 || mode == YAK_LOAD_SEXCHR3) {
			if (ch0 == 0 && mode != YAK_LOAD_SEXCHR1) mode_err = 1;
		} else mode_err = 1;
		va_end(ap);
	}
	if (mode_err) return 0;
	YakFile yf;
	const int rc = read_yak_file(fn, mode, min_cnt, mid_cnt, yf, 0, true); // the header first: no table for a file that is not one
	if (rc == -2) fprintf(stderr, "ERROR: wrong file magic.\n");
	if (rc == -3) fprintf(stderr, "ERROR: saved counter bits: %d; compile-time counter bits: %d\n", (int)yf.counter_bits, YAK_COUNTER_BITS);
	if (rc != 0) return 0;
	yak_ch_t *ch = ch0 ? ch0 : yak_ch_init((int)yf.k, (int)yf.pre, 0, 0);
	if (!ch) return 0;
	assert((int)yf.k == ch->k && (int)yf.pre == ch->pre); // htab.c:437
	if (yf.caps.empty() && read_yak_file(fn, mode, min_cnt, mid_cnt, yf, 0) != 0) { if (!ch0) yak_ch_destroy(ch); return 0; }
	ChBox *b = box_single(ch, __func__);
	std::lock_guard<std::mutex> lk(b->mu);
	const int P = 1 << ch->pre;
	std::vector<uint32_t> &caps = yf.caps;
	std::vector<uint64_t> &off = yf.off;
	KeyBuf &keys = yf.keys;
	uint64_t n_new = keys.size();
	const bool engine_sharded = b->eng->P != P; // shards hold part of the sub-tables: not a restore target
	if (engine_sharded) throw CudaError("yak_ch_restore_core on a shard of a multi-GPU table");
	if (ch0 == 0 && mode == YAK_LOAD_ALL) b->eng->load_subtables(caps, off, keys.data());
	else n_new = b->eng->upsert(caps, off, keys.data(), mode != YAK_LOAD_ALL);
	if (ch0 == 0) ch->tot = 0; // the reference leaves tot untouched on restore (htab.c:441)
	fprintf(stderr, "[M::%s] inserted %ld k-mers, of which %ld are new\n", __func__, (long)keys.size(), (long)n_new);
	return ch;
	GUARD_END(0)
}

extern "C" yak_ch_t *yak_ch_restore(const char *fn) { return yak_ch_restore_core(0, fn, YAK_LOAD_ALL); }

// ------------------------------------------------------------------ chunk-level entry points

static int run_ascii_dev(ChBox *b, const uint8_t *d_asc, uint64_t n, int create_new, uint64_t stats[4])
{
	ChunkStats st = b->eng->count_ascii(d_asc, n, create_new);
	b->pub.tot = b->eng->tot;
	if (stats) { stats[0] = st.n_events; stats[1] = st.n_pending; stats[2] = st.n_put; stats[3] = st.n_new; }
	return 0;
}

extern "C" int yakb_count_ascii_dev(yak_ch_t *h, const void *d_asc, uint64_t n, int create_new, uint64_t stats[4])
{
	GUARD_BEGIN
	ChBox *b = box_single(h, __func__);
	std::lock_guard<std::mutex> lk(b->mu);
	return run_ascii_dev(b, (const uint8_t*)d_asc, n, create_new, stats);
	GUARD_END(-1)
}

extern "C" int yakb_count_ascii_host(yak_ch_t *h, const char *asc, uint64_t n, int create_new, uint64_t stats[4])
{
	GUARD_BEGIN
	ChBox *b = box_single(h, __func__);
	std::lock_guard<std::mutex> lk(b->mu);
	uint8_t *d = b->d_in.as<uint8_t>(n + 64);
	YAKB_CUDA(cudaMemcpyAsync(d, asc, n, cudaMemcpyHostToDevice, b->eng->stream));
	return run_ascii_dev(b, d, n, create_new, stats);
	GUARD_END(-1)
}

extern "C" int yakb_count_events_dev(yak_ch_t *h, const uint64_t *d_ev, uint64_t n, int create_new, uint64_t stats[4])
{
	GUARD_BEGIN
	ChBox *b = box_single(h, __func__);
	std::lock_guard<std::mutex> lk(b->mu);
	ChunkStats st = b->eng->count_events(d_ev, n, create_new, -1);
	b->pub.tot = b->eng->tot;
	if (stats) { stats[0] = st.n_events; stats[1] = st.n_pending; stats[2] = st.n_put; stats[3] = st.n_new; }
	return 0;
	GUARD_END(-1)
}

static RouteScratch g_route_scratch;
static std::mutex g_route_mu;
extern "C" int yakb_extract_route_dev(const void *d_asc, uint64_t n, int k, int pre, int world, uint64_t *d_out, uint64_t *counts, void *cuda_stream)
{
	GUARD_BEGIN
	std::lock_guard<std::mutex> lk(g_route_mu);
	return extract_events((const uint8_t*)d_asc, n, k, pre, world, d_out, counts, (cudaStream_t)cuda_stream, g_route_scratch);
	GUARD_END(-1)
}

extern "C" int yakb_extract_route_async(const void *d_asc, uint64_t n, int k, int pre, int world, uint64_t *d_out, uint64_t *d_counts, void *cuda_stream)
{
	GUARD_BEGIN
	std::lock_guard<std::mutex> lk(g_route_mu);
	return extract_events_async((const uint8_t*)d_asc, n, k, pre, world, d_out, d_counts, (cudaStream_t)cuda_stream, g_route_scratch);
	GUARD_END(-1)
}

extern "C" int yakb_ch_reserve(yak_ch_t *h, uint64_t keys_per_subtable)
{
	GUARD_BEGIN
	ChBox *b = box_single(h, __func__);
	std::lock_guard<std::mutex> lk(b->mu);
	b->eng->reserve(keys_per_subtable);
	return 0;
	GUARD_END(-1)
}

extern "C" void *yakb_ch_stream(const yak_ch_t *h) { return (void*)box_of(h)->eng->stream; }
extern "C" uint64_t yakb_ch_device_bytes(const yak_ch_t *h)
{
	uint64_t tot = 0;
	each_shard(box_of(h), false, [&](Engine *e, int) { tot += e->device_bytes(); });
	return tot;
}
extern "C" int yakb_ch_gpus(const yak_ch_t *h) { ChBox *b = box_of(h); return b->shards.empty() ? 1 : (int)b->shards.size(); }
extern "C" const char *yakb_version(void) { return YAKS_VERSION; }
extern "C" int yakb_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }
extern "C" uint64_t yakb_kernel_launches(void) { return Engine::launches(); }
extern "C" uint64_t yakb_device_cache_bytes(void) { return (uint64_t)dev_pool_idle(); }
extern "C" void yakb_device_cache_trim(void) { dev_trim(); }
extern "C" void yakb_prof_enable(int on) { Prof::enable(on != 0); if (on) Prof::reset(); }
extern "C" int yakb_prof_json(char *buf, uint64_t cap)
{
	Prof::resolve(); // event pairs recorded outside the per-chunk path (layout, shrink); waits for their completion
	std::string j = Prof::json();
	if (j.size() + 1 > cap) return -1;
	memcpy(buf, j.c_str(), j.size() + 1);
	return (int)j.size();
}

extern "C" int yakb_synth_genome_dev(uint64_t seed_g, uint64_t G, uint64_t *d_genome2, void *cuda_stream)
{
	GUARD_BEGIN
	synth_genome(seed_g, G, d_genome2, (cudaStream_t)cuda_stream);
	return 0;
	GUARD_END(-1)
}
extern "C" int yakb_synth_reads_dev(const uint64_t *d_genome2, uint64_t G, uint64_t seed_r, uint64_t first, uint64_t n_reads,
                                    int L, double err, int n_pct, int fmt, uint8_t *d_asc, void *cuda_stream)
{
	GUARD_BEGIN
	synth_reads(d_genome2, G, seed_r, first, n_reads, L, err, n_pct, fmt, d_asc, (cudaStream_t)cuda_stream);
	return 0;
	GUARD_END(-1)
}

// ------------------------------------------------------------------ host-only helpers (no GPU needed)

extern "C" void *yakb_fastx_open(const char *fn)
{
	FastxReader *r = new FastxReader;
	if (!r->open(fn)) { delete r; return 0; }
	return r;
}
// bgzf_threads < 0: zlib's sequential reader even for a BGZF file; 0: one inflating thread per core (csrc/bgzf.h)
extern "C" void *yakb_fastx_open_bgzf(const char *fn, int bgzf_threads, uint64_t job_bytes)
{
	FastxReader *r = new FastxReader;
	if (!r->open(fn, bgzf_threads, (size_t)job_bytes)) { delete r; return 0; }
	return r;
}
extern "C" int yakb_fastx_bgzf_threads(void *reader) { return ((FastxReader*)reader)->bgzf_threads(); }
// test hooks (no GPU) for csrc/textcache.h: what yak_count does around its first pass over a compressed file ...
namespace { std::mutex g_tee_mu; std::map<void*, std::pair<std::string, int>> g_tee; }
extern "C" void *yakb_fastx_open_tee(const char *fn)
{
	FastxReader *r = new FastxReader;
	if (!r->open(fn)) { delete r; return 0; }
	std::string key;
	const int fd = text_cache_begin(fn, &key);
	if (fd >= 0) { r->tee_to(fd, text_cache_budget()); std::lock_guard<std::mutex> lk(g_tee_mu); g_tee[r] = {key, fd}; }
	return r;
}
extern "C" int yakb_fastx_tee_commit(void *reader) // after the last fill(): 1 if the text was kept
{
	std::pair<std::string, int> t;
	{ std::lock_guard<std::mutex> lk(g_tee_mu); auto it = g_tee.find(reader); if (it == g_tee.end()) return 0; t = it->second; g_tee.erase(it); }
	const uint64_t n = ((FastxReader*)reader)->tee_bytes();
	text_cache_end(t.first, t.second, n);
	return n != UINT64_MAX;
}
// ... and before its second: the path that maps the kept text ("" if none), and the release that follows the open
extern "C" int yakb_text_cache_path(const char *fn, char *buf, int len)
{
	const std::string p = text_cache_lookup(fn);
	if ((int)p.size() + 1 > len) return -1;
	memcpy(buf, p.c_str(), p.size() + 1);
	return (int)p.size();
}
extern "C" void yakb_text_cache_release(void) { text_cache_release(); }
extern "C" int64_t yakb_fastx_next(void *reader, const char **seq, const char **name)
{
	FastxReader *r = (FastxReader*)reader;
	int64_t len = r->next();
	if (seq) *seq = r->seq().c_str();
	if (name) *name = r->name().c_str();
	return len;
}
extern "C" void yakb_fastx_close(void *reader) { delete (FastxReader*)reader; }
extern "C" int64_t yakb_fastx_fill(void *reader, char *buf, uint64_t cap, uint64_t target, int min_len, int64_t *n_seq, int *done, uint64_t *need)
{
	FastxReader *r = (FastxReader*)reader;
	bool d = false;
	size_t nd = 0;
	int64_t ns = 0;
	size_t n = r->fill((uint8_t*)buf, cap, target, min_len, &ns, &d, &nd);
	if (n_seq) *n_seq = ns;
	if (done) *done = d;
	if (need) *need = nd;
	return (int64_t)n;
}

// which records of a stream the reference reads: lens[i] = record length, or -2 for a truncated FASTQ record (csrc/ref_flow.h)
extern "C" void yakb_ref_flow_sim(const int64_t *lens, int64_t n, int workers, int64_t chunk_size, int min_len, uint8_t *read_out)
{
	yakb_ref_flow_t f;
	yakb_ref_flow_init(&f, workers, chunk_size, min_len);
	int64_t i = 0;
	for (; i < n; ++i) {
		if (lens[i] == -2) { read_out[i] = 0; if (!yakb_ref_flow_bad(&f)) { ++i; break; } }
		else { read_out[i] = 1; yakb_ref_flow_record(&f, lens[i]); }
	}
	for (; i < n; ++i) read_out[i] = 0;
}

extern "C" void yakb_fastx_set_chunk(void *reader, int64_t chunk_size) { ((FastxReader*)reader)->set_ref_chunk(chunk_size); }
extern "C" void yakb_pfastx_set_chunk(void *reader, int64_t chunk_size) { ((ParallelFastx*)reader)->set_ref_chunk(chunk_size); }
extern "C" void yakb_fastx_set_workers(void *reader, int workers) { ((FastxReader*)reader)->set_ref_workers(workers); }
extern "C" void yakb_pfastx_set_flow(void *reader, int64_t chunk_size, int workers, int flow_min_len) { ((ParallelFastx*)reader)->set_ref_flow(chunk_size, workers, flow_min_len); }

extern "C" void *yakb_pfastx_open(const char *fn, uint64_t block_bytes, int threads)
{
	ParallelFastx *r = new ParallelFastx;
	if (!r->open(fn, (size_t)block_bytes, threads)) { delete r; return 0; }
	return r;
}
extern "C" int64_t yakb_pfastx_fill(void *reader, char *buf, uint64_t cap, uint64_t target, int min_len, int64_t *n_seq, int *done, uint64_t *need)
{
	ParallelFastx *r = (ParallelFastx*)reader;
	bool d = false;
	size_t nd = 0;
	int64_t ns = 0;
	size_t n = r->fill((uint8_t*)buf, cap, target, min_len, &ns, &d, &nd);
	if (n_seq) *n_seq = ns;
	if (done) *done = d;
	if (need) *need = nd;
	return (int64_t)n;
}
extern "C" uint64_t yakb_pfastx_redo(void *reader) { return ((ParallelFastx*)reader)->mis_speculations(); }
extern "C" void yakb_pfastx_close(void *reader) { delete (ParallelFastx*)reader; }

// Skip n_skip records, then append up to n_take records of length >= min_len to buf as "SEQ\n".
// Returns the number of records CONSUMED (skipped + taken + dropped short ones counted among the taken);
// fewer than n_skip + n_take means end of input.  *n_bytes / *n_seq describe what was appended.
extern "C" int64_t yakb_fastx_read_slice(void *reader, int64_t n_skip, int64_t n_take, int min_len,
                                         char *buf, uint64_t cap, uint64_t *n_bytes, int64_t *n_seq)
{
	FastxReader *r = (FastxReader*)reader;
	int64_t used = 0;
	uint64_t n = 0;
	*n_bytes = 0; *n_seq = 0;
	for (int64_t i = 0; i < n_skip; ++i) { if (r->next() < 0) return used; ++used; }
	for (int64_t i = 0; i < n_take; ++i) {
		int64_t len = r->next();
		if (len < 0) break;
		++used;
		if (len < min_len) continue;
		if (n + (uint64_t)len + 1 > cap) return -1; // caller's buffer too small
		memcpy(buf + n, r->seq().data(), len);
		n += len; buf[n++] = '\n';
		++*n_seq;
	}
	*n_bytes = n;
	return used;
}

// ------------------------------------------------------------------ yak_count / yak_recount

// bases per device batch on the host-fed paths.  Results do not depend on it; 64 M
// keeps the two pinned staging buffers cheap to allocate while each batch still fills the GPU.
static uint64_t batch_bases(int64_t chunk_size)
{
	uint64_t b = 64ull << 20;
	const char *e = getenv("YAKB_BATCH");
	if (e && atoll(e) > 0) b = (uint64_t)atoll(e);
	if ((uint64_t)chunk_size > b) b = (uint64_t)chunk_size;
	if (b > 0x70000000ull) b = 0x70000000ull;
	return b;
}


// count.c:147-166 with 85-145 folded in: read records, drop those shorter than k (count.c:95),
// concatenate with '\n' separators into pinned memory, ship the batch, run the device path.
// Three things overlap: the reader's pool parses ahead of the consumer, the producer thread stitches
// batch i+1 into pinned memory and copies it to the device on its own stream, the main thread runs
// the kernels of batch i.
// ---- one batch on a multi-GPU table (inside one process: one host thread per GPU).
//      The batch ("SEQ\nSEQ\n..." in pinned host memory) is cut into G contiguous parts at record boundaries.  Thread r
//      copies part r to GPU r and extracts its k-mers grouped by owner (extract_events: count.c:28-60 + the stable
//      partition); after a barrier every GPU pulls the runs it owns from its peers' buffers - source ranks in order, i.e.
//      file order per sub-table (count.c:120,133) - with peer copies over NVLink, and counts them on its shard.
namespace {
struct Barrier {
	std::mutex mu; std::condition_variable cv; int n, waiting = 0; uint64_t gen = 0;
	explicit Barrier(int n_) : n(n_) {}
	void wait()
	{
		std::unique_lock<std::mutex> lk(mu);
		const uint64_t g = gen;
		if (++waiting == n) { waiting = 0; ++gen; cv.notify_all(); }
		else cv.wait(lk, [&] { return gen != g; });
	}
};
}

static void multi_batch(ChBox *b, const uint8_t *host, size_t n, int create_new)
{
	const int G = (int)b->shards.size(), k = b->pub.k, pre = b->pub.pre;
	std::vector<size_t> cut(G + 1, 0);
	cut[G] = n;
	for (int r = 1; r < G; ++r) { // the last record boundary at or before r*n/G
		size_t p = std::max(cut[r - 1], n * (size_t)r / G);
		while (p > cut[r - 1] && host[p - 1] != '\n') --p;
		cut[r] = p;
	}
	std::vector<std::vector<uint64_t>> counts(G, std::vector<uint64_t>(G, 0)); // counts[s][d]: events of part s owned by GPU d
	std::vector<const uint64_t*> d_ev(G, nullptr);
	std::vector<std::string> err(G);
	std::atomic<int> failed{0};
	Barrier mid(G);
	std::vector<uint64_t> n_new(G, 0);
	auto work = [&](int r) {
		Shard *sh = b->shards[r];
		cudaSetDevice(sh->dev);
		try {
			const size_t nr = cut[r + 1] - cut[r];
			uint8_t *d_in = sh->d_in.as<uint8_t>(nr + 64);
			uint64_t *ev = sh->d_ev.as<uint64_t>(std::max<size_t>(nr, 1));
			if (nr) YAKB_CUDA(cudaMemcpyAsync(d_in, host + cut[r], nr, cudaMemcpyHostToDevice, sh->stream));
			if (extract_events(d_in, nr, k, pre, G, ev, counts[r].data(), sh->stream, sh->route) != 0) throw CudaError("extract_events failed");
			YAKB_CUDA(cudaStreamSynchronize(sh->stream));
			d_ev[r] = ev;
		} catch (const std::exception &e) { err[r] = e.what(); failed = 1; }
		mid.wait(); // every part is extracted, every count is known
		if (failed) { mid.wait(); return; }
		try {
			uint64_t total = 0;
			for (int s = 0; s < G; ++s) total += counts[s][r];
			uint64_t *recv = sh->d_recv.as<uint64_t>(std::max<uint64_t>(total, 1));
			uint64_t off = 0;
			for (int s = 0; s < G; ++s) {
				uint64_t src_off = 0;
				for (int d = 0; d < r; ++d) src_off += counts[s][d];
				if (counts[s][r]) YAKB_CUDA(cudaMemcpyPeerAsync(recv + off, sh->dev, d_ev[s] + src_off, b->shards[s]->dev, counts[s][r] * 8, sh->eng->stream));
				off += counts[s][r];
			}
			n_new[r] = sh->eng->count_events(recv, total, create_new, -1).n_new; // on the same stream, behind the copies
		} catch (const std::exception &e) { err[r] = e.what(); failed = 1; }
		mid.wait(); // nobody reuses its event buffer before every peer has pulled from it
	};
	std::vector<std::thread> th;
	for (int r = 0; r < G; ++r) th.emplace_back(work, r);
	for (auto &t : th) t.join();
	for (auto &e : err) if (!e.empty()) throw CudaError(e);
	for (int r = 0; r < G; ++r) b->pub.tot += n_new[r];
}

// ---- yak_count with the text parsed on the device (csrc/ingest.cu).  For plain files in the strict 4-line FASTQ / 2-line
//      FASTA layout the host does nothing but move bytes: a few threads copy the next batch of the memory-mapped file into
//      page-locked memory and on to the GPU while the GPU turns the previous batch into the dense base stream (validating
//      the layout of every record) and counts it.  Batches are cut where a record starts; that cut is only a guess (a line
//      that starts with the marker and, for FASTQ, is followed two lines later by a '+' line) - the device check decides.
//      Returns 0 = the whole file was counted; 1 = the layout does not hold and NOTHING was counted (the caller takes the host
//      parser); 2 = the layout broke after batches had been counted (the caller must start the pass over).
namespace {
struct FileId {
	dev_t dev; ino_t ino; off_t size; int64_t mtime_ns;
	bool operator<(const FileId &o) const { return std::tie(dev, ino, size, mtime_ns) < std::tie(o.dev, o.ino, o.size, o.mtime_ns); }
};
std::mutex g_strict_mu;
std::map<FileId, bool> g_strict_files; // files a whole pass went through the device check (main.c:57 reads the same file again)
FileId file_id(const struct stat &st) { return FileId{st.st_dev, st.st_ino, st.st_size, (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec}; }
}

// YAKB_GPU_INGEST: 0 = never, 1 = whenever the file qualifies (tests), unset = files of at least 256 MB
static int gpu_ingest_mode() { static int v = -2; if (v == -2) { const char *e = getenv("YAKB_GPU_INGEST"); v = e ? atoi(e) : -1; } return v; }

// the last record start at or before `pos` (and behind `lo`), or `lo` when the window holds none
static uint64_t record_start_before(const uint8_t *m, uint64_t lo, uint64_t pos, uint64_t size, int lpr, uint8_t marker)
{
	if (pos >= size) return size;
	for (uint64_t p = pos; p > lo; --p) {
		if (m[p - 1] != '\n' || m[p] != marker) continue;
		if (lpr == 2) return p;
		const uint8_t *e1 = (const uint8_t*)memchr(m + p, '\n', size - p);
		if (!e1) continue;
		const uint8_t *e2 = (const uint8_t*)memchr(e1 + 1, '\n', size - (uint64_t)(e1 + 1 - m));
		if (!e2 || (uint64_t)(e2 + 1 - m) >= size) continue;
		if (e2[1] == '+') return p; // a quality line that starts with '@' is followed two lines later by bases, not by '+'
	}
	return lo;
}

static void parallel_copy(uint8_t *dst, const uint8_t *src, size_t n, int threads)
{
	if (n < (8u << 20) || threads <= 1) { memcpy(dst, src, n); return; }
	std::vector<std::thread> th;
	for (int t = 0; t < threads; ++t) {
		const size_t a = n * (size_t)t / threads, e = n * (size_t)(t + 1) / threads;
		th.emplace_back([=] { memcpy(dst + a, src + a, e - a); });
	}
	for (auto &t : th) t.join();
}

static int count_gpu_ingest(yak_ch_t *h, const char *src, const struct stat &st, int create_new)
{
	StageTimer tm("yak_count(device ingest)");
	ChBox *b = box_of(h);
	std::lock_guard<std::mutex> lk(b->mu);
	const uint64_t size = (uint64_t)st.st_size;
	const int fd = open(src, O_RDONLY);
	if (fd < 0) return 1;
	const uint8_t *m = (const uint8_t*)mmap(nullptr, size, PROT_READ, MAP_SHARED, fd, 0);
	close(fd);
	if (m == MAP_FAILED) return 1;
	struct Unmap { const uint8_t *p; uint64_t n; ~Unmap() { munmap((void*)p, n); } } unmap{m, size};
	madvise((void*)m, size, MADV_SEQUENTIAL);
	const uint8_t marker = m[0];
	const int lpr = marker == '@' ? 4 : marker == '>' ? 2 : 0;
	if (lpr == 0 || m[size - 1] != '\n') return 1;
	// raw bytes per batch: ~1/32 of the file, between 128 MB and 1.5 GB (the dense stream of a batch must stay below 2^31 bytes)
	uint64_t B = std::min<uint64_t>(std::max<uint64_t>(size / 32, 128ull << 20), 1536ull << 20);
	if (const char *e = getenv("YAKB_INGEST_BATCH")) if (atoll(e) > 0) B = (uint64_t)atoll(e);
	B = std::min<uint64_t>(B, size);
	int dev = 0;
	cudaGetDevice(&dev);
	if (!b->copy_stream) YAKB_CUDA(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
	uint8_t *pinned[2] = {nullptr, nullptr};
	size_t pinned_cap[2] = {0, 0};
	struct Release { uint8_t **p; size_t *c; ~Release() { for (int i = 0; i < 2; ++i) g_pinned.put(p[i], c[i]); } } release{pinned, pinned_cap};
	auto ensure_pinned = [&](int i, size_t bytes) {
		if (pinned_cap[i] < bytes) { g_pinned.put(pinned[i], pinned_cap[i]); pinned[i] = nullptr; pinned_cap[i] = 0; pinned[i] = g_pinned.get(bytes, &pinned_cap[i]); }
	};
	const int cthreads = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 8u);
	DBuf *d_raw[2] = {&b->d_in, &b->d_in2};
	struct Batch { uint64_t a = 0, e = 0; std::string err; };
	Batch batch[2];
	auto produce = [&](int slot, uint64_t a) { // [a, e): whole records, about B bytes (one record at least)
		Batch &t = batch[slot];
		t = Batch();
		t.a = a;
		try {
			cudaSetDevice(dev);
			const uint64_t e = a + B >= size ? size : record_start_before(m, a, a + B, size, lpr, marker);
			if (e <= a) { t.err = "a record longer than a batch"; return; } // a chromosome on one line: the host parser's business
			t.e = e;
			const uint64_t n = e - a;
			ensure_pinned(slot, n + 4096);
			parallel_copy(pinned[slot], m + a, n, cthreads);
			uint8_t *d = d_raw[slot]->as<uint8_t>(n + 64);
			YAKB_CUDA(cudaMemcpyAsync(d, pinned[slot], n, cudaMemcpyHostToDevice, b->copy_stream));
			YAKB_CUDA(cudaStreamSynchronize(b->copy_stream));
		} catch (const std::exception &ex) { t.err = ex.what(); }
	};
	int slot = 0, n_batches = 0;
	double t_dev = 0, t_wait = 0;
	produce(0, 0);
	for (;;) {
		Batch cur = batch[slot];
		if (!cur.err.empty()) { if (cur.err[0] == 'a') return n_batches ? 2 : 1; throw CudaError(cur.err); }
		std::thread producer;
		if (cur.e < size) producer = std::thread(produce, slot ^ 1, cur.e);
		struct Join { std::thread &t; ~Join() { if (t.joinable()) t.join(); } } join_on_unwind{producer};
		const uint64_t n = cur.e - cur.a;
		const double td = wall_now();
		uint8_t *dense = b->d_aux.as<uint8_t>(n + 64);
		unsigned long long *d_res = b->d_aux2.as<unsigned long long>(4), res[3] = {0, 0, 0};
		if (ingest_strict((const uint8_t*)d_raw[slot]->p, n, lpr, dense, d_res, b->eng->stream, b->ing) != 0) return n_batches ? 2 : 1;
		YAKB_CUDA(cudaMemcpyAsync(res, d_res, sizeof(res), cudaMemcpyDeviceToHost, b->eng->stream));
		YAKB_CUDA(cudaStreamSynchronize(b->eng->stream));
		if (res[0] != 0 || res[1] >= 0x7FFFFF00ull) {
			if (yak_verbose >= 3) fprintf(stderr, "[M::%s] '%s' is not in the strict %d-line layout (check %llx at batch %d): using the host parser\n", __func__, src, lpr, res[0], n_batches + 1);
			return n_batches ? 2 : 1;
		}
		run_ascii_dev(b, dense, res[1], create_new, nullptr);
		t_dev += wall_now() - td; ++n_batches;
		if (timing_on()) fprintf(stderr, "[T::yak_count] batch %d: %llu bytes of text, %llu of bases, device %.4f s\n", n_batches, (unsigned long long)n, res[1], wall_now() - td);
		fprintf(stderr, "[M::%s::%.3f*%.2f] processed %d sequences; %ld distinct k-mers in the hash table\n", "count_impl",
		        wall_now() - g_t0, cpu_now() / (wall_now() - g_t0 + 1e-9), (int)(res[2] / lpr), (long)h->tot);
		{ const double tw = wall_now(); if (producer.joinable()) producer.join(); t_wait += wall_now() - tw; }
		if (cur.e >= size) break;
		slot ^= 1;
	}
	{ std::lock_guard<std::mutex> g(g_strict_mu); g_strict_files[file_id(st)] = true; }
	if (timing_on()) fprintf(stderr, "[T::yak_count] device ingest: %d batches, device %.3f s, waiting for the copies %.3f s\n", n_batches, t_dev, t_wait);
	return 0;
}

// the two halves of the device ingest as calls of their own (yak_b200/dist.py: every rank of a multi-GPU job takes its own
// byte range of the file): where to cut, and text -> dense base stream + layout check on the current device
extern "C" uint64_t yakb_record_start_before(const void *text, uint64_t lo, uint64_t pos, uint64_t size, int lines_per_record)
{
	return record_start_before((const uint8_t*)text, lo, pos, size, lines_per_record, lines_per_record == 4 ? '@' : '>');
}
static IngestScratch g_ingest_scratch;
static std::mutex g_ingest_mu;
extern "C" int yakb_ingest_dev(const void *d_raw, uint64_t n, int lines_per_record, void *d_out, uint64_t *d_res, void *cuda_stream)
{
	GUARD_BEGIN
	std::lock_guard<std::mutex> lk(g_ingest_mu);
	return ingest_strict((const uint8_t*)d_raw, n, lines_per_record, (uint8_t*)d_out, (unsigned long long*)d_res, (cudaStream_t)cuda_stream, g_ingest_scratch);
	GUARD_END(-1)
}

static yak_ch_t *count_impl(const char *fn, const yak_copt_t *opt, yak_ch_t *h0, int ref_workers);
extern "C" yak_ch_t *yak_count(const char *fn, const yak_copt_t *opt, yak_ch_t *h0) { return count_impl(fn, opt, h0, 3); } // count.c:162

// ref_workers: the reference's pipeline threads for this entry point (what follows a truncated FASTQ record depends on them,
// csrc/ref_flow.h); 0 = yak_recount's plain read loop
static yak_ch_t *count_impl(const char *fn, const yak_copt_t *opt, yak_ch_t *h0, int ref_workers)
{
	GUARD_BEGIN
	StageTimer tm("yak_count");
	FastxReader rd;
	ParallelFastx prd; // plain regular files are parsed by several threads; gzip / stdin by the sequential reader
	// the second pass over a compressed file reads the text the first pass kept (csrc/textcache.h; YAKB_TEXT_CACHE_GB, off by default)
	const std::string cached = h0 ? text_cache_lookup(fn) : std::string();
	const char *src = cached.empty() ? fn : cached.c_str();
	const bool par = !getenv("YAKB_SERIAL_PARSE") && prd.open(src);
	if (!cached.empty()) text_cache_release(); // one use: the mapping keeps the text alive until prd closes
	if (!par && !rd.open(fn)) return 0;
	struct Tee { std::string key; int fd = -1; ~Tee() { if (fd >= 0) close(fd); } } tee;
	if (!par && !h0 && (tee.fd = text_cache_begin(fn, &tee.key)) >= 0) rd.tee_to(tee.fd, text_cache_budget());
	prd.set_ref_flow(opt->chunk_size, ref_workers, -1); // -K and the pipeline decide what follows a truncated FASTQ record
	rd.set_ref_chunk(opt->chunk_size); rd.set_ref_workers(ref_workers);
	yak_ch_t *h = h0;
	if (h0) assert(h0->k == opt->k && h0->pre == opt->pre);
	else {
		const int gpus = multi_gpus_wanted(); // YAKB_GPUS > 1: the new table is spread over that many GPUs of this process
		h = gpus > 1 && (1 << opt->pre) >= gpus ? ch_init_multi(opt->k, opt->pre, opt->bf_n_hash, opt->bf_shift, gpus)
		                                          : yak_ch_init(opt->k, opt->pre, opt->bf_n_hash, opt->bf_shift);
	}
	if (!h) return 0;
	// Large plain files in the strict FASTQ / FASTA layout are parsed on the device (count_gpu_ingest above).  A counting-only
	// pass (h0 given: counts cannot be taken back) goes that way only for a file an earlier pass of this process has checked.
	{
		struct stat sti;
		const int gi = gpu_ingest_mode();
		if (par && gi != 0 && box_of(h)->shards.empty() && stat(src, &sti) == 0 && S_ISREG(sti.st_mode) && sti.st_size > 0 &&
		    (gi == 1 || sti.st_size >= (256ll << 20))) {
			bool known = false;
			{ std::lock_guard<std::mutex> g(g_strict_mu); known = g_strict_files.count(file_id(sti)) != 0; }
			if (!h0 || known) {
				const int rc = count_gpu_ingest(h, src, sti, h0 == 0);
				if (rc == 0) return h;
				if (rc == 2) {
					if (h0) { fprintf(stderr, "[yakb] FATAL: '%s' changed while it was being counted\n", src); abort(); }
					yak_ch_destroy(h); // our own table: start the pass over with the host parser
					h = yak_ch_init(opt->k, opt->pre, opt->bf_n_hash, opt->bf_shift);
					if (!h) return 0;
				}
			}
		}
	}
	ChBox *b = box_of(h);
	std::lock_guard<std::mutex> lk(b->mu);
	const bool multi = !b->shards.empty();
	const int create_new = h0 == 0;
	uint64_t cap = batch_bases(opt->chunk_size);
	struct stat st;
	if (par && stat(src, &st) == 0 && S_ISREG(st.st_mode)) {
		// large files get large batches (the partitioned probe gains with the events per chunk; ~32 batches keep parser,
		// copies and kernels overlapped), and a batch never holds more than the file: no 2 x 1.9 GB of pinned memory for
		// `cntasm -K1.9g` on a small assembly
		if (!getenv("YAKB_BATCH")) cap = std::max<uint64_t>(cap, std::min<uint64_t>((uint64_t)st.st_size / 32, 1ull << 30));
		cap = std::min<uint64_t>(cap, std::max<uint64_t>((uint64_t)st.st_size + 4096, 1u << 20));
	}
	int dev = 0;
	cudaGetDevice(&dev);
	if (!b->copy_stream) YAKB_CUDA(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
	uint8_t *pinned[2] = {nullptr, nullptr};
	size_t pinned_cap[2] = {0, 0};
	auto ensure_pinned = [&](size_t bytes) {
		for (int i = 0; i < 2; ++i)
			if (pinned_cap[i] < bytes) { g_pinned.put(pinned[i], pinned_cap[i]); pinned[i] = g_pinned.get(bytes, &pinned_cap[i]); }
	};
	struct Release { uint8_t **p; size_t *c; ~Release() { for (int i = 0; i < 2; ++i) g_pinned.put(p[i], c[i]); } } release{pinned, pinned_cap};
	{ StageTimer tp("yak_count:pinned"); ensure_pinned(cap + (cap >> 4) + 4096); }
	struct Batch { size_t n = 0; int64_t n_seq = 0; bool done = false; size_t need = 0; std::string err; };
	Batch batch[2];
	DBuf *d_in[2] = {&b->d_in, &b->d_in2};
	auto produce = [&](int slot, uint64_t target) {
		Batch &t = batch[slot];
		t = Batch();
		try {
			cudaSetDevice(dev);
			const size_t pc = std::min(pinned_cap[0], pinned_cap[1]);
			t.n = par ? prd.fill(pinned[slot], pc, target, opt->k, &t.n_seq, &t.done, &t.need)
			          : rd.fill(pinned[slot], pc, target, opt->k, &t.n_seq, &t.done, &t.need);
			if (t.n && !t.need && !multi) {
				uint8_t *d = d_in[slot]->as<uint8_t>(t.n + 64);
				YAKB_CUDA(cudaMemcpyAsync(d, pinned[slot], t.n, cudaMemcpyHostToDevice, b->copy_stream));
				YAKB_CUDA(cudaStreamSynchronize(b->copy_stream));
			}
		} catch (const std::exception &e) { t.err = e.what(); t.done = true; }
	};
	int slot = 0;
	// (YAKB_RAMP=<bases> makes the first batch that small and doubles from there; with the pooled parser the
	// first full batch is ready within ~10 ms, so the default is no ramp)
	const char *env_ramp = getenv("YAKB_RAMP");
	uint64_t target = env_ramp && atoll(env_ramp) > 0 ? std::min<uint64_t>(cap, (uint64_t)atoll(env_ramp)) : cap;
	double t_first = wall_now(), t_dev = 0, t_wait = 0;
	int n_batches = 0;
	produce(0, target);
	t_first = wall_now() - t_first;
	for (;;) {
		Batch cur = batch[slot];
		if (!cur.err.empty()) throw CudaError(cur.err);
		if (cur.need) { // one record larger than the staging buffers: grow them and parse again
			ensure_pinned(cur.need + (cur.need >> 3) + 4096);
			produce(slot, target);
			continue;
		}
		const uint64_t next_target = std::min<uint64_t>(cap, target * 2);
		std::thread producer;
		if (!cur.done) producer = std::thread(produce, slot ^ 1, next_target);
		struct Join { std::thread &t; ~Join() { if (t.joinable()) t.join(); } } join_on_unwind{producer};
		if (cur.n) {
			const double td = wall_now();
			if (multi) multi_batch(b, pinned[slot], cur.n, create_new);
			else run_ascii_dev(b, (const uint8_t*)d_in[slot]->p, cur.n, create_new, nullptr);
			t_dev += wall_now() - td; ++n_batches;
			if (timing_on()) fprintf(stderr, "[T::yak_count] batch %d: %zu bytes, kernels %.4f s\n", n_batches, cur.n, wall_now() - td);
			fprintf(stderr, "[M::%s::%.3f*%.2f] processed %d sequences; %ld distinct k-mers in the hash table\n", __func__,
			        wall_now() - g_t0, cpu_now() / (wall_now() - g_t0 + 1e-9), (int)cur.n_seq, (long)h->tot);
		}
		{ const double tw = wall_now(); if (producer.joinable()) producer.join(); t_wait += wall_now() - tw; }
		if (cur.done) break;
		slot ^= 1;
		target = next_target;
	}
	if (tee.fd >= 0) { text_cache_end(tee.key, tee.fd, rd.tee_bytes()); tee.fd = -1; }
	if (timing_on()) fprintf(stderr, "[T::yak_count] %d batches: first batch ready after %.3f s, kernels %.3f s, waiting for the producer %.3f s\n", n_batches, t_first, t_dev, t_wait);
	return h;
	GUARD_END(0)
}

extern "C" void yak_recount(const char *fn, yak_ch_t *h) // count.c:168-193: clear, then count existing k-mers only
{
	yak_copt_t o;
	yak_copt_init(&o);
	o.k = h->k; o.pre = h->pre;
	FILE *probe = (fn && strcmp(fn, "-")) ? fopen(fn, "rb") : stdin;
	if (probe == 0) return; // count.c:172
	if (probe != stdin) fclose(probe);
	yak_ch_clear(h, 1);
	// NB the reference does not drop reads shorter than k here, which changes nothing: they hold no k-mer.  Its loop is a
	// plain `while (kseq_read(ks) >= 0)` (count.c:176): the first truncated FASTQ record ends it.
	count_impl(fn, &o, h, 0);
}

// ------------------------------------------------------------------ qv

// one batch of sequences already concatenated with '\n' separators in pinned/host memory
static void qv_batch(ChBox *b, const uint8_t *cat, uint64_t n, const std::vector<uint64_t> &seq_off, int min_len, double min_frac,
                     unsigned long long *d_hist, int32_t *tot, int32_t *non0, std::vector<int16_t> *cnt_back)
{
	Engine *e = b->eng;
	const uint64_t n_seq = seq_off.size() - 1;
	uint8_t *d = b->d_in.as<uint8_t>(n + 64);
	YAKB_CUDA(cudaMemcpyAsync(d, cat, n, cudaMemcpyHostToDevice, e->stream));
	// layout of d_aux: int16 cnt[n] | u64 seq_off[n_seq+1] | i32 tot[n_seq] | i32 non0[n_seq] | u8 pass[n_seq]
	const size_t o_cnt = 0, o_off = (n * 2 + 15) & ~(size_t)15, o_tot = o_off + (n_seq + 1) * 8, o_non0 = o_tot + n_seq * 4, o_pass = o_non0 + n_seq * 4;
	uint8_t *base = b->d_aux.as<uint8_t>(o_pass + n_seq + 16);
	int16_t *d_cnt = (int16_t*)(base + o_cnt);
	uint64_t *d_off = (uint64_t*)(base + o_off);
	int32_t *d_tot = (int32_t*)(base + o_tot), *d_non0 = (int32_t*)(base + o_non0);
	uint8_t *d_pass = base + o_pass;
	YAKB_CUDA(cudaMemcpyAsync(d_off, seq_off.data(), (n_seq + 1) * 8, cudaMemcpyHostToDevice, e->stream));
	const double tb0 = wall_now();
	if (timing_on()) YAKB_CUDA(cudaStreamSynchronize(e->stream));
	const double tb1 = wall_now();
	qv_scan_ascii(e, d, n, d_cnt);
	if (timing_on()) fprintf(stderr, "[T::qv_batch] H2D %.4f s, pack + scan %.4f s\n", tb1 - tb0, wall_now() - tb1);
	qv_stats(d_cnt, d_off, n_seq, n, min_len, min_frac, d_tot, d_non0, d_pass, d_hist, e->stream);
	if (tot) YAKB_CUDA(cudaMemcpyAsync(tot, d_tot, n_seq * 4, cudaMemcpyDeviceToHost, e->stream));
	if (non0) YAKB_CUDA(cudaMemcpyAsync(non0, d_non0, n_seq * 4, cudaMemcpyDeviceToHost, e->stream));
	if (cnt_back) { cnt_back->resize(n); YAKB_CUDA(cudaMemcpyAsync(cnt_back->data(), d_cnt, n * 2, cudaMemcpyDeviceToHost, e->stream)); }
	YAKB_CUDA(cudaStreamSynchronize(e->stream));
}

extern "C" int yakb_qv_seqs(const yak_ch_t *h, int64_t n_seq, const int64_t *lens, const char *cat,
                            int min_len, double min_frac, int64_t cnt[YAK_N_COUNTS], int32_t *tot, int32_t *non0)
{
	GUARD_BEGIN
	ChBox *b = box_single(h, __func__);
	std::lock_guard<std::mutex> lk(b->mu);
	assert(h->k < 32); // qv.c:43
	unsigned long long *d_hist = (unsigned long long*)b->d_aux2.need(1024 * 8);
	YAKB_CUDA(cudaMemsetAsync(d_hist, 0, 1024 * 8, b->eng->stream));
	const uint64_t cap = batch_bases(0);
	int64_t s = 0, src = 0;
	std::vector<uint8_t> buf;
	while (s < n_seq) {
		std::vector<uint64_t> off(1, 0);
		buf.clear();
		int64_t s_first = s;
		while (s < n_seq && (buf.empty() || buf.size() + lens[s] + 1 <= cap)) {
			buf.insert(buf.end(), cat + src, cat + src + lens[s]);
			buf.push_back('\n');
			off.push_back(buf.size());
			src += lens[s]; ++s;
		}
		qv_batch(b, buf.data(), buf.size(), off, min_len, min_frac, d_hist, tot ? tot + s_first : nullptr, non0 ? non0 + s_first : nullptr, nullptr);
	}
	YAKB_CUDA(cudaMemcpyAsync(cnt, d_hist, 1024 * 8, cudaMemcpyDeviceToHost, b->eng->stream));
	YAKB_CUDA(cudaStreamSynchronize(b->eng->stream));
	return 0;
	GUARD_END(-1)
}

// the lookups of the other scanners (triobin.c:62-86, trioeval.c:61-89, chkerr.c:35-56, sexchr.c:42-66), batched
extern "C" int yakb_scan_seqs(const yak_ch_t *h, int64_t n_seq, const int64_t *lens, const char *cat, int16_t *out)
{
	GUARD_BEGIN
	ChBox *b = box_single(h, __func__);
	std::lock_guard<std::mutex> lk(b->mu);
	Engine *e = b->eng;
	const uint64_t cap = batch_bases(0);
	int64_t s = 0, src = 0;
	std::vector<uint8_t> buf;
	std::vector<int16_t> back;
	while (s < n_seq) {
		buf.clear();
		const int64_t s_first = s, src_first = src;
		while (s < n_seq && (buf.empty() || buf.size() + lens[s] + 1 <= cap)) {
			buf.insert(buf.end(), cat + src, cat + src + lens[s]);
			buf.push_back('\n'); // a separator byte: no k-mer spans two sequences
			src += lens[s]; ++s;
		}
		const uint64_t n = buf.size();
		uint8_t *d = b->d_in.as<uint8_t>(n + 64);
		int16_t *d_cnt = (int16_t*)b->d_aux.need(n * 2 + 16);
		YAKB_CUDA(cudaMemcpyAsync(d, buf.data(), n, cudaMemcpyHostToDevice, e->stream));
		qv_scan_ascii(e, d, n, d_cnt, 1);
		back.resize(n);
		YAKB_CUDA(cudaMemcpyAsync(back.data(), d_cnt, n * 2, cudaMemcpyDeviceToHost, e->stream));
		YAKB_CUDA(cudaStreamSynchronize(e->stream));
		int64_t o = src_first, p = 0;
		for (int64_t i = s_first; i < s; ++i) { // drop the separators again
			memcpy(out + o, back.data() + p, (size_t)lens[i] * 2);
			o += lens[i]; p += lens[i] + 1;
		}
	}
	return 0;
	GUARD_END(-1)
}

// qv.c:116-135 (+ the SQ / EK printing of worker_qv, qv.c:62-81, done in input order)
extern "C" void yak_qv(const yak_qopt_t *opt, const char *fn, const yak_ch_t *ch, int64_t *cnt)
{
	GUARD_BEGIN
	ChBox *b = box_single(ch, __func__);
	std::lock_guard<std::mutex> lk(b->mu);
	assert(ch->k < 32); // qv.c:43
	memset(cnt, 0, YAK_N_COUNTS * sizeof(int64_t));
	unsigned long long *d_hist = (unsigned long long*)b->d_aux2.need(1024 * 8);
	YAKB_CUDA(cudaMemsetAsync(d_hist, 0, 1024 * 8, b->eng->stream));
	const uint64_t cap = std::min<uint64_t>(batch_bases(0), (uint64_t)std::max<int64_t>(opt->chunk_size, 1));
	// Without per-sequence output nothing needs the names: plain files then come through the parser pool as
	// "SEQ\n" batches (sequences shorter than min_len take no part in anything, qv.c:44, so the pool may drop them)
	ParallelFastx prd;
	if (!opt->print_each && !opt->print_err_kmer && !getenv("YAKB_SERIAL_PARSE") && prd.open(fn)) {
		prd.set_ref_flow(opt->chunk_size, 2, 0); // behind a truncated FASTQ record: bseq_read's batches, two pipeline workers (qv.c:94,126)
		size_t pcap = cap + (cap >> 4) + 4096;
		std::unique_ptr<uint8_t[]> pbuf(new uint8_t[pcap]); // not a vector: no need to zero 68 MB first
		bool pdone = false;
		const int min_len = std::max(opt->min_len, 0);
		while (!pdone) {
			int64_t n_seq = 0;
			size_t need = 0;
			const double tq0 = wall_now();
			const size_t n = prd.fill(pbuf.get(), pcap, cap, min_len, &n_seq, &pdone, &need);
			const double tq1 = wall_now();
			if (need) { pcap = need + (need >> 3) + 4096; pbuf.reset(new uint8_t[pcap]); continue; }
			if (n == 0) continue;
			std::vector<uint64_t> off(1, 0);
			off.reserve((size_t)n_seq + 1);
			for (const uint8_t *p = pbuf.get(), *e = p + n; p < e; ) { // one '\n' ends every sequence
				const uint8_t *nl = (const uint8_t*)memchr(p, '\n', e - p);
				p = nl ? nl + 1 : e;
				off.push_back((uint64_t)(p - pbuf.get()));
			}
			fprintf(stderr, "[M::%s] read %d sequences\n", "yak_qv_cb", (int)n_seq);
			const double tq2 = wall_now();
			qv_batch(b, pbuf.get(), n, off, opt->min_len, opt->min_frac, d_hist, nullptr, nullptr, nullptr);
			if (timing_on()) fprintf(stderr, "[T::yak_qv] batch of %zu bytes: parse %.4f s, offsets %.4f s, device %.4f s\n", n, tq1 - tq0, tq2 - tq1, wall_now() - tq2);
			fprintf(stderr, "[M::%s@%.2f*%.2f] processed %d sequences\n", "yak_qv_cb", wall_now() - g_t0, cpu_now() / (wall_now() - g_t0 + 1e-6), (int)n_seq);
		}
		YAKB_CUDA(cudaMemcpyAsync(cnt, d_hist, 1024 * 8, cudaMemcpyDeviceToHost, b->eng->stream));
		YAKB_CUDA(cudaStreamSynchronize(b->eng->stream));
		return;
	}
	FastxReader rd;
	if (!rd.open(fn)) return;
	bool done = false;
	yakb_ref_flow_t flow;
	yakb_ref_flow_init(&flow, 2, opt->chunk_size, 0); // bseq_read under kt_pipeline(2, ...): qv.c:94,126
	std::vector<uint8_t> buf;
	std::vector<std::string> names;
	std::vector<int32_t> tot, non0;
	std::vector<int16_t> back;
	while (!done) {
		std::vector<uint64_t> off(1, 0);
		buf.clear(); names.clear();
		while (buf.size() < cap) { // bseq.c:33-57: a batch closes once it holds >= chunk_size bases
			int64_t len = rd.next();
			if (len == -2) { // a truncated FASTQ record: csrc/ref_flow.h
				if (!yakb_ref_flow_bad(&flow)) { done = true; break; }
				continue;
			}
			if (len < 0) { done = true; break; }
			yakb_ref_flow_record(&flow, len);
			buf.insert(buf.end(), rd.seq().begin(), rd.seq().end());
			buf.push_back('\n');
			off.push_back(buf.size());
			if (opt->print_each || opt->print_err_kmer) names.push_back(rd.name());
		}
		const size_t n_seq = off.size() - 1;
		fprintf(stderr, "[M::%s] read %d sequences\n", "yak_qv_cb", (int)n_seq);
		if (n_seq == 0) break;
		tot.resize(n_seq); non0.resize(n_seq);
		qv_batch(b, buf.data(), buf.size(), off, opt->min_len, opt->min_frac, d_hist, tot.data(), non0.data(),
		         opt->print_err_kmer ? &back : nullptr);
		for (size_t s = 0; s < n_seq; ++s) {
			const int l_seq = (int)(off[s + 1] - off[s] - 1);
			if (l_seq < opt->min_len) continue;
			if (opt->print_err_kmer)
				for (uint64_t p = off[s]; p + 1 < off[s + 1]; ++p)
					if (back[p] == 0) printf("EK\t%s\t%d\n", names[s].c_str(), (int)(p - off[s]) + 1 - ch->k);
			if (opt->print_each) { // qv.c:70-81
				double qv = -1.0;
				if (tot[s] > 0) {
					if (non0[s] > 0) {
						if (tot[s] > non0[s]) { qv = log((double)tot[s] / non0[s]) / ch->k; qv = -4.3429448190325175 * log(qv); }
						else qv = 99.0;
					} else qv = 0.0;
				}
				printf("SQ\t%s\t%d\t%d\t%d\t%.2f\n", names[s].c_str(), l_seq, tot[s], non0[s], qv);
			}
		}
		fprintf(stderr, "[M::%s@%.2f*%.2f] processed %d sequences\n", "yak_qv_cb", wall_now() - g_t0, cpu_now() / (wall_now() - g_t0 + 1e-6), (int)n_seq);
	}
	YAKB_CUDA(cudaMemcpyAsync(cnt, d_hist, 1024 * 8, cudaMemcpyDeviceToHost, b->eng->stream));
	YAKB_CUDA(cudaStreamSynchronize(b->eng->stream));
	GUARD_END_VOID
}

// ------------------------------------------------------------------ bbf.c as a stand-alone device filter

extern "C" yak_bf_t *yak_bf_init(int n_shift, int n_hashes) // bbf.c:5-18
{
	GUARD_BEGIN
	if (n_shift + YAK_BLK_SHIFT > 64 || n_shift < YAK_BLK_SHIFT) return 0;
	if (yakb_device_count() == 0) { fprintf(stderr, "[yakb] ERROR: no CUDA device; this library has no CPU path\n"); return 0; }
	yak_bf_t *b = (yak_bf_t*)calloc(1, sizeof(yak_bf_t));
	b->n_shift = n_shift; b->n_hashes = n_hashes;
	b->b = (uint8_t*)dev_alloc((size_t)1 << (n_shift - 3));
	YAKB_CUDA(cudaMemset(b->b, 0, (size_t)1 << (n_shift - 3)));
	return b;
	GUARD_END(0)
}
extern "C" void yak_bf_destroy(yak_bf_t *b) { if (b == 0) return; dev_free(b->b); free(b); } // bbf.c:20-24
extern "C" int yak_bf_insert(yak_bf_t *b, uint64_t hash) // bbf.c:25-42
{
	GUARD_BEGIN
	return bf_insert_one(b->b, b->n_shift, b->n_hashes, hash);
	GUARD_END(-1)
}
