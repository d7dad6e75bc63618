// engine.cuh - host-side view of the device-resident count table and its per-chunk pipeline.
//
// One Engine backs one yak_ch_t.  HBM layout (all device memory unless noted):
//   slots[P*cap]      u64   count table; region s = sub-table s; slot = (v>>pre)<<10 | count
//   nkeys[P]          u32   distinct keys per sub-table
//   bloom             u8    2^(n_shift-3) bytes; the 64-byte block of hash v sits at
//                     (v & (2^(n_shift-9)-1))*64, i.e. blocks are ordered like the group sort, so a
//                     chunk sweeps the filter front to back (block b of sub-filter s = (b<<pre|s)*64)
//   journal           list of segments; segment = new keys of one chunk ordered by
//                     (sub-table, first-put time) + P+1 offsets.  Concatenating a sub-table's
//                     runs over segments gives the order in which the reference's khashl saw
//                     its distinct keys, which (with `presize` and `trailing`) determines the
//                     .yak slot layout.
//   last_put/last_new[P] u64 time of the last put-event / last new-key put (trailing flag, Q3)
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <vector>
#include <string>
#include <stdexcept>
#include <algorithm>
#include <functional>
#include <cuda_runtime.h>
#include "dbuf.cuh"
#include "radix.cuh"

namespace yakb {

// optional per-kernel timing with CUDA events on the engine's stream (bench.py roofline leg)
struct Prof {
	static void enable(bool on);
	static bool on();
	static void reset();
	static void host(const char *name, double ms);     // host-side time inside the timed region that no kernel accounts for (cudaMalloc)
	static void begin(const char *name, cudaStream_t s);
	static void end(cudaStream_t s);
	static void units(const char *name, uint64_t n);   // work items processed under this name (events, positions)
	static void resolve();                       // call after the stream was synchronised
	static std::string json();                   // {"name": [total_ms, launches, units], ...}
};
struct ProfScope {
	cudaStream_t s; bool live;
	ProfScope(const char *name, cudaStream_t st) : s(st), live(Prof::on()) { if (live) Prof::begin(name, s); }
	~ProfScope() { if (live) Prof::end(s); }
};

struct Segment { uint64_t *keys; uint64_t *off; uint64_t n; };

struct ChunkStats { uint64_t n_events, n_pending, n_put, n_new; };

// result of rebuilding the reference layout for a range of sub-tables (host memory)
struct LayoutOut {
	std::vector<uint32_t> cap, size;    // per sub-table in the range
	std::vector<uint64_t> off;          // size+1 offsets into keys
	std::vector<uint64_t> keys;         // stored keys (with counts) in slot order
};

// which sub-tables a (sharded) engine owns: those whose top `lw` index bits equal `rank`
struct Own { int shift; uint32_t mask, rank; };

struct Engine {
	// P = number of sub-tables held HERE (2^pre, or 2^pre / world for one shard of a multi-GPU
	// table: shard `rank` owns the contiguous sub-table range [rank*P, (rank+1)*P))
	int k = 0, pre = 0, n_hash = 0, n_shift = 0, P = 0;
	int lw = 0, rank = 0;
	Own own() const { Own o; o.shift = pre - lw; o.mask = (1u << lw) - 1; o.rank = (uint32_t)rank; return o; }
	uint64_t *slots = nullptr; uint32_t cap = 0;
	uint32_t *nkeys = nullptr;
	uint8_t *bloom = nullptr; int nb = 0;
	uint64_t *last_put = nullptr, *last_new = nullptr;
	std::vector<uint8_t> presize_flag;      // host
	std::vector<uint32_t> presize_val;      // host
	// Journal entries: low 10 bits 0 = put of key (entry>>10); otherwise an operation on the khashl
	// set that is replayed at that point of the sequence (table set-ops):
	//   1 resize(entry>>10)            htab.c:108 / khashl.h:152
	//   2 load check of a put of an existing key (khashl.h:202-205, quirk Q3)
	//   3 tighten: if (size*3 < capacity) resize(size*3)             htab.c:107-108
	//   4 resize(r) if r > capacity, r = entry>>10                   htab.c:250-254 (merge pre_resize)
	enum { OP_PUT = 0, OP_RESIZE = 1, OP_CHECK = 2, OP_TIGHTEN = 3, OP_RESIZE_IF_LARGER = 4 };
	std::vector<uint32_t> nops;             // host: operation entries per sub-table
	std::vector<uint32_t> max_req;          // host: largest explicit resize request per sub-table
	std::vector<Segment> journal;
	// journal segments are carved from big slabs (a cudaMalloc per chunk costs milliseconds)
	struct Slab { char *p; size_t cap, used; };
	std::vector<Slab> slabs;
	void *journal_alloc(size_t bytes);
	void journal_free_all();
	uint32_t chunk_seq = 0;
	uint64_t tot = 0;
	cudaStream_t stream = nullptr;
	double load_limit = 0.6;
	// scratch (grow-only, reused by every chunk)
	DBuf b_w2, b_wm, b_flags, b_tilecnt, b_tileoff, b_pv, b_ppos, b_sv, b_sj, b_sv2, b_sj2, b_pflag, b_newv, b_newsorted,
	     b_tmp, b_pend, b_lput, b_lnew, b_stats, b_misc, b_lay[12], b_segp, b_zev, b_zpos, b_zsp, b_zspp, b_zfill;
	RadixScratch rs;

	static Engine *create(int k, int pre, int n_hash, int n_shift, int rank = 0, int world = 1);
	~Engine();
	void destroy_bloom();
	void free_slots();                             // the table allocation starts YAKB_SAT_BYTES in front of `slots` (yakb_dev.cuh)

	// ---- per-chunk hot path ----
	// bases: device ASCII (any non-ACGTU byte separates reads), n bytes
	ChunkStats count_ascii(const uint8_t *d_asc, uint64_t n, int create_new);
	// events: device array of hashed k-mers in file order; only_s >= 0 keeps one sub-table (htab.c:61)
	ChunkStats count_events(const uint64_t *d_ev, uint64_t n, int create_new, int only_s, bool ignore_bloom = false);
	// append one operation entry (0 = none) per sub-table to the journal
	void append_ops(const std::vector<uint64_t> &op);
	// replace the table by `keys` (stored form, per sub-table runs given by off) put in that order into
	// khashl sets pre-sized to caps[] (htab.c:183 shrink, 295 subtract, 326 isec, 438 restore)
	void rebuild(const std::vector<uint32_t> &caps, const std::vector<uint64_t> &off, const uint64_t *keys);
	void sizes(std::vector<uint32_t> &out);  // distinct keys per sub-table

	// ---- table-wide ops ----
	void clear();                                  // htab.c:116-130
	void hist(int64_t cnt[1024]);                  // htab.c:136-169
	void shrink(int min, int max);                 // htab.c:175-208
	void get_batch(const uint64_t *d_x, uint64_t n, int32_t *d_out);  // htab.c:93-100, device in/out
	void reserve(uint64_t keys_per_subtable);      // grow regions so every sub-table can take this many
	// rebuild the reference's khashl layout of sub-tables [s0, s1) and bring it to the host
	void layout(int s0, int s1, LayoutOut &out, bool with_counts = true);
	// the same, left on the device: fn(b0, ns, voff, d_dense, cap, size) per batch of sub-tables b0 .. b0+ns (relative to s0) - the stored
	// keys of sub-table b0+t in slot order are d_dense[voff[t] .. voff[t+1]) (device memory, valid inside the call), cap / size as khashl's
	typedef std::function<void(int, int, const std::vector<uint64_t>&, const uint64_t*, const std::vector<uint32_t>&, const std::vector<uint32_t>&)> LayoutFn;
	void layout_device(int s0, int s1, bool with_counts, const LayoutFn &fn);
	// restore: sub-table s gets `keys` (stored form, with counts) in file order, khashl pre-sized to cap
	void load_subtables(const std::vector<uint32_t> &caps, const std::vector<uint64_t> &off, const uint64_t *keys, bool keys_on_device = false);
	// restore into the table as it is (htab.c:436-472): resize every sub-table to caps[s], then put the keys
	// (stored form; low bits = bits to set) in file order; existing keys get the bits OR-ed in when or_bits.
	// Returns the number of keys that were new.
	uint64_t upsert(const std::vector<uint32_t> &caps, const std::vector<uint64_t> &off, const uint64_t *keys, bool or_bits);
	void rebuild_dev(const std::vector<uint32_t> &caps, const std::vector<uint64_t> &off, const uint64_t *d_keys);
	template<class F> void layout_batches(int s0, int s1, bool with_counts, uint64_t reserve_bytes, F &&fn);
	uint64_t device_bytes() const;
	static uint64_t launches();
	static void note_launch(int n);

private:
	ChunkStats finish_chunk(uint64_t n_words, int create_new, const uint64_t *w2, const uint32_t *wm,
	                        const uint64_t *d_ev, uint64_t n_units, int only_s, bool ignore_bloom = false);
	bool probe_partitioned(uint64_t nwords, int create_new, const uint64_t *w2, const uint32_t *wm, const uint64_t *d_ev, uint64_t n_units,
	                       int only_s, uint32_t *flags, uint32_t *tilecnt, uint32_t *lput, unsigned long long *stats, int nsm);
	uint64_t pending_range(uint64_t t0, uint64_t t1, uint32_t off0, uint32_t n_pending, uint64_t nwords, const uint64_t *w2, const uint32_t *wm,
	                       const uint64_t *d_ev, const uint32_t *flags, const uint32_t *tileoff, uint32_t *lput, uint32_t *lnew,
	                       unsigned long long *stats, uint8_t *bloom, int nsm);
	// k-mer events per input position of the last large chunk (sizes the partition lists of the next one)
	double ev_ratio = 1.0;
	void note_ratio(uint64_t n_events, uint64_t n_pos) { if (n_pos >= (1u << 20)) ev_ratio = std::min(1.0, (double)n_events / (double)n_pos * 1.03 + 0.01); }
	void grow(uint32_t new_cap);
	void reset_table(const std::vector<uint64_t> &off);
};

// lookups for qv (qv.c:34-86): per position count (-1 = no k-mer event there), device in/out
void qv_scan_ascii(Engine *e, const uint8_t *d_asc, uint64_t n, int16_t *d_cnt, int raw = 0);

} // namespace yakb
