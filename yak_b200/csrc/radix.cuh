// radix.cuh - our own device primitives for the ordered parts of the path: exclusive scan,
// stable LSD radix sort of (u64 key, u32 payload) records on a bit range, ordered compaction.
// Stability is what the path needs (file order inside a sub-table / bloom block), so every pass
// ranks elements in (warp, round, lane) order inside a tile and tiles in index order.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "dbuf.cuh"

namespace yakb {

struct RadixScratch {
	DBuf hist, lvl[4];
	void release() { hist.release(); for (DBuf &b : lvl) b.release(); }
};
// our own kernels launched by these helpers are counted here (Engine::launches adds them)
void radix_note_launch(int n);
uint64_t radix_launches();

// out[i] = sum of in[0..i) ; out may alias in.  n up to 2^32.
void exclusive_scan_u32(const uint32_t *in, uint32_t *out, uint64_t n, cudaStream_t st, RadixScratch &rs);

// Stable sort on key bits [begin_bit, end_bit).  src keys are left untouched; (a,b) are two work
// buffers of n records each.  Payloads start as 0..n-1 when v_src is null.  Returns which work
// buffer holds the result (0 = a, 1 = b).  With end_bit <= begin_bit the records are copied to a.
int radix_sort_pairs(const uint64_t *k_src, const uint32_t *v_src, uint64_t *k_a, uint32_t *v_a, uint64_t *k_b, uint32_t *v_b,
                     uint64_t n, int begin_bit, int end_bit, cudaStream_t st, RadixScratch &rs);

// out = in[i] for every i with flag[i] != 0, order kept; *d_count (device) receives the number kept.
void compact_flagged_u64(const uint64_t *in, const uint8_t *flag, uint64_t n, uint64_t *out, uint32_t *d_count,
                         cudaStream_t st, RadixScratch &rs);

} // namespace yakb
