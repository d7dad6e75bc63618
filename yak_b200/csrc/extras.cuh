// extras.cuh - host entry points of extras.cu
#pragma once
#include "engine.cuh"

namespace yakb {

// hashed canonical k-mers of a device ASCII stream in file order, stably grouped by owner rank;
// the grow-only scratch is owned by the caller
struct RouteScratch { DBuf b[8]; RadixScratch rs; };
int extract_events(const uint8_t *d_asc, uint64_t n, int k, int pre, int world, uint64_t *d_out, uint64_t *counts,
                   cudaStream_t stream, RouteScratch &sc);
// the same without waiting for the device: d_counts (device, world entries) receives the events per owner
int extract_events_async(const uint8_t *d_asc, uint64_t n, int k, int pre, int world, uint64_t *d_out, uint64_t *d_counts,
                         cudaStream_t stream, RouteScratch &sc);
// n = positions in the batch (= seq_off[n_seq]); every sequence is followed by one separator position
void qv_stats(const int16_t *d_cnt, const uint64_t *d_seq_off, uint64_t n_seq, uint64_t n, int min_len, double min_frac,
              int32_t *d_tot, int32_t *d_non0, uint8_t *d_pass, unsigned long long *d_hist, cudaStream_t stream);
int inc_one(Engine *e, uint64_t v);
void setcnt(Engine *e, int c);
int bf_insert_one(uint8_t *d_bits, int n_shift, int n_hashes, uint64_t hash);
void synth_genome(uint64_t seed_g, uint64_t G, uint64_t *d_g2, cudaStream_t stream);
void synth_reads(const uint64_t *d_g2, uint64_t G, uint64_t seed_r, uint64_t first, uint64_t n_reads, int L, double err, int n_pct,
                 int fmt, uint8_t *d_asc, cudaStream_t stream);

} // namespace yakb
