// ingest.cu - FASTA/FASTQ text to the dense "SEQ\nSEQ\n..." base stream ON THE GPU, for files in the strict layout
// sequencers write: every record exactly 4 lines (FASTQ: @name / bases / + / qualities of the same length) or exactly
// 2 lines (FASTA: >name / bases).  The host parser (fastx_par.cpp, the exact restatement of kseq.h:192-232 for ANY input)
// tops out at ~10-20 GB/s of text on the host's cores; here the host only moves bytes, and the line structure is resolved
// with two prefix sums on the device at a few hundred GB/s.
//
// What kseq_read does with such a record (kseq.h:192-232): the header line is skipped to its newline; sequence lines are
// collected until a line starts with '>', '+' or '@'; after '+', quality lines are read until they hold as many characters
// as the sequence; the next record is then searched from the next '@' / '>'.  For a file in the strict layout this is
// "the second line of every record" - PROVIDED the layout really holds, which the device checks for every record:
//   * line 0 of a record starts with the marker ('@' / '>'), line 2 (FASTQ) with '+'
//   * line 1 is not empty and does not start with '>', '+', '@'  (kseq would take it for a header / separator)
//   * FASTQ: line 3 is as long as line 1, and both end the same way (CR or no CR: kseq strips one '\r' per line)
//   * the batch ends with the newline that closes a record
// One violation anywhere and the caller falls back to the host parser for the whole pass (capi.cu): this path never has
// to reproduce kseq's behaviour on irregular input, it only has to recognise regular input with certainty.
//
// Kernels (tile = 4096 bytes, 256 threads x 16 bytes):
//   ingest_nl_count   newlines per tile                                   -> scan = line number at every tile start
//   ingest_mark       per byte: its line number; per tile: bytes of sequence lines (their '\n' included); all checks
//   ingest_scatter    the bytes of the sequence lines, densely, in file order
#include "ingest.cuh"
#include "kernels.cuh"
#include <stdio.h>

namespace yakb {

static inline uint32_t cdiv64(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

#define ING_TILE 4096

// 16 bytes of the thread as a bit mask of newlines (bit i = byte i is '\n'); bytes at or beyond n read as 0
__device__ __forceinline__ uint32_t ing_load16(const uint8_t *__restrict__ raw, uint64_t off, uint64_t n, uint8_t (&c)[16])
{
	uint32_t m = 0;
	if (off + 16 <= n && ((uintptr_t)(raw + off) & 15) == 0) {
		const uint4 q = __ldg((const uint4*)(raw + off));
		const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
		for (int i = 0; i < 16; ++i) c[i] = (uint8_t)(u[i >> 2] >> (8 * (i & 3)));
	} else {
#pragma unroll
		for (int i = 0; i < 16; ++i) c[i] = off + i < n ? raw[off + i] : 0;
	}
#pragma unroll
	for (int i = 0; i < 16; ++i) if (c[i] == '\n' && off + i < n) m |= 1u << i;
	return m;
}

__global__ void __launch_bounds__(256) ingest_nl_count(const uint8_t *__restrict__ raw, uint64_t n, uint32_t *__restrict__ cnt)
{
	uint8_t c[16];
	const uint64_t off = blockIdx.x * (uint64_t)ING_TILE + threadIdx.x * 16ull;
	const uint32_t m = ing_load16(raw, off, n, c);
	uint32_t tot;
	block_excl_scan_256(__popc(m), &tot);
	if (threadIdx.x == 0) cnt[blockIdx.x] = tot;
}

// line number (from the start of the batch) of the first byte of the thread's 16 bytes
__device__ __forceinline__ uint32_t ing_line0(const uint32_t *__restrict__ tile_line, uint32_t nl_mask)
{
	return tile_line[blockIdx.x] + block_excl_scan_256(__popc(nl_mask), nullptr);
}

// res[0] |= violation bits; res[1] = bytes kept (set by the host from the scan), res[2] = newlines
__global__ void __launch_bounds__(256) ingest_mark(const uint8_t *__restrict__ raw, uint64_t n, int lpr, uint8_t marker,
                                                   const uint32_t *__restrict__ tile_line, uint32_t *__restrict__ keepcnt,
                                                   uint32_t *__restrict__ nlpos, uint32_t nl_cap, unsigned long long *res)
{
	uint8_t c[16];
	const uint64_t off = blockIdx.x * (uint64_t)ING_TILE + threadIdx.x * 16ull;
	const uint32_t m = ing_load16(raw, off, n, c);
	uint32_t line = ing_line0(tile_line, m);
	uint32_t keep = 0, bad = 0;
	const uint32_t lmask = (uint32_t)lpr - 1; // lpr is 2 or 4
	if (off == 0 && n && c[0] != marker) bad |= 1;
#pragma unroll
	for (int i = 0; i < 16; ++i) {
		if (off + i >= n) break;
		if ((line & lmask) == 1) ++keep;
		if (m >> i & 1) {
			const uint64_t p = off + i;         // this newline closes line `line`
			const uint32_t nxt = (line + 1) & lmask;
			if (p + 1 < n) {
				const uint8_t f = i < 15 ? c[i + 1] : raw[p + 1];
				if (nxt == 0 && f != marker) bad |= 2;
				if (nxt == 1 && (f == '>' || f == '+' || f == '@' || f == '\n' || f == '\r')) bad |= 4;
				if (lpr == 4 && nxt == 2 && f != '+') bad |= 8;
			} else if (nxt != 0) bad |= 16;     // the batch must end with the newline that closes a record
			if (lpr == 4 && line < nl_cap) nlpos[line] = (uint32_t)p; // line lengths are compared per record by ingest_lengths
			++line;
		}
	}
	if (off + 16 >= n && off < n && !(m >> ((n - 1 - off) & 15) & 1)) bad |= 128; // the last byte of the batch is a newline
	uint32_t tot;
	block_excl_scan_256(keep, &tot);
	if (threadIdx.x == 0) keepcnt[blockIdx.x] = tot;
	if (bad) atomicOr(res, (unsigned long long)bad);
}

__global__ void __launch_bounds__(256) ingest_scatter(const uint8_t *__restrict__ raw, uint64_t n, int lpr,
                                                      const uint32_t *__restrict__ tile_line, const uint32_t *__restrict__ keepoff,
                                                      uint8_t *__restrict__ out)
{
	// the tile's kept bytes are first packed in shared memory, then leave as one contiguous run: consecutive lanes write
	// consecutive bytes / words (a thread storing its own 16 bytes one by one cost 16 scattered sector writes per warp store)
	__shared__ __align__(16) uint8_t s_out[ING_TILE + 16];
	uint8_t c[16];
	const uint64_t off = blockIdx.x * (uint64_t)ING_TILE + threadIdx.x * 16ull;
	const uint32_t m = ing_load16(raw, off, n, c);
	uint32_t line = ing_line0(tile_line, m);
	const uint32_t lmask = (uint32_t)lpr - 1;
	uint32_t keep = 0, l2 = line;
#pragma unroll
	for (int i = 0; i < 16; ++i) {
		if (off + i >= n) break;
		if ((l2 & lmask) == 1) ++keep;
		if (m >> i & 1) ++l2;
	}
	uint32_t tot;
	uint32_t o = block_excl_scan_256(keep, &tot);
	const uint64_t gbase = keepoff[blockIdx.x];
	const uint32_t lead = (uint32_t)((4 - (gbase & 3)) & 3); // bytes in front of the first 4-byte aligned output address
	// place so that shared-memory word w (w >= 1) is output word (gbase + lead) / 4 + w - 1: byte j of the run sits at 4 - lead + j
	const uint32_t shift = 4 - lead;
#pragma unroll
	for (int i = 0; i < 16; ++i) {
		if (off + i >= n) break;
		if ((line & lmask) == 1) s_out[shift + o++] = c[i];
		if (m >> i & 1) ++line;
	}
	__syncthreads();
	// head bytes, whole words, tail bytes
	if (threadIdx.x < lead && threadIdx.x < tot) out[gbase + threadIdx.x] = s_out[shift + threadIdx.x];
	if (tot > lead) {
		const uint32_t body = tot - lead, nw = body / 4;
		uint32_t *ow = (uint32_t*)(out + gbase + lead);
		const uint32_t *sw = (const uint32_t*)(s_out + 4);
		for (uint32_t w = threadIdx.x; w < nw; w += 256) ow[w] = sw[w];
		const uint32_t done = lead + nw * 4;
		if (threadIdx.x < tot - done) out[gbase + done + threadIdx.x] = s_out[shift + done + threadIdx.x];
	}
}

// FASTQ: the quality line of every record is as long as its bases, and both end the same way (kseq strips one CR per line).
// nlpos[l] = offset of the newline that closes line l
__global__ void __launch_bounds__(256) ingest_lengths(const uint8_t *__restrict__ raw, const uint32_t *__restrict__ nlpos,
                                                      const uint32_t *__restrict__ tile_line, uint32_t ntiles, uint32_t nl_cap, unsigned long long *res)
{
	const uint32_t lines = min(tile_line[ntiles], nl_cap);
	const uint64_t r = blockIdx.x * 256ull + threadIdx.x;
	if (r * 4 + 3 >= lines) return;
	const int64_t p0 = nlpos[r * 4], p1 = nlpos[r * 4 + 1], p2 = nlpos[r * 4 + 2], p3 = nlpos[r * 4 + 3];
	const int64_t slen = p1 - p0 - 1, qlen = p3 - p2 - 1;
	const bool cr_s = slen > 0 && raw[p1 - 1] == '\r', cr_q = qlen > 0 && raw[p3 - 1] == '\r';
	if (slen != qlen || cr_s != cr_q) atomicOr(res, 64ull);
}

__global__ void ingest_finish(const uint32_t *__restrict__ tile_line, const uint32_t *__restrict__ keepoff, uint32_t ntiles, int lpr,
                              uint32_t nl_cap, unsigned long long *res)
{
	res[1] = keepoff[ntiles];
	res[2] = tile_line[ntiles];
	if (tile_line[ntiles] % (uint32_t)lpr) atomicOr(res, 256ull); // a whole number of records
	if (lpr == 4 && tile_line[ntiles] > nl_cap) atomicOr(res, 512ull); // lines of fewer than 16 bytes on average: not what this path is for
}

int ingest_strict(const uint8_t *d_raw, uint64_t n, int lpr, uint8_t *d_out, unsigned long long *d_res, cudaStream_t stream, IngestScratch &sc)
{
	YAKB_CUDA(cudaMemsetAsync(d_res, 0, 3 * sizeof(unsigned long long), stream));
	if (n == 0) return 0;
	if ((lpr != 2 && lpr != 4) || n >= 0xFFFFFF00ull) return -1;
	const uint32_t ntiles = cdiv64(n, ING_TILE);
	uint32_t *tile_line = sc.b[0].as<uint32_t>((uint64_t)ntiles + 1), *keep = sc.b[1].as<uint32_t>((uint64_t)ntiles + 1);
	ProfScope ps("ingest", stream);
	YAKB_CUDA(cudaMemsetAsync(tile_line + ntiles, 0, 4, stream));
	YAKB_CUDA(cudaMemsetAsync(keep + ntiles, 0, 4, stream));
	ingest_nl_count<<<ntiles, 256, 0, stream>>>(d_raw, n, tile_line);
	exclusive_scan_u32(tile_line, tile_line, (uint64_t)ntiles + 1, stream, sc.rs);
	const uint32_t nl_cap = (uint32_t)(n / 16 + 1024);
	uint32_t *nlpos = lpr == 4 ? sc.b[2].as<uint32_t>(nl_cap) : nullptr;
	ingest_mark<<<ntiles, 256, 0, stream>>>(d_raw, n, lpr, lpr == 4 ? '@' : '>', tile_line, keep, nlpos, nl_cap, d_res);
	if (lpr == 4) ingest_lengths<<<cdiv64((uint64_t)nl_cap / 4 + 1, 256), 256, 0, stream>>>(d_raw, nlpos, tile_line, ntiles, nl_cap, d_res);
	exclusive_scan_u32(keep, keep, (uint64_t)ntiles + 1, stream, sc.rs);
	ingest_scatter<<<ntiles, 256, 0, stream>>>(d_raw, n, lpr, tile_line, keep, d_out);
	ingest_finish<<<1, 1, 0, stream>>>(tile_line, keep, ntiles, lpr, nl_cap, d_res);
	YAKB_CUDA(cudaGetLastError());
	Engine::note_launch(lpr == 4 ? 5 : 4);
	Prof::units("ingest", n);
	return 0;
}

} // namespace yakb
