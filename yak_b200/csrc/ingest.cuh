// ingest.cuh - host entry point of ingest.cu (FASTA/FASTQ text in the strict 2- / 4-line layout -> dense base stream, on the device)
#pragma once
#include "engine.cuh"

namespace yakb {

struct IngestScratch { DBuf b[3]; RadixScratch rs; void release() { for (DBuf &d : b) d.release(); rs.release(); } };

// d_raw: n bytes of text (n < 2^32) that start at the first byte of a record and end behind the newline of a record;
// lpr = lines per record: 4 (FASTQ, '@') or 2 (FASTA, '>').  d_out (capacity n) receives "SEQ\nSEQ\n...".
// d_res (device, 3 words): [0] = 0 if every record is in the strict layout, else a mask of what was violated (d_out is then
// meaningless and the caller falls back to the host parser); [1] = bytes written to d_out; [2] = lines seen.
// Nothing here waits for the device.
int ingest_strict(const uint8_t *d_raw, uint64_t n, int lpr, uint8_t *d_out, unsigned long long *d_res, cudaStream_t stream, IngestScratch &sc);

} // namespace yakb
