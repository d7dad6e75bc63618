"""Seeded synthetic genomes / reads (SURVEY.md section 8(d) generator).

Counter-based (splitmix64 finaliser) so that every base of every read is a pure function of
(seed, read index, position): the numpy implementation here and the CUDA generator in
csrc/synth.cu produce the identical stream, on any number of ranks, in any order.

    mix(z)        = splitmix64 finaliser of z + 0x9E3779B97F4A7C15
    genome[j]     = mix(seed_g * GMUL + j) >> 62
    a_i           = mix(seed_r * GMUL + i)                      (per-read state)
    start_i       = mix(a_i + 1) % (G - L + 1)
    f_i           = mix(a_i + 2): strand = f_i & 1; has_N = ((f_i >> 8) % 100) < n_pct;
                    N position = (f_i >> 32) % L
    e_ij          = mix(a_i + 16 + j): substitute iff (e_ij & 0xFFFFFF) < err * 2^24,
                    new base = (e_ij >> 24) & 3   (uniform over 4, so 3/4 of them change the base)
"""
from __future__ import annotations

import numpy as np

GMUL = np.uint64(0xD1342543DE82EF95)
_C0 = np.uint64(0x9E3779B97F4A7C15)
_C1 = np.uint64(0xBF58476D1CE4E5B9)
_C2 = np.uint64(0x94D049BB133111EB)
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def mix(z):
    """splitmix64 finaliser, vectorised over uint64 arrays (wraps mod 2^64)."""
    with np.errstate(over="ignore"):
        z = (np.asarray(z, dtype=np.uint64) + _C0).astype(np.uint64)
        z = (z ^ (z >> np.uint64(30))) * _C1
        z = (z ^ (z >> np.uint64(27))) * _C2
        return z ^ (z >> np.uint64(31))


def genome_codes(seed_g: int, G: int, start: int = 0, n: int | None = None) -> np.ndarray:
    """2-bit codes (0..3 = A,C,G,T) of genome[start:start+n]."""
    n = G - start if n is None else n
    with np.errstate(over="ignore"):
        j = np.arange(start, start + n, dtype=np.uint64)
        return (mix(np.uint64(seed_g) * GMUL + j) >> np.uint64(62)).astype(np.uint8)


def read_codes(seed_g: int, G: int, seed_r: int, first: int, n: int, L: int = 150,
               err: float = 0.005, n_pct: int = 1, genome: np.ndarray | None = None) -> np.ndarray:
    """(n, L) uint8 array of codes 0..3, or 4 for an injected N, for reads first..first+n-1."""
    if genome is None:
        genome = genome_codes(seed_g, G)
    with np.errstate(over="ignore"):
        i = np.arange(first, first + n, dtype=np.uint64)
        a = mix(np.uint64(seed_r) * GMUL + i)
        start = (mix(a + np.uint64(1)) % np.uint64(G - L + 1)).astype(np.int64)
        f = mix(a + np.uint64(2))
        strand = (f & np.uint64(1)).astype(bool)
        has_n = ((f >> np.uint64(8)) % np.uint64(100)) < np.uint64(n_pct)
        npos = ((f >> np.uint64(32)) % np.uint64(L)).astype(np.int64)
        j = np.arange(L, dtype=np.int64)
        fwd_idx = start[:, None] + j[None, :]
        rev_idx = start[:, None] + (L - 1 - j)[None, :]
        idx = np.where(strand[:, None], rev_idx, fwd_idx)
        b = genome[idx]
        b = np.where(strand[:, None], 3 - b, b).astype(np.uint8)
        e = mix(a[:, None] + np.uint64(16) + j[None, :].astype(np.uint64))
        thr = np.uint64(int(err * (1 << 24)))
        sub = (e & np.uint64(0xFFFFFF)) < thr
        b = np.where(sub, ((e >> np.uint64(24)) & np.uint64(3)).astype(np.uint8), b)
        rows = np.nonzero(has_n)[0]
        b[rows, npos[rows]] = 4
    return b


def codes_to_ascii(codes: np.ndarray) -> np.ndarray:
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    return lut[codes]


def _records(names: list[bytes], seqs: np.ndarray, fastq: bool) -> bytes:
    out = bytearray()
    L = seqs.shape[1]
    qual = b"I" * L
    for nm, row in zip(names, seqs):
        if fastq:
            out += b"@" + nm + b"\n" + row.tobytes() + b"\n+\n" + qual + b"\n"
        else:
            out += b">" + nm + b"\n" + row.tobytes() + b"\n"
    return bytes(out)


def reads_file_bytes(seed_g: int, G: int, seed_r: int, n_reads: int, L: int = 150, err: float = 0.005,
                     n_pct: int = 1, fastq: bool = False, first: int = 0, genome: np.ndarray | None = None,
                     block: int = 1 << 16) -> bytes:
    """FASTA (``>r<i>\\nSEQ\\n``) or FASTQ (constant quality ``I``) text of reads first..first+n-1."""
    if genome is None:
        genome = genome_codes(seed_g, G)
    parts = []
    for s in range(first, first + n_reads, block):
        m = min(block, first + n_reads - s)
        asc = codes_to_ascii(read_codes(seed_g, G, seed_r, s, m, L, err, n_pct, genome))
        # vectorised assembly per name-length group
        ids = np.arange(s, s + m)
        names = np.char.add("r", ids.astype(str)).astype("S")
        nl = np.char.str_len(names)
        for length in np.unique(nl):
            sel = np.nonzero(nl == length)[0]
            k = len(sel)
            if fastq:
                rec = np.empty((k, 1 + length + 1 + L + 3 + L + 1), dtype=np.uint8)
                rec[:, 0] = ord("@")
            else:
                rec = np.empty((k, 1 + length + 1 + L + 1), dtype=np.uint8)
                rec[:, 0] = ord(">")
            nm = np.frombuffer(names[sel].astype(f"S{length}").tobytes(), dtype=np.uint8).reshape(k, length)
            rec[:, 1:1 + length] = nm
            rec[:, 1 + length] = 10
            rec[:, 2 + length:2 + length + L] = asc[sel]
            rec[:, 2 + length + L] = 10
            if fastq:
                o = 3 + length + L
                rec[:, o] = ord("+"); rec[:, o + 1] = 10
                rec[:, o + 2:o + 2 + L] = ord("I")
                rec[:, o + 2 + L] = 10
            parts.append((sel + s, rec))
    # restore read order across the name-length groups (groups are contiguous id ranges except
    # at powers of ten, so a stable sort by first id of each group suffices per block)
    parts.sort(key=lambda p: int(p[0][0]))
    return b"".join(p[1].tobytes() for p in parts)


def count_events(codes: np.ndarray, k: int) -> int:
    """Number of k-mer events (windows of k valid bases) in an (n, L) code matrix."""
    valid = (codes < 4).astype(np.int32)
    c = np.cumsum(valid, axis=1)
    L = codes.shape[1]
    if L < k:
        return 0
    win = c[:, k - 1:] - np.concatenate([np.zeros((codes.shape[0], 1), np.int32), c[:, :L - k]], axis=1)
    return int((win == k).sum())


def contigs_bytes(seed_g: int, G: int, seed_c: int, n_contigs: int, length: int, sub: float = 1e-4,
                  width: int = 0, genome: np.ndarray | None = None) -> bytes:
    """FASTA of n contigs cut from the genome with `sub` substitutions (cfg 3 queries)."""
    if genome is None:
        genome = genome_codes(seed_g, G)
    out = bytearray()
    with np.errstate(over="ignore"):
        for c in range(n_contigs):
            a = mix(np.uint64(seed_c) * GMUL + np.uint64(c))
            start = int(mix(a + np.uint64(1)) % np.uint64(G - length + 1))
            b = genome[start:start + length].copy()
            e = mix(a + np.uint64(16) + np.arange(length, dtype=np.uint64))
            thr = np.uint64(int(sub * (1 << 24)))
            m = (e & np.uint64(0xFFFFFF)) < thr
            b[m] = (b[m] + 1 + ((e[m] >> np.uint64(24)) % np.uint64(3)).astype(np.uint8)) & 3
            asc = codes_to_ascii(b).tobytes()
            out += b">ctg%d\n" % c
            if width > 0:
                for p in range(0, length, width):
                    out += asc[p:p + width] + b"\n"
            else:
                out += asc + b"\n"
    return bytes(out)
