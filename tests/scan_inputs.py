"""Seeded inputs of the scanner tests (triobin / trioeval / chkerr / sexchr): a 'paternal' genome A, a 'maternal'
genome B = A with 1 % substitutions, a third unrelated genome C, 20x reads of each, and 'child' contigs that
switch between A and B (plus a few substitutions, an N run and a short contig).  Pure functions of the seeds, so
only the reference's stdout needs committing (tests/golden/scan_*.txt, made by tests/golden/make_golden_scan.py)."""
import os

import numpy as np

from yak_b200 import synth

G = 60_000
SEED_A, SEED_C = 21, 23
N_READS = 8000          # 20x of 60 kbp


def genomes():
    a = synth.genome_codes(SEED_A, G)
    rng = np.random.default_rng(5)
    b = a.copy()
    pos = rng.choice(G, size=G // 100, replace=False)
    b[pos] = (b[pos] + rng.integers(1, 4, size=pos.size).astype(np.uint8)) & 3
    c = synth.genome_codes(SEED_C, G)
    return a, b, c


def reads_fasta(genome, seed_r):
    return synth.reads_file_bytes(0, G, seed_r, N_READS, 150, 0.003, 1, fastq=False, genome=genome)


def _fa(name, codes, width=0):
    asc = synth.codes_to_ascii(codes).tobytes()
    if width:
        asc = b"\n".join(asc[p:p + width] for p in range(0, len(asc), width))
    return b">" + name + b" some comment\n" + asc + b"\n"


def child_contigs(a, b, seed=9, n=6):
    """contigs of 8 kbp switching parent every 1-3 kbp, 0.05 % substitutions; one has an N run, one is 25 bp"""
    rng = np.random.default_rng(seed)
    out = bytearray()
    for c in range(n):
        start = int(rng.integers(0, G - 8000))
        segs, p, hap = [], 0, int(rng.integers(0, 2))
        while p < 8000:
            ln = int(min(8000 - p, rng.integers(1000, 3000)))
            segs.append((b if hap else a)[start + p:start + p + ln])
            p += ln
            hap ^= 1
        x = np.concatenate(segs).copy()
        m = rng.random(x.size) < 5e-4
        x[m] = (x[m] + 1) & 3
        asc = bytearray(synth.codes_to_ascii(x).tobytes())
        if c == 2:
            asc[3000:3040] = b"N" * 40
        if c == 3:
            asc = asc.lower()
        out += b">child%d hap\n" % c + (b"\n".join(bytes(asc[q:q + 70]) for q in range(0, len(asc), 70)) if c == 4 else bytes(asc)) + b"\n"
    out += b">tiny\nACGTACGTACGTACGTACGTACGTA\n>empty\n\n"
    return bytes(out)


def bad_fastq(a, b):
    """FASTQ contigs with truncated-quality records between the good ones (quality one character too long, so the bad
    record does not swallow the next header): good bad good bad bad good bad bad good.  The reference's two-worker
    pipeline (kt_pipeline(2, ...)) reads on behind a bad record, loses a worker on every batch that starts with one, and
    never reaches the last record."""
    def rec(name, codes, ok=True):
        s = synth.codes_to_ascii(codes).tobytes()
        return b"@" + name + b"\n" + s + b"\n+\n" + b"I" * (len(s) + (0 if ok else 1)) + b"\n"
    parts = [rec(b"g1", a[100:3100]), rec(b"b1", a[5000:5100], False), rec(b"g2", b[7000:9500]), rec(b"b2", b[100:200], False),
             rec(b"b3", a[300:400], False), rec(b"g3", a[12000:15000]), rec(b"b4", b[400:500], False), rec(b"b5", a[600:700], False),
             rec(b"g4", b[20000:22000])]
    return b"".join(parts)


def write_all(d):
    """writes pat.fa mat.fa third.fa child.fa hapA.fa hapB.fa bad.fq into d; returns their paths"""
    a, b, c = genomes()
    files = {"pat.fa": reads_fasta(a, 31), "mat.fa": reads_fasta(b, 32), "third.fa": reads_fasta(c, 33),
             "child.fa": child_contigs(a, b), "bad.fq": bad_fastq(a, b),
             "hapA.fa": _fa(b"hA1", a[1000:9000]) + _fa(b"hA2", c[500:4500], 60) + _fa(b"hA3", b[20000:26000]),
             "hapB.fa": _fa(b"hB1", b[30000:38000]) + _fa(b"hB2", c[10000:13000])}
    paths = {}
    for name, data in files.items():
        paths[name] = os.path.join(d, "yakb_scan_" + name)
        if not os.path.exists(paths[name]) or os.path.getsize(paths[name]) != len(data):
            with open(paths[name], "wb") as f:
                f.write(data)
    return paths


# (golden file, table-building counts, command line after the executable); {x} = path of x
COUNTS = {"pat.yak": ("pat.fa", 31), "mat.yak": ("mat.fa", 31), "third.yak": ("third.fa", 31),
          "pat47.yak": ("pat.fa", 47), "mat47.yak": ("mat.fa", 47)}
CASES = [
    ("scan_triobin.txt", ["triobin", "-t1", "{pat.yak}", "{mat.yak}", "{child.fa}"]),
    ("scan_triobin_p.txt", ["triobin", "-t1", "-p", "-c3", "-d6", "-r0.5", "{pat.yak}", "{mat.yak}", "{child.fa}"]),
    ("scan_triobin_k47.txt", ["triobin", "-t1", "{pat47.yak}", "{mat47.yak}", "{child.fa}"]),
    ("scan_trioeval.txt", ["trioeval", "-t1", "{pat.yak}", "{mat.yak}", "{child.fa}"]),
    ("scan_trioeval_e.txt", ["trioeval", "-t1", "-e", "-n3", "-c3", "{pat.yak}", "{mat.yak}", "{child.fa}"]),
    ("scan_trioeval_F.txt", ["trioeval", "-t1", "-F", "{mat.yak}", "{pat.yak}", "{child.fa}"]),
    ("scan_chkerr.txt", ["chkerr", "-t1", "{pat.yak}", "{child.fa}"]),
    ("scan_chkerr_c.txt", ["chkerr", "-t1", "-c8", "-s2", "{mat47.yak}", "{child.fa}"]),
    ("scan_sexchr.txt", ["sexchr", "-t1", "{pat.yak}", "{mat.yak}", "{third.yak}", "{hapA.fa}", "{hapB.fa}"]),
    ("scan_sexchr_K.txt", ["sexchr", "-t1", "-K5k", "{third.yak}", "{pat.yak}", "{mat.yak}", "{hapB.fa}", "{hapA.fa}"]),
    # truncated FASTQ records: which records are still read depends on the reference's pipeline (bseq.c:40, kthread.c:119)
    ("scan_chkerr_badq.txt", ["chkerr", "-t1", "{pat.yak}", "{bad.fq}"]),
    ("scan_triobin_badq.txt", ["triobin", "-t1", "{pat.yak}", "{mat.yak}", "{bad.fq}"]),
    ("scan_sexchr_badq.txt", ["sexchr", "-t1", "-K2k", "{pat.yak}", "{mat.yak}", "{third.yak}", "{bad.fq}", "{hapB.fa}"]),
]


def argv(cmd, paths):
    return [paths[a[1:-1]] if a.startswith("{") else a for a in cmd]
