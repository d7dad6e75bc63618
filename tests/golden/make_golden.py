#!/usr/bin/env python
"""Generate tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref/yak, built from
/root/reference by oracle/Makefile).  Run in the build container (the GPU box has no reference):

    python tests/golden/make_golden.py

Inputs are the seeded synthetic reads of yak_b200/synth.py (pure functions of the seeds below), so
only digests of the reference's output need committing, plus one small .yak file kept whole.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from yak_b200 import synth  # noqa: E402
import oracle_lib  # noqa: E402

INPUTS = {
    # name: (seed_g, G, seed_r, n_reads, fastq)
    "reads_a": (7, 200_000, 11, 6667, False),
    "reads_q": (7, 200_000, 12, 3000, True),
    "cfg1": (101, 1_000_000, 102, 66_667, False),     # BASELINE.json configs[0]: 10 MB FASTA
    "reads_c": (7, 100_000, 13, 10_000, False),       # 15x of a 100 kbp genome: the qv fixture
}
CASES = [
    # (input, k, pre, bf_shift, second input or None)
    ("reads_a", 31, 12, 0, None), ("reads_a", 31, 10, 0, None), ("reads_a", 21, 11, 0, None), ("reads_a", 15, 10, 0, None),
    ("reads_a", 47, 12, 0, None), ("reads_a", 63, 10, 0, None), ("reads_a", 31, 12, 22, None), ("reads_a", 31, 10, 20, None),
    ("reads_a", 31, 12, 24, None), ("reads_a", 27, 11, 21, None), ("reads_a", 63, 10, 21, None), ("reads_a", 31, 12, 12, None),
    ("reads_q", 31, 12, 23, "reads_a"),
    ("cfg1", 31, 12, 0, None), ("cfg1", 31, 12, 24, None),
    ("reads_c", 31, 10, 22, None),
]


def input_bytes(name):
    sg, G, sr, n, fq = INPUTS[name]
    return synth.reads_file_bytes(sg, G, sr, n, fastq=fq)


def main():
    oracle_lib.build()
    tmp = tempfile.mkdtemp()
    files = {}
    out = {"inputs": {}, "cases": []}
    for name in INPUTS:
        data = input_bytes(name)
        files[name] = os.path.join(tmp, name + (".fq" if INPUTS[name][4] else ".fa"))
        open(files[name], "wb").write(data)
        out["inputs"][name] = {"params": INPUTS[name], "sha256": hashlib.sha256(data).hexdigest(), "bytes": len(data)}
    for name, k, pre, b, second in CASES:
        y = os.path.join(tmp, "out.yak")
        cmd = [oracle_lib.REF_YAK, "count", f"-k{k}", f"-p{pre}", "-t4", "-o", y]
        if b:
            cmd.append(f"-b{b}")
        cmd.append(files[name])
        if second:
            cmd.append(files[second])
        subprocess.run(cmd, check=True, capture_output=True)
        data = open(y, "rb").read()
        insp = subprocess.run([oracle_lib.REF_YAK, "inspect", y], check=True, capture_output=True).stdout
        out["cases"].append({"input": name, "second": second, "k": k, "pre": pre, "bf_shift": b, "bytes": len(data),
                             "sha256": hashlib.sha256(data).hexdigest(),
                             "inspect_sha256": hashlib.sha256(insp).hexdigest()})
        if (name, k, pre, b) == ("reads_c", 31, 10, 22):
            open(os.path.join(HERE, "reads_c_k31_p10_b22.yak"), "wb").write(data)
            # qv of contigs (cfg 3 in miniature) against this table: the reference's full stdout
            ctg = os.path.join(tmp, "ctg.fa")
            open(ctg, "wb").write(synth.contigs_bytes(7, 100_000, 3, 8, 20_000, sub=2e-3))
            qv = subprocess.run([oracle_lib.REF_YAK, "qv", "-t1", "-p", y, ctg], check=True, capture_output=True).stdout
            open(os.path.join(HERE, "qv_reads_c_ctg.txt"), "wb").write(qv)
    # primitive KATs straight from the reference's own inline functions (oracle/_ref/libyakref.so)
    import ctypes as C
    R = C.CDLL(oracle_lib.REF_LIB)
    R.ref_hash64.restype = C.c_uint64; R.ref_hash64.argtypes = [C.c_uint64, C.c_uint64]
    R.ref_hash64_64.restype = C.c_uint64; R.ref_hash64_64.argtypes = [C.c_uint64]
    R.ref_hash_long.restype = C.c_uint64; R.ref_hash_long.argtypes = [C.POINTER(C.c_uint64)]
    R.ref_h2b.restype = C.c_uint32; R.ref_h2b.argtypes = [C.c_uint32, C.c_uint32]
    kat = {"hash64": [], "hash64_64": [], "hash_long": [], "h2b": []}
    xs = [0, 1, 0x123456789abcdef, (1 << 62) - 1, 0x2a, 0xdeadbeefcafef00d & ((1 << 62) - 1)]
    for kk in (15, 21, 31):
        m = (1 << 2 * kk) - 1
        for x in xs:
            kat["hash64"].append([x & m, m, R.ref_hash64(x & m, m)])
    for x in xs + [(1 << 64) - 1]:
        kat["hash64_64"].append([x, R.ref_hash64_64(x)])
    for q in ([5, 9, 3, 7], [5, 7, 3, 7], [1, 2, 3, 4], [9, 9, 9, 8], [0, 0, 0, 0]):
        arr = (C.c_uint64 * 4)(*q)
        kat["hash_long"].append([q, R.ref_hash_long(arr)])
    for hsh, bits in ((1, 2), (0xdeadbeef, 20), (12345, 10), (0xffffffff, 31)):
        kat["h2b"].append([hsh, bits, R.ref_h2b(hsh, bits)])
    out["kat"] = kat
    json.dump(out, open(os.path.join(HERE, "golden.json"), "w"), indent=1)
    print("wrote", os.path.join(HERE, "golden.json"), len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
