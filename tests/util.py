"""Shared helpers for the parity tests: seeded inputs and readable .yak diffs."""
import ctypes as C
import os
import struct
import tempfile

import numpy as np

from yak_b200 import synth

TMP = os.environ.get("YAKB_TEST_TMP", tempfile.gettempdir())


def write_reads(path, seed_g=7, G=200_000, seed_r=11, n_reads=6667, fastq=False, **kw):
    data = synth.reads_file_bytes(seed_g, G, seed_r, n_reads, fastq=fastq, **kw)
    with open(path, "wb") as f:
        f.write(data)
    return path


def parse_yak(b: bytes):
    assert b[:4] == b"YAK\x02", b[:4]
    k, pre, cb = struct.unpack_from("<III", b, 4)
    off = 16
    subs = []
    for _ in range(1 << pre):
        cap, size = struct.unpack_from("<II", b, off)
        off += 8
        keys = np.frombuffer(b, dtype="<u8", count=size, offset=off)
        off += 8 * size
        subs.append((cap, size, keys))
    assert off == len(b), (off, len(b))
    return k, pre, cb, subs


def explain_diff(mine: bytes, ref: bytes) -> str:
    if mine == ref:
        return "identical"
    try:
        a, b = parse_yak(mine), parse_yak(ref)
    except Exception as e:  # noqa: BLE001
        return f"unparsable ({e}); len {len(mine)} vs {len(ref)}"
    if a[:3] != b[:3]:
        return f"header {a[:3]} vs {b[:3]}"
    msgs = []
    nbad = 0
    for s, (x, y) in enumerate(zip(a[3], b[3])):
        if x[0] != y[0] or x[1] != y[1] or not np.array_equal(x[2], y[2]):
            nbad += 1
            if len(msgs) < 5:
                same_set = x[1] == y[1] and np.array_equal(np.sort(x[2] >> 10), np.sort(y[2] >> 10))
                same_cnt = x[1] == y[1] and np.array_equal(np.sort(x[2]), np.sort(y[2]))
                first = next((i for i in range(min(x[1], y[1])) if x[2][i] != y[2][i]), None)
                msgs.append(f"sub {s}: cap {x[0]}/{y[0]} size {x[1]}/{y[1]} same_keyset={same_set} "
                            f"same_counts={same_cnt} first_diff_at={first}")
    return f"{nbad} sub-tables differ; " + "; ".join(msgs)


def u64_array(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint64))
