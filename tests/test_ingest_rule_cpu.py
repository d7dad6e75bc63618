"""The rule behind the device-side text ingest (csrc/ingest.cu), checked on the CPU: whenever a file passes the strict-layout check the
device makes per record, "the second line of every record" IS what the reference's parser (kseq.h:192-232, restated in the oracle's reader,
itself pinned to the reference binary) reads from it - for any content of the lines.  The check is restated here line by line from the
kernel (ingest_mark / ingest_lengths / ingest_finish); the GPU tests (tests/test_gpu_ingest.py) hold the kernel to the same outcomes."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
import util


def strict_ok(text: bytes):
    """the device's verdict on a whole file given as one batch; lines-per-record from the first byte"""
    if not text or text[-1:] != b"\n":
        return False, 0
    marker = text[:1]
    lpr = 4 if marker == b"@" else 2 if marker == b">" else 0
    if lpr == 0:
        return False, 0
    lines = text[:-1].split(b"\n")
    if len(lines) % lpr:                                   # a whole number of records (ingest_finish)
        return False, lpr
    for i, ln in enumerate(lines):
        ph = i % lpr
        first = ln[:1]
        if ph == 0 and first != marker:                    # bad |= 2 (and |= 1 for the very first byte)
            return False, lpr
        if ph == 1 and (first in (b">", b"+", b"@", b"\r") or ln == b""):   # bad |= 4: kseq would take it for a header / separator, or skip it
            return False, lpr
        if lpr == 4 and ph == 2 and first != b"+":         # bad |= 8
            return False, lpr
        if lpr == 4 and ph == 3:                           # ingest_lengths: as long as the bases, same line ending
            s = lines[i - 2]
            if len(ln) != len(s) or (len(ln) > 0 and ln[-1:] == b"\r") != (len(s) > 0 and s[-1:] == b"\r"):
                return False, lpr
    return True, lpr


def reader_seqs(path):
    L = O.lib()
    L.yo_reader_open.restype = C.c_void_p; L.yo_reader_open.argtypes = [C.c_char_p]
    L.yo_reader_next.restype = C.c_int64; L.yo_reader_next.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]
    L.yo_reader_close.argtypes = [C.c_void_p]
    r = L.yo_reader_open(path.encode())
    out = []
    while True:
        seq, name = C.c_char_p(), C.c_char_p()
        n = L.yo_reader_next(r, C.byref(seq), C.byref(name))
        if n < 0:
            out.append(int(n))                             # -1 end of file, -2 truncated quality
            break
        out.append(C.string_at(seq, n))
    L.yo_reader_close(r)
    return out


ALPH = np.frombuffer(b"ACGTNacgtn@+>IF#!~ \t\r.-_:/0123456789", dtype=np.uint8)


def _line(rng, n, first=None):
    s = bytes(rng.choice(ALPH, n).astype(np.uint8)).replace(b"\n", b"A")
    if first is not None and n:
        s = first + s[1:]
    return s


def _random_file(rng, fastq):
    recs = []
    for _ in range(int(rng.integers(1, 12))):
        n = int(rng.integers(1, 40))
        seq = _line(rng, n)
        if rng.random() < 0.7:                             # mostly plain bases, so that many files pass
            seq = bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), n).astype(np.uint8))
        crlf = rng.random() < 0.15
        nl = b"\r\n" if crlf else b"\n"
        if fastq:
            qual = _line(rng, n)
            r = rng.random()
            if r < 0.05: qual = qual[:-1]                  # truncated
            elif r < 0.10: qual += b"I"                    # too long
            elif r < 0.13: nl_q = b"\n"                   # (mixed line ends handled below)
            rec = [b"@" + _line(rng, int(rng.integers(0, 10))), seq, b"+" + (_line(rng, 3) if rng.random() < 0.3 else b""), qual]
            body = nl.join(rec) + (b"\n" if (crlf and rng.random() < 0.2) else nl)
        else:
            rec = [b">" + _line(rng, int(rng.integers(0, 10))), seq]
            body = nl.join(rec) + nl
        r = rng.random()
        if r < 0.04: body = b"\n" + body                   # blank line
        elif r < 0.08 and fastq: body = body.replace(b"\n+", b"\n", 1)   # '+' line lost
        elif r < 0.11: body = body.replace(seq, seq[:len(seq) // 2] + nl + seq[len(seq) // 2:], 1)   # bases on two lines
        recs.append(body)
    text = b"".join(recs)
    if rng.random() < 0.05:
        text = text[:-1]                                   # no final newline
    return text


@pytest.mark.parametrize("fastq", [True, False])
def test_files_that_pass_the_strict_check_parse_to_their_second_lines(fastq):
    rng = np.random.default_rng(11 if fastq else 12)
    fn = os.path.join(util.TMP, f"yakb_rule_{int(fastq)}.txt")
    n_pass = n_fail = 0
    for _ in range(3000):
        text = _random_file(rng, fastq)
        ok, lpr = strict_ok(text)
        if not ok:
            n_fail += 1
            continue
        n_pass += 1
        with open(fn, "wb") as f:
            f.write(text)
        lines = text[:-1].split(b"\n")
        want = []
        for s in lines[1::lpr]:
            want.append(s[:-1] if len(s) > 1 and s.endswith(b"\r") else s)     # kseq strips one CR from a line longer than one character
        got = reader_seqs(fn)
        assert got[-1] == -1 and got[:-1] == want, (text, got, want)
    assert n_pass > 500 and n_fail > 300, (n_pass, n_fail)                     # both sides of the rule were exercised
