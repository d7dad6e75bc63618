"""GPU parity of the device-side text ingest (csrc/ingest.cu + count_gpu_ingest in csrc/capi.cu): FASTA/FASTQ files in the strict
2- / 4-line layout are turned into the base stream on the GPU; everything else must be recognised as such and handed to the
host parser (the exact restatement of kseq.h:192-232).  Either way the .yak bytes are the oracle's."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import util
from yak_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "yak_b200", "bin", "yak-b200")


def _cli_count(fn, k, pre, b, batch, fn2=None):
    out = fn + f".k{k}p{pre}b{b}.yak"
    env = dict(os.environ, YAKB_GPU_INGEST="1", YAKB_INGEST_BATCH=str(batch), YAKB_TIMING="1")
    cmd = [CLI, "count", f"-k{k}", f"-p{pre}", f"-b{b}", "-o", out, fn] + ([fn2] if fn2 else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    return open(out, "rb").read(), r.stderr


def _oracle(fn, k, pre, b, fn2=None):
    h, _ = O.count_file(fn, k=k, pre=pre, bf_shift=b, fn2=fn2)
    want = O.dump_bytes(h)
    O.lib().yo_ch_destroy(h)
    return want


def _strict_fastq(path, n=6000, crlf=False, seed=3):
    rng = np.random.default_rng(seed)
    reads = synth.codes_to_ascii(synth.read_codes(7, 300_000, seed, 0, n, 150, 0.01, 3))
    nl = b"\r\n" if crlf else b"\n"
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            L = int(rng.integers(20, 151))                      # ragged lengths, some shorter than k
            qual = bytes(rng.integers(33, 74, L).astype(np.uint8))   # every quality character, '@', '+', '>' at line starts included
            f.write(b"@r%d some comment" % i + nl + bytes(r[:L]) + nl + b"+" + (b"r%d" % i if i % 3 == 0 else b"") + nl + qual + nl)
    return path


@pytest.mark.parametrize("k,pre,b,batch", [(31, 10, 0, 50_000), (31, 12, 22, 300_000), (47, 11, 21, 7_000), (21, 10, 0, 1 << 30)])
def test_strict_fastq_is_parsed_on_the_device(yakb, k, pre, b, batch):
    fn = _strict_fastq(os.path.join(util.TMP, "yakb_ing_strict.fq"))
    got, err = _cli_count(fn, k, pre, b, batch)
    assert "device ingest:" in err and "using the host parser" not in err, err[-2000:]
    if b > 0:   # the second pass (main.c:57) reads the file the first pass has checked: on the device again
        assert err.count("device ingest:") == 2, err[-2000:]
    want = _oracle(fn, k, pre, b)
    assert got == want, util.explain_diff(got, want)


def test_strict_fastq_crlf_and_two_files(yakb):
    fq = _strict_fastq(os.path.join(util.TMP, "yakb_ing_crlf.fq"), n=3000, crlf=True, seed=5)
    fa = os.path.join(util.TMP, "yakb_ing_strict.fa")
    with open(fa, "wb") as f:
        f.write(synth.reads_file_bytes(7, 300_000, 9, 4000))                # ">r<i>\nSEQ\n": the strict 2-line layout
    for fn, fn2 in ((fq, None), (fa, None), (fq, fa)):
        got, err = _cli_count(fn, 31, 12, 21, 40_000, fn2)
        assert "device ingest:" in err, err[-2000:]
        want = _oracle(fn, 31, 12, 21, fn2)
        assert got == want, (fn, fn2, util.explain_diff(got, want))


def _variants(base: bytes):
    recs = base.split(b"\n@r")
    recs = [recs[0]] + [b"@r" + r for r in recs[1:]]
    mid = len(recs) // 2
    out = {}
    out["no-final-newline"] = base[:-1]
    r = recs[mid].split(b"\n")
    out["short-quality"] = b"\n".join(recs[:mid] + [b"\n".join([r[0], r[1], r[2], r[3][:-5]])] + recs[mid + 1:])
    out["long-quality"] = b"\n".join(recs[:mid] + [b"\n".join([r[0], r[1], r[2], r[3] + b"IIII"])] + recs[mid + 1:])
    out["two-line-bases"] = b"\n".join(recs[:mid] + [b"\n".join([r[0], r[1][:40], r[1][40:], r[2], r[3]])] + recs[mid + 1:])
    out["blank-line"] = b"\n".join(recs[:mid] + [b""] + recs[mid:])
    out["missing-plus"] = b"\n".join(recs[:mid] + [b"\n".join([r[0], r[1], r[3]])] + recs[mid + 1:])
    out["fasta-record-inside"] = b"\n".join(recs[:mid] + [b">x\nACGTACGTACGTACGTACGTACGTACGTACGTACGTAAA"] + recs[mid:])
    out["bases-start-with-plus"] = b"\n".join(recs[:mid] + [b"\n".join([r[0], b"+" + r[1][1:], r[2], r[3]])] + recs[mid + 1:])
    out["mixed-line-ends"] = b"\n".join(recs[:mid] + [b"\n".join([r[0], r[1] + b"\r", r[2], r[3][:-1] + b"I"])] + recs[mid + 1:])
    return out


@pytest.mark.parametrize("batch", [30_000, 1 << 30])
def test_irregular_files_fall_back_to_the_host_parser(yakb, batch):
    """every way the strict layout can break, in the middle of a file of good records (so batches before it were already counted
    on the device and the pass starts over): the result is the oracle's, i.e. the reference's"""
    src = _strict_fastq(os.path.join(util.TMP, "yakb_ing_base.fq"), n=1500, seed=8)
    base = open(src, "rb").read()
    for name, data in _variants(base).items():
        fn = os.path.join(util.TMP, f"yakb_ing_{name}.fq")
        with open(fn, "wb") as f:
            f.write(data)
        for b in (0, 20):
            got, err = _cli_count(fn, 31, 10, b, batch)
            want = _oracle(fn, 31, 10, b)
            assert got == want, (name, b, util.explain_diff(got, want))
        assert "device ingest:" not in err, (name, err[-1500:])     # none of these may pass the device check
    # multi-line FASTA (an assembly): refused at the first batch
    fa = os.path.join(util.TMP, "yakb_ing_multi.fa")
    with open(fa, "wb") as f:
        f.write(synth.contigs_bytes(7, 300_000, 3, 5, 30_000, width=60))
    got, err = _cli_count(fa, 31, 10, 0, batch)
    assert "device ingest:" not in err
    assert got == _oracle(fa, 31, 10, 0)
