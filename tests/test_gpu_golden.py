"""GPU parity against the committed golden digests of the UNMODIFIED reference's output
(tests/golden/golden.json, produced by tests/golden/make_golden.py), through the C ABI and the CLI."""
import hashlib
import os
import subprocess

import pytest

import golden_util as G
import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "yak_b200", "bin", "yak-b200")


@pytest.mark.parametrize("case", G.GOLD["cases"], ids=G.case_id)
def test_count_equals_reference_digest(yakb, case):
    fn = G.input_path(case["input"])
    fn2 = G.input_path(case["second"]) if case["second"] else None
    h = yakb.count_file(fn, k=case["k"], pre=case["pre"], bf_shift=case["bf_shift"], fn2=fn2)
    assert h
    data = yakb.dump_bytes(h)
    yakb.lib().yak_ch_destroy(h)
    assert len(data) == case["bytes"]
    assert hashlib.sha256(data).hexdigest() == case["sha256"]


def test_cli_count_inspect_qv_equal_reference_output():
    """`yak count -b -o`, `yak inspect`, `yak qv -p` of the CLI: file bytes and stdout identical."""
    case = next(c for c in G.GOLD["cases"] if c["input"] == "reads_c")
    fn = G.input_path("reads_c")
    y = os.path.join(util.TMP, "yakb_cli_out.yak")
    r = subprocess.run([EXE, "count", f"-k{case['k']}", f"-p{case['pre']}", f"-b{case['bf_shift']}", "-t4", "-o", y, fn],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "distinct k-mers after shrinking" in r.stderr and "Real time" in r.stderr
    data = open(y, "rb").read()
    assert hashlib.sha256(data).hexdigest() == case["sha256"]
    assert data == open(os.path.join(G.HERE, "reads_c_k31_p10_b22.yak"), "rb").read()
    r = subprocess.run([EXE, "inspect", y], capture_output=True, text=True)
    assert hashlib.sha256(r.stdout.encode()).hexdigest() == case["inspect_sha256"]
    from yak_b200 import synth
    ctg = os.path.join(util.TMP, "yakb_ctg.fa")
    open(ctg, "wb").write(synth.contigs_bytes(7, 100_000, 3, 8, 20_000, sub=2e-3))
    r = subprocess.run([EXE, "qv", "-t1", "-p", y, ctg], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout == open(os.path.join(G.HERE, "qv_reads_c_ctg.txt")).read()
    # stdout dump ("-") carries the same bytes (htab.c:378)
    r = subprocess.run([EXE, "count", "-k31", "-p10", "-b22", "-o", "-", fn], capture_output=True)
    assert r.returncode == 0 and hashlib.sha256(r.stdout).hexdigest() == case["sha256"]


def test_device_generator_equals_numpy_generator(yakb):
    import ctypes as C
    import numpy as np
    import torch
    from yak_b200 import synth
    L = yakb.lib()
    Gn, n, Lr = 300_000, 5000, 150
    g2 = torch.empty((Gn + 31) // 32 + 1, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    L.yakb_synth_genome_dev(5, Gn, g2.data_ptr(), st)
    for fmt, fastq in ((1, False), (2, True)):
        rec = Lr + 4 if fmt == 1 else 2 * Lr + 7
        buf = torch.empty(n * rec, dtype=torch.uint8, device="cuda")
        L.yakb_synth_reads_dev(g2.data_ptr(), Gn, 9, 100, n, Lr, 0.005, 1, fmt, buf.data_ptr(), st)
        torch.cuda.synchronize()
        got = buf.cpu().numpy().reshape(n, rec)
        codes = synth.read_codes(5, Gn, 9, 100, n, Lr, 0.005, 1)
        want = synth.codes_to_ascii(codes)
        assert np.array_equal(got[:, 3:3 + Lr], want)
        assert bytes(got[0, :3]) == (b"@r\n" if fastq else b">r\n")
