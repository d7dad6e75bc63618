"""GPU side of SURVEY 8(f) rank 2: the lookup scanners (triobin / trioeval / chkerr / sexchr) and the flag modes of
yak_ch_restore_core.  The CLI's stdout must equal the UNMODIFIED reference's (tests/golden/scan_*.txt); the library's
tables after mode loads must dump the same bytes as the oracle's; the batched lookup must equal the oracle's loop."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import scan_inputs as S
import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "yak_b200", "bin", "yak-b200")
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def paths(yakb):
    p = S.write_all(util.TMP)
    for y, (fa, k) in S.COUNTS.items():   # tables built by the CLI under test (byte-exact counts are test_gpu_golden's job)
        p[y] = os.path.join(util.TMP, "yakb_gscan_" + y)
        r = subprocess.run([EXE, "count", f"-k{k}", "-p10", "-t4", "-o", p[y], p[fa]], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    return p


@pytest.mark.parametrize("gold,cmd", S.CASES, ids=[c[0] for c in S.CASES])
def test_cli_scanner_equals_reference_stdout(paths, gold, cmd):
    r = subprocess.run([EXE] + S.argv(cmd, paths), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    assert r.stdout == open(os.path.join(GOLD, gold), "rb").read()


def _oracle_core():
    L = O.lib()
    L.yo_ch_restore_core.restype = C.POINTER(O.YoCh)
    L.yo_ch_restore_core.argtypes = [C.POINTER(O.YoCh), C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]
    L.yo_scan_seq.argtypes = [C.POINTER(O.YoCh), C.c_int64, C.c_char_p, C.POINTER(C.c_int16)]
    return L


@pytest.mark.parametrize("script", [
    [("pat.yak", 2, 2, 5), ("mat.yak", 3, 2, 5)],
    [("pat.yak", 2, 3, 9), ("mat.yak", 3, 1, 4)],
    [("pat.yak", 4), ("mat.yak", 5), ("third.yak", 6)],
    [("third.yak", 4), ("pat.yak", 6), ("mat.yak", 5)],
    [("pat.yak", 1), ("mat.yak", 1), ("third.yak", 1)],        # YAK_LOAD_ALL into an existing table
    [("mat47.yak", 2, 2, 5), ("pat47.yak", 3, 2, 5)],
], ids=["triobin", "triobin-c3d9", "sexchr", "sexchr-perm", "all-into-existing", "triobin-k47"])
def test_restore_core_modes_dump_like_the_oracle(yakb, paths, script):
    """the table after each load: same .yak bytes (khashl layout included: resize to the file's capacity, puts in
    file order, trailing put) and same histogram as the oracle's restatement of htab.c:396-476"""
    L, Lo = yakb.lib(), _oracle_core()
    L.yak_ch_restore_core.restype = C.POINTER(yakb.YakCh)
    hg, ho = None, None
    for st in script:
        fn, mode = paths[st[0]].encode(), st[1]
        a = [C.c_int(x) for x in st[2:]]
        L.yak_ch_restore_core.argtypes = [C.POINTER(yakb.YakCh), C.c_char_p, C.c_int] + [C.c_int] * len(a)
        hg2 = L.yak_ch_restore_core(hg, fn, mode, *a)
        assert hg2
        hg = hg2
        ho = Lo.yo_ch_restore_core(ho, fn, mode, st[2] if len(st) > 2 else 0, st[3] if len(st) > 3 else 0, None)
        mine, ref = yakb.dump_bytes(hg), O.dump_bytes(ho)
        assert mine == ref, util.explain_diff(mine, ref)
        hm, hr = (C.c_int64 * 1024)(), (C.c_int64 * 1024)()
        L.yak_ch_hist(hg, hm, 4)
        Lo.yo_ch_hist(ho, hr)
        assert list(hm) == list(hr)
    L.yak_ch_destroy(hg)
    Lo.yo_ch_destroy(ho)


def test_restore_core_mode_errors(yakb, paths):
    L = yakb.lib()
    L.yak_ch_restore_core.restype = C.POINTER(yakb.YakCh)
    L.yak_ch_restore_core.argtypes = [C.POINTER(yakb.YakCh), C.c_char_p, C.c_int, C.c_int, C.c_int]
    fn = paths["pat.yak"].encode()
    assert not L.yak_ch_restore_core(None, fn, 3, 2, 5)      # TRIOBIN2 without a table (htab.c:412)
    assert not L.yak_ch_restore_core(None, fn, 5, 0, 0)      # SEXCHR2 without a table (htab.c:416)
    assert not L.yak_ch_restore_core(None, fn, 9, 0, 0)      # unknown mode (htab.c:419)
    assert not L.yak_ch_restore_core(None, b"/nonexistent.yak", 1, 0, 0)


@pytest.mark.parametrize("k", [31, 47])
def test_batched_scan_equals_the_oracle_loop(yakb, paths, k):
    L, Lo = yakb.lib(), _oracle_core()
    y = paths["pat.yak" if k == 31 else "pat47.yak"]
    hg = L.yak_ch_restore(y.encode())
    ho = Lo.yo_ch_restore(y.encode())
    seqs = []
    for fn in ("child.fa", "hapA.fa"):
        for rec in open(paths[fn], "rb").read().split(b">")[1:]:
            seqs.append(b"".join(rec.split(b"\n")[1:]))
    seqs += [b"", b"ACGT", b"N" * 50, seqs[0][:k], seqs[0][:k - 1]]
    cat = b"".join(seqs)
    lens = (C.c_int64 * len(seqs))(*[len(s) for s in seqs])
    got = (C.c_int16 * len(cat))()
    assert L.yakb_scan_seqs(hg, len(seqs), lens, cat, got) == 0
    want = (C.c_int16 * len(cat))()
    off = 0
    for s in seqs:
        Lo.yo_scan_seq(ho, len(s), s, C.cast(C.byref(want, off * 2), C.POINTER(C.c_int16)))
        off += len(s)
    g, w = np.frombuffer(got, dtype=np.int16), np.frombuffer(want, dtype=np.int16)
    assert np.array_equal(g, w), np.nonzero(g != w)[0][:10]
    assert (w >= 0).sum() > 10_000 and (w == -1).sum() > 100 and (w == -2).sum() > 100
    L.yak_ch_destroy(hg)
    Lo.yo_ch_destroy(ho)
