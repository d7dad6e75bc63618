"""Table set-ops: tighten / setcnt / merge / subtract / isec.

CPU: the oracle's restatements against the reference's OWN functions called through
oracle/_ref/libyakref.so (skipped where oracle/_ref is not built).
GPU: the library against the oracle, byte-exact dumps after each operation.
"""
import ctypes as C
import os

import pytest

import golden_util as G
import oracle_lib as O
import util


def _yak_files():
    """two overlapping tables as .yak files written by the oracle"""
    fa, fq = G.input_path("reads_a"), G.input_path("reads_q")
    out = []
    for tag, fn, b in (("A", fa, 0), ("B", fq, 0), ("C", fa, 22)):
        p = os.path.join(util.TMP, f"yakb_setops_{tag}.yak")
        if not os.path.exists(p):
            h, _ = O.count_file(fn, k=31, pre=10, bf_shift=b)
            assert O.lib().yo_ch_dump(h, p.encode()) == 0
            O.lib().yo_ch_destroy(h)
        out.append(p)
    return out


SCRIPTS = [
    [("tighten",)],
    [("merge", "B", 0, 1023, 0)],
    [("merge", "B", 2, 6, 1)],
    [("merge", "B", 1, 1023, 1), ("tighten",), ("merge", "C", 0, 1023, 0)],
    [("subtract", "B")],
    [("isec", "B")],
    [("isec", "C"), ("setcnt", 7), ("merge", "B", 0, 1023, 1), ("subtract", "C"), ("tighten",)],
    [("shrink", 2, 5), ("merge", "B", 0, 1023, 0), ("shrink", 2, 1023), ("tighten",)],
]


def _run_oracle(script, files):
    L = O.lib()
    fa, fb, fc = files
    name = {"A": fa, "B": fb, "C": fc}
    h = L.yo_ch_restore(fa.encode())
    outs = []
    for st in script:
        if st[0] == "tighten": L.yo_ch_tighten(h)
        elif st[0] == "setcnt": L.yo_ch_setcnt(h, st[1])
        elif st[0] == "shrink": L.yo_ch_shrink(h, st[1], st[2])
        elif st[0] == "merge": L.yo_ch_merge(h, L.yo_ch_restore(name[st[1]].encode()), st[2], st[3], st[4])
        else:
            o = L.yo_ch_restore(name[st[1]].encode())
            (L.yo_ch_subtract if st[0] == "subtract" else L.yo_ch_isec)(h, o)
            L.yo_ch_destroy(o)
        outs.append((O.dump_bytes(h), h.contents.tot))
    L.yo_ch_destroy(h)
    return outs


def _run_clib(L, prefix, script, files, dump):
    """the same script on a library with the reference's API (the reference itself or libyakb200)"""
    fa, fb, fc = files
    name = {"A": fa, "B": fb, "C": fc}
    h = L.yak_ch_restore(fa.encode())
    assert h
    outs = []
    for st in script:
        if st[0] == "tighten": L.yak_ch_tighten(h)
        elif st[0] == "setcnt": L.yak_ch_setcnt(h, st[1], 2)
        elif st[0] == "shrink": L.yak_ch_shrink(h, st[1], st[2], 2)
        elif st[0] == "merge": L.yak_ch_merge(h, L.yak_ch_restore(name[st[1]].encode()), st[2], st[3], 2, st[4])
        else:
            o = L.yak_ch_restore(name[st[1]].encode())
            (L.yak_ch_subtract if st[0] == "subtract" else L.yak_ch_isec)(h, o, 2)
            L.yak_ch_destroy(o)
        outs.append((dump(h), None))
    L.yak_ch_destroy(h)
    return outs


@pytest.mark.skipif(not os.path.exists(O.REF_LIB), reason="oracle/_ref not built")
@pytest.mark.parametrize("script", SCRIPTS, ids=lambda s: "+".join(x[0] for x in s))
def test_oracle_setops_equal_reference_functions(script):
    from yak_b200.capi import YakCh
    R = C.CDLL(O.REF_LIB)
    ChP = C.POINTER(YakCh)
    R.yak_ch_restore.restype = ChP; R.yak_ch_restore.argtypes = [C.c_char_p]
    R.yak_ch_dump.argtypes = [ChP, C.c_char_p]
    R.yak_ch_destroy.argtypes = [ChP]
    R.yak_ch_tighten.argtypes = [ChP]
    R.yak_ch_setcnt.argtypes = [ChP, C.c_int, C.c_int]
    R.yak_ch_shrink.argtypes = [ChP, C.c_int, C.c_int, C.c_int]
    R.yak_ch_merge.argtypes = [ChP, ChP, C.c_int, C.c_int, C.c_int, C.c_int]
    R.yak_ch_subtract.argtypes = [ChP, ChP, C.c_int]
    R.yak_ch_isec.argtypes = [ChP, ChP, C.c_int]
    tmp = os.path.join(util.TMP, "yakb_setops_ref.yak")

    def dump(h):
        R.yak_ch_dump(h, tmp.encode())
        return open(tmp, "rb").read()
    files = _yak_files()
    want = _run_clib(R, "ref", script, files, dump)
    got = _run_oracle(script, files)
    for i, ((g, _), (w, _)) in enumerate(zip(got, want)):
        assert g == w, f"step {i} {script[i]}: " + util.explain_diff(g, w)


@pytest.mark.gpu
@pytest.mark.parametrize("script", SCRIPTS, ids=lambda s: "+".join(x[0] for x in s))
def test_gpu_setops_equal_oracle(yakb, script):
    files = _yak_files()
    want = _run_oracle(script, files)
    got = _run_clib(yakb.lib(), "yakb", script, files, yakb.dump_bytes)
    for i, ((g, _), (w, _)) in enumerate(zip(got, want)):
        assert g == w, f"step {i} {script[i]}: " + util.explain_diff(g, w)


@pytest.mark.gpu
def test_gpu_counting_continues_after_setops(oracle, yakb):
    """merge + tighten in the middle of counting: later inserts keep replaying on top of the recorded operations"""
    import numpy as np
    OL, L = oracle.lib(), yakb.lib()
    files = _yak_files()
    ho, hg = OL.yo_ch_restore(files[0].encode()), L.yak_ch_restore(files[0].encode())
    OL.yo_ch_merge(ho, OL.yo_ch_restore(files[1].encode()), 0, 1023, 1)
    L.yak_ch_merge(hg, L.yak_ch_restore(files[1].encode()), 0, 1023, 2, 1)
    OL.yo_ch_tighten(ho); L.yak_ch_tighten(hg)
    seqs = [ln.strip() for ln in open(G.input_path("reads_c")) if not ln.startswith(">")][:800]
    ev = []
    for s in seqs:
        buf = (C.c_uint64 * len(s))()
        n = OL.yo_extract(31, len(s), s.encode(), buf)
        ev.append(np.frombuffer(buf, dtype=np.uint64, count=n).copy())
    ev = np.concatenate(ev)
    sub = (ev & np.uint64(1023)).astype(np.int64)
    order = np.argsort(sub, kind="stable")
    bounds = np.flatnonzero(np.diff(sub[order])) + 1
    for lst in np.split(ev[order], bounds):
        a, p = util.u64_array(lst)
        assert OL.yo_ch_insert_list(ho, 1, len(a), p) == L.yak_ch_insert_list(hg, 1, len(a), p)
    assert yakb.dump_bytes(hg) == oracle.dump_bytes(ho)
    OL.yo_ch_tighten(ho); L.yak_ch_tighten(hg)
    assert yakb.dump_bytes(hg) == oracle.dump_bytes(ho)
    L.yak_ch_destroy(hg); OL.yo_ch_destroy(ho)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(O.REF_YAK), reason="oracle/_ref not built")
def test_cli_setop_commands_equal_reference_binary():
    """recount / subtract / isec / print of the CLI against the reference binary's output"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "yak_b200", "bin", "yak-b200")
    fa, fb, fc = _yak_files()
    reads = G.input_path("reads_q")
    for args in (["subtract", fa, fb], ["isec", fa, fb, fc], ["recount", fc, reads]):
        o1, o2 = os.path.join(util.TMP, "yakb_cli_a.yak"), os.path.join(util.TMP, "yakb_cli_b.yak")
        subprocess.run([O.REF_YAK, args[0], "-o", o1] + args[1:], check=True, capture_output=True)
        subprocess.run([exe, args[0], "-o", o2] + args[1:], check=True, capture_output=True)
        a, b = open(o1, "rb").read(), open(o2, "rb").read()
        assert a == b, args[0] + ": " + util.explain_diff(b, a)
    for flags in ([], ["-c"]):
        r1 = subprocess.run([O.REF_YAK, "print"] + flags + [fc], check=True, capture_output=True).stdout
        r2 = subprocess.run([exe, "print"] + flags + [fc], check=True, capture_output=True).stdout
        assert r1 == r2 and len(r1) > 1000


@pytest.mark.skipif(not os.path.exists(O.REF_LIB), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", range(12))
def test_oracle_random_scripts_equal_reference_functions(seed):
    """random sequences of table operations - tighten / setcnt / shrink / merge / subtract / isec and yak_ch_restore_core
    into the table in every mode - on the reference's own functions and on the oracle: same .yak bytes after every step"""
    import numpy as np
    from yak_b200.capi import YakCh
    rng = np.random.default_rng(300 + seed)
    R = C.CDLL(O.REF_LIB)
    ChP = C.POINTER(YakCh)
    R.yak_ch_restore.restype = ChP; R.yak_ch_restore.argtypes = [C.c_char_p]
    R.yak_ch_restore_core.restype = ChP
    R.yak_ch_dump.argtypes = [ChP, C.c_char_p]
    R.yak_ch_destroy.argtypes = [ChP]
    R.yak_ch_tighten.argtypes = [ChP]
    R.yak_ch_setcnt.argtypes = [ChP, C.c_int, C.c_int]
    R.yak_ch_shrink.argtypes = [ChP, C.c_int, C.c_int, C.c_int]
    R.yak_ch_merge.argtypes = [ChP, ChP, C.c_int, C.c_int, C.c_int, C.c_int]
    R.yak_ch_subtract.argtypes = [ChP, ChP, C.c_int]
    R.yak_ch_isec.argtypes = [ChP, ChP, C.c_int]
    L = O.lib()
    L.yo_ch_restore_core.restype = C.POINTER(O.YoCh)
    L.yo_ch_restore_core.argtypes = [C.POINTER(O.YoCh), C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]
    files = dict(zip("ABC", _yak_files()))
    tmp = os.path.join(util.TMP, "yakb_setops_rnd.yak")
    start = str(rng.choice(list("ABC")))
    hr, ho = R.yak_ch_restore(files[start].encode()), L.yo_ch_restore(files[start].encode())
    for step in range(int(rng.integers(2, 7))):
        op = str(rng.choice(["tighten", "setcnt", "shrink", "merge", "subtract", "isec", "load"]))
        x = files[str(rng.choice(list("ABC")))].encode()
        if op == "tighten":
            R.yak_ch_tighten(hr); L.yo_ch_tighten(ho)
        elif op == "setcnt":
            c = int(rng.integers(0, 1024))
            R.yak_ch_setcnt(hr, c, 2); L.yo_ch_setcnt(ho, c)
        elif op == "shrink":
            lo, hi = int(rng.integers(0, 6)), int(rng.choice([3, 20, 1023]))
            R.yak_ch_shrink(hr, lo, hi, 2); L.yo_ch_shrink(ho, lo, hi)
        elif op == "merge":
            lo, hi, pr = int(rng.integers(0, 4)), int(rng.choice([5, 1023])), int(rng.integers(0, 2))
            R.yak_ch_merge(hr, R.yak_ch_restore(x), lo, hi, 2, pr); L.yo_ch_merge(ho, L.yo_ch_restore(x), lo, hi, pr)
        elif op in ("subtract", "isec"):
            a, b = R.yak_ch_restore(x), L.yo_ch_restore(x)
            (R.yak_ch_subtract if op == "subtract" else R.yak_ch_isec)(hr, a, 2)
            (L.yo_ch_subtract if op == "subtract" else L.yo_ch_isec)(ho, b)
            R.yak_ch_destroy(a); L.yo_ch_destroy(b)
        else:   # yak_ch_restore_core(ch0 != NULL, mode): htab.c:436-472
            mode = int(rng.integers(1, 7))
            mn, md = int(rng.integers(1, 4)), int(rng.integers(3, 9))
            R.yak_ch_restore_core.argtypes = [ChP, C.c_char_p, C.c_int, C.c_int, C.c_int]
            assert R.yak_ch_restore_core(hr, x, mode, mn, md)
            assert L.yo_ch_restore_core(ho, x, mode, mn, md, None)
        R.yak_ch_dump(hr, tmp.encode())
        got, want = O.dump_bytes(ho), open(tmp, "rb").read()
        assert got == want, f"seed {seed} step {step} {op}: " + util.explain_diff(got, want)
    R.yak_ch_destroy(hr); L.yo_ch_destroy(ho)
