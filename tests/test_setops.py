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


CNTASM_CASES = [
    # (options, number of inputs, start from the table of an earlier run via -i)
    ([], 3, False),
    (["-c1", "-x3", "-e1", "-s2"], 4, False),
    (["-c1", "-x2", "-r", "-e0", "-s1", "-k27", "-p11"], 3, False),
    (["-c2", "-x1023", "-e2", "-s1"], 4, True),
]


def _cntasm_inputs(n):
    """n small 'assemblies' sharing most of their sequence: windows of one seeded genome, a few bases changed in each"""
    import numpy as np
    from yak_b200 import synth
    g = synth.genome_codes(4242, 60_000)
    out = []
    for i in range(n):
        p = os.path.join(util.TMP, f"yakb_cntasm_{i}.fa")
        rng = np.random.default_rng(900 + i)
        s = g[i * 2000: 40_000 + i * 5000].copy()
        pos = rng.integers(0, len(s), 150)
        s[pos] = (s[pos] + rng.integers(1, 4, 150)) & 3
        seq = np.frombuffer(b"ACGT", dtype=np.uint8)[s]
        with open(p, "wb") as f:
            for j, a in enumerate(range(0, len(seq), 9000)):   # a few contigs, a duplicated one so that counts above 1 occur
                f.write(b">c%d\n" % j + seq[a:a + 9000].tobytes() + b"\n")
            f.write(b">dup\n" + seq[100:3100].tobytes() + b"\n")
        out.append(p)
    return out


def _cntasm_oracle(opts, fns, fn_in):
    """main.c:90-162 on the oracle's functions"""
    import getopt
    L = O.lib()
    o = dict(getopt.getopt(opts, "k:p:c:x:e:s:r")[0])
    k, pre = int(o.get("-k", 31)), int(o.get("-p", 10))
    mn, mx, max_out, check_n, pr = int(o.get("-c", 1)), int(o.get("-x", 1)), int(o.get("-e", 0)), int(o.get("-s", 10)), int("-r" in o)
    h = L.yo_ch_restore(fn_in.encode()) if fn_in else None
    for n, fn in enumerate(fns, 1):
        h1, _ = O.count_file(fn, k=k, pre=pre, bf_shift=0)
        if not h:
            h = h1
            L.yo_ch_shrink(h, mn, mx); L.yo_ch_setcnt(h, 1)
        else:
            L.yo_ch_merge(h, h1, mn, mx, pr)
        if n == len(fns) or (n > max_out and n % check_n == 0):
            L.yo_ch_shrink(h, n - max_out, 1023)
    L.yo_ch_tighten(h)
    b = O.dump_bytes(h)
    L.yo_ch_destroy(h)
    return b


def _cntasm_run(exe, opts, fns, fn_in, out):
    import subprocess
    cmd = [exe, "cntasm", "-o", out] + (opts if any(x.startswith("-p") for x in opts) else opts + ["-p10"]) + (["-i", fn_in] if fn_in else []) + fns
    subprocess.run(cmd, check=True, capture_output=True)
    return open(out, "rb").read()


@pytest.mark.skipif(not os.path.exists(O.REF_YAK), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", CNTASM_CASES, ids=lambda c: "cntasm" + "".join(c[0]))
def test_oracle_cntasm_flow_equals_reference_binary(case):
    """`yak cntasm` of the reference binary against the same sequence of operations on the oracle"""
    opts, n, chain = case
    fns = _cntasm_inputs(n)
    prev = os.path.join(util.TMP, "yakb_cntasm_prev.yak")
    if chain:
        _cntasm_run(O.REF_YAK, ["-c1", "-x1"], fns[:2], None, prev)
    want = _cntasm_run(O.REF_YAK, opts, fns, prev if chain else None, os.path.join(util.TMP, "yakb_cntasm_ref.yak"))
    got = _cntasm_oracle(opts, fns, prev if chain else None)
    assert len(want) > 16 + 8 * 1024 + 8 * 1000          # the case keeps k-mers
    assert got == want, util.explain_diff(got, want)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(O.REF_YAK), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", CNTASM_CASES, ids=lambda c: "cntasm" + "".join(c[0]))
def test_cli_cntasm_equals_reference_binary(case):
    opts, n, chain = case
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "yak_b200", "bin", "yak-b200")
    fns = _cntasm_inputs(n)
    prev = os.path.join(util.TMP, "yakb_cntasm_prev.yak")
    if chain:
        _cntasm_run(O.REF_YAK, ["-c1", "-x1"], fns[:2], None, prev)
    want = _cntasm_run(O.REF_YAK, opts, fns, prev if chain else None, os.path.join(util.TMP, "yakb_cntasm_ref.yak"))
    got = _cntasm_run(exe, opts, fns, prev if chain else None, os.path.join(util.TMP, "yakb_cntasm_cli.yak"))
    assert got == want, util.explain_diff(got, want)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(O.REF_YAK), reason="oracle/_ref not built")
def test_cli_inspect_two_files_equals_reference_binary():
    """`inspect [-m] in1.yak in2.yak`: the lookups of stored keys (quirk Q7) as batched device lookups, stdout of the reference"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "yak_b200", "bin", "yak-b200")
    fa, fb, fc = _yak_files()
    for args in ([fa, fb], ["-m7", fb, fa], ["-m40", fc, fa], [fa, fa]):
        r = subprocess.run([O.REF_YAK, "inspect"] + args, check=True, capture_output=True).stdout
        m = subprocess.run([exe, "inspect"] + args, check=True, capture_output=True).stdout
        assert m == r and r.count(b"\n") > 3, args
