"""CPU checks of the scanner path against the UNMODIFIED reference's stdout
(tests/golden/scan_*.txt): the oracle's restatements of yak_ch_restore_core's flag modes and of the shared lookup
loop feed yak_b200/cli/scan_logic.c - the per-sequence logic the CLI runs behind the batched device lookup - and
the text must equal the reference's byte for byte.  The GPU side of the same cases is tests/test_gpu_scan.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
import scan_inputs as S
import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CHUNK = {"triobin": 200_000_000, "trioeval": 1_000_000_000, "chkerr": 1_000_000_000, "sexchr": 1_000_000_000}


class Batch(C.Structure):
    _fields_ = [("n_seq", C.c_int64), ("names", C.POINTER(C.c_char_p)), ("lens", C.POINTER(C.c_int64)), ("vals", C.POINTER(C.c_int16))]


class TbOpt(C.Structure):
    _fields_ = [("k", C.c_int), ("print_diff", C.c_int), ("ratio_thres", C.c_double)]


class TeOpt(C.Structure):
    _fields_ = [("k", C.c_int), ("min_n", C.c_int), ("print_err", C.c_int), ("print_frag", C.c_int)]


class TeSum(C.Structure):
    _fields_ = [("v", C.c_int64 * 6)]


class CeOpt(C.Structure):
    _fields_ = [("k", C.c_int), ("min_cnt", C.c_int), ("min_streak", C.c_int)]


@pytest.fixture(scope="module")
def env():
    oracle_lib.build()
    so = os.path.join(util.TMP, "yakb_scan_logic_test.so")
    subprocess.run(["gcc", "-O2", "-Wall", "-fPIC", "-shared", "-o", so, os.path.join(ROOT, "yak_b200", "cli", "scan_logic.c")], check=True)
    logic = C.CDLL(so)
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p; libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    for f in ("yakb_triobin_batch", "yakb_chkerr_batch"):
        getattr(logic, f).argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Batch)]
    logic.yakb_trioeval_batch.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Batch), C.POINTER(TeSum)]
    logic.yakb_trioeval_header.argtypes = [C.c_void_p]
    logic.yakb_trioeval_footer.argtypes = [C.c_void_p, C.POINTER(TeSum)]
    logic.yakb_sexchr_header.argtypes = [C.c_void_p]
    logic.yakb_sexchr_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(Batch)]
    O = oracle_lib.lib()
    O.yo_ch_restore_core.restype = C.POINTER(oracle_lib.YoCh)
    O.yo_ch_restore_core.argtypes = [C.POINTER(oracle_lib.YoCh), C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]
    O.yo_scan_seq.argtypes = [C.POINTER(oracle_lib.YoCh), C.c_int64, C.c_char_p, C.POINTER(C.c_int16)]
    O.yo_reader_open.restype = C.c_void_p; O.yo_reader_open.argtypes = [C.c_char_p]
    O.yo_reader_next.restype = C.c_int64; O.yo_reader_next.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]
    O.yo_reader_close.argtypes = [C.c_void_p]
    paths = S.write_all(util.TMP)
    for y, (fa, k) in S.COUNTS.items():   # the oracle's count is byte-exact against the reference (test_oracle_cpu.py)
        paths[y] = os.path.join(util.TMP, "yakb_scan_" + y)
        h, _ = oracle_lib.count_file(paths[fa], k=k, pre=10, bf_shift=0)
        assert O.yo_ch_dump(h, paths[y].encode()) == 0
        O.yo_ch_destroy(h)
    return logic, libc, O, paths


def batches(O, fn, chunk, workers=2):
    """bseq.c:33-57 under kt_pipeline(2, ...): records until the batch holds >= chunk bases or kseq_read fails; a truncated
    FASTQ record (-2) only ends the batch, and a batch that read nothing retires one pipeline worker (kthread.c:119)"""
    r = O.yo_reader_open(fn.encode())
    assert r
    names, seqs, size = [], [], 0
    seq, name = C.c_char_p(), C.c_char_p()
    while True:
        ln = O.yo_reader_next(r, C.byref(seq), C.byref(name))
        if ln >= 0:
            names.append(name.value)
            seqs.append(C.string_at(seq, ln))
            size += ln
        if ln < 0 or size >= chunk:
            if names:
                yield names, seqs
            elif ln == -2:
                workers -= 1
            names, seqs, size = [], [], 0
            if ln == -1 or workers == 0:
                break
    O.yo_reader_close(r)


def make_batch(O, ch, names, seqs):
    lens = (C.c_int64 * len(seqs))(*[len(s) for s in seqs])
    tot = sum(len(s) for s in seqs)
    vals = (C.c_int16 * max(tot, 1))()
    off = 0
    for s in seqs:
        O.yo_scan_seq(ch, len(s), s, C.cast(C.byref(vals, off * 2), C.POINTER(C.c_int16)))
        off += len(s)
    nm = (C.c_char_p * len(names))(*names)
    b = Batch(len(seqs), nm, lens, vals)
    b._keep = (nm, lens, vals)
    return b


def getopt(args, spec):
    """tiny getopt for the -xVALUE / -x forms the cases use; returns (dict, positional)"""
    o, pos = {}, []
    for a in args:
        if a.startswith("-") and len(a) > 1:
            o[a[1]] = a[2:] if spec[a[1]] else True
        else:
            pos.append(a)
    return o, pos


def parse_num(s):
    mul = {"k": 1e3, "m": 1e6, "g": 1e9}.get(s[-1].lower())
    return int(float(s[:-1]) * mul + .499) if mul else int(float(s) + .499)


def run_case(env, cmd):
    logic, libc, O, paths = env
    name, args = cmd[0], S.argv(cmd[1:], paths)
    out_path = os.path.join(util.TMP, "yakb_scan_out.txt")
    fp = libc.fopen(out_path.encode(), b"w")
    if name in ("triobin", "trioeval"):
        o, pos = getopt(args, {"c": 1, "d": 1, "t": 1, "p": 0, "r": 1, "n": 1, "e": 0, "F": 0})
        mn, md = int(o.get("c", 2)), int(o.get("d", 5))
        ch = O.yo_ch_restore_core(None, pos[0].encode(), 2, mn, md, None)
        assert ch and O.yo_ch_restore_core(ch, pos[1].encode(), 3, mn, md, None)
        if name == "triobin":
            opt = TbOpt(ch.contents.k, 1 if "p" in o else 0, float(o.get("r", 0.33)))
            for names, seqs in batches(O, pos[2], CHUNK[name]):
                logic.yakb_triobin_batch(fp, C.byref(opt), C.byref(make_batch(O, ch, names, seqs)))
        else:
            opt = TeOpt(ch.contents.k, int(o.get("n", 2)), 1 if "e" in o else 0, 0 if "F" in o else 1)
            sm = TeSum()
            logic.yakb_trioeval_header(fp)
            for names, seqs in batches(O, pos[2], CHUNK[name]):
                logic.yakb_trioeval_batch(fp, C.byref(opt), C.byref(make_batch(O, ch, names, seqs)), C.byref(sm))
            logic.yakb_trioeval_footer(fp, C.byref(sm))
    elif name == "chkerr":
        o, pos = getopt(args, {"t": 1, "c": 1, "s": 1})
        ch = O.yo_ch_restore_core(None, pos[0].encode(), 1, 0, 0, None)
        opt = CeOpt(ch.contents.k, int(o.get("c", 3)), int(o.get("s", 5)))
        for names, seqs in batches(O, pos[1], CHUNK[name]):
            logic.yakb_chkerr_batch(fp, C.byref(opt), C.byref(make_batch(O, ch, names, seqs)))
    else:
        o, pos = getopt(args, {"t": 1, "K": 1})
        ch = O.yo_ch_restore_core(None, pos[0].encode(), 4, 0, 0, None)
        assert ch and O.yo_ch_restore_core(ch, pos[1].encode(), 5, 0, 0, None) and O.yo_ch_restore_core(ch, pos[2].encode(), 6, 0, 0, None)
        logic.yakb_sexchr_header(fp)
        for hap in (1, 2):
            for names, seqs in batches(O, pos[2 + hap], parse_num(o["K"]) if "K" in o else CHUNK[name]):
                logic.yakb_sexchr_batch(fp, hap, C.byref(make_batch(O, ch, names, seqs)))
    libc.fclose(fp)
    O.yo_ch_destroy(ch)
    return open(out_path, "rb").read()


@pytest.mark.parametrize("gold,cmd", S.CASES, ids=[c[0] for c in S.CASES])
def test_scan_logic_on_oracle_lookups_equals_reference_stdout(env, gold, cmd):
    got = run_case(env, cmd)
    want = open(os.path.join(GOLD, gold), "rb").read()
    assert got == want


def test_restore_core_modes_of_the_oracle(env):
    """flag bits after TRIOBIN1+2 and SEXCHR1+2+3 loads (htab.c:448-470) and the mode errors (htab.c:412-419)"""
    _, _, O, paths = env
    n_io = (C.c_int64 * 2)()
    assert not O.yo_ch_restore_core(None, paths["pat.yak"].encode(), 3, 2, 5, None)   # TRIOBIN2 needs a table
    assert not O.yo_ch_restore_core(None, paths["pat.yak"].encode(), 5, 0, 0, None)
    assert not O.yo_ch_restore_core(None, paths["pat.yak"].encode(), 7, 0, 0, None)
    ch = O.yo_ch_restore_core(None, paths["pat.yak"].encode(), 2, 2, 5, n_io)
    first = n_io[1]
    assert first == n_io[0] > 0
    assert O.yo_ch_restore_core(ch, paths["mat.yak"].encode(), 3, 2, 5, n_io)
    assert 0 < n_io[1] < n_io[0]                       # most maternal k-mers are shared with the father
    hist = (C.c_int64 * 1024)()
    O.yo_ch_hist(ch, hist)
    assert sum(hist) == first + n_io[1] and all(hist[i] == 0 for i in range(16, 1024)) and hist[0] == 0
    assert hist[2] > 0 and hist[8] > 0 and hist[10] > hist[2]   # specific to one parent / class 2 in both
    O.yo_ch_destroy(ch)


def _random_contigs(rng, a, b, c, path):
    """contigs cut from the three genomes, mixed within a contig, with Ns / lower case / short ones, FASTA or FASTQ"""
    from yak_b200 import synth
    fastq = rng.random() < 0.3
    out = bytearray()
    for i in range(int(rng.integers(2, 9))):
        segs = []
        for _ in range(int(rng.integers(1, 5))):
            g = (a, b, c)[int(rng.integers(0, 3))]
            st, ln = int(rng.integers(0, S.G - 4000)), int(rng.integers(10, 4000))
            segs.append(g[st:st + ln])
        x = np.concatenate(segs).copy()
        m = rng.random(x.size) < 0.002
        x[m] = (x[m] + 1) & 3
        asc = bytearray(synth.codes_to_ascii(x).tobytes())
        for _ in range(int(rng.integers(0, 3))):
            p = int(rng.integers(0, len(asc)))
            asc[p:p + int(rng.integers(1, 30))] = b"N" * min(int(rng.integers(1, 30)), len(asc) - p)
        if rng.random() < 0.3:
            asc = asc.lower()
        if fastq:
            out += b"@c%d x\n" % i + bytes(asc) + b"\n+\n" + b"I" * len(asc) + b"\n"
        else:
            w = int(rng.choice([0, 60, 1000]))
            body = bytes(asc) if not w else b"\n".join(bytes(asc[q:q + w]) for q in range(0, len(asc), w))
            out += b">c%d x\n" % i + body + b"\n"
    open(path, "wb").write(bytes(out))


@pytest.mark.parametrize("seed", range(16))
def test_scan_logic_randomised_against_reference_binary(env, seed):
    """random contigs and random options: the reference binary's stdout (-t1) against oracle lookups + cli/scan_logic.c"""
    if not os.path.exists(oracle_lib.REF_YAK):
        pytest.skip("oracle/_ref not built")
    logic, libc, O, paths = env
    rng = np.random.default_rng(500 + seed)
    a, b, c = S.genomes()
    paths = dict(paths)
    for tag in ("rc1.fx", "rc2.fx"):
        paths[tag] = os.path.join(util.TMP, f"yakb_scan_rnd{seed}_{tag}")
        _random_contigs(rng, a, b, c, paths[tag])
    k47 = rng.random() < 0.25
    pat, mat = ("{pat47.yak}", "{mat47.yak}") if k47 else ("{pat.yak}", "{mat.yak}")
    which = seed % 4
    if which == 0:
        cmd = ["triobin", "-t1", f"-c{int(rng.integers(1, 5))}", f"-d{int(rng.integers(2, 12))}", f"-r{float(rng.choice([0.1, 0.33, 0.9]))}"]
        cmd += (["-p"] if rng.random() < 0.5 else []) + [pat, mat, "{rc1.fx}"]
    elif which == 1:
        cmd = ["trioeval", "-t1", f"-c{int(rng.integers(1, 5))}", f"-d{int(rng.integers(2, 12))}", f"-n{int(rng.integers(1, 6))}"]
        cmd += (["-e"] if rng.random() < 0.5 else []) + (["-F"] if rng.random() < 0.3 else []) + [pat, mat, "{rc1.fx}"]
    elif which == 2:
        cmd = ["chkerr", "-t1", f"-c{int(rng.integers(0, 9))}", f"-s{int(rng.integers(0, 9))}", mat if rng.random() < 0.5 else pat, "{rc1.fx}"]
    else:
        cmd = ["sexchr", "-t1"] + ([f"-K{int(rng.choice([500, 3000, 100000]))}"] if rng.random() < 0.6 else [])
        cmd += ["{pat.yak}", "{third.yak}", "{mat.yak}", "{rc1.fx}", "{rc2.fx}"]
    ref = subprocess.run([oracle_lib.REF_YAK] + S.argv(cmd, paths), capture_output=True)
    assert ref.returncode == 0, ref.stderr.decode()[-500:]
    env2 = (logic, libc, O, paths)
    got = run_case(env2, cmd)
    assert got == ref.stdout, (cmd, len(got), len(ref.stdout))


@pytest.mark.parametrize("seed", range(10))
def test_qv_oracle_randomised_against_reference_binary(env, seed):
    """`yak qv` end to end on the CPU: the reference binary's stdout (-t1, random -l / -f / -e / -K / -p, contigs with
    truncated FASTQ records among them) against oracle reader + bseq flow + yo_qv_seqs + cli/qv_solve.c"""
    import math
    if not os.path.exists(oracle_lib.REF_YAK):
        pytest.skip("oracle/_ref not built")
    logic, libc, O, paths = env
    rng = np.random.default_rng(900 + seed)
    a, b, c = S.genomes()
    fx = os.path.join(util.TMP, f"yakb_qv_rnd{seed}.fx")
    _random_contigs(rng, a, b, c, fx)
    if open(fx, "rb").read(1) == b"@" and seed % 2:          # FASTQ: add records with a quality one character too long
        recs = open(fx, "rb").read().split(b"\n@")
        bad = b"bad\nACGTACGTTTGACCA\n+\nIIIIIIIIIIIIIIII"
        for _ in range(int(rng.integers(1, 4))):
            recs.insert(int(rng.integers(0, len(recs) + 1)), bad)
        open(fx, "wb").write(b"\n@".join(recs) if not recs[0].startswith(b"bad") else b"@" + b"\n@".join(recs))
    min_len = int(rng.choice([0, 50, 2000])); min_frac = float(rng.choice([0.5, 0.9, 0.0])); fpr = float(rng.choice([4e-5, 1e-3]))
    chunk = int(rng.choice([1_000_000_000, 3000])); each = bool(rng.random() < 0.6)
    y = paths["pat.yak"]
    cmd = [oracle_lib.REF_YAK, "qv", "-t1", f"-l{min_len}", f"-f{min_frac}", f"-e{fpr}", f"-K{chunk}"] + (["-p"] if each else []) + [y, fx]
    ref = subprocess.run(cmd, capture_output=True)
    assert ref.returncode == 0, ref.stderr.decode()[-400:]
    # the same on our side of the fence
    O.yo_ch_restore.restype = C.POINTER(oracle_lib.YoCh)
    ch = O.yo_ch_restore(y.encode())
    k = ch.contents.k
    hist = (C.c_int64 * 1024)()
    O.yo_ch_hist(ch, hist)
    names, seqs = [], []
    for nm, ss in batches(O, fx, chunk):
        names += nm; seqs += ss
    lens = (C.c_int64 * max(len(seqs), 1))(*[len(s) for s in seqs])
    cnt = (C.c_int64 * 1024)()
    tot, non0 = (C.c_int32 * max(len(seqs), 1))(), (C.c_int32 * max(len(seqs), 1))()
    O.yo_qv_seqs.argtypes = [C.POINTER(oracle_lib.YoCh), C.c_int64, C.POINTER(C.c_int64), C.c_char_p, C.c_int, C.c_double,
                             C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    O.yo_qv_seqs(ch, len(seqs), lens, b"".join(seqs), min_len, min_frac, cnt, tot, non0)
    out = ["CC\tCT  kmer_occurrence    short_read_kmer_count  raw_input_kmer_count  adjusted_input_kmer_count", "CC\tFR  fpr_lower_bound    fpr_upper_bound",
           "CC\tER  total_input_kmers  adjusted_error_kmers", "CC\tCV  coverage", "CC\tQV  raw_quality_value  adjusted_quality_value", "CC"]
    if each:
        for i, s in enumerate(seqs):
            if len(s) < min_len:
                continue
            qv = -1.0
            if tot[i] > 0:
                qv = 0.0 if non0[i] == 0 else 99.0 if tot[i] == non0[i] else -4.3429448190325175 * math.log(math.log(tot[i] / non0[i]) / k)
            out.append("SQ\t%s\t%d\t%d\t%d\t%.2f" % (names[i].decode(), len(s), tot[i], non0[i], qv))
    so = os.path.join(util.TMP, "yakb_qvsolve3.so")
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-o", so, os.path.join(ROOT, "yak_b200", "cli", "qv_solve.c"), "-lm"], check=True)

    class Qs(C.Structure):
        _fields_ = [("tot", C.c_int64), ("qv_raw", C.c_double), ("qv", C.c_double), ("cov", C.c_double), ("err", C.c_double),
                    ("fpr_lower", C.c_double), ("fpr_upper", C.c_double), ("adj_cnt", C.c_double * 1024)]
    Sv = C.CDLL(so)
    Sv.yak_qv_solve.argtypes = [C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int, C.c_double, C.POINTER(Qs)]
    qs = Qs()
    Sv.yak_qv_solve(hist, cnt, k, fpr, C.byref(qs))
    def f3(x):  # C's printf writes the sign of a NaN, Python's % does not
        return ("-nan" if math.copysign(1.0, x) < 0 else "nan") if math.isnan(x) else "%.3f" % x
    out += ["CT\t%d\t%d\t%d\t%s" % (i, hist[i], cnt[i], f3(qs.adj_cnt[i])) for i in range(1023, -1, -1)]
    out += ["FR\t%.3g\t%.3g" % (qs.fpr_lower, qs.fpr_upper), "ER\t%d\t%s" % (qs.tot, f3(qs.err)), "CV\t%s" % f3(qs.cov), "QV\t%s\t%s" % (f3(qs.qv_raw), f3(qs.qv))]
    O.yo_ch_destroy(ch)
    want = ref.stdout.decode().split("\n")
    got = out + [""]
    if max(cnt[2:1023]) == 0:   # the reference reads cnt[-1] / hist[-1] for the coverage then (undefined): leave the CV line out
        got, want = [ln for ln in got if not ln.startswith("CV")], [ln for ln in want if not ln.startswith("CV")]
    assert got == want, (cmd, next((i, g, w) for i, (g, w) in enumerate(zip(got, want)) if g != w) if len(got) == len(want) else (len(got), len(want)))
