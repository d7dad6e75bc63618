"""CPU tests: the oracle against the reference's golden vectors; host logic; the C ABI surface."""
import ctypes as C
import hashlib
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import golden_util as G
import oracle_lib as O
import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------- oracle pinned to the reference

def test_inputs_are_reproducible():
    for name, meta in G.GOLD["inputs"].items():
        data = open(G.input_path(name), "rb").read()
        assert hashlib.sha256(data).hexdigest() == meta["sha256"], name


def test_primitive_kats():
    L = O.lib()
    kat = G.GOLD["kat"]
    for x, m, want in kat["hash64"]:
        assert L.yo_hash64(x, m) == want
        assert L.yo_hash64_inv(want, m) == x
    for x, want in kat["hash64_64"]:
        assert L.yo_hash64_64(x) == want
    for q, want in kat["hash_long"]:
        assert L.yo_hash_long((C.c_uint64 * 4)(*q)) == want
    for h, bits, want in kat["h2b"]:
        assert L.yo_slot_home(h << 10, bits) == want
    # SURVEY section 4 table (captured from the reference in the survey session)
    m62 = (1 << 62) - 1
    assert [L.yo_hash64(x, m62) for x in (0, 1, 0x123456789abcdef, m62)] == \
        [0x1df3e87bbc06f2a4, 0x1bca7c69b794f8ce, 0x2437e41bd0ec327b, 0x37ba6eccef93ff51]
    assert L.yo_hash64(0x2a, (1 << 42) - 1) == 0x2f1e7b6f7a
    assert [L.yo_hash64_64(x) for x in (0, 1)] == [0x77cfa1eef01bca90, 0x5bca7c69b794f8ce]
    assert L.yo_hash_long((C.c_uint64 * 4)(5, 9, 3, 7)) == 0x95ea2abb2bd45540
    assert [L.yo_slot_home(h << 10, b) for h, b in ((1, 2), (0xdeadbeef, 20), (12345, 10))] == [2, 598639, 644]
    b = L.yo_bloom_init(25, 4)
    assert (L.yo_bloom_insert(b, 0x1234567) , L.yo_bloom_insert(b, 0x1234567)) == (0, 4)
    L.yo_bloom_destroy(b)
    assert not L.yo_bloom_init(8, 4) and not L.yo_bloom_init(56, 4)          # bbf.c:9


def test_khashl_quirks():
    L = O.lib()
    pre = 10
    keys = np.array([(i * 7919 + 3) << pre | 5 for i in range(1, 4)], dtype=np.uint64)
    import struct
    for extra, want in ((False, (4, 3)), (True, (8, 3))):                    # Q3 trailing-put doubling
        h = L.yo_ch_init(31, pre, 4, 0)
        a, p = util.u64_array(np.concatenate([keys, keys[:1]]) if extra else keys)
        L.yo_ch_insert_list(h, 1, len(a), p)
        data = O.dump_bytes(h)
        off = 16
        for _ in range(5):
            cap, size = struct.unpack_from("<II", data, off); off += 8 + 8 * size
        assert struct.unpack_from("<II", data, off) == want
        assert data[:16] == bytes.fromhex("59414b02" "1f000000" "0a000000" "0a000000")
        L.yo_ch_shrink(h, 1, 1023)                                            # Q9
        assert struct.unpack_from("<II", O.dump_bytes(h), 16) == (4, 0)
        L.yo_ch_destroy(h)
    assert not L.yo_ch_init(31, 9, 4, 0)


@pytest.mark.parametrize("case", G.GOLD["cases"], ids=G.case_id)
def test_oracle_matches_reference_golden(case):
    if case["input"] == "cfg1" and case["bf_shift"] == 0 and os.environ.get("YAKB_FAST_TESTS"):
        pytest.skip("fast mode")
    fn = G.input_path(case["input"])
    fn2 = G.input_path(case["second"]) if case["second"] else None
    h, ne = O.count_file(fn, k=case["k"], pre=case["pre"], bf_shift=case["bf_shift"], fn2=fn2)
    data = O.dump_bytes(h)
    O.lib().yo_ch_destroy(h)
    assert len(data) == case["bytes"]
    assert hashlib.sha256(data).hexdigest() == case["sha256"]


def test_oracle_restore_and_qv_against_reference_output():
    """yo_ch_restore + yo_qv_seqs reproduce the SQ lines and CT histogram the reference printed."""
    L = O.lib()
    y = os.path.join(G.HERE, "reads_c_k31_p10_b22.yak")
    h = L.yo_ch_restore(y.encode())
    assert h
    from yak_b200 import synth
    ctg = synth.contigs_bytes(7, 100_000, 3, 8, 20_000, sub=2e-3)
    seqs = [ln for ln in ctg.split(b"\n") if ln and not ln.startswith(b">")]
    lens = (C.c_int64 * len(seqs))(*[len(s) for s in seqs])
    cnt = (C.c_int64 * 1024)()
    tot, non0 = (C.c_int32 * len(seqs))(), (C.c_int32 * len(seqs))()
    L.yo_qv_seqs(h, len(seqs), lens, b"".join(seqs), 0, 0.5, cnt, tot, non0)
    hist = (C.c_int64 * 1024)()
    L.yo_ch_hist(h, hist)
    ref = open(os.path.join(G.HERE, "qv_reads_c_ctg.txt")).read().splitlines()
    sq = [ln.split("\t") for ln in ref if ln.startswith("SQ")]
    assert [(int(x[3]), int(x[4])) for x in sq] == list(zip(tot, non0))
    ct = {int(x[1]): (int(x[2]), int(x[3])) for x in (ln.split("\t") for ln in ref if ln.startswith("CT"))}
    assert all(ct[i] == (hist[i], cnt[i]) for i in range(1024))
    L.yo_ch_destroy(h)


@pytest.mark.skipif(not os.path.exists(O.REF_YAK), reason="oracle/_ref not built (no /root/reference here)")
def test_oracle_vs_live_reference_edge_cases():
    import test_gpu_parity as T
    fn = T._edge_file(os.path.join(util.TMP, "yakb_edge_cpu.fa"))
    for k, pre, b in ((31, 10, 0), (31, 10, 19), (33, 10, 0), (5, 10, 0)):
        y = os.path.join(util.TMP, "yakb_edge_ref.yak")
        O.ref_count(fn, y, k=k, pre=pre, bf_shift=b)
        h, _ = O.count_file(fn, k=k, pre=pre, bf_shift=b)
        assert O.dump_bytes(h) == open(y, "rb").read()
        O.lib().yo_ch_destroy(h)


# ---------------------------------------------------------------- host logic of the product (no GPU)

def test_library_loads_and_exports_every_declared_symbol():
    from yak_b200 import capi
    L = capi.lib()
    names = set()
    for hdr in ("yak.h", "yak_b200.h"):
        txt = open(os.path.join(ROOT, "include", hdr)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names |= set(re.findall(r"\b(yakb?_[a-z0-9_]+)\s*\(", txt))
    assert len(names) > 40
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing
    for var in ("yak_verbose", "seq_nt4_table"):
        assert C.c_int.in_dll(L, var) is not None
    assert L.yakb_version().startswith(b"0.1")


def test_nt4_table_and_option_defaults():
    from yak_b200 import capi
    L = capi.lib()
    tab = (C.c_ubyte * 256).in_dll(L, "seq_nt4_table")
    want = (C.c_ubyte * 256).in_dll(O.lib(), "yo_nt4")
    assert bytes(tab) == bytes(want)
    o = capi.YakCopt()
    L.yak_copt_init(C.byref(o))
    assert (o.k, o.pre, o.bf_shift, o.bf_n_hash, o.n_thread, o.chunk_size) == (31, 10, 0, 4, 4, 10_000_000)   # misc.c:23-32
    q = capi.YakQopt()
    L.yak_qopt_init(C.byref(q))
    assert (q.chunk_size, q.n_threads, q.min_frac, q.fpr, q.min_len) == (1_000_000_000, 4, 0.5, 0.00004, 0)  # qv.c:137-144


def test_no_gpu_means_loud_failure_not_fallback():
    from yak_b200 import capi
    L = capi.lib()
    if L.yakb_device_count() > 0:
        pytest.skip("a GPU is present")
    assert not L.yak_ch_init(31, 12, 4, 0)
    o = capi.copt(31, 12)
    assert not L.yak_count(G.input_path("reads_a").encode(), C.byref(o), None)
    with pytest.raises(RuntimeError):
        capi.require_gpu()


def _records_product(path):
    from yak_b200 import capi
    L = capi.lib()
    r = L.yakb_fastx_open(path.encode())
    assert r
    out = []
    seq, name = C.c_char_p(), C.c_char_p()
    while True:
        n = L.yakb_fastx_next(r, C.byref(seq), C.byref(name))
        if n < 0:
            out.append(n)
            break
        out.append((name.value, seq.value))
        assert len(seq.value) == n
    L.yakb_fastx_close(r)
    return out


def _records_oracle(path):
    L = O.lib()
    L.yo_reader_open.restype = C.c_void_p; L.yo_reader_open.argtypes = [C.c_char_p]
    L.yo_reader_next.restype = C.c_int64; L.yo_reader_next.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]
    L.yo_reader_close.argtypes = [C.c_void_p]
    r = L.yo_reader_open(path.encode())
    out = []
    seq, name = C.c_char_p(), C.c_char_p()
    while True:
        n = L.yo_reader_next(r, C.byref(seq), C.byref(name))
        if n < 0:
            out.append(n)
            break
        out.append((name.value, seq.value))
    L.yo_reader_close(r)
    return out


def _ref_flow(path, min_len, chunk=10_000_000, workers=3):
    """the sequences `yak count` processes, in order: count.c:88-110 over kseq_read (the oracle's restatement of it) -
    a truncated FASTQ record ends the step-0 call, the third call that collects nothing ends the input (kthread.c:119)"""
    L = O.lib()
    L.yo_reader_open.restype = C.c_void_p; L.yo_reader_open.argtypes = [C.c_char_p]
    L.yo_reader_next.restype = C.c_int64; L.yo_reader_next.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]
    L.yo_reader_close.argtypes = [C.c_void_p]
    r = L.yo_reader_open(path.encode())
    out, sum_len, n_rec = [], 0, 0
    seq, name = C.c_char_p(), C.c_char_p()
    while True:
        n = L.yo_reader_next(r, C.byref(seq), C.byref(name))
        if n == -1:
            break
        if n == -2:
            # "the call collected nothing": count.c:109 tests sum_len == 0 over records of >= k >= 1 bases, bseq.c:56 / qv.c:97
            # the number of records - the same thing wherever the reference can be (a call made of empty records only exists
            # for min_len = 0, where it is bseq_read's rule that applies)
            if n_rec == 0:
                workers -= 1
                if workers == 0:
                    break
            sum_len = n_rec = 0
            continue
        if n < min_len:
            continue
        out.append(seq.value)
        sum_len += n
        n_rec += 1
        if sum_len >= chunk:
            sum_len = n_rec = 0
    L.yo_reader_close(r)
    return out


def test_fastx_reader_matches_kseq_semantics():
    import gzip
    import test_gpu_parity as T
    fn = T._edge_file(os.path.join(util.TMP, "yakb_edge_rd.fa"))
    a, b = _records_product(fn), _records_oracle(fn)
    assert a == b and len(a) > 15 and a[-1] == -1
    gz = fn + ".gz"
    with gzip.open(gz, "wb") as f:
        f.write(open(fn, "rb").read())
    assert _records_product(gz) == a
    trunc = os.path.join(util.TMP, "yakb_trunc.fq")
    open(trunc, "w").write("@r1\nACGTACGT\n+\nIIII\n")
    assert _records_product(trunc)[-1] == -2 and _records_oracle(trunc)[-1] == -2   # kseq.h:190
    big = G.input_path("reads_q")
    assert _records_product(big) == _records_oracle(big)
    from yak_b200 import capi
    assert not capi.lib().yakb_fastx_open(b"/nonexistent/x.fa")


def test_cli_front_end_without_gpu():
    exe = os.path.join(ROOT, "yak_b200", "bin", "yak-b200")
    assert os.path.exists(exe)
    r = subprocess.run([exe, "version"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("0.1")
    assert subprocess.run([exe], capture_output=True).returncode == 1
    assert subprocess.run([exe, "count"], capture_output=True).returncode == 1
    r = subprocess.run([exe, "count", "-p9", "x.fa"], capture_output=True, text=True)          # main.c:43-46
    assert r.returncode == 1 and "-p should be at least 10" in r.stderr
    r = subprocess.run([exe, "count", "-k64", "x.fa"], capture_output=True, text=True)         # main.c:47-49
    assert r.returncode == 1 and "smaller than 64" in r.stderr
    y = os.path.join(G.HERE, "reads_c_k31_p10_b22.yak")
    r = subprocess.run([exe, "inspect", y], capture_output=True, text=True)                     # inspect.c:96-103
    assert r.returncode == 0
    case = next(c for c in G.GOLD["cases"] if c["input"] == "reads_c")
    assert hashlib.sha256(r.stdout.encode()).hexdigest() == case["inspect_sha256"]
    assert subprocess.run([exe, "bogus"], capture_output=True).returncode == 1


def test_qv_solver_matches_reference_output():
    """qv_solve.c (host FP64, CLI side) against the CT/FR/ER/CV/QV lines the reference printed."""
    src = os.path.join(ROOT, "yak_b200", "cli", "qv_solve.c")
    so = os.path.join(util.TMP, "yakb_qvsolve.so")
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-o", so, src, "-lm"], check=True)
    S = C.CDLL(so)

    class Qs(C.Structure):
        _fields_ = [("tot", C.c_int64), ("qv_raw", C.c_double), ("qv", C.c_double), ("cov", C.c_double), ("err", C.c_double),
                    ("fpr_lower", C.c_double), ("fpr_upper", C.c_double), ("adj_cnt", C.c_double * 1024)]
    ref = open(os.path.join(G.HERE, "qv_reads_c_ctg.txt")).read().splitlines()
    ct = {int(x[1]): x for x in (ln.split("\t") for ln in ref if ln.startswith("CT"))}
    hist = (C.c_int64 * 1024)(*[int(ct[i][2]) for i in range(1024)])
    cnt = (C.c_int64 * 1024)(*[int(ct[i][3]) for i in range(1024)])
    qs = Qs()
    S.yak_qv_solve.argtypes = [C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int, C.c_double, C.POINTER(Qs)]
    S.yak_qv_solve(hist, cnt, 31, 0.00004, C.byref(qs))
    lines = ["CT\t%d\t%d\t%d\t%.3f" % (i, hist[i], cnt[i], qs.adj_cnt[i]) for i in range(1023, -1, -1)]
    lines += ["FR\t%.3g\t%.3g" % (qs.fpr_lower, qs.fpr_upper), "ER\t%d\t%.3f" % (qs.tot, qs.err), "CV\t%.3f" % qs.cov,
              "QV\t%.3f\t%.3f" % (qs.qv_raw, qs.qv)]
    want = [ln for ln in ref if ln[:2] in ("CT", "FR", "ER", "CV", "QV")]
    assert lines == want


def _fill_all(path, cap, target, min_len, chunk=0, workers=None):
    """Drive yakb_fastx_fill the way yak_count does; returns the concatenated 'SEQ\\n' stream."""
    from yak_b200 import capi
    L = capi.lib()
    r = L.yakb_fastx_open(path.encode())
    if chunk:
        L.yakb_fastx_set_chunk(r, chunk)
    if workers is not None:
        L.yakb_fastx_set_workers(r, workers)
    out, nseq = bytearray(), 0
    buf = C.create_string_buffer(cap)
    while True:
        ns, done, need = C.c_int64(), C.c_int(), C.c_uint64()
        n = L.yakb_fastx_fill(r, buf, cap, target, min_len, C.byref(ns), C.byref(done), C.byref(need))
        if need.value:
            cap = need.value + 7
            buf = C.create_string_buffer(cap)
            continue
        out += buf.raw[:n]
        nseq += ns.value
        if done.value:
            break
    L.yakb_fastx_close(r)
    return bytes(out), nseq


@pytest.mark.parametrize("cap,target", [(1 << 20, 1 << 20), (4096, 1000), (700, 100), (64, 64)])
def test_bulk_fill_equals_record_reader(cap, target):
    import gzip
    import test_gpu_parity as T
    fn = T._edge_file(os.path.join(util.TMP, "yakb_edge_fill.fa"))
    for min_len in (0, 31):
        recs = [r for r in _records_oracle(fn) if not isinstance(r, int)]
        want = b"".join(s + b"\n" for _, s in recs if len(s) >= min_len)
        got, nseq = _fill_all(fn, cap, target, min_len)
        assert got == want
        assert nseq == sum(1 for _, s in recs if len(s) >= min_len)
    gz = fn + ".gz"
    with gzip.open(gz, "wb") as f:
        f.write(open(fn, "rb").read())
    assert _fill_all(gz, cap, target, 0)[0] == _fill_all(fn, cap, target, 0)[0]
    # FASTQ: final record without a trailing newline is valid; a short quality is dropped (kseq -2)
    p = os.path.join(util.TMP, "yakb_tail.fq")
    open(p, "w").write("@a\nACGTAC\n+\nIIIIII\n@b\nGGGTTT\n+\nIIIIII")
    assert _fill_all(p, cap, target, 0)[0] == b"ACGTAC\nGGGTTT\n"
    open(p, "w").write("@a\nACGTAC\n+\nIIIIII\n@b\nGGGTTT\n+\nIII\n@c\nAAAA\n+\nIIII\n")
    want = b"".join(s + b"\n" for s in _ref_flow(p, 0))     # the short quality of b swallows the header of c (kseq.h:224)
    assert _fill_all(p, cap, target, 0)[0] == want == b"ACGTAC\n"
    # reading goes on behind a truncated record the way the reference's pipeline does (count.c:93,109; kthread.c:119)
    open(p, "w").write("@a\nACGTAC\n+\nIIIIII\n@b\nGGGTTT\n+\nIIIIIIII\n@c\nAAAA\n+\nIIII\n@d\nCC\n+\nI\n@e\nTTTT\n+\nIIII\n")
    for chunk in (10_000_000, 6, 4):
        want = b"".join(s + b"\n" for s in _ref_flow(p, 0, chunk))
        assert _fill_all(p, cap, target, 0, chunk)[0] == want, chunk
    assert _ref_flow(p, 0) == [b"ACGTAC", b"AAAA"]    # b is dropped, c follows; d's short quality swallows the header of e
    big = G.input_path("reads_q")
    want = b"".join(s + b"\n" for _, s in (r for r in _records_oracle(big) if not isinstance(r, int)))
    assert _fill_all(big, max(cap, 4096), target, 0)[0] == want


def _pfill_all(path, block, threads, cap, target, min_len, chunk=0, workers=None):
    from yak_b200 import capi
    L = capi.lib()
    r = L.yakb_pfastx_open(path.encode(), block, threads)
    assert r
    if chunk:
        L.yakb_pfastx_set_chunk(r, chunk)
    if workers is not None:
        L.yakb_pfastx_set_flow(r, chunk or 10_000_000, workers, -1)
    out, nseq = bytearray(), 0
    buf = C.create_string_buffer(cap)
    while True:
        ns, done, need = C.c_int64(), C.c_int(), C.c_uint64()
        n = L.yakb_pfastx_fill(r, buf, cap, target, min_len, C.byref(ns), C.byref(done), C.byref(need))
        if need.value:
            cap = need.value + 11
            buf = C.create_string_buffer(cap)
            continue
        out += buf.raw[:n]
        nseq += ns.value
        if done.value:
            break
    redo = L.yakb_pfastx_redo(r)
    L.yakb_pfastx_close(r)
    return bytes(out), nseq, redo


@pytest.mark.parametrize("block,threads", [(64, 3), (257, 4), (4096, 8), (1 << 20, 2)])
def test_parallel_reader_is_exactly_the_sequential_reader(block, threads):
    """speculative block parsing + stitching must reproduce kseq semantics on any input, also when
    blocks are far smaller than records and every guess is wrong"""
    import test_gpu_parity as T
    files = [T._edge_file(os.path.join(util.TMP, "yakb_edge_par.fa")), G.input_path("reads_q")]
    p = os.path.join(util.TMP, "yakb_par_mixed.fq")   # quality lines starting with '>' and '@', CRLF, no final newline
    rng = np.random.default_rng(3)
    with open(p, "w", newline="") as f:
        for i in range(300):
            L_ = int(rng.integers(1, 120))
            s = "".join("ACGTN"[j] for j in rng.integers(0, 5, L_))
            q = "".join(">@+I#"[j] for j in rng.integers(0, 5, L_))
            eol = "\r\n" if i % 7 == 0 else "\n"
            f.write(f"@r{i} x{eol}{s}{eol}+{eol}{q}" + (eol if i < 299 else ""))
    files.append(p)
    from yak_b200 import capi
    for fn in files:
        for min_len in (0, 31):
            want, wn = _fill_all(fn, 1 << 20, 1 << 20, min_len)
            got, gn, redo = _pfill_all(fn, block, threads, 1 << 16, 1 << 15, min_len)
            assert got == want and gn == wn, (fn, block, threads, min_len, redo)
    trunc = os.path.join(util.TMP, "yakb_par_trunc.fq")
    open(trunc, "w").write("@a\nACGTAC\n+\nIIIIII\n@b\nGGGTTT\n+\nIII\n@c\nAAAA\n+\nIIII\n" * 3)
    want = b"".join(s + b"\n" for s in _ref_flow(trunc, 0))      # every b is dropped, reading goes on behind it
    assert _pfill_all(trunc, block, threads, 4096, 4096, 0)[0] == _fill_all(trunc, 4096, 4096, 0)[0] == want == b"ACGTAC\n" * 3
    assert not capi.lib().yakb_pfastx_open((files[0] + ".gz").encode(), 0, 0) or True


def test_parallel_reader_pool_copies_and_parse_ahead():
    """the streaming pool: block outputs >= 64 KB are copied into the caller's buffer by the workers while other
    blocks are parsed ahead; many small fills and few large ones must both reproduce the sequential reader"""
    from yak_b200 import synth
    p = os.path.join(util.TMP, "yakb_par_big.fq")
    with open(p, "wb") as f:
        f.write(synth.reads_file_bytes(5, 200_000, 9, 60_000, 150, 0.01, 3, fastq=True))
    want, wn = _fill_all(p, 1 << 22, 1 << 22, 31)
    for block, threads, cap, target in [(256 << 10, 8, 3 << 20, 2 << 20), (128 << 10, 5, 1 << 20, 300_000), (1 << 20, 3, 40 << 20, 40 << 20)]:
        for _ in range(3):
            got, gn, redo = _pfill_all(p, block, threads, cap, target, 31)
            assert got == want and gn == wn, (block, threads, cap, target, redo)


def test_parallel_reader_long_records_stay_linear():
    """a chromosome-sized record spans hundreds of blocks; the open record is carried, not re-copied, block after block"""
    import time
    rng = np.random.default_rng(8)
    big = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 24_000_000)])
    p = os.path.join(util.TMP, "yakb_par_long.fa")
    with open(p, "wb") as f:
        f.write(b">one line\n" + big + b"\n>wrapped\r\n" + b"\r\n".join(big[i:i + 80] for i in range(0, 3_000_000, 80)) + b"\r\n>tail\nACGT")
    want = big + b"\n" + big[:3_000_000] + b"\nACGT\n"
    t0 = time.time()
    got, gn, _ = _pfill_all(p, 64 << 10, 8, 64 << 20, 64 << 20, 0)
    dt = time.time() - t0
    assert got == want and gn == 3
    assert _fill_all(p, 64 << 20, 64 << 20, 0)[0] == want
    assert dt < 5.0, dt      # 24 MB in 64 KB blocks: quadratic re-copying took many seconds
    # long FASTQ reads: blocks that lie inside a quality string look like sequence to the mid-record speculation and
    # must be refused by the true state; quality characters include every character the grammar cares about
    q = os.path.join(util.TMP, "yakb_par_long.fq")
    with open(q, "wb") as f:
        for i, n in enumerate((300_000, 70_000, 500, 1_000_000)):
            seq = big[i * 1000:i * 1000 + n]
            qual = bytes(np.frombuffer(b"@+>I#", dtype=np.uint8)[rng.integers(0, 5, n)])
            if i == 1:   # multi-line sequence and quality (kseq allows both)
                seq_txt = b"\n".join(seq[j:j + 100] for j in range(0, n, 100))
                qual_txt = b"\n".join(qual[j:j + 7000] for j in range(0, n, 7000))
            else:
                seq_txt, qual_txt = seq, qual
            f.write(b"@r%d\n" % i + seq_txt + b"\n+\n" + qual_txt + b"\n")
    want_q, wn = _fill_all(q, 64 << 20, 64 << 20, 0)
    assert wn == 4 and len(want_q) == 300_000 + 70_000 + 500 + 1_000_000 + 4
    for block, threads in ((16 << 10, 8), (64 << 10, 3), (5000, 4)):
        got, gn, _ = _pfill_all(q, block, threads, 64 << 20, 64 << 20, 0)
        assert got == want_q and gn == wn, (block, threads)


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs) on a tiny bounded sample, here without a GPU"""
    import json
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1",
                        "--genome", "400000", "--chunk-reads", "3000", "--bf-shift", "24"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "events/s" and d["steps"] == 3 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["e2e"]["h2d_bytes_per_step"] == 0


def _random_fastx(rng, n_rec, long_lines=False, bad=0.0):
    """adversarial FASTA/FASTQ text: multi-line records, CRLF, blank lines, quality lines starting with '>' '@' '+',
    lower case / IUPAC, empty sequences, optional missing final newline"""
    out = []
    for i in range(n_rec):
        eol = "\r\n" if rng.random() < 0.2 else "\n"
        n = int(rng.integers(0, 60000 if long_lines and rng.random() < 0.3 else 400))
        seq = "".join("ACGTNacgtRY"[j] for j in rng.integers(0, 11, n))
        width = int(rng.integers(20, 5000 if long_lines else 120)) if rng.random() < 0.5 else max(n, 1)
        lines = [seq[j:j + width] for j in range(0, max(n, 1), width)]
        if rng.random() < 0.1:
            lines.insert(int(rng.integers(0, len(lines) + 1)), "")          # a blank line inside the sequence
        if rng.random() < 0.5:
            out.append(f">r{i} c{eol}" + eol.join(lines) + eol)
        else:
            qual = "".join("I>@+#"[j] for j in rng.integers(0, 5, n))
            if bad and n > 3 and rng.random() < bad:
                qual = qual[:int(rng.integers(1, n))] if rng.random() < 0.6 else qual + "I" * int(rng.integers(1, 9))   # truncated / too long
            qw = int(rng.integers(20, 5000 if long_lines else 120)) if rng.random() < 0.4 else max(n, 1)
            qlines = [qual[j:j + qw] for j in range(0, max(n, 1), qw)]
            out.append(f"@q{i}{eol}" + eol.join(lines) + eol + "+" + ("x" if rng.random() < 0.3 else "") + eol + eol.join(qlines) + eol)
    txt = "".join(out)
    if rng.random() < 0.3:
        txt = txt.rstrip("\r\n")
    if rng.random() < 0.15:
        txt = "\n\n" + txt
    return txt.encode()


@pytest.mark.parametrize("seed", list(range(24)) + [29, 31, 59, 107])   # the last four: calls made of empty records only
def test_parallel_reader_randomised_against_sequential(seed):
    """differential test of the parser pool (speculative record starts, mid-sequence speculation, carried records,
    pool copies, spill) against the sequential reader on adversarial inputs, block sizes and buffer sizes"""
    rng = np.random.default_rng(1000 + seed)
    long_lines = seed % 3 == 0
    p = os.path.join(util.TMP, f"yakb_par_rand{seed}.fx")
    with open(p, "wb") as f:
        f.write(_random_fastx(rng, int(rng.integers(5, 120)), long_lines, bad=(0.0, 0.1, 0.5)[seed % 3 if seed % 2 else 0]))
    for min_len in (0, 31):
        chunk = int(rng.choice([10_000_000, 3000, 200]))     # the reference's -K: decides what follows a truncated record
        want = b"".join(s + b"\n" for s in _ref_flow(p, min_len, chunk))
        wn = len(_ref_flow(p, min_len, chunk))
        assert _fill_all(p, 1 << 22, 1 << 22, min_len, chunk) == (want, wn), (seed, min_len, chunk)
        for _ in range(4):
            block = int(rng.choice([16, 61, 300, 4096, 9000, 70000])) if not long_lines else int(rng.choice([4096, 6000, 20000, 70000]))
            threads = int(rng.integers(1, 9))
            cap = int(rng.choice([1 << 22, 1 << 17, 70000]))
            target = int(rng.integers(1, cap + 1))
            got, gn, redo = _pfill_all(p, block, threads, cap, target, min_len, chunk)
            assert got == want and gn == wn, (seed, block, threads, cap, target, min_len, chunk, redo)
    # yak_recount's plain loop (count.c:176) ends at the first truncated record: workers = 0
    plain = b"".join(s + b"\n" for _, s in (r for r in _records_oracle(p) if not isinstance(r, int)))
    assert _fill_all(p, 1 << 22, 1 << 22, 0, workers=0)[0] == plain
    assert _pfill_all(p, 4096, 4, 1 << 22, 1 << 22, 0, workers=0)[0] == plain
    # the record readers themselves on the same adversarial file: product == oracle (kseq restatement) record by record,
    # and the oracle's whole count == the UNMODIFIED reference binary's .yak (where oracle/_ref is built), for several -K
    assert _records_product(p) == _records_oracle(p)
    if os.path.exists(O.REF_YAK):
        y = os.path.join(util.TMP, f"yakb_par_rand{seed}.yak")
        for chunk in (10_000_000, 3000, 200):
            O.ref_count(p, y, k=15, pre=10, extra=(f"-K{chunk}",))
            h, _ = O.count_file(p, k=15, pre=10, bf_shift=0, chunk_size=chunk)
            assert O.dump_bytes(h) == open(y, "rb").read(), (seed, chunk)
            O.lib().yo_ch_destroy(h)


@pytest.mark.parametrize("seed", range(10))
def test_oracle_randomised_parameters_against_reference_binary(seed):
    """the checker itself, checked: random k / -p / -b / -H / -K, one- and two-file runs, clean and adversarial reads -
    the oracle's .yak must be the unmodified reference's, byte for byte"""
    if not os.path.exists(O.REF_YAK):
        pytest.skip("oracle/_ref not built")
    from yak_b200 import synth
    rng = np.random.default_rng(7000 + seed)
    k = int(rng.choice([5, 13, 21, 27, 31, 32, 33, 47, 63]))
    pre = int(rng.integers(10, 15))
    b = int(rng.choice([0, 0, pre - 1, pre + 9, pre + 10, pre + 12, pre + 14]))
    nh = int(rng.choice([1, 2, 4, 4, 7, 8]))
    chunk = int(rng.choice([10_000_000, 50_000, 4000]))
    f1 = os.path.join(util.TMP, f"yakb_rnd{seed}_1.fx")
    f2 = os.path.join(util.TMP, f"yakb_rnd{seed}_2.fx")
    G_, n = 30_000, int(rng.integers(200, 1500))
    with open(f1, "wb") as f:
        f.write(synth.reads_file_bytes(50 + seed, G_, 60 + seed, n, 150, 0.01, 3, fastq=bool(seed % 2)))
        if seed % 3 == 0:
            f.write(_random_fastx(rng, 40, bad=0.2))
    with open(f2, "wb") as f:
        f.write(synth.reads_file_bytes(50 + seed, G_, 80 + seed, n // 2, 150, 0.02, 3, fastq=not seed % 2))
    two = b > 0 and seed % 2 == 0
    y = os.path.join(util.TMP, f"yakb_rnd{seed}.yak")
    O.ref_count(f1, y, k=k, pre=pre, bf_shift=b, bf_n_hash=nh, extra=(f"-K{chunk}",))
    if two:   # second input for pass 2 (main.c:57)
        cmd = [O.REF_YAK, "count", f"-k{k}", f"-p{pre}", "-t3", f"-H{nh}", f"-b{b}", f"-K{chunk}", "-o", y, f1, f2]
        subprocess.run(cmd, check=True, capture_output=True)
    h, _ = O.count_file(f1, k=k, pre=pre, bf_shift=b, bf_n_hash=nh, fn2=f2 if two else None, chunk_size=chunk)
    got, want = O.dump_bytes(h), open(y, "rb").read()
    assert got == want, (k, pre, b, nh, chunk, two, util.explain_diff(got, want))
    O.lib().yo_ch_destroy(h)


@pytest.mark.parametrize("seed", range(8))
def test_ref_flow_rule_of_the_library(seed):
    """csrc/ref_flow.h (used by the sequential reader and yak_qv) against the flows pinned to the reference: `_ref_flow`
    (count.c, checked against the reference binary above) and test_scan_cpu.batches (bseq_read, checked against the
    reference's scanner stdout), on random streams of good and truncated FASTQ records"""
    import test_scan_cpu as TS
    from yak_b200 import capi
    L = capi.lib()
    rng = np.random.default_rng(9000 + seed)
    n = int(rng.integers(5, 60))
    lens = [(-2 if rng.random() < (0.15, 0.4, 0.7)[seed % 3] else int(rng.integers(1, 80))) for _ in range(n)]
    p = os.path.join(util.TMP, f"yakb_flow{seed}.fq")
    with open(p, "w") as f:
        for i, ln in enumerate(lens):
            m = ln if ln > 0 else int(rng.integers(4, 40))
            s = "".join("ACGT"[j] for j in rng.integers(0, 4, m))
            f.write(f"@r{i}\n{s}\n+\n" + "I" * (m + (1 if ln < 0 else 0)) + "\n")     # one quality character too many = kseq's -2
    seqs = [ln.strip() for ln in open(p).read().split("\n")[1::4]]
    arr = (C.c_int64 * n)(*lens)
    for chunk in (10_000_000, 150, 40):
        for min_len in (0, 31):
            out = (C.c_uint8 * n)()
            L.yakb_ref_flow_sim(arr, n, 3, chunk, min_len, out)
            mine = [seqs[i].encode() for i in range(n) if out[i] and lens[i] >= min_len]
            assert mine == _ref_flow(p, min_len, chunk), (seed, chunk, min_len)
        out = (C.c_uint8 * n)()
        L.yakb_ref_flow_sim(arr, n, 2, chunk, 0, out)
        O_ = O.lib()
        O_.yo_reader_open.restype = C.c_void_p; O_.yo_reader_open.argtypes = [C.c_char_p]
        O_.yo_reader_next.restype = C.c_int64; O_.yo_reader_next.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]
        O_.yo_reader_close.argtypes = [C.c_void_p]
        want = [s for _, ss in TS.batches(O_, p, chunk) for s in ss]
        assert [seqs[i].encode() for i in range(n) if out[i]] == want, (seed, chunk)


def test_qv_solver_randomised_against_reference_function():
    """cli/qv_solve.c against the reference's own yak_qv_solve (qv.c:146-244, called through oracle/_ref/libyakref.so) on
    random k-mer spectra: every printed figure identical, every double within 1e-9 relative"""
    if not os.path.exists(O.REF_LIB):
        pytest.skip("oracle/_ref not built")
    src = os.path.join(ROOT, "yak_b200", "cli", "qv_solve.c")
    so = os.path.join(util.TMP, "yakb_qvsolve2.so")
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-o", so, src, "-lm"], check=True)

    class Qs(C.Structure):
        _fields_ = [("tot", C.c_int64), ("qv_raw", C.c_double), ("qv", C.c_double), ("cov", C.c_double), ("err", C.c_double),
                    ("fpr_lower", C.c_double), ("fpr_upper", C.c_double), ("adj_cnt", C.c_double * 1024)]
    libs = [C.CDLL(so), C.CDLL(O.REF_LIB)]
    for S in libs:
        S.yak_qv_solve.argtypes = [C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int, C.c_double, C.POINTER(Qs)]
        S.yak_qv_solve.restype = C.c_int
    rng = np.random.default_rng(77)
    for trial in range(60):
        cov = float(rng.choice([4, 12, 21, 35, 80, 300]))
        G_ = int(rng.choice([1e5, 3e6, 2e8]))
        x = np.arange(1024)
        with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
            logp = x * np.log(cov) - cov - np.array([float(np.sum(np.log(np.arange(1, i + 1)))) for i in x])
        peak = np.exp(logp)
        hist = (G_ * peak + G_ * 0.02 * 0.5 ** x * (x > 0)).astype(np.int64)     # genomic peak + error k-mers
        hist[1023] += int(G_ * 1e-4)
        if trial % 7 == 0:
            hist = hist // int(rng.choice([1000, 100000]))                       # thin spectra too (an all-zero one makes the reference read cnt[-1])
        if trial % 11 == 0:
            hist[int(rng.integers(2, 1000))] = 0
        cnt = (hist * float(rng.uniform(0.5, 1.2))).astype(np.int64)
        cnt[0] = int(rng.integers(0, max(2, G_ // 1000)))
        if cnt[2:1023].max() <= 0:
            continue
        k = int(rng.choice([15, 21, 31]))
        fpr = float(rng.choice([0.00004, 0.001, 0.0]))
        res = []
        for S in libs:
            qs = Qs()
            rc = S.yak_qv_solve((C.c_int64 * 1024)(*hist.tolist()), (C.c_int64 * 1024)(*cnt.tolist()), k, fpr, C.byref(qs))
            res.append((rc, qs))
        (r0, a), (r1, b) = res
        assert r0 == r1, trial
        fa = ["%.3f" % a.adj_cnt[i] for i in range(1024)] + ["%.3g %.3g" % (a.fpr_lower, a.fpr_upper), "%d %.3f" % (a.tot, a.err), "%.3f" % a.cov, "%.3f %.3f" % (a.qv_raw, a.qv)]
        fb = ["%.3f" % b.adj_cnt[i] for i in range(1024)] + ["%.3g %.3g" % (b.fpr_lower, b.fpr_upper), "%d %.3f" % (b.tot, b.err), "%.3f" % b.cov, "%.3f %.3f" % (b.qv_raw, b.qv)]
        assert fa == fb, trial
        for u, v in zip([a.qv_raw, a.qv, a.cov, a.err, a.fpr_lower, a.fpr_upper] + list(a.adj_cnt), [b.qv_raw, b.qv, b.cov, b.err, b.fpr_lower, b.fpr_upper] + list(b.adj_cnt)):
            assert (u == v) or (np.isnan(u) and np.isnan(v)) or abs(u - v) <= 1e-9 * max(abs(u), abs(v)), (trial, u, v)


def test_cli_inspect_randomised_against_reference_binary():
    """`yak-b200 inspect x.yak` (inspect.c:8-106, single-file mode: no table, no GPU) against the reference binary's stdout on
    tables of random k / -p / -b"""
    if not os.path.exists(O.REF_YAK):
        pytest.skip("oracle/_ref not built")
    from yak_b200 import synth
    exe = os.path.join(ROOT, "yak_b200", "bin", "yak-b200")
    rng = np.random.default_rng(31)
    fa = os.path.join(util.TMP, "yakb_insp.fa")
    with open(fa, "wb") as f:
        f.write(synth.reads_file_bytes(3, 20_000, 4, 3000, 150, 0.01, 2))
    for trial in range(6):
        k, pre = int(rng.choice([15, 31, 47, 63])), int(rng.integers(10, 14))
        b = int(rng.choice([0, pre + 10]))
        y = os.path.join(util.TMP, "yakb_insp.yak")
        h, _ = O.count_file(fa, k=k, pre=pre, bf_shift=b)
        assert O.lib().yo_ch_dump(h, y.encode()) == 0
        O.lib().yo_ch_destroy(h)
        mine = subprocess.run([exe, "inspect", y], capture_output=True)
        ref = subprocess.run([O.REF_YAK, "inspect", y], capture_output=True)
        assert mine.returncode == ref.returncode == 0
        assert mine.stdout == ref.stdout and len(ref.stdout) > 100, (k, pre, b)
    # error paths: missing file, wrong magic (inspect.c:18-30)
    bad = os.path.join(util.TMP, "yakb_insp_bad.yak")
    open(bad, "wb").write(b"NOPE" + b"\0" * 64)
    for args in (["inspect"], ["inspect", "/nonexistent.yak"], ["inspect", bad]):
        m = subprocess.run([exe] + args, capture_output=True)
        r = subprocess.run([O.REF_YAK] + args, capture_output=True)
        assert (m.returncode != 0) == (r.returncode != 0) and m.stdout == r.stdout, args


def test_inspect_two_file_logic_against_reference_binary():
    """`yak inspect [-m] in1.yak in2.yak` (inspect.c:31-95, stored keys passed to yak_ch_get: quirk Q7): cli/inspect_logic.c over
    the oracle's yak_ch_get equals the reference binary's stdout on tables of different inputs, k, -p and -m"""
    if not os.path.exists(O.REF_YAK):
        pytest.skip("oracle/_ref not built")
    from yak_b200 import synth
    so = os.path.join(util.TMP, "yakb_inspect_logic_test.so")
    cli = os.path.join(ROOT, "yak_b200", "cli")
    subprocess.run(["gcc", "-O2", "-Wall", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-o", so,
                    os.path.join(cli, "inspect_logic.c"), os.path.join(cli, "qv_solve.c"), "-lm"], check=True)
    logic = C.CDLL(so)
    LOOKUP = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_int32))
    logic.yakb_inspect_run.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int64), LOOKUP, C.c_void_p, C.c_int64]
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p; libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    L = O.lib()
    rng = np.random.default_rng(77)
    fa, fb = os.path.join(util.TMP, "yakb_insp2_a.fa"), os.path.join(util.TMP, "yakb_insp2_b.fa")
    with open(fa, "wb") as f:
        f.write(synth.reads_file_bytes(3, 8_000, 4, 1200, 150, 0.01, 2))
    with open(fb, "wb") as f:
        f.write(synth.reads_file_bytes(3, 8_000, 5, 900, 150, 0.02, 2))
    y1, y2, outp = (os.path.join(util.TMP, n) for n in ("yakb_insp2_1.yak", "yakb_insp2_2.yak", "yakb_insp2.out"))
    for trial in range(6):
        k = int(rng.choice([15, 21, 31]))
        p1, p2 = int(rng.integers(10, 13)), int(rng.integers(10, 13))
        for fn, y, pre in ((fa, y1, p1), (fb, y2, p2)):
            h, _ = O.count_file(fn, k=k, pre=pre, bf_shift=0)
            assert L.yo_ch_dump(h, y.encode()) == 0
            L.yo_ch_destroy(h)
        h2 = L.yo_ch_restore(y2.encode())
        hist = (C.c_int64 * 1024)()
        L.yo_ch_hist(h2, hist)

        def lookup(ctx, n, x, out):
            for i in range(n):
                out[i] = L.yo_ch_get(h2, x[i])
            return 0
        flags = [] if trial % 2 == 0 else ["-m%d" % int(rng.integers(1, 40))]
        m = int(flags[0][2:]) if flags else 20
        fp = libc.fopen(outp.encode(), b"w")
        assert logic.yakb_inspect_run(fp, y1.encode(), m, hist, LOOKUP(lookup), None, int(rng.choice([1, 1000, 1 << 20]))) == 0
        libc.fclose(fp)
        L.yo_ch_destroy(h2)
        ref = subprocess.run([O.REF_YAK, "inspect"] + flags + [y1, y2], capture_output=True)
        assert ref.returncode == 0
        got = open(outp, "rb").read()
        assert got == ref.stdout and got.count(b"\n") > 3, (trial, k, p1, p2, flags)


@pytest.mark.parametrize("seed", range(4))
def test_fill_shortcut_for_four_line_fastq_equals_the_state_machine(seed):
    """fill()'s shortcut for whole four-line FASTQ records (fastx.cpp) against the oracle's record reader: records of random
    length next to everything that must NOT take the shortcut (CRLF, multi-line, blank lines, short / long qualities that
    are legal, FASTA records in between), small destination buffers, the length filter"""
    rng = np.random.default_rng(900 + seed)
    recs = []
    for i in range(3000):
        ln = int(rng.integers(1, 260))
        s = bytes(rng.choice(np.frombuffer(b"ACGTNacgtn", dtype=np.uint8), ln))
        kind = int(rng.integers(0, 12))
        if kind == 0:      # CRLF
            recs.append(b"@r%d x\r\n" % i + s + b"\r\n+\r\n" + b"I" * ln + b"\r\n")
        elif kind == 1:    # two sequence lines, two quality lines
            h = ln // 2
            recs.append(b"@r%d\n" % i + s[:h] + b"\n" + s[h:] + b"\n+\n" + b"I" * h + b"\n" + b"I" * (ln - h) + b"\n")
        elif kind == 2:    # blank line before the sequence
            recs.append(b"@r%d\n\n" % i + s + b"\n+r%d\n" % i + b"I" * ln + b"\n")
        elif kind == 3:    # FASTA
            recs.append(b">f%d\n" % i + s + b"\n")
        elif kind == 4:    # quality characters that look like headers
            recs.append(b"@r%d\n" % i + s + b"\n+\n" + b"@" * ln + b"\n")
        else:
            recs.append(b"@r%d some comment\n" % i + s + b"\n+\n" + b"I" * ln + b"\n")
    fn = os.path.join(util.TMP, "yakb_fastpath.fq")
    with open(fn, "wb") as f:
        f.write(b"".join(recs))
    want_recs = [r for r in _records_oracle(fn) if not isinstance(r, int)]
    assert len(want_recs) == 3000
    for min_len in (0, 31, 200):
        want = b"".join(q + b"\n" for _, q in want_recs if len(q) >= min_len)
        for cap, target in ((1 << 20, 1 << 20), (300, 100), (5000, 5000)):
            got, nseq = _fill_all(fn, cap, target, min_len)
            assert got == want, (min_len, cap)
            assert nseq == sum(1 for _, q in want_recs if len(q) >= min_len)
            got, nseq, _ = _pfill_all(fn, 3000, 3, cap, target, min_len)
            assert got == want, ("pool", min_len, cap)


@pytest.mark.parametrize("tail", [b"\n", b"", b"\r\n"])
def test_last_record_waiting_in_the_carry_is_not_lost(tail):
    """a last record that does not fit what is left of the caller's buffer is closed by the end of the input while it sits in the
    reader's carry buffer: fill() must not report `done` before it has handed it over (FASTA and FASTQ, both readers, gzip)"""
    import gzip
    rng = np.random.default_rng(5)
    big = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 9000))
    for text in (b">a\nACGTACGT\n>b\n" + big + tail, b"@a\nACGTACGT\n+\nIIIIIIII\n@b\n" + big + b"\n+\n" + b"I" * 9000 + tail,
                 b">only\n" + big[:5000] + b"\n" + big[5000:] + tail):
        fn = os.path.join(util.TMP, "yakb_lastcarry.fx")
        with open(fn, "wb") as f:
            f.write(text)
        with gzip.open(fn + ".gz", "wb") as f:
            f.write(text)
        want = b"".join(q + b"\n" for _, q in (r for r in _records_oracle(fn) if not isinstance(r, int)))
        assert want.endswith(big + b"\n")
        for cap, target in ((9004, 9004), (9010, 100), (1 << 20, 1 << 20), (9001, 9001)):
            assert _fill_all(fn, cap, target, 0)[0] == want, (cap, "sequential")
            assert _fill_all(fn + ".gz", cap, target, 0)[0] == want, (cap, "gzip")
            assert _pfill_all(fn, 2000, 3, cap, target, 0)[0] == want, (cap, "pool")


def _parse_yak_like_the_reference(data: bytes, mode, min_cnt, mid_cnt):
    """htab.c:419-472 on bytes, with its unchecked freads: a missing header is an empty sub-table, a short key array ends the file"""
    if len(data) < 4:
        return -1
    if data[:4] != b"YAK\x02":
        return -2
    if len(data) < 16:
        return -1
    k, pre, cb = struct.unpack_from("<III", data, 4)
    if cb != 10:
        return -3
    caps, off, keys, o = [], [0], [], 16
    for s in range(1 << pre):
        cap = size = 0
        if o + 8 <= len(data):
            cap, size = struct.unpack_from("<II", data, o)
            o += 8
        else:
            o = len(data)
        got = min(size, (len(data) - o) // 8)
        ks = np.frombuffer(data, dtype="<u8", count=got, offset=o).copy()
        o = o + 8 * got if got == size else len(data)
        if mode in (2, 3):        # YAK_LOAD_TRIOBIN1/2
            cnt = (ks & np.uint64(1023)).astype(np.int64)
            shift = 0 if mode == 2 else 2
            x = np.where(cnt >= mid_cnt, 2 << shift, np.where(cnt >= min_cnt, 1 << shift, -1))
            keep = x >= 0
            ks = (ks[keep] & ~np.uint64(1023)) | x[keep].astype(np.uint64)
        elif mode in (4, 5, 6):   # YAK_LOAD_SEXCHR1/2/3
            ks = (ks & ~np.uint64(1023)) | np.uint64(1 << (mode - 4))
        caps.append(cap)
        keys.append(ks)
        off.append(off[-1] + len(ks))
    return k, pre, caps, off, np.concatenate(keys) if keys else np.zeros(0, np.uint64)


def test_yak_file_reader_of_restore_equals_a_plain_parse():
    """the host side of yak_ch_restore_core (header walk + key arrays read by several threads, capi.cu) against a plain parse:
    whole files, every load mode, files cut anywhere, a pipe"""
    import threading
    from yak_b200 import capi, synth
    L = capi.lib()
    L.yakb_yak_file_read.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                     C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.POINTER(C.c_uint64))]
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]

    def read(fn, mode=1, mn=0, md=0, threads=0):
        k, pre = C.c_uint32(), C.c_uint32()
        caps, off, keys = C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint64)()
        rc = L.yakb_yak_file_read(fn.encode(), mode, mn, md, threads, C.byref(k), C.byref(pre), C.byref(caps), C.byref(off), C.byref(keys))
        if rc != 0:
            return rc
        P = 1 << pre.value
        o = [off[i] for i in range(P + 1)]
        out = (k.value, pre.value, [caps[i] for i in range(P)], o, np.ctypeslib.as_array(keys, shape=(max(o[-1], 1),))[:o[-1]].copy())
        for p in (caps, off, keys):
            libc.free(p)
        return out

    def same(a, b):
        if isinstance(a, int) or isinstance(b, int):
            return a == b
        return a[:4] == b[:4] and np.array_equal(a[4], b[4])
    rng = np.random.default_rng(17)
    fa = os.path.join(util.TMP, "yakb_yfr.fa")
    with open(fa, "wb") as f:
        f.write(synth.reads_file_bytes(3, 400_000, 4, 60_000, 150, 0.01, 2))
    y = os.path.join(util.TMP, "yakb_yfr.yak")
    for k, pre in ((31, 10), (21, 12)):
        h, _ = O.count_file(fa, k=k, pre=pre, bf_shift=0)
        assert O.lib().yo_ch_dump(h, y.encode()) == 0
        O.lib().yo_ch_destroy(h)
        data = open(y, "rb").read()
        assert len(data) > (8 << 20)          # several threads take part
        for mode, mn, md in ((1, 0, 0), (2, 2, 5), (3, 1, 3), (4, 0, 0), (5, 0, 0), (6, 0, 0)):
            want = _parse_yak_like_the_reference(data, mode, mn, md)
            assert same(read(y, mode, mn, md, 0), want) and same(read(y, mode, mn, md, 1), want) and same(read(y, mode, mn, md, 5), want)
        cut = os.path.join(util.TMP, "yakb_yfr_cut.yak")
        for n in [0, 3, 4, 15, 16, 20, 24, 31] + [int(x) for x in rng.integers(16, len(data), 12)]:
            open(cut, "wb").write(data[:n])
            assert same(read(cut, 1, 0, 0, 3), _parse_yak_like_the_reference(data[:n], 1, 0, 0)), n
        open(cut, "wb").write(b"NOPE" + data[4:100]);
        assert read(cut) == -2
        open(cut, "wb").write(data[:12] + struct.pack("<I", 9) + data[16:100])
        assert read(cut) == -3
        assert read("/nonexistent/x.yak") == -1
    # a pipe cannot be read by offsets: front to back, same result
    fifo = os.path.join(util.TMP, "yakb_yfr.fifo")
    if os.path.exists(fifo):
        os.unlink(fifo)
    os.mkfifo(fifo)
    small = data[:16 + 8 * 4096 + 200_000]
    t = threading.Thread(target=lambda: open(fifo, "wb").write(small))
    t.start()
    got = read(fifo, 1, 0, 0, 4)
    t.join()
    os.unlink(fifo)
    assert same(got, _parse_yak_like_the_reference(small, 1, 0, 0))


def test_cpu_read_generator_equals_the_python_specification():
    """oracle/_bin/synthgen (bench.py's reference arm writes its input with it) against yak_b200/synth.py: same bytes, same event count"""
    import subprocess
    from yak_b200 import synth
    gen = os.path.join(ROOT, "oracle", "_bin", "synthgen")
    if not os.path.exists(gen):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    for fmt, fastq in ((1, False), (2, True)):
        out = os.path.join(util.TMP, f"yakb_synthgen{fmt}.txt")
        r = subprocess.run([gen, "7", "200000", "11", "5", "3000", "150", "0.005", "1", str(fmt), "31", out, "3"], capture_output=True, text=True, check=True)
        codes = synth.read_codes(7, 200000, 11, 5, 3000, 150, 0.005, 1)
        exp = bytearray()
        for row in synth.codes_to_ascii(codes):
            exp += (b"@r\n" if fastq else b">r\n") + row.tobytes() + b"\n" + ((b"+\n" + b"I" * 150 + b"\n") if fastq else b"")
        assert open(out, "rb").read() == bytes(exp)
        assert int(r.stdout) == synth.count_events(codes, 31)


def test_record_boundary_guess_on_adversarial_quality_lines():
    """yakb_record_start_before (csrc/capi.cu; host code, no GPU): where the device ingest cuts its batches.  Quality lines that start
    with '@' must not be taken for headers: two lines behind a header comes a '+' line, two lines behind such a quality line come bases."""
    import ctypes as C
    from yak_b200 import capi
    L = capi.lib()
    rng = np.random.default_rng(3)
    recs, starts, off = [], [], 0
    for i in range(400):
        n = int(rng.integers(1, 80))
        seq = bytes(rng.choice(list(b"ACGT"), n).astype(np.uint8))
        qual = (b"@" if i % 2 else b"+") + bytes(rng.integers(33, 74, n - 1).astype(np.uint8)) if n > 1 else b"@"
        rec = b"@r%d\n" % i + seq + b"\n+\n" + qual + b"\n"
        starts.append(off); off += len(rec); recs.append(rec)
    text = b"".join(recs)
    buf = C.create_string_buffer(text, len(text))
    starts = np.array(starts)
    for pos in list(rng.integers(1, len(text), 300)) + [len(text) - 1]:
        lo = int(rng.integers(0, pos))
        got = L.yakb_record_start_before(buf, lo, int(pos), len(text), 4)
        cand = starts[(starts > lo) & (starts <= pos)]
        # the last record start in (lo, pos] whose '+' line is still inside the text; lo when there is none
        want = lo
        for s in cand[::-1]:
            want = int(s)
            break
        assert got == want, (lo, pos, got, want)
    fa = b"".join(b">c%d\n" % i + b"ACGT" * 5 + b"\n" for i in range(50))
    fbuf = C.create_string_buffer(fa, len(fa))
    assert L.yakb_record_start_before(fbuf, 0, 100, len(fa), 2) == max(i for i in range(0, 101) if fa[i:i + 1] == b">" and (i == 0 or fa[i - 1:i] == b"\n"))
