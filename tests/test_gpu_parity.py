"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit-exact .yak bytes."""
import ctypes as C
import os

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def _count_both(oracle, yakb, fn, k, pre, b, fn2=None, chunk_size=10_000_000):
    ho, ne = oracle.count_file(fn, k=k, pre=pre, bf_shift=b, fn2=fn2, chunk_size=chunk_size)
    ref = oracle.dump_bytes(ho)
    hg = yakb.count_file(fn, k=k, pre=pre, bf_shift=b, fn2=fn2, chunk_size=chunk_size)
    assert hg
    mine = yakb.dump_bytes(hg)
    return ho, hg, ref, mine, ne


@pytest.fixture(scope="module")
def reads_fa():
    return util.write_reads(os.path.join(util.TMP, "yakb_reads.fa"))


@pytest.fixture(scope="module")
def reads_fq():
    return util.write_reads(os.path.join(util.TMP, "yakb_reads.fq"), seed_r=12, n_reads=3000, fastq=True)


@pytest.mark.parametrize("k,pre,b", [(31, 12, 0), (31, 10, 0), (21, 11, 0), (15, 10, 0), (47, 12, 0), (63, 10, 0),
                                     (31, 12, 22), (31, 10, 20), (31, 12, 24), (27, 11, 21), (63, 10, 21), (31, 12, 12),
                                     (31, 14, 0), (31, 14, 25), (31, 16, 0), (31, 16, 26), (32, 10, 0), (4, 10, 0)])
def test_count_matches_oracle(oracle, yakb, reads_fa, k, pre, b):
    ho, hg, ref, mine, ne = _count_both(oracle, yakb, reads_fa, k, pre, b)
    try:
        assert mine == ref, util.explain_diff(mine, ref)
        assert hg.contents.tot == ho.contents.tot
    finally:
        yakb.lib().yak_ch_destroy(hg)
        oracle.lib().yo_ch_destroy(ho)


def test_count_fastq_two_files(oracle, yakb, reads_fa, reads_fq):
    # pass 2 over a different file (main.c:57)
    ho, hg, ref, mine, _ = _count_both(oracle, yakb, reads_fq, 31, 12, 23, fn2=reads_fa)
    try:
        assert mine == ref, util.explain_diff(mine, ref)
    finally:
        yakb.lib().yak_ch_destroy(hg)
        oracle.lib().yo_ch_destroy(ho)


@pytest.mark.parametrize("b", [0, 22])
def test_many_small_chunks(oracle, yakb, reads_fa, b, monkeypatch):
    # results must not depend on how the input is cut into batches
    # (a batch is max(YAKB_BATCH, opt->chunk_size) bases: both must be small; 1 Mbp of reads in 20+ batches, and the
    # parser pool hands each out in pieces of the staging buffer)
    monkeypatch.setenv("YAKB_BATCH", "50000")
    ho, hg, ref, mine, _ = _count_both(oracle, yakb, reads_fa, 31, 12, b, chunk_size=40_000)
    try:
        assert mine == ref, util.explain_diff(mine, ref)
    finally:
        yakb.lib().yak_ch_destroy(hg)
        oracle.lib().yo_ch_destroy(ho)


def _edge_file(path):
    rng = np.random.default_rng(5)
    def rnd(n):
        return "".join("ACGT"[i] for i in rng.integers(0, 4, n))
    recs = []
    recs.append(">short\n" + rnd(20) + "\n")                       # shorter than k: dropped (count.c:95)
    recs.append(">exact31\n" + rnd(31) + "\n")
    recs.append(">lower case\n" + rnd(200).lower() + "\n")
    s = rnd(500)
    recs.append(">multi line comment here\n" + "\n".join(s[i:i + 60] for i in range(0, 500, 60)) + "\n")
    recs.append(">withN\n" + rnd(40) + "N" + rnd(30) + "NN" + rnd(100) + "\n")
    recs.append(">iupac\n" + rnd(50) + "RYKM" + rnd(50) + "\n")
    recs.append(">uracil\n" + rnd(80).replace("T", "U") + "\n")
    recs.append(">crlf\r\n" + rnd(100) + "\r\n" + rnd(50) + "\r\n")
    recs.append(">polyA\n" + "A" * 3000 + "\n")                    # one k-mer 2970 times: saturates at 1023
    recs.append(">dinuc\n" + "AC" * 700 + "\n")
    recs.append(">empty\n\n")
    recs.append(">dup1\n" + s + "\n>dup2\n" + s + "\n")
    recs.append("\n\n>blank lines before\n\n" + rnd(90) + "\n\n")
    recs.append(">palin\n" + "ACGT" * 40 + "\n")
    fq = "@q1 c\n" + rnd(100) + "\n+\n" + "@" * 100 + "\n@q2\n" + rnd(64) + "\n+q2\n" + "+" * 64 + "\n"
    with open(path, "w", newline="") as f:
        f.write("".join(recs) + fq)
    return path


@pytest.mark.parametrize("k,pre,b", [(31, 10, 0), (31, 10, 19), (21, 10, 0), (33, 10, 0), (5, 10, 0)])
def test_edge_case_records(oracle, yakb, k, pre, b):
    fn = _edge_file(os.path.join(util.TMP, "yakb_edge.fa"))
    ho, hg, ref, mine, ne = _count_both(oracle, yakb, fn, k, pre, b)
    try:
        assert ne > 0
        assert mine == ref, util.explain_diff(mine, ref)
    finally:
        yakb.lib().yak_ch_destroy(hg)
        oracle.lib().yo_ch_destroy(ho)


def test_record_larger_than_staging_buffer(oracle, yakb, monkeypatch):
    """batches of 1000 bases against records of 20 and 45 kbp: yak_count grows its staging buffers and parses again"""
    monkeypatch.setenv("YAKB_BATCH", "1000")
    rng = np.random.default_rng(6)
    fn = os.path.join(util.TMP, "yakb_bigrec.fa")
    with open(fn, "w") as f:
        for i, n in enumerate((300, 20_000, 150, 45_000, 31, 6000)):
            f.write(f">r{i}\n" + "".join("ACGT"[j] for j in rng.integers(0, 4, n)) + "\n")
    for b in (0, 20):
        ho, hg, ref, mine, _ = _count_both(oracle, yakb, fn, 31, 10, b, chunk_size=1000)
        try:
            assert mine == ref, util.explain_diff(mine, ref)
        finally:
            yakb.lib().yak_ch_destroy(hg)
            oracle.lib().yo_ch_destroy(ho)


@pytest.mark.parametrize("seed,chunk", [(1, 10_000_000), (1, 3000), (2, 200), (5, 10_000_000), (7, 700)])
def test_truncated_fastq_records_like_the_reference(oracle, yakb, seed, chunk, monkeypatch):
    """behind a FASTQ record with a truncated quality the reference reads on or stops depending on its pipeline state and -K
    (count.c:93,109,162; kthread.c:119; the oracle's restatement is checked against the reference binary on the CPU)"""
    import test_oracle_cpu as T
    rng = np.random.default_rng(4000 + seed)
    fn = os.path.join(util.TMP, f"yakb_badq{seed}.fq")
    with open(fn, "wb") as f:
        f.write(T._random_fastx(rng, 150, long_lines=seed == 5, bad=0.3))
    for serial in ("", "1"):
        if serial:
            monkeypatch.setenv("YAKB_SERIAL_PARSE", "1")
        for b in (0, 19):
            ho, hg, ref, mine, _ = _count_both(oracle, yakb, fn, 15, 10, b, chunk_size=chunk)
            try:
                assert mine == ref, (serial, b, util.explain_diff(mine, ref))
            finally:
                yakb.lib().yak_ch_destroy(hg)
                oracle.lib().yo_ch_destroy(ho)


def test_empty_and_missing_input(yakb):
    L = yakb.lib()
    o = yakb.copt(31, 10)
    assert not L.yak_count(b"/nonexistent/file.fa", C.byref(o), None)      # count.c:152
    fn = os.path.join(util.TMP, "yakb_empty.fa")
    open(fn, "w").close()
    h = L.yak_count(fn.encode(), C.byref(o), None)
    assert h and h.contents.tot == 0
    data = yakb.dump_bytes(h)
    assert len(data) == 16 + 8 * 1024 and data[16:] == bytes(8 * 1024)     # capacity 0, size 0 everywhere
    L.yak_ch_destroy(h)
    assert not L.yak_ch_init(31, 9, 4, 0)                                    # htab.c:17


def test_insert_list_get_hist_clear(oracle, yakb, reads_fa):
    """The reference's own call pattern: per-sub-table lists into yak_ch_insert_list (count.c:82)."""
    OL, L = oracle.lib(), yakb.lib()
    k, pre = 31, 10
    seqs = [ln.strip() for ln in open(reads_fa) if not ln.startswith(">")][:1500]
    ev = []
    for s in seqs:
        buf = (C.c_uint64 * len(s))()
        n = OL.yo_extract(k, len(s), s.encode(), buf)
        ev.append(np.frombuffer(buf, dtype=np.uint64, count=n).copy())
    ev = np.concatenate(ev)
    ho = OL.yo_ch_init(k, pre, 4, 0)
    hg = L.yak_ch_init(k, pre, 4, 0)
    half = len(ev) // 2
    for part in (ev[:half], ev[half:]):                      # two "chunks"
        sub = (part & np.uint64((1 << pre) - 1)).astype(np.int64)
        order = np.argsort(sub, kind="stable")
        part_s, sub_s = part[order], sub[order]
        bounds = np.flatnonzero(np.diff(sub_s)) + 1
        for lst in np.split(part_s, bounds):
            a, p = util.u64_array(lst)
            n1 = OL.yo_ch_insert_list(ho, 1, len(a), p)
            n2 = L.yak_ch_insert_list(hg, 1, len(a), p)
            assert n1 == n2
    # a list with foreign elements: only those sharing a[0]'s sub-table count (htab.c:61)
    a, p = util.u64_array(ev[:4000])
    assert OL.yo_ch_insert_list(ho, 1, len(a), p) == L.yak_ch_insert_list(hg, 1, len(a), p)
    assert yakb.dump_bytes(hg) == oracle.dump_bytes(ho)
    # lookups, present and absent
    probe = np.concatenate([ev[::97], ev[::89] ^ np.uint64(0x5555555555)])
    a, p = util.u64_array(probe)
    got = np.zeros(len(a), dtype=np.int32)
    assert L.yakb_ch_get_batch(hg, len(a), p, got.ctypes.data_as(C.POINTER(C.c_int32))) == 0
    want = np.array([OL.yo_ch_get(ho, int(x)) for x in a], dtype=np.int32)
    assert np.array_equal(got, want)
    assert L.yak_ch_get(hg, int(a[0])) == want[0] and L.yak_ch_get(hg, int(a[-1])) == want[-1]
    h1, h2 = (C.c_int64 * 1024)(), (C.c_int64 * 1024)()
    OL.yo_ch_hist(ho, h1); L.yak_ch_hist(hg, h2, 4)
    assert list(h1) == list(h2)
    # count-existing mode then clear
    a, p = util.u64_array(ev[:5000])
    OL.yo_ch_insert_list(ho, 0, len(a), p); L.yak_ch_insert_list(hg, 0, len(a), p)
    assert yakb.dump_bytes(hg) == oracle.dump_bytes(ho)
    OL.yo_ch_clear(ho); L.yak_ch_clear(hg, 4)
    assert yakb.dump_bytes(hg) == oracle.dump_bytes(ho)
    L.yak_ch_destroy(hg); OL.yo_ch_destroy(ho)


def test_trailing_put_and_shrink_quirks(oracle, yakb):
    """SURVEY section 4 KATs: Q3 trailing-put doubling {4,3}->{8,3}; Q9 shrink of empty sub-tables."""
    L = yakb.lib()
    pre = 10
    keys = np.array([(i * 7919 + 3) << pre | 5 for i in range(1, 4)], dtype=np.uint64)
    import struct
    for extra, want in ((False, (4, 3)), (True, (8, 3))):
        h = L.yak_ch_init(31, pre, 4, 0)
        lst = np.concatenate([keys, keys[:1]]) if extra else keys
        a, p = util.u64_array(lst)
        L.yak_ch_insert_list(h, 1, len(a), p)
        data = yakb.dump_bytes(h)
        off = 16
        for s in range(5):
            cap, size = struct.unpack_from("<II", data, off); off += 8 + 8 * size
        assert struct.unpack_from("<II", data, off) == want
        L.yak_ch_shrink(h, 1, 1023, 1)                        # Q9: empty sub-tables now have capacity 4
        data = yakb.dump_bytes(h)
        assert struct.unpack_from("<II", data, 16) == (4, 0)
        L.yak_ch_destroy(h)


def test_restore_roundtrip_and_qv(oracle, yakb, reads_fa):
    OL, L = oracle.lib(), yakb.lib()
    k, pre, b = 31, 12, 22
    ho, _ = oracle.count_file(reads_fa, k=k, pre=pre, bf_shift=b)
    fn = os.path.join(util.TMP, "yakb_sr.yak")
    assert OL.yo_ch_dump(ho, fn.encode()) == 0
    hg = L.yak_ch_restore(fn.encode())
    assert hg and hg.contents.k == k and hg.contents.pre == pre
    ho2 = OL.yo_ch_restore(fn.encode())
    assert yakb.dump_bytes(hg) == oracle.dump_bytes(ho2)      # restore -> dump goes through khashl re-placement
    # qv scan of contigs cut from the genome
    from yak_b200 import synth
    ctg = synth.contigs_bytes(7, 200_000, 3, 6, 20_000, sub=1e-3)
    seqs = [ln for ln in ctg.split(b"\n") if ln and not ln.startswith(b">")] + [b"ACGTNNNN", b"ACGT" * 20]
    lens = (C.c_int64 * len(seqs))(*[len(s) for s in seqs])
    cat = b"".join(seqs)
    for min_len, min_frac in ((0, 0.5), (100, 0.999)):
        c1, c2 = (C.c_int64 * 1024)(), (C.c_int64 * 1024)()
        t1, z1 = (C.c_int32 * len(seqs))(), (C.c_int32 * len(seqs))()
        t2, z2 = (C.c_int32 * len(seqs))(), (C.c_int32 * len(seqs))()
        OL.yo_qv_seqs(ho2, len(seqs), lens, cat, min_len, min_frac, c1, t1, z1)
        assert L.yakb_qv_seqs(hg, len(seqs), lens, cat, min_len, min_frac, c2, t2, z2) == 0
        assert list(t1) == list(t2) and list(z1) == list(z2)
        assert list(c1) == list(c2)
        assert sum(c1) > 0 or min_frac > 0.9
    # yak_qv(file) without per-sequence output reads plain files through the parser pool, gzip through the sequential reader
    import gzip
    fa = os.path.join(util.TMP, "yakb_qv_ctg.fa")
    with open(fa, "wb") as f:
        f.write(synth.contigs_bytes(7, 200_000, 5, 40, 9_000, sub=2e-3, width=60) + b">short\nACGTACGT\n>n\n" + b"N" * 200 + b"\n")
    with gzip.open(fa + ".gz", "wb") as f:
        f.write(open(fa, "rb").read())
    recs = [b"".join(r.split(b"\n")[1:]) for r in open(fa, "rb").read().split(b">")[1:]]
    lens = (C.c_int64 * len(recs))(*[len(s) for s in recs])
    for min_len in (0, 5000):
        qo = yakb.YakQopt()
        L.yak_qopt_init(C.byref(qo))
        qo.min_len = min_len
        want = (C.c_int64 * 1024)()
        OL.yo_qv_seqs(ho2, len(recs), lens, b"".join(recs), min_len, qo.min_frac, want, None, None)
        for path in (fa, fa + ".gz"):
            got = (C.c_int64 * 1024)()
            L.yak_qv(C.byref(qo), path.encode(), hg, got)
            assert list(got) == list(want) and sum(want) > 100_000, (path, min_len)
    L.yak_ch_destroy(hg); OL.yo_ch_destroy(ho); OL.yo_ch_destroy(ho2)


@pytest.mark.parametrize("smem_max,warp", [("0", "1"), ("0", "0"), ("1024", "1")])
def test_layout_replay_variants_agree(oracle, yakb, reads_fa, monkeypatch, smem_max, warp):
    """the khashl layout rebuild has three code paths (shared memory, one thread, one warp per sub-table);
    force the large-table paths on small data: count, two-pass bloom, restore and shrink must stay byte-exact"""
    monkeypatch.setenv("YAKB_LAYOUT_SMEM_MAX", smem_max)
    monkeypatch.setenv("YAKB_LAYOUT_WARP", warp)
    for k, pre, b in ((31, 10, 0), (31, 10, 20), (21, 11, 0)):
        ho, hg, ref, mine, _ = _count_both(oracle, yakb, reads_fa, k, pre, b)
        try:
            assert mine == ref, util.explain_diff(mine, ref)
            oracle.lib().yo_ch_shrink(ho, 3, 900)
            yakb.lib().yak_ch_shrink(hg, 3, 900, 1)
            a, bb = yakb.dump_bytes(hg), oracle.dump_bytes(ho)
            assert a == bb, util.explain_diff(a, bb)
        finally:
            yakb.lib().yak_ch_destroy(hg)
            oracle.lib().yo_ch_destroy(ho)


ZONE_SNIPPET = r"""
import sys, os
import ctypes as C
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np
import oracle_lib
from yak_b200 import capi, synth
capi.require_gpu()
fn = {fn!r}
rng = np.random.default_rng(11)
with open(fn, "wb") as f:
    f.write(synth.reads_file_bytes(3, 300_000, 4, 30_000, 150, 0.01, 2))
    for i in range(3000):                      # skew: one k-mer (and its neighbours) thousands of times -> zone lists overflow
        f.write(b">pa%d\n" % i + b"A" * int(rng.integers(31, 200)) + b"\n")
    f.write(synth.reads_file_bytes(3, 300_000, 5, 20_000, 150, 0.01, 2))
for k, pre, b in ((31, 10, 0), (31, 12, 24), (47, 11, 23), (21, 10, 22)):
    hg = capi.count_file(fn, k=k, pre=pre, bf_shift=b, chunk_size=1_500_000)
    ho, _ = oracle_lib.count_file(fn, k=k, pre=pre, bf_shift=b)
    assert capi.dump_bytes(hg) == oracle_lib.dump_bytes(ho), (k, pre, b)
    capi.lib().yak_ch_destroy(hg); oracle_lib.lib().yo_ch_destroy(ho)
# the array front end (yak_ch_insert_list, count.c:82; the multi-GPU receive side) through the same partition
OL, L = oracle_lib.lib(), capi.lib()
k, pre = 31, 10
seqs = [ln.strip() for ln in open(fn, "rb") if not ln.startswith(b">")][:4000]
ev = []
for s in seqs:
    buf = (C.c_uint64 * len(s))()
    n = OL.yo_extract(k, len(s), s, buf)
    ev.append(np.frombuffer(buf, dtype=np.uint64, count=n).copy())
ev = np.concatenate(ev)
for b in (0, 21):
    ho, hg = OL.yo_ch_init(k, pre, 4, b), L.yak_ch_init(k, pre, 4, b)
    for part in np.array_split(ev, 3):
        sub = (part & np.uint64((1 << pre) - 1)).astype(np.int64)
        order = np.argsort(sub, kind="stable")
        part_s, sub_s = part[order], sub[order]
        for lst in np.split(part_s, np.flatnonzero(np.diff(sub_s)) + 1):
            a = np.ascontiguousarray(lst, dtype=np.uint64)
            p = a.ctypes.data_as(C.POINTER(C.c_uint64))
            assert OL.yo_ch_insert_list(ho, 1, len(a), p) == L.yak_ch_insert_list(hg, 1, len(a), p)
    a = np.ascontiguousarray(ev[:50_000], dtype=np.uint64)         # a long list with foreign elements (htab.c:61)
    p = a.ctypes.data_as(C.POINTER(C.c_uint64))
    assert OL.yo_ch_insert_list(ho, 1, len(a), p) == L.yak_ch_insert_list(hg, 1, len(a), p)
    assert capi.dump_bytes(hg) == oracle_lib.dump_bytes(ho), ("insert_list", b)
    L.yak_ch_destroy(hg); OL.yo_ch_destroy(ho)
print("zone ok")
"""


@pytest.mark.parametrize("env", [
    {"YAKB_ZONE_MB": "0.02"},                                  # several sub-tables per zone
    {"YAKB_ZONE_MB": "0.0001"},                                # one sub-table per zone (2048 zones at -p12: the maximum)
    {"YAKB_ZONE_MB": "0.02", "YAKB_ZONE_SLACK": "0"},          # zone lists overflow into the spill list, the spill list into the fallback
    {"YAKB_ZONE_MB": "4"},                                     # few zones
    {"YAKB_ZONE_MB": "0.02", "YAKB_PEND_MAX": "20000"},        # the pending events of a chunk in many ranges
    {"YAKB_ZONE": "0", "YAKB_PEND_MAX": "5000"},               # ranges behind the unpartitioned probe
], ids=["zones", "zone-per-subtable", "spill", "few-zones", "pending-ranges", "ranges-unpartitioned"])
def test_partitioned_front_end_is_bit_exact(env):
    """the partitioned probe path (partition.cuh: part_scatter, engine.cu: zone_probe) is chosen for tables of gigabytes and
    chunks of tens of millions of positions; force it on small inputs - both passes, several chunks, skewed k-mers, packed
    reads and event arrays - and compare the .yak bytes with the oracle"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ, YAKB_BATCH="1500000", YAKB_ZONE="1")
    e.update(env)
    r = subprocess.run([sys.executable, "-c", ZONE_SNIPPET.format(root=root, fn=os.path.join(util.TMP, "yakb_zone.fa"))],
                       env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "zone ok" in r.stdout, r.stderr[-3000:]


def test_ch_inc_matches_oracle(oracle, yakb, reads_fa):
    """yak_ch_inc (htab.c:80-91): saturating ++ of an existing key, -1 for an absent one; then the bytes"""
    OL, L = oracle.lib(), yakb.lib()
    OL.yo_ch_inc.restype = C.c_int
    OL.yo_ch_inc.argtypes = [C.POINTER(oracle.YoCh), C.c_uint64]
    k, pre = 31, 10
    ho, _ = oracle.count_file(reads_fa, k=k, pre=pre, bf_shift=0)
    hg = yakb.count_file(reads_fa, k=k, pre=pre, bf_shift=0)
    seqs = [ln.strip() for ln in open(reads_fa) if not ln.startswith(">")][:40]
    ev = []
    for s in seqs:
        buf = (C.c_uint64 * len(s))()
        n = OL.yo_extract(k, len(s), s.encode(), buf)
        ev.append(np.frombuffer(buf, dtype=np.uint64, count=n).copy())
    ev = np.concatenate(ev)[:600]
    absent = ev[:50] ^ np.uint64(0x2AAAAAAAAAAA)
    for x in np.concatenate([ev, absent, ev[:20]]):
        assert L.yak_ch_inc(hg, int(x)) == OL.yo_ch_inc(ho, int(x))
    # a counter at the cap stays there (htab.c:88): polyA-like saturation by repeated increments of one key
    x0 = int(ev[0])
    for _ in range(1030):
        OL.yo_ch_inc(ho, x0)
    OL.yo_ch_inc.restype = C.c_int
    a, p = util.u64_array(np.full(1030, x0, dtype=np.uint64))
    L.yak_ch_insert_list(hg, 0, len(a), p)               # the batched form of the same increments
    assert L.yak_ch_inc(hg, x0) == OL.yo_ch_inc(ho, x0) == 1023
    assert yakb.dump_bytes(hg) == oracle.dump_bytes(ho)
    L.yak_ch_destroy(hg); OL.yo_ch_destroy(ho)


@pytest.mark.parametrize("n_shift,n_hashes", [(9, 4), (12, 4), (20, 2), (16, 12), (24, 40)])
def test_public_bloom_filter_matches_oracle(oracle, yakb, n_shift, n_hashes):
    """yak_bf_init / yak_bf_insert / yak_bf_destroy (bbf.c:5-42) as a stand-alone device filter: the same return value
    (number of bits already set) for every insert of a stream with repeats; init refuses n_shift < 9 and > 55"""
    OL, L = oracle.lib(), yakb.lib()
    assert not L.yak_bf_init(8, 4) and not L.yak_bf_init(56, 4)          # bbf.c:9
    bo, bg = OL.yo_bloom_init(n_shift, n_hashes), L.yak_bf_init(n_shift, n_hashes)
    assert bo and bg
    rng = np.random.default_rng(n_shift * 100 + n_hashes)
    xs = rng.integers(0, 1 << 62, 400, dtype=np.uint64)
    xs = np.concatenate([xs, xs[::3], (xs[:64] & np.uint64((1 << n_shift) - 1)) | np.uint64(32 << n_shift)])  # h2 a multiple of 32: bbf.c:33
    for x in xs:
        assert L.yak_bf_insert(bg, int(x)) == OL.yo_bloom_insert(bo, int(x))
    L.yak_bf_destroy(bg); OL.yo_bloom_destroy(bo)
    L.yak_bf_destroy(None)                                                 # bbf.c:22: NULL-safe


@pytest.mark.parametrize("k", [32, 63])
def test_the_key_that_looks_like_an_empty_slot(oracle, yakb, k):
    """k >= 32 with -p10: the stored key whose 54 id bits are all ones, at count 1023, is the bit pattern the device table uses for an
    empty slot (csrc/yakb_dev.cuh YAKB_SAT_BYTES).  Every way a counter is written or read, on exactly that key: counting up to and past
    the cap, lookups, histogram, increments one by one, dump, restore of a file that holds the pattern, shrink, clear - against the oracle
    (khashl has a used-bit per slot, khashl.h:92-96, and no such special value)."""
    OL, L = oracle.lib(), yakb.lib()
    OL.yo_ch_inc.restype = C.c_int
    OL.yo_ch_inc.argtypes = [C.POINTER(oracle.YoCh), C.c_uint64]
    pre = 10
    star = np.uint64(0xFFFFFFFFFFFFFFFF)                         # id bits all ones, sub-table 1023
    near = [np.uint64(0xFFFFFFFFFFFFFFFF ^ (1 << b)) for b in (10, 11, 40, 63)] + [np.uint64(0x7FFFFFFFFFFFFBFF)]   # neighbours in the same and other sub-tables
    rng = np.random.default_rng(k)
    filler = rng.integers(0, 1 << 63, 3000, dtype=np.uint64) | np.uint64(1023)   # more keys of sub-table 1023: the table grows around it
    ho, hg = OL.yo_ch_init(k, pre, 4, 0), L.yak_ch_init(k, pre, 4, 0)

    def feed(ev, create_new=1):
        ev = np.ascontiguousarray(ev, dtype=np.uint64)
        sub = (ev & np.uint64((1 << pre) - 1)).astype(np.int64)
        order = np.argsort(sub, kind="stable")
        evs, subs = ev[order], sub[order]
        for lst in np.split(evs, np.flatnonzero(np.diff(subs)) + 1):
            a, p = util.u64_array(lst)
            assert OL.yo_ch_insert_list(ho, create_new, len(a), p) == L.yak_ch_insert_list(hg, create_new, len(a), p)

    def same(tag):
        a, b = yakb.dump_bytes(hg), oracle.dump_bytes(ho)
        assert a == b, (tag, util.explain_diff(a, b))
        h1, h2 = (C.c_int64 * 1024)(), (C.c_int64 * 1024)()
        OL.yo_ch_hist(ho, h1); L.yak_ch_hist(hg, h2, 1)
        assert list(h1) == list(h2), tag
        for x in [star] + near:
            assert L.yak_ch_get(hg, int(x)) == OL.yo_ch_get(ho, int(x)), (tag, hex(int(x)))

    feed(np.concatenate([np.full(1020, star), near, filler[:1000], np.full(700, near[0])]))
    same("1020")
    feed(np.concatenate([np.full(2, star), filler[1000:2000]]))           # 1022: the last value a slot holds
    same("1022")
    feed(np.full(1, star)); same("1023")                                  # the step that would write the empty pattern
    feed(np.concatenate([np.full(50, star), filler[2000:]])); same("past the cap, table grown")
    assert L.yak_ch_inc(hg, int(star)) == OL.yo_ch_inc(ho, int(star)) == 1023
    fn = os.path.join(util.TMP, f"yakb_star_{k}.yak")
    assert OL.yo_ch_dump(ho, fn.encode()) == 0                            # the file holds the pattern as a key
    hr = L.yak_ch_restore(fn.encode())
    ho2 = OL.yo_ch_restore(fn.encode())
    assert yakb.dump_bytes(hr) == oracle.dump_bytes(ho2)
    assert L.yak_ch_get(hr, int(star)) == 1023
    L.yak_ch_destroy(hr); OL.yo_ch_destroy(ho2)
    OL.yo_ch_shrink(ho, 1000, 1023); L.yak_ch_shrink(hg, 1000, 1023, 1); same("shrink keeps it")
    OL.yo_ch_clear(ho); L.yak_ch_clear(hg, 1); same("clear")
    feed(np.full(1022, star), create_new=0); same("count-only pass to 1022")
    for want in (1023, 1023):                                             # one by one over the edge (htab.c:80-91)
        assert L.yak_ch_inc(hg, int(star)) == OL.yo_ch_inc(ho, int(star)) == want
    same("inc over the edge")
    L.yak_ch_destroy(hg); OL.yo_ch_destroy(ho)
