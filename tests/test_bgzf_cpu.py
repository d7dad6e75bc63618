"""Blocked-gzip input (yak_b200/csrc/bgzf.h): the pool of inflating threads must hand the parser exactly the byte
stream zlib's gzread produces (which is what the reference parses: kseq over gzread, count.c:150-151) - on well-formed
BGZF files, on files that continue with plain gzip members, on truncated and on corrupted files."""
import ctypes as C
import gzip
import os
import struct
import zlib

import numpy as np
import pytest

import util
from yak_b200 import capi


def bgzf_block(data: bytes, level=6, extra_before=b"", flg_name=b"") -> bytes:
    """one BGZF member; extra_before = other subfields in front of 'BC', flg_name = an FNAME field (both legal gzip)"""
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    body = co.compress(data) + co.flush()
    xlen = len(extra_before) + 6
    flg = 4 | (8 if flg_name else 0)
    name = flg_name + b"\0" if flg_name else b""
    total = 12 + xlen + len(name) + len(body) + 8
    assert total <= 65536
    head = struct.pack("<BBBBIBBH", 31, 139, 8, flg, 0, 0, 255, xlen) + extra_before + b"BC" + struct.pack("<HH", 2, total - 1)
    return head + name + body + struct.pack("<II", zlib.crc32(data), len(data) & 0xFFFFFFFF)


EOF_MARK = bgzf_block(b"")


def bgzf_bytes(data: bytes, block=65280, rng=None, **kw) -> bytes:
    out, i = [], 0
    while i < len(data):
        n = block if rng is None else int(rng.integers(1, block + 1))
        out.append(bgzf_block(data[i:i + n], **kw))
        i += n
    return b"".join(out) + EOF_MARK


def fastq_text(rng, n_rec, fasta=False) -> bytes:
    out = []
    for i in range(n_rec):
        ln = int(rng.integers(1, 400))
        s = bytes(rng.choice(np.frombuffer(b"ACGTNacgt", dtype=np.uint8), ln))
        if fasta:
            w = int(rng.integers(20, 90))
            out.append(b">s%d some comment\n" % i + b"\n".join(s[j:j + w] for j in range(0, ln, w)) + b"\n")
        else:
            out.append(b"@r%d\n" % i + s + b"\n+\n" + b"I" * ln + b"\n")
    return b"".join(out)


def lib():
    L = capi.lib()
    L.yakb_fastx_open_bgzf.restype = C.c_void_p
    L.yakb_fastx_open_bgzf.argtypes = [C.c_char_p, C.c_int, C.c_uint64]
    L.yakb_fastx_bgzf_threads.argtypes = [C.c_void_p]
    L.yakb_fastx_fill.restype = C.c_int64
    L.yakb_fastx_fill.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
    L.yakb_fastx_close.argtypes = [C.c_void_p]
    L.yakb_fastx_next.restype = C.c_int64
    L.yakb_fastx_next.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]
    return L


def read_all(fn, threads, job=4 << 20, cap=1 << 20, expect_pool=None):
    """(bytes of every record as SEQ\\n through the bulk path, number of records)"""
    L = lib()
    r = L.yakb_fastx_open_bgzf(fn.encode(), threads, job)
    assert r
    if expect_pool is not None:
        assert (L.yakb_fastx_bgzf_threads(r) > 0) == expect_pool
    out, nseq = bytearray(), 0
    buf = C.create_string_buffer(cap)
    while True:
        ns, done, need = C.c_int64(), C.c_int(), C.c_uint64()
        n = L.yakb_fastx_fill(r, buf, cap, cap, 0, C.byref(ns), C.byref(done), C.byref(need))
        assert not need.value
        out += buf.raw[:n]
        nseq += ns.value
        if done.value:
            break
    L.yakb_fastx_close(r)
    return bytes(out), nseq


def records(fn, threads, job=4 << 20):
    L = lib()
    r = L.yakb_fastx_open_bgzf(fn.encode(), threads, job)
    out = []
    seq, name = C.c_char_p(), C.c_char_p()
    while True:
        n = L.yakb_fastx_next(r, C.byref(seq), C.byref(name))
        if n < 0:
            out.append(n)
            if n == -1:
                break
            continue
        out.append((name.value, seq.value))
    L.yakb_fastx_close(r)
    return out


def _write(name, data):
    p = os.path.join(util.TMP, name)
    with open(p, "wb") as f:
        f.write(data)
    return p


@pytest.mark.parametrize("fasta", [False, True])
def test_bgzf_pool_equals_plain_and_zlib_readers(fasta):
    rng = np.random.default_rng(5 + fasta)
    text = fastq_text(rng, 6000, fasta)
    plain = _write("yakb_bgzf_plain.fx", text)
    want = read_all(plain, 0, expect_pool=False)
    assert want[1] == 6000
    for block, threads, job in ((65280, 0, 4 << 20), (65280, 3, 100_000), (4000, 8, 1), (100, 2, 5000), (65280, 1, 1 << 20)):
        fn = _write("yakb_bgzf_a.fx.gz", bgzf_bytes(text, block))
        assert gzip.decompress(open(fn, "rb").read()) == text          # the writer above makes valid gzip
        assert read_all(fn, threads, job, expect_pool=True) == want
        assert read_all(fn, -1, expect_pool=False) == want              # zlib's reader on the same file
    fn = _write("yakb_bgzf_a.fx.gz", bgzf_bytes(text, 30000, rng))          # ragged blocks
    assert read_all(fn, 4, 70_000, cap=3000) == want
    assert records(fn, 4, 70_000) == records(plain, 0)
    # legal variations of the member header: another subfield in front of 'BC', a file name field
    fn = _write("yakb_bgzf_a.fx.gz", bgzf_bytes(text, 50000, extra_before=b"XY" + struct.pack("<H", 3) + b"abc", flg_name=b"reads.fq"))
    assert gzip.decompress(open(fn, "rb").read()) == text
    assert read_all(fn, 4, expect_pool=True) == want
    # an ordinary gzip file never goes through the pool
    fn = _write("yakb_bgzf_b.fx.gz", gzip.compress(text))
    assert read_all(fn, 4, expect_pool=False) == want
    # empty members anywhere, nothing but empty members
    fn = _write("yakb_bgzf_a.fx.gz", EOF_MARK * 3 + bgzf_bytes(text[:100_000], 7000) + EOF_MARK * 5000 + bgzf_bytes(text[100_000:], 9000))
    assert read_all(fn, 4, 20_000, expect_pool=True) == want
    fn = _write("yakb_bgzf_a.fx.gz", EOF_MARK * 10)
    assert read_all(fn, 2, expect_pool=True) == (b"", 0)


def test_bgzf_followed_by_other_members_or_garbage():
    rng = np.random.default_rng(11)
    a, b, c = fastq_text(rng, 1500), fastq_text(rng, 1500), fastq_text(rng, 700)
    # BGZF blocks, then a plain gzip member, then BGZF blocks again: zlib reads all three; so must we
    data = bgzf_bytes(a, 20000) + gzip.compress(b) + bgzf_bytes(c, 20000)
    fn = _write("yakb_bgzf_mix.fq.gz", data)
    assert gzip.decompress(data) == a + b + c
    want = read_all(fn, -1)
    assert want[1] == 3700
    for threads, job in ((4, 50_000), (1, 1 << 20), (8, 1)):
        assert read_all(fn, threads, job, expect_pool=True) == want
    # whatever is not a gzip member after the last one is ignored (zlib gz_look)
    for junk in (b"\0" * 100, b"this is not gzip", b"\x1f", b"\x1f\x8b"):
        fn = _write("yakb_bgzf_junk.fq.gz", bgzf_bytes(a, 20000) + junk)
        want = read_all(fn, -1)
        assert want[1] == 1500
        assert read_all(fn, 4, 50_000, expect_pool=True) == want


def zlib_stream(data: bytes) -> bytes:
    """what a member-by-member inflate of the file yields: every byte that decodes before an error or the end of the data;
    members are read as long as the next bytes are a gzip magic (zlib gz_look), anything else is ignored"""
    out, pos = [], 0
    while len(data) - pos >= 2 and data[pos:pos + 2] == b"\x1f\x8b":
        d = zlib.decompressobj(31)
        try:
            o = d.decompress(data[pos:])
        except zlib.error:       # Python drops the partial output of the failing call: feed this member byte by byte
            d = zlib.decompressobj(31)
            for i in range(pos, len(data)):
                try:
                    out.append(d.decompress(data[i:i + 1]))
                except zlib.error:
                    break
            break
        out.append(o)
        if not d.eof:            # cut short
            break
        pos = len(data) - len(d.unused_data)
    return b"".join(out)


@pytest.mark.parametrize("seed", range(6))
def test_bgzf_truncated_and_corrupted_files(seed):
    """the pool on damaged files: exactly the bytes that inflate before the damage (zlib's gzread delivers the same stream
    minus the output of the read call that met the error), and it reads on where only the BGZF fields are wrong"""
    rng = np.random.default_rng(100 + seed)
    text = fastq_text(rng, 1200)
    good = bgzf_bytes(text, 8000, rng)
    assert zlib_stream(good) == text
    for trial in range(24):
        data = bytearray(good)
        kind = trial % 4
        if kind == 0:      # cut anywhere
            data = data[:int(rng.integers(1, len(data)))]
        elif kind == 1:    # flip one byte anywhere (header, BSIZE, deflate data, CRC, ISIZE)
            i = int(rng.integers(28, len(data)))
            data[i] ^= int(rng.integers(1, 256))
        elif kind == 2:    # both
            i = int(rng.integers(28, len(data)))
            data[i] ^= int(rng.integers(1, 256))
            data = data[:int(rng.integers(i, len(data)))]
        else:              # a BSIZE field that lies in some member behind the first: gzread never looks at it
            offs, o = [], 0
            while o < len(data):
                offs.append(o)
                o += struct.unpack_from("<H", data, o + 16)[0] + 1
            o = offs[int(rng.integers(1, len(offs)))]
            struct.pack_into("<H", data, o + 16, int(rng.integers(0, 65536)))
        fn = _write("yakb_bgzf_bad.fq.gz", bytes(data))
        stream = zlib_stream(bytes(data))
        if kind == 3:
            assert stream == text
        want = read_all(_write("yakb_bgzf_bad_expect.fq", stream), 0)
        got = read_all(fn, int(rng.integers(1, 6)), int(rng.choice([1, 30_000, 1 << 20])))
        assert got == want, (seed, trial, kind)
        if kind in (0, 3):   # nothing is dropped by gzread either when the file is merely cut short / mislabelled
            assert read_all(fn, -1) == want


@pytest.mark.parametrize("seed", list(range(12)) + [27, 53, 73])
def test_sequential_reader_buffer_boundaries_on_adversarial_input(seed):
    """the sequential reader's block buffer is 4 MB, so small test files never put a record across two buffers.  BGZF members
    of a few bytes with one member per job do: every member becomes one parser buffer.  Adversarial FASTA/FASTQ (CRLF, blank
    lines, multi-line, truncated qualities) read that way must equal the same text read from a plain file: the bulk path
    (shortcut included), the record path, and what follows truncated records for several -K"""
    import test_oracle_cpu as T
    rng = np.random.default_rng(4000 + seed)
    text = T._random_fastx(rng, int(rng.integers(20, 150)), False, bad=(0.0, 0.15)[seed % 2])
    plain = _write(f"yakb_bgzf_adv{seed}.fx", text)          # per-seed names: the cases may run side by side (pytest -n)
    fn = _write(f"yakb_bgzf_adv{seed}.fx.gz", bgzf_bytes(text, int(rng.choice([1, 2, 7, 33, 150, 1000])), rng if seed % 3 else None))
    L = lib()
    L.yakb_fastx_set_chunk.argtypes = [C.c_void_p, C.c_int64]

    def fill_all(path, threads, job, cap, target, min_len, chunk):
        r = L.yakb_fastx_open_bgzf(path.encode(), threads, job)
        L.yakb_fastx_set_chunk(r, chunk)
        out, nseq = bytearray(), 0
        buf = C.create_string_buffer(cap)
        while True:
            ns, done, need = C.c_int64(), C.c_int(), C.c_uint64()
            n = L.yakb_fastx_fill(r, buf, cap, target, min_len, C.byref(ns), C.byref(done), C.byref(need))
            if need.value:
                cap = need.value + 5
                buf = C.create_string_buffer(cap)
                continue
            out += buf.raw[:n]
            nseq += ns.value
            if done.value:
                break
        L.yakb_fastx_close(r)
        return bytes(out), nseq
    for min_len in (0, 31):
        for chunk in (10_000_000, 3000, 200):
            want = fill_all(plain, 0, 1, 1 << 22, 1 << 22, min_len, chunk)
            assert want[0] == b"".join(s + b"\n" for s in T._ref_flow(plain, min_len, chunk))     # the plain read is the pinned one
            for cap, target in ((1 << 22, 1 << 22), (700, 300), (64, 64)):
                assert fill_all(fn, int(rng.integers(1, 5)), 1, cap, target, min_len, chunk) == want, (seed, min_len, chunk, cap)
    assert records(fn, 3, 1) == records(plain, 0)


def test_readers_under_thread_sanitizer():
    """tools/tsan_readers.cpp: the sequential reader with its read-ahead thread, the BGZF pool (also closed while busy) and the
    parser pool compiled with -fsanitize=thread - no data race reported, same bytes on every path"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(util.TMP, "yakb_tsan")
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(3)
    text = fastq_text(rng, 8000)
    for name, data in (("t.fq", text), ("t.fq.gz", bgzf_bytes(text, 3000, rng)),
                       ("t_mix.fq.gz", bgzf_bytes(text[:len(text) // 2], 3000, rng) + gzip.compress(text[len(text) // 2:]))):
        open(os.path.join(d, name), "wb").write(data)
    bad = bytearray(bgzf_bytes(text, 3000, rng))
    bad[len(bad) // 2] ^= 0x55
    open(os.path.join(d, "t_bad.fq.gz"), "wb").write(bytes(bad))
    csrc = os.path.join(root, "yak_b200", "csrc")
    exe = os.path.join(d, "tsan_readers")
    cc = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-pthread", "-fsanitize=thread", "-I" + csrc, os.path.join(root, "tools", "tsan_readers.cpp")] +
                        [os.path.join(csrc, f) for f in ("fastx.cpp", "bgzf.cpp", "fastx_par.cpp")] + ["-lz", "-o", exe], capture_output=True, text=True)
    if cc.returncode != 0:
        pytest.skip("no ThreadSanitizer runtime here: " + cc.stderr[-300:])
    r = subprocess.run([exe, d], capture_output=True, text=True, env=dict(os.environ, TSAN_OPTIONS="halt_on_error=1"))
    assert r.returncode == 0 and r.stdout.startswith("ok") and "ThreadSanitizer" not in r.stderr, (r.stdout + r.stderr)[-3000:]


def test_inputs_of_many_parser_buffers():
    """20 MB of FASTQ (five 4 MB parser buffers; records across every buffer end): plain file, plain gzip through the read-ahead
    thread, BGZF through the pool, and the parser pool - one stream"""
    from yak_b200 import synth
    text = synth.reads_file_bytes(1, 500_000, 2, 64_000, fastq=True)
    assert len(text) > (18 << 20)
    plain = _write("yakb_big.fq", text)
    want = read_all(plain, 0, cap=8 << 20)
    assert want[1] == 64_000
    gz = _write("yakb_big.fq.gz", gzip.compress(text, 1))
    assert read_all(gz, 0, cap=8 << 20, expect_pool=False) == want
    bz = _write("yakb_big.bgzf.fq.gz", bgzf_bytes(text, 65280, level=1))
    assert read_all(bz, 4, cap=8 << 20, expect_pool=True) == want
    assert read_all(bz, 3, job=300_000, cap=3 << 20) == want
    L = capi.lib()
    r = L.yakb_pfastx_open(plain.encode(), 1 << 20, 4)
    out, nseq = bytearray(), 0
    buf = C.create_string_buffer(8 << 20)
    while True:
        ns, done, need = C.c_int64(), C.c_int(), C.c_uint64()
        n = L.yakb_pfastx_fill(r, buf, 8 << 20, 8 << 20, 0, C.byref(ns), C.byref(done), C.byref(need))
        out += buf.raw[:n]
        nseq += ns.value
        if done.value:
            break
    L.yakb_pfastx_close(r)
    assert (bytes(out), nseq) == want
    for p in (plain, gz, bz):
        os.unlink(p)


def test_text_cache_for_the_second_pass(monkeypatch):
    """csrc/textcache.h: the first pass over a compressed file keeps the inflated text in a memfd, the second pass maps it and
    parses it with the parser pool - same records as reading the file again; nothing is kept when the text exceeds the budget,
    when the cache is off, or for a file that changed in between"""
    import time
    L = lib()
    L.yakb_fastx_open_tee.restype = C.c_void_p
    L.yakb_fastx_open_tee.argtypes = [C.c_char_p]
    L.yakb_fastx_tee_commit.argtypes = [C.c_void_p]
    L.yakb_text_cache_path.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    L.yakb_pfastx_open.restype = C.c_void_p
    L.yakb_pfastx_open.argtypes = [C.c_char_p, C.c_uint64, C.c_int]
    L.yakb_pfastx_fill.restype = C.c_int64
    L.yakb_pfastx_fill.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
    L.yakb_pfastx_close.argtypes = [C.c_void_p]

    def drain(r, fill, cap=1 << 20):
        out, nseq = bytearray(), 0
        buf = C.create_string_buffer(cap)
        while True:
            ns, done, need = C.c_int64(), C.c_int(), C.c_uint64()
            n = fill(r, buf, cap, cap, 0, C.byref(ns), C.byref(done), C.byref(need))
            assert not need.value
            out += buf.raw[:n]
            nseq += ns.value
            if done.value:
                return bytes(out), nseq

    def first_pass(fn):
        r = L.yakb_fastx_open_tee(fn.encode())
        assert r
        got = drain(r, L.yakb_fastx_fill)
        kept = L.yakb_fastx_tee_commit(r)
        L.yakb_fastx_close(r)
        return got, kept

    def cache_path(fn):
        buf = C.create_string_buffer(256)
        assert L.yakb_text_cache_path(fn.encode(), buf, 256) >= 0
        return buf.value.decode()

    rng = np.random.default_rng(21)
    text = fastq_text(rng, 9000) + fastq_text(rng, 500, fasta=True)
    plain = _write("yakb_tc.fx", text)
    want = read_all(plain, 0)
    gz = _write("yakb_tc.fx.gz", gzip.compress(text))
    bz = _write("yakb_tc.bgzf.gz", bgzf_bytes(text, 20000, rng))
    monkeypatch.delenv("YAKB_TEXT_CACHE_GB", raising=False)
    assert first_pass(gz) == (want, 0) and cache_path(gz) == ""                    # off by default
    monkeypatch.setenv("YAKB_TEXT_CACHE_GB", "0.5")
    for fn in (gz, bz):
        got, kept = first_pass(fn)
        assert got == want and kept == 1
        p = cache_path(fn)
        assert p.startswith("/proc/self/fd/") and cache_path(plain) == ""
        assert open(p, "rb").read() == text                                      # the inflated text, byte for byte
        r = L.yakb_pfastx_open(p.encode(), 50_000, 3)                              # what the second pass does
        assert r
        L.yakb_text_cache_release()
        assert cache_path(fn) == ""                                               # one use
        assert drain(r, L.yakb_pfastx_fill) == want
        L.yakb_pfastx_close(r)
    monkeypatch.setenv("YAKB_TEXT_CACHE_GB", "0.000001")                         # 1 KB: the text does not fit
    assert first_pass(gz) == (want, 0) and cache_path(gz) == ""
    monkeypatch.setenv("YAKB_TEXT_CACHE_GB", "0.5")
    assert first_pass(gz)[1] == 1 and cache_path(gz) != ""
    time.sleep(0.01)
    with open(gz, "wb") as f:                                                     # the file changed: the copy is not its text any more
        f.write(gzip.compress(text[:100_000]))
    assert cache_path(gz) == ""
    L.yakb_text_cache_release()
