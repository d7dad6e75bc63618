"""ctypes binding of oracle/liboracle.so (the CPU restatement) - test infrastructure only."""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_YAK = os.path.join(ORACLE_DIR, "_ref", "yak")
REF_LIB = os.path.join(ORACLE_DIR, "_ref", "libyakref.so")


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)


class YoCh(C.Structure):
    _fields_ = [("k", C.c_int), ("pre", C.c_int), ("n_hash", C.c_int), ("n_shift", C.c_int),
                ("tot", C.c_uint64), ("h", C.c_void_p), ("b", C.c_void_p)]


class YoSet(C.Structure):
    _fields_ = [("bits", C.c_uint32), ("count", C.c_uint32), ("used", C.c_void_p), ("keys", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(ORACLE_DIR, "liboracle.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    u64, i64, p = C.c_uint64, C.c_int64, C.c_void_p
    L.yo_hash64.restype = u64; L.yo_hash64.argtypes = [u64, u64]
    L.yo_hash64_64.restype = u64; L.yo_hash64_64.argtypes = [u64]
    L.yo_hash_long.restype = u64; L.yo_hash_long.argtypes = [C.POINTER(u64)]
    L.yo_hash64_inv.restype = u64; L.yo_hash64_inv.argtypes = [u64, u64]
    L.yo_slot_home.restype = C.c_uint32; L.yo_slot_home.argtypes = [u64, C.c_uint32]
    L.yo_bloom_init.restype = p; L.yo_bloom_init.argtypes = [C.c_int, C.c_int]
    L.yo_bloom_insert.restype = C.c_int; L.yo_bloom_insert.argtypes = [p, u64]
    L.yo_bloom_destroy.argtypes = [p]
    L.yo_ch_init.restype = C.POINTER(YoCh); L.yo_ch_init.argtypes = [C.c_int] * 4
    L.yo_ch_destroy.argtypes = [C.POINTER(YoCh)]
    L.yo_ch_destroy_bf.argtypes = [C.POINTER(YoCh)]
    L.yo_ch_insert_list.restype = C.c_int
    L.yo_ch_insert_list.argtypes = [C.POINTER(YoCh), C.c_int, C.c_int, C.POINTER(u64)]
    L.yo_ch_insert_events.argtypes = [C.POINTER(YoCh), C.c_int, i64, C.POINTER(u64)]
    L.yo_ch_get.restype = C.c_int; L.yo_ch_get.argtypes = [C.POINTER(YoCh), u64]
    L.yo_ch_clear.argtypes = [C.POINTER(YoCh)]
    L.yo_ch_hist.argtypes = [C.POINTER(YoCh), C.POINTER(i64)]
    L.yo_ch_shrink.argtypes = [C.POINTER(YoCh), C.c_int, C.c_int]
    L.yo_ch_tighten.argtypes = [C.POINTER(YoCh)]
    L.yo_ch_setcnt.argtypes = [C.POINTER(YoCh), C.c_int]
    L.yo_ch_merge.argtypes = [C.POINTER(YoCh), C.POINTER(YoCh), C.c_int, C.c_int, C.c_int]
    L.yo_ch_subtract.argtypes = [C.POINTER(YoCh), C.POINTER(YoCh)]
    L.yo_ch_isec.argtypes = [C.POINTER(YoCh), C.POINTER(YoCh)]
    L.yo_ch_dump.restype = C.c_int; L.yo_ch_dump.argtypes = [C.POINTER(YoCh), C.c_char_p]
    L.yo_ch_dump_mem.restype = i64; L.yo_ch_dump_mem.argtypes = [C.POINTER(YoCh), C.POINTER(p)]
    L.yo_ch_restore.restype = C.POINTER(YoCh); L.yo_ch_restore.argtypes = [C.c_char_p]
    L.yo_extract.restype = i64; L.yo_extract.argtypes = [C.c_int, i64, C.c_char_p, C.POINTER(u64)]
    L.yo_count_file.restype = C.POINTER(YoCh)
    L.yo_count_file.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(YoCh), C.POINTER(i64)]
    L.yo_count_seqs.restype = C.POINTER(YoCh)
    L.yo_count_seqs.argtypes = [i64, C.POINTER(i64), C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.POINTER(YoCh), C.POINTER(i64)]
    L.yo_qv_seqs.argtypes = [C.POINTER(YoCh), i64, C.POINTER(i64), C.c_char_p, C.c_int, C.c_double,
                             C.POINTER(i64), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.yo_set_resize.argtypes = [C.POINTER(YoSet), C.c_uint32]
    L.yo_set_put.restype = C.c_uint32; L.yo_set_put.argtypes = [C.POINTER(YoSet), u64, C.POINTER(C.c_int)]
    L.yo_set_get.restype = C.c_uint32; L.yo_set_get.argtypes = [C.POINTER(YoSet), u64]
    L.yo_set_capacity.restype = C.c_uint32; L.yo_set_capacity.argtypes = [C.POINTER(YoSet)]
    _lib = L
    return L


def dump_bytes(h) -> bytes:
    L = lib()
    out = C.c_void_p()
    n = L.yo_ch_dump_mem(h, C.byref(out))
    data = C.string_at(out, n)
    C.CDLL(None).free(out)
    return data


def count_file(fn: str, k=31, pre=12, bf_shift=0, bf_n_hash=4, two_pass=None, fn2=None, chunk_size=10_000_000):
    """The `yak count` protocol of main.c:53-60 on the oracle; returns (handle, n_events).  chunk_size = the reference's -K
    (it only matters behind a truncated FASTQ record, count.c:93,109)."""
    L = lib()
    ne = C.c_int64()
    L.yo_count_file_chunked.restype = C.POINTER(YoCh)
    L.yo_count_file_chunked.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(YoCh), C.POINTER(C.c_int64), C.c_int64]
    h = L.yo_count_file_chunked(fn.encode(), k, pre, bf_shift, bf_n_hash, None, C.byref(ne), chunk_size)
    if not h:
        return None, 0
    if two_pass is None:
        two_pass = bf_shift > 0
    if two_pass:
        L.yo_ch_destroy_bf(h)
        L.yo_ch_clear(h)
        L.yo_count_file_chunked((fn2 or fn).encode(), k, pre, bf_shift, bf_n_hash, h, None, chunk_size)
        L.yo_ch_shrink(h, 2, 1023)
    return h, ne.value


def ref_count(fn: str, out: str, k=31, pre=12, bf_shift=0, bf_n_hash=4, threads=4, extra=()):
    """Run the unmodified reference binary (oracle/_ref/yak count)."""
    cmd = [REF_YAK, "count", f"-k{k}", f"-p{pre}", f"-t{threads}", f"-H{bf_n_hash}", "-o", out]
    if bf_shift > 0:
        cmd.append(f"-b{bf_shift}")
    cmd += list(extra) + [fn]
    return subprocess.run(cmd, check=True, capture_output=True)
