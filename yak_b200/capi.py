"""ctypes binding of libyakb200.so - the C-ABI drop-in library (include/yak.h, include/yak_b200.h).

The Python side is plumbing only: it passes file names, host buffers or raw device pointers to
the C entry points.  There is no Python or CPU implementation of the path behind it; a missing
library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libyakb200.so")

u64, i64, i32, vp = C.c_uint64, C.c_int64, C.c_int32, C.c_void_p


class YakCopt(C.Structure):  # yak.h yak_copt_t
    _fields_ = [("bf_shift", i32), ("bf_n_hash", i32), ("k", i32), ("pre", i32), ("n_thread", i32),
                ("chunk_size", i64)]


class YakQopt(C.Structure):  # yak.h yak_qopt_t
    _fields_ = [("print_each", i32), ("print_err_kmer", i32), ("min_len", i32), ("n_threads", i32),
                ("min_frac", C.c_double), ("fpr", C.c_double), ("chunk_size", i64)]


class YakCh(C.Structure):  # yak.h yak_ch_t (public head)
    _fields_ = [("k", C.c_int), ("pre", C.c_int), ("n_hash", C.c_int), ("n_shift", C.c_int),
                ("tot", u64), ("h", vp)]


class YakKnt(C.Structure):
    _fields_ = [("x", u64), ("c", C.c_int)]


ChP = C.POINTER(YakCh)
_lib = None


def build(verbose: bool = False) -> None:
    subprocess.run(["make", "-j8", "-C", os.path.join(HERE, "csrc")], check=True,
                   stdout=None if verbose else subprocess.DEVNULL, stderr=None if verbose else subprocess.DEVNULL)


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no fallback implementation)")
    L = C.CDLL(LIB_PATH)
    sig = {
        "yak_copt_init": (None, [C.POINTER(YakCopt)]),
        "yak_qopt_init": (None, [C.POINTER(YakQopt)]),
        "yak_ch_init": (ChP, [C.c_int] * 4),
        "yak_ch_destroy": (None, [ChP]),
        "yak_ch_destroy_bf": (None, [ChP]),
        "yak_ch_insert_list": (C.c_int, [ChP, C.c_int, C.c_int, C.POINTER(u64)]),
        "yak_ch_get": (C.c_int, [ChP, u64]),
        "yak_ch_inc": (C.c_int, [ChP, u64]),
        "yak_ch_getseq": (C.POINTER(YakKnt), [ChP, C.c_int, C.POINTER(C.c_uint32)]),
        "yak_ch_clear": (None, [ChP, C.c_int]),
        "yak_ch_hist": (None, [ChP, C.POINTER(i64), C.c_int]),
        "yak_ch_shrink": (None, [ChP, C.c_int, C.c_int, C.c_int]),
        "yak_ch_setcnt": (None, [ChP, C.c_int, C.c_int]),
        "yak_ch_dump": (C.c_int, [ChP, C.c_char_p]),
        "yak_ch_restore": (ChP, [C.c_char_p]),
        "yak_count": (ChP, [C.c_char_p, C.POINTER(YakCopt), ChP]),
        "yak_recount": (None, [C.c_char_p, ChP]),
        "yak_qv": (None, [C.POINTER(YakQopt), C.c_char_p, ChP, C.POINTER(i64)]),
        "yak_bf_init": (vp, [C.c_int, C.c_int]),
        "yak_bf_destroy": (None, [vp]),
        "yak_bf_insert": (C.c_int, [vp, u64]),
        "yakb_version": (C.c_char_p, []),
        "yakb_device_count": (C.c_int, []),
        "yakb_count_ascii_dev": (C.c_int, [ChP, vp, u64, C.c_int, C.POINTER(u64)]),
        "yakb_count_ascii_host": (C.c_int, [ChP, C.c_char_p, u64, C.c_int, C.POINTER(u64)]),
        "yakb_count_events_dev": (C.c_int, [ChP, vp, u64, C.c_int, C.POINTER(u64)]),
        "yakb_extract_route_dev": (C.c_int, [vp, u64, C.c_int, C.c_int, C.c_int, vp, C.POINTER(u64), vp]),
        "yakb_extract_route_async": (C.c_int, [vp, u64, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
        "yakb_record_start_before": (u64, [vp, u64, u64, u64, C.c_int]),
        "yakb_ingest_dev": (C.c_int, [vp, u64, C.c_int, vp, vp, vp]),
        "yakb_ch_get_batch": (C.c_int, [ChP, u64, C.POINTER(u64), C.POINTER(i32)]),
        "yakb_ch_get_batch_dev": (C.c_int, [ChP, u64, vp, vp]),
        "yakb_qv_seqs": (C.c_int, [ChP, i64, C.POINTER(i64), C.c_char_p, C.c_int, C.c_double, C.POINTER(i64),
                                   C.POINTER(i32), C.POINTER(i32)]),
        "yakb_scan_seqs": (C.c_int, [ChP, i64, C.POINTER(i64), C.c_char_p, C.POINTER(C.c_int16)]),
        "yakb_ch_dump_mem": (i64, [ChP, C.POINTER(vp)]),
        "yakb_ch_init_shard": (ChP, [C.c_int] * 6),
        "yakb_ch_dump_shard_mem": (i64, [ChP, C.c_int, C.POINTER(vp)]),
        "yakb_ch_dump_shard_size": (i64, [ChP, C.c_int]),
        "yakb_ch_dump_shard_at": (i64, [ChP, C.c_int, C.c_char_p, u64]),
        "yakb_ch_reserve": (C.c_int, [ChP, u64]),
        "yakb_ch_stream": (vp, [ChP]),
        "yakb_ch_device_bytes": (u64, [ChP]),
        "yakb_kernel_launches": (u64, []),
        "yakb_ch_gpus": (C.c_int, [ChP]),
        "yakb_device_cache_bytes": (u64, []),
        "yakb_device_cache_trim": (None, []),
        "yakb_fastx_open": (vp, [C.c_char_p]),
        "yakb_fastx_next": (i64, [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]),
        "yakb_fastx_close": (None, [vp]),
        "yakb_fastx_fill": (i64, [vp, vp, u64, u64, C.c_int, C.POINTER(i64), C.POINTER(C.c_int), C.POINTER(u64)]),
        "yakb_pfastx_open": (vp, [C.c_char_p, u64, C.c_int]),
        "yakb_pfastx_fill": (i64, [vp, vp, u64, u64, C.c_int, C.POINTER(i64), C.POINTER(C.c_int), C.POINTER(u64)]),
        "yakb_pfastx_redo": (u64, [vp]),
        "yakb_fastx_set_chunk": (None, [vp, i64]),
        "yakb_ref_flow_sim": (None, [C.POINTER(i64), i64, C.c_int, i64, C.c_int, C.POINTER(C.c_uint8)]),
        "yakb_pfastx_set_chunk": (None, [vp, i64]),
        "yakb_fastx_set_workers": (None, [vp, C.c_int]),
        "yakb_pfastx_set_flow": (None, [vp, i64, C.c_int, C.c_int]),
        "yakb_pfastx_close": (None, [vp]),
        "yakb_fastx_read_slice": (i64, [vp, i64, i64, C.c_int, vp, u64, C.POINTER(u64), C.POINTER(i64)]),
        "yakb_prof_enable": (None, [C.c_int]),
        "yakb_prof_json": (C.c_int, [C.c_char_p, u64]),
        "yakb_synth_genome_dev": (C.c_int, [u64, u64, vp, vp]),
        "yakb_synth_reads_dev": (C.c_int, [vp, u64, u64, u64, u64, C.c_int, C.c_double, C.c_int, C.c_int, vp, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


def require_gpu() -> None:
    if lib().yakb_device_count() <= 0:
        raise RuntimeError("libyakb200 needs a CUDA device; there is no CPU path")


def dump_bytes(h) -> bytes:
    out = vp()
    n = lib().yakb_ch_dump_mem(h, C.byref(out))
    if n < 0:
        raise RuntimeError("yakb_ch_dump_mem failed")
    data = C.string_at(out, n)
    C.CDLL(None).free(out)
    return data


def copt(k=31, pre=10, bf_shift=0, bf_n_hash=4, n_thread=4, chunk_size=10_000_000) -> YakCopt:
    o = YakCopt()
    lib().yak_copt_init(C.byref(o))
    o.k, o.pre, o.bf_shift, o.bf_n_hash, o.n_thread, o.chunk_size = k, pre, bf_shift, bf_n_hash, n_thread, chunk_size
    return o


def count_file(fn: str, k=31, pre=12, bf_shift=0, bf_n_hash=4, fn2: str | None = None, chunk_size=10_000_000):
    """`yak count` (main.c:53-60) through the C API; returns the yak_ch_t* (caller destroys)."""
    L = lib()
    o = copt(k, pre, bf_shift, bf_n_hash, chunk_size=chunk_size)
    h = L.yak_count(fn.encode(), C.byref(o), None)
    if not h:
        return None
    if bf_shift > 0:
        L.yak_ch_destroy_bf(h)
        L.yak_ch_clear(h, o.n_thread)
        h = L.yak_count((fn2 or fn).encode(), C.byref(o), h)
        L.yak_ch_shrink(h, 2, 1023, o.n_thread)
    return h
