"""world_size-2/4 gloo runs of the multi-GPU routing logic (yak_b200/dist.py) on CPU.

The per-rank compute is stood in by the oracle (tests only); what is exercised is the product's
host logic for N>1: contiguous read slices per rank, owner-grouped events, the count + payload
all-to-all, source-rank-ordered receive, shard-local two-pass protocol, rank-ordered dump.
The result must be byte-identical to a single-table count of the same file.
"""
import ctypes as C
import os
import socket
import struct
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_util as G
import oracle_lib as O
import util


class OracleBackend:
    device = torch.device("cpu")

    def __init__(self, k, pre, bf_shift, bf_n_hash, rank, world):
        self.L = O.lib()
        self.k, self.pre, self.rank, self.world = k, pre, rank, world
        self.lw = world.bit_length() - 1
        self.h = self.L.yo_ch_init(k, pre, bf_n_hash, bf_shift)

    def to_device(self, asc):
        return asc

    def extract_route(self, asc):
        data = bytes(asc)
        evs = []
        for s in data.split(b"\n"):
            if len(s) < self.k:
                continue
            buf = (C.c_uint64 * len(s))()
            n = self.L.yo_extract(self.k, len(s), s, buf)
            evs.append(np.frombuffer(buf, dtype=np.uint64, count=n).copy())
        ev = np.concatenate(evs) if evs else np.zeros(0, dtype=np.uint64)
        owner = ((ev & np.uint64((1 << self.pre) - 1)) >> np.uint64(self.pre - self.lw)).astype(np.int64)
        order = np.argsort(owner, kind="stable")
        counts = np.bincount(owner, minlength=self.world).tolist()
        return torch.from_numpy(ev[order].view(np.int64).copy()), torch.tensor([int(c) for c in counts], dtype=torch.int64)

    def count_events(self, ev, create_new):
        a = np.ascontiguousarray(ev.numpy().view(np.uint64))
        mine = ((a & np.uint64((1 << self.pre) - 1)) >> np.uint64(self.pre - self.lw)) == np.uint64(self.rank)
        assert mine.all(), "received an event of a sub-table this rank does not own"
        self.L.yo_ch_insert_events(self.h, create_new, len(a), a.ctypes.data_as(C.POINTER(C.c_uint64)))
        return len(a)

    def destroy_bf(self): self.L.yo_ch_destroy_bf(self.h)
    def clear(self): self.L.yo_ch_clear(self.h)
    def shrink(self, lo, hi): self.L.yo_ch_shrink(self.h, lo, hi)
    def tot(self): return int(self.h.contents.tot)

    def dump_shard(self, with_header):
        full = O.dump_bytes(self.h)
        per = (1 << self.pre) >> self.lw
        off = 16
        start = None
        for s in range(1 << self.pre):
            if s == self.rank * per:
                start = off
            if s == (self.rank + 1) * per:
                break
            cap, size = struct.unpack_from("<II", full, off)
            off += 8 + 8 * size
        return (full[:16] if with_header else b"") + full[start:off]


def _worker(rank, world, port, fn, k, pre, b, out, batch_bases=0):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    from yak_b200 import dist as yd
    be = OracleBackend(k, pre, b, 4, rank, world)
    sc = yd.count_file_sharded(fn, be, records_per_chunk=1000, k=k, two_pass=b > 0, batch_bases=batch_bases)
    tot = sc.total_distinct()
    data = sc.dump_bytes()
    assert sc.dump_file(out + ".par") == os.path.getsize(out + ".par")     # every rank writes its part of the file itself
    if rank == 0:
        open(out, "wb").write(data)
        open(out + ".tot", "w").write(str(tot))
        assert open(out + ".par", "rb").read() == data
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,k,pre,b", [(2, 31, 12, 0), (2, 31, 10, 20), (4, 21, 11, 21)])
def test_sharded_count_equals_single_table(world, k, pre, b):
    fn = G.input_path("reads_q")
    out = os.path.join(util.TMP, f"yakb_dist_{world}_{k}_{pre}_{b}.yak")
    mp.spawn(_worker, args=(world, _free_port(), fn, k, pre, b, out), nprocs=world, join=True)
    h, _ = O.count_file(fn, k=k, pre=pre, bf_shift=b)
    want = O.dump_bytes(h)
    got = open(out, "rb").read()
    assert got == want, util.explain_diff(got, want)
    assert int(open(out + ".tot").read()) == h.contents.tot
    O.lib().yo_ch_destroy(h)


@pytest.mark.parametrize("world,k,pre,b,batch", [(2, 31, 12, 0, 5000), (2, 31, 10, 20, 100_000), (4, 21, 11, 21, 20_000)])
def test_sharded_count_through_the_parser_pool(world, k, pre, b, batch):
    """the fast file path of N>1 (bench.py's e2e at N GPUs): every rank parses the same batches with the library's
    parser pool and keeps its contiguous part of each; same bytes as a single-table count"""
    fn = G.input_path("reads_q")
    out = os.path.join(util.TMP, f"yakb_distpool_{world}_{k}_{pre}_{b}.yak")
    mp.spawn(_worker, args=(world, _free_port(), fn, k, pre, b, out, batch), nprocs=world, join=True)
    h, _ = O.count_file(fn, k=k, pre=pre, bf_shift=b)
    want = O.dump_bytes(h)
    got = open(out, "rb").read()
    assert got == want, util.explain_diff(got, want)
    assert int(open(out + ".tot").read()) == h.contents.tot


@pytest.mark.parametrize("kind", ["bgzf", "gzip"])
def test_sharded_count_of_compressed_input_is_read_once_by_rank0(kind):
    """gzip / blocked-gzip input on N > 1: rank 0 inflates and parses (csrc/bgzf.cpp, fastx.cpp) into the shared staging
    buffer, every rank takes its part; same bytes as a single-table count of the plain file"""
    import gzip
    import test_bgzf_cpu as B
    world, k, pre, b = 2, 31, 10, 20
    plain = G.input_path("reads_q")
    text = open(plain, "rb").read()
    fn = os.path.join(util.TMP, f"yakb_dist_{kind}.fq.gz")
    with open(fn, "wb") as f:
        f.write(B.bgzf_bytes(text, 40_000) if kind == "bgzf" else gzip.compress(text))
    out = os.path.join(util.TMP, f"yakb_dist_{kind}.yak")
    mp.spawn(_worker, args=(world, _free_port(), fn, k, pre, b, out, 50_000), nprocs=world, join=True)
    h, _ = O.count_file(plain, k=k, pre=pre, bf_shift=b)
    want = O.dump_bytes(h)
    got = open(out, "rb").read()
    assert got == want, util.explain_diff(got, want)
    O.lib().yo_ch_destroy(h)


def test_slice_bounds_cut_at_record_boundaries():
    """_slice_bounds (the per-rank cut of a parsed batch): parts are contiguous, cover the batch, and every cut is right behind a
    newline - for short reads, for records longer than a part, for a batch that is one record, for an empty batch"""
    from yak_b200.dist import _slice_bounds
    rng = np.random.default_rng(9)
    for trial in range(300):
        G_ = int(rng.choice([1, 2, 4, 8]))
        kind = trial % 4
        if kind == 0:
            lens = rng.integers(1, 300, int(rng.integers(0, 400)))
        elif kind == 1:
            lens = rng.integers(1, 200_000, int(rng.integers(1, 6)))           # records longer than a part and than the search window
        elif kind == 2:
            lens = np.array([int(rng.integers(1, 100_000))])
        else:
            lens = np.concatenate([rng.integers(1, 50, 200), [150_000], rng.integers(1, 50, 3)])
        buf = np.concatenate([np.concatenate([np.full(int(n), 65, np.uint8), [10]]) for n in lens]) if len(lens) else np.zeros(0, np.uint8)
        buf = buf.astype(np.uint8)
        n = len(buf)
        b = _slice_bounds(buf, n, G_)
        assert len(b) == G_ + 1 and b[0] == 0 and b[-1] == n
        assert all(b[i] <= b[i + 1] for i in range(G_))
        for x in b[1:-1]:
            assert x == 0 or buf[x - 1] == 10, (trial, x)
