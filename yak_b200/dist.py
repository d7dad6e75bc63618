"""Multi-GPU k-mer counting: sub-table shards + one all-to-all per chunk.

One process per GPU (torchrun).  Rank r owns the contiguous sub-table range
[r*P/G, (r+1)*P/G).  For every chunk each rank extracts the hashed k-mers of ITS contiguous slice
of the chunk's reads, stably grouped by owner rank; one all-to-all (counts, then payload) routes
them; the receiver sees source ranks 0..G-1 in order, i.e. file order per sub-table, and feeds the
run to its shard.  torch.distributed is the plumbing (NCCL on GPUs; gloo in the CPU tests); the
compute is the C library behind `GpuBackend` (tests substitute an oracle-backed backend to
exercise exactly this routing logic without a GPU).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


class GpuBackend:
    """The product: libyakb200 on the current CUDA device."""

    def __init__(self, k, pre, bf_shift, bf_n_hash, rank, world):
        from . import capi
        capi.require_gpu()
        self.capi, self.lib = capi, capi.lib()
        self.k, self.pre, self.rank, self.world = k, pre, rank, world
        self.h = self.lib.yakb_ch_init_shard(k, pre, bf_n_hash, bf_shift, rank, world)
        if not self.h:
            raise RuntimeError("yakb_ch_init_shard failed")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.stats = (C.c_uint64 * 4)()
        self._out = self._recv = self._out1 = self._recv1 = None   # grow-only exchange buffers (two sets: count_rounds works a round ahead)

    def _buf(self, name: str, n: int) -> torch.Tensor:
        cur = getattr(self, name)
        if cur is None or cur.numel() < n:
            setattr(self, name, None)
            cur = torch.empty(int(n * 1.1) + 1024, dtype=torch.int64, device=self.device)
            setattr(self, name, cur)
        return cur

    def recv_buffer(self, n: int, slot: int = 0) -> torch.Tensor:
        return self._buf("_recv" if slot == 0 else "_recv1", n)[:n]

    def to_device(self, asc: bytes | np.ndarray | torch.Tensor) -> torch.Tensor:
        if isinstance(asc, torch.Tensor):
            return asc.to(self.device)
        a = np.frombuffer(asc, dtype=np.uint8) if isinstance(asc, (bytes, bytearray)) else asc
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.device)

    def extract_route(self, asc: torch.Tensor, slot: int = 0):
        """hashed k-mers of `asc` grouped by owner rank (file order inside a group) and the events per owner as a DEVICE
        tensor: nothing here waits for the GPU (the host learns the counts once, from the all-gather of the exchange)"""
        n = asc.numel()
        out = self._buf("_out" if slot == 0 else "_out1", max(n, 1))
        counts = torch.zeros(self.world, dtype=torch.int64, device=self.device)
        rc = self.lib.yakb_extract_route_async(asc.data_ptr(), n, self.k, self.pre, self.world, out.data_ptr(), counts.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            raise RuntimeError("yakb_extract_route_async failed")
        return out, counts

    def count_events(self, ev: torch.Tensor, create_new: int) -> int:
        torch.cuda.current_stream().synchronize()
        rc = self.lib.yakb_count_events_dev(self.h, ev.data_ptr(), ev.numel(), create_new, self.stats)
        if rc != 0:
            raise RuntimeError("yakb_count_events_dev failed")
        return int(self.stats[0])

    def destroy_bf(self): self.lib.yak_ch_destroy_bf(self.h)
    def clear(self): self.lib.yak_ch_clear(self.h, 1)
    def shrink(self, lo, hi): self.lib.yak_ch_shrink(self.h, lo, hi, 1)
    def tot(self): return int(self.h.contents.tot)

    def dump_shard(self, with_header: bool) -> bytes:
        out = C.c_void_p()
        n = self.lib.yakb_ch_dump_shard_mem(self.h, int(with_header), C.byref(out))
        if n < 0:
            raise RuntimeError("yakb_ch_dump_shard_mem failed")
        data = C.string_at(out, n)
        C.CDLL(None).free(out)
        return data

    def dump_shard_size(self, with_header: bool) -> int:
        n = self.lib.yakb_ch_dump_shard_size(self.h, int(with_header))
        if n < 0:
            raise RuntimeError("yakb_ch_dump_shard_size failed")
        return int(n)

    def dump_shard_at(self, with_header: bool, path: str, offset: int) -> int:
        n = self.lib.yakb_ch_dump_shard_at(self.h, int(with_header), path.encode(), offset)
        if n < 0:
            raise RuntimeError("yakb_ch_dump_shard_at failed")
        return int(n)

    def close(self):
        self._out = self._recv = self._out1 = self._recv1 = None
        if self.h:
            self.lib.yak_ch_destroy(self.h)
            self.h = None


class ShardedCounter:
    """`yak count` over `world` shards.  `backend` does the per-rank compute."""

    def __init__(self, backend, group=None):
        self.b = backend
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.events = 0
        self._a2a_events, self._a2a_total, self.a2a_bytes = [], 0.0, 0

    def count_chunk(self, asc_local, create_new: int = 1) -> int:
        """asc_local: this rank's contiguous slice of the chunk (ASCII, any non-ACGTU byte separates reads)."""
        import os
        import time
        dbg = self.rank == 0 and os.environ.get("YAKB_DIST_TIMING") == "1"
        t = [time.time()]

        def mark():
            if dbg:
                torch.cuda.synchronize()
                t.append(time.time())
        ev, c_out = self.b.extract_route(self.b.to_device(asc_local))
        mark()
        if self.world == 1:
            recv = ev[:int(c_out[0])]
        else:
            dev = ev.device
            cuda = dev.type == "cuda"
            # every rank's per-owner counts in one small all-gather; its copy to the host is the only point of the
            # exchange where the host waits for the device.  m[s][d] = events rank s sends to rank d
            allc = torch.empty(self.world * self.world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allc, c_out.contiguous(), group=self.group)
            m = allc.view(self.world, self.world).cpu()
            mark()
            out_splits = [int(x) for x in m[self.rank].tolist()]
            in_splits = [int(x) for x in m[:, self.rank].tolist()]
            recv = self.b.recv_buffer(sum(in_splits)) if hasattr(self.b, "recv_buffer") else torch.empty(sum(in_splits), dtype=torch.int64, device=dev)
            if cuda:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            dist.all_to_all_single(recv, ev[:sum(out_splits)], output_split_sizes=in_splits, input_split_sizes=out_splits,
                                   group=self.group)                        # the one payload all-to-all of the chunk
            if cuda:
                e1.record()
                self._a2a_events.append((e0, e1))
                self.a2a_bytes += 8 * (sum(out_splits) - out_splits[self.rank])
            mark()
        n = self.b.count_events(recv, create_new)
        mark()
        if dbg:
            import sys
            print("[T::count_chunk] extract %.1f ms, counts %.1f ms, all-to-all %.1f ms, count %.1f ms (%d events in)" %
                  tuple([(b - a) * 1e3 for a, b in zip(t, t[1:])] + [recv.numel()]), file=sys.stderr)
        self.events += n
        return n

    def count_rounds(self, slices, create_new: int = 1) -> int:
        """Several chunks in a row, software-pipelined: while the shard counts round i (the library's own stream), a helper thread
        extracts round i+1 and runs its exchange (its own stream, NCCL's stream) - the compute-bound extraction and the NVLink
        transfer hide behind the memory-bound count.  Two sets of exchange buffers; a round is extracted only when the count of
        the round before last has finished with its buffers.  Results are those of count_chunk per slice."""
        slices = list(slices)
        if self.world == 1 or len(slices) < 2 or self._dev().type != "cuda" or not hasattr(self.b, "recv_buffer"):
            return sum(self.count_chunk(s, create_new) for s in slices)
        import queue
        import threading
        ready: "queue.Queue" = queue.Queue()
        free = threading.Semaphore(2)
        err = []
        dev = self._dev()

        def producer():
            try:
                torch.cuda.set_device(dev)
                stream = torch.cuda.Stream(device=dev)
                with torch.cuda.stream(stream):
                    for i, sl in enumerate(slices):
                        free.acquire()
                        recv = self._extract_exchange(sl, i & 1)
                        stream.synchronize()            # the routed events of round i are all here
                        ready.put(recv)
            except BaseException as e:                  # noqa: BLE001 - handed to the caller
                err.append(e)
                ready.put(None)

        th = threading.Thread(target=producer, daemon=True)
        th.start()
        n = 0
        for _ in slices:
            recv = ready.get()
            if recv is None:
                break
            n += self.b.count_events(recv, create_new)
            free.release()
        th.join()
        if err:
            raise err[0]
        self.events += n
        return n

    def _extract_exchange(self, asc_local, slot: int) -> torch.Tensor:
        """count_chunk without the count, on the buffers of `slot`: this rank's share of the events after the exchange"""
        ev, c_out = self.b.extract_route(self.b.to_device(asc_local), slot)
        dev = ev.device
        allc = torch.empty(self.world * self.world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, c_out.contiguous(), group=self.group)
        m = allc.view(self.world, self.world).cpu()
        out_splits = [int(x) for x in m[self.rank].tolist()]
        in_splits = [int(x) for x in m[:, self.rank].tolist()]
        recv = self.b.recv_buffer(sum(in_splits), slot)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_to_all_single(recv, ev[:sum(out_splits)], output_split_sizes=in_splits, input_split_sizes=out_splits, group=self.group)
        e1.record()
        self._a2a_events.append((e0, e1))
        self.a2a_bytes += 8 * (sum(out_splits) - out_splits[self.rank])
        return recv

    def a2a_ms(self) -> float:
        """device time of the payload all-to-alls so far (call after a synchronize)"""
        t = sum(a.elapsed_time(b) for a, b in self._a2a_events)
        self._a2a_events.clear()
        self._a2a_total += t
        return self._a2a_total

    def second_pass_prepare(self):
        """main.c:55-56 on every shard."""
        self.b.destroy_bf()
        self.b.clear()

    def shrink(self, lo=2, hi=1023):
        self.b.shrink(lo, hi)

    def total_distinct(self) -> int:
        t = torch.tensor([self.b.tot()], dtype=torch.int64, device=self._dev())
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
        return int(t[0])

    def _dev(self):
        return getattr(self.b, "device", torch.device("cpu"))

    def dump_bytes(self) -> bytes | None:
        """The whole .yak image on rank 0 (rank-ordered concatenation of the shard images)."""
        part = self.b.dump_shard(self.rank == 0)
        if self.world == 1:
            return part
        parts = [None] * self.world if self.rank == 0 else None
        dist.gather_object(part, parts, dst=0, group=self.group)
        return b"".join(parts) if self.rank == 0 else None


    def dump_file(self, path: str) -> int:
        """The .yak file written by all ranks of one node side by side: every rank writes its shard image at its own offset
        (rank-ordered concatenation, htab.c:373-394); returns the file size.  No shard travels between processes - or through
        Python: a backend with `dump_shard_size` / `dump_shard_at` (the GPU library) writes straight from C."""
        import os
        direct = hasattr(self.b, "dump_shard_at")
        part = None if direct else self.b.dump_shard(self.rank == 0)
        mine = self.b.dump_shard_size(self.rank == 0) if direct else len(part)
        if self.world == 1:
            with open(path, "wb") as f:
                if not direct:
                    f.write(part)
            if direct:
                self.b.dump_shard_at(True, path, 0)
            return mine
        lens = torch.zeros(self.world, dtype=torch.int64, device=self._dev())
        lens[self.rank] = mine
        dist.all_reduce(lens, group=self.group)
        off = [0] + [int(x) for x in torch.cumsum(lens, 0).tolist()]
        if self.rank == 0:
            with open(path, "wb") as f:
                f.truncate(off[-1])
        dist.barrier(group=self.group)
        if direct:
            self.b.dump_shard_at(self.rank == 0, path, off[self.rank])
        else:
            fd = os.open(path, os.O_WRONLY)
            try:
                done = 0
                view = memoryview(part)
                while done < len(part):
                    done += os.pwrite(fd, view[done:done + (1 << 30)], off[self.rank] + done)
            finally:
                os.close(fd)
        dist.barrier(group=self.group)
        return off[-1]


_STAGES: dict = {}


def _slice_bounds(view: np.ndarray, n: int, G: int) -> list[int]:
    """cut [0, n) of a "SEQ\nSEQ\n..." buffer into G contiguous parts at record boundaries, about equal in bytes"""
    b = [0]
    for r in range(1, G):
        p = max(b[-1], r * n // G)
        w = 1 << 12
        while True:                       # last newline before p (records can be long: widen the window)
            lo = max(b[-1], p - w)
            hit = np.flatnonzero(view[lo:p] == 10)
            if hit.size or lo == b[-1]:
                break
            w <<= 2
        b.append(lo + int(hit[-1]) + 1 if hit.size else b[-1])
    b.append(n)
    return b


def _count_file_sharded_pool(fn, sc: ShardedCounter, k, create_new, batch_bases, group):
    """One pass over a file read ONCE, by rank 0: a plain file with the library's parser pool (csrc/fastx_par.cpp, all the
    host's cores), gzip / blocked gzip with the sequential reader (csrc/fastx.cpp; BGZF blocks inflated by a pool of
    threads, csrc/bgzf.cpp).  Rank 0 parses into a staging buffer in shared memory that every rank of the node maps; per
    batch it broadcasts the batch length, every rank cuts the batch into G contiguous parts at record boundaries (same
    arithmetic everywhere), and ships only ITS part to its GPU.  While the ranks work on batch i, rank 0 already fills
    the other half of the buffer with batch i+1 (a rank has read its part of batch i-1 before it entered that batch's
    all-to-all, so the half is free).  False if rank 0 cannot open the file: the caller falls back (and reports it)."""
    import os
    import tempfile
    import threading
    from . import capi
    L = capi.lib()
    G, r, dev = sc.world, sc.rank, sc._dev()
    cap = batch_bases + (batch_bases >> 4) + 4096
    rd, fill_fn, close_fn = None, L.yakb_pfastx_fill, L.yakb_pfastx_close
    if r == 0:
        rd = L.yakb_pfastx_open(fn.encode(), 0, 0)
        if not rd and fn != "-":          # gzip / BGZF: the sequential reader has the same bulk call
            rd, fill_fn, close_fn = L.yakb_fastx_open(fn.encode()), L.yakb_fastx_fill, L.yakb_fastx_close
    ok = torch.tensor([1 if (r != 0 or rd) else 0], dtype=torch.int64, device=dev)
    if G > 1:
        dist.broadcast(ok, 0, group=group)
    if int(ok[0]) == 0:
        return False
    use_shm = 0
    if r == 0:
        try:
            st = os.statvfs("/dev/shm")
            use_shm = int(os.access("/dev/shm", os.W_OK) and st.f_bavail * st.f_frsize >= 2 * cap + (64 << 20))
        except OSError:
            pass
    where = torch.tensor([use_shm], dtype=torch.int64, device=dev)
    if G > 1:
        dist.broadcast(where, 0, group=group)     # one decision for the node
    stage_dir = "/dev/shm" if int(where[0]) else tempfile.gettempdir()
    mm = _STAGES.get((id(group), cap))        # mapped (and page-locked for DMA) once per process, reused by later passes
    fresh = torch.tensor([0 if mm is not None else 1], dtype=torch.int64, device=dev)
    if G > 1:
        dist.all_reduce(fresh, op=dist.ReduceOp.MAX, group=group)
    if int(fresh[0]):
        # rank 0 creates the file under a fresh random name (O_EXCL, no symlink is followed) and tells the others:
        # two jobs of one user can never map each other's staging buffer
        name = [None]
        if r == 0:
            fd, name[0] = tempfile.mkstemp(prefix="yakb_stage_", dir=stage_dir)
            os.ftruncate(fd, 2 * cap)
            os.close(fd)
        if G > 1:
            dist.broadcast_object_list(name, 0, group=group)
        stage = name[0]
        mm = np.memmap(stage, dtype=np.uint8, mode="r+", shape=(2 * cap,))
        if dev.type == "cuda":
            torch.cuda.cudart().cudaHostRegister(mm.ctypes.data, 2 * cap, 0)   # best effort: pageable copies work too
        _STAGES[(id(group), cap)] = mm
        if G > 1:
            dist.barrier(group=group)
        if r == 0:
            os.unlink(stage)                  # the mappings keep the memory alive; nothing is left behind in /dev/shm
    state = [None, None]

    def fill(slot):
        ns, done, need = C.c_int64(), C.c_int(), C.c_uint64()
        n = fill_fn(rd, mm.ctypes.data + slot * cap, cap, batch_bases, k, C.byref(ns), C.byref(done), C.byref(need))
        state[slot] = (int(n), 1 if done.value else 0, int(need.value))

    try:
        slot = 0
        if r == 0:
            fill(0)
        while True:
            hdr = torch.tensor(list(state[slot]) if r == 0 else [0, 0, 0], dtype=torch.int64, device=dev)
            if G > 1:
                dist.broadcast(hdr, 0, group=group)
            n, done, need = (int(x) for x in hdr.tolist())
            if need:
                raise RuntimeError("a record larger than the staging buffer; raise batch_bases")
            th = None
            if r == 0 and not done:       # the pool parses the next batch while this one is on the devices
                th = threading.Thread(target=fill, args=(slot ^ 1,))
                th.start()
            view = mm[slot * cap: slot * cap + n]
            b = _slice_bounds(view, n, G)
            sc.count_chunk(view[b[r]:b[r + 1]], create_new)   # also for an empty part: every rank joins every all-to-all
            if th:
                th.join()
            if done:
                break
            slot ^= 1
    finally:
        if rd:
            close_fn(rd)
    return True


def _count_file_sharded_ingest(fn, sc: ShardedCounter, k, create_new, batch_bytes, group):
    """One pass over a plain file in the strict 4-line FASTQ / 2-line FASTA layout with NO host parsing: every rank maps the
    file, takes its own byte range of every batch (cut where a record starts: csrc/capi.cu record_start_before, the same
    arithmetic on every rank), copies it to its GPU and has the device turn it into the base stream, checking the layout of
    every record (csrc/ingest.cu).  Returns True when the whole file went through; False when the file does not qualify or
    its FIRST batch fails the check (nothing was counted: the caller takes the parser path).  A later failure raises."""
    import os
    import threading
    from . import capi
    L = capi.lib()
    G, r, dev = sc.world, sc.rank, sc._dev()
    if dev.type != "cuda" or os.environ.get("YAKB_GPU_INGEST", "") == "0":
        return False
    try:
        size = os.path.getsize(fn)
        mm = np.memmap(fn, dtype=np.uint8, mode="r") if size > 0 else None
    except (OSError, ValueError):
        return False
    if mm is None or size < (256 << 20) and os.environ.get("YAKB_GPU_INGEST", "") != "1":
        return False
    marker = int(mm[0])
    lpr = 4 if marker == ord("@") else 2 if marker == ord(">") else 0
    if lpr == 0 or int(mm[size - 1]) != 10:
        return False
    base = mm.ctypes.data
    B = max(1 << 16, int(batch_bytes))

    def cut(lo, pos):
        return int(L.yakb_record_start_before(base, lo, min(pos, size), size, lpr)) if pos < size else size

    pinned = [torch.empty(0, dtype=torch.uint8).pin_memory(), torch.empty(0, dtype=torch.uint8).pin_memory()]
    d_res = torch.zeros(3, dtype=torch.int64, device=dev)
    state = {}

    def stage(slot, a):
        """my part of the batch that starts at a: -> (next batch start, bytes staged in pinned[slot])"""
        e = cut(a, a + B)
        if e <= a:
            e = -1                       # a record longer than a batch: not this path's business
            state[slot] = (e, 0)
            return
        n = e - a
        lo = a if r == 0 else cut(a, a + n * r // G)
        hi = e if r == G - 1 else cut(a, a + n * (r + 1) // G)
        m = max(0, hi - lo)
        if pinned[slot].numel() < m:
            pinned[slot] = torch.empty(int(m * 1.1) + 4096, dtype=torch.uint8).pin_memory()
        if m:                            # a few threads: one memcpy out of the page cache runs at 5-8 GB/s
            dst = pinned[slot].numpy()
            parts = 4 if m >= (32 << 20) else 1
            ths = [threading.Thread(target=np.copyto, args=(dst[m * i // parts:m * (i + 1) // parts], mm[lo + m * i // parts:lo + m * (i + 1) // parts]))
                   for i in range(parts)]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
        state[slot] = (e, m)

    a, slot, n_batches = 0, 0, 0
    stage(0, 0)
    while True:
        e, m = state[slot]
        th = None
        if 0 < e < size:
            th = threading.Thread(target=stage, args=(slot ^ 1, e))
            th.start()
        ok = 1
        dense = None
        if e < 0:
            ok = 0
        else:
            raw = pinned[slot][:m].to(dev, non_blocking=True)
            dense = torch.empty(max(m, 1), dtype=torch.uint8, device=dev)
            if L.yakb_ingest_dev(raw.data_ptr(), m, lpr, dense.data_ptr(), d_res.data_ptr(), torch.cuda.current_stream().cuda_stream) != 0:
                ok = 0
            res = d_res.cpu().tolist()
            if res[0] != 0:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int64, device=dev)
        if G > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag[0]) == 0:
            if th:
                th.join()
            if n_batches == 0:
                return False
            raise RuntimeError(f"{fn}: the strict {lpr}-line layout breaks after {n_batches} batches; rerun with YAKB_GPU_INGEST=0")
        sc.count_chunk(dense[:int(res[1])], create_new)
        n_batches += 1
        if th:
            th.join()
        if e >= size:
            break
        slot ^= 1
    return True


def count_file_sharded(fn: str, backend, records_per_chunk: int = 1 << 20, k: int = 31, two_pass: bool = False,
                       fn2: str | None = None, group=None, batch_bases: int = 0, timings: dict | None = None) -> ShardedCounter:
    """`yak count` of one shared file on all ranks of ONE node.  batch_bases > 0: plain files are parsed once by rank 0's
    parser pool into shared memory, each rank taking its contiguous part of every batch (the fast path, bench.py's e2e
    at N GPUs); otherwise, and for gzip / stdin, every rank walks the file record by record and keeps its slice of
    every chunk."""
    from . import capi
    L = capi.lib()
    sc = ShardedCounter(backend, group)
    G, r = sc.world, sc.rank
    per = max(1, records_per_chunk // G)

    def one_pass(path, create_new):
        # plain files in the strict layout: every rank ingests its own byte range on its GPU (2 bytes of FASTQ text per base)
        if batch_bases > 0 and _count_file_sharded_ingest(path, sc, k, create_new, 2 * batch_bases, group):
            return
        if batch_bases > 0 and _count_file_sharded_pool(path, sc, k, create_new, batch_bases, group):
            return
        rd = L.yakb_fastx_open(path.encode())
        if not rd:
            raise FileNotFoundError(path)
        cap = 1 << 24
        buf = np.empty(cap, dtype=np.uint8)
        nb, ns = C.c_uint64(), C.c_int64()
        while True:
            while True:
                used = L.yakb_fastx_read_slice(rd, r * per, per, k, buf.ctypes.data, cap, C.byref(nb), C.byref(ns))
                if used != -1:
                    break
                raise RuntimeError("record slice larger than the staging buffer; raise records_per_chunk granularity")
            tail = L.yakb_fastx_read_slice(rd, (G - 1 - r) * per, 0, k, buf.ctypes.data + nb.value, 0, C.byref(C.c_uint64()), C.byref(C.c_int64()))
            sc.count_chunk(buf[:nb.value].copy(), create_new)
            # every rank must take part in every all-to-all: stop only when ALL ranks ran dry
            more = torch.tensor([1 if used + max(tail, 0) == G * per else 0], dtype=torch.int64, device=sc._dev())
            if G > 1:
                dist.all_reduce(more, op=dist.ReduceOp.MAX, group=group)
            if int(more[0]) == 0:
                break
        L.yakb_fastx_close(rd)

    import time
    t0 = time.time()
    one_pass(fn, 1)
    if timings is not None:                   # bench.py's e2e: pass 1 up to the result every rank reads back
        timings["distinct_after_pass1"] = sc.total_distinct()
        if sc._dev().type == "cuda":
            torch.cuda.synchronize()
        timings["pass1_seconds"] = time.time() - t0
    if two_pass:
        sc.second_pass_prepare()
        one_pass(fn2 or fn, 0)
        sc.shrink(2, 1023)
    return sc


def main(argv=None) -> int:
    """`yak count` on all GPUs of one node: torchrun --nproc-per-node N -m yak_b200.dist count [options] in.fq [in2.fq]
    (options and two-pass protocol of the reference's main_count, main.c:13-64; one process per GPU, sub-tables sharded,
    one all-to-all per batch, every rank writes its part of the .yak file)."""
    import argparse
    import os
    import sys
    ap = argparse.ArgumentParser(prog="yak_b200.dist")
    ap.add_argument("command", choices=["count"])
    ap.add_argument("-k", type=int, default=31)
    ap.add_argument("-p", type=int, default=10)
    ap.add_argument("-b", type=int, default=0)
    ap.add_argument("-H", type=int, default=4)
    ap.add_argument("-t", type=int, default=4, help="accepted, unused")
    ap.add_argument("-K", default="100m", help="accepted; batches are sized for the GPUs")
    ap.add_argument("-o", default=None)
    ap.add_argument("files", nargs="+")
    a = ap.parse_args(argv)
    if a.p < 10:
        print("ERROR: -p should be at least 10", file=sys.stderr)
        return 1
    if a.k >= 64:
        print("ERROR: -k must be smaller than 64", file=sys.stderr)
        return 1
    rank, world, local = (int(os.environ.get(v, d)) for v, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    if world & (world - 1) or world > (1 << a.p):
        print("ERROR: the number of ranks must be a power of two", file=sys.stderr)
        return 1
    from . import capi
    capi.require_gpu()                    # no CPU path: fail loudly
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    be = GpuBackend(a.k, a.p, a.b, a.H, rank, world)
    try:
        batch = min(64 << 20, (512 << 20) // world) * world
        sc = count_file_sharded(a.files[0], be, k=a.k, two_pass=a.b > 0, fn2=a.files[1] if len(a.files) > 1 else None, batch_bases=batch)
        tot = sc.total_distinct()
        if rank == 0:
            print("[M::%s] %d distinct k-mers%s on %d GPUs" % ("yak_b200.dist", tot, " after shrinking" if a.b > 0 else "", world), file=sys.stderr)
        if a.o:
            sc.dump_file(a.o)
    finally:
        be.close()
        if world > 1:
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
