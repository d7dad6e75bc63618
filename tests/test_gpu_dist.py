"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the NCCL-routed sharded count
through the C ABI equals the oracle's single-table bytes."""
import os
import socket

import pytest

import golden_util as G
import oracle_lib as O
import util

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, fn, k, pre, b, out, batch_bases=0):
    import torch
    if batch_bases < 0:          # the per-rank device ingest (yak_b200/dist.py _count_file_sharded_ingest), forced on a small file
        os.environ["YAKB_GPU_INGEST"] = "1"
        batch_bases = -batch_bases
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    from yak_b200 import dist as yd
    be = yd.GpuBackend(k, pre, b, 4, rank, world)
    sc = yd.count_file_sharded(fn, be, records_per_chunk=2000, k=k, two_pass=b > 0, batch_bases=batch_bases)
    tot = sc.total_distinct()
    data = sc.dump_bytes()
    if rank == 0:
        open(out, "wb").write(data)
        open(out + ".tot", "w").write(str(tot))
    be.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("k,pre,b,batch", [(31, 12, 0, 0), (31, 10, 20, 0), (47, 12, 0, 0), (31, 12, 22, 60_000), (31, 10, 0, 1 << 20),
                                           (31, 12, 22, -40_000), (31, 10, 0, -(1 << 20))])
def test_nccl_sharded_count_equals_oracle(yakb, k, pre, b, batch):
    import torch
    import torch.multiprocessing as mp
    world = 1
    while world * 2 <= min(torch.cuda.device_count(), 8):
        world *= 2
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    fn = G.input_path("reads_q")
    out = os.path.join(util.TMP, f"yakb_nccl_{k}_{pre}_{b}.yak")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(world, port, fn, k, pre, b, out, batch), nprocs=world, join=True)
    h, _ = O.count_file(fn, k=k, pre=pre, bf_shift=b)
    want = O.dump_bytes(h)
    got = open(out, "rb").read()
    assert got == want, util.explain_diff(got, want)
    assert int(open(out + ".tot").read()) == h.contents.tot
    O.lib().yo_ch_destroy(h)


@pytest.mark.parametrize("b,compressed", [(0, False), (22, False), (20, True)])
def test_multi_gpu_count_command_equals_oracle(yakb, b, compressed):
    """`torchrun -m yak_b200.dist count ...` on every GPU of the box (1 works too): the .yak file the ranks write side by side"""
    import subprocess
    import sys
    import torch
    world = 1
    while world * 2 <= min(torch.cuda.device_count(), 8):
        world *= 2
    fn = G.input_path("reads_q")
    if compressed:
        import test_bgzf_cpu as B
        gz = os.path.join(util.TMP, "yakb_distcli.fq.gz")
        open(gz, "wb").write(B.bgzf_bytes(open(fn, "rb").read(), 50_000))
    out = os.path.join(util.TMP, f"yakb_distcli_{b}.yak")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), "-m", "yak_b200.dist", "count", "-k31", "-p12", f"-b{b}", "-o", out, gz if compressed else fn]
    env = dict(os.environ)
    if b == 22:                  # this case through the per-rank device ingest (forced: the file is small)
        env["YAKB_GPU_INGEST"] = "1"
    r = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    h, _ = O.count_file(fn, k=31, pre=12, bf_shift=b)
    want = O.dump_bytes(h)
    got = open(out, "rb").read()
    assert got == want, util.explain_diff(got, want)
    O.lib().yo_ch_destroy(h)


@pytest.mark.parametrize("k,pre", [(31, 12), (21, 10), (47, 11), (63, 12)])
def test_route_extraction_is_the_oracles_stable_grouping(yakb, k, pre):
    """yakb_extract_route_dev on ONE GPU for world = 1..16 owners (csrc/extras.cu route_tile_kernel / route_gather_kernel): the
    hashed k-mers of a read batch grouped by owner rank, file order inside a group (count.c:120,133 is what that order
    protects) - against the oracle's event stream of the same reads"""
    import ctypes as C
    import numpy as np
    import torch
    from yak_b200 import synth
    L, OL = yakb.lib(), O.lib()
    rng = np.random.default_rng(k * 100 + pre)
    reads = synth.codes_to_ascii(synth.read_codes(5, 400_000, 9, 0, 5000, 150, 0.01, 3))
    seqs = [bytes(r) for r in reads] + [b"ACGT" * 5, b"", b"A" * 700, bytes(rng.choice(list(b"ACGTN"), 3000).astype(np.uint8))]
    asc = b"\n".join(seqs) + b"\n"
    want = []
    for s in seqs:
        for part in s.split(b"N"):
            if len(part) >= k:
                buf = (C.c_uint64 * len(part))()
                n = OL.yo_extract(k, len(part), part, buf)
                want.append(np.frombuffer(buf, dtype=np.uint64, count=n).copy())
    want = np.concatenate(want)
    d_asc = torch.frombuffer(bytearray(asc), dtype=torch.uint8).cuda()
    for world in (1, 2, 4, 8, 16):
        lw = world.bit_length() - 1
        out = torch.zeros(len(asc), dtype=torch.int64, device="cuda")
        counts = (C.c_uint64 * world)()
        torch.cuda.synchronize()
        assert L.yakb_extract_route_dev(d_asc.data_ptr(), len(asc), k, pre, world, out.data_ptr(), counts, torch.cuda.current_stream().cuda_stream) == 0
        owner = ((want & np.uint64((1 << pre) - 1)) >> np.uint64(pre - lw)).astype(np.int64)
        order = np.argsort(owner, kind="stable")
        assert [int(c) for c in counts] == np.bincount(owner, minlength=world).tolist(), world
        got = out[:len(want)].cpu().numpy().view(np.uint64)
        assert np.array_equal(got, want[order]), world
    assert L.yakb_extract_route_dev(d_asc.data_ptr(), len(asc), k, pre, 32, out.data_ptr(), (C.c_uint64 * 32)(), torch.cuda.current_stream().cuda_stream) != 0
    assert L.yakb_extract_route_dev(d_asc.data_ptr(), len(asc), k, pre, 3, out.data_ptr(), (C.c_uint64 * 3)(), torch.cuda.current_stream().cuda_stream) != 0


def _gpus_pow2():
    import torch
    world = 1
    while world * 2 <= min(torch.cuda.device_count(), 8):
        world *= 2
    return world


@pytest.mark.parametrize("b,batch", [(0, "0"), (22, "70000"), (20, "1500")])
def test_in_library_multi_gpu_count_command_equals_oracle(yakb, b, batch):
    """`yak-b200 count -g G` (YAKB_GPUS): ONE process, one host thread per GPU, the routed k-mers pulled from the peers'
    buffers with peer copies (csrc/capi.cu multi_batch) - the plain-C boundary on several GPUs, no Python, no torchrun.
    Batches of 70 kB and 1.5 kB of bases cut the input into hundreds of exchanges."""
    import subprocess
    world = _gpus_pow2()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    fn = G.input_path("reads_q")
    out = os.path.join(util.TMP, f"yakb_multi_cli_{b}.yak")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    if batch != "0":
        env["YAKB_BATCH"] = batch
    r = subprocess.run([os.path.join(root, "yak_b200", "bin", "yak-b200"), "count", "-k31", "-p12", f"-b{b}", "-g", str(world), "-K", "1000", "-o", out, fn],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    assert f"on {world} GPUs" in r.stderr
    h, _ = O.count_file(fn, k=31, pre=12, bf_shift=b)
    want = O.dump_bytes(h)
    got = open(out, "rb").read()
    assert got == want, util.explain_diff(got, want)
    O.lib().yo_ch_destroy(h)


MULTI_SNIPPET = r"""
import sys, os
import ctypes as C
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np
import oracle_lib as O
from yak_b200 import capi
capi.require_gpu()
L, OL = capi.lib(), O.lib()
fn = {fn!r}
for k, pre, b in ((31, 10, 0), (47, 12, 21)):
    hg = capi.count_file(fn, k=k, pre=pre, bf_shift=b)
    assert L.yakb_ch_gpus(hg) == {world}
    ho, _ = O.count_file(fn, k=k, pre=pre, bf_shift=b)
    assert hg.contents.tot == ho.contents.tot
    assert capi.dump_bytes(hg) == O.dump_bytes(ho), (k, pre, b)
    h1, h2 = (C.c_int64 * 1024)(), (C.c_int64 * 1024)()
    OL.yo_ch_hist(ho, h1); L.yak_ch_hist(hg, h2, 4)
    assert list(h1) == list(h2)
    # lookups across the shards, present and absent (htab.c:93-100)
    seqs = [ln.strip() for ln in open(fn, "rb") if ln[:1] in b"ACGT"][:300]
    ev = []
    for s in seqs:
        buf = (C.c_uint64 * len(s))()
        n = OL.yo_extract(k, len(s), s, buf)
        ev.append(np.frombuffer(buf, dtype=np.uint64, count=n).copy())
    ev = np.concatenate(ev)
    probe = np.ascontiguousarray(np.concatenate([ev[::7], ev[::11] ^ np.uint64(0x5555555555)]), dtype=np.uint64)
    got = np.zeros(len(probe), dtype=np.int32)
    assert L.yakb_ch_get_batch(hg, len(probe), probe.ctypes.data_as(C.POINTER(C.c_uint64)), got.ctypes.data_as(C.POINTER(C.c_int32))) == 0
    want = np.array([OL.yo_ch_get(ho, int(x)) for x in probe], dtype=np.int32)
    assert np.array_equal(got, want)
    assert L.yak_ch_get(hg, int(probe[0])) == want[0]
    OL.yo_ch_shrink(ho, 3, 900); L.yak_ch_shrink(hg, 3, 900, 1)
    assert capi.dump_bytes(hg) == O.dump_bytes(ho), ("shrink", k, pre, b)
    assert hg.contents.tot == ho.contents.tot
    OL.yo_ch_clear(ho); L.yak_ch_clear(hg, 1)
    assert capi.dump_bytes(hg) == O.dump_bytes(ho), ("clear", k, pre, b)
    L.yak_ch_destroy(hg); OL.yo_ch_destroy(ho)
print("multi ok")
"""


def test_in_library_multi_gpu_table_operations(yakb):
    """the `yak count` flow of the C API on a table spread over the GPUs of one process (YAKB_GPUS): tot, dump, hist, get across
    shards, shrink, clear - against the oracle's single table"""
    import subprocess
    import sys
    world = _gpus_pow2()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ, YAKB_GPUS=str(world), YAKB_BATCH="200000")
    r = subprocess.run([sys.executable, "-c", MULTI_SNIPPET.format(root=root, fn=G.input_path("reads_q"), world=world)],
                       env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "multi ok" in r.stdout, r.stderr[-3000:]


def _rounds_worker(rank, world, port, fn, k, pre, b, out):
    import numpy as np
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from yak_b200 import dist as yd
    seqs = [ln.strip() for ln in open(fn, "rb") if ln[:1] in b"ACGTN"]
    be = yd.GpuBackend(k, pre, b, 4, rank, world)
    sc = yd.ShardedCounter(be)
    n_chunks = 7

    def my_part(c):      # chunk c of the file = reads [c*n/7, (c+1)*n/7); rank r takes its world-th of it (file order = chunk, then rank)
        lo, hi = len(seqs) * c // n_chunks, len(seqs) * (c + 1) // n_chunks
        a, e = lo + (hi - lo) * rank // world, lo + (hi - lo) * (rank + 1) // world
        return torch.from_numpy(np.frombuffer(b"".join(s + b"\n" for s in seqs[a:e]) or b"\n", dtype=np.uint8).copy())

    for create_new in ((1, 0) if b > 0 else (1,)):
        if create_new == 0:
            sc.second_pass_prepare()
        sc.count_rounds([my_part(c) for c in range(0, 4)], create_new)      # four rounds pipelined, then three
        sc.count_rounds([my_part(c) for c in range(4, 7)], create_new)
    if b > 0:
        sc.shrink(2, 1023)
    data = sc.dump_bytes()
    if rank == 0:
        open(out, "wb").write(data)
    be.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("k,pre,b", [(31, 12, 0), (31, 10, 21)])
def test_pipelined_rounds_equal_oracle(yakb, k, pre, b):
    """ShardedCounter.count_rounds: extraction + exchange of round i+1 behind the count of round i (helper thread, two buffer sets)"""
    import torch
    import torch.multiprocessing as mp
    world = _gpus_pow2()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    fn = G.input_path("reads_q")
    out = os.path.join(util.TMP, f"yakb_rounds_{k}_{pre}_{b}.yak")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_rounds_worker, args=(world, port, fn, k, pre, b, out), nprocs=world, join=True)
    h, _ = O.count_file(fn, k=k, pre=pre, bf_shift=b)
    want = O.dump_bytes(h)
    got = open(out, "rb").read()
    assert got == want, util.explain_diff(got, want)
    O.lib().yo_ch_destroy(h)
