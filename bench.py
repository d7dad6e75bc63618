#!/usr/bin/env python
"""bench.py - k-mer events/s of the count hot path at BASELINE.json configs[1] geometry.

Workload (config.workload = "cfg2"): k=31, -p12, -b37 -H4, 150 bp reads sampled from a 3 Gbp
uniform genome (e=0.005, 1 % reads with an N), pass 1 of `yak count` (bloom + insert), the same
synthetic stream as yak_b200/synth.py generated on the device.  A *step* is one batch of reads
(--chunk-reads, default 2 M reads = 300 Mbp > L2) through the whole per-chunk path
(pack -> fused extract+probe -> ordered bloom/insert of pending events -> journal).  Steps are
consecutive batches from the start of the job; the table, bloom and journal persist across steps.
The default 3 + 297 steps are the WHOLE first pass of cfg2 (600 M reads = 90 Gbp = 30x of 3 Gbp),
table growth and rehash included (N ranks take N batches per step, so the default is 300 / N steps there:
the -b37 filter is sized for 90 Gbp, not for N times that).
Each step is bracketed by its own pair of CUDA events; the step's input batch is generated on the
device just before it, outside the timed region (max over ranks of the summed step times).

  value     events/s with the batch's ASCII bases already resident in HBM (CUDA events on the
            library's stream around exactly K steps, max over ranks)
  e2e       the same metric through the reference-facing C call yak_count(file) on a FASTQ file in
            host memory (tmpfs): parse + H2D + both passes + shrink + dump, D2H included
  roofline  dominant kernel of the timed region (per-kernel CUDA events inside the library)
  cpu_baseline  oracle/_ref/yak (the unmodified reference) `count` on the same sample file

`--impl reference` times only the reference's CPU implementation (oracle/_ref/yak, else the port).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout must carry exactly ONE JSON line, but libraries print there too (NCCL's version banner, for one):
# keep a private handle on the real stdout for the result and point fd 1 at stderr for everybody else.
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit_result(line: dict) -> None:
    _RESULT_OUT.write(json.dumps(line) + "\n")
    _RESULT_OUT.flush()


K, PRE, BF, NH, L = 31, 12, 37, 4, 150
SEED_G, SEED_R = 20260925, 7
ERR, NPCT = 0.005, 1
# SURVEY 8(d) algorithmic bytes per k-mer event
ALG = {"extract": 0.31 + 8.0, "insert": 24.0, "bloom": 128.0, "pass1_bloom": 160.3, "plain": 32.3, "lookup": 8.31}


def shm_dir(need_bytes: int = 4 << 30):
    """a directory in memory for the e2e sample file (and the reference's output), with room for it"""
    for d in ("/dev/shm", tempfile.gettempdir()):
        try:
            if os.path.isdir(d) and os.access(d, os.W_OK):
                st = os.statvfs(d)
                if st.f_bavail * st.f_frsize >= need_bytes:
                    return d
        except OSError:
            pass
    return tempfile.gettempdir()


class ClockSampler:
    """nvidia-smi clocks/throttle reasons while the timed region runs (B200_PROFILING.md)."""

    def __init__(self, gpu: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:  # noqa: BLE001
                self.proc.kill()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback"


def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "yak")
    return p if os.path.exists(p) else None


def cpu_reference_run(fn: str, n_events: int, threads: int, bf: int):
    """`yak count` of the unmodified reference on the host cores; returns (events/s, seconds, kind)."""
    out = os.path.join(shm_dir(), f"yakb_ref_{os.getpid()}.yak")
    ref = ref_binary()
    t0 = time.time()
    if ref:
        cmd = [ref, "count", f"-k{K}", f"-p{PRE}", f"-t{threads}", f"-H{NH}", "-o", out]
        if bf > 0:
            cmd.append(f"-b{bf}")
        subprocess.run(cmd + [fn], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        kind = "reference"
    else:  # the oracle port (single thread)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        h, _ = oracle_lib.count_file(fn, k=K, pre=PRE, bf_shift=bf, bf_n_hash=NH)
        oracle_lib.lib().yo_ch_dump(h, out.encode())
        oracle_lib.lib().yo_ch_destroy(h)
        kind, threads = "port", 1
    dt = time.time() - t0
    try:
        os.unlink(out)
    except OSError:
        pass
    return n_events / dt, dt, kind, threads


def make_sample_file(torch, lib, genome2, G, n_reads: int, first: int, path: str):
    """FASTQ sample of the workload written from the device generator; returns #k-mer events."""
    rec = 2 * L + 7
    buf = torch.empty(n_reads * rec, dtype=torch.uint8, device="cuda")
    lib.yakb_synth_reads_dev(genome2.data_ptr(), G, SEED_R, first, n_reads, L, ERR, NPCT, 2, buf.data_ptr(),
                             torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    host = buf.cpu().numpy()
    host.tofile(path)
    # events = windows of k valid bases: count from the sequence lines
    import numpy as np
    seq = host.reshape(n_reads, rec)[:, 3:3 + L]
    valid = (seq != ord("N")).astype(np.int32)
    c = np.cumsum(valid, axis=1)
    win = c[:, K - 1:] - np.concatenate([np.zeros((n_reads, 1), np.int32), c[:, :L - K]], axis=1)
    return int((win == K).sum())


def run_reference_arm(args):
    """--impl reference: the UNMODIFIED reference (oracle/_ref/yak count) on the host cores, same metric.

    One run of `yak count -k31 -p12 -b<bf> -K<step bases> -t<cores>` over (W+K) batches of the cfg2 read
    stream; the reference prints one progress line per batch with its wall clock (count.c:140), so the
    time of step i is the difference of consecutive pass-1 lines.  The batches are a bounded sample
    (at most ~12 M reads in total) so the whole run ends within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import re
    import numpy as np
    W, KS = args.warmup, args.steps
    total_reads = int(min(12_000_000, (W + KS) * args.chunk_reads))
    per = max(2000, total_reads // (W + KS))
    n_reads = per * (W + KS)
    G = args.genome
    fn = os.path.join(shm_dir(), f"yakb_refarm_{os.getpid()}.fq")
    n_ev = None
    try:  # the device generator (not on the timed path) writes the sample when a GPU is present
        import torch
        from yak_b200 import capi
        if capi.lib().yakb_device_count() > 0:
            lib = capi.lib()
            genome2 = torch.empty((G + 31) // 32 + 1, dtype=torch.int64, device="cuda")
            lib.yakb_synth_genome_dev(SEED_G, G, genome2.data_ptr(), torch.cuda.current_stream().cuda_stream)
            n_ev = make_sample_file(torch, lib, genome2, G, n_reads, 0, fn)
            del genome2
            torch.cuda.empty_cache()
    except Exception:  # noqa: BLE001
        n_ev = None
    if n_ev is None:  # no GPU: numpy generator on a window of the genome (same read model)
        from yak_b200 import synth
        G = min(G, 50_000_000)
        genome = synth.genome_codes(SEED_G, G)
        with open(fn, "wb") as f:
            f.write(synth.reads_file_bytes(SEED_G, G, SEED_R, n_reads, L, ERR, NPCT, fastq=True, genome=genome))
        n_ev = synth.count_events(synth.read_codes(SEED_G, G, SEED_R, 0, n_reads, L, ERR, NPCT, genome), K)
    threads = os.cpu_count() or 1
    bf = args.bf_shift
    ref = ref_binary()
    out = os.path.join(shm_dir(), f"yakb_refarm_{os.getpid()}.yak")
    kind = "reference"
    if ref:
        cmd = [ref, "count", f"-k{K}", f"-p{PRE}", f"-t{threads}", f"-H{NH}", f"-K{per * L}", "-o", out]
        if bf > 0:
            cmd.append(f"-b{bf}")
        t0 = time.time()
        r = subprocess.run(cmd + [fn], check=True, capture_output=True, text=True)
        wall = time.time() - t0
        stamps = [float(m.group(1)) for m in re.finditer(r"\[M::worker_pipeline::([0-9.]+)\*", r.stderr)]
        stamps = stamps[:W + KS]  # pass 1 (the second pass prints the same number of lines again)
        if len(stamps) == W + KS and KS > 0:
            t_start = stamps[W - 1] if W > 0 else 0.0
            ms_total = (stamps[-1] - t_start) * 1000.0
        else:  # fewer lines than expected (reads shorter than a batch): fall back to the whole pass
            ms_total = wall * 1000.0 * KS / (2 * (W + KS) if bf > 0 else (W + KS))
    else:  # the oracle port, single thread, whole run
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        t0 = time.time()
        h = oracle_lib.lib().yo_count_file(fn.encode(), K, PRE, bf, NH, None, None)
        ms_total = (time.time() - t0) * 1000.0 * KS / (W + KS)
        oracle_lib.lib().yo_ch_destroy(h)
        kind, threads = "port", 1
    for p in (fn, out):
        try:
            os.unlink(p)
        except OSError:
            pass
    ev_per_step = n_ev / (W + KS)
    val = ev_per_step * KS / (ms_total / 1000.0)
    line = {"impl": "reference", "metric": f"k-mer events/s (k={K}, pass 1 of `yak count -b{args.bf_shift}`, chunk steps)", "value": val,
            "unit": "events/s", "n_gpus": args.gpus, "steps": KS, "warmup": W, "ms_per_step": ms_total / KS,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "cfg2", "k": K, "pre": PRE, "bf_shift": bf, "bf_n_hash": NH, "read_len": L, "genome_bp": G,
                       "reads_per_step": per, "sample": f"{n_reads} reads in {W + KS} batches of {per}"},
            "cpu_baseline": {"value": val, "unit": "events/s", "cores": threads, "kind": kind,
                             "sample": f"yak count -k{K} -p{PRE} -b{bf} -t{threads} -K{per * L}: pass-1 batches {W + 1}..{W + KS} of {per} reads ({int(ev_per_step)} events each), timed from the reference's own progress lines"},
            "e2e": {"value": val, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_result(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None,
                    help="timed steps; default = the whole first pass of cfg2 (600 M reads): 300 / gpus batches per rank, minus the warm-up")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome", type=int, default=3_000_000_000)
    ap.add_argument("--k", type=int, default=K, help="k-mer length (configs[4] sweeps 21/31/47/63)")
    ap.add_argument("--chunk-reads", type=int, default=2_000_000)
    ap.add_argument("--bf-shift", type=int, default=BF)
    ap.add_argument("--e2e-reads", type=int, default=4_000_000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    globals()["K"] = args.k
    if args.steps is None:   # N ranks consume N batches per step: the default job stays cfg2's 600 M reads at every N
        world_env = max(1, int(os.environ.get("WORLD_SIZE", str(args.gpus))))
        args.steps = max(1, 300 // world_env - args.warmup)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from yak_b200 import capi
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    capi.require_gpu()
    lib = capi.lib()
    G = args.genome
    W, KS = args.warmup, args.steps
    nr = args.chunk_reads
    rec = L + 1
    genome2 = torch.empty((G + 31) // 32 + 1, dtype=torch.int64, device="cuda")
    cur = torch.cuda.current_stream().cuda_stream
    lib.yakb_synth_genome_dev(SEED_G, G, genome2.data_ptr(), cur)
    # every step takes the next batch of the read stream; with N ranks, step i is the N consecutive
    # slices i*N .. i*N+N-1, one per rank (weak scaling: per-GPU work fixed).  The batch is generated
    # on the device before its step and is not part of the timed region.
    buf = torch.empty(nr * rec, dtype=torch.uint8, device="cuda")

    def make_input(i):
        first = (i * world + rank) * nr
        lib.yakb_synth_reads_dev(genome2.data_ptr(), G, SEED_R, first, nr, L, ERR, NPCT, 0, buf.data_ptr(), cur)
        torch.cuda.synchronize()

    stats = (C.c_uint64 * 4)()
    ev_total = 0
    per_step = []
    if world == 1:
        h = lib.yak_ch_init(K, PRE, NH, args.bf_shift)
        assert h, "yak_ch_init failed"
        stream = torch.cuda.ExternalStream(lib.yakb_ch_stream(h))

        def step(i):
            rc = lib.yakb_count_ascii_dev(h, buf.data_ptr(), nr * rec, 1, stats)
            assert rc == 0
            return list(stats)
    else:
        # sub-tables sharded over the ranks, one all-to-all per step (yak_b200/dist.py)
        from yak_b200 import dist as ydist
        be = ydist.GpuBackend(K, PRE, args.bf_shift, NH, rank, world)
        sc = ydist.ShardedCounter(be)
        h = be.h
        stream = torch.cuda.current_stream()

        def step(i):
            n = sc.count_chunk(buf, 1)
            return [n, int(be.stats[1]), int(be.stats[2]), int(be.stats[3])]
    for i in range(W):
        make_input(i)
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = lib.yakb_kernel_launches()
    lib.yakb_prof_enable(1)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms = 0.0
    for i in range(W, W + KS):
        make_input(i)                                   # untimed
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        st = step(i)                                    # exactly one step between the two events
        e1.record(stream)
        torch.cuda.synchronize()
        dt = e0.elapsed_time(e1)
        ms += dt
        ev_total += st[0]
        per_step.append(st + [round(dt, 3)])
    clk = clocks.stop() if rank == 0 else None
    launches = lib.yakb_kernel_launches() - launches0
    pj = C.create_string_buffer(1 << 16)
    lib.yakb_prof_json(pj, 1 << 16)
    prof = json.loads(pj.value.decode())
    lib.yakb_prof_enable(0)
    if world > 1:
        t = torch.tensor([ms, float(ev_total)], dtype=torch.float64, device="cuda")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, ev_all = float(tmax[0]), float(tsum[1])
    else:
        ev_all = float(ev_total)
    value = ev_all / (ms / 1000.0)
    dev_bytes = lib.yakb_ch_device_bytes(h)
    if world == 1:
        lib.yak_ch_destroy(h)
    else:
        be.close()
    lib.yakb_device_cache_trim()   # the library keeps freed device blocks for its next table; the legs below start clean
    del buf
    torch.cuda.empty_cache()
    e2e_multi = None
    if world > 1 and not args.no_e2e:
        # e2e at N GPUs: the same pass-1 metric from a FASTQ file on the host through the sharded file path
        # (yak_b200/dist.py count_file_sharded: rank 0's parser pool fills a shared-memory staging buffer, every rank takes
        # its contiguous part of every batch: H2D, extraction, one NCCL all-to-all per batch, count on its shard)
        from yak_b200 import dist as ydist
        fn = os.path.join(shm_dir(), f"yakb_bench_e2e_{os.environ.get('MASTER_PORT', '0')}.fq")
        nev = torch.zeros(1, dtype=torch.int64, device="cuda")
        if rank == 0:
            nev[0] = make_sample_file(torch, lib, genome2, G, args.e2e_reads, 0, fn)
        dist.broadcast(nev, 0)
        batch = min(64 << 20, (512 << 20) // world) * world
        dt = 0.0
        for rep in range(2):            # warm-up (pinned buffers, page cache), then the timed run
            be2 = ydist.GpuBackend(K, PRE, args.bf_shift, NH, rank, world)
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.time()
            sc2 = ydist.count_file_sharded(fn, be2, k=K, batch_bases=batch)
            tot2 = sc2.total_distinct()  # the result every rank reads back
            torch.cuda.synchronize()
            dt = time.time() - t0
            be2.close()
        tmax = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.barrier()
        if rank == 0:
            os.unlink(fn)
            e2e_multi = {"value": int(nev[0]) / float(tmax[0]), "unit": "events/s", "n_gpus_used": world,
                         "h2d_bytes_per_step": args.e2e_reads * (L + 1), "d2h_bytes_per_step": 8 * world, "seconds": float(tmax[0]),
                         "distinct_after_pass1": tot2,
                         "what": f"pass 1 (-b{args.bf_shift}) of {args.e2e_reads} FASTQ reads in tmpfs on {world} GPUs: parse + H2D + extract + all-to-all + count, {int(nev[0])} events (max over ranks)"}
        lib.yakb_device_cache_trim()
    if world > 1:  # every rank leaves the process group together; the legs below are rank 0's alone
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    # ---- roofline of the dominant kernel (live CUDA-event times from inside the library)
    peak, peak_kind = measured_peaks()
    n_pending = sum(s[1] for s in per_step)
    alg_bytes = {  # algorithmic bytes per launch set = SURVEY 8(d) per-event figure x events the kernel processed
        "k1_fused": (0.31 + 16.0) * ev_total,           # read 2-bit bases; slot read + counter write per event
        "k1_array": (8.0 + 16.0) * ev_total,            # read the routed event; slot read + counter write
        "group_insert": (8.0 + 128.0 + 16.0) * n_pending,  # sorted event + bloom block RMW + slot read/write
        "group_sort": 4 * 2 * 12.0 * n_pending,      # 4 passes, read + write of a 12-byte record
        "compact": 8.0 * n_pending,
        "pack_ascii": 1.375 * nr * rec * KS,
    }
    # DRAM bytes per launch from the committed `ncu --set full` captures (dram__bytes_read.sum + dram__bytes_write.sum):
    # k1_fused in the steady state (profiles/r01_k1_fused_steady.md), group_insert at step ~10 (profiles/r01_top_kernels_ncu.md)
    ncu_traffic = {"k1_fused": 8.218429e9 + 7.537115e9, "group_insert": 20.139495e9 + 15.937980e9}
    ncu_source = {"k1_fused": "profiles/r01_k1_fused_steady.md", "group_insert": "profiles/r01_top_kernels_ncu.md"}
    dom = max(prof.items(), key=lambda kv: kv[1][0]) if prof else (None, [0, 0])
    roof = None
    if dom[0]:
        nm, (tms, nl) = dom
        ab = alg_bytes.get(nm, 0.0)
        ach = ab / (tms / 1000.0) / 1e9 if tms > 0 else 0.0
        roof = {"bound": "hbm", "kernel": nm, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": ncu_traffic.get(nm), "traffic_source": ncu_source.get(nm),
                "peak_source": peak_kind, "launches": nl, "kernel_ms_total": tms,
                "algorithmic_bytes_per_launch": ab / max(nl, 1),
                # the binding limit of a hash-table update is the random 32-byte read-modify-write rate, not bytes:
                # measured with tools/gups.cu on this pool (profiles/r01_gups_random_access.txt, 64 GiB footprint)
                "random_rmw": {"achieved_gops": (ev_total / (tms / 1000.0) / 1e9) if nm.startswith("k1_") and tms > 0 else None,
                               "peak_gops": 14.7, "unit": "G read-modify-writes/s"},
                "share_of_step": tms / ms if ms > 0 else None}
    # the whole step against the same peak, by SURVEY 8(d)'s fixed accounting for pass 1 with a filter
    # (extract 8.31 + insert 24 + bloom block 128 = 160.3 B per k-mer event, whatever the implementation skips)
    roof_step = {"bound": "hbm", "achieved": ALG["pass1_bloom"] * ev_all / (ms / 1000.0) / 1e9 / world, "peak": peak, "unit": "GB/s",
                 "algorithmic_bytes_per_event": ALG["pass1_bloom"], "per": "GPU"} if args.bf_shift > PRE else None
    if roof_step:
        roof_step["frac"] = roof_step["achieved"] / peak
    line = {"metric": f"k-mer events/s (k={K}, pass 1 of `yak count -b{args.bf_shift}`, chunk steps)", "value": value, "unit": "events/s",
            "n_gpus": world, "steps": KS, "warmup": W, "ms_per_step": ms / KS, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "cfg2", "k": K, "pre": PRE, "bf_shift": args.bf_shift, "bf_n_hash": NH, "read_len": L,
                       "genome_bp": G, "reads_per_step": nr, "bases_per_step": nr * L, "l2": "inputs larger than L2 (no flush)",
                       "parallelism": f"{world} GPUs: reads split by rank, 2^{PRE}/{world} sub-tables per rank, one NCCL all-to-all per step" if world > 1 else "1 GPU",
                       "device_bytes": int(dev_bytes)},
            "input_gbp_per_s": (nr * L * KS * world) / (ms / 1000.0) / 1e9,
            "gpu_launches": int(launches), "kernels_ms": {k: round(v[0], 3) for k, v in prof.items()},
            "clocks": clk, "roofline": roof, "roofline_step": roof_step}

    # ---- e2e: the same metric (pass-1 events/s) through the reference-facing C call on a HOST file:
    #      yak_count(fn, opt, NULL) = parse + H2D + every kernel of the pass, result (h->tot) read back.
    #      The whole `yak count -b37` job (both passes + shrink + dump) on the same file is reported next
    #      to it together with the unmodified reference's time for that job (cpu_baseline).
    if e2e_multi:
        line["e2e"] = e2e_multi
    elif not args.no_e2e:
        fn = os.path.join(shm_dir(), f"yakb_bench_{os.getpid()}.fq")
        n_ev = make_sample_file(torch, lib, genome2, G, args.e2e_reads, 0, fn)
        out = os.path.join(shm_dir(), f"yakb_bench_{os.getpid()}.yak")
        fsz = os.path.getsize(fn)
        o = capi.copt(K, PRE, args.bf_shift, NH)
        lib.yak_ch_destroy(lib.yak_count(fn.encode(), C.byref(o), None))   # warm-up: page cache, pinned buffers, context
        t0 = time.time()
        hh = lib.yak_count(fn.encode(), C.byref(o), None)
        tot1 = int(hh.contents.tot)
        dt1 = time.time() - t0
        lib.yak_ch_destroy(hh)
        t0 = time.time()
        hh = capi.count_file(fn, k=K, pre=PRE, bf_shift=args.bf_shift, bf_n_hash=NH)
        lib.yak_ch_dump(hh, out.encode())
        dt = time.time() - t0
        osz = os.path.getsize(out)
        lib.yak_ch_destroy(hh)
        line["e2e"] = {"value": n_ev / dt1, "unit": "events/s", "n_gpus_used": 1, "h2d_bytes_per_step": args.e2e_reads * (L + 1),
                       "d2h_bytes_per_step": 8, "seconds": dt1, "distinct_after_pass1": tot1,
                       "what": f"yak_count(pass 1, -b{args.bf_shift}) of {args.e2e_reads} FASTQ reads ({fsz} B in tmpfs): parse + H2D + kernels, {n_ev} events"}
        line["e2e_full_job"] = {"value": n_ev / dt, "unit": "input events/s", "seconds": dt, "h2d_bytes": 2 * args.e2e_reads * (L + 1),
                                "d2h_bytes": osz, "what": "yak_count x2 + yak_ch_shrink + yak_ch_dump of the same file (= `yak count -b37 -o`)"}
        if not args.no_cpu:
            threads = os.cpu_count() or 1
            v, dtc, kind, threads = cpu_reference_run(fn, n_ev, threads, args.bf_shift)
            line["cpu_baseline"] = {"value": v, "unit": "input events/s", "cores": threads, "kind": kind, "seconds": dtc,
                                    "sample": f"whole job `yak count -k{K} -p{PRE} -b{args.bf_shift} -t{threads} -o` on the e2e sample file ({n_ev} events); compare with e2e_full_job"}
        for p in (fn, out):
            try:
                os.unlink(p)
            except OSError:
                pass
    if args.verbose:
        sys.stderr.write(json.dumps(per_step) + "\n")
    emit_result(line)


if __name__ == "__main__":
    main()
