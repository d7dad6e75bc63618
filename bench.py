#!/usr/bin/env python
"""bench.py - k-mer events/s of the count hot path at BASELINE.json configs[1] geometry.

Workload (config.workload = "cfg2"): k=31, -p12, -b37 -H4, 150 bp reads sampled from a 3 Gbp
uniform genome (e=0.005, 1 % reads with an N), pass 1 of `yak count` (bloom + insert), the same
synthetic stream as yak_b200/synth.py generated on the device.

A *step* is one global batch of --reads-per-step reads (default 16 M reads = 2.4 Gbp, far larger than
L2) through the whole per-chunk path (pack -> partition -> zone-by-zone probe -> ordered bloom/insert
of the pending events -> journal).  Steps are consecutive batches from the start of the job; table,
bloom and journal persist, so the default 3 + 34 steps are (all but the last 1.3 % of) the first pass
of cfg2 (600 M reads = 90 Gbp = 30x of 3 Gbp), table growth and rehash included.

STRONG scaling: every N consumes the IDENTICAL read stream.  With N ranks, rank r takes the r-th
contiguous N-th of each step's batch (16/N M reads), extracts its k-mers, one NCCL all-to-all routes
them to the owners of their sub-tables, each rank counts on its shard.  A rank's share of a step is
cut into rounds of at most --chunk-reads reads (one engine chunk / one all-to-all per round; default: the whole step).
Each step is bracketed by its own pair of CUDA events; its input is generated on the device before
it, outside the timed region (max over ranks of the summed step times).

  value     events/s with the step's ASCII bases already resident in HBM
  e2e       the same metric (pass-1 events/s) through the reference-facing C call yak_count(file) on a
            FASTQ file in host memory (tmpfs): parse + H2D + kernels, result read back.  At N > 1:
            yak_b200.dist.count_file_sharded (the multi-GPU form of the same call)
  parity    sha256 of the .yak file of the whole two-pass job on the e2e sample (N GPUs) against the
            unmodified reference's (oracle/_ref/yak count -b37 -o) on the same file; a mismatch exits 1
  roofline  dominant kernel of the timed region: SURVEY 8(d) algorithmic bytes per unit x the units the
            kernel processed (counted inside the library) / its CUDA-event time, against the measured peak
  cpu_baseline  the unmodified reference's whole job on the e2e sample file, pass-1 rate from its own
            progress lines

`--impl reference` times only the reference's CPU implementation (oracle/_ref/yak, else the port).
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import re
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
# SURVEY 8(d) algorithmic bytes per unit (event / pending event / position), independent of implementation tricks:
#   E extract+partition: 0.31 B of 2-bit bases read + 8 B hashed k-mer written;  I insert/update: 8 B partition
#   buffer + 8 B slot read + 8 B slot written;  bloom: one 64-B block read + written
ALG_UNIT = {
    "pack_ascii": 1.375,                 # per position: 1 B ASCII read, 0.25 + 0.125 B written
    "part_scatter": 0.31 + 8.0,          # E
    "zone_probe": 24.0,                  # I
    "k1_fused": 0.31 + 16.0,             # E + I without a materialised partition
    "k1_array": 8.0 + 16.0,              # I on routed events
    "compact": 8.0,                      # per pending event: its hash written in file order
    "group_sort": 24.0,                  # per pending event: one read + one write of a 12-byte record (the algorithmic minimum; we make 4 passes)
    "group_insert": 8.0 + 128.0 + 16.0,  # per pending event: sorted record + bloom block RMW + slot read/write
    "post_pending": 8.0,
}
ALG_STEP = {"pass1_bloom": 160.3, "plain": 32.3}


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
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per unit from this round's committed `ncu --set full` captures (profiles/r02_ncu_traffic.json, written
    by tools/ncu_summary.py from the .ncu-rep of the same bench command): kernel -> {dram_bytes_per_unit, source, commit}"""
    p = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    try:
        return json.load(open(p))
    except Exception:  # noqa: BLE001
        return {}


def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "yak")
    return p if os.path.exists(p) else None


def sha256_file(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


PROGRESS_RE = re.compile(r"\[M::worker_pipeline::([0-9.]+)\*")


def cpu_reference_job(fn: str, n_events: int, threads: int, bf: int, out: str, batch_bases: int | None = None):
    """the whole `yak count` job of the unmodified reference on the host cores (output kept for the byte comparison);
    returns dict(seconds, pass1_seconds from its own progress lines, kind, threads)"""
    ref = ref_binary()
    t0 = time.time()
    p1 = None
    if ref:
        cmd = [ref, "count", f"-k{K}", f"-p{PRE}", f"-t{threads}", f"-H{NH}", "-o", out]
        if batch_bases:
            cmd.append(f"-K{batch_bases}")
        if bf > 0:
            cmd.append(f"-b{bf}")
        r = subprocess.run(cmd + [fn], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        stamps = [float(m.group(1)) for m in PROGRESS_RE.finditer(r.stderr)]
        if stamps:   # pass 1 = the first half of the progress lines when a second pass follows (main.c:57)
            n1 = len(stamps) // 2 if bf > 0 else len(stamps)
            p1 = stamps[max(n1, 1) - 1]
        kind = "reference"
    else:  # the oracle port (single thread)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        h, _ = oracle_lib.count_file(fn, k=K, pre=PRE, bf_shift=bf, bf_n_hash=NH)
        oracle_lib.lib().yo_ch_dump(h, out.encode())
        oracle_lib.lib().yo_ch_destroy(h)
        kind, threads = "port", 1
    dt = time.time() - t0
    return {"seconds": dt, "pass1_seconds": p1, "kind": kind, "threads": threads}


def count_events_dev(torch, buf, n_reads: int, rec: int, seq_off: int) -> int:
    """k-mer events (windows of K bases without N) of n_reads fixed-size records on the device"""
    tot = 0
    view = buf[: n_reads * rec].view(n_reads, rec)
    for s in range(0, n_reads, 1 << 20):
        seq = view[s:s + (1 << 20), seq_off:seq_off + L]
        c = torch.cumsum((seq != ord("N")).to(torch.int32), dim=1)
        win = c[:, K - 1:].clone()
        win[:, 1:] -= c[:, :L - K]
        tot += int((win == K).sum())
    return tot


def make_sample_file(torch, lib, genome2, G, n_reads: int, first: int, path: str) -> int:
    """FASTQ sample of the workload written from the device generator in slices; returns #k-mer events."""
    rec = 2 * L + 7
    per = 4_000_000
    buf = torch.empty(min(per, n_reads) * rec, dtype=torch.uint8, device="cuda")
    n_ev = 0
    with open(path, "wb") as f:
        for s in range(0, n_reads, per):
            m = min(per, n_reads - s)
            lib.yakb_synth_reads_dev(genome2.data_ptr(), G, SEED_R, first + s, m, L, ERR, NPCT, 2, buf.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            n_ev += count_events_dev(torch, buf, m, rec, 3)
            buf[: m * rec].cpu().numpy().tofile(f)
    del buf
    return n_ev


def base_config(args, world):
    nr = args.reads_per_step
    return {"workload": "cfg2", "k": K, "pre": PRE, "bf_shift": args.bf_shift, "bf_n_hash": NH, "read_len": L, "genome_bp": args.genome,
            "reads_per_step": nr, "bases_per_step": nr * L, "l2": "inputs larger than L2 (no flush)"}


def run_reference_arm(args):
    """--impl reference: the UNMODIFIED reference (oracle/_ref/yak count) on the host cores, same metric and config.

    One run of `yak count -k31 -p12 -b<bf> -K<sample bases> -t<cores>` over (W+K) batches of the cfg2 read stream; each
    batch is a BOUNDED SAMPLE of the step's reads (the first `per` reads of every step's 16 M, `per` chosen so that the whole
    run ends within a few minutes); the reference prints one progress line per batch with its wall clock (count.c:140), so
    the time of step i is the difference of consecutive pass-1 lines.  The input is written by oracle/_bin/synthgen (the
    CPU twin of the device generator): this process never loads the product library."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W, KS = args.warmup, args.steps
    total_reads = int(min(12_000_000, (W + KS) * args.reads_per_step))
    per = max(2000, total_reads // (W + KS))
    n_reads = per * (W + KS)
    G = args.genome
    fn = os.path.join(shm_dir(), f"yakb_refarm_{os.getpid()}.fq")
    threads = os.cpu_count() or 1
    gen = os.path.join(ROOT, "oracle", "_bin", "synthgen")
    if not os.path.exists(gen):      # normally built by __graft_entry__.build(); gcc is all it needs
        os.makedirs(os.path.dirname(gen), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-o", gen, os.path.join(ROOT, "oracle", "synthgen.c"), "-lpthread"], check=True)
    # batch i of the sample = the first `per` reads of step i of the workload: one generator call per batch, concatenated
    n_ev = 0
    with open(fn, "wb") as out_f:
        for i in range(W + KS):
            part = fn + ".part"
            r = subprocess.run([gen, str(SEED_G), str(G), str(SEED_R), str(i * args.reads_per_step), str(per), str(L), str(ERR), str(NPCT),
                                "2", str(K), part, str(threads)], check=True, capture_output=True, text=True)
            n_ev += int(r.stdout.strip())
            with open(part, "rb") as pf:
                while True:
                    b = pf.read(1 << 24)
                    if not b:
                        break
                    out_f.write(b)
            os.unlink(part)
    bf = args.bf_shift
    ref = ref_binary()
    out = os.path.join(shm_dir(), f"yakb_refarm_{os.getpid()}.yak")
    kind = "reference"
    best = None
    tried = []
    # report the best of several -t (steps 0/1 of the reference's pipeline are single-threaded, -t saturates early)
    tlist = sorted({threads, max(1, threads // 2), min(threads, 8)}, reverse=True) if ref else [1]
    for t in tlist:
        if ref:
            cmd = [ref, "count", f"-k{K}", f"-p{PRE}", f"-t{t}", f"-H{NH}", f"-K{per * L}", "-o", out]
            if bf > 0:
                cmd.append(f"-b{bf}")
            t0 = time.time()
            # pass 1 is what is timed: the run is stopped once its W+K progress lines are out (the second pass would print
            # the same number of lines again and double the wall time of this arm for nothing)
            pr = subprocess.Popen(cmd + [fn], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
            stamps = []
            for ln in pr.stderr:
                m = PROGRESS_RE.search(ln)
                if m:
                    stamps.append(float(m.group(1)))
                    if len(stamps) >= W + KS:
                        break
            pr.terminate()
            try:
                pr.wait(timeout=20)
            except Exception:  # noqa: BLE001
                pr.kill()
            pr.stderr.close()
            wall = time.time() - t0
            if len(stamps) == W + KS and KS > 0:
                t_start = stamps[W - 1] if W > 0 else 0.0
                ms_total = (stamps[-1] - t_start) * 1000.0
            else:  # fewer lines than expected: fall back to the whole pass
                ms_total = wall * 1000.0 * KS / (W + KS)
        else:  # the oracle port, single thread, whole run
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib
            t0 = time.time()
            h = oracle_lib.lib().yo_count_file(fn.encode(), K, PRE, bf, NH, None, None)
            ms_total = (time.time() - t0) * 1000.0 * KS / (W + KS)
            oracle_lib.lib().yo_ch_destroy(h)
            kind = "port"
        tried.append({"threads": t, "ms_total": ms_total})
        if best is None or ms_total < best[1]:
            best = (t, ms_total)
    for p in (fn, out):
        try:
            os.unlink(p)
        except OSError:
            pass
    threads, ms_total = best
    ev_per_step = n_ev / (W + KS)
    val = ev_per_step * KS / (ms_total / 1000.0)
    line = {"impl": "reference", "metric": f"k-mer events/s (k={K}, pass 1 of `yak count -b{args.bf_shift}`, chunk steps)", "value": val,
            "unit": "events/s", "n_gpus": args.gpus, "steps": KS, "warmup": W, "ms_per_step": ms_total / KS,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": base_config(args, 1),
            "sample": {"reads_per_step_sampled": per, "reads_total": n_reads, "events_per_step_sampled": int(ev_per_step),
                       "what": f"each of the {W + KS} steps is the first {per} reads of the step's {args.reads_per_step}; ms_per_step is per sampled step"},
            "cpu_baseline": {"value": val, "unit": "events/s", "cores": threads, "kind": kind, "threads_tried": tried,
                             "sample": f"yak count -k{K} -p{PRE} -b{bf} -t{threads} -K{per * L}: pass-1 batches {W + 1}..{W + KS} of {per} reads ({int(ev_per_step)} events each), timed from the reference's own progress lines; best of -t {tlist}"},
            "e2e": {"value": val, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_result(line)


def roofline_objects(prof, ms, peak, peak_kind):
    """per-kernel achieved bandwidth by SURVEY 8(d) accounting; the dominant KERNEL (largest time among those with a byte
    model) is `roofline`, the partition+insert pair north_star names is reported next to it"""
    traffic = ncu_traffic()
    per = {}
    for nm, ent in prof.items():
        tms, nl = ent[0], ent[1]
        units = ent[2] if len(ent) > 2 else 0
        if nm not in ALG_UNIT or tms <= 0:
            continue
        ab = ALG_UNIT[nm] * units
        ach = ab / (tms / 1000.0) / 1e9
        tr = traffic.get(nm, {})
        # a table update is one random 32-byte sector read-modify-write: what binds such kernels is how many sector accesses per
        # second DRAM serves, not bytes (tools/gups.cu on this pool: 37.7 G random sector loads/s, 2 x 15.8 G for load+store pairs
        # spread over 32 GiB, profiles/r01_gups_random_access.txt); reported next to the byte roofline when ncu traffic is known
        acc = None
        if tr.get("dram_bytes_per_unit"):
            acc = {"achieved_G_sectors_per_s": tr["dram_bytes_per_unit"] * units / 32.0 / (tms / 1000.0) / 1e9,
                   "unconfined_random_G_sectors_per_s": 37.7, "source": "ncu DRAM bytes / 32 B over the live kernel time; peak: profiles/r01_gups_random_access.txt"}
        per[nm] = {"kernel": nm, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "dram_access_rate": acc,
                   "traffic": (tr.get("dram_bytes_per_unit") * units / max(nl, 1)) if tr.get("dram_bytes_per_unit") else None,
                   "traffic_source": tr.get("source"), "peak_source": peak_kind, "launches": nl, "kernel_ms_total": tms,
                   "units": units, "algorithmic_bytes_per_unit": ALG_UNIT[nm], "algorithmic_bytes_per_launch": ab / max(nl, 1),
                   "share_of_step": tms / ms if ms > 0 else None}
    dom = max(per.values(), key=lambda d: d["kernel_ms_total"]) if per else None
    pi = None
    if "part_scatter" in per and "zone_probe" in per:
        t = per["part_scatter"]["kernel_ms_total"] + per["zone_probe"]["kernel_ms_total"]
        ab = per["part_scatter"]["algorithmic_bytes_per_unit"] * per["part_scatter"]["units"] + per["zone_probe"]["algorithmic_bytes_per_unit"] * per["zone_probe"]["units"]
        ach = ab / (t / 1000.0) / 1e9
        pi = {"kernels": ["part_scatter", "zone_probe"], "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
              "kernel_ms_total": t, "events_per_s": per["zone_probe"]["units"] / (t / 1000.0), "share_of_step": t / ms if ms > 0 else None}
    return dom, per, pi


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None,
                    help="timed steps; default = the first pass of cfg2 (600 M reads): 37 steps of 16 M reads minus the warm-up")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome", type=int, default=3_000_000_000)
    ap.add_argument("--k", type=int, default=K, help="k-mer length (configs[4] sweeps 21/31/47/63)")
    ap.add_argument("--reads-per-step", type=int, default=16_000_000, help="the global batch of one step (all ranks together)")
    ap.add_argument("--chunk-reads", type=int, default=16_000_000,
                    help="reads per engine chunk / all-to-all round on one rank (16 M reads = 2.4 G positions: one chunk per step at N = 1; "
                         "12.6 vs 11.9 G events/s for chunks of 8 M, profiles/r02_chunk_size.md)")
    ap.add_argument("--min-rounds", type=int, default=1,
                    help="N > 1: rounds per step at least; with 2 or more, extraction + exchange of a round run behind the count of the one before "
                         "(ShardedCounter.count_rounds) - measured at N = 2: 17.30 vs 17.41 G events/s, the count saturates the memory system and the "
                         "overlapped extraction just takes longer (profiles/r02_bench_2gpu.md), so the default is one round")
    ap.add_argument("--bf-shift", type=int, default=BF)
    ap.add_argument("--e2e-reads", type=int, default=32_000_000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the reference run on the e2e sample (no cpu_baseline, no parity)")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    globals()["K"] = args.k
    if args.steps is None:
        args.steps = max(1, 600_000_000 // args.reads_per_step - args.warmup)
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
    R = args.reads_per_step
    assert R % world == 0, "--reads-per-step must be a multiple of the number of ranks"
    mine = R // world                                   # this rank's reads of every step (strong scaling)
    rounds = max(1, -(-mine // args.chunk_reads))
    if world > 1:
        rounds = max(rounds, args.min_rounds)
    per_round = -(-mine // rounds)
    rec = L + 1
    genome2 = torch.empty((G + 31) // 32 + 1, dtype=torch.int64, device="cuda")
    cur = torch.cuda.current_stream().cuda_stream
    lib.yakb_synth_genome_dev(SEED_G, G, genome2.data_ptr(), cur)
    # step i = reads [i*R, (i+1)*R) of the stream at EVERY N; rank r takes [i*R + r*mine, +mine).  The rank's share is
    # generated on the device before the step and is not part of the timed region.
    buf = torch.empty(mine * rec, dtype=torch.uint8, device="cuda")

    def make_input(i):
        lib.yakb_synth_reads_dev(genome2.data_ptr(), G, SEED_R, i * R + rank * mine, mine, L, ERR, NPCT, 0, buf.data_ptr(), cur)
        torch.cuda.synchronize()

    stats = (C.c_uint64 * 4)()
    ev_total = 0
    pend_total = 0
    per_step = []
    if world == 1:
        h = lib.yak_ch_init(K, PRE, NH, args.bf_shift)
        assert h, "yak_ch_init failed"
        stream = torch.cuda.ExternalStream(lib.yakb_ch_stream(h))

        def step(i):
            tot = [0, 0, 0, 0]
            for c in range(rounds):
                a, b = c * per_round, min(mine, (c + 1) * per_round)
                rc = lib.yakb_count_ascii_dev(h, buf.data_ptr() + a * rec, (b - a) * rec, 1, stats)
                assert rc == 0
                tot = [x + int(y) for x, y in zip(tot, stats)]
            return tot
    else:
        # sub-tables sharded over the ranks, one all-to-all per round (yak_b200/dist.py)
        from yak_b200 import dist as ydist
        be = ydist.GpuBackend(K, PRE, args.bf_shift, NH, rank, world)
        sc = ydist.ShardedCounter(be)
        h = be.h
        stream = torch.cuda.current_stream()

        class _Acc:          # the backend's per-chunk stats, summed over the rounds of a step
            def __init__(self, inner): self.inner, self.tot = inner, [0, 0, 0, 0]
            def __getattr__(self, k): return getattr(self.inner, k)
            def count_events(self, ev, create_new):
                n = self.inner.count_events(ev, create_new)
                self.tot = [x + int(y) for x, y in zip(self.tot, self.inner.stats)]
                return n
        acc = _Acc(be)
        sc.b = acc

        def step(i):
            acc.tot = [0, 0, 0, 0]
            sc.count_rounds([buf[c * per_round * rec:min(mine, (c + 1) * per_round) * rec] for c in range(rounds)], 1)
            return list(acc.tot)
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
        pend_total += st[1]
        per_step.append(st + [round(dt, 3)])
    clk = clocks.stop() if rank == 0 else None
    launches = lib.yakb_kernel_launches() - launches0
    pj = C.create_string_buffer(1 << 16)
    lib.yakb_prof_json(pj, 1 << 16)
    prof = json.loads(pj.value.decode())
    lib.yakb_prof_enable(0)
    if world > 1:
        t = torch.tensor([ms, float(ev_total), float(pend_total)], dtype=torch.float64, device="cuda")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, ev_all, pend_all = float(tmax[0]), float(tsum[1]), float(tsum[2])
    else:
        ev_all, pend_all = float(ev_total), float(pend_total)
    value = ev_all / (ms / 1000.0)
    exchange = None
    if world > 1:   # rank 0's payload all-to-alls (device time between CUDA events around the NCCL call; includes waiting for peers)
        exchange = {"a2a_ms_total": round(sc.a2a_ms(), 3), "a2a_bytes_sent": int(sc.a2a_bytes), "rounds": (W + KS) * rounds,
                    "note": "whole run incl. warm-up steps; bytes = 8 B per event routed to another rank"}
    dev_bytes = lib.yakb_ch_device_bytes(h)
    if world == 1:
        lib.yak_ch_destroy(h)
    else:
        be.close()
    lib.yakb_device_cache_trim()   # the library keeps freed device blocks for its next table; the legs below start clean
    del buf
    torch.cuda.empty_cache()

    # ---- e2e + parity on a FASTQ sample in host memory
    e2e = e2e_full = cpu_base = parity = None
    sample = os.path.join(shm_dir(args.e2e_reads * (2 * L + 7) * 2), f"yakb_bench_e2e_{os.environ.get('MASTER_PORT', str(os.getpid()))}.fq")
    out_gpu, out_ref = sample + ".gpu.yak", sample + ".ref.yak"
    if not args.no_e2e:
        nev = torch.zeros(1, dtype=torch.int64, device="cuda")
        if rank == 0:
            nev[0] = make_sample_file(torch, lib, genome2, G, args.e2e_reads, 0, sample)
        if world > 1:
            dist.broadcast(nev, 0)
        n_ev = int(nev[0])
        del genome2
        torch.cuda.empty_cache()
        if world == 1:
            fsz = os.path.getsize(sample)
            o = capi.copt(K, PRE, args.bf_shift, NH)
            lib.yak_ch_destroy(lib.yak_count(sample.encode(), C.byref(o), None))   # warm-up: page cache, pinned buffers, context
            t0 = time.time()
            hh = lib.yak_count(sample.encode(), C.byref(o), None)
            tot1 = int(hh.contents.tot)
            dt1 = time.time() - t0
            lib.yak_ch_destroy(hh)
            t0 = time.time()
            hh = capi.count_file(sample, k=K, pre=PRE, bf_shift=args.bf_shift, bf_n_hash=NH)
            lib.yak_ch_dump(hh, out_gpu.encode())
            dt = time.time() - t0
            lib.yak_ch_destroy(hh)
            what1 = f"yak_count(pass 1, -b{args.bf_shift}) of {args.e2e_reads} FASTQ reads ({fsz} B in tmpfs): parse + H2D + kernels, {n_ev} events"
        else:
            from yak_b200 import dist as ydist
            batch = max(256 << 20, (64 << 20) * world)   # bases per global batch: several batches, so copies and kernels overlap
            dt1 = dt = 0.0
            tot1 = 0
            for rep in range(2):            # warm-up (pinned buffers, page cache), then the timed run
                be2 = ydist.GpuBackend(K, PRE, args.bf_shift, NH, rank, world)
                dist.barrier()
                torch.cuda.synchronize()
                t0 = time.time()
                tm = {}
                sc2 = ydist.count_file_sharded(sample, be2, k=K, batch_bases=batch, two_pass=args.bf_shift > 0 and rep == 1, timings=tm)
                if rep == 1:
                    sc2.dump_file(out_gpu)
                torch.cuda.synchronize()
                dt = time.time() - t0
                dt1, tot1 = tm.get("pass1_seconds", dt), tm.get("distinct_after_pass1", 0)
                be2.close()
            tmax = torch.tensor([dt1, dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dt1, dt = float(tmax[0]), float(tmax[1])
            what1 = f"pass 1 (-b{args.bf_shift}) of {args.e2e_reads} FASTQ reads in tmpfs on {world} GPUs: parse + H2D + extract + all-to-all + count, {n_ev} events (max over ranks)"
        # the text itself crosses PCIe (device-side ingest of files >= 256 MB); with YAKB_GPU_INGEST=0 it would be the L+1 bytes per read the host parser keeps
        text_bytes = args.e2e_reads * (2 * L + 7)
        h2d = text_bytes if text_bytes >= (256 << 20) and os.environ.get("YAKB_GPU_INGEST", "") != "0" else args.e2e_reads * (L + 1)
        e2e = {"value": n_ev / dt1, "unit": "events/s", "n_gpus_used": world, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 8 * world, "seconds": dt1, "distinct_after_pass1": tot1, "events": n_ev, "what": what1}
        e2e_full = {"value": n_ev / dt, "unit": "input events/s", "seconds": dt, "h2d_bytes": 2 * h2d,
                    "what": f"the whole `yak count -b{args.bf_shift} -o` job on the same file on {world} GPU(s): both passes + shrink + dump"}
    lib.yakb_device_cache_trim()
    if world > 1:  # every rank leaves the process group together; the legs below are rank 0's alone
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    cj = None
    if not args.no_e2e and not args.no_cpu:
        threads = os.cpu_count() or 1
        try:
            cj = cpu_reference_job(sample, n_ev, threads, args.bf_shift, out_ref)
        except Exception as e:  # noqa: BLE001 - without the reference's output there is no parity object; the timed numbers stand
            sys.stderr.write(f"[bench] the reference job on the e2e sample failed: {e!r}\n")
    if cj:
        p1 = cj["pass1_seconds"]
        cpu_base = {"value": n_ev / p1 if p1 else n_ev / cj["seconds"], "unit": "events/s", "cores": cj["threads"], "kind": cj["kind"],
                    "seconds_whole_job": cj["seconds"], "pass1_seconds": p1,
                    "sample": f"`yak count -k{K} -p{PRE} -b{args.bf_shift} -t{cj['threads']} -o` on the e2e sample file ({args.e2e_reads} reads, {n_ev} events): "
                              "pass-1 rate from the reference's own progress lines; whole job in seconds_whole_job (compare e2e_full_job)"}
        ga, rb = sha256_file(out_gpu), sha256_file(out_ref)
        parity = {"sha256_equal": ga == rb, "n_gpus": world, "sha256_gpu": ga, "sha256_reference": rb, "bytes": os.path.getsize(out_gpu),
                  "what": f".yak of the two-pass job on the e2e sample ({args.e2e_reads} reads): {world} GPU(s) vs oracle/_ref/yak count -b{args.bf_shift} -o"}
        if world > 1:
            # the same job behind the plain-C boundary: ONE process (the C command line over libyakb200), one host thread per GPU,
            # peer copies instead of NCCL (csrc/capi.cu multi_batch).  The other ranks have left; their GPUs are free.
            cli = os.path.join(ROOT, "yak_b200", "bin", "yak-b200")
            out_cli = sample + ".cli.yak"
            t0 = time.time()
            try:
                r = subprocess.run([cli, "count", f"-k{K}", f"-p{PRE}", f"-b{args.bf_shift}", f"-H{NH}", "-g", str(world), "-o", out_cli, sample],
                                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, timeout=900)
                rc, err_tail = r.returncode, r.stderr[-600:]
            except Exception as e:  # noqa: BLE001 - a secondary leg must not take the bench line with it
                rc, err_tail = -1, repr(e)[-600:]
            dtc = time.time() - t0
            ok = rc == 0 and os.path.exists(out_cli)
            gc = sha256_file(out_cli) if ok else None
            parity["c_api_multi_gpu"] = {"sha256_equal": bool(ok and gc == rb), "n_gpus": world, "seconds": dtc, "rc": rc,
                                         "input_events_per_s": n_ev / dtc if dtc > 0 else None,
                                         "what": f"`yak-b200 count -g {world} -b{args.bf_shift} -o` on the same file: process start, both passes, shrink, dump"}
            if not ok:
                parity["c_api_multi_gpu"]["stderr_tail"] = err_tail
            # reported, not folded into sha256_equal / the exit code: that flag is about the arm that was timed above
            try:
                os.unlink(out_cli)
            except OSError:
                pass
    for p in (sample, out_gpu, out_ref):
        try:
            os.unlink(p)
        except OSError:
            pass

    # ---- roofline (live CUDA-event times and unit counts from inside the library, rank 0's)
    peak, peak_kind = measured_peaks()
    roof, per_kernel, part_insert = roofline_objects(prof, ms, peak, peak_kind)
    # the whole step against the same peak, by SURVEY 8(d)'s fixed accounting for pass 1 with a filter
    # (extract 8.31 + insert 24 + bloom block 128 = 160.3 B per k-mer event, whatever the implementation skips)
    alg_step = ALG_STEP["pass1_bloom"] if args.bf_shift > PRE else ALG_STEP["plain"]
    roof_step = {"bound": "hbm", "achieved": alg_step * ev_all / (ms / 1000.0) / 1e9 / world, "peak": peak, "unit": "GB/s",
                 "algorithmic_bytes_per_event": alg_step, "per": "GPU"}
    roof_step["frac"] = roof_step["achieved"] / peak
    cfg = base_config(args, world)
    line = {"metric": f"k-mer events/s (k={K}, pass 1 of `yak count -b{args.bf_shift}`, chunk steps)", "value": value, "unit": "events/s",
            "n_gpus": world, "steps": KS, "warmup": W, "ms_per_step": ms / KS, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": cfg,
            "parallelism": (f"{world} GPUs: every step's {R} reads split by rank, 2^{PRE}/{world} sub-tables per rank, one NCCL all-to-all per round of {per_round} reads per rank"
                            if world > 1 else f"1 GPU, {rounds} chunk(s) of {per_round} reads per step"),
            "device_bytes": int(dev_bytes), "events": int(ev_all), "n_pending": int(pend_all),
            "input_gbp_per_s": (R * L * KS) / (ms / 1000.0) / 1e9,
            "gpu_launches": int(launches), "kernels_ms": {k: round(v[0], 3) for k, v in prof.items()},
            "clocks": clk, "roofline": roof, "roofline_partition_insert": part_insert, "roofline_step": roof_step,
            "roofline_kernels": {k: {"achieved": round(v["achieved"], 1), "frac": round(v["frac"], 4), "ms": round(v["kernel_ms_total"], 2), "units": v["units"]}
                                 for k, v in per_kernel.items()}}
    if exchange:
        line["exchange"] = exchange
    if e2e:
        line["e2e"] = e2e
        line["e2e_full_job"] = e2e_full
    if cpu_base:
        line["cpu_baseline"] = cpu_base
    if parity:
        line["parity"] = parity
    if args.verbose:
        sys.stderr.write(json.dumps(per_step) + "\n")
    emit_result(line)
    if parity and not parity["sha256_equal"]:
        sys.stderr.write("[bench] PARITY FAILURE: the .yak bytes differ from the reference's\n")
        sys.exit(1)


if __name__ == "__main__":
    main()
