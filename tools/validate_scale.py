"""Medium-scale exactness check on the GPU box: `yak-b200 count` against the UNMODIFIED reference binary
(oracle/_ref/yak) on a device-generated FASTQ, comparing .yak bytes (sha256).  Exercises many chunks, table
growth, large sub-tables (warp layout path), multi-segment journals."""
import hashlib, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from yak_b200 import capi
lib = capi.lib()
G = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
n_reads = int(float(sys.argv[2])) if len(sys.argv) > 2 else 5_000_000
g2 = torch.empty((G + 31) // 32 + 1, dtype=torch.int64, device="cuda")
lib.yakb_synth_genome_dev(bench.SEED_G, G, g2.data_ptr(), torch.cuda.current_stream().cuda_stream)
fn = "/dev/shm/yakb_val.fq"
n_ev = bench.make_sample_file(torch, lib, g2, G, n_reads, 0, fn)
del g2; torch.cuda.empty_cache()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ref, exe = os.path.join(root, "oracle", "_ref", "yak"), os.path.join(root, "yak_b200", "bin", "yak-b200")
ok = True
for k, p, b in ((31, 10, 30), (31, 12, 0), (47, 12, 31)):
    outs = []
    for tag, binary in (("ref", ref), ("b200", exe)):
        out = f"/dev/shm/yakb_val_{tag}.yak"
        cmd = [binary, "count", f"-k{k}", f"-p{p}", "-t128", "-o", out] + ([f"-b{b}"] if b else []) + [fn]
        t0 = time.time()
        subprocess.run(cmd, check=True, capture_output=True, env=dict(os.environ, YAKB_BATCH="40000000"))
        dt = time.time() - t0
        h = hashlib.sha256(open(out, "rb").read()).hexdigest()
        outs.append((tag, dt, os.path.getsize(out), h))
    same = outs[0][3] == outs[1][3]
    ok &= same
    print(f"k={k} p={p} b={b} events={n_ev}: identical={same} ref {outs[0][1]:.1f}s b200 {outs[1][1]:.1f}s bytes={outs[0][2]}", flush=True)
for f in (fn, "/dev/shm/yakb_val_ref.yak", "/dev/shm/yakb_val_b200.yak"):
    os.unlink(f)
sys.exit(0 if ok else 1)
