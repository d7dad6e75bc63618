"""per-kernel device time of pass 1 through yak_count(file) at several host batch sizes (YAKB_BATCH)"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from yak_b200 import capi
lib = capi.lib()
G = 3_000_000_000
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
g2 = torch.empty((G + 31) // 32 + 1, dtype=torch.int64, device="cuda")
lib.yakb_synth_genome_dev(bench.SEED_G, G, g2.data_ptr(), torch.cuda.current_stream().cuda_stream)
fn = "/dev/shm/yakb_probe.fq"
n_ev = bench.make_sample_file(torch, lib, g2, G, n_reads, 0, fn)
del g2; torch.cuda.empty_cache()
o = capi.copt(31, 12, 37, 4)
for batch in sys.argv[2:] or ["0"]:
    if batch != "0":
        os.environ["YAKB_BATCH"] = batch
    for rep in range(2):
        lib.yakb_prof_enable(1)
        t0 = time.time()
        h = lib.yak_count(fn.encode(), C.byref(o), None)
        t1 = time.time()
        pj = C.create_string_buffer(1 << 16)
        lib.yakb_prof_json(pj, 1 << 16)
        lib.yakb_prof_enable(0)
        prof = json.loads(pj.value.decode())
        lib.yak_ch_destroy(h)
        tot = sum(v[0] for v in prof.values())
        print(f"batch {batch} rep {rep}: yak_count {t1-t0:.3f} s -> {n_ev/(t1-t0)/1e6:.0f} M events/s; kernels {tot:.1f} ms: "
              + ", ".join(f"{k} {v[0]:.1f}/{v[1]}" for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])), flush=True)
os.unlink(fn)
