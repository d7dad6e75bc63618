"""time the pieces of `yak count -b37 -o` through the C API on a device-generated FASTQ sample (YAKB_TIMING=1)"""
import ctypes as C, os, sys, time
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
for rep in range(2):
    t0 = time.time()
    h = capi.count_file(fn, k=31, pre=12, bf_shift=37)
    t1 = time.time()
    lib.yak_ch_dump(h, b"/dev/shm/yakb_probe.yak")
    t2 = time.time()
    lib.yak_ch_destroy(h)
    print(f"rep {rep}: count+shrink {t1-t0:.3f} s, dump {t2-t1:.3f} s, total {t2-t0:.3f} s -> {n_ev/(t2-t0)/1e6:.1f} M input events/s", flush=True)
os.unlink(fn); os.unlink("/dev/shm/yakb_probe.yak")
