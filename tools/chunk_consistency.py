"""chunk_consistency.py - results must not depend on how the read stream is cut into chunks, also beyond 2^31 positions
per chunk: the same 32 M reads of the cfg2 stream counted as 4 chunks of 8 M reads and as 2 chunks of 16 M (2.4 G positions each), both
passes + shrink; the two .yak images must have the same sha256 (the 8 M-read path is the one checked against the reference)."""
import ctypes as C, hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from yak_b200 import capi
lib = capi.lib()
G, L, total = 3_000_000_000, 150, 32_000_000
cur = torch.cuda.current_stream().cuda_stream
g2 = torch.empty((G + 31) // 32 + 1, dtype=torch.int64, device="cuda")
lib.yakb_synth_genome_dev(bench.SEED_G, G, g2.data_ptr(), cur)
stats = (C.c_uint64 * 4)()
digests = []
for nr in (8_000_000, 16_000_000):
    buf = torch.empty(nr * (L + 1), dtype=torch.uint8, device="cuda")
    h = lib.yak_ch_init(31, 12, 4, 37)
    t0 = time.time()
    for create_new in (1, 0):
        if create_new == 0:
            lib.yak_ch_destroy_bf(h); lib.yak_ch_clear(h, 1)
        for i in range(total // nr):
            lib.yakb_synth_reads_dev(g2.data_ptr(), G, bench.SEED_R, i * nr, nr, L, bench.ERR, bench.NPCT, 0, buf.data_ptr(), cur)
            torch.cuda.synchronize()
            assert lib.yakb_count_ascii_dev(h, buf.data_ptr(), nr * (L + 1), create_new, stats) == 0
    lib.yak_ch_shrink(h, 2, 1023, 1)
    out = f"/dev/shm/yakb_chunk_{nr}.yak"
    assert lib.yak_ch_dump(h, out.encode()) == 0
    sha = hashlib.sha256(open(out, "rb").read()).hexdigest()
    print(f"chunks of {nr} reads ({nr * (L + 1)} positions): {time.time() - t0:.1f} s, {os.path.getsize(out)} bytes, sha256 {sha}", file=sys.stderr, flush=True)
    digests.append(sha)
    os.unlink(out)
    lib.yak_ch_destroy(h)
    del buf
    lib.yakb_device_cache_trim(); torch.cuda.empty_cache()
print("identical:", digests[0] == digests[1], file=sys.stderr)
sys.exit(0 if digests[0] == digests[1] else 1)
