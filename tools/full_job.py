"""The complete config-2 job on one GPU with device-generated reads: pass 1, destroy_bf, clear, pass 2,
shrink (layout of every sub-table), hist.  Prints stage times; optional partial dump."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from yak_b200 import capi
lib = capi.lib()
G = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_000_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 75
bf = int(sys.argv[3]) if len(sys.argv) > 3 else 37
nr, L = 8_000_000, 150   # chunks of 8 M reads: 75 of them are config 2's 600 M reads
cur = torch.cuda.current_stream().cuda_stream
g2 = torch.empty((G + 31) // 32 + 1, dtype=torch.int64, device="cuda")
lib.yakb_synth_genome_dev(bench.SEED_G, G, g2.data_ptr(), cur)
buf = torch.empty(nr * (L + 1), dtype=torch.uint8, device="cuda")
h = lib.yak_ch_init(31, 12, 4, bf)
stats = (C.c_uint64 * 4)()
out = {}
def one_pass(create_new):
    t_dev, ev = 0.0, 0
    for i in range(steps):
        lib.yakb_synth_reads_dev(g2.data_ptr(), G, bench.SEED_R, i * nr, nr, L, bench.ERR, bench.NPCT, 0, buf.data_ptr(), cur)
        torch.cuda.synchronize()
        t0 = time.time()
        assert lib.yakb_count_ascii_dev(h, buf.data_ptr(), nr * (L + 1), create_new, stats) == 0
        t_dev += time.time() - t0
        ev += stats[0]
    return t_dev, ev
t, ev = one_pass(1); out["pass1_s"] = t; out["events"] = ev; out["distinct_pass1"] = int(h.contents.tot)
out["dev_GB_after_pass1"] = lib.yakb_ch_device_bytes(h) / 1e9
print(json.dumps(out), flush=True)
t0 = time.time(); lib.yak_ch_destroy_bf(h); lib.yak_ch_clear(h, 1); out["destroy_bf_clear_s"] = time.time() - t0
t, ev = one_pass(0); out["pass2_s"] = t
print(json.dumps(out), flush=True)
lib.yakb_prof_enable(1)
t0 = time.time(); lib.yak_ch_shrink(h, 2, 1023, 1); out["shrink_s"] = time.time() - t0; out["distinct_final"] = int(h.contents.tot)
pj = C.create_string_buffer(1 << 16); lib.yakb_prof_json(pj, 1 << 16); lib.yakb_prof_enable(0)
out["shrink_kernels_ms"] = {k: round(v[0], 1) for k, v in json.loads(pj.value.decode()).items()}
print(json.dumps(out), flush=True)
hist = (C.c_int64 * 1024)()
t0 = time.time(); lib.yak_ch_hist(h, hist, 1); out["hist_s"] = time.time() - t0
out["hist_peak"] = max(range(2, 1024), key=lambda i: hist[i]); out["hist_sum"] = sum(hist)
out["total_s"] = out["pass1_s"] + out["destroy_bf_clear_s"] + out["pass2_s"] + out["shrink_s"]
out["input_events_per_s"] = out["events"] / out["total_s"]
print(json.dumps(out), flush=True)
if "qv" in sys.argv[4:]:
    # BASELINE configs[2]: yak qv of 50 x 1 Mbp contigs (1e-4 substitutions) against the resident table, through
    # yak_qv(opt, file, ch, cnt): parse + H2D + batched device lookups + histogram
    nctg, lctg = 50, 1_000_000
    cbuf = torch.empty(nctg * (lctg + 4), dtype=torch.uint8, device="cuda")
    lib.yakb_synth_reads_dev(g2.data_ptr(), G, 99, 0, nctg, lctg, 1e-4, 0, 1, cbuf.data_ptr(), cur)
    torch.cuda.synchronize()
    fn = "/dev/shm/yakb_ctg.fa"
    cbuf.cpu().numpy().tofile(fn)
    qo = capi.YakQopt()
    lib.yak_qopt_init(C.byref(qo))
    cnt = (C.c_int64 * 1024)()
    for rep in range(2):
        t0 = time.time(); lib.yak_qv(C.byref(qo), fn.encode(), h, cnt); out["qv_s"] = time.time() - t0
    out["qv_lookups"] = sum(cnt); out["qv_lookups_per_s"] = sum(cnt) / out["qv_s"]; out["qv_absent_or_zero"] = cnt[0]
    os.unlink(fn)
    print(json.dumps(out), flush=True)
if "dump" in sys.argv[4:]:
    t0 = time.time(); rc = lib.yak_ch_dump(h, b"/dev/shm/yakb_full.yak"); out["dump_s"] = time.time() - t0
    out["dump_bytes"] = os.path.getsize("/dev/shm/yakb_full.yak") if rc == 0 else -1
    os.unlink("/dev/shm/yakb_full.yak")
    print(json.dumps(out), flush=True)
lib.yak_ch_destroy(h)
