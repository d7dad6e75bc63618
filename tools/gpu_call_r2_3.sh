#!/bin/bash
# round 2, call 3: the reworked bench.py as the driver runs it (ours + reference arm), zone_probe occupancy variants
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c3_ref.json 2> gpurun_out/r2c3_ref.err; echo "ref rc=$?"
YAKB_ZPROBE_OCC=2 timeout 600 python bench.py --no-e2e --steps 20 --warmup 5 > gpurun_out/r2c3_occ2.json 2> gpurun_out/r2c3_occ2.err
timeout 600 python bench.py --no-e2e --steps 20 --warmup 5 --chunk-reads 16000000 > /dev/null 2> gpurun_out/r2c3_16m.err; echo "16m rc=$? (expected: chunk too large)"
timeout 600 python bench.py --no-e2e --steps 20 --warmup 5 --chunk-reads 14000000 --reads-per-step 14000000 > gpurun_out/r2c3_14m.json 2> gpurun_out/r2c3_14m.err
for f in gpurun_out/r2c3_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "%.2f G events/s" % (d["value"] / 1e9), {k: round(v) for k, v in d.get("kernels_ms", {}).items()})
    for k in ("e2e", "e2e_full_job", "cpu_baseline", "parity", "roofline_step", "roofline_partition_insert"):
        if k in d: print("   ", k, json.dumps(d[k])[:400])
    if d.get("roofline"): print("    roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 4))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
    import subprocess
    print(subprocess.run("grep -v '^\\[M::' " + sys.argv[1].replace(".json", ".err") + " | tail -n 12", shell=True, capture_output=True, text=True).stdout)
PY
done
