#!/bin/bash
# round 2, call 19 (1 GPU): final records on the final tree - GPU suite, bench as the driver runs it + reference arm, launch list of the same command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c19_pytest.log 2>&1; echo "pytest gpu rc=$?"; tail -n 3 gpurun_out/r2c19_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c19_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r2c19_smoke.log
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c19_bench.json 2> gpurun_out/r2c19_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c19_ref.json 2> gpurun_out/r2c19_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c19_launches.csv python bench.py --gpus 1 --steps 20 --warmup 5 --no-e2e > gpurun_out/r2c19_launches_bench.json 2> /dev/null; echo "launch list rc=$?"
for f in gpurun_out/r2c19_bench.json gpurun_out/r2c19_ref.json gpurun_out/r2c19_launches_bench.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "%.3f G events/s" % (d["value"] / 1e9), {k: round(v) for k, v in d.get("kernels_ms", {}).items()})
    for k in ("e2e", "e2e_full_job", "cpu_baseline", "parity", "roofline_step", "roofline_partition_insert"):
        if k in d: print("   ", k, json.dumps(d[k])[:500])
    if d.get("roofline"): print("    roofline", json.dumps(d["roofline"])[:900])
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
