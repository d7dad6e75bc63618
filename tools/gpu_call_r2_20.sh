#!/bin/bash
# round 2, call 20 (1 GPU): journal slab sized from the filter, x4 growth of small tables, cudaMalloc accounted: parity subset, bench twice
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_ingest.py -x -q -m gpu > gpurun_out/r2c20_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2c20_pytest.log
timeout 600 python bench.py --no-e2e --steps 20 --warmup 5 > gpurun_out/r2c20_noe2e.json 2> gpurun_out/r2c20_noe2e.err
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c20_bench.json 2> gpurun_out/r2c20_bench.err; echo "bench rc=$?"
for f in gpurun_out/r2c20_noe2e.json gpurun_out/r2c20_bench.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "%.3f G events/s, %.1f ms/step, kernels %.0f ms" % (d["value"] / 1e9, d["ms_per_step"], sum(v for k, v in d["kernels_ms"].items() if not k.startswith("host"))), {k: round(v) for k, v in d.get("kernels_ms", {}).items()})
    for k in ("e2e", "e2e_full_job", "parity"):
        if k in d: print("   ", k, json.dumps(d[k])[:300])
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
