#!/bin/bash
# round 2, call 1: the GPU suite as the driver runs it, the never-run staged zone scatter (parity, then speed over zone and chunk sizes)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | sed -n 2p; df -h /dev/shm | tail -1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest gpu rc=$?"
YAKB_TEST_UNVERIFIED=1 timeout 600 python -m pytest tests/test_gpu_parity.py -k "zone_staged" -q > gpurun_out/r2c1_pytest_unverified.log 2>&1; echo "unverified rc=$?"
tail -n 3 gpurun_out/r2c1_pytest.log gpurun_out/r2c1_pytest_unverified.log
timeout 600 python bench.py --no-e2e --no-cpu > gpurun_out/r2c1_default_2m.json 2> gpurun_out/r2c1_default_2m.err
timeout 600 python bench.py --no-e2e --no-cpu --chunk-reads 5000000 --steps 117 > gpurun_out/r2c1_default_5m.json 2> gpurun_out/r2c1_default_5m.err
for mb in 32 64 128; do
	YAKB_ZONE=2 YAKB_ZONE_MB=$mb YAKB_ZONE_STAGED=1 timeout 600 python bench.py --no-e2e --no-cpu > gpurun_out/r2c1_zone${mb}_2m.json 2> gpurun_out/r2c1_zone${mb}_2m.err
	YAKB_ZONE=2 YAKB_ZONE_MB=$mb YAKB_ZONE_STAGED=1 timeout 600 python bench.py --no-e2e --no-cpu --chunk-reads 5000000 --steps 117 > gpurun_out/r2c1_zone${mb}_5m.json 2> gpurun_out/r2c1_zone${mb}_5m.err
done
for f in gpurun_out/r2c1_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "%.2f G events/s" % (d["value"] / 1e9), {k: round(v) for k, v in d.get("kernels_ms", {}).items()})
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
