#!/bin/bash
# round 2, call 16 (1 GPU): the key that looks like an empty slot (parity), smoke(), zone_probe without / with the L2 prefetch at 2 and 3 CTAs per SM
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_setops.py tests/test_gpu_scan.py -x -q -m gpu > gpurun_out/r2c16_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/r2c16_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c16_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r2c16_smoke.log
for v in "0 2" "0 3" "1 2"; do set -- $v
	YAKB_ZPROBE_PREFETCH=$1 YAKB_ZPROBE_OCC=$2 timeout 600 python bench.py --no-e2e --steps 20 --warmup 5 > gpurun_out/r2c16_pf$1_occ$2.json 2> gpurun_out/r2c16_pf$1_occ$2.err
done
for f in gpurun_out/r2c16_pf*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "%.3f G events/s" % (d["value"] / 1e9), {k: round(v) for k, v in d.get("kernels_ms", {}).items()})
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
