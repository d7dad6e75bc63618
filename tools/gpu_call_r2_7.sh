#!/bin/bash
# round 2, call 7 (1 GPU): batched multi-bucket probe (probe_inc4): parity + speed; ncu --set full of the three top kernels in the steady state
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q > gpurun_out/r2c7_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/r2c7_pytest.log
timeout 600 python bench.py --no-e2e --steps 20 --warmup 5 > gpurun_out/r2c7_bench.json 2> gpurun_out/r2c7_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c7_bench.json"))
print("%.2f G events/s" % (d["value"] / 1e9), {k: round(v) for k, v in d.get("kernels_ms", {}).items()})
PY
# steady state: 10 warm-up steps (160 M reads = 8x), then capture the first launch of each kernel in the timed steps
for kern in zone_probe part_scatter group_insert; do
	timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kern -s 22 -c 1 -f -o gpurun_out/r2c7_$kern \
		python bench.py --no-e2e --steps 2 --warmup 10 > gpurun_out/r2c7_ncu_$kern.log 2>&1; echo "ncu $kern rc=$?"
done
ls -la gpurun_out/*.ncu-rep
