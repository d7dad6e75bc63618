#!/bin/bash
# round 2, call 21 (1 GPU): the dump streamed from the device through page-locked pieces: the whole GPU suite (every parity test dumps), then the bench with its full-job leg
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c21_pytest.log 2>&1; echo "pytest gpu rc=$?"; tail -n 3 gpurun_out/r2c21_pytest.log
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c21_bench.json 2> gpurun_out/r2c21_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c21_bench.json"))
print("%.3f G events/s, %.1f ms/step" % (d["value"] / 1e9, d["ms_per_step"]), {k: round(v) for k, v in d.get("kernels_ms", {}).items()})
for k in ("e2e", "e2e_full_job", "parity"):
    if k in d: print("   ", k, json.dumps(d[k])[:300])
PY
python - <<'PY'
import subprocess
subprocess.run(["oracle/_bin/synthgen", "20260925", "3000000000", "7", "0", "32000000", "150", "0.005", "1", "2", "31", "/dev/shm/r2c21.fq", "16"], check=True)
PY
YAKB_TIMING=1 yak_b200/bin/yak-b200 count -k31 -p12 -b37 -o /dev/shm/r2c21.yak /dev/shm/r2c21.fq 2>&1 | grep "T::\|Real time" | grep -v "batch [0-9]*:" | head -20
sha256sum /dev/shm/r2c21.yak; rm -f /dev/shm/r2c21.*
