#!/bin/bash
# round 2, call 9 (1 GPU): device-side text ingest: parity (strict and irregular files), then the bench with its e2e legs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ingest.py -x -q > gpurun_out/r2c9_pytest.log 2>&1; echo "pytest ingest rc=$?"
tail -n 25 gpurun_out/r2c9_pytest.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c9_bench.json 2> gpurun_out/r2c9_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2c9_bench.json"))
    print("%.2f G events/s" % (d["value"] / 1e9), {k: round(v) for k, v in d.get("kernels_ms", {}).items()})
    for k in ("e2e", "e2e_full_job", "cpu_baseline", "parity"):
        if k in d: print("   ", k, json.dumps(d[k])[:400])
except Exception as e:
    print("unreadable:", e)
    import subprocess
    print(subprocess.run("grep -v '^\\[M::' gpurun_out/r2c9_bench.err | tail -n 20", shell=True, capture_output=True, text=True).stdout)
PY
python - <<'PY'
import subprocess
subprocess.run(["oracle/_bin/synthgen", "20260925", "3000000000", "7", "0", "16000000", "150", "0.005", "1", "2", "31", "/dev/shm/r2c9.fq", "16"], check=True)
PY
for i in 1 2; do YAKB_TIMING=1 yak_b200/bin/yak-b200 count -k31 -p12 -b37 -o /dev/shm/r2c9.yak /dev/shm/r2c9.fq 2>&1 | grep "T::\|Real time" | grep -v "batch [0-9]*:" | head -40; echo ==; done
YAKB_GPU_INGEST=0 YAKB_TIMING=1 yak_b200/bin/yak-b200 count -k31 -p12 -b37 -o /dev/shm/r2c9b.yak /dev/shm/r2c9.fq 2>&1 | grep "T::\|Real time" | grep -v "batch [0-9]*:" | head -40
sha256sum /dev/shm/r2c9.yak /dev/shm/r2c9b.yak
rm -f /dev/shm/r2c9*
