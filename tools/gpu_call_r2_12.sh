#!/bin/bash
# round 2, call 12 (8 GPUs): the strong-scaling bench at N=8 with parity on a reduced sample (8 M reads: the reference's job is charged 8x)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l; nproc; free -g | sed -n 2p; df -h /dev/shm | tail -1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 5 --e2e-reads 8000000 > gpurun_out/r2c12_bench8.json 2> gpurun_out/r2c12_bench8.err; echo "bench8 rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2c12_bench8.json"))
    print("N=8 %.2f G events/s, %.1f ms/step" % (d["value"] / 1e9, d["ms_per_step"]), {k: round(v) for k, v in d.get("kernels_ms", {}).items()})
    for k in ("e2e", "e2e_full_job", "cpu_baseline", "parity", "exchange"):
        if k in d: print("   ", k, json.dumps(d[k])[:900])
except Exception as e:
    print("unreadable:", e)
    import subprocess
    print(subprocess.run("grep -v '^\\[M::' gpurun_out/r2c12_bench8.err | tail -n 30", shell=True, capture_output=True, text=True).stdout)
PY
