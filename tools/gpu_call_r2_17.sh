#!/bin/bash
# round 2, call 17 (4 GPUs): bench --gpus 4 with all legs on a 16 M-read sample (parity through NCCL and through the C path)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 10 --warmup 5 --e2e-reads 16000000 > gpurun_out/r2c17_bench4.json 2> gpurun_out/r2c17_bench4.err; echo "bench4 rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2c17_bench4.json"))
    print("N=4 %.2f G events/s, %.1f ms/step" % (d["value"] / 1e9, d["ms_per_step"]), {k: round(v) for k, v in d.get("kernels_ms", {}).items()})
    for k in ("e2e", "e2e_full_job", "cpu_baseline", "parity", "exchange"):
        if k in d: print("   ", k, json.dumps(d[k])[:900])
except Exception as e:
    print("unreadable:", e)
    import subprocess
    print(subprocess.run("grep -v '^\\[M::' gpurun_out/r2c17_bench4.err | tail -n 30", shell=True, capture_output=True, text=True).stdout)
PY
