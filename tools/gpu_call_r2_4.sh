#!/bin/bash
# round 2, call 4 (2 GPUs): multi-GPU parity tests (new route kernels, torchrun entry point) and the strong-scaling bench at N=2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -4; nproc; free -g | sed -n 2p
timeout 1200 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r2c4_pytest.log 2>&1; echo "pytest dist rc=$?"
tail -n 12 gpurun_out/r2c4_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2c4_bench2.json 2> gpurun_out/r2c4_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2c4_bench2.json"))
    print("N=2 %.2f G events/s" % (d["value"] / 1e9), {k: round(v) for k, v in d.get("kernels_ms", {}).items()})
    for k in ("e2e", "e2e_full_job", "cpu_baseline", "parity", "roofline_step", "exchange"):
        if k in d: print("   ", k, json.dumps(d[k])[:500])
except Exception as e:
    print("unreadable:", e)
    import subprocess
    print(subprocess.run("grep -v '^\\[M::' gpurun_out/r2c4_bench2.err | tail -n 25", shell=True, capture_output=True, text=True).stdout)
PY
