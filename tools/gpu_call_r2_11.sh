#!/bin/bash
# round 2, call 11 (2 GPUs): bench --gpus 2 with all legs (strong scaling, per-rank ingest e2e, parity NCCL path + C path), torchrun count command test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -k "count_command" > gpurun_out/r2c11_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2c11_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 5 > gpurun_out/r2c11_bench2.json 2> gpurun_out/r2c11_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2c11_bench2.json"))
    print("N=2 %.2f G events/s" % (d["value"] / 1e9), {k: round(v) for k, v in d.get("kernels_ms", {}).items()})
    for k in ("e2e", "e2e_full_job", "cpu_baseline", "parity", "exchange"):
        if k in d: print("   ", k, json.dumps(d[k])[:900])
except Exception as e:
    print("unreadable:", e)
    import subprocess
    print(subprocess.run("grep -v '^\\[M::' gpurun_out/r2c11_bench2.err | tail -n 25", shell=True, capture_output=True, text=True).stdout)
PY
