#!/bin/bash
# round 2, call 14 (2 GPUs): pipelined rounds (extraction + exchange behind the count): parity test, then bench --gpus 2 with and without
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -k "pipelined or nccl_sharded" > gpurun_out/r2c14_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2c14_pytest.log
for mr in 2 1; do
	timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2955$mr bench.py --gpus 2 --steps 10 --warmup 5 --no-e2e --min-rounds $mr > gpurun_out/r2c14_bench2_mr$mr.json 2> gpurun_out/r2c14_bench2_mr$mr.err; echo "bench2 min-rounds $mr rc=$?"
done
python - <<'PY'
import json
for mr in (2, 1):
    try:
        d = json.load(open(f"gpurun_out/r2c14_bench2_mr{mr}.json"))
        print("N=2 min-rounds %d: %.2f G events/s, %.1f ms/step" % (mr, d["value"] / 1e9, d["ms_per_step"]), {k: round(v) for k, v in d.get("kernels_ms", {}).items()}, d.get("exchange"))
    except Exception as e:
        print("unreadable:", e)
        import subprocess
        print(subprocess.run(f"grep -v '^\\[M::' gpurun_out/r2c14_bench2_mr{mr}.err | tail -n 25", shell=True, capture_output=True, text=True).stdout)
PY
