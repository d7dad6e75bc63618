#!/bin/bash
# round 2, call 13 (1 GPU): the whole GPU suite on the current tree; kernel split of the all-pending first steps; config 2 at a tenth of its size with the reference beside it
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c13_pytest.log 2>&1; echo "pytest gpu rc=$?"; tail -n 4 gpurun_out/r2c13_pytest.log
timeout 600 python bench.py --no-e2e --steps 2 --warmup 0 > gpurun_out/r2c13_first2.json 2> gpurun_out/r2c13_first2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c13_first2.json"))
print("first 2 steps: %.2f G events/s, pending %d of %d" % (d["value"] / 1e9, d["n_pending"], d["events"]), {k: round(v) for k, v in d.get("kernels_ms", {}).items()})
PY
timeout 1500 python tools/validate_cfg2.py 60000000 34 1 > gpurun_out/r2c13_cfg2_9gbp.log 2>&1; echo "cfg2 rc=$?"
cat gpurun_out/r2c13_cfg2_9gbp.log
