#!/bin/bash
# round 2, call 22 (1 GPU): the whole first pass of config 2 with the final defaults (37 steps of 16 M reads), smoke
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c22_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r2c22_smoke.log
timeout 900 python bench.py --no-e2e --verbose > gpurun_out/r2c22_wholepass.json 2> gpurun_out/r2c22_wholepass.err; echo "whole pass rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c22_wholepass.json"))
print("whole pass: %.3f G events/s, %d steps + %d warm-up, %.1f ms/step" % (d["value"] / 1e9, d["steps"], d["warmup"], d["ms_per_step"]), {k: round(v) for k, v in d.get("kernels_ms", {}).items()})
print("roofline_step", d["roofline_step"]["frac"], "pending", d["n_pending"], "events", d["events"])
PY
grep -a "^\[\[" gpurun_out/r2c22_wholepass.err | tail -n 1 | cut -c1-3000
