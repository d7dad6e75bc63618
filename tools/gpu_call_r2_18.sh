#!/bin/bash
# round 2, call 18 (1 GPU): 32-bit positions - chunks of 16 M reads (2.4 G positions): self-consistency with the 8 M-read path, then speed
mkdir -p gpurun_out
timeout 900 python tools/chunk_consistency.py > /dev/null 2> gpurun_out/r2c18_chunks.log; echo "chunk consistency rc=$?"; grep -v "^\[M::" gpurun_out/r2c18_chunks.log | tail -n 4
for cr in 16000000 8000000; do
	timeout 600 python bench.py --no-e2e --steps 20 --warmup 5 --chunk-reads $cr > gpurun_out/r2c18_cr$cr.json 2> gpurun_out/r2c18_cr$cr.err
done
for f in gpurun_out/r2c18_cr*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "%.3f G events/s" % (d["value"] / 1e9), {k: round(v) for k, v in d.get("kernels_ms", {}).items()}, d["device_bytes"] / 1e9)
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
    import subprocess
    print(subprocess.run("grep -v '^\\[M::' " + sys.argv[1].replace(".json", ".err") + " | tail -n 8", shell=True, capture_output=True, text=True).stdout)
PY
done
