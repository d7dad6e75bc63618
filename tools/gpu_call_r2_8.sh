#!/bin/bash
# round 2, call 8 (1 GPU): pipelined zone_probe (L2 prefetch, 256-bit bucket loads), warp-cooperative group_insert: parity + speed; e2e time split
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_setops.py -x -q -m gpu > gpurun_out/r2c8_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/r2c8_pytest.log
for occ in 3 2; do
	YAKB_ZPROBE_OCC=$occ timeout 600 python bench.py --no-e2e --steps 20 --warmup 5 > gpurun_out/r2c8_occ$occ.json 2> gpurun_out/r2c8_occ$occ.err
done
YAKB_ZPROBE_OCC=2 YAKB_ZONE_MB=32 timeout 600 python bench.py --no-e2e --steps 20 --warmup 5 > gpurun_out/r2c8_occ2_z32.json 2> gpurun_out/r2c8_occ2_z32.err
YAKB_ZPROBE_OCC=2 YAKB_ZONE_MB=128 timeout 600 python bench.py --no-e2e --steps 20 --warmup 5 > gpurun_out/r2c8_occ2_z128.json 2> gpurun_out/r2c8_occ2_z128.err
for f in gpurun_out/r2c8_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "%.2f G events/s" % (d["value"] / 1e9), {k: round(v) for k, v in d.get("kernels_ms", {}).items()})
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
    import subprocess
    print(subprocess.run("grep -v '^\\[M::' " + sys.argv[1].replace(".json", ".err") + " | tail -n 6", shell=True, capture_output=True, text=True).stdout)
PY
done
# e2e time split: a 16 M-read FASTQ file through the CLI with stage timing
python - <<'PY'
import subprocess, os
gen = "oracle/_bin/synthgen"
subprocess.run([gen, "20260925", "3000000000", "7", "0", "16000000", "150", "0.005", "1", "2", "31", "/dev/shm/r2c8.fq", "16"], check=True)
PY
for i in 1 2; do YAKB_TIMING=1 yak_b200/bin/yak-b200 count -k31 -p12 -b37 -o /dev/shm/r2c8.yak /dev/shm/r2c8.fq 2>&1 | grep "T::\|Real time" | grep -v "batch [0-9]*:" | head -40; echo ==; done
rm -f /dev/shm/r2c8.*
