#!/bin/bash
# round 2, call 2: the new one-roll partition kernel (part_scatter): parity, then speed over chunk and zone sizes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "partitioned or inc or bloom_filter or count_matches" > gpurun_out/r2c2_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 15 gpurun_out/r2c2_pytest.log
run() { # name, env..., -- args
	name=$1; shift
	env "$@" > /dev/null 2>&1 || true
}
b() { name=$1; shift; timeout 600 env $ENVV python bench.py --no-e2e --no-cpu "$@" > gpurun_out/r2c2_$name.json 2> gpurun_out/r2c2_$name.err; }
ENVV="YAKB_X=1" b part_2m
ENVV="YAKB_X=1" b part_4m --chunk-reads 4000000 --steps 147
ENVV="YAKB_X=1" b part_8m --chunk-reads 8000000 --steps 72
ENVV="YAKB_ZONE_MB=32" b part32_8m --chunk-reads 8000000 --steps 72
ENVV="YAKB_ZONE_MB=128" b part128_8m --chunk-reads 8000000 --steps 72
ENVV="YAKB_ZONE=0" b nopart_8m --chunk-reads 8000000 --steps 72
ENVV="YAKB_X=1" b part_12m --chunk-reads 12000000 --steps 47
for f in gpurun_out/r2c2_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "%.2f G events/s" % (d["value"] / 1e9), {k: round(v) for k, v in d.get("kernels_ms", {}).items()})
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
    import subprocess
    print(subprocess.run("grep -v '^\\[M::' " + sys.argv[1].replace(".json", ".err") + " | tail -n 4", shell=True, capture_output=True, text=True).stdout)
PY
done
