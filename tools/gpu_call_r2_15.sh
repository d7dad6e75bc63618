#!/bin/bash
# round 2, call 15 (1 GPU): the final records - GPU suite, bench as the driver runs it (+ reference arm), ncu launch list and --set full captures
# of the three top kernels, k sweep (config 5), the complete config-2 job with qv (configs 2 and 3)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c15_pytest.log 2>&1; echo "pytest gpu rc=$?"; tail -n 3 gpurun_out/r2c15_pytest.log
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c15_bench.json 2> gpurun_out/r2c15_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c15_ref.json 2> gpurun_out/r2c15_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c15_launches.csv python bench.py --no-e2e --steps 2 --warmup 1 > gpurun_out/r2c15_launches_bench.json 2> /dev/null; echo "launch list rc=$?"
for kern in zone_probe part_scatter group_insert; do
	timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kern -s 12 -c 1 -f -o gpurun_out/r2c15_$kern \
		python bench.py --no-e2e --steps 2 --warmup 6 --verbose > gpurun_out/r2c15_ncu_$kern.json 2> gpurun_out/r2c15_ncu_$kern.err; echo "ncu $kern rc=$?"
done
for k in 21 47 63; do
	timeout 600 python bench.py --no-e2e --steps 20 --warmup 5 --k $k > gpurun_out/r2c15_k$k.json 2> gpurun_out/r2c15_k$k.err; echo "k=$k rc=$?"
done
timeout 900 python tools/full_job.py 3000000000 75 37 qv > gpurun_out/r2c15_full_job.json 2> gpurun_out/r2c15_full_job.err; echo "full job rc=$?"; tail -n 1 gpurun_out/r2c15_full_job.json
for f in gpurun_out/r2c15_bench.json gpurun_out/r2c15_ref.json gpurun_out/r2c15_k21.json gpurun_out/r2c15_k47.json gpurun_out/r2c15_k63.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "%.3f G events/s" % (d["value"] / 1e9), {k: round(v) for k, v in d.get("kernels_ms", {}).items()})
    for k in ("e2e", "e2e_full_job", "cpu_baseline", "parity", "roofline_step", "roofline_partition_insert"):
        if k in d: print("   ", k, json.dumps(d[k])[:500])
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
