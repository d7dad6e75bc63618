#!/bin/bash
# The first GPU call of the next round, in one go (round 1 ended without GPU minutes: everything committed after
# profiles/r01_* is CPU-tested only).  Run as:  gpurun --timeout 2400 -- 'bash tools/next_gpu_call.sh'
# 1. the GPU suite as the driver runs it; 2. the parity cases of code that has never run on a GPU; 3. the default bench line;
# 4. the zone-blocked front end with the direct and the staged scatter over zone sizes and chunk sizes.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/n_pytest.log 2>&1; echo "pytest gpu rc=$?"
YAKB_TEST_UNVERIFIED=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -k "zone_staged or count_command" -q > gpurun_out/n_pytest_unverified.log 2>&1; echo "unverified rc=$?"
timeout 900 python bench.py > gpurun_out/n_bench_default.json 2> gpurun_out/n_bench_default.err; echo "bench rc=$?"
for mb in 32 64 128; do
	for st in 0 1; do
		YAKB_ZONE=2 YAKB_ZONE_MB=$mb YAKB_ZONE_STAGED=$st timeout 600 python bench.py --no-e2e --no-cpu > gpurun_out/n_zone${mb}_s${st}.json 2> gpurun_out/n_zone${mb}_s${st}.err
		YAKB_ZONE=2 YAKB_ZONE_MB=$mb YAKB_ZONE_STAGED=$st timeout 600 python bench.py --no-e2e --no-cpu --chunk-reads 5000000 --steps 117 > gpurun_out/n_zone${mb}_s${st}_5m.json 2> gpurun_out/n_zone${mb}_s${st}_5m.err
	done
done
timeout 600 python bench.py --no-e2e --no-cpu --chunk-reads 5000000 --steps 117 > gpurun_out/n_bench_5m.json 2> gpurun_out/n_bench_5m.err
# 4b. compressed input end to end: plain / gzip / BGZF, with and without the text cache for pass 2 - same .yak, wall times
python - <<'PY'
import gzip, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from yak_b200 import synth
import test_bgzf_cpu as B
t = synth.reads_file_bytes(1, 5_000_000, 2, 1_000_000, fastq=True)
open("/dev/shm/n_reads.fq", "wb").write(t)
open("/dev/shm/n_reads.fq.gz", "wb").write(gzip.compress(t, 1))
open("/dev/shm/n_reads.bgzf.gz", "wb").write(B.bgzf_bytes(t, 65280, level=1))
PY
for f in fq fq.gz bgzf.gz; do
	for c in 0 8; do
		/usr/bin/time -f "count -b30 of n_reads.$f, YAKB_TEXT_CACHE_GB=$c: %e s" env YAKB_TEXT_CACHE_GB=$c timeout 600 yak_b200/bin/yak-b200 count -k31 -p12 -b30 -o /dev/shm/n_${f}_$c.yak /dev/shm/n_reads.$f 2> gpurun_out/n_count_${f}_$c.err
		tail -n 1 gpurun_out/n_count_${f}_$c.err
	done
done
sha256sum /dev/shm/n_*.yak | tee gpurun_out/n_compressed_sha.txt
rm -f /dev/shm/n_reads.* /dev/shm/n_*.yak
# 5. the kernels under compute-sanitizer: the smoke run, memcheck then racecheck
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/n_memcheck.log 2>&1; echo "memcheck rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/n_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -n 5 gpurun_out/n_memcheck.log gpurun_out/n_racecheck.log
tail -n 3 gpurun_out/n_pytest.log gpurun_out/n_pytest_unverified.log
for f in gpurun_out/n_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "%.2f G events/s" % (d["value"] / 1e9), {k: round(v) for k, v in d.get("kernels_ms", {}).items()})
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
