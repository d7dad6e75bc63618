#!/bin/bash
# round 2, call 23 (1 GPU): the new kernels under compute-sanitizer memcheck - smoke() (partitioned probe, device ingest, layout, dump) and the route extraction test
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r2c23_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"
tail -n 4 gpurun_out/r2c23_memcheck_smoke.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_dist.py -x -q -k "route_extraction and 31-12" > gpurun_out/r2c23_memcheck_route.log 2>&1; echo "memcheck route rc=$?"
tail -n 4 gpurun_out/r2c23_memcheck_route.log
