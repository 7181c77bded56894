#!/bin/bash
# One GPU-box round of this repo's checks: parity tests, the bench line, per-row timings.  Usage: gpu_r02.sh <tag> [tests]
tag=${1:-x}
cd $GRAFT_REPO_ROOT
if [ "$2" != "notests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_$tag.log
  tail -8 gpurun_out/r02_pytest_gpu_$tag.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1_$tag.log 2> gpurun_out/r02_bench_n1_$tag.err; echo "bench rc=$?"
tail -c 3500 gpurun_out/r02_bench_n1_$tag.log
PM_ROWS=1 timeout 600 python scripts/explore.py 26 > gpurun_out/r02_explore26_$tag.log 2>&1
grep -v "^       " gpurun_out/r02_explore26_$tag.log | tail -30
