#!/bin/bash
# last validation of the round's final build on one GPU: parity tests, THE bench line with the driver's flags, per-row times
cd $GRAFT_REPO_ROOT
python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu_final.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final.log 2> gpurun_out/r02_bench_n1_final.err; echo "bench rc=$?"
PM_ROWS=1 timeout 600 python scripts/explore.py 26 > gpurun_out/r02_explore26_final.log 2>&1
grep "^  [a-z]" gpurun_out/r02_explore26_final.log
