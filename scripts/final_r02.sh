#!/bin/bash
# The evidence of the round with the final build, one GPU: smoke, parity tests, bench line (driver flags), per-row timings,
# reference arm, BASELINE configs[1] (scale 25 tree, LCC only is the tree workload here), hub-class workload, fuzzy prototypes
cd $GRAFT_REPO_ROOT
python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu_final.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final.log 2> gpurun_out/r02_bench_n1_final.err; echo "bench rc=$?"
PM_ROWS=1 timeout 600 python scripts/explore.py 26 > gpurun_out/r02_explore26_final.log 2>&1
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.log 2>&1; echo "reference rc=$?"
timeout 600 python bench.py --scale 25 --workload tree --no-cpu-baseline --e2e-steps 1 --steps 10 --warmup 3 > gpurun_out/r02_bench_s25_tree_config1.log 2>&1; echo "s25 rc=$?"
timeout 600 python bench.py --scale 24 --workload hubs --no-cpu-baseline --e2e-steps 0 --steps 5 --warmup 3 > gpurun_out/r02_bench_s24_hubs.log 2>&1; echo "hubs rc=$?"
timeout 900 python scripts/fuzzy_prototypes.py 24 1024 2 > gpurun_out/r02_fuzzy_prototypes_s24.log 2>&1; echo "fuzzy rc=$?"
tail -c 600 gpurun_out/r02_bench_s24_hubs.log; tail -3 gpurun_out/r02_fuzzy_prototypes_s24.log
