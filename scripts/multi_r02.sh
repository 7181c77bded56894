#!/bin/bash
# multi-GPU round: parity check (tests/multi_gpu_check.py), per-row timings, bench line.  Usage: multi_r02.sh <N> <tag> [explore scale] [bench steps]
N=${1:-2}; tag=${2:-x}
cd $GRAFT_REPO_ROOT
make -C oracle -s
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 17 4 > gpurun_out/r02_multi_check_n${N}_$tag.log 2>&1
echo "check rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r02_multi_check_n${N}_$tag.log | tail -12
if [ -n "$3" ]; then
PM_DEBUG_HOPS=${PM_DEBUG_HOPS:-} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/multi_explore.py $3 1024 > gpurun_out/r02_multi_explore_n${N}_$tag.log 2>&1
echo "explore rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r02_multi_explore_n${N}_$tag.log | tail -70
fi
if [ -n "$4" ]; then
PM_DEBUG_BUILD=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps $4 --warmup 3 > gpurun_out/r02_bench_n${N}_$tag.log 2> gpurun_out/r02_bench_n${N}_$tag.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/r02_bench_n${N}_$tag.log; tail -5 gpurun_out/r02_bench_n${N}_$tag.err
fi
