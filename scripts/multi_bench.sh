cd $GRAFT_REPO_ROOT
N=${1:-2}
export PM_BENCH_SCALE=${2:-26}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps ${3:-3} --warmup 3 > gpurun_out/bench_n$N.log 2>&1
echo "rc=$?"
tail -5 gpurun_out/bench_n$N.log
