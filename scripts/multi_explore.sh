cd $GRAFT_REPO_ROOT
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/multi_explore.py ${2:-27} 1024 > gpurun_out/multi_explore.log 2>&1
echo "rc=$?"
grep -v "^\*\|OMP_NUM" gpurun_out/multi_explore.log | tail -80
