cd $GRAFT_REPO_ROOT
make -C oracle -s
nvidia-smi -L
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py ${2:-17} ${3:-4} > gpurun_out/multi_check.log 2>&1
echo "rc=$?"
tail -60 gpurun_out/multi_check.log
