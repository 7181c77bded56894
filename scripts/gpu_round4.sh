cd $GRAFT_REPO_ROOT
make -C oracle -s
nvidia-smi -L | wc -l
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 17 4 > gpurun_out/multi_check4.log 2>&1
echo "rc=$?"
grep -v "^\*\|OMP_NUM" gpurun_out/multi_check4.log | tail -12
