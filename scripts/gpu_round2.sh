cd $GRAFT_REPO_ROOT
timeout 62 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 17 4 > gpurun_out/multi_check2.log 2>&1
echo "rc=$?"
grep -v "^\*\|OMP_NUM" gpurun_out/multi_check2.log | tail -5
