cd $GRAFT_REPO_ROOT
make -C oracle -s
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 17 4 > gpurun_out/r02_multi_check_n2_fz.log 2>&1
echo "check rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r02_multi_check_n2_fz.log | tail -28
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fuzzy or label_stream" 2>&1 | tail -3
