cd $GRAFT_REPO_ROOT
make -C oracle -s
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/tests_multi.log
cat gpurun_out/tests_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/bench_n2.log 2>&1
echo "rc=$?"
tail -1 gpurun_out/bench_n2.log | cut -c1-300
