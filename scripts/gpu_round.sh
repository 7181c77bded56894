cd $GRAFT_REPO_ROOT
make -C oracle -s
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/tests.log
cat gpurun_out/tests.log
PM_ROWS=1 timeout 600 python scripts/explore.py 26 1024 > gpurun_out/explore23.log 2>&1
grep -E "^  (tree|tri|cyc|kstat)" gpurun_out/explore23.log | cut -c1-120
