set -x
cd $GRAFT_REPO_ROOT
make -C oracle -s
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/tests.log
cat gpurun_out/tests.log
timeout 600 python scripts/explore.py 24,26 1024 > gpurun_out/explore5.log 2>&1
cat gpurun_out/explore5.log
