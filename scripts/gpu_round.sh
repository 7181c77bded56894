cd $GRAFT_REPO_ROOT
make -C oracle -s
timeout 900 python -m pytest tests -x -q -m gpu --durations=6 2>&1 | tail -25 > gpurun_out/tests.log
cat gpurun_out/tests.log
