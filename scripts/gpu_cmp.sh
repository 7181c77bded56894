cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_typed.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_typed.log
PM_FUSE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rmat or small_random or quirks or baseline or bench_templates or edge_cases or planted" > gpurun_out/r02_pytest_gpu_fuse.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_fuse.log
for mode in default fuse notyped; do
  if [ $mode = fuse ]; then export PM_FUSE=1; else unset PM_FUSE; fi
  if [ $mode = notyped ]; then export PM_NO_TYPED=1; else unset PM_NO_TYPED; fi
  PM_ROWS=1 timeout 600 python scripts/explore.py 26 > gpurun_out/r02_explore26_$mode.log 2>&1
  echo "== $mode"; grep "^  [a-z]" gpurun_out/r02_explore26_$mode.log
done
