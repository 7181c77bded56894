cd $GRAFT_REPO_ROOT
make -C oracle -s
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/tests.log
cat gpurun_out/tests.log
PM_ROWS=1 timeout 600 python scripts/explore.py 26 1024 > gpurun_out/explore17.log 2>&1
grep -E "^  (tree|tri|cyc|kstat)" gpurun_out/explore17.log
PM_KEEP_LABW=1 PM_ROWS=1 timeout 600 python scripts/explore.py 26 1024 > gpurun_out/explore17_labw.log 2>&1
grep -E "^  (tree|tri|cyc|kstat)" gpurun_out/explore17_labw.log
PM_DEBUG_BUILD=1 timeout 600 python scripts/e2e_breakdown.py 26 > gpurun_out/e2e_breakdown2.log 2>&1
tail -12 gpurun_out/e2e_breakdown2.log
CMD="python bench.py --scale 26 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_lcc_scan$' -s 8 -c 1 -o gpurun_out/prof_first_r01g -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
