# every launch of one bench step with its device time (cold-cache, serialised: compare SHARES)
cd $GRAFT_REPO_ROOT
CMD="python bench.py --scale 26 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$1.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/plain.log
