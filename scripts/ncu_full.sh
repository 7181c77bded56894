# one ncu --set full capture of the kernels matching $1 (regex), launches skipped $2, count $3
cd $GRAFT_REPO_ROOT
CMD="python bench.py --scale 26 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$1 -s ${2:-0} -c ${3:-2} -o gpurun_out/prof_$4 -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
