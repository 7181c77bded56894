# launch list + full captures of the LCC scan kernels of ONE bench step (after a plain run of the same command)
cd $GRAFT_REPO_ROOT
TAG=$1
CMD="python bench.py --scale 26 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_lcc_scan<.bool.1' -c 4 -o gpurun_out/prof_scanfirst_$TAG -f $CMD > gpurun_out/ncu_full.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_lcc_scan<.bool.0' -c 8 -o gpurun_out/prof_scanlater_$TAG -f $CMD > gpurun_out/ncu_full2.log 2>&1
tail -2 gpurun_out/ncu_full2.log
ls -la gpurun_out/
