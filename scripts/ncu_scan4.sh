#!/bin/bash
# ncu --set full of the four first-superstep scans and the four renaming scans of ONE timed bench step (after the warm-up step)
cd $GRAFT_REPO_ROOT
CMD="python bench.py --scale 26 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0"
$CMD > gpurun_out/r02_ncu_plain_scan4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_lcc_first_packed' -s 4 -c 4 -o gpurun_out/r02_prof_first4 -f $CMD > gpurun_out/r02_ncu_first4.log 2>&1
ncu -i gpurun_out/r02_prof_first4.ncu-rep --page raw --csv > gpurun_out/r02_prof_first4_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:'k_lcc_scan<0, 0, 1' -s 4 -c 4 -o gpurun_out/r02_prof_xlate4 -f $CMD > gpurun_out/r02_ncu_xlate4.log 2>&1
ncu -i gpurun_out/r02_prof_xlate4.ncu-rep --page raw --csv > gpurun_out/r02_prof_xlate4_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_prof_first4.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r02_prof_first4_source.csv 2>/dev/null
ls -la gpurun_out/ | grep "first4\|xlate4"
