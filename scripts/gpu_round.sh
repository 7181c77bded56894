cd $GRAFT_REPO_ROOT
PM_ROWS=1 timeout 600 python scripts/explore.py 26 1024 > gpurun_out/explore16.log 2>&1
CMD="python bench.py --scale 26 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_lcc_scan$' -s 9 -c 1 -o gpurun_out/prof_xlate_r01f -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
timeout 600 ncu --set full --clock-control none -k regex:'^k_lcc_scan$' -s 8 -c 10 -o /tmp/prof_scans -f $CMD > gpurun_out/ncu_full2.log 2>&1
ncu -i /tmp/prof_scans.ncu-rep --page raw --csv > gpurun_out/prof_scans_r01f_raw.csv
ls -la gpurun_out /tmp/prof_scans.ncu-rep
