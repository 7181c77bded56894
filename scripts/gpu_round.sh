cd $GRAFT_REPO_ROOT
make -C oracle -s
python bench.py > gpurun_out/bench_n1.log 2>&1
tail -1 gpurun_out/bench_n1.log | cut -c1-400
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
tail -1 gpurun_out/bench_ref.log | cut -c1-200
CMD="python bench.py --scale 26 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r01h.csv $CMD > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'^k_lcc_scan$' -s 8 -c 12 -o /tmp/prof_scans -f $CMD > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/prof_scans.ncu-rep --page raw --csv > gpurun_out/prof_scans_r01h_raw.csv
ls -la gpurun_out
