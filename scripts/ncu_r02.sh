#!/bin/bash
# ncu evidence of one bench step (1 warm-up + 1 timed step, 4 templates, R-MAT scale 26): launch list + --set full of the hot kernels
tag=${1:-final}
cd $GRAFT_REPO_ROOT
CMD="python bench.py --scale 26 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0"
$CMD > gpurun_out/r02_ncu_plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_$tag.csv $CMD > gpurun_out/r02_ncu_list_$tag.log 2>&1
tail -2 gpurun_out/r02_ncu_plain_$tag.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:'k_lcc_first_packed|k_lcc_scan|k_nem1_expand|k_nem1_close_cycle|k_init_flags|k_init_assign|k_lcc_commit' -s 160 -c 170 -o gpurun_out/r02_prof_$tag -f $CMD > gpurun_out/r02_ncu_full_$tag.log 2>&1
tail -3 gpurun_out/r02_ncu_full_$tag.log
ncu -i gpurun_out/r02_prof_$tag.ncu-rep --page raw --csv > gpurun_out/r02_prof_${tag}_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -8
timeout 600 python bench.py --scale 22 --gen-ranks 4 --workload hubs --no-cpu-baseline --e2e-steps 0 --steps 5 --warmup 3 > gpurun_out/r02_bench_s22_hubs.log 2>&1; echo "hubs rc=$?"
