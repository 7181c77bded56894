cd $GRAFT_REPO_ROOT
PM_WORKLOAD=hubs PM_ROWS=1 timeout 600 python scripts/explore.py 20,22,24 4 > gpurun_out/r02_explore_hubs.log 2>&1
grep -v "^       (1\|kstat 3\|kstat 4" gpurun_out/r02_explore_hubs.log | tail -40
