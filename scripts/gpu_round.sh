cd $GRAFT_REPO_ROOT
run() { echo "== $1 gps=$2"; PM_SCAN_GPS=$2 PMGPU_LIB=$PWD/fuzzypatternmatching_b200/$1 timeout 600 python scripts/explore.py 26 1024 > gpurun_out/explore22.log 2>&1; grep -E "^  (tree|tri|cyc|kstat [014])" gpurun_out/explore22.log | cut -c1-110; }
run libpmgpu.so 8
run libpmgpu.so 4
run libpmgpu.so 16
run libpmgpu_b5.so 5
run libpmgpu_b5.so 10
run libpmgpu_b6.so 6
run libpmgpu_b6.so 12
