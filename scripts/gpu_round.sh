cd $GRAFT_REPO_ROOT
PM_ROWS=1 timeout 600 python scripts/explore.py 26 1024 > gpurun_out/explore24.log 2>&1
grep -E "^  (tree|tri|cyc|kstat)" gpurun_out/explore24.log | cut -c1-120
python bench.py --no-cpu-baseline > gpurun_out/bench_n1b.log 2>&1
tail -1 gpurun_out/bench_n1b.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k in ('value','ms_per_step','e2e'): print(k, d[k])
print(d['roofline']['frac'])"
