cd $GRAFT_REPO_ROOT
make -C oracle -s
timeout 100 python bench.py --scale 25 --workload tree --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_s25_tree.log 2>&1
tail -1 gpurun_out/bench_s25_tree.log | cut -c1-400
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
