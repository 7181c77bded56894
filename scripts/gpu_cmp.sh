# A/B of environment switches on the bench workload: gpu_cmp.sh "<name>=<ENV=1 ...>" ...
cd $GRAFT_REPO_ROOT
for spec in "$@"; do
  name=${spec%%=*}; envs=${spec#*=}
  env $envs PM_ROWS=1 timeout 600 python scripts/explore.py 26 > gpurun_out/r02_explore26_$name.log 2>&1
  echo "== $name ($envs)"; grep "^  [a-z]" gpurun_out/r02_explore26_$name.log
done
