cd $GRAFT_REPO_ROOT
timeout 40 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "host_csr or edge_cases or graph_store or golden" 2>&1 | tail -4
