"""Multi-GPU parity (needs >= 2 GPUs): one process per GPU under torchrun, 1-D partition
owner(v) = v mod G, results compared with the CPU oracle run with n_ranks = G
(tests/multi_gpu_check.py does the work; this wrapper makes it part of `pytest -m gpu`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_partitioned_search_matches_oracle(world, oracle):
    if _n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29531 + world), os.path.join(ROOT, "tests", "multi_gpu_check.py"), "17", "4"]
    p = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert p.returncode == 0 and "MULTI-GPU PARITY OK" in p.stdout, p.stdout[-4000:]
