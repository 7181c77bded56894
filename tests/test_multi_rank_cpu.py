"""Host-side logic of the N > 1 path on CPU (gloo, world_size 2): the per-rank result layout the
multi-GPU engine is checked against, and bench.py's rank handling for the reference arm."""
import json
import os
import subprocess
import sys

import numpy as np

from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from fuzzypatternmatching_b200 import patterns as PT
    # rank 0 fixes the inputs, everybody gets the same ones (what bench.py / tests/multi_gpu_check.py do with the NCCL id)
    box = [None]
    if rank == 0:
        edges = cases.random_multigraph(5, 300, 1800)
        labels = cases.random_labels(5, 300, [1, 2, 3, 4]).tolist()
        box = [(edges, labels, PT.write_pattern_dir(os.path.join(tmp, "pat"), PT.cycle4(1, 2, 3, 4)))]
    dist.broadcast_object_list(box, src=0)
    edges, labels, d = box[0]
    g = O.Graph.from_undirected(300, edges)
    labels = np.asarray(labels, dtype=np.uint64)
    one = O.Run(g, labels, O.Pattern(d), n_ranks=1, tds_from_pl=1, max_iterations=50)
    two = O.Run(g, labels, O.Pattern(d), n_ranks=world, tds_from_pl=1, max_iterations=50)
    # results do not depend on the number of ranks ...
    assert one.rows == two.rows and np.array_equal(one.active_edges, two.active_edges)
    # ... and this rank's share is the vertices / sources it owns: owner(v) = v mod world
    v, t = two.active_vertices()
    mine = [(int(a), int(b)) for a, b in zip(v, t) if a % world == rank]
    mine_e = [tuple(map(int, e)) for e in two.active_edges.tolist() if e[0] % world == rank]
    got = [None] * world
    dist.all_gather_object(got, (mine, mine_e))
    all_v = sorted(x for part in got for x in part[0])
    all_e = sorted(x for part in got for x in part[1])
    assert all_v == [(int(a), int(b)) for a, b in zip(v, t)]
    assert all_e == sorted(tuple(map(int, e)) for e in two.active_edges.tolist())
    # the written per-rank files carry exactly that share
    out = os.path.join(tmp, "res")
    if rank == 0:
        os.makedirs(out, exist_ok=True)
        O.make_result_tree(out)
        two.write_results(out)
    dist.barrier()
    rows = open(os.path.join(out, "0", "all_ranks_active_vertices", "active_vertices_%d" % rank)).read().splitlines()
    assert sorted(int(r.split(",")[1]) for r in rows) == [a for a, _ in mine]
    dist.destroy_process_group()


def test_two_rank_result_layout(tmp_path, oracle):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, 29541, str(tmp_path)), nprocs=2, join=True)


def test_reference_arm_runs_on_rank0_only(oracle):
    """bench.py --impl reference under a 2-rank launch: rank 0 prints the JSON line, rank 1 exits 0 silently."""
    env = dict(os.environ, PM_BENCH_SCALE="17", PM_BENCH_CPU_SCALE="17", PM_BENCH_REF_SCALE="13", MASTER_ADDR="127.0.0.1",
               MASTER_PORT="29542", WORLD_SIZE="2")
    outs = []
    for rank in (0, 1):
        e = dict(env, RANK=str(rank), LOCAL_RANK=str(rank))
        p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                            "--steps", "1", "--warmup", "1", "--workload", "tree"], cwd=ROOT, env=e,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append(p.stdout.strip())
    assert outs[1] == ""
    line = json.loads(outs[0].splitlines()[-1])
    # the reference's own binary where oracle/_ref was built (the container with the reference tree), else the oracle port
    from oracle import reference_run as R
    assert line["cpu_baseline"]["kind"] == ("reference" if R.available() else "port")
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["n_gpus"] == 2
