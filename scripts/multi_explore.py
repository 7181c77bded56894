"""Per-row device times of a partitioned search (run under torchrun, one rank per GPU)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import bench  # noqa: E402
from fuzzypatternmatching_b200.engine import Engine  # noqa: E402

scale = int(sys.argv[1])
gen_ranks = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
eng = Engine(local)
ids = [Engine.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
eng.comm_init(rank, world, ids[0])
pats = bench.write_patterns("cyclic")
t = time.time()
eng.graph_rmat(scale, gen_ranks)
eng.labels_degree_log2()
if rank == 0:
    print("scale", scale, "ranks", world, "build %.2fs" % (time.time() - t), eng.graph_info(), flush=True)
for name, d, tds in pats:
    eng.pattern_load_dir(d)
    for rep in range(3):
        dist.barrier()
        t = time.time()
        sm = eng.run(tds_from_pl=tds, keep_subgraphs=False)
        dt = time.time() - t
    if rank == 0:
        print("  %-22s %.3f ms (dev %.3f ms) iters %d launches %d" % (name, dt * 1e3, sm["device_seconds"] * 1e3,
                                                                      sm["iterations"], eng.kernel_launches()), flush=True)
        for r in eng.rows_timed():
            print("       %s" % (r,), flush=True)
eng.close()
dist.destroy_process_group()
