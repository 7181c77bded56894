"""Where the end-to-end step (host CSR -> device store -> labels -> search -> results) spends its time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fuzzypatternmatching_b200.engine import Engine

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 26
eng = Engine(0)
eng.graph_rmat(scale, 1024); eng.labels_degree_log2(); gi = eng.graph_info()
pats = bench.write_patterns("cyclic")
rowptr, col = eng.graph_csr(); degm = eng.graph_degree()
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
rowptr, col, degm = pin(rowptr), pin(col), pin(degm)
# raw pinned host -> device copy of the same bytes, for scale
tcol = torch.from_numpy(col)
dcol = torch.empty_like(tcol, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); dcol.copy_(tcol, non_blocking=True); torch.cuda.synchronize()
    dt = time.perf_counter() - t
print("raw pinned H2D of col: %.1f ms, %.1f GB/s" % (dt * 1e3, col.nbytes / dt / 1e9), flush=True)
del dcol
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.graph_from_csr(rowptr, col, degm, n_vertices=gi["n_vertices"]); torch.cuda.synchronize(); t1 = time.perf_counter()
    eng.labels_degree_log2(); torch.cuda.synchronize(); t2 = time.perf_counter()
    for _, d, tds in pats:
        eng.pattern_load_dir(d); eng.run(tds_from_pl=tds, keep_subgraphs=False)
    t3 = time.perf_counter()
    for _, d, tds in pats[-1:]:
        v, b = eng.active_vertices(); e = eng.active_edges()
    t4 = time.perf_counter()
    print("rep %d: graph_from_csr %.1f ms  labels %.1f ms  search %.1f ms  fetch %.1f ms  total %.1f ms" % (
        rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t4 - t0) * 1e3), flush=True)
