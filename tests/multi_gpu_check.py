"""Multi-GPU parity check: run under torchrun with one rank per GPU.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tests/multi_gpu_check.py [scale] [gen_ranks]

Every rank searches its partition (owner(v) = v mod G); rank 0 gathers the per-rank rows, vertex /
edge lists and enumerated subgraphs and compares them with the CPU oracle run with n_ranks = G.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from fuzzypatternmatching_b200.engine import Engine  # noqa: E402
from fuzzypatternmatching_b200 import patterns as PT  # noqa: E402
from tests import cases  # noqa: E402


def _oracle_rank_rows(ref, world):
    """per-rank (vertices, edges) of every row from the count files the oracle writes (beta.cpp:504-535 layout)"""
    import tempfile
    from oracle import oracle as O
    d = tempfile.mkdtemp(prefix="pm_orc_")
    O.make_result_tree(d)
    ref.write_results(d)
    out = []
    for r in range(world):
        nv = [int(l.split(",")[-1]) for l in open(os.path.join(d, "0", "all_ranks_active_vertices_count", "active_vertices_%d" % r))]
        ne = [int(l.split(",")[-1]) for l in open(os.path.join(d, "0", "all_ranks_active_edges_count", "active_edges_%d" % r))]
        out.append(list(zip(nv, ne)))
    return out


def partition_parity(eng, dist, rank, world, name, build_graph, oracle_graph, labels, spec, tds_from, log=print,
                     delegate_threshold=0):
    """Collective.  Every rank searches its partition with `eng`; rank 0 compares the per-rank rows, vertex / edge
    lists and enumerated subgraphs with the CPU oracle run with n_ranks = world.  Returns True / False on rank 0
    (None elsewhere).  `dist` is torch.distributed with an initialised group (any backend)."""
    d = cases.pattern_dir(spec) if rank == 0 else None
    box = [d]
    dist.broadcast_object_list(box, src=0)
    d = box[0]
    build_graph()
    n_hubs = eng.graph_set_delegate_threshold(delegate_threshold) if delegate_threshold else 0
    if labels is None:
        eng.labels_degree_log2()
    else:
        eng.labels_set(labels)
    eng.pattern_load_dir(d)
    eng.run(tds_from_pl=tds_from, max_iterations=50)
    ncons = len(spec["constraints"])
    mine = dict(rows=eng.rows(), iterations=int(eng.summary["iterations"]),
                vertices=[tuple(map(int, x)) for x in zip(*eng.active_vertices())],
                edges=[tuple(map(int, x)) for x in eng.active_edges().tolist()],
                subgraphs=[sorted(map(tuple, eng.subgraphs(pl).tolist())) for pl in range(ncons)],
                labels=eng.labels_get().tolist() if labels is None else None)
    got = [None] * world
    dist.gather_object(mine, got if rank == 0 else None, dst=0)
    if rank != 0:
        return None
    from oracle import oracle as O
    g = oracle_graph()
    lab = g.labels_degree_log2() if labels is None else labels
    ref = O.Run(g, lab, O.Pattern(d), n_ranks=world, tds_from_pl=tds_from, max_iterations=50,
                delegate_threshold=delegate_threshold)
    want = cases.run_summary(ref)
    ok = True
    # who writes a vertex's rows: v mod ranks, or — for a hub — its controller, delegate id mod ranks
    hubs = np.nonzero(g.degree >= delegate_threshold)[0].tolist() if delegate_threshold else []
    ctl = {v: i % world for i, v in enumerate(hubs)}
    owner = lambda v: ctl.get(v, v % world)  # noqa: E731
    if delegate_threshold:
        ok &= n_hubs == len(hubs) and len(hubs) > 0
        want_rows = _oracle_rank_rows(ref, world)
        for r, gr in enumerate(got):  # every rank's count file, row by row
            ok &= [(x[3], x[4]) for x in gr["rows"]] == want_rows[r]
    if labels is None:
        ok &= all(gr["labels"] == lab.tolist() for gr in got)
    # rows: per-rank counts sum to the oracle's totals
    rows = [(r[0], r[1], r[2], sum(gr["rows"][i][3] for gr in got), sum(gr["rows"][i][4] for gr in got))
            for i, r in enumerate(got[0]["rows"])]
    ok &= rows == want["rows"]
    ok &= all(gr["iterations"] == want["iterations"] for gr in got)
    for r, gr in enumerate(got):
        ok &= gr["vertices"] == [x for x in want["vertices"] if owner(x[0]) == r]
        ok &= gr["edges"] == [x for x in want["edges"] if owner(x[0]) == r]
        for pl in range(ncons):
            ok &= gr["subgraphs"][pl] == [w for w in want["subgraphs"][pl] if owner(w[-1]) == r]
    log("%-28s %s  rows %d final (%d, %d) subgraphs %s" % (
        name, "ok" if ok else "MISMATCH", len(rows), rows[-1][3] if rows else -1, rows[-1][4] if rows else -1,
        [len(x) for x in want["subgraphs"]]))
    if not ok and rows != want["rows"]:
        for a, b in zip(rows, want["rows"]):
            if a != b:
                log("   first differing row: got %s want %s" % (a, b))
                break
    return bool(ok)


def partition_parity_fuzzy(eng, dist, rank, world, name, build_graph, oracle_graph, labels, spec, log=print):
    """The run_fuzzy_pattern_matching path (pm_run_fuzzy) over `world` ranks against the oracle: per-superstep map sizes
    summed over the ranks, iteration count, and every rank's final (vertex, 1 << vertex_pattern_index) list."""
    d = cases.pattern_dir(spec) if rank == 0 else None
    box = [d]
    dist.broadcast_object_list(box, src=0)
    d = box[0]
    build_graph()
    if labels is None:
        eng.labels_degree_log2()
    else:
        eng.labels_set(labels)
    eng.pattern_load_dir(d)
    eng.run_fuzzy(max_iterations=50)
    mine = dict(rows=eng.rows(), iterations=int(eng.summary["iterations"]),
                vertices=[tuple(map(int, x)) for x in zip(*eng.active_vertices())])
    got = [None] * world
    dist.gather_object(mine, got if rank == 0 else None, dst=0)
    if rank != 0:
        return None
    from oracle import oracle as O
    g = oracle_graph()
    lab = g.labels_degree_log2() if labels is None else labels
    ref = O.Run(g, lab, O.Pattern(d), fuzzy=True, max_iterations=50)
    rv, rt = ref.active_vertices()
    want_v = list(zip(rv.tolist(), rt.tolist()))
    rows = [(r[0], r[1], r[2], sum(gr["rows"][i][3] for gr in got), 0) for i, r in enumerate(got[0]["rows"])]
    ok = rows == ref.rows and all(gr["iterations"] == ref.iterations for gr in got)
    for r, gr in enumerate(got):
        ok &= gr["vertices"] == [x for x in want_v if x[0] % world == r]
    log("%-28s %s  rows %d final %d vertices" % (name, "ok" if ok else "MISMATCH", len(rows), len(want_v)))
    if not ok and rows != ref.rows:
        for a, b in zip(rows, ref.rows):
            if a != b:
                log("   first differing row: got %s want %s" % (a, b))
                break
    return bool(ok)


def main():
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 17
    gen_ranks = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    eng = Engine(local)
    ids = [Engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    eng.comm_init(rank, world, ids[0])
    failures = []

    def check(name, build_graph, oracle_graph, labels, spec, tds_from, delegate_threshold=0):
        ok = partition_parity(eng, dist, rank, world, name, build_graph, oracle_graph, labels, spec, tds_from,
                              log=lambda m: print(m, flush=True), delegate_threshold=delegate_threshold)
        if rank == 0 and not ok:
            failures.append(name)

    from oracle import oracle as O
    # small random graphs, every template family
    for name, spec, labelset, tds_from in cases.SPECS:
        for seed in range(4):
            n, m = 60 + 10 * (seed % 4), 220 + 60 * (seed % 5)
            if name == "cycle6":  # random graphs of this size never hold the template: plant it
                edges, labels = cases.planted(seed, n, m, spec, labelset)
            else:
                edges = cases.random_multigraph(seed, n, m)
                labels = cases.random_labels(seed, n, labelset)
            src, dst = cases.slots_of(edges)
            check("%s/seed%d" % (name, seed), lambda: eng.graph_from_slots(n, src, dst),
                  lambda: O.Graph.from_undirected(n, edges), labels, spec, tds_from)
    # R-MAT, generated across the GPUs and shuffled to the owners
    check("rmat%d/tree" % scale, lambda: eng.graph_rmat(scale, gen_ranks), lambda: O.Graph.rmat(scale, gen_ranks),
          None, PT.RMAT_LOG2_TREE, 4)
    for nm, spec, tds in (("triangle", PT.triangle(6, 7, 8), 1), ("cycle4", PT.cycle4(5, 6, 7, 8), 1)):
        check("rmat%d/%s" % (scale, nm), lambda: eng.graph_rmat(scale, gen_ranks), lambda: O.Graph.rmat(scale, gen_ranks),
              None, spec, tds)
    # host CSR round trip: every rank reads its rows back and re-opens the graph from them (pm_graph_from_csr)
    def reopen_from_host_csr():
        eng.graph_rmat(scale, gen_ranks)
        rowptr, col = eng.graph_csr()
        degm = eng.graph_degree()
        nv = eng.graph_info()["n_vertices"]
        eng.graph_from_csr(rowptr, col, degm, n_vertices=nv)
    check("rmat%d/host_csr/triangle" % scale, reopen_from_host_csr, lambda: O.Graph.rmat(scale, gen_ranks),
          None, PT.triangle(6, 7, 8), 1)
    # delegates: hubs (multigraph degree >= threshold) are attributed to their controller ranks in every per-rank output
    for nm, spec, tds, thr in (("tree", PT.RMAT_LOG2_TREE, 4, 64), ("triangle", PT.triangle(6, 7, 8), 1, 96),
                               ("cycle4", PT.cycle4(5, 6, 7, 8), 1, 40)):
        check("delegates%d/rmat%d/%s" % (thr, scale, nm), lambda: eng.graph_rmat(scale, gen_ranks),
              lambda: O.Graph.rmat(scale, gen_ranks), None, spec, tds, delegate_threshold=thr)
    # the run_fuzzy path over the same partition (unique-label LCC + cycle token passing over the unpruned adjacency)
    def check_fuzzy(name, build_graph, oracle_graph, labels, spec):
        ok = partition_parity_fuzzy(eng, dist, rank, world, name, build_graph, oracle_graph, labels, spec,
                                    log=lambda m: print(m, flush=True))
        if rank == 0 and not ok:
            failures.append(name)

    for nm, spec, labelset in (("triangle", PT.triangle(1, 2, 3), [1, 2, 3]), ("cycle4", PT.cycle4(1, 2, 3, 4), [1, 2, 3, 4])):
        for seed in range(6):
            n, m = 60 + 10 * (seed % 4), 220 + 60 * (seed % 5)
            edges = cases.random_multigraph(seed, n, m)
            labels = cases.random_labels(seed, n, labelset)
            src, dst = cases.slots_of(edges)
            check_fuzzy("fuzzy/%s/seed%d" % (nm, seed), lambda: eng.graph_from_slots(n, src, dst),
                        lambda: O.Graph.from_undirected(n, edges), labels, spec)
    check_fuzzy("fuzzy/rmat%d/cycle4" % scale, lambda: eng.graph_rmat(scale, gen_ranks), lambda: O.Graph.rmat(scale, gen_ranks),
                None, PT.cycle4(5, 6, 7, 8))
    flag = [len(failures)]
    dist.broadcast_object_list(flag, src=0)
    eng.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI-GPU PARITY", "FAILED: %s" % failures if failures else "OK (%d ranks)" % world, flush=True)
    sys.exit(1 if flag[0] else 0)


if __name__ == "__main__":
    main()
