"""The oracle's multi-rank attribution with delegates (what tests/multi_gpu_check.py holds the engine to at 2/4/8 GPUs), against
the reference's rules restated independently here, straight from the edge list:
  * hubs = vertices whose out-degree — directed slots, parallel edges and self loops counted — reaches the threshold
    (impl/delegate_partitioned_graph.ipp:501-512); delegate id = position in the ascending list of hubs (:681-683);
    controller = delegate id mod ranks (delegate_partitioned_graph.hpp:231-233);
  * a vertex of the final vertex_state_map is written, with its edges, by its controller if it is a hub and by rank
    `v mod ranks` otherwise (beta.cpp:1380-1410), and counted in that rank's count files (ee.hpp:1112-1125);
  * nothing else changes: the union over the ranks is the single-rank result and the per-rank rows add up to its rows."""
import collections
import os

import numpy as np
import pytest

from fuzzypatternmatching_b200 import patterns as PT
from tests import cases


def _rows(path):
    return [[t.strip() for t in l.split(",")] for l in open(path).read().splitlines() if l.strip()]


@pytest.mark.parametrize("ranks,threshold", [(2, 14), (3, 15), (4, 16), (8, 17)])
def test_hub_rows_belong_to_their_controller(oracle, tmp_path, ranks, threshold):
    spec = PT.triangle(1, 2, 3)
    d = cases.pattern_dir(spec)
    n = 160
    edges = cases.random_multigraph(ranks * 31 + threshold, n, 900, dup=0.2, loops=0.1)
    labels = cases.random_labels(ranks * 31 + threshold, n, [1, 2, 3])
    out_degree = collections.Counter()
    for a, b in edges:  # an undirected input edge is two directed slots; a self loop is two slots of the same vertex
        out_degree[a] += 1
        out_degree[b] += 1
    hubs = sorted(v for v in range(n) if out_degree[v] >= threshold)
    assert 5 <= len(hubs) < (2 * n) // 3
    controller = {v: i % ranks for i, v in enumerate(hubs)}
    owner = lambda v: controller[v] if v in controller else v % ranks  # noqa: E731

    g = oracle.Graph.from_undirected(n, edges)
    assert np.array_equal(g.degree, np.array([out_degree[v] for v in range(n)], dtype=np.uint64))
    one = oracle.Run(g, labels, oracle.Pattern(d), n_ranks=1, tds_from_pl=1, max_iterations=50)
    many = oracle.Run(g, labels, oracle.Pattern(d), n_ranks=ranks, tds_from_pl=1, max_iterations=50, delegate_threshold=threshold)
    assert many.rows == one.rows  # the aggregate rows of a multi-rank run are the single-rank rows
    out = str(tmp_path / "tree")
    oracle.make_result_tree(out)
    many.write_results(out)

    v1, t1 = one.active_vertices()
    final_vertices = dict(zip(v1.tolist(), t1.tolist()))
    final_edges = set(map(tuple, one.active_edges.tolist()))
    assert len(final_vertices) > 10 and any(v in controller and owner(v) != v % ranks for v in final_vertices), \
        "the input must end with hubs whose controller is not their modulo owner"
    seen_v, seen_e = {}, set()
    vertex_rows = [None] * ranks
    for r in range(ranks):
        rows = _rows(os.path.join(out, "0", "all_ranks_active_vertices", "active_vertices_%d" % r))
        for t in rows:  # "rank, vertex, pattern index, label, bitset"
            v = int(t[1])
            assert int(t[0]) == r == owner(v), (r, v)
            assert v not in seen_v
            seen_v[v] = int(t[4], 2)
            assert int(t[3]) == int(labels[v])
        vertex_rows[r] = len(rows)
        erows = _rows(os.path.join(out, "0", "all_ranks_active_edges", "active_edges_%d" % r))
        for t in erows:  # "rank, vertex, neighbour"
            assert int(t[0]) == r == owner(int(t[1]))
            seen_e.add((int(t[1]), int(t[2])))
        # the count files: one row per superstep / constraint; the last row counts what this rank wrote
        vc = _rows(os.path.join(out, "0", "all_ranks_active_vertices_count", "active_vertices_%d" % r))
        ec = _rows(os.path.join(out, "0", "all_ranks_active_edges_count", "active_edges_%d" % r))
        assert int(vc[-1][3]) == len(rows) and int(ec[-1][3]) == len(erows)
    assert seen_v == final_vertices and seen_e == final_edges
    # row by row, the ranks' counts add up to the single-rank rows
    per_rank = [_rows(os.path.join(out, "0", "all_ranks_active_vertices_count", "active_vertices_%d" % r)) for r in range(ranks)]
    per_rank_e = [_rows(os.path.join(out, "0", "all_ranks_active_edges_count", "active_edges_%d" % r)) for r in range(ranks)]
    assert all(len(p) == len(one.rows) for p in per_rank)
    for i, row in enumerate(one.rows):
        assert sum(int(p[i][3]) for p in per_rank) == row[3] and sum(int(p[i][3]) for p in per_rank_e) == row[4], i
        assert all((int(p[i][0]), p[i][1], int(p[i][2])) == row[:3] for p in per_rank)
