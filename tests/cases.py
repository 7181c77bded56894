"""Seeded inputs shared by the CPU (oracle) and GPU (parity) tests."""
import random
import tempfile

import numpy as np

from fuzzypatternmatching_b200 import patterns as PT


def random_multigraph(seed, n, m, dup=0.1, loops=0.05):
    """generated undirected edges with duplicates and a few self loops"""
    rng = random.Random(seed)
    e = []
    for _ in range(m):
        a, b = rng.randrange(n), rng.randrange(n)
        if a == b and rng.random() > loops:
            b = (a + 1) % n
        e.append((a, b))
        if rng.random() < dup:
            e.append((a, b))
    return e


def random_labels(seed, n, labelset):
    rng = random.Random(seed * 7919 + 13)
    return np.array([rng.choice(labelset) for _ in range(n)], dtype=np.uint64)


def slots_of(edges):
    e = np.asarray(edges, dtype=np.uint64).reshape(-1, 2)
    src = np.empty(2 * len(e), dtype=np.uint64)
    dst = np.empty(2 * len(e), dtype=np.uint64)
    src[0::2], dst[0::2] = e[:, 0], e[:, 1]
    src[1::2], dst[1::2] = e[:, 1], e[:, 0]
    return src, dst


# (name, spec, labelset, tds_from_pl)
SPECS = [
    ("tree", PT.RMAT_LOG2_TREE, [2, 3, 4, 5, 7], 4),
    ("triangle", PT.triangle(1, 2, 3), [1, 2, 3], 1),
    ("cycle4", PT.cycle4(1, 2, 3, 4), [1, 2, 3, 4], 1),
    ("cycle6", PT.cycle6_chords([1, 2, 3, 4, 5, 6]), [1, 2, 3, 4, 5, 6], 3),
    # a single LCC superstep per call (pattern_stat "diameter : 1"): edge maps keep neighbours that left the
    # vertex map until the next call
    ("triangle_d1", dict(PT.triangle(1, 2, 3), diameter=1), [1, 2, 3], 1),
]


def pattern_dir(spec):
    return PT.write_pattern_dir(tempfile.mkdtemp(prefix="pmpat_"), spec)


def run_summary(run):
    """comparable view of an oracle Run"""
    v, t = run.active_vertices()
    return dict(rows=run.rows, iterations=run.iterations,
                vertices=list(zip(v.tolist(), t.tolist())),
                edges=[tuple(x) for x in run.active_edges.tolist()],
                subgraphs=[sorted(map(tuple, s.tolist())) for s in run.subgraphs])


def engine_summary(eng, n_constraints):
    """the same view of a GPU Engine after run()"""
    v, t = eng.active_vertices()
    return dict(rows=eng.rows(), iterations=int(eng.summary["iterations"]),
                vertices=list(zip(v.tolist(), t.tolist())),
                edges=[tuple(x) for x in eng.active_edges().tolist()],
                subgraphs=[sorted(map(tuple, eng.subgraphs(pl).tolist())) for pl in range(n_constraints)])


def edge_cases():
    """(name, n, undirected edges, labels, spec, tds_from_pl): degenerate and ragged inputs — nothing matches, a single
    edge, only self loops, isolated vertices, vertex counts that straddle the 16-slot words and 4096-slot tiles of the
    compact id tables, a template without non-local constraints."""
    tri = PT.triangle(1, 2, 3)
    lcc_only = {"labels": [1, 2, 1], "edges": [(0, 1), (1, 2)], "diameter": 2, "constraints": []}
    out = [
        ("no_template_label", 50, random_multigraph(1, 50, 200), np.full(50, 9, dtype=np.uint64), tri, 1),
        ("labels_match_structure_never", 30, [(i, i + 1) for i in range(29)],
         np.array([1 + i % 3 for i in range(30)], dtype=np.uint64), tri, 1),
        ("single_edge", 2, [(0, 1)], np.array([1, 2], dtype=np.uint64), lcc_only, 0),
        ("self_loops_only", 5, [(i, i) for i in range(5)], np.array([1, 2, 1, 2, 1], dtype=np.uint64), lcc_only, 0),
    ]
    for n in (17, 4097, 8193):
        labels = np.full(n, 7, dtype=np.uint64)
        a, b, c = n - 1, n // 2, 0
        labels[[a, b, c]] = [1, 2, 3]
        out.append(("one_triangle_among_%d_isolated" % n, n, [(a, b), (b, c), (a, c)], labels, tri, 1))
    for seed in range(4):
        out.append(("lcc_only_seed%d" % seed, 200, random_multigraph(seed + 70, 200, 900),
                    random_labels(seed + 70, 200, [1, 2]), lcc_only, 0))
    return out
