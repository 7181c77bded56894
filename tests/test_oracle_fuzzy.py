"""run_fuzzy path: the C++ oracle against the literal transliteration with randomised
message delivery, plus hand-derived known answers."""
import numpy as np
import pytest

from tests import cases
from fuzzypatternmatching_b200 import patterns as PT
from oracle import ref_literal_fuzzy as LF


def _literal(edges, n, labels, spec, seed):
    adj = [[] for _ in range(n)]
    for a, b in edges:
        adj[a].append(b)
        adj[b].append(a)
    k = len(spec["labels"])
    pat_adj = [set() for _ in range(k)]
    for a, b in spec["edges"]:
        pat_adj[a].add(b)
        pat_adj[b].add(a)
    cons = [([spec["labels"][w] for w in c["walk"]], list(c["walk"]), len(c["walk"]) - 2, bool(c.get("cycle")))
            for c in spec["constraints"]]
    return LF.run(adj, [int(x) for x in labels], spec["labels"], pat_adj, spec["diameter"], cons, seed=seed)


@pytest.mark.parametrize("name,spec,labelset", [("triangle", PT.triangle(1, 2, 3), [1, 2, 3]),
                                                ("cycle4", PT.cycle4(1, 2, 3, 4), [1, 2, 3, 4])])
def test_fuzzy_oracle_matches_literal_under_random_delivery(oracle, name, spec, labelset):
    nontrivial = 0
    for seed in range(10):
        n, m = 50 + 7 * (seed % 3), 200 + 50 * (seed % 4)
        edges = cases.random_multigraph(seed, n, m)
        labels = cases.random_labels(seed, n, labelset)
        g = oracle.Graph.from_undirected(n, edges)
        ref = oracle.Run(g, labels, oracle.Pattern(cases.pattern_dir(spec)), fuzzy=True, max_iterations=50)
        for order in (1, 2):
            rows, vpi, itr = _literal(edges, n, labels, spec, seed * 10 + order)
            assert rows == ref.rows and itr == ref.iterations
            v, t = ref.active_vertices()
            assert {int(a): int(b).bit_length() - 1 for a, b in zip(v, t)} == vpi
        nontrivial += ref.rows[-1][3] > 0
    assert nontrivial >= 3


def test_fuzzy_known_answers(oracle):
    spec = PT.triangle(1, 2, 3)
    d = cases.pattern_dir(spec)
    # a triangle 0-1-2 with labels 1,2,3 plus a pendant path 2-3-4 (labels 1, 2): LCC keeps 3 and 4 alive
    # (each hears every template neighbour's label) ... except that 4 (label 2) needs a label-3 neighbour: dies,
    # then 3 (label 1) loses its label-2 neighbour: dies.  The triangle survives token passing.
    edges = [(0, 1), (1, 2), (2, 0), (2, 3), (3, 4)]
    labels = np.array([1, 2, 3, 1, 2], dtype=np.uint64)
    g = oracle.Graph.from_undirected(5, edges)
    ref = oracle.Run(g, labels, oracle.Pattern(d), fuzzy=True)
    v, t = ref.active_vertices()
    assert v.tolist() == [0, 1, 2] and t.tolist() == [1, 2, 4]
    # an open path 1-2-3 (no closing edge): every vertex misses one template neighbour in the first superstep
    g2 = oracle.Graph.from_undirected(3, [(0, 1), (1, 2)])
    ref2 = oracle.Run(g2, np.array([1, 2, 3], dtype=np.uint64), oracle.Pattern(d), fuzzy=True)
    assert ref2.active_vertices()[0].size == 0
    # a 6-ring labelled 1,2,3,1,2,3 passes LCC everywhere and holds no triangle — but token passing only runs
    # after an LCC call that removed a vertex (run_pattern_matching.cpp:511), so the ring alone survives ...
    ring = [(i, (i + 1) % 6) for i in range(6)]
    g3 = oracle.Graph.from_undirected(6, ring)
    ref3 = oracle.Run(g3, np.array([1, 2, 3, 1, 2, 3], dtype=np.uint64), oracle.Pattern(d), fuzzy=True)
    assert ref3.active_vertices()[0].size == 6 and ref3.iterations == 1
    assert ref3.rows[0][3] == 6  # all six vertices are in the map after the first superstep
    # ... while a pendant vertex that LCC removes (label 1, no label-3 neighbour) triggers token passing: the two
    # label-1 sources on the ring find no closed walk and are erased, and the next LCC call removes the rest
    g4 = oracle.Graph.from_undirected(7, ring + [(1, 6)])
    ref4 = oracle.Run(g4, np.array([1, 2, 3, 1, 2, 3, 1], dtype=np.uint64), oracle.Pattern(d), fuzzy=True)
    assert ref4.active_vertices()[0].size == 0 and ref4.iterations >= 2
    assert ("TP" in [r[1] for r in ref4.rows])


def test_prototype_set_generation(tmp_path):
    """BASELINE configs[3]: connected edge-deleted variants within edit distance k, as pattern directories."""
    import os
    protos = PT.edit_distance_prototypes(PT.cycle4(5, 6, 7, 8), 2)
    assert [len(g) for g, _ in protos] == [0, 1, 1, 1, 1]  # two deletions disconnect a 4-cycle
    assert len(protos[0][1]["constraints"]) == 2 and all(not p["constraints"] for _, p in protos[1:])
    assert [p["diameter"] for _, p in protos] == [3, 3, 3, 3, 3]
    six = PT.edit_distance_prototypes(PT.cycle6_chords([4, 5, 6, 7, 8, 9]), 2)
    assert len(six) == 1 + 8 + 26 and all(len(p["edges"]) == 8 - len(g) for g, p in six)
    dirs = PT.write_prototype_set(str(tmp_path), PT.cycle4(5, 6, 7, 8), 1)
    assert [os.path.basename(d) for _, _, d in dirs] == ["0", "1", "2", "3", "4"]
    assert open(os.path.join(dirs[1][2], "pattern_stat")).read().strip() == "diameter : 3"
    assert open(os.path.join(dirs[1][2], "pattern_nlc")).read().strip() == ""
