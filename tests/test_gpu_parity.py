"""GPU parity: libpmgpu.so (through the C ABI) against the CPU oracle on the same
seeded inputs.  Bit-exact: per-superstep active vertex / edge counts, final
vertex -> template bitset map, final edge set, iteration count and the enumerated
subgraph rows."""
import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from fuzzypatternmatching_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _compare(oracle, eng, n, edges, labels, spec, tds_from):
    d = cases.pattern_dir(spec)
    g = oracle.Graph.from_undirected(n, edges)
    pat = oracle.Pattern(d)
    ref = oracle.Run(g, labels, pat, tds_from_pl=tds_from, max_iterations=50)
    assert not ref.hazards[:4].any(), "input makes the reference order dependent"
    src, dst = cases.slots_of(edges)
    eng.graph_from_slots(n, src, dst)
    eng.labels_set(labels)
    eng.pattern_load_dir(d)
    eng.run(tds_from_pl=tds_from, max_iterations=50)
    got, want = cases.engine_summary(eng, pat.n_constraints), cases.run_summary(ref)
    for k in ("rows", "iterations", "vertices", "edges", "subgraphs"):
        assert got[k] == want[k], k
    return ref


def test_graph_store_matches_oracle(oracle, eng):
    edges = cases.random_multigraph(3, 500, 4000, dup=0.2, loops=0.1)
    g = oracle.Graph.from_undirected(500, edges)
    src, dst = cases.slots_of(edges)
    eng.graph_from_slots(500, src, dst)
    assert np.array_equal(eng.graph_degree(), g.degree)
    rowptr, col = eng.graph_csr()
    assert np.array_equal(rowptr, g.rowptr)
    assert np.array_equal(col, g.col)
    eng.labels_degree_log2()
    assert np.array_equal(eng.labels_get(), g.labels_degree_log2())


@pytest.mark.parametrize("name,spec,labelset,tds_from", cases.SPECS, ids=[s[0] for s in cases.SPECS])
def test_small_random_graphs(oracle, eng, name, spec, labelset, tds_from):
    nontrivial = 0
    for seed in range(12):
        n, m = 60 + 10 * (seed % 4), 220 + 60 * (seed % 5)
        edges = cases.random_multigraph(seed, n, m)
        labels = cases.random_labels(seed, n, labelset)
        ref = _compare(oracle, eng, n, edges, labels, spec, tds_from)
        nontrivial += ref.rows[-1][3] > 0
    assert nontrivial >= 3


@pytest.mark.parametrize("name,spec,labelset,tds_from", cases.SPECS, ids=[s[0] for s in cases.SPECS])
def test_medium_graphs_with_hubs(oracle, eng, name, spec, labelset, tds_from):
    # a few very high degree vertices exercise the warp- and CTA-per-vertex kernels
    n = 20000
    edges = cases.random_multigraph(101, n, 60000, dup=0.05, loops=0.01)
    import random
    rng = random.Random(5)
    for hub, deg in ((7, 9000), (11, 5000), (13, 300), (17, 100)):
        edges += [(hub, rng.randrange(n)) for _ in range(deg)]
    labels = cases.random_labels(9, n, labelset)
    _compare(oracle, eng, n, edges, labels, spec, tds_from)


def test_rmat_graph_is_bit_exact(oracle, eng):
    g = oracle.Graph.rmat(17, 4)
    eng.graph_rmat(17, 4)
    gi = eng.graph_info()
    assert gi["n_slots_multi"] == g.n_slots_multi and gi["n_slots"] == g.n_slots
    assert np.array_equal(eng.graph_degree(), g.degree)
    rowptr, col = eng.graph_csr()
    assert np.array_equal(rowptr, g.rowptr) and np.array_equal(col, g.col)


@pytest.mark.parametrize("scale,gen_ranks", [(17, 4), (18, 8)])
def test_rmat_tree_search(oracle, eng, scale, gen_ranks):
    from fuzzypatternmatching_b200 import patterns as PT
    d = cases.pattern_dir(PT.RMAT_LOG2_TREE)
    g = oracle.Graph.rmat(scale, gen_ranks)
    labels = g.labels_degree_log2()
    pat = oracle.Pattern(d)
    ref = oracle.Run(g, labels, pat, tds_from_pl=4)
    eng.graph_rmat(scale, gen_ranks)
    eng.labels_degree_log2()
    assert np.array_equal(eng.labels_get(), labels)
    eng.pattern_load_dir(d)
    eng.run(tds_from_pl=4)
    got, want = cases.engine_summary(eng, pat.n_constraints), cases.run_summary(ref)
    for k in ("rows", "iterations", "vertices", "edges", "subgraphs"):
        assert got[k] == want[k], k
