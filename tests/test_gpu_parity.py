"""GPU parity: libpmgpu.so (through the C ABI) against the CPU oracle on the same
seeded inputs.  Bit-exact: per-superstep active vertex / edge counts, final
vertex -> template bitset map, final edge set, iteration count and the enumerated
subgraph rows."""
import os

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


def _compare(oracle, eng, n, edges, labels, spec, tds_from, quirks=False):
    d = cases.pattern_dir(spec)
    g = oracle.Graph.from_undirected(n, edges)
    pat = oracle.Pattern(d)
    ref = oracle.Run(g, labels, pat, tds_from_pl=tds_from, max_iterations=50)
    # counters 0..2: the reference itself is order dependent on this input (nothing to compare against);
    # 3 (a bit resurrected, A.6 #4) and 5 (an edge kept by a flag set outside LCC, A.6 #11) are deterministic
    # quirks the GPU must reproduce — only the quirk tests feed such inputs
    assert not ref.hazards[:3].any() and not ref.hazards[4], "input makes the reference order dependent"
    assert quirks or not ref.hazards[3], "resurrection outside the quirk tests"
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


@pytest.mark.parametrize("name,spec,labelset,tds_from,div,counter", cases.QUIRK_SPECS, ids=[q[0] for q in cases.QUIRK_SPECS])
def test_reference_quirks_are_reproduced(oracle, eng, name, spec, labelset, tds_from, div, counter):
    """SURVEY A.6 #4 (twin template vertices: a bit NLCC cleared in T_arr is resurrected from T_state by the next
    LCC post step) and A.6 #11 (an edge flagged by a successful cycle token outside LCC survives one post step
    after its neighbour was deactivated): per-superstep rows and final sets equal the oracle's, on inputs where
    the oracle's counters show the quirk occurred."""
    fired = 0
    for seed, n, m in cases.quirk_inputs(name, div):
        edges = cases.random_multigraph(seed, n, m)
        labels = cases.random_labels(seed, n, labelset)
        ref = _compare(oracle, eng, n, edges, labels, spec, tds_from, quirks=True)
        fired += int(ref.hazards[counter] > 0)
    assert fired >= 5


def test_planted_cycle6_ends_non_trivially(oracle, eng):
    name, spec, labelset, tds_from = cases.SPECS[3]
    nontrivial = 0
    for seed in range(8):
        n, m = 60 + 10 * (seed % 4), 220 + 60 * (seed % 5)
        edges, labels = cases.planted(seed, n, m, spec, labelset)
        ref = _compare(oracle, eng, n, edges, labels, spec, tds_from)
        nontrivial += ref.rows[-1][3] > 0 and len(ref.subgraphs[3]) > 0
    assert nontrivial >= 6


def test_reference_grid_graph(oracle, eng):
    """the reference's own graph fixture (test/include/input_graph.hpp:8-68) through the GPU graph store, and the
    hand-derived LCC rows of tests/test_oracle_kat.py::test_grid_graph_lcc_by_hand"""
    slots = cases.grid_graph_slots()
    src = np.array([a for a, _ in slots], dtype=np.uint64)
    dst = np.array([b for _, b in slots], dtype=np.uint64)
    eng.graph_from_slots(15, src, dst)
    assert eng.graph_degree().tolist() == cases.GRID_DEGREE
    rowptr, col = eng.graph_csr()
    assert rowptr.tolist() == cases.GRID_OFFSET and col.tolist() == [b for _, b in slots]
    eng.labels_set(np.array(cases.GRID_DEGREE, dtype=np.uint64))
    spec = {"labels": [2, 3, 4], "edges": [(0, 1), (1, 2)], "diameter": 2, "constraints": []}
    eng.pattern_load_dir(cases.pattern_dir(spec))
    eng.run(tds_from_pl=-1)
    assert eng.rows() == [(0, "LP", 0, 13, 30), (0, "LP", 1, 12, 28)]
    v, t = eng.active_vertices()
    assert v.tolist() == [0, 1, 3, 4, 5, 6, 8, 9, 10, 11, 13, 14] and t.tolist() == [1, 2, 2, 1, 2, 4, 4, 2, 1, 2, 2, 1]


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


def test_baseline_config1_scale21_reference_pattern_dir(oracle, eng):
    """BASELINE configs[0] exactly: generate_rmat -s 21 with 4 generating ranks, degree-log2 labels, the reference's
    examples/rmat_log2_tree_pattern/0 verbatim (tests/golden/rmat_log2_tree_pattern/0; checked against
    /root/reference by tests/test_oracle_kat.py), tree search + enumeration: per-superstep counts, final vertex and
    edge sets, enumerated subgraph rows and their count."""
    import os
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rmat_log2_tree_pattern", "0")
    g = oracle.Graph.rmat(21, 4)
    labels = g.labels_degree_log2()
    pat = oracle.Pattern(d)
    ref = oracle.Run(g, labels, pat, tds_from_pl=4)
    assert not ref.hazards[:5].any()
    eng.graph_rmat(21, 4)
    gi = eng.graph_info()
    assert gi["n_slots_multi"] == g.n_slots_multi == 2 ** 26 and gi["n_slots"] == g.n_slots
    assert np.array_equal(eng.graph_degree(), g.degree)
    eng.labels_degree_log2()
    assert np.array_equal(eng.labels_get(), labels)
    eng.pattern_load_dir(d)
    eng.run(tds_from_pl=4)
    got, want = cases.engine_summary(eng, pat.n_constraints), cases.run_summary(ref)
    for k in ("rows", "iterations", "vertices", "edges", "subgraphs"):
        assert got[k] == want[k], k
    assert len(want["vertices"]) > 0 and len(want["subgraphs"][4]) > 0
    assert eng.subgraph_count(4) == len(want["subgraphs"][4])
    # ... and against what THE REFERENCE ITSELF wrote for this input (its own driver + visitor headers over the single-rank
    # runtime stand-in; fixture written by oracle/make_reference_golden.py --large)
    _, golden = cases.reference_golden_load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_runs",
                                                         "rmat21_config0_reference_pattern_dir.json"))
    cases.assert_equals_reference_golden(got, golden)


def test_bench_templates_on_rmat_scale20(oracle, eng):
    """bench.py's cyclic templates (BASELINE configs[2]: triangle, 4-cycle, 6-cycle with chords) with the bench's
    label choices on R-MAT scale 20: rows, sets and enumerated walks against the oracle."""
    from fuzzypatternmatching_b200 import patterns as PT
    g = oracle.Graph.rmat(20, 4)
    labels = g.labels_degree_log2()
    eng.graph_rmat(20, 4)
    eng.labels_degree_log2()
    for nm, spec in (("triangle_678", PT.triangle(6, 7, 8)), ("cycle4_5678", PT.cycle4(5, 6, 7, 8)),
                     ("cycle6_chords_456789", PT.cycle6_chords([4, 5, 6, 7, 8, 9]))):
        d = cases.pattern_dir(spec)
        tds = PT.tds_from_pl(spec)
        pat = oracle.Pattern(d)
        ref = oracle.Run(g, labels, pat, tds_from_pl=tds)
        assert not ref.hazards[:5].any(), nm
        eng.pattern_load_dir(d)
        eng.run(tds_from_pl=tds)
        got, want = cases.engine_summary(eng, pat.n_constraints), cases.run_summary(ref)
        for k in ("rows", "iterations", "vertices", "edges", "subgraphs"):
            assert got[k] == want[k], (nm, k)
        assert len(want["rows"]) >= 8, nm
        # ... and against what THE REFERENCE ITSELF wrote for this graph and template (its driver enumerates from constraint
        # 4 on, so it ran the template with the enumeration walk moved there: same final sets and walks)
        _, golden = cases.reference_golden_load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                                             "reference_runs", "rmat20_bench_%s.json" % nm))
        cases.assert_final_sets_equal_reference_golden(got, tds, golden)


def test_label_stream_path_without_packed_labels(oracle, eng):
    """Neighbour labels normally ride in the high bits of the adjacency slots (id bits + label bits <= 32); larger
    graphs fall back to a parallel byte stream.  PM_NO_PACK forces that path here; switching back re-packs."""
    import os
    from fuzzypatternmatching_b200 import patterns as PT
    g = oracle.Graph.rmat(17, 4)
    labels = g.labels_degree_log2()
    eng.graph_rmat(17, 4)
    for mode in ("1", None, "1"):
        if mode:
            os.environ["PM_NO_PACK"] = mode
        else:
            os.environ.pop("PM_NO_PACK", None)
        try:
            eng.labels_degree_log2()
        finally:
            os.environ.pop("PM_NO_PACK", None)
        rowptr, col = eng.graph_csr()
        assert np.array_equal(col, g.col)
        for spec, tds in ((PT.RMAT_LOG2_TREE, 4), (PT.cycle4(5, 6, 7, 8), 1)):
            d = cases.pattern_dir(spec)
            pat = oracle.Pattern(d)
            ref = oracle.Run(g, labels, pat, tds_from_pl=tds)
            eng.pattern_load_dir(d)
            eng.run(tds_from_pl=tds)
            got, want = cases.engine_summary(eng, pat.n_constraints), cases.run_summary(ref)
            for k in ("rows", "iterations", "vertices", "edges", "subgraphs"):
                assert got[k] == want[k], (mode, k)
    # the run_fuzzy path builds its byte label stream on demand from the packed slots
    eng.labels_degree_log2()
    d = cases.pattern_dir(PT.cycle4(5, 6, 7, 8))
    ref = oracle.Run(g, labels, oracle.Pattern(d), fuzzy=True)
    eng.pattern_load_dir(d)
    eng.run_fuzzy()
    assert eng.rows() == ref.rows


def test_golden_fixtures_on_gpu(oracle, eng):
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "runs.json")))
    specs = {s[0]: s for s in cases.SPECS}
    for case in gold:
        _, spec, labelset, tds_from = specs[case["spec"]]
        edges = cases.random_multigraph(case["seed"], case["n"], case["m"])
        labels = cases.random_labels(case["seed"], case["n"], labelset)
        src, dst = cases.slots_of(edges)
        eng.graph_from_slots(case["n"], src, dst)
        eng.labels_set(labels)
        eng.pattern_load_dir(cases.pattern_dir(spec))
        eng.run(tds_from_pl=tds_from, max_iterations=50)
        s = cases.engine_summary(eng, len(spec["constraints"]))
        assert [list(x) for x in s["rows"]] == case["rows"]
        assert [list(x) for x in s["vertices"]] == case["vertices"]
        assert [list(x) for x in s["edges"]] == case["edges"]
        assert [[list(w) for w in sg] for sg in s["subgraphs"]] == case["subgraphs"]


def test_host_csr_upload_equals_slot_build(oracle, eng):
    edges = cases.random_multigraph(8, 3000, 20000, dup=0.15, loops=0.05)
    g = oracle.Graph.from_undirected(3000, edges)
    eng.graph_from_csr(g.rowptr, g.col, g.degree)
    assert np.array_equal(eng.graph_degree(), g.degree)
    rowptr, col = eng.graph_csr()
    assert np.array_equal(rowptr, g.rowptr) and np.array_equal(col, g.col)
    assert eng.graph_info()["n_slots_multi"] == g.n_slots_multi


def test_step_by_step_operators_match_the_driver_loop(oracle, eng):
    """pm_lcc / pm_nlcc called one by one, the way the reference main does, give the rows pm_run gives."""
    from fuzzypatternmatching_b200 import patterns as PT
    spec = PT.RMAT_LOG2_TREE
    edges = cases.random_multigraph(21, 400, 2400)
    labels = cases.random_labels(21, 400, [2, 3, 4, 5, 7])
    d = cases.pattern_dir(spec)
    g = oracle.Graph.from_undirected(400, edges)
    ref = oracle.Run(g, labels, oracle.Pattern(d), tds_from_pl=4)
    src, dst = cases.slots_of(edges)
    eng.graph_from_slots(400, src, dst)
    eng.labels_set(labels)
    eng.pattern_load_dir(d)
    eng.state_reset()
    rows, itr, init = [], 0, True
    while True:
        nf, counts = eng.label_propagation_pattern_matching_bsp(init)
        rows += [(itr, "LP", k, c[0], c[1]) for k, c in enumerate(counts)]
        init = False
        if itr == 0:
            nf = True
        if nf:
            nf = False
            for pl in range(5):
                found, deleted, c = eng.token_passing_pattern_matching(pl, tds=pl >= 4)
                rows.append((itr, "TP", pl, c[0], c[1]))
                if deleted:
                    nf = True
                    nf2, counts = eng.label_propagation_pattern_matching_bsp(False, nf)
                    nf = nf or nf2
                    rows += [(itr, "LP", k, c[0], c[1]) for k, c in enumerate(counts)]
        eng._chk(eng._lib.pm_end_iteration(eng._h, 0.0))
        itr += 1
        if not nf:
            break
    assert rows == ref.rows and itr == ref.iterations


def test_cli_writes_the_reference_result_tree(oracle, tmp_path):
    """generate_rmat + run_pattern_matching_beta (the C++ host driver) against the oracle's result tree."""
    import os
    import subprocess
    from fuzzypatternmatching_b200 import patterns as PT
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bindir = os.path.join(root, "fuzzypatternmatching_b200", "bin")
    gbase = str(tmp_path / "rmat")
    subprocess.check_call([os.path.join(bindir, "generate_rmat"), "-s", "17", "-r", "4", "-o", gbase],
                          stdout=subprocess.DEVNULL)
    pdir = str(tmp_path / "pattern")
    PT.write_pattern_dir(pdir, PT.RMAT_LOG2_TREE)
    out_gpu, out_ref = str(tmp_path / "gpu"), str(tmp_path / "ref")
    for o in (out_gpu, out_ref):
        os.makedirs(o)
        oracle.make_result_tree(o)
    subprocess.check_call([os.path.join(bindir, "run_pattern_matching_beta"), "-i", gbase, "-p", pdir, "-o", out_gpu],
                          stdout=subprocess.DEVNULL)
    g = oracle.Graph.rmat(17, 4)
    labels = g.labels_degree_log2()
    ref = oracle.Run(g, labels, oracle.Pattern(os.path.join(pdir, "0")), n_ranks=1, tds_from_pl=4)
    ref.write_results(out_ref)
    for rel in ("0/all_ranks_active_vertices/active_vertices_0", "0/all_ranks_active_edges/active_edges_0",
                "0/all_ranks_active_vertices_count/active_vertices_0", "0/all_ranks_active_edges_count/active_edges_0",
                "0/all_ranks_subgraphs/subgraphs_4_0"):
        a = sorted(open(os.path.join(out_gpu, rel)).read().splitlines())
        b = sorted(open(os.path.join(out_ref, rel)).read().splitlines())
        assert a == b and len(a) > 0, rel
    # time carrying files: same row keys
    key = lambda p: [",".join(l.split(",")[:3]) for l in open(p).read().splitlines()]  # noqa: E731
    assert key(os.path.join(out_gpu, "0/result_superstep")) == key(os.path.join(out_ref, "0/result_superstep"))
    a = open(os.path.join(out_gpu, "result_pattern_set")).read().split(",")
    b = open(os.path.join(out_ref, "result_pattern_set")).read().split(",")
    assert [x.strip() for x in a[:3] + a[4:]] == [x.strip() for x in b[:3] + b[4:]]


def test_large_label_values_take_the_gather_path(oracle, eng):
    """labels >= 64 (e.g. string hashes, VertexData = uint64_t in beta.cpp:263) disable the byte-label
    streams and the signature filter; the class-gather kernels must give the same answers."""
    import copy
    from fuzzypatternmatching_b200 import patterns as PT
    off = 1000003
    for base_spec, labelset, tds_from in ((PT.RMAT_LOG2_TREE, [2, 3, 4, 5, 7], 4), (PT.cycle4(1, 2, 3, 4), [1, 2, 3, 4], 1)):
        spec = copy.deepcopy(base_spec)
        spec["labels"] = [l + off for l in spec["labels"]]
        for seed in range(4):
            n, m = 300, 1500
            edges = cases.random_multigraph(seed + 40, n, m)
            labels = cases.random_labels(seed + 40, n, labelset) + np.uint64(off)
            _compare(oracle, eng, n, edges, labels, spec, tds_from)


def test_run_fuzzy_path_matches_oracle(oracle, eng):
    """unique-label LCC + cycle token passing over the unpruned adjacency (pm_run_fuzzy)."""
    from fuzzypatternmatching_b200 import patterns as PT
    nontrivial = 0
    for spec, labelset in ((PT.triangle(1, 2, 3), [1, 2, 3]), (PT.cycle4(1, 2, 3, 4), [1, 2, 3, 4])):
        d = cases.pattern_dir(spec)
        for seed in range(10):
            n, m = 60 + 10 * (seed % 4), 220 + 60 * (seed % 5)
            edges = cases.random_multigraph(seed, n, m)
            labels = cases.random_labels(seed, n, labelset)
            g = oracle.Graph.from_undirected(n, edges)
            ref = oracle.Run(g, labels, oracle.Pattern(d), fuzzy=True, max_iterations=50)
            src, dst = cases.slots_of(edges)
            eng.graph_from_slots(n, src, dst)
            eng.labels_set(labels)
            eng.pattern_load_dir(d)
            eng.run_fuzzy(max_iterations=50)
            assert eng.rows() == ref.rows and int(eng.summary["iterations"]) == ref.iterations
            v, t = eng.active_vertices()
            rv, rt = ref.active_vertices()
            assert np.array_equal(v, rv) and np.array_equal(t, rt)
            nontrivial += len(rv) > 0
    assert nontrivial >= 5
    # R-MAT with degree labels
    spec = PT.cycle4(5, 6, 7, 8)
    d = cases.pattern_dir(spec)
    g = oracle.Graph.rmat(17, 4)
    labels = g.labels_degree_log2()
    ref = oracle.Run(g, labels, oracle.Pattern(d), fuzzy=True)
    eng.graph_rmat(17, 4)
    eng.labels_degree_log2()
    eng.pattern_load_dir(d)
    eng.run_fuzzy()
    assert eng.rows() == ref.rows
    v, t = eng.active_vertices()
    rv, rt = ref.active_vertices()
    assert np.array_equal(v, rv) and np.array_equal(t, rt) and len(rv) > 0
    # ... and against what the reference's own run_pattern_matching driver wrote for this input (vertex counts per row; one
    # template vertex index per vertex)
    _, golden = cases.reference_golden_load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_runs",
                                                         "fuzzy_rmat17_cycle4_5678.json"))
    assert [(a, b, c, nv, 0) for a, b, c, nv, _ in eng.rows()] == golden["rows"]
    assert sorted((int(a), int(b).bit_length() - 1) for a, b in zip(v, t)) == golden["vertices"]


def test_edge_cases_match_oracle(oracle, eng):
    """Degenerate and ragged inputs (tests/cases.py: edge_cases)."""
    for name, n, edges, labels, spec, tds_from in cases.edge_cases():
        ref = _compare(oracle, eng, n, edges, labels, spec, tds_from)
        if name.startswith("one_triangle"):
            assert ref.rows[-1][3] == 3 and ref.rows[-1][4] == 6, name


@pytest.mark.parametrize("which", ["tree", "cycle4"])
def test_full_size_properties(eng, which):
    """BASELINE configs[1] / [2] sizes (R-MAT scale 25), where the oracle is out of reach: size-independent
    properties of the reference's fixed point — determinism, symmetry of the active edge set, every edge
    joins two active vertices, the local constraint holds at every active vertex, one more LCC call removes
    nothing, and every enumerated walk runs over active edges."""
    from fuzzypatternmatching_b200 import patterns as PT
    spec, tds_from = (PT.RMAT_LOG2_TREE, 4) if which == "tree" else (PT.cycle4(5, 6, 7, 8), 1)
    d = cases.pattern_dir(spec)
    eng.graph_rmat(25, 1024)
    eng.labels_degree_log2()
    labels = eng.labels_get()
    eng.pattern_load_dir(d)
    eng.run(tds_from_pl=tds_from)
    rows = eng.rows()
    v, t = eng.active_vertices()
    e = eng.active_edges()
    sub = [eng.subgraphs(pl) for pl in range(len(spec["constraints"]))]
    assert rows[-1][3] == len(v) and rows[-1][4] == len(e)
    # vertex and edge counts never grow along the rows
    assert all(rows[i + 1][3] <= rows[i][3] and rows[i + 1][4] <= rows[i][4] for i in range(len(rows) - 1))
    # determinism: a second search gives the same rows and sets
    eng.run(tds_from_pl=tds_from)
    v2, t2 = eng.active_vertices()
    assert eng.rows() == rows and np.array_equal(v, v2) and np.array_equal(t, t2) and np.array_equal(e, eng.active_edges())
    # the fixed point: one more LCC call neither removes a vertex nor an edge
    nf, counts = eng.label_propagation_pattern_matching_bsp(False)
    assert not nf and all(c[0] == len(v) and c[1] == len(e) for c in counts)
    if len(v) == 0:
        return
    assert np.all(t != 0) and np.all(np.diff(v.astype(np.int64)) > 0)
    mask = dict(zip(v.tolist(), t.tolist()))
    es = set(map(tuple, e.tolist()))
    assert len(es) == len(e)
    nbrs = {}
    for a, b in es:
        assert (b, a) in es and a in mask and b in mask
        nbrs.setdefault(a, []).append(b)
    # template side
    n_t = len(spec["labels"])
    tn = [[] for _ in range(n_t)]
    for a, b in spec["edges"]:
        tn[a].append(b)
        tn[b].append(a)
    for x, m in mask.items():
        heard = 0
        for y in nbrs.get(x, []):
            heard |= mask[y]
        for p in range(n_t):
            if (m >> p) & 1:
                assert int(labels[x]) == spec["labels"][p]
                assert all((heard >> q) & 1 for q in tn[p]), "local constraint violated"
    # enumerated walks (TDS constraints) run over active edges and carry the walk's labels
    for pl, k in enumerate(spec["constraints"]):
        if not k.get("tds") or pl < tds_from:
            continue
        w = k["walk"]
        for r in sub[pl].tolist():
            assert len(r) == len(w)
            for i, x in enumerate(r):
                assert int(labels[x]) == spec["labels"][w[i]]
            for i in range(len(r) - 1):
                assert (r[i], r[i + 1]) in es


def _cli_compare(out_gpu, out_ref, ps, ranks, subgraph_pls):
    import os
    rels = []
    for r in range(ranks):
        rels += ["%d/all_ranks_active_vertices/active_vertices_%d" % (ps, r), "%d/all_ranks_active_edges/active_edges_%d" % (ps, r),
                 "%d/all_ranks_active_vertices_count/active_vertices_%d" % (ps, r),
                 "%d/all_ranks_active_edges_count/active_edges_%d" % (ps, r)]
        rels += ["%d/all_ranks_subgraphs/subgraphs_%d_%d" % (ps, pl, r) for pl in subgraph_pls]
    total = 0
    for rel in rels:
        a = sorted(open(os.path.join(out_gpu, rel)).read().splitlines())
        b = sorted(open(os.path.join(out_ref, rel)).read().splitlines())
        assert a == b, rel
        total += len(a)
    return total


def test_cli_ingest_vertex_metadata_and_pattern_set(oracle, tmp_path):
    """ingest_edge_list -u 1 + run_pattern_matching_beta -v -e over a pattern set <p>/0, <p>/1: the
    result trees of both elements against the oracle's, one result_pattern_set row per element."""
    import os
    import subprocess
    from fuzzypatternmatching_b200 import patterns as PT
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bindir = os.path.join(root, "fuzzypatternmatching_b200", "bin")
    n, m = 400, 2600
    edges = cases.random_multigraph(31, n, m)
    labels = cases.random_labels(31, n, [1, 2, 3, 4])
    labels[n - 1] = 3
    edges.append((n - 1, 0))  # the largest id appears in the edge list: ingest sizes the graph by it
    with open(tmp_path / "edges_a.txt", "w") as f:
        for a, b in edges[:1500]:
            f.write("%d %d\n" % (a, b))
    with open(tmp_path / "edges_b.txt", "w") as f:
        for a, b in edges[1500:]:
            f.write("%d %d 1\n" % (a, b))
    gbase = str(tmp_path / "graph")
    subprocess.check_call([os.path.join(bindir, "ingest_edge_list"), "-o", gbase, "-u", "1", "-d", "64",
                           str(tmp_path / "edges_a.txt"), str(tmp_path / "edges_b.txt")], stdout=subprocess.DEVNULL)
    meta = tmp_path / "meta"
    meta.mkdir()
    for part in range(2):
        with open(meta / ("vlabel_%d" % part), "w") as f:
            for v in range(part, n, 2):
                if labels[v] != 0:
                    f.write("%d %d\n" % (v, labels[v]))
    with open(meta / "elabel_0", "w") as f:
        for a, b in edges[:50]:
            f.write("%d %d 7\n" % (a, b))
    pdir = str(tmp_path / "pattern")
    specs = [PT.triangle(1, 2, 3), PT.cycle4(1, 2, 3, 4)]
    for ps, spec in enumerate(specs):
        PT.write_pattern_dir(pdir, spec, ps=ps)
    out_gpu = str(tmp_path / "gpu")
    os.makedirs(out_gpu)
    for ps in range(2):
        oracle.make_result_tree(out_gpu, ps)
    subprocess.check_call([os.path.join(bindir, "run_pattern_matching_beta"), "-i", gbase, "-p", pdir, "-o", out_gpu,
                           "-v", str(meta / "vlabel"), "-e", str(meta / "elabel"), "-t", "1"], stdout=subprocess.DEVNULL)
    g = oracle.Graph.from_undirected(n, edges)
    for ps, spec in enumerate(specs):
        out_ref = str(tmp_path / ("ref%d" % ps))
        os.makedirs(out_ref)
        oracle.make_result_tree(out_ref)
        ref = oracle.Run(g, labels, oracle.Pattern(os.path.join(pdir, str(ps))), n_ranks=1, tds_from_pl=1)
        ref.write_results(out_ref)
        os.rename(os.path.join(out_ref, "0"), os.path.join(out_ref, "x"))
        os.rename(os.path.join(out_ref, "x"), os.path.join(out_ref, str(ps)))
        assert _cli_compare(out_gpu, out_ref, ps, 1, [1]) > 0
    rows = open(os.path.join(out_gpu, "result_pattern_set")).read().splitlines()
    assert [r.split(",")[0].strip() for r in rows] == ["0", "1"]


@pytest.mark.parametrize("threshold", [0, 64])
def test_cli_two_ranks_write_every_rank_file(oracle, tmp_path, threshold):
    """run_pattern_matching_beta -n 2: one process per GPU, per-rank files *_0 and *_1 equal to a 2-rank oracle run.
    threshold: generate_rmat -d — hubs above it appear in their CONTROLLER's files (delegate id mod ranks)."""
    import os
    import subprocess
    import torch
    from fuzzypatternmatching_b200 import patterns as PT
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bindir = os.path.join(root, "fuzzypatternmatching_b200", "bin")
    gbase = str(tmp_path / "rmat")
    subprocess.check_call([os.path.join(bindir, "generate_rmat"), "-s", "17", "-r", "4", "-o", gbase] +
                          (["-d", str(threshold)] if threshold else []), stdout=subprocess.DEVNULL)
    pdir = str(tmp_path / "pattern")
    PT.write_pattern_dir(pdir, PT.RMAT_LOG2_TREE)
    out_gpu, out_ref = str(tmp_path / "gpu"), str(tmp_path / "ref")
    for o in (out_gpu, out_ref):
        os.makedirs(o)
        oracle.make_result_tree(o)
    subprocess.check_call([os.path.join(bindir, "run_pattern_matching_beta"), "-i", gbase, "-p", pdir, "-o", out_gpu, "-n", "2"],
                          stdout=subprocess.DEVNULL, timeout=600)
    g = oracle.Graph.rmat(17, 4)
    ref = oracle.Run(g, g.labels_degree_log2(), oracle.Pattern(os.path.join(pdir, "0")), n_ranks=2, tds_from_pl=4,
                     delegate_threshold=threshold)
    ref.write_results(out_ref)
    assert _cli_compare(out_gpu, out_ref, 0, 2, [4]) > 0


@pytest.mark.parametrize("name,spec,labelset,tds_from", cases.APPROX_SPECS, ids=[s[0] for s in cases.APPROX_SPECS])
def test_approximate_patterns_match_oracle(oracle, eng, name, spec, labelset, tds_from):
    """optional template edges, vertex_min_optional_edge_count and the mandatory / optional coverage test
    (approximate_pattern_matching/local_constraint_checking.hpp:641-651, 1062-1113) — small-label signature path,
    large-label gather path and an R-MAT graph."""
    import copy
    nontrivial = 0
    for seed in range(10):
        n, m = 60 + 10 * (seed % 4), 160 + 40 * (seed % 5)
        edges = cases.random_multigraph(seed + 500, n, m)
        labels = cases.random_labels(seed + 500, n, labelset)
        ref = _compare(oracle, eng, n, edges, labels, spec, tds_from)
        nontrivial += ref.rows[-1][3] > 0
    assert nontrivial >= (0 if name == "impossible_min_optional" else 3)
    big = copy.deepcopy(spec)
    big["labels"] = [l + 1000 for l in spec["labels"]]
    for seed in range(3):
        edges = cases.random_multigraph(seed + 600, 200, 700)
        labels = cases.random_labels(seed + 600, 200, labelset) + np.uint64(1000)
        _compare(oracle, eng, 200, edges, labels, big, tds_from)


def test_approximate_pattern_on_rmat(oracle, eng):
    g = oracle.Graph.rmat(17, 4)
    labels = g.labels_degree_log2()
    eng.graph_rmat(17, 4)
    eng.labels_degree_log2()
    for min_optional in ({}, {0: 1}, {2: 1}):
        spec = {"labels": [5, 6, 7, 8], "edges": [(0, 1), (1, 2), (2, 3), (0, 3), (0, 2)], "optional_edges": [(0, 2)],
                "diameter": 3, "constraints": [{"walk": [0, 1, 2, 3, 0], "cycle": True}]}
        if min_optional:
            spec["min_optional"] = min_optional
        d = cases.pattern_dir(spec)
        pat = oracle.Pattern(d)
        ref = oracle.Run(g, labels, pat, tds_from_pl=-1)
        eng.pattern_load_dir(d)
        eng.run(tds_from_pl=-1)
        got, want = cases.engine_summary(eng, pat.n_constraints), cases.run_summary(ref)
        for k in ("rows", "iterations", "vertices", "edges"):
            assert got[k] == want[k], (min_optional, k)
        assert min_optional == {2: 1} or len(want["vertices"]) > 0


def test_hub_class_template_on_rmat_scale20(oracle, eng):
    """A template over hub classes (degree labels 12, 14, 16: rows of 2^11 .. 2^16 slots) on R-MAT scale 20: the
    CTA-per-row kernels (rows above 4096 slots) carry the first scan, the renaming scan and the later scans; rows, final
    sets and the number of enumerated walks against the oracle."""
    from fuzzypatternmatching_b200 import patterns as PT
    g = oracle.Graph.rmat(20, 4)
    labels = g.labels_degree_log2()
    eng.graph_rmat(20, 4)
    eng.labels_degree_log2()
    assert eng.graph_info()["max_degree"] > 4096
    spec = PT.triangle(12, 14, 16)
    d = cases.pattern_dir(spec)
    ref = oracle.Run(g, labels, oracle.Pattern(d), tds_from_pl=1, keep_subgraphs=False)
    assert not ref.hazards[:5].any()
    eng.pattern_load_dir(d)
    s = eng.run(tds_from_pl=1, keep_subgraphs=False)
    assert eng.rows() == ref.rows and int(s["iterations"]) == ref.iterations
    v, t = eng.active_vertices()
    rv, rt = ref.active_vertices()
    assert np.array_equal(v, rv) and np.array_equal(t, rt) and len(rv) > 100
    assert np.array_equal(eng.active_edges(), ref.active_edges)
    assert int(s["path_count"]) == ref.path_count > 0
    assert eng.kernel_stats(2)["launches"] > 0  # the CTA-per-row class ran
    # ... and against what THE REFERENCE ITSELF wrote for this graph and template (enumeration walk moved to constraint 4,
    # where its driver starts template-driven search: same final sets, same number of enumerated walks)
    _, golden = cases.reference_golden_load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_runs",
                                                         "rmat20_hubs_triangle_12_14_16.json"))
    assert sorted(zip(v.tolist(), t.tolist())) == golden["vertices"]
    assert sorted(map(tuple, eng.active_edges().tolist())) == golden["edges"]
    assert int(s["path_count"]) == golden["subgraphs_digest"][4]["count"]
