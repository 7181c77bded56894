"""CPU tests that PIN the oracle (the reference ships no golden vectors for this
path, SURVEY §4/§8c, so the pins are independent restatements and hand-derived
answers): Mersenne-Twister stream vs numpy's MT19937, R-MAT edges vs a pure-Python
restatement, label formula, hand-computed LCC/NLCC outcomes, and the committed
golden fixtures."""
import json
import os

import numpy as np
import pytest

from fuzzypatternmatching_b200 import patterns as PT
from tests import cases

HERE = os.path.dirname(os.path.abspath(__file__))


# ---- independent pure-Python restatement of the R-MAT stream -----------------------
def _h16(a):
    a &= 0xFFFF
    a = ((a + 0x5D16) + (a << 6)) & 0xFFFF
    a = ((a ^ 0xC23C) ^ (a >> 9)) & 0xFFFF
    a = ((a + 0x67B1) + (a << 5)) & 0xFFFF
    a = ((a + 0x646C) ^ (a << 7)) & 0xFFFF
    a = ((a + 0x46C5) + (a << 3)) & 0xFFFF
    a = ((a ^ 0x4F09) ^ (a >> 8)) & 0xFFFF
    return a


def _hash_nbits_lt32(x, n):
    k = n - 16
    order = list(range(k + 1)) + list(range(k, -1, -1))
    for i in order:
        h = _h16((x >> i) & 0xFFFF)
        x = (x & ~(0xFFFF << i)) | (h << i)
    return x


def _py_rmat(scale, rank, n_edges):
    bg = np.random.MT19937()
    bg._legacy_seeding(5489 + 3 * rank)  # init_genrand == std::mt19937(seed) == boost::mt19937(seed)
    raw = bg.random_raw(n_edges * 5 * scale)
    out, k = [], 0
    for _ in range(n_edges):
        a, b, c, d = 0.57, 0.19, 0.19, 0.05
        u = v = 0
        step = (1 << scale) >> 1
        for _j in range(scale):
            p = float(raw[k]) * 2.0 ** -32
            if p < a:
                pass
            elif p < a + b:
                v += step
            elif p < a + b + c:
                u += step
            else:
                u += step
                v += step
            step >>= 1
            a *= 0.9 + 0.2 * (float(raw[k + 1]) * 2.0 ** -32)
            b *= 0.9 + 0.2 * (float(raw[k + 2]) * 2.0 ** -32)
            c *= 0.9 + 0.2 * (float(raw[k + 3]) * 2.0 ** -32)
            d *= 0.9 + 0.2 * (float(raw[k + 4]) * 2.0 ** -32)
            k += 5
            s = a + b + c + d
            a /= s
            b /= s
            c /= s
            d = 1.0 - a - b - c
        out.append((_hash_nbits_lt32(u, scale), _hash_nbits_lt32(v, scale)))
    return out


def test_mt19937_first_output_is_the_textbook_value():
    bg = np.random.MT19937()
    bg._legacy_seeding(5489)
    assert int(bg.random_raw(1)[0]) == 3499211612  # mt19937 default-seed known answer


@pytest.mark.parametrize("scale,rank", [(17, 0), (17, 3), (21, 0), (21, 2)])
def test_rmat_stream_matches_independent_restatement(oracle, scale, rank):
    got = oracle.rmat_stream(scale, rank, 64).tolist()
    assert [tuple(x) for x in got] == _py_rmat(scale, rank, 64)


def test_rmat_stream_golden(oracle):
    gold = json.load(open(os.path.join(HERE, "golden", "rmat_kat.json")))
    for key, edges in gold.items():
        scale, rank = [int(x) for x in key.split("_")]
        assert oracle.rmat_stream(scale, rank, len(edges)).tolist() == edges


def test_rmat_stream_and_hash_equal_the_reference_generators_output(oracle):
    """tests/golden/rmat_reference_generator.json was written by the reference's OWN rmat_edge_generator.hpp and
    detail/hash.hpp (oracle/_ref/rmat_edge_dump via oracle/make_reference_golden.py): first edges of generating ranks of
    BASELINE's graphs (scale 21 / 4 ranks; scales 25, 26, 28 / 1024 ranks as the bench generates them) and hash_nbits for
    every supported width up to 32.  tests/test_oracle_vs_reference.py compares whole streams where the binary is present."""
    gold = json.load(open(os.path.join(HERE, "golden", "rmat_reference_generator.json")))
    assert len(gold["streams"]) >= 8
    for s in gold["streams"]:
        assert oracle.rmat_stream(s["scale"], s["rank"], len(s["edges"])).tolist() == s["edges"], (s["scale"], s["rank"], s["ranks"])
    for h in gold["hash_nbits"]:
        assert [oracle.hash_nbits(x, h["n"]) for x in h["x"]] == h["hash"], h["n"]


def test_hash_nbits_is_a_permutation_of_17_bits(oracle):
    xs = {oracle.hash_nbits(x, 17) for x in range(0, 1 << 17, 7)}
    assert len(xs) == len(range(0, 1 << 17, 7)) and max(xs) < (1 << 17)
    assert oracle.hash_nbits(12345, 17) == _hash_nbits_lt32(12345, 17)


def test_rmat_graph_chunked_generation_equals_stream(oracle):
    g = oracle.Graph.rmat(17, 4, threads=3)
    deg = np.zeros(1 << 17, dtype=np.uint64)
    for r in range(4):
        e = oracle.rmat_stream(17, r, (1 << 17) * 16 // 4)
        np.add.at(deg, e[:, 0], 1)
        np.add.at(deg, e[:, 1], 1)
    assert np.array_equal(deg, g.degree)
    assert g.n_slots_multi == 2 * 16 * (1 << 17)


def test_degree_labels_are_bit_lengths(oracle):
    # ceil(log2(d+1)) evaluated in double == bit length of d (vertex_data_db_degree.hpp:109)
    n = 5000
    edges = [(0, i) for i in range(1, 40)] + [(1, 2), (1, 2), (100, 100)]
    g = oracle.Graph.from_undirected(n, edges)
    lab = g.labels_degree_log2()
    assert all(int(l) == int(d).bit_length() for l, d in zip(lab, g.degree))
    assert g.degree[100] == 2 and g.degree[1] == 3 and g.degree[0] == 39  # self loop counts twice, duplicates count


# ---- hand-derived LCC / NLCC answers -----------------------------------------------------
def _run(oracle, n, edges, labels, spec, tds_from=None, **kw):
    pat = oracle.Pattern(cases.pattern_dir(spec))
    g = oracle.Graph.from_undirected(n, edges)
    return oracle.Run(g, np.asarray(labels, dtype=np.uint64), pat,
                      tds_from_pl=PT.tds_from_pl(spec) if tds_from is None else tds_from, **kw)


def test_triangle_present_and_absent(oracle):
    spec = PT.triangle(1, 2, 3)
    # vertices 0,1,2 form a 1-2-3 triangle; 3,4,5 form an open path 1-2-3 (no closing edge)
    edges = [(0, 1), (1, 2), (0, 2), (3, 4), (4, 5)]
    r = _run(oracle, 6, edges, [1, 2, 3, 1, 2, 3], spec)
    v, t = r.active_vertices()
    assert v.tolist() == [0, 1, 2] and t.tolist() == [1, 2, 4]
    assert sorted(map(tuple, r.active_edges.tolist())) == [(0, 1), (0, 2), (1, 0), (1, 2), (2, 0), (2, 1)]
    # LCC alone removes the open path 3-4-5: in superstep 0 the end points 3 and 5 miss one
    # required template neighbour each and leave; the middle vertex 4 heard both and stays,
    # still holding its two edges (4 vertices, 6 + 2 edges); in superstep 1 it hears nobody.
    assert r.rows[0] == (0, "LP", 0, 4, 8)
    assert r.rows[1] == (0, "LP", 1, 3, 6)
    assert sorted(map(tuple, r.subgraphs[1].tolist())) == [(0, 1, 2, 0)]


def test_lcc_keeps_a_hexagon_that_only_nlcc_can_reject(oracle):
    spec = PT.triangle(1, 2, 3)
    # 6-cycle labelled 1,2,3,1,2,3: locally every vertex sees both other labels, but there is no triangle
    edges = [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 0)]
    r = _run(oracle, 6, edges, [1, 2, 3, 1, 2, 3], spec)
    assert r.rows[0][3:] == (6, 12) and r.rows[1][3:] == (6, 12)       # LCC keeps everything
    tp0 = [x for x in r.rows if x[1] == "TP" and x[2] == 0][0]
    # both label-1 sources fail the cycle check and leave the map; the TP row counts the edge
    # maps of the 4 vertices still in the map (their entries towards the dead sources included)
    assert tp0[3:] == (4, 8)
    assert r.in_map.sum() == 0 and len(r.active_edges) == 0             # and the rest unravels
    assert r.iterations == 2


def test_parallel_edges_and_self_loops(oracle):
    spec = PT.triangle(1, 2, 3)
    edges = [(0, 1), (0, 1), (0, 1), (1, 2), (0, 2), (2, 2), (0, 0)]
    r = _run(oracle, 3, edges, [1, 2, 3], spec)
    # edge maps are keyed by neighbour: duplicates collapse, self loops are dropped (labels differ)
    assert r.rows[0][3:] == (3, 6)
    assert len(r.active_edges) == 6


def test_path_constraint_needs_a_second_end_point(oracle):
    # template: path 1 - 2 - 1 (two template vertices share label 1); the path constraint
    # 0 -> 1 -> 2 must end at a vertex different from its source
    spec = {"labels": [1, 2, 1], "edges": [(0, 1), (1, 2)], "diameter": 2,
            "constraints": [{"walk": [0, 1, 2]}, {"walk": [2, 1, 0]}]}
    # star with ONE leaf: 0(label 1) - 1(label 2): the only length-2 walk returns to the source
    r = _run(oracle, 2, [(0, 1)], [1, 2], spec, tds_from=-1)
    assert r.in_map.sum() == 0
    # star with TWO leaves: both leaves can be the two ends
    r = _run(oracle, 3, [(0, 1), (1, 2)], [1, 2, 1], spec, tds_from=-1)
    v, t = r.active_vertices()
    assert v.tolist() == [0, 1, 2] and t.tolist() == [0b101, 0b010, 0b101]


def test_planted_tree_is_found_exactly_once(oracle):
    spec = PT.RMAT_LOG2_TREE
    labels = spec["labels"] + [9, 9, 9]
    edges = list(spec["edges"]) + [(7, 8), (8, 9), (0, 7)]
    r = _run(oracle, 10, edges, labels, spec)
    v, _ = r.active_vertices()
    assert v.tolist() == list(range(7))
    assert len(r.active_edges) == 12
    # template vertices 0/4 (label 3) and 2/6 (label 7) are interchangeable only where the tree allows it
    assert r.subgraphs[4].tolist() == [[0, 1, 2, 1, 3, 5, 4, 5, 6]]
    assert r.hazards[:5].tolist() == [0, 0, 0, 0, 0]


def test_reference_grid_graph_fixture(oracle):
    """The one fixture the reference holds for this path's graph store (test/include/input_graph.hpp:8-68,
    test/test_delegate_graph_static.cpp:146-152): 3 x 5 grid, expected degrees, CSR offsets and hubs."""
    slots = cases.grid_graph_slots()
    assert len(slots) == 44
    src = np.array([a for a, _ in slots], dtype=np.uint64)
    dst = np.array([b for _, b in slots], dtype=np.uint64)
    g = oracle.Graph.from_slots(15, src, dst)
    assert g.degree.tolist() == cases.GRID_DEGREE
    assert g.rowptr.tolist() == cases.GRID_OFFSET
    assert [v for v in range(15) if g.degree[v] >= 4] == cases.GRID_HUBS_AT_THRESHOLD_4
    assert g.labels_degree_log2().tolist() == [2, 2, 2, 2, 2, 2, 3, 3, 3, 2, 2, 2, 2, 2, 2]


def test_grid_graph_lcc_by_hand(oracle):
    """LCC on the reference's grid graph, labels = degree, template path 2 - 3 - 4 (two supersteps, no NLCC).
    Superstep 0: the degree-3 vertices 2 and 12 have no degree-2 neighbour, enter the map (they hear 7) and leave
    it.  Superstep 1: vertex 7 only heard 2 and 12 and leaves.  Edge maps hold template-adjacent label pairs only."""
    slots = cases.grid_graph_slots()
    spec = {"labels": [2, 3, 4], "edges": [(0, 1), (1, 2)], "diameter": 2, "constraints": []}
    pat = oracle.Pattern(cases.pattern_dir(spec))
    src = np.array([a for a, _ in slots], dtype=np.uint64)
    dst = np.array([b for _, b in slots], dtype=np.uint64)
    g = oracle.Graph.from_slots(15, src, dst)
    r = oracle.Run(g, np.array(cases.GRID_DEGREE, dtype=np.uint64), pat, tds_from_pl=-1)
    assert r.rows == [(0, "LP", 0, 13, 30), (0, "LP", 1, 12, 28)]
    v, t = r.active_vertices()
    assert v.tolist() == [0, 1, 3, 4, 5, 6, 8, 9, 10, 11, 13, 14]
    assert t.tolist() == [1, 2, 2, 1, 2, 4, 4, 2, 1, 2, 2, 1]
    e = set(map(tuple, r.active_edges.tolist()))
    assert (6, 7) not in e and (1, 2) not in e and (6, 1) in e and (1, 6) in e and len(e) == 28


def test_twin_template_bit_is_resurrected_by_hand(oracle):
    """SURVEY A.6 #4 on two vertices v(1) - u(2), template path 1 - 2 - 1 (twins 0 and 2), TDS walk 0 1 2.
    The walk fails (no second label-1 vertex): bit 0 of T_arr(v) is cleared.  In the next LCC superstep u hears
    only {2} and leaves, while v's post step rewrites T_arr from T_state = {0, 2}: the bit is back for one
    superstep, then v hears nobody and leaves."""
    r = _run(oracle, 2, [(0, 1)], [1, 2], cases.TWIN, tds_from=0, max_iterations=10)
    assert r.rows == [(0, "LP", 0, 2, 2), (0, "LP", 1, 2, 2), (0, "TP", 0, 2, 2),
                      (0, "LP", 0, 1, 1), (0, "LP", 1, 0, 0), (1, "LP", 0, 0, 0), (1, "LP", 1, 0, 0)]
    assert r.iterations == 2 and r.hazards[3] == 1 and not r.hazards[:3].any()


@pytest.mark.parametrize("name,spec,labelset,tds_from,div,counter", cases.QUIRK_SPECS, ids=[s[0] for s in cases.QUIRK_SPECS])
def test_quirk_inputs_equal_literal_execution(oracle, name, spec, labelset, tds_from, div, counter):
    """The quirk inputs (A.6 #4 resurrection, A.6 #11 flag set outside LCC): the oracle equals the literal
    transliteration of the reference visitors under randomised delivery, and the quirk really occurs."""
    from oracle.ref_literal import LiteralRun, PatternFiles
    d = cases.pattern_dir(spec)
    fired = 0
    for seed, n, m in cases.quirk_inputs(name, div, range(10)):
        n, m = min(n, 50), m * 2 // 3
        edges = cases.random_multigraph(seed, n, m)
        labels = cases.random_labels(seed, n, labelset)
        g = oracle.Graph.from_undirected(n, edges)
        r = oracle.Run(g, labels, oracle.Pattern(d), tds_from_pl=tds_from, max_iterations=50)
        assert not r.hazards[:3].any() and not r.hazards[4]
        fired += int(r.hazards[counter] > 0)
        want = cases.run_summary(r)
        slots = []
        for a, b in edges:
            slots += [(a, b), (b, a)]
        for s in (0, 1):
            L = LiteralRun(n, slots, labels.tolist(), PatternFiles(d), seed=seed * 31 + s, tds_from_pl=tds_from)
            assert not L.errors
            assert L.rows == want["rows"] and L.iterations == want["iterations"]
            assert L.final_vertices() == want["vertices"] and L.final_edges() == want["edges"]
    assert fired >= 2


def test_reference_example_pattern_fixture_is_verbatim():
    """tests/golden/rmat_log2_tree_pattern/0 is the reference's examples/rmat_log2_tree_pattern/0 (data files,
    BASELINE configs[0]); where the reference tree is present (not on the GPU box) the copy is checked byte for
    byte, and the in-code spec PT.RMAT_LOG2_TREE must describe the same template."""
    gold = os.path.join(HERE, "golden", "rmat_log2_tree_pattern", "0")
    ref = "/root/reference/examples/rmat_log2_tree_pattern/0"
    names = ["pattern_edge", "pattern_edge_data", "pattern_nlc", "pattern_non_local_constraint", "pattern_stat",
             "pattern_vertex", "pattern_vertex_data"]
    if os.path.isdir(ref):
        for n in names:
            assert open(os.path.join(gold, n), "rb").read() == open(os.path.join(ref, n), "rb").read(), n
    mine = cases.pattern_dir(PT.RMAT_LOG2_TREE)
    num = lambda p: [[int(x) for x in l.replace(":", " ").split()] for l in open(p).read().splitlines() if l.strip()]  # noqa: E731
    for n in ("pattern_edge", "pattern_vertex_data", "pattern_nlc"):
        assert num(os.path.join(gold, n)) == num(os.path.join(mine, n)), n
    assert num(os.path.join(gold, "pattern_non_local_constraint"))[4] == num(os.path.join(mine, "pattern_non_local_constraint"))[4]


def test_lcc_only_runs_to_a_vertex_fixed_point(oracle):
    spec = PT.RMAT_LOG2_TREE
    edges = cases.random_multigraph(4, 80, 300)
    labels = cases.random_labels(4, 80, [2, 3, 4, 5, 7])
    r = _run(oracle, 80, edges, labels, spec, lcc_only=True)
    assert all(k == "LP" for _, k, _, _, _ in r.rows)
    d = spec["diameter"]
    assert r.rows[-1][3] == r.rows[-1 - d][3] if len(r.rows) > d else True


def test_golden_fixtures(oracle):
    gold = json.load(open(os.path.join(HERE, "golden", "runs.json")))
    specs = {s[0]: s for s in cases.SPECS}
    for case in gold:
        _, spec, labelset, tds_from = specs[case["spec"]]
        edges = cases.random_multigraph(case["seed"], case["n"], case["m"])
        labels = cases.random_labels(case["seed"], case["n"], labelset)
        r = _run(oracle, case["n"], edges, labels, spec, tds_from=tds_from, max_iterations=50)
        s = cases.run_summary(r)
        assert [list(x) for x in s["rows"]] == case["rows"]
        assert [list(x) for x in s["vertices"]] == case["vertices"]
        assert [list(x) for x in s["edges"]] == case["edges"]
        assert [[list(w) for w in sg] for sg in s["subgraphs"]] == case["subgraphs"]
