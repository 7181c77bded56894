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


def planted(seed, n, m, spec, labelset, copies=3):
    """random multigraph + labels with `copies` instances of the template planted on random distinct vertices, so
    that templates which random graphs of this size never contain (6-cycle with chords) end non-trivially"""
    rng = random.Random(seed * 104729 + 7)
    edges = random_multigraph(seed, n, m)
    labels = random_labels(seed, n, labelset)
    k = len(spec["labels"])
    for _ in range(copies):
        vs = rng.sample(range(n), k)
        for i, v in enumerate(vs):
            labels[v] = spec["labels"][i]
        edges += [(vs[a], vs[b]) for a, b in spec["edges"]]
    return edges, labels


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


# Inputs on which the reference's QUIRKS are exercised; deterministic in the reference, so the GPU must
# reproduce them bit for bit.  (name, spec, labelset, tds_from_pl, edge divisor, index of the oracle counter that must fire)
#  * twin templates: template vertices 0 and 2 share label and neighbourhood.  NLCC clears a bit in T_arr only; the
#    next LCC post step rewrites T_arr from T_state and RESURRECTS it (A.6 #4, oracle counter [3]).
#  * bowtie (two triangles sharing an edge) with interleave_lcc off: a successful cycle token flags E_s[sender]
#    outside LCC; the sender is deactivated by the second constraint before the next LCC call, and the flagged edge
#    survives that call's first post step (A.6 #11, oracle counter [5]).
TWIN = {"labels": [1, 2, 1], "edges": [(0, 1), (1, 2)], "diameter": 2, "constraints": [{"walk": [0, 1, 2], "tds": True}]}
TWIN_BRANCH = {"labels": [1, 2, 1, 3], "edges": [(0, 1), (1, 2), (1, 3)], "diameter": 3,
               "constraints": [{"walk": [0, 1, 2], "tds": True}]}
BOWTIE = {"labels": [1, 2, 3, 4], "edges": [(0, 1), (1, 2), (0, 2), (1, 3), (2, 3)], "diameter": 2,
          "constraints": [{"walk": [0, 1, 2, 0], "cycle": True, "interleave": False},
                          {"walk": [2, 3, 1, 2], "cycle": True, "interleave": False}]}
QUIRK_SPECS = [
    ("twin", TWIN, [1, 2], 0, 3, 3),
    ("twin_branch", TWIN_BRANCH, [1, 2, 3], 0, 3, 3),
    ("bowtie_no_interleave", BOWTIE, [1, 2, 3, 4], -1, 1, 5),
]


# Approximate matching (approximate_pattern_matching/local_constraint_checking.hpp:641-651, 1062-1113):
# optional template edges and minimum optional edge counts.  (name, spec, labelset, tds_from_pl)
APPROX_SPECS = [
    # a square with one optional diagonal: the diagonal's end points need not see each other
    ("square_optional_diagonal", {"labels": [1, 2, 3, 4], "edges": [(0, 1), (1, 2), (2, 3), (0, 3), (0, 2)],
                                  "optional_edges": [(0, 2)], "diameter": 3,
                                  "constraints": [{"walk": [0, 1, 2, 3, 0], "cycle": True}]}, [1, 2, 3, 4], -1),
    # a star whose centre must keep both optional leaves once a minimum optional count is set, next to a mandatory one
    ("star_min_optional", {"labels": [1, 2, 3, 4], "edges": [(0, 1), (0, 2), (0, 3)], "optional_edges": [(0, 2), (0, 3)],
                           "min_optional": {0: 2}, "diameter": 2, "constraints": []}, [1, 2, 3, 4], -1),
    # a path whose middle vertex has only optional edges (nothing is required of it locally)
    ("path_all_optional_middle", {"labels": [1, 2, 3, 2], "edges": [(0, 1), (1, 2), (2, 3)],
                                  "optional_edges": [(0, 1), (1, 2)], "diameter": 3, "constraints": []}, [1, 2, 3], -1),
    # a minimum optional count that cannot be met: the vertex can never stay
    ("impossible_min_optional", {"labels": [1, 2, 3], "edges": [(0, 1), (1, 2)], "optional_edges": [(1, 2)],
                                 "min_optional": {1: 2}, "diameter": 2, "constraints": []}, [1, 2, 3], -1),
]


def quirk_inputs(name, divisor, seeds=range(24)):
    for seed in seeds:
        n, m = 60 + 10 * (seed % 4), (220 + 60 * (seed % 5)) // divisor
        if name.startswith("bowtie"):
            seed += 130
        yield seed, n, m


# the 3 x 5 grid graph the reference's own graph-construction test is written against
# (/root/reference/test/include/input_graph.hpp:8-56: 44 directed slots; :58-68: the degrees and CSR offsets it
# expects; test_delegate_graph_static.cpp:146-152: with delegate threshold 4 the hubs are 6, 7 and 8)
def grid_graph_slots():
    slots = []
    for r in range(3):
        for c in range(5):
            v = 5 * r + c
            for dr, dc in ((-1, 0), (0, -1), (0, 1), (1, 0)):
                rr, cc = r + dr, c + dc
                if 0 <= rr < 3 and 0 <= cc < 5:
                    slots.append((v, 5 * rr + cc))
    return sorted(slots)


GRID_DEGREE = [2, 3, 3, 3, 2, 3, 4, 4, 4, 3, 2, 3, 3, 3, 2]
GRID_OFFSET = [0, 2, 5, 8, 11, 13, 16, 20, 24, 28, 31, 33, 36, 39, 42, 44]
GRID_HUBS_AT_THRESHOLD_4 = [6, 7, 8]


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


# ---------------------------------------------------------------------------------------------------------------------
# Inputs of the committed reference outputs (tests/golden/reference_runs/*.json, written by oracle/make_reference_golden.py
# from the reference's own binary).  A case is a small dict that regenerates its input here; "gpu": the engine is run on it
# too (only inputs whose settings the other GPU tests already use: template-driven search from constraint 4, or none).
def _spec_by_name(name):
    for n, spec, labelset, _ in SPECS:
        if n == name:
            return spec, labelset
    for q in QUIRK_SPECS:
        if q[0] == name:
            return q[1], q[2]
    for n, spec, labelset, _ in APPROX_SPECS:
        if n == name:
            return spec, labelset
    if name in BENCH_TEMPLATES:
        return BENCH_TEMPLATES[name], sorted(set(BENCH_TEMPLATES[name]["labels"]))
    raise KeyError(name)


# the cyclic templates of bench.py's default workload (BASELINE configs[2]), with its label choices
BENCH_TEMPLATES = {"triangle_678": PT.triangle(6, 7, 8), "cycle4_5678": PT.cycle4(5, 6, 7, 8),
                   "cycle6_chords_456789": PT.cycle6_chords([4, 5, 6, 7, 8, 9])}


HUB_TEMPLATES = {"triangle_hubs_12_14_16": PT.triangle(12, 14, 16)}


def rmat20_template(name):
    return BENCH_TEMPLATES[name] if name in BENCH_TEMPLATES else HUB_TEMPLATES[name]


def bench_template_at_constraint_4(name):
    """(spec as the reference driver ran it, index of its enumeration constraint there)"""
    from oracle import reference_run as R
    return R.tds_at_constraint_4(rmat20_template(name))


def subgraphs_digest(rows):
    """count and SHA-256 of the sorted enumerated walks (one "a,b,c,..." line per walk): what a fixture keeps when the
    walks themselves are too many to commit"""
    import hashlib
    rows = sorted(tuple(int(x) for x in r) for r in rows)
    return {"count": len(rows), "sha256": hashlib.sha256("\n".join(",".join(map(str, r)) for r in rows).encode()).hexdigest()}


def reference_golden_cases():
    out = []
    for name, _, _, _ in SPECS:
        for seed in range(4):
            out.append({"name": "%s_random_%d" % (name, seed), "kind": "random", "spec": name, "seed": seed,
                        "labels": "random", "gpu": name == "tree"})
        out.append({"name": "%s_planted_1" % name, "kind": "planted", "spec": name, "seed": 1, "labels": "random",
                    "gpu": name == "tree"})
    for seed, n, m in list(quirk_inputs("bowtie_no_interleave", 1))[:4]:
        out.append({"name": "bowtie_no_interleave_%d" % seed, "kind": "quirk", "spec": "bowtie_no_interleave", "seed": seed,
                    "n": n, "m": m, "labels": "random", "gpu": True})
    for seed, n, m in list(quirk_inputs("twin", 3))[:4]:
        out.append({"name": "twin_at_constraint_4_%d" % seed, "kind": "quirk", "spec": "twin", "seed": seed, "n": n, "m": m,
                    "labels": "random", "pad_to_constraint_4": True, "gpu": False})
    out.append({"name": "rmat15_tree_degree_labels", "kind": "rmat", "spec": "tree", "scale": 15, "gen_ranks": 4,
                "labels": "degree_log2", "gpu": True})
    # the run_fuzzy path (the reference's src/run_pattern_matching.cpp): inputs of the GPU fuzzy parity test
    for name in ("triangle", "cycle4"):
        for seed in range(5):
            out.append({"name": "fuzzy_%s_random_%d" % (name, seed), "kind": "random", "spec": name, "seed": seed,
                        "labels": "random", "path": "run_fuzzy", "gpu": True})
    # ... and its R-MAT input: scale 17 with the reference's own degree labels, the 4-cycle of bench.py's workload (the engine is
    # held to this fixture inside tests/test_gpu_parity.py::test_run_fuzzy_path_matches_oracle)
    out.append({"name": "fuzzy_rmat17_cycle4_5678", "kind": "rmat", "spec": "cycle4_5678", "scale": 17, "gen_ranks": 4,
                "labels": "degree_log2", "path": "run_fuzzy", "gpu": False})
    # BASELINE configs[0] exactly: R-MAT scale 21 with 4 generating ranks, the reference's own degree labels and its
    # examples/rmat_log2_tree_pattern (tests/golden/rmat_log2_tree_pattern).  The graph comes from the oracle's generator
    # (Graph.rmat), not from an edge list in Python; written by `make_reference_golden.py --large` (a 1 GB slot file, minutes).
    # The engine is held to this fixture inside tests/test_gpu_parity.py::test_baseline_config1_scale21_reference_pattern_dir.
    out.append({"name": "rmat21_config0_reference_pattern_dir", "kind": "rmat_generated", "scale": 21, "gen_ranks": 4,
                "labels": "degree_log2", "pattern_dir": "golden/rmat_log2_tree_pattern", "gpu": False, "large": True})
    # BASELINE configs[2]'s templates as bench.py writes them (triangle, 4-cycle, 6-cycle with chords over populous degree
    # classes) on R-MAT scale 20 with 4 generating ranks and the reference's own degree labels — the input of
    # tests/test_gpu_parity.py::test_bench_templates_on_rmat_scale20.  The driver starts template-driven search at constraint
    # 4 (beta.cpp:725-730), so the reference ran the template with its enumeration walk moved there
    # (oracle/reference_run.py::tds_at_constraint_4: copies of the first, already satisfied constraint in front of it).
    for name in sorted(BENCH_TEMPLATES):
        out.append({"name": "rmat20_bench_%s" % name, "kind": "rmat_generated", "scale": 20, "gen_ranks": 4,
                    "labels": "degree_log2", "bench_template": name, "gpu": False, "large": True})
    # the hub-class template of tests/test_gpu_parity.py::test_hub_class_template_on_rmat_scale20 (degree labels 12, 14, 16:
    # rows of 2^11 .. 2^16 slots) on the same graph; its 135 126 enumerated walks are stored as a count and a digest
    out.append({"name": "rmat20_hubs_triangle_12_14_16", "kind": "rmat_generated", "scale": 20, "gen_ranks": 4,
                "labels": "degree_log2", "bench_template": "triangle_hubs_12_14_16", "digest_subgraphs": True,
                "gpu": False, "large": True})
    # approximate matching (the reference's src/run_pattern_matching_beta_2.cpp): the rows of the first local
    # constraint checking call, on inputs of the GPU approximate-pattern test
    for name, _, _, _ in APPROX_SPECS:
        for seed in range(3):
            out.append({"name": "approx_%s_%d" % (name, seed), "kind": "approx", "spec": name, "seed": seed,
                        "labels": "random", "path": "approx_first_lcc", "gpu": True})
    return out


def reference_golden_input(case, oracle):
    """(n, undirected edges, labels or None, spec) of a case; `oracle` supplies the R-MAT stream (rmat kind only)"""
    spec, labelset = _spec_by_name(case["spec"])
    if case.get("pad_to_constraint_4"):
        spec = dict(spec, constraints=[{"walk": [0, 1]} for _ in range(4)] + list(spec["constraints"]))
    if case["kind"] == "rmat":
        scale, gen = case["scale"], case["gen_ranks"]
        e = np.concatenate([oracle.rmat_stream(scale, r, (16 << scale) // gen) for r in range(gen)])
        return 1 << scale, [tuple(x) for x in e.tolist()], None, spec
    seed = case["seed"]
    if case["kind"] == "planted":
        edges, labels = planted(seed, 300, 900, spec, labelset)
        return 300, edges, labels, spec
    if case["kind"] == "approx":
        n, m = 60 + 10 * (seed % 4), 160 + 40 * (seed % 5)
        return n, random_multigraph(seed + 500, n, m), random_labels(seed + 500, n, labelset), spec
    if case["kind"] == "quirk":
        n, m = case["n"], case["m"]
    else:
        n, m = 60 + 10 * (seed % 4), 220 + 60 * (seed % 5)
    return n, random_multigraph(seed, n, m), random_labels(seed, n, labelset), spec


def reference_golden_load(path):
    """the stored reference output in the comparable form of run_summary / engine_summary"""
    import json
    with open(path) as f:
        doc = json.load(f)
    r = doc["reference"]
    if doc["case"].get("path") == "approx_first_lcc":  # the count rows of the first local constraint checking call
        return doc["case"], dict(rows=[(a, b, c, d, e) for a, b, c, d, e in r["rows"]])
    if doc["case"].get("path") == "run_fuzzy":  # count rows without edge counts, vertex -> template vertex index
        return doc["case"], dict(rows=[(a, b, c, d, e) for a, b, c, d, e in r["rows"]], iterations=r["iterations"],
                                 vertices=[tuple(x) for x in r["vertices"]])
    out = dict(rows=[(a, b, c, d, e) for a, b, c, d, e in r["rows"]], iterations=r["iterations"],
               vertices=[tuple(x) for x in r["vertices"]], edges=[tuple(x) for x in r["edges"]],
               subgraphs={int(k): [tuple(x) for x in v] for k, v in r.get("subgraphs", {}).items()})
    if "subgraphs_digest" in r:  # {constraint: {"count", "sha256"}} instead of the walks (subgraphs_digest above)
        out["subgraphs_digest"] = {int(k): v for k, v in r["subgraphs_digest"].items()}
    return doc["case"], out


def assert_equals_reference_golden(summary, golden):
    """summary: run_summary(oracle run) or engine_summary(engine); only template-driven search writes subgraph files"""
    assert summary["rows"] == golden["rows"], "rows"
    assert summary["iterations"] == golden["iterations"], "iterations"
    assert sorted(summary["vertices"]) == golden["vertices"], "vertices"
    assert sorted(summary["edges"]) == golden["edges"], "edges"
    for pl in range(4, len(summary["subgraphs"])):
        if pl in golden.get("subgraphs_digest", {}):
            assert subgraphs_digest(summary["subgraphs"][pl]) == golden["subgraphs_digest"][pl], "walks of constraint %d" % pl
        else:
            assert sorted(summary["subgraphs"][pl]) == golden["subgraphs"].get(pl, []), "subgraphs of constraint %d" % pl


def assert_final_sets_equal_reference_golden(summary, tds_pl, golden):
    """A template run with its enumeration walk at its OWN constraint index `tds_pl`, against a fixture of the reference
    driver, which ran the same template with that walk moved to constraint 4 behind copies of an already satisfied
    constraint (oracle/reference_run.py::tds_at_constraint_4): the count rows differ by those extra constraints' rows, the
    final vertex -> template bitset map, the final edge set and the enumerated walks do not."""
    assert sorted(summary["vertices"]) == golden["vertices"], "vertices"
    assert sorted(summary["edges"]) == golden["edges"], "edges"
    if 4 in golden.get("subgraphs_digest", {}):
        if summary["subgraphs"][tds_pl] is not None:  # None: the caller kept no walks and compares the count itself
            assert subgraphs_digest(summary["subgraphs"][tds_pl]) == golden["subgraphs_digest"][4], "enumerated walks"
    else:
        assert sorted(summary["subgraphs"][tds_pl]) == golden["subgraphs"].get(4, []), "enumerated walks"
    first_call = lambda rows: [r for r in rows if r[0] == 0 and r[1] == "LP"]  # noqa: E731
    assert first_call(summary["rows"]) == first_call(golden["rows"]), "rows of the first local constraint checking call"


def check_multi_rank_attribution(oracle, n, edges, labels, spec, tds_from, ranks, threshold, workdir):
    """The oracle's per-rank files of a `ranks`-rank run with delegate threshold `threshold` (0: no hubs) against the reference's
    rules restated from the raw edge list (see tests/test_oracle_delegates.py).  Returns a few counts for the caller's
    non-triviality checks; skips nothing — raises AssertionError on any deviation."""
    import collections
    import os

    def rows_of(path):
        return [[t.strip() for t in l.split(",")] for l in open(path).read().splitlines() if l.strip()]

    out_degree = collections.Counter()
    for a, b in edges:  # an undirected input edge is two directed slots; a self loop is two slots of the same vertex
        out_degree[a] += 1
        out_degree[b] += 1
    hubs = sorted(v for v in range(n) if threshold and out_degree[v] >= threshold)
    controller = {v: i % ranks for i, v in enumerate(hubs)}
    owner = lambda v: controller[v] if v in controller else v % ranks  # noqa: E731
    d = pattern_dir(spec)
    g = oracle.Graph.from_undirected(n, edges)
    assert np.array_equal(g.degree, np.array([out_degree[v] for v in range(n)], dtype=np.uint64))
    one = oracle.Run(g, labels, oracle.Pattern(d), n_ranks=1, tds_from_pl=tds_from, max_iterations=50)
    many = oracle.Run(g, labels, oracle.Pattern(d), n_ranks=ranks, tds_from_pl=tds_from, max_iterations=50,
                      delegate_threshold=threshold)
    assert many.rows == one.rows and many.iterations == one.iterations  # aggregate rows = the single-rank rows
    os.makedirs(workdir, exist_ok=True)
    oracle.make_result_tree(workdir)
    many.write_results(workdir)
    v1, t1 = one.active_vertices()
    final_vertices = dict(zip(v1.tolist(), t1.tolist()))
    final_edges = set(map(tuple, one.active_edges.tolist()))
    seen_v, seen_e = {}, set()
    base = os.path.join(workdir, "0")
    for r in range(ranks):
        rows = rows_of(os.path.join(base, "all_ranks_active_vertices", "active_vertices_%d" % r))
        for t in rows:  # "rank, vertex, pattern index, label, bitset"
            v = int(t[1])
            assert int(t[0]) == r == owner(v), (r, v)
            assert v not in seen_v
            seen_v[v] = int(t[4], 2)
            assert int(t[3]) == int(labels[v])
        erows = rows_of(os.path.join(base, "all_ranks_active_edges", "active_edges_%d" % r))
        for t in erows:  # "rank, vertex, neighbour"
            assert int(t[0]) == r == owner(int(t[1]))
            seen_e.add((int(t[1]), int(t[2])))
        vc = rows_of(os.path.join(base, "all_ranks_active_vertices_count", "active_vertices_%d" % r))
        ec = rows_of(os.path.join(base, "all_ranks_active_edges_count", "active_edges_%d" % r))
        assert int(vc[-1][3]) == len(rows) and int(ec[-1][3]) == len(erows)  # the last row counts what this rank wrote
    assert seen_v == final_vertices and seen_e == final_edges  # the union over the ranks is the single-rank result
    # enumerated walks: "[rank], v0, ..., v_last, [v_last]" written where the walk completes (tds_batch_1.hpp:684-689) — the
    # rank that runs the final vertex's visit: its controller for a hub, its modulo owner otherwise
    for pl, want_rows in enumerate(one.subgraphs):
        got_rows = []
        for r in range(ranks):
            path = os.path.join(base, "all_ranks_subgraphs", "subgraphs_%d_%d" % (pl, r))
            for l in (open(path).read().splitlines() if os.path.exists(path) else []):
                toks = [t.strip() for t in l.split(",")]
                assert toks[0] == "[%d]" % r and toks[-1] == "[%s]" % toks[-2]
                assert owner(int(toks[-2])) == r, (pl, r, l)
                got_rows.append(tuple(int(t) for t in toks[1:-1]))
        assert sorted(got_rows) == sorted(map(tuple, want_rows.tolist())), pl
    per_v = [rows_of(os.path.join(base, "all_ranks_active_vertices_count", "active_vertices_%d" % r)) for r in range(ranks)]
    per_e = [rows_of(os.path.join(base, "all_ranks_active_edges_count", "active_edges_%d" % r)) for r in range(ranks)]
    assert all(len(p) == len(one.rows) for p in per_v + per_e)
    for i, row in enumerate(one.rows):  # row by row, the ranks' counts add up to the single-rank rows
        assert sum(int(p[i][3]) for p in per_v) == row[3] and sum(int(p[i][3]) for p in per_e) == row[4], i
        assert all((int(p[i][0]), p[i][1], int(p[i][2])) == tuple(row[:3]) for p in per_v)
    return {"hubs": len(hubs), "final_vertices": len(final_vertices),
            "hubs_moved": sum(1 for v in final_vertices if v in controller and controller[v] != v % ranks)}
