"""Outputs of THE REFERENCE ITSELF as committed fixtures: tests/golden/reference_runs/*.json were written by
oracle/make_reference_golden.py from oracle/_ref/run_pattern_matching_beta — the reference's own driver and visitor
headers over the single-rank runtime stand-in of oracle/ref_shim.  The inputs are regenerated from seeds (tests/cases.py),
so the fixtures only hold what the reference printed into its result files.

CPU: the oracle equals every fixture (template-driven search from constraint 4, like the driver, beta.cpp:725-730); inputs
the oracle flags as order dependent in the reference are skipped (a fixture then records one of several valid outcomes).
GPU: the engine equals the fixtures whose settings the other GPU tests use."""
import glob
import os

import pytest

from fuzzypatternmatching_b200 import patterns as PT
from tests import cases

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_runs", "*.json")))


def _ids(paths):
    return [os.path.basename(p)[:-5] for p in paths]


def test_fixture_set_is_complete():
    assert _ids(GOLDEN) == sorted(c["name"] for c in cases.reference_golden_cases())


@pytest.mark.parametrize("path", GOLDEN, ids=_ids(GOLDEN))
def test_oracle_equals_reference_output(oracle, path):
    case, golden = cases.reference_golden_load(path)
    if case["kind"] == "rmat_generated" and "bench_template" in case:
        # BASELINE configs[2]'s templates on R-MAT scale 20: the reference ran the template with its enumeration walk at
        # constraint 4; the oracle equals it on that template — and on the template as bench.py and the GPU test run it
        # (enumeration at its own index) the final sets and the enumerated walks are the same
        g, labels = _rmat_graph(oracle, case["scale"], case["gen_ranks"])
        spec4, at4 = cases.bench_template_at_constraint_4(case["bench_template"])
        run = oracle.Run(g, labels, oracle.Pattern(cases.pattern_dir(spec4)), tds_from_pl=4)
        assert not run.hazards[:5].any()
        cases.assert_equals_reference_golden(cases.run_summary(run), golden)
        spec = cases.rmat20_template(case["bench_template"])
        own = oracle.Run(g, labels, oracle.Pattern(cases.pattern_dir(spec)), tds_from_pl=PT.tds_from_pl(spec))
        cases.assert_final_sets_equal_reference_golden(cases.run_summary(own), PT.tds_from_pl(spec), golden)
        return
    if case["kind"] == "rmat_generated":  # BASELINE configs[0]: the oracle's own generator and the reference's pattern directory
        g = oracle.Graph.rmat(case["scale"], case["gen_ranks"])
        d = os.path.join(os.path.dirname(os.path.abspath(__file__)), case["pattern_dir"], "0")
        run = oracle.Run(g, g.labels_degree_log2(), oracle.Pattern(d), tds_from_pl=4)
        assert not run.hazards[:5].any()
        cases.assert_equals_reference_golden(cases.run_summary(run), golden)
        assert len(golden["vertices"]) == 147 and len(golden["edges"]) == 262 and len(golden["subgraphs"][4]) == 74
        return
    n, edges, labels, spec = cases.reference_golden_input(case, oracle)
    g = oracle.Graph.from_undirected(n, edges)
    if labels is None:
        labels = g.labels_degree_log2()
    if case.get("path") == "approx_first_lcc":
        run = oracle.Run(g, labels, oracle.Pattern(cases.pattern_dir(spec)), tds_from_pl=-1, max_iterations=50)
        assert run.rows[:spec["diameter"]] == golden["rows"] and len(golden["rows"]) == spec["diameter"]
        return
    if case.get("path") == "run_fuzzy":
        run = oracle.Run(g, labels, oracle.Pattern(cases.pattern_dir(spec)), fuzzy=True, max_iterations=50)
        _assert_fuzzy(run.rows, run.iterations, *run.active_vertices(), golden)
        return
    run = oracle.Run(g, labels, oracle.Pattern(cases.pattern_dir(spec)), tds_from_pl=4, max_iterations=50)
    if run.hazards[:3].any() or run.hazards[4]:
        pytest.skip("the reference is order dependent on this input")
    cases.assert_equals_reference_golden(cases.run_summary(run), golden)


_RMAT = {}


def _rmat_graph(oracle, scale, gen_ranks):
    if (scale, gen_ranks) not in _RMAT:
        g = oracle.Graph.rmat(scale, gen_ranks)
        _RMAT.clear()
        _RMAT[(scale, gen_ranks)] = (g, g.labels_degree_log2())
    return _RMAT[(scale, gen_ranks)]


def _assert_fuzzy(rows, iterations, v, t, golden):
    """run_fuzzy path: the reference driver writes vertex counts only and one template vertex index per vertex"""
    assert [(a, b, c, nv, 0) for a, b, c, nv, _ in rows] == golden["rows"], "rows"
    assert int(iterations) == golden["iterations"], "iterations"
    assert sorted((int(a), int(b).bit_length() - 1) for a, b in zip(v, t)) == golden["vertices"], "vertices"


def test_fixtures_are_not_trivial(oracle):
    ends, enumerated, multi = 0, 0, 0
    for path in GOLDEN:
        _, golden = cases.reference_golden_load(path)
        if "vertices" not in golden:
            continue
        ends += len(golden["vertices"]) > 0
        enumerated += any(len(v) > 0 for v in golden.get("subgraphs", {}).values())
        multi += golden["iterations"] > 1
    assert ends >= 15 and enumerated >= 4 and multi >= 5


GPU_GOLDEN = [p for p in GOLDEN if cases.reference_golden_load(p)[0].get("gpu")]


@pytest.fixture(scope="module")
def eng():
    from fuzzypatternmatching_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", GPU_GOLDEN, ids=_ids(GPU_GOLDEN))
def test_engine_equals_reference_output(oracle, eng, path):
    """libpmgpu.so through the C ABI against what the reference itself wrote for the same input (the call sequence of
    tests/test_gpu_parity.py::_compare, with the fixture in the oracle's place)"""
    from fuzzypatternmatching_b200 import patterns as PT
    case, golden = cases.reference_golden_load(path)
    n, edges, labels, spec = cases.reference_golden_input(case, oracle)
    d = cases.pattern_dir(spec)
    src, dst = cases.slots_of(edges)
    eng.graph_from_slots(n, src, dst)
    if labels is None:
        eng.labels_degree_log2()
    else:
        eng.labels_set(labels)
    eng.pattern_load_dir(d)
    if case.get("path") == "approx_first_lcc":  # the calls of test_gpu_parity.py::_compare on an approximate template
        eng.run(tds_from_pl=-1, max_iterations=50)
        assert eng.rows()[:spec["diameter"]] == golden["rows"]
        return
    if case.get("path") == "run_fuzzy":  # the calls of test_gpu_parity.py::test_run_fuzzy_path_matches_oracle
        eng.run_fuzzy(max_iterations=50)
        _assert_fuzzy(eng.rows(), eng.summary["iterations"], *eng.active_vertices(), golden)
        return
    tds_from = PT.tds_from_pl(spec)  # 4 for the tree template, -1 where no constraint enumerates: the driver's own choice
    assert tds_from in (4, -1)
    eng.run(tds_from_pl=tds_from, max_iterations=50)
    cases.assert_equals_reference_golden(cases.engine_summary(eng, len(spec["constraints"])), golden)
