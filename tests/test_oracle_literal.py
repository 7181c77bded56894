"""The C++ oracle against oracle/ref_literal.py: a dict/set transliteration of the
reference visitors executed with a RANDOMISED asynchronous scheduler.  Equality over
seeds shows the oracle's level-synchronous evaluation is faithful to any message
order; a template with repeated interior labels shows the hazard counters fire
exactly when the reference itself becomes order dependent."""
import numpy as np
import pytest

from oracle.ref_literal import LiteralRun, PatternFiles
from tests import cases


def _both(oracle, spec, labelset, tds_from, seed, n, m, sched_seeds=(0, 1)):
    d = cases.pattern_dir(spec)
    edges = cases.random_multigraph(seed, n, m)
    labels = cases.random_labels(seed, n, labelset)
    g = oracle.Graph.from_undirected(n, edges)
    r = oracle.Run(g, labels, oracle.Pattern(d), tds_from_pl=tds_from, max_iterations=50)
    slots = []
    for a, b in edges:
        slots += [(a, b), (b, a)]
    lits = [LiteralRun(n, slots, labels.tolist(), PatternFiles(d), seed=seed * 31 + s, tds_from_pl=tds_from)
            for s in sched_seeds]
    return r, lits


@pytest.mark.parametrize("name,spec,labelset,tds_from", cases.SPECS, ids=[s[0] for s in cases.SPECS])
def test_oracle_equals_literal_async_execution(oracle, name, spec, labelset, tds_from):
    nontrivial = multi_iter = 0
    for seed in range(10):
        r, lits = _both(oracle, spec, labelset, tds_from, seed, 50 + 5 * (seed % 3), 200 + 40 * (seed % 4))
        want = cases.run_summary(r)
        assert not r.hazards[:5].any()
        for L in lits:
            assert not L.errors
            assert L.rows == want["rows"]
            assert L.iterations == want["iterations"]
            assert L.final_vertices() == want["vertices"]
            assert L.final_edges() == want["edges"]
            assert [sorted(s) for s in L.subgraphs] == want["subgraphs"]
        nontrivial += r.rows[-1][3] > 0
        multi_iter += r.iterations > 1
    assert nontrivial >= 2


def test_hazard_counters_fire_on_order_dependent_template(oracle):
    # 4-cycle with alternating labels a b a b: interior hops 1 and 3 share a label, so the
    # (vertex, source) aggregation of nem_1 makes the reference order dependent
    from fuzzypatternmatching_b200 import patterns as PT
    spec = PT.cycle4(1, 2, 1, 2)
    spec["constraints"] = spec["constraints"][:1]
    fired = 0
    for seed in range(8):
        r, _ = _both(oracle, spec, [1, 2], -1, seed, 40, 200, sched_seeds=())
        fired += int(r.hazards[0] + r.hazards[1] > 0)
    assert fired > 0


def test_edge_cases_equal_literal_execution(oracle):
    """The degenerate inputs of the GPU parity suite, oracle against the literal asynchronous execution."""
    for name, n, edges, labels, spec, tds_from in cases.edge_cases():
        if n > 5000:
            continue  # the literal executor is pure Python
        d = cases.pattern_dir(spec)
        g = oracle.Graph.from_undirected(n, edges)
        r = oracle.Run(g, labels, oracle.Pattern(d), tds_from_pl=tds_from, max_iterations=50)
        want = cases.run_summary(r)
        assert not r.hazards[:5].any(), name
        slots = []
        for a, b in edges:
            slots += [(a, b), (b, a)]
        for s in (0, 1):
            L = LiteralRun(n, slots, labels.tolist(), PatternFiles(d), seed=17 + s, tds_from_pl=tds_from)
            assert not L.errors, name
            assert L.rows == want["rows"], name
            assert L.iterations == want["iterations"], name
            assert L.final_vertices() == want["vertices"], name
            assert L.final_edges() == want["edges"], name
            assert [sorted(x) for x in L.subgraphs] == want["subgraphs"], name
