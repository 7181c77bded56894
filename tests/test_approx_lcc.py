"""Approximate matching: the oracle's approximate local constraint — optional template edges,
vertex_min_optional_edge_count, mandatory / optional coverage (approximate_pattern_matching/local_constraint_checking.hpp:
641-651, 1062-1113; pattern files of pattern_graph.hpp:282-337, 604-622) — against an independent plain-Python
restatement of the superstep loop, and hand-derived cases."""
import numpy as np
import pytest

from tests import cases


def _python_lcc(n, edges, labels, spec):
    """Jacobi supersteps of SURVEY A.2 over dicts and sets, repeated by the outer loop until nothing is removed
    (LCC only).  Returns (rows, final {vertex: mask}, final edge set)."""
    tl = spec["labels"]
    nt = len(tl)
    opt = set()
    for a, b in spec.get("optional_edges", []):
        opt |= {(a, b), (b, a)}
    approx = bool(opt) or "min_optional" in spec
    Nm, No = [0] * nt, [0] * nt
    for a, b in spec["edges"]:
        for x, y in ((a, b), (b, a)):
            if (x, y) in opt:
                No[x] |= 1 << y
            else:
                Nm[x] |= 1 << y
    min_opt = [max(spec.get("min_optional", {}).get(i, 0), 0) for i in range(nt)]
    nbrs = {v: set() for v in range(n)}
    for a, b in edges:
        nbrs[a].add(b)
        nbrs[b].add(a)

    def nb(T):
        return sum_or(Nm[p] | No[p] for p in range(nt) if (T >> p) & 1)

    def sum_or(it):
        r = 0
        for x in it:
            r |= x
        return r

    def cover(T, heard):
        out = 0
        for p in range(nt):
            if not (T >> p) & 1:
                continue
            if not approx:
                ok = Nm[p] != 0 and (Nm[p] & ~heard) == 0
            else:
                ok = (Nm[p] & ~heard) == 0
                if min_opt[p] > 0:
                    ok = ok and (No[p] & ~heard) == 0 and bin(No[p]).count("1") >= min_opt[p]
            if ok:
                out |= 1 << p
        return out

    T = {}      # T_arr == T_state on this path (no NLCC in between)
    E = {}
    rows = []
    itr, init = 0, True
    while True:
        removed = False
        for k in range(spec["diameter"]):
            first = init and k == 0
            if first:
                lm = {v: sum_or(1 << p for p in range(nt) if tl[p] == int(labels[v])) for v in range(n)}
                send = {v: m for v, m in lm.items() if m}
                newT, newE = {}, {}
                for v, Tv in send.items():
                    heard, kept = 0, set()
                    for u in nbrs[v]:
                        m = send.get(u, 0)
                        if m and (m & nb(Tv)):
                            heard |= m
                            kept.add(u)
                    if not kept:
                        continue  # never entered the map
                    ts = cover(Tv, heard)
                    if ts:
                        newT[v], newE[v] = ts, kept
                    else:
                        removed = True
                T, E = newT, newE
            else:
                newT, newE = {}, {}
                for v, Tv in T.items():
                    heard, kept = 0, set()
                    for u in E[v]:
                        m = T.get(u, 0)
                        if m and (m & nb(Tv)):
                            heard |= m
                            kept.add(u)
                    ts = cover(Tv, heard)
                    if ts:
                        newT[v], newE[v] = ts, kept
                    else:
                        removed = True
                T, E = newT, newE
            rows.append((itr, "LP", k, len(T), sum(len(x) for x in E.values())))
        init = False
        itr += 1
        if not removed:
            break
    return rows, T, {(v, u) for v, s in E.items() for u in s}


@pytest.mark.parametrize("name,spec,labelset,tds_from", cases.APPROX_SPECS, ids=[s[0] for s in cases.APPROX_SPECS])
def test_oracle_approximate_lcc_equals_python_restatement(oracle, name, spec, labelset, tds_from):
    nontrivial = 0
    for seed in range(8):
        n, m = 60 + 10 * (seed % 4), 160 + 40 * (seed % 5)
        edges = cases.random_multigraph(seed + 500, n, m)
        labels = cases.random_labels(seed + 500, n, labelset)
        g = oracle.Graph.from_undirected(n, edges)
        r = oracle.Run(g, labels, oracle.Pattern(cases.pattern_dir(spec)), tds_from_pl=tds_from, lcc_only=True, max_iterations=50)
        rows, T, E = _python_lcc(n, edges, labels, spec)
        assert r.rows == rows
        v, t = r.active_vertices()
        assert dict(zip(v.tolist(), t.tolist())) == T
        assert set(map(tuple, r.active_edges.tolist())) == E
        nontrivial += len(T) > 0
    assert nontrivial >= (0 if name == "impossible_min_optional" else 3)


def test_exact_patterns_unchanged_by_the_python_restatement(oracle):
    """the same restatement on an exact template: an independent check of the oracle's LCC loop"""
    name, spec, labelset, tds_from = cases.SPECS[0]
    for seed in range(6):
        n, m = 70, 300
        edges = cases.random_multigraph(seed + 40, n, m)
        labels = cases.random_labels(seed + 40, n, labelset)
        g = oracle.Graph.from_undirected(n, edges)
        r = oracle.Run(g, labels, oracle.Pattern(cases.pattern_dir(spec)), lcc_only=True, max_iterations=50)
        rows, T, E = _python_lcc(n, edges, labels, spec)
        assert r.rows == rows and set(map(tuple, r.active_edges.tolist())) == E


def test_optional_edge_by_hand(oracle):
    """Square 1-2-3-4 with an optional diagonal 1-3.  Graph A has the diagonal, graph B does not: both survive.  With the
    diagonal mandatory, B dies.  A minimum optional count of 1 on template vertex 0 makes the diagonal required at the
    label-1 vertex only (its far end keeps no such requirement)."""
    sq = [(0, 1), (1, 2), (2, 3), (0, 3)]
    labels = np.array([1, 2, 3, 4], dtype=np.uint64)
    base = {"labels": [1, 2, 3, 4], "edges": sq + [(0, 2)], "diameter": 3, "constraints": []}
    approx = dict(base, optional_edges=[(0, 2)])
    need0 = dict(approx, min_optional={0: 1})

    def run(spec, edges):
        g = oracle.Graph.from_undirected(4, edges)
        r = oracle.Run(g, labels, oracle.Pattern(cases.pattern_dir(spec)), lcc_only=True, tds_from_pl=-1)
        return r.in_map.sum(), len(r.active_edges)

    assert run(base, sq + [(0, 2)]) == (4, 10)
    assert run(base, sq) == (0, 0)
    assert run(approx, sq + [(0, 2)]) == (4, 10)
    assert run(approx, sq) == (4, 8)
    assert run(need0, sq + [(0, 2)]) == (4, 10)
    assert run(need0, sq) == (0, 0)   # vertex 0 leaves, then the rest unravels
