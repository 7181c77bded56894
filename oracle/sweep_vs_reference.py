"""A long randomized sweep of the oracle against THE REFERENCE ITSELF (oracle/_ref/run_pattern_matching_beta, see
oracle/ref_shim/README.md): many more seeds, sizes and densities than tests/test_oracle_vs_reference.py runs in the CPU
suite.  Every input the oracle does not flag as order dependent in the reference must give the same count rows, iteration
count, final vertex -> bitset map, final edge set and enumerated walks.  TEST INFRASTRUCTURE.
    python oracle/sweep_vs_reference.py [--path beta|fuzzy|approx] [--seeds N] [--jobs J]
(prints one summary line per template, exit 1 on a mismatch)
"""
import argparse
import multiprocessing as mp
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != os.path.join(ROOT, "oracle")]
sys.path.insert(0, ROOT)


def templates():
    from tests import cases
    from oracle import reference_run as R
    out = []
    for name, spec, labelset, _ in cases.SPECS:
        out.append((name, spec, labelset))
    for name, spec, labelset, tds_from, _, _ in cases.QUIRK_SPECS:
        if tds_from >= 0:  # the driver enumerates from constraint 4 on: one-hop path checks in front
            spec = dict(spec, constraints=[{"walk": [0, 1]} for _ in range(4)] + list(spec["constraints"]))
        out.append((name, spec, labelset))
    for name, spec in cases.BENCH_TEMPLATES.items():
        out.append(("bench_" + name, R.tds_at_constraint_4(spec)[0], sorted(set(spec["labels"]))))
    return out


def fuzzy_templates():
    from fuzzypatternmatching_b200 import patterns as PT
    return [("fuzzy_triangle", PT.triangle(1, 2, 3), [1, 2, 3]), ("fuzzy_cycle4", PT.cycle4(1, 2, 3, 4), [1, 2, 3, 4])]


def approx_templates():
    from tests import cases
    return [("approx_" + name, spec, labelset) for name, spec, labelset, _ in cases.APPROX_SPECS]


def _random_input(cases, rng, seed, labelset):
    n = rng.choice([12, 40, 90, 200, 500])
    m = int(n * rng.choice([1.0, 2.0, 3.5, 6.0]))
    edges = cases.random_multigraph(seed, n, m, dup=rng.choice([0.0, 0.1, 0.4]), loops=rng.choice([0.0, 0.05, 0.3]))
    labels = cases.random_labels(seed, n, labelset + ([labelset[0] + 50] if rng.random() < 0.3 else []))
    return n, edges, labels


def one_fuzzy(args):
    """the run_fuzzy path: the reference's src/run_pattern_matching.cpp"""
    t_index, seed = args
    from oracle import oracle as O
    from oracle import reference_run as R
    from tests import cases
    name, spec, labelset = fuzzy_templates()[t_index]
    n, edges, labels = _random_input(cases, random.Random(seed * 104729 + t_index), seed, labelset)
    if not len(edges):
        return name, "skipped", None
    d = cases.pattern_dir(spec)
    run = O.Run(O.Graph.from_undirected(n, edges), labels, O.Pattern(d), fuzzy=True, max_iterations=60)
    src, dst = cases.slots_of(edges)
    got = R.run_fuzzy(n, src.tolist(), dst.tolist(), os.path.dirname(d), labels.tolist())
    v, t = run.active_vertices()
    ok = (got["rows"] == [(a, b, c, nv, 0) for a, b, c, nv, _ in run.rows] and got["iterations"] == run.iterations
          and got["vertices"] == sorted((int(a), int(b).bit_length() - 1) for a, b in zip(v, t)))
    if not ok:
        return name, "MISMATCH", dict(seed=seed, n=n)
    return name, ("nontrivial" if run.rows[-1][3] > 0 else "empty") + ("+walked" if any(r[1] == "TP" for r in run.rows) else ""), None


def one_approx(args):
    """approximate matching: the rows of the first local constraint checking call of run_pattern_matching_beta_2.cpp"""
    t_index, seed = args
    from oracle import oracle as O
    from oracle import reference_run as R
    from tests import cases
    name, spec, labelset = approx_templates()[t_index]
    n, edges, labels = _random_input(cases, random.Random(seed * 1299709 + t_index), seed, labelset)
    if not len(edges):
        return name, "skipped", None
    d = cases.pattern_dir(spec)
    run = O.Run(O.Graph.from_undirected(n, edges), labels, O.Pattern(d), tds_from_pl=-1, max_iterations=60)
    want = [r for r in run.rows[:spec["diameter"]]]
    src, dst = cases.slots_of(edges)
    got = R.run_approx_first_lcc(n, src.tolist(), dst.tolist(), os.path.dirname(d), labels.tolist(), spec)
    if got != want:
        return name, "MISMATCH", dict(seed=seed, n=n)
    return name, "pruned" if want[-1][3] < n else "kept_all", None


def random_template(rng):
    """A random connected template of 3..6 vertices: a random tree plus up to two extra edges; labels distinct or with
    repeats; one cycle constraint per extra edge (tree path + the closing edge), one path constraint per pair of equally
    labelled vertices (both directions, like the README tree), then — at constraint index >= 4, where the driver starts
    template-driven search — a closed depth-first walk over every template edge as the enumeration constraint."""
    k = rng.randint(3, 6)
    parent = {i: rng.randrange(i) for i in range(1, k)}
    edges = [(parent[i], i) for i in range(1, k)]
    for _ in range(rng.randint(0, 2)):
        a, b = sorted(rng.sample(range(k), 2))
        if (a, b) not in edges:
            edges.append((a, b))
    labels = list(range(1, k + 1))
    if rng.random() < 0.5:
        for _ in range(rng.randint(1, 2)):
            labels[rng.randrange(k)] = labels[rng.randrange(k)]
    adj = {i: set() for i in range(k)}
    for a, b in edges:
        adj[a].add(b); adj[b].add(a)  # noqa: E702

    def tree_path(a, b):
        up = lambda x: [x] + (up(parent[x]) if x else [])  # noqa: E731
        pa, pb = up(a), up(b)
        common = next(x for x in pa if x in pb)
        return pa[:pa.index(common) + 1] + pb[:pb.index(common)][::-1]

    def bfs_ecc(src):
        dist, todo = {src: 0}, [src]
        for x in todo:
            for y in adj[x]:
                if y not in dist:
                    dist[y] = dist[x] + 1
                    todo.append(y)
        return max(dist.values())
    cons = []
    for a, b in edges[k - 1:]:
        walk = tree_path(a, b)
        if len(walk) >= 3:
            cons.append({"walk": walk + [a], "cycle": True})
    for a in range(k):
        for b in range(k):
            if a != b and labels[a] == labels[b] and len(tree_path(a, b)) >= 3:
                cons.append({"walk": tree_path(a, b)})
    cons = cons[:6]
    while len(cons) < 4:
        cons.append({"walk": [0, 1] if (0, 1) in edges else list(edges[0])})  # one-hop checks LCC has already settled
    walk, seen, last_new = [0], set(), [0]

    def dfs(x):
        for y in sorted(adj[x]):
            e = (min(x, y), max(x, y))
            if e in seen:
                continue
            seen.add(e)
            walk.append(y)
            last_new[0] = len(walk)
            dfs(y)
            walk.append(x)
    dfs(0)
    walk = walk[:last_new[0]]  # no backtracking behind the last new edge (the README walk ends at its last discovered vertex)
    tds = {"walk": walk, "tds": True}
    if walk[-1] == walk[0]:
        tds["cycle"] = True  # a walk that closes at its source is a cycle constraint (valid_cycle = 1)
    cons.append(tds)
    return {"labels": labels, "edges": edges, "diameter": max(1, max(bfs_ecc(i) for i in range(k))) + rng.randint(0, 1),
            "constraints": cons}


def one_random_template(args):
    _, seed = args
    from oracle import oracle as O
    from oracle import reference_run as R
    from tests import cases
    rng = random.Random(seed * 15485863 + 11)
    spec = random_template(rng)
    labelset = sorted(set(spec["labels"]))
    name = "random_templates"
    n = rng.choice([12, 40, 90, 200])
    m = int(n * rng.choice([1.0, 2.0, 3.5, 6.0]))
    if rng.random() < 0.4 and n >= 40:
        edges, labels = cases.planted(seed, n, m, spec, labelset, copies=rng.choice([1, 3]))
    else:
        edges = cases.random_multigraph(seed, n, m, dup=rng.choice([0.0, 0.1, 0.4]), loops=rng.choice([0.0, 0.05, 0.3]))
        labels = cases.random_labels(seed, n, labelset)
    if not len(edges):
        return name, "skipped", None
    d = cases.pattern_dir(spec)
    try:
        pat = O.Pattern(d)
    except Exception:  # noqa: BLE001 — a template the pattern reader refuses
        return name, "refused_by_the_reader", None
    run = O.Run(O.Graph.from_undirected(n, edges), labels, pat, tds_from_pl=4, max_iterations=60)
    if run.hazards[:3].any() or run.hazards[4]:
        return name, "order_dependent", None
    want = cases.run_summary(run)
    src, dst = cases.slots_of(edges)
    try:
        got = R.run(n, src.tolist(), dst.tolist(), os.path.dirname(d), labels=labels.tolist(), timeout=120)
    except Exception as e:  # noqa: BLE001
        return name, "reference_failed", dict(seed=seed, spec=spec, error=str(e)[-200:])
    if not R.template_read_intact(got["stdout"], spec):
        return name, "reference_misread_its_template", None
    ok = (got["rows"] == want["rows"] and got["iterations"] == want["iterations"] and got["vertices"] == sorted(want["vertices"])
          and got["edges"] == sorted(want["edges"])
          and all(got["subgraphs"].get(pl, []) == sorted(want["subgraphs"][pl]) for pl in range(4, len(want["subgraphs"]))))
    if not ok:
        return name, "MISMATCH", dict(seed=seed, n=n, m=m, spec=spec)
    return name, ("nontrivial" if want["rows"][-1][3] > 0 else "empty") + ("+multi_iteration" if want["iterations"] > 1 else "") + \
        ("+enumerated" if any(len(x) for x in want["subgraphs"]) else ""), None


_RMAT_STATE = {}


def _rmat_state(scale=17, gen=4):
    """the R-MAT graph every `rmat_templates` input searches: built once per worker; the slot file once per sweep"""
    if not _RMAT_STATE:
        import numpy as np
        from oracle import oracle as O
        from oracle import reference_run as R
        path = os.path.join(os.environ["PM_SWEEP_DIR"], "rmat%d_%d.slots" % (scale, gen))
        g = O.Graph.rmat(scale, gen)
        if not os.path.exists(path):
            e = np.concatenate([O.rmat_stream(scale, r, (16 << scale) // gen) for r in range(gen)])
            src = np.empty(2 * len(e), dtype=np.uint64)
            dst = np.empty(2 * len(e), dtype=np.uint64)
            src[0::2], dst[0::2] = e[:, 0], e[:, 1]
            src[1::2], dst[1::2] = e[:, 1], e[:, 0]
            R.write_slot_file(path + ".tmp%d" % os.getpid(), 1 << scale, src, dst)
            os.replace(path + ".tmp%d" % os.getpid(), path)
        _RMAT_STATE.update(g=g, labels=g.labels_degree_log2(), path=path)
    return _RMAT_STATE


def one_rmat_template(args):
    """random templates over the populous degree classes of an R-MAT graph (scale 17, 4 generating ranks), labelled by the
    reference's OWN degree labels (no -v): skewed degrees, hubs, parallel edges and self loops as in BASELINE's configurations"""
    _, seed = args
    from oracle import oracle as O
    from oracle import reference_run as R
    from tests import cases
    name = "rmat17_random_templates"
    st = _rmat_state()
    rng = random.Random(seed * 32452843 + 5)
    spec = random_template(rng)
    classes = rng.sample(range(2, 10), 6)  # degree classes 2..9 hold thousands of vertices each at scale 17
    # existence checks only (cycle / path constraints with work aggregation): a generated enumeration walk over populous degree
    # classes is combinatorial on an R-MAT graph; enumeration on R-MAT is covered by the committed scale-17..21 fixtures
    spec = dict(spec, labels=[classes[l - 1] for l in spec["labels"]], constraints=[c for c in spec["constraints"] if not c.get("tds")][:4])  # the driver enumerates from constraint 4 on
    d = cases.pattern_dir(spec)
    try:
        pat = O.Pattern(d)
    except Exception:  # noqa: BLE001
        return name, "refused_by_the_reader", None
    run = O.Run(st["g"], st["labels"], pat, tds_from_pl=4, max_iterations=60)
    if run.hazards[:3].any() or run.hazards[4]:
        return name, "order_dependent", None
    want = cases.run_summary(run)
    out = os.path.join(os.environ["PM_SWEEP_DIR"], "out_%d_%d" % (os.getpid(), seed))
    p = R.launch(st["path"], os.path.dirname(d), out)
    so, se = p.communicate(timeout=900)
    if p.returncode != 0:
        return name, "reference_failed", dict(seed=seed, spec=spec, error=se[-200:])
    got = R.parse_result_tree(out)
    import shutil
    shutil.rmtree(out, ignore_errors=True)
    if not R.template_read_intact(so, spec):
        return name, "reference_misread_its_template", None
    ok = (got["rows"] == want["rows"] and got["iterations"] == want["iterations"] and got["vertices"] == sorted(want["vertices"])
          and got["edges"] == sorted(want["edges"])
          and all(got["subgraphs"].get(pl, []) == sorted(want["subgraphs"][pl]) for pl in range(4, len(want["subgraphs"]))))
    if not ok:
        return name, "MISMATCH", dict(seed=seed, spec=spec)
    return name, ("nontrivial" if want["rows"][-1][3] > 0 else "empty") + ("+multi_iteration" if want["iterations"] > 1 else "") + \
        ("+enumerated" if any(len(x) for x in want["subgraphs"]) else ""), None


def one(args):
    t_index, seed = args
    import numpy as np  # noqa: F401
    from oracle import oracle as O
    from oracle import reference_run as R
    from tests import cases
    name, spec, labelset = templates()[t_index]
    rng = random.Random(seed * 7919 + t_index)
    n = rng.choice([12, 40, 90, 200, 500])
    m = int(n * rng.choice([1.0, 2.0, 3.5, 6.0]))
    kind = rng.choice(["random", "random", "planted"])
    if kind == "planted" and n >= 90:
        edges, labels = cases.planted(seed, n, m, spec, labelset, copies=rng.choice([1, 3, 6]))
    else:
        edges = cases.random_multigraph(seed, n, m, dup=rng.choice([0.0, 0.1, 0.4]), loops=rng.choice([0.0, 0.05, 0.3]))
        labels = cases.random_labels(seed, n, labelset + ([labelset[0] + 50] if rng.random() < 0.3 else []))
    if not len(edges):
        return name, "skipped", None
    d = cases.pattern_dir(spec)
    g = O.Graph.from_undirected(n, edges)
    run = O.Run(g, labels, O.Pattern(d), tds_from_pl=4, max_iterations=60)
    if run.hazards[:3].any() or run.hazards[4]:
        return name, "order_dependent", None
    want = cases.run_summary(run)
    src, dst = cases.slots_of(edges)
    got = R.run(n, src.tolist(), dst.tolist(), os.path.dirname(d), labels=labels.tolist())
    ok = (got["rows"] == want["rows"] and got["iterations"] == want["iterations"] and got["vertices"] == sorted(want["vertices"])
          and got["edges"] == sorted(want["edges"])
          and all(got["subgraphs"].get(pl, []) == sorted(want["subgraphs"][pl]) for pl in range(4, len(want["subgraphs"]))))
    if not R.template_read_intact(got["stdout"], spec):
        return name, "reference_misread_its_template", None  # undefined behaviour in graph.hpp, see template_read_intact
    if not ok:
        return name, "MISMATCH", dict(seed=seed, n=n, m=m, kind=kind)
    flags = ("nontrivial" if want["rows"][-1][3] > 0 else "empty") + ("+multi_iteration" if want["iterations"] > 1 else "") + \
        ("+resurrection" if run.hazards[3] else "") + ("+flag_outside_lcc" if run.hazards[5] else "")
    return name, flags, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=100)
    ap.add_argument("--jobs", type=int, default=max(1, (os.cpu_count() or 2) - 1))
    ap.add_argument("--path", default="beta", choices=["beta", "fuzzy", "approx", "random_templates", "rmat_templates"],
                    help="beta: run_pattern_matching_beta (LCC / NLCC); fuzzy: run_pattern_matching (run_fuzzy path); "
                         "approx: run_pattern_matching_beta_2 (first local constraint checking call); random_templates: the beta driver over "
                         "random 3..6 vertex templates; rmat_templates: the same templates over the degree classes of an R-MAT scale-17 graph")
    a = ap.parse_args()
    from oracle import oracle as O
    from oracle import reference_run as R
    O.build()
    if R.build() is None:
        sys.exit("oracle/_ref is not built and /root/reference is not here")
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import tempfile
    os.environ.setdefault("PM_SWEEP_DIR", tempfile.mkdtemp(prefix="pm_sweep_"))
    if a.path == "rmat_templates":
        _rmat_state()  # the slot file is written once, before the workers start
    ts, fn = {"beta": (templates, one), "fuzzy": (fuzzy_templates, one_fuzzy), "approx": (approx_templates, one_approx),
              "random_templates": (lambda: [("random_templates", None, None)], one_random_template),
              "rmat_templates": (lambda: [("rmat17_random_templates", None, None)], one_rmat_template)}[a.path]
    ts = ts()
    work = [(t, s) for t in range(len(ts)) for s in range(a.seeds)]
    stats, bad = {}, []
    with mp.Pool(a.jobs) as pool:
        for name, flags, info in pool.imap_unordered(fn, work, chunksize=8):
            st = stats.setdefault(name, {})
            st[flags] = st.get(flags, 0) + 1
            if info:
                bad.append((name, info))
    total = 0
    for name, _, _ in ts:
        st = stats.get(name, {})
        compared = sum(v for k, v in st.items() if k not in ("skipped", "order_dependent", "MISMATCH", "reference_misread_its_template",
                                                             "refused_by_the_reader", "reference_failed"))
        total += compared
        print("%-28s compared %4d  %s" % (name, compared, " ".join("%s=%d" % kv for kv in sorted(st.items()))))
    print("path %s: oracle == reference on %d inputs, %d mismatches" % (a.path, total, len(bad)))
    for b in bad:
        print("MISMATCH", b)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
