"""Engine against the oracle over RANDOM TEMPLATES (run as a script in its own process by
tests/test_zz_random_templates_gpu.py): the generator of oracle/sweep_vs_reference.py — random trees of 3..6 template vertices
with up to two extra edges, distinct or repeated labels, generated cycle / path constraints and a depth-first enumeration walk
at constraint 4 — on random and planted multigraphs.  The oracle was held to the reference's own driver on 19 152 such inputs
(profiles/r02_oracle_vs_reference_sweep.log).  Prints one JSON line; exit status 1 on a mismatch.
    python tests/random_templates_gpu_check.py [n_seeds]"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from fuzzypatternmatching_b200.engine import Engine, PmError
    from oracle import oracle as O
    from oracle import sweep_vs_reference as SW
    from tests import cases
    O.build()
    n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    eng = Engine(0)
    out = {"compared": 0, "nontrivial": 0, "enumerated": 0, "order_dependent": 0, "refused": 0, "mismatches": []}
    for seed in range(n_seeds):
        rng = random.Random(seed * 15485863 + 11)
        spec = SW.random_template(rng)
        labelset = sorted(set(spec["labels"]))
        n = rng.choice([12, 40, 90, 200])
        m = int(n * rng.choice([1.0, 2.0, 3.5, 6.0]))
        if rng.random() < 0.4 and n >= 40:
            edges, labels = cases.planted(seed, n, m, spec, labelset, copies=rng.choice([1, 3]))
        else:
            edges = cases.random_multigraph(seed, n, m, dup=rng.choice([0.0, 0.1, 0.4]), loops=rng.choice([0.0, 0.05, 0.3]))
            labels = cases.random_labels(seed, n, labelset)
        if not len(edges):
            continue
        d = cases.pattern_dir(spec)
        try:
            pat = O.Pattern(d)
        except Exception:  # noqa: BLE001 — a template both readers refuse (a walk longer than 16 vertices)
            continue
        ref = O.Run(O.Graph.from_undirected(n, edges), labels, pat, tds_from_pl=4, max_iterations=60)
        if ref.hazards[:3].any() or ref.hazards[4]:
            out["order_dependent"] += 1  # the reference itself is order dependent here: nothing to compare with
            continue
        src, dst = cases.slots_of(edges)
        try:
            eng.graph_from_slots(n, src, dst)
            eng.labels_set(labels)
            eng.pattern_load_dir(d)
            eng.run(tds_from_pl=4, max_iterations=60)
        except PmError as e:
            # a nem_1 constraint whose result depends on message order in the reference is refused
            if "UNSUPPORTED" in str(e).upper() or "order" in str(e).lower():
                out["refused"] += 1
                continue
            out["mismatches"].append({"seed": seed, "error": str(e)[-300:]})
            continue
        got, want = cases.engine_summary(eng, len(spec["constraints"])), cases.run_summary(ref)
        bad = [k for k in ("rows", "iterations", "vertices", "edges", "subgraphs") if got[k] != want[k]]
        if bad:
            out["mismatches"].append({"seed": seed, "differs": bad, "spec": spec, "n": n, "m": m})
            continue
        out["compared"] += 1
        out["nontrivial"] += want["rows"][-1][3] > 0
        out["enumerated"] += any(len(x) for x in want["subgraphs"])
    # the same generator over the degree classes of an R-MAT graph (scale 17, 4 generating ranks, degree labels): skewed
    # degrees, hubs, parallel edges; existence constraints only (oracle/sweep_vs_reference.py::one_rmat_template)
    out.update({"rmat_compared": 0, "rmat_nontrivial": 0})
    g = O.Graph.rmat(17, 4)
    glabels = g.labels_degree_log2()
    eng.graph_rmat(17, 4)
    eng.labels_degree_log2()
    for seed in range(max(1, n_seeds // 6)):
        rng = random.Random(seed * 32452843 + 5)
        spec = SW.random_template(rng)
        classes = rng.sample(range(2, 10), 6)
        spec = dict(spec, labels=[classes[l - 1] for l in spec["labels"]],
                    constraints=[c for c in spec["constraints"] if not c.get("tds")][:4])
        d = cases.pattern_dir(spec)
        ref = O.Run(g, glabels, O.Pattern(d), tds_from_pl=4, max_iterations=60)
        if ref.hazards[:3].any() or ref.hazards[4]:
            out["order_dependent"] += 1
            continue
        try:
            eng.pattern_load_dir(d)
            eng.run(tds_from_pl=4, max_iterations=60)
        except PmError as e:
            if "UNSUPPORTED" in str(e).upper() or "order" in str(e).lower():
                out["refused"] += 1
                continue
            out["mismatches"].append({"rmat_seed": seed, "error": str(e)[-300:]})
            continue
        got, want = cases.engine_summary(eng, len(spec["constraints"])), cases.run_summary(ref)
        bad = [k for k in ("rows", "iterations", "vertices", "edges") if got[k] != want[k]]
        if bad:
            out["mismatches"].append({"rmat_seed": seed, "differs": bad, "spec": spec})
            continue
        out["rmat_compared"] += 1
        out["rmat_nontrivial"] += want["rows"][-1][3] > 0
    eng.close()
    print(json.dumps(out))
    return 1 if out["mismatches"] else 0


if __name__ == "__main__":
    sys.exit(main())
