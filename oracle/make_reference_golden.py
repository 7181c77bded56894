"""Writes tests/golden/reference_runs/*.json: outputs of THE REFERENCE ITSELF (oracle/_ref/run_pattern_matching_beta, see
oracle/ref_shim/README.md) on seeded inputs that tests/cases.py regenerates anywhere.  Run it in the container that holds
/root/reference:   python oracle/make_reference_golden.py
(`--large` also writes the R-MAT scale-20 / scale-21 fixtures, minutes each; `--only=<name>` restricts that to one of them.)
Each file names its input generator and carries the reference's count rows, iteration count, final vertex -> template
bitset map, final edge set and enumerated subgraphs (template-driven search starts at constraint 4 in the driver,
beta.cpp:725-730)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != os.path.join(ROOT, "oracle")]  # `oracle` is the package
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import reference_run as R  # noqa: E402
from tests import cases  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "reference_runs")


def main():
    if R.build() is None:
        sys.exit("oracle/_ref is not built and /root/reference is not here")
    O.build()
    os.makedirs(OUT, exist_ok=True)
    n_files = 0
    large = "--large" in sys.argv
    for case in cases.reference_golden_cases():
        if case.get("large"):
            if not large:
                continue  # kept as committed; regenerate with --large
            only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
            if only and case["name"] not in only:
                continue
            scale, gen = case["scale"], case["gen_ranks"]
            e = np.concatenate([O.rmat_stream(scale, r, (16 << scale) // gen) for r in range(gen)])
            src = np.empty(2 * len(e), dtype=np.uint64)
            dst = np.empty(2 * len(e), dtype=np.uint64)
            src[0::2], dst[0::2] = e[:, 0], e[:, 1]
            src[1::2], dst[1::2] = e[:, 1], e[:, 0]
            if "bench_template" in case:  # a template of bench.py's workload, enumeration walk at constraint 4
                pdir = os.path.dirname(cases.pattern_dir(cases.bench_template_at_constraint_4(case["bench_template"])[0]))
            else:
                pdir = os.path.join(ROOT, "tests", case["pattern_dir"])
            got = R.run(1 << scale, src, dst, pdir, labels=None, timeout=3600)
            got.pop("stdout")
            if case.get("digest_subgraphs"):  # too many walks to commit: their count and digest
                got["subgraphs_digest"] = {str(k): cases.subgraphs_digest(v) for k, v in got["subgraphs"].items() if v}
                got["subgraphs"] = {}
            got["subgraphs"] = {str(k): v for k, v in got["subgraphs"].items() if v}
            with open(os.path.join(OUT, case["name"] + ".json"), "w") as f:
                json.dump({"case": case, "reference": got,
                           "produced_by": "oracle/_ref/run_pattern_matching_beta (reference driver + visitor headers, single-rank "
                                          "runtime stand-in) via oracle/make_reference_golden.py --large"}, f, separators=(",", ":"))
            n_files += 1
            continue
        n, edges, labels, spec = cases.reference_golden_input(case, O)
        d = cases.pattern_dir(spec)
        src, dst = cases.slots_of(edges)
        if case.get("path") == "approx_first_lcc":
            got = {"rows": R.run_approx_first_lcc(n, src.tolist(), dst.tolist(), os.path.dirname(d), np.asarray(labels).tolist(), spec)}
        elif case.get("path") == "run_fuzzy":
            got = R.run_fuzzy(n, src.tolist(), dst.tolist(), os.path.dirname(d), np.asarray(labels).tolist())
            got.pop("stdout")
        else:
            got = R.run(n, src.tolist(), dst.tolist(), os.path.dirname(d),
                        labels=None if case["labels"] == "degree_log2" else np.asarray(labels).tolist())
            got.pop("stdout")
            got["subgraphs"] = {str(k): v for k, v in got["subgraphs"].items() if v}
        with open(os.path.join(OUT, case["name"] + ".json"), "w") as f:
            json.dump({"case": case, "reference": got,
                       "produced_by": "oracle/_ref/%s (reference driver + visitor headers, single-rank runtime stand-in) via "
                                      "oracle/make_reference_golden.py"
                                      % {"run_fuzzy": "run_pattern_matching", "approx_first_lcc": "run_pattern_matching_beta_2"}.get(case.get("path"), "run_pattern_matching_beta")}, f, separators=(",", ":"))
        n_files += 1
    print("wrote %d files to %s" % (n_files, OUT))


if __name__ == "__main__":
    main()
