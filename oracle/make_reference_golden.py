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
    write_rmat_fixture()


RMAT_FIXTURE = os.path.join(ROOT, "tests", "golden", "rmat_reference_generator.json")
RMAT_FIXTURE_STREAMS = [(17, 0, 4), (17, 3, 4), (21, 0, 4), (21, 2, 4), (25, 0, 1024), (26, 0, 1024), (26, 1023, 1024),
                        (28, 517, 1024), (32, 1, 4)]


def write_rmat_fixture():
    """tests/golden/rmat_reference_generator.json: the first 64 generated edges of a few generating ranks and hash_nbits of
    a few values, from the reference's OWN rmat_edge_generator.hpp / detail/hash.hpp (oracle/_ref/rmat_edge_dump)"""
    if not os.access(R.BINARY_RMAT, os.X_OK):
        return
    import random
    rng = random.Random(7)
    doc = {"produced_by": "oracle/_ref/rmat_edge_dump (the reference's rmat_edge_generator.hpp + detail/hash.hpp, constructed like "
                          "src/generate_rmat.cpp:202-205) via oracle/make_reference_golden.py",
           "streams": [], "hash_nbits": []}
    for scale, rank, ranks in RMAT_FIXTURE_STREAMS:
        pairs = R.rmat_edge_dump(scale, rank, ranks, 64)
        assert np.array_equal(pairs[0::2], pairs[1::2][:, ::-1])
        doc["streams"].append({"scale": scale, "rank": rank, "ranks": ranks, "edges": pairs[0::2].tolist()})
    for n in range(17, 33):
        xs = [0, 1, (1 << n) - 1] + [rng.randrange(1 << n) for _ in range(13)]
        doc["hash_nbits"].append({"n": n, "x": xs, "hash": R.reference_hash_nbits(xs, n)})
    with open(RMAT_FIXTURE, "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    print("wrote " + RMAT_FIXTURE)


if __name__ == "__main__":
    main()
