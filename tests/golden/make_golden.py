"""Generates tests/golden/*.json from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference cannot be built or imported
here (MPI + Boost), so these fixtures freeze the ORACLE's answers after it has been
pinned by tests/test_oracle_kat.py and tests/test_oracle_literal.py; the GPU parity
tests and later refactors of the oracle are then held to them."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests import cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

rmat = {}
for scale, rank in ((17, 0), (17, 1), (21, 0), (26, 5)):
    rmat["%d_%d" % (scale, rank)] = O.rmat_stream(scale, rank, 8).tolist()
json.dump(rmat, open(os.path.join(HERE, "rmat_kat.json"), "w"), indent=0)

runs = []
for name, spec, labelset, tds_from in cases.SPECS:
    for seed, n, m in ((1, 60, 240), (2, 70, 300), (5, 90, 420)):
        edges = cases.random_multigraph(seed, n, m)
        labels = cases.random_labels(seed, n, labelset)
        pat = O.Pattern(cases.pattern_dir(spec))
        g = O.Graph.from_undirected(n, edges)
        r = O.Run(g, labels, pat, tds_from_pl=tds_from, max_iterations=50)
        assert not r.hazards[:5].any()
        s = cases.run_summary(r)
        runs.append(dict(spec=name, seed=seed, n=n, m=m, rows=s["rows"], vertices=s["vertices"],
                         edges=s["edges"], subgraphs=s["subgraphs"], iterations=s["iterations"]))
json.dump(runs, open(os.path.join(HERE, "runs.json"), "w"))
print("wrote", len(runs), "runs")
