"""BASELINE configs[3]: run_fuzzy_pattern_matching over an edit-distance-k prototype set on R-MAT.

  python scripts/fuzzy_prototypes.py [scale=24] [gen_ranks=1024] [k=2] [check]

For every connected template within k edge deletions of the base template (4-cycle with degree
labels 5,6,7,8 and the 6-cycle with chords 4..9) the run_fuzzy path prunes the graph and the
surviving vertices are counted.  `check` compares every run with the CPU oracle (small scales): that mode is test
tooling (the oracle as the checker); without it the script only drives the GPU path."""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fuzzypatternmatching_b200 import patterns as PT  # noqa: E402
from fuzzypatternmatching_b200.engine import Engine  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
gen_ranks = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
k = int(sys.argv[3]) if len(sys.argv) > 3 else 2
check = len(sys.argv) > 4 and sys.argv[4] == "check"
eng = Engine(0)
t = time.time()
eng.graph_rmat(scale, gen_ranks)
eng.labels_degree_log2()
gi = eng.graph_info()
print("R-MAT scale %d (%d generating ranks): %d vertices, %d directed slots, build %.2f s" % (
    scale, gen_ranks, gi["n_vertices"], gi["n_slots_multi"], time.time() - t), flush=True)
if check:
    from oracle import oracle as O
    g = O.Graph.rmat(scale, gen_ranks)
    lab = g.labels_degree_log2()
total_ms, n_runs = 0.0, 0
for name, base in (("cycle4_5678", PT.cycle4(5, 6, 7, 8)), ("cycle6_chords_456789", PT.cycle6_chords([4, 5, 6, 7, 8, 9]))):
    base = dict(base, constraints=[c for c in base["constraints"] if not c.get("tds")])  # this path has no TDS walker
    protos = PT.write_prototype_set(tempfile.mkdtemp(prefix="pm_proto_"), base, k)
    print("%s: %d prototypes within edit distance %d" % (name, len(protos), k), flush=True)
    for i, gone, d in protos:
        eng.pattern_load_dir(d)
        eng.run_fuzzy()  # warm (sizes the token pool)
        s = eng.run_fuzzy()
        rows = eng.rows()
        total_ms += s["device_seconds"] * 1e3
        n_runs += 1
        line = "  #%-3d deleted %-22s iterations %d  active vertices %-9d search %.3f ms" % (
            i, str(gone), s["iterations"], rows[-1][3], s["device_seconds"] * 1e3)
        if check:
            ref = O.Run(g, lab, O.Pattern(d), fuzzy=True)
            line += "  oracle rows equal: %s" % (ref.rows == rows)
        print(line, flush=True)
print("%d prototype searches, %.3f ms device time in total, %.3g directed slots searched per second" % (
    n_runs, total_ms, gi["n_slots_multi"] * n_runs / (total_ms * 1e-3)), flush=True)
eng.close()
