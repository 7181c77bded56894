"""GPU exploration: graph build and per-template search times at several scales.
Development tooling, not product: the optional `check` argument uses the CPU oracle as a checker, like the tests do."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from fuzzypatternmatching_b200.engine import Engine

scales = [int(x) for x in sys.argv[1].split(",")]
gen_ranks = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
check = len(sys.argv) > 3 and sys.argv[3] == "check"
pats = bench.write_patterns(os.environ.get("PM_WORKLOAD", "cyclic"))
eng = Engine(0)
for s in scales:
    t = time.time(); eng.graph_rmat(s, gen_ranks); eng.labels_degree_log2(); gi = eng.graph_info()
    print("scale", s, "build %.2fs" % (time.time() - t), gi, flush=True)
    if check:
        from oracle import oracle as O
        g = O.Graph.rmat(s, gen_ranks); lab = g.labels_degree_log2()
    for name, d, tds in pats:
        eng.pattern_load_dir(d)
        for rep in range(2):
            t = time.time(); sm = eng.run(tds_from_pl=tds, keep_subgraphs=False); dt = time.time() - t
        rows = eng.rows()
        print("  %-22s %.3f ms (dev %.3f ms) iters %d rows %d final %s paths %d scanned %d launches %d" % (
            name, dt * 1e3, sm["device_seconds"] * 1e3, sm["iterations"], len(rows), rows[-1][3:], sm["path_count"],
            sm["edges_processed"], eng.kernel_launches()), flush=True)
        print("     first rows", [(r[1], r[2], r[3], r[4]) for r in rows[:4]], flush=True)
        if os.environ.get("PM_ROWS"):
            for r in eng.rows_timed():
                print("       %s" % (r,), flush=True)
        if check:
            t = time.time(); ref = O.Run(g, lab, O.Pattern(d), tds_from_pl=tds, keep_subgraphs=False)
            print("     oracle %.2fs rows_equal %s hazards %s" % (time.time() - t, ref.rows == rows, ref.hazards[:5]), flush=True)
    for b in range(5):
        print("  kstat", b, eng.kernel_stats(b))
