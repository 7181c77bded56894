"""The per-rank count files of a G-rank run add up to the single-rank run's rows
(fuzzypatternmatching_b200/total_active_count.py, the py3 stand-in for the reference's
examples/scripts/total_active_count.py)."""
import os
import subprocess
import sys

from fuzzypatternmatching_b200 import patterns as PT
from fuzzypatternmatching_b200.total_active_count import total_counts
from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rank_files_add_up_to_the_single_rank_rows(oracle, tmp_path):
    spec = PT.triangle(1, 2, 3)
    d = cases.pattern_dir(spec)
    edges = cases.random_multigraph(3, 80, 400)
    labels = cases.random_labels(3, 80, [1, 2, 3])
    g = oracle.Graph.from_undirected(80, edges)
    one = oracle.Run(g, labels, oracle.Pattern(d), n_ranks=1, tds_from_pl=1, max_iterations=50)
    for ranks in (2, 3):
        out = str(tmp_path / ("r%d" % ranks))
        oracle.make_result_tree(out)
        oracle.Run(g, labels, oracle.Pattern(d), n_ranks=ranks, tds_from_pl=1, max_iterations=50).write_results(out)
        for sub, col in (("all_ranks_active_vertices_count", 3), ("all_ranks_active_edges_count", 4)):
            rows, n_files = total_counts(os.path.join(out, "0", sub))
            assert n_files == ranks
            assert [(int(p[0]), p[1], int(p[2]), t) for p, t in rows] == [(r[0], r[1], r[2], r[col]) for r in one.rows]
    # the command line prints the reference script's lines
    p = subprocess.run([sys.executable, "-m", "fuzzypatternmatching_b200.total_active_count",
                        os.path.join(out, "0", "all_ranks_active_vertices_count")], cwd=ROOT,
                       stdout=subprocess.PIPE, text=True, check=True)
    lines = p.stdout.strip().split("\n")
    assert lines[0].startswith("3 files to process")
    assert lines[3] == "Total number of iterations: %d" % len(one.rows) and lines[-1] == "Done."
    assert lines[4] == "%d,%s,%d,%d" % (one.rows[0][0], one.rows[0][1], one.rows[0][2], one.rows[0][3])


REF_SCRIPT = "/root/reference/examples/scripts/total_active_count.py"


def test_prints_what_the_reference_script_prints(oracle, tmp_path):
    """The reference's own examples/scripts/total_active_count.py (Python 2: it imports `thread` and `urlparse` without using
    them, everything else runs under Python 3) executed from where it lies, with those two module names pre-seeded, on the
    per-rank count files of a 3-rank run: our stand-in prints the same lines."""
    import pytest
    if not os.path.exists(REF_SCRIPT):
        pytest.skip("needs the reference tree")
    spec = PT.RMAT_LOG2_TREE
    d = cases.pattern_dir(spec)
    edges, labels = cases.planted(2, 300, 900, spec, [2, 3, 4, 5, 7])
    g = oracle.Graph.from_undirected(300, edges)
    out = str(tmp_path / "r3")
    oracle.make_result_tree(out)
    oracle.Run(g, labels, oracle.Pattern(d), n_ranks=3, tds_from_pl=4, max_iterations=50).write_results(out)
    launcher = ("import sys, types, runpy; sys.modules['thread'] = types.ModuleType('thread'); "
                "u = types.ModuleType('urlparse'); u.urlparse = None; sys.modules['urlparse'] = u; "
                "sys.argv = [%r, sys.argv[1]]; runpy.run_path(%r, run_name='__main__')" % (REF_SCRIPT, REF_SCRIPT))
    for sub in ("all_ranks_active_vertices_count", "all_ranks_active_edges_count"):
        target = os.path.join(out, "0", sub)
        ref = subprocess.run([sys.executable, "-c", launcher, target], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert ref.returncode == 0, ref.stderr[-800:]
        ours = subprocess.run([sys.executable, "-m", "fuzzypatternmatching_b200.total_active_count", target], cwd=ROOT,
                              stdout=subprocess.PIPE, text=True, check=True)
        assert ours.stdout == ref.stdout and "Total number of iterations" in ref.stdout, sub
