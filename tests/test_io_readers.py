"""Host-side readers of the reference's text inputs (csrc/pm_io.hpp through the C ABI, no GPU): the -v vertex metadata
files (include/havoqgt/vertex_data_db.hpp:139-262), the -e edge metadata files and the edge lists of ingest_edge_list
(include/havoqgt/parallel_edge_list_reader.hpp:236-262), against plain Python restatements; the result merger against
an oracle-written multi-rank result tree."""
import os
import random

import numpy as np
import pytest

from fuzzypatternmatching_b200 import engine as E
from fuzzypatternmatching_b200 import merge_results as MR
from tests import cases


def test_vertex_data_files_by_prefix(tmp_path):
    rng = random.Random(3)
    n = 500
    want = np.zeros(n, dtype=np.uint64)
    d = tmp_path / "meta"
    d.mkdir()
    # the base names a directory and a file-name PREFIX: vdata_0, vdata_1, vdata.extra are read, other_0 is not
    pairs = [(rng.randrange(n), rng.randrange(1, 40)) for _ in range(900)]
    for i, name in enumerate(["vdata_0", "vdata_1", "vdata.extra"]):
        with open(d / name, "w") as f:
            for v, l in pairs[i * 300:(i + 1) * 300]:
                f.write("%d %d\n" % (v, l))
            f.write("\n")
    with open(d / "other_0", "w") as f:
        f.write("0 999\n")
    for name in sorted(["vdata_0", "vdata_1", "vdata.extra"]):  # name order, later pairs win
        i = ["vdata_0", "vdata_1", "vdata.extra"].index(name)
        for v, l in pairs[i * 300:(i + 1) * 300]:
            want[v] = l
    got, n_pairs = E.read_vertex_data(str(d / "vdata"), n)
    assert n_pairs == 900 and np.array_equal(got, want)


def test_vertex_data_errors(tmp_path):
    with pytest.raises(ValueError, match="Invalid directory"):
        E.read_vertex_data(str(tmp_path / "nope" / "x"), 10)
    with pytest.raises(ValueError, match="Failed to read input files"):
        E.read_vertex_data(str(tmp_path / "x"), 10)
    (tmp_path / "bad_0").write_text("3 4\nfoo 1\n")
    with pytest.raises(ValueError, match="bad_0:2"):
        E.read_vertex_data(str(tmp_path / "bad"), 10)
    (tmp_path / "far_0").write_text("30 4\n")
    with pytest.raises(ValueError, match="not in the graph"):
        E.read_vertex_data(str(tmp_path / "far"), 10)
    (tmp_path / "huge_0").write_text("3 123456789012345678901234567890\n")
    with pytest.raises(ValueError, match="huge_0:1"):
        E.read_vertex_data(str(tmp_path / "huge"), 10)


def test_edge_data_files_are_validated(tmp_path):
    (tmp_path / "edata_0").write_text("0 1 7\n1 0 7\n")
    (tmp_path / "edata_1").write_text("2 3 9\n")
    assert E.check_edge_data(str(tmp_path / "edata"), 4) == 3
    (tmp_path / "ebad_0").write_text("0 1\n")
    with pytest.raises(ValueError, match="ebad_0:1"):
        E.check_edge_data(str(tmp_path / "ebad"), 4)
    with pytest.raises(ValueError, match="edata_1:1"):
        E.check_edge_data(str(tmp_path / "edata"), 3)  # vertex 3 is outside a 3-vertex graph


def test_edge_lists_directed_and_undirected(tmp_path):
    rng = random.Random(5)
    edges = [(rng.randrange(300), rng.randrange(300)) for _ in range(2000)]
    with open(tmp_path / "a.txt", "w") as f:
        f.write("# comment\n")
        for s, t in edges[:1000]:
            f.write("%d %d\n" % (s, t))
    with open(tmp_path / "b.txt", "w") as f:
        for s, t in edges[1000:]:
            f.write("%d\t%d 17\n" % (s, t))  # weights are read and dropped
        f.write("\n")
    files = [str(tmp_path / "a.txt"), str(tmp_path / "b.txt")]
    nv, src, dst = E.read_edge_lists(files, undirected=False)
    assert nv == max(max(e) for e in edges) + 1
    assert list(zip(src.tolist(), dst.tolist())) == edges
    nv2, src, dst = E.read_edge_lists(files, undirected=True)
    want = []
    for s, t in edges:
        want += [(s, t), (t, s)]
    assert nv2 == nv and list(zip(src.tolist(), dst.tolist())) == want
    (tmp_path / "c.txt").write_text("1 x\n")
    with pytest.raises(ValueError, match="c.txt:1"):
        E.read_edge_lists([str(tmp_path / "c.txt")])
    with pytest.raises(ValueError, match="cannot open"):
        E.read_edge_lists([str(tmp_path / "missing.txt")])


def test_result_merger_over_ranks_and_pattern_set(oracle, tmp_path):
    """Two elements of a pattern set written by the oracle with 3 ranks: the merger gives back the oracle's sets."""
    from fuzzypatternmatching_b200 import patterns as PT
    n, m = 80, 400
    want = {}
    out = str(tmp_path / "out")
    os.makedirs(out)
    for ps, (name, spec, labelset, tds) in enumerate([cases.SPECS[1], cases.SPECS[2]]):
        oracle.make_result_tree(out, ps)
        edges = cases.random_multigraph(2, n, m)
        labels = cases.random_labels(2, n, [1, 2, 3, 4])
        g = oracle.Graph.from_undirected(n, edges)
        r = oracle.Run(g, labels, oracle.Pattern(cases.pattern_dir(spec)), n_ranks=3, tds_from_pl=tds, max_iterations=50)
        elem = str(tmp_path / ("elem%d" % ps))
        os.makedirs(elem)
        oracle.make_result_tree(elem)
        r.write_results(elem)
        os.rename(os.path.join(elem, "0"), os.path.join(out, str(ps) + "_tmp"))
        import shutil
        shutil.rmtree(os.path.join(out, str(ps)))
        os.rename(os.path.join(out, str(ps) + "_tmp"), os.path.join(out, str(ps)))
        want[ps] = cases.run_summary(r)
    merged = MR.write_merged(out, str(tmp_path / "merged"))
    assert sorted(merged["elements"]) == [0, 1]
    for ps in (0, 1):
        el = merged["elements"][ps]
        assert [(v, int(b[1], 2)) for v, b in el["vertices"].items()] == want[ps]["vertices"]
        assert el["edges"] == want[ps]["edges"]
        for pl, rows in el["subgraphs"].items():
            assert rows == want[ps]["subgraphs"][pl]
    assert set(merged["union_vertices"]) == {v for ps in (0, 1) for v, _ in want[ps]["vertices"]}
    assert merged["union_edges"] == sorted(set(want[0]["edges"]) | set(want[1]["edges"]))
    assert os.path.exists(str(tmp_path / "merged" / "union_active_edges"))
