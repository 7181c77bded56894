"""The library's host-side pattern directory reader (csrc/pm_pattern.hpp through pm_pattern_check_dir — no GPU
needed) against the oracle's reader on the same directories, and on malformed directories
(graph.hpp:195-270, 337-358; pattern_util.hpp:172-278)."""
import os
import shutil

import pytest

from fuzzypatternmatching_b200 import patterns as PT
from fuzzypatternmatching_b200.engine import pattern_check_dir
from tests import cases


def _specs():
    out = [(n, s) for n, s, _, _ in cases.SPECS]
    out.append(("lcc_only", {"labels": [1, 2, 1], "edges": [(0, 1), (1, 2)], "diameter": 2, "constraints": []}))
    return out


@pytest.mark.parametrize("name,spec", _specs(), ids=[n for n, _ in _specs()])
def test_reader_agrees_with_the_oracle(oracle, name, spec):
    d = cases.pattern_dir(spec)
    got = pattern_check_dir(d)
    want = oracle.Pattern(d)
    assert (got["n_vertices"], got["n_edges"], got["diameter"], got["n_constraints"]) == \
        (want.n_vertices, want.n_edges, want.diameter, want.n_constraints)
    assert got["n_vertices"] == len(spec["labels"]) and got["n_edges"] == 2 * len(spec["edges"])
    assert len(got["constraints"]) == len(spec["constraints"])
    for k, c in zip(spec["constraints"], got["constraints"]):
        assert c["walk_length"] == len(k["walk"])
        assert c["valid_cycle"] == int(bool(k.get("cycle")))


def _broken(tmp_path, name, edit):
    d = str(tmp_path / name)
    shutil.copytree(cases.pattern_dir(PT.triangle(1, 2, 3)), d)
    edit(d)
    return d


def _rewrite(fname, fn):
    def edit(d):
        p = os.path.join(d, fname)
        text = open(p).read()
        open(p, "w").write(fn(text))
    return edit


# (name, edit, oracle_rejects).  Where the reference itself has no defined behaviour the oracle restates its
# leniency and only the library refuses: an empty pattern_edge makes ::graph index edge_list[-1]
# (graph.hpp:226), `iss >> s >> t` on garbage reads zeros (:199-201), and a missing or zero diameter silently
# runs zero supersteps (:337-358, ee.hpp:1069) — a drop-in should say so instead.
BROKEN = [
    ("no_edge_file", lambda d: os.remove(os.path.join(d, "pattern_edge")), True),
    ("empty_edge_file", _rewrite("pattern_edge", lambda t: ""), False),
    ("unsorted_edges", _rewrite("pattern_edge", lambda t: "\n".join(reversed(t.strip().split("\n"))) + "\n"), True),
    ("vertex_id_16", _rewrite("pattern_edge", lambda t: t + "16 0\n"), True),
    ("edge_not_numeric", _rewrite("pattern_edge", lambda t: t.replace("0 1", "0 x", 1)), False),
    ("no_stat", lambda d: os.remove(os.path.join(d, "pattern_stat")), False),
    ("diameter_zero", _rewrite("pattern_stat", lambda t: "diameter : 0\n"), False),
    ("diameter_missing", _rewrite("pattern_stat", lambda t: "radius : 3\n"), False),
    ("no_vertex_data", lambda d: os.remove(os.path.join(d, "pattern_vertex_data")), True),
    ("nlc_too_few_fields", _rewrite("pattern_nlc", lambda t: "1 2 3 1 : 0 1 2 0 : 2\n"), True),
    ("nlc_walk_length_mismatch", _rewrite("pattern_nlc", lambda t: t.replace(": 2 :", ": 3 :", 1)), True),
]


@pytest.mark.parametrize("name,edit,oracle_rejects", BROKEN, ids=[b[0] for b in BROKEN])
def test_malformed_directories_are_rejected(oracle, tmp_path, name, edit, oracle_rejects):
    d = _broken(tmp_path, name, edit)
    with pytest.raises(ValueError) as e:
        pattern_check_dir(d)
    assert str(e.value)
    if oracle_rejects:
        with pytest.raises(ValueError):
            oracle.Pattern(d)
    else:
        oracle.Pattern(d)  # restates the reference's leniency


def test_stat_key_is_case_insensitive_like_the_reference(oracle, tmp_path):
    # graph.hpp:337-358 lower-cases the key before comparing
    d = _broken(tmp_path, "upper", _rewrite("pattern_stat", lambda t: "Diameter : 2\n"))
    assert pattern_check_dir(d)["diameter"] == 2 and oracle.Pattern(d).diameter == 2
