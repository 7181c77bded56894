"""The drop-in boundary on a CPU-only box: libpmgpu.so builds for sm_100a, loads, and
exports exactly the entry points include/pmgpu.h declares; no compute without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from fuzzypatternmatching_b200 import build
    build.build()
    from fuzzypatternmatching_b200 import _lib
    return _lib


def _declared():
    src = open(os.path.join(ROOT, "include", "pmgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pm_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(lib):
    assert _declared() == sorted(lib.SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    so = C.CDLL(lib.LIB_PATH)
    for name in _declared():
        assert hasattr(so, name), name
    lib.load()


def test_every_entry_point_cites_the_reference_interface_it_replaces():
    src = open(os.path.join(ROOT, "include", "pmgpu.h")).read()
    for needle in ("label_propagation_pattern_matching_nonunique_ee.hpp:1029-1040",
                   "token_passing_pattern_matching_nonunique_nem_1.hpp:908-922",
                   "token_passing_pattern_matching_nonunique_tds_batch_1.hpp:976-984",
                   "beta.cpp:544-1351", "vertex_data_db_degree.hpp:109", "graph.hpp:73-110"):
        assert needle in src, needle


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fuzzypatternmatching_b200.engine import Engine, PmError
    with pytest.raises(PmError):
        Engine(0)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "fuzzypatternmatching_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pm_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_pattern_directory_roundtrip(tmp_path, lib):
    from fuzzypatternmatching_b200 import patterns as PT
    d = PT.write_pattern_dir(str(tmp_path), PT.RMAT_LOG2_TREE)
    nlc = open(os.path.join(d, "pattern_nlc")).read().splitlines()
    assert nlc[0] == "3 5 2 4 3 : 4 5 3 1 0 : 3 : 0 : 1 : 0"
    assert nlc[4] == "3 4 7 4 2 5 3 5 7 : 0 1 2 1 3 5 4 5 6 : 7 : 0 : 1 : 0"
    enum = open(os.path.join(d, "pattern_non_local_constraint")).read().splitlines()
    assert enum[4] == "0 1 2 1 3 5 4 5 6 : 0 1 2 1 4 5 6 5 8 : 0 1 1 1 1 1 1 1 1"
    assert open(os.path.join(d, "pattern_stat")).read().strip() == "diameter : 8"
    assert len(open(os.path.join(d, "pattern_edge")).read().split()) == 24
