"""The C++ CLI twins (fuzzypatternmatching_b200/bin, sources csrc/cli/) on a box WITHOUT a GPU: the command-line contract of
the reference drivers (flags, usage text, exit codes: src/run_pattern_matching_beta.cpp:67-142, src/generate_rmat.cpp:78-150,
src/ingest_edge_list.cpp) and the loud failure of a product that has no CPU path.  The compute side of the same binaries is
covered by the GPU suite (tests/test_gpu_parity.py: test_cli_*)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "fuzzypatternmatching_b200", "bin")
REF = "/root/reference/src"
CLIS = ("run_pattern_matching_beta", "generate_rmat", "ingest_edge_list")


@pytest.fixture(scope="module", autouse=True)
def _built():
    from fuzzypatternmatching_b200 import build as B
    B.build()
    for c in CLIS:
        assert os.access(os.path.join(BIN, c), os.X_OK), c


def _run(name, *args, timeout=60):
    return subprocess.run([os.path.join(BIN, name)] + list(args), capture_output=True, text=True, timeout=timeout)


def _has_gpu():
    import ctypes
    from fuzzypatternmatching_b200 import _lib
    h = ctypes.c_void_p()
    L = _lib.load()
    if L.pm_create(ctypes.byref(h), 0) != 0:
        return False
    L.pm_destroy(h)
    return True


@pytest.mark.parametrize("name", CLIS)
def test_help_and_missing_required_flags_print_usage_and_fail(name):
    """the reference returns -1 from main after usage() (beta.cpp:96, 138-141): exit status 255, usage on stderr"""
    for args in (("-h",), ()):
        p = _run(name, *args)
        assert p.returncode == 255, (name, args, p.returncode)
        assert p.stderr.startswith("Usage: "), (name, args)
        assert " -h            - print help and exit" in p.stderr


def _reference_usage_lines(source):
    """the string literals of the reference's usage(): one flag per line"""
    text = open(os.path.join(REF, source)).read()
    body = text[text.index("void usage()"):]
    body = body[:body.index("\n}\n")]
    lits = re.findall(r'"((?:[^"\\]|\\.)*)"', body)
    lines = "".join(l.replace('\\n', "\n").replace('\\"', '"') for l in lits).splitlines()
    return [l for l in lines if l.strip()]


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree")
@pytest.mark.parametrize("name,source", [("run_pattern_matching_beta", "run_pattern_matching_beta.cpp"),
                                         ("generate_rmat", "generate_rmat.cpp"), ("ingest_edge_list", "ingest_edge_list.cpp")])
def test_usage_text_keeps_every_flag_line_of_the_reference(name, source):
    """every flag line of the reference's usage() appears in the twin's (which adds its own -t / -n / -r lines); the one line the
    reference completes at run time with its MPI world size is compared up to that number"""
    squeeze = lambda t: re.sub(r"[ \t]+", " ", t)  # noqa: E731 — the twins align the columns the reference leaves ragged
    ours = squeeze(_run(name, "-h").stderr)
    ref = _reference_usage_lines(source)
    assert len(ref) >= 5, ref
    for line in ref:
        probe = line.split("batch size is")[0] if "batch size is" in line else line
        assert squeeze(probe).rstrip() in ours, (name, line)


def test_without_a_gpu_every_cli_fails_loudly_and_at_once(tmp_path):
    """no CPU fallback: the drivers say so on stderr and exit non-zero instead of computing anything on the host; the
    multi-rank launcher (-n: one forked process per GPU) reports every rank and does not wait for absent peers"""
    if _has_gpu():
        pytest.skip("this box has a CUDA device")
    from fuzzypatternmatching_b200 import patterns as PT
    pdir = str(tmp_path / "pattern")
    PT.write_pattern_dir(pdir, PT.RMAT_LOG2_TREE)
    out = str(tmp_path / "out")
    os.makedirs(out)
    p = _run("generate_rmat", "-s", "10", "-o", str(tmp_path / "g"))
    assert p.returncode == 1 and "no CUDA device" in p.stderr and not os.path.exists(str(tmp_path / "g"))
    p = _run("run_pattern_matching_beta", "-i", "rmat:10:4", "-p", pdir, "-o", out)
    assert p.returncode == 1 and "no CUDA device 0" in p.stderr
    p = _run("run_pattern_matching_beta", "-n", "2", "-i", "rmat:10:4", "-p", pdir, "-o", out, timeout=30)
    assert p.returncode == 1 and "no CUDA device 0" in p.stderr and "no CUDA device 1" in p.stderr
    assert os.listdir(out) == []  # nothing was written
    edges = tmp_path / "edges_0.txt"
    edges.write_text("0 1\n1 2\n")
    p = _run("ingest_edge_list", "-o", str(tmp_path / "h"), "-u", "1", str(edges))
    assert p.returncode != 0 and "CUDA" in p.stderr
