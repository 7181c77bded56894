"""Merges the per-rank result files of a run ("result merger").

The reference writes one file per MPI rank and kind under `<out>/<ps>/all_ranks_*/` (src/run_pattern_matching_beta.cpp:
504-535, 1386-1425; row grammar in SURVEY A.5) and leaves putting them together to the user; its only helper,
examples/scripts/total_active_count.py, sums the count files (ported in total_active_count.py).  This module reads the
row files of every rank of one pattern-set element, or of the whole set, into sorted duplicate-free tables:

    python -m fuzzypatternmatching_b200.merge_results <out> [<merged dir>]

  active vertices   {vertex: (label, 16-character template bitset)}
  active edges      sorted (vertex, neighbour) pairs
  subgraphs         per constraint, sorted walks (the vertex ids between the two bracketed fields)
Over a pattern set (`<out>/0`, `<out>/1`, ... — the edit-distance prototypes of BASELINE configs[3]) the union of the
elements' vertex and edge sets is the approximate-match solution subgraph.
"""
import os
import re
import sys


def _rank_files(d, stem):
    pat = re.compile(r"^%s_(\d+)$" % re.escape(stem))
    if not os.path.isdir(d):
        return []
    out = []
    for name in os.listdir(d):
        m = pat.match(name)
        if m:
            out.append((int(m.group(1)), os.path.join(d, name)))
    return [p for _, p in sorted(out)]


def _fields(line):
    return [f.strip() for f in line.strip().split(",") if f.strip() != ""]


def merge_element(outdir, ps=0):
    """One element of the pattern set: every rank's rows of <outdir>/<ps>/all_ranks_*."""
    base = os.path.join(outdir, str(ps))
    vertices, edges, subgraphs = {}, set(), {}
    for path in _rank_files(os.path.join(base, "all_ranks_active_vertices"), "active_vertices"):
        for line in open(path):
            f = _fields(line)
            if len(f) >= 5:  # rank, vertex, 0, label, bitset (beta.cpp:1390-1394)
                vertices[int(f[1])] = (int(f[3]), f[4])
    for path in _rank_files(os.path.join(base, "all_ranks_active_edges"), "active_edges"):
        for line in open(path):
            f = _fields(line)
            if len(f) >= 3:  # rank, vertex, neighbour (beta.cpp:1398-1403)
                edges.add((int(f[1]), int(f[2])))
    sub_dir = os.path.join(base, "all_ranks_subgraphs")
    if os.path.isdir(sub_dir):
        for name in sorted(os.listdir(sub_dir)):
            m = re.match(r"^subgraphs_(\d+)_(\d+)$", name)
            if not m:
                continue
            pl = int(m.group(1))
            rows = subgraphs.setdefault(pl, [])
            for line in open(os.path.join(sub_dir, name)):
                f = _fields(line)
                if len(f) >= 3:  # [rank], v0, ..., vh, [vh] (tds_batch_1.hpp:685-689)
                    rows.append(tuple(int(x) for x in f[1:-1]))
    return {"vertices": dict(sorted(vertices.items())), "edges": sorted(edges),
            "subgraphs": {pl: sorted(rows) for pl, rows in sorted(subgraphs.items())}}


def pattern_set_elements(outdir):
    return sorted(int(n) for n in os.listdir(outdir) if n.isdigit() and os.path.isdir(os.path.join(outdir, n)))


def merge_set(outdir):
    """Every element of the set plus the union of their vertex and edge sets."""
    elements = {ps: merge_element(outdir, ps) for ps in pattern_set_elements(outdir)}
    union_v, union_e = {}, set()
    for ps, el in elements.items():
        for v, (label, _bits) in el["vertices"].items():
            union_v.setdefault(v, (label, []))[1].append(ps)
        union_e.update(el["edges"])
    return {"elements": elements, "union_vertices": dict(sorted(union_v.items())), "union_edges": sorted(union_e)}


def write_merged(outdir, merged_dir):
    os.makedirs(merged_dir, exist_ok=True)
    m = merge_set(outdir)
    for ps, el in m["elements"].items():
        with open(os.path.join(merged_dir, "active_vertices_%d" % ps), "w") as f:
            for v, (label, bits) in el["vertices"].items():
                f.write("%d, %d, %s\n" % (v, label, bits))
        with open(os.path.join(merged_dir, "active_edges_%d" % ps), "w") as f:
            for a, b in el["edges"]:
                f.write("%d, %d\n" % (a, b))
        for pl, rows in el["subgraphs"].items():
            with open(os.path.join(merged_dir, "subgraphs_%d_%d" % (ps, pl)), "w") as f:
                for r in rows:
                    f.write(", ".join(str(x) for x in r) + "\n")
    with open(os.path.join(merged_dir, "union_active_vertices"), "w") as f:
        for v, (label, pss) in m["union_vertices"].items():
            f.write("%d, %d, %s\n" % (v, label, " ".join(str(p) for p in pss)))
    with open(os.path.join(merged_dir, "union_active_edges"), "w") as f:
        for a, b in m["union_edges"]:
            f.write("%d, %d\n" % (a, b))
    return m


def main(argv):
    if len(argv) < 2:
        print(__doc__)
        return 1
    out = argv[1]
    merged = argv[2] if len(argv) > 2 else os.path.join(out, "merged")
    m = write_merged(out, merged)
    for ps, el in m["elements"].items():
        print("pattern [%d]: %d active vertices, %d active edges, subgraphs %s" % (
            ps, len(el["vertices"]), len(el["edges"]), {pl: len(r) for pl, r in el["subgraphs"].items()}))
    print("union: %d vertices, %d edges -> %s" % (len(m["union_vertices"]), len(m["union_edges"]), merged))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
