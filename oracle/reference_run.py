"""Runs the reference's own driver (oracle/_ref/run_pattern_matching_beta: src/run_pattern_matching_beta.cpp and the
visitor headers it includes, compiled from /root/reference over the single-rank runtime stand-in of oracle/ref_shim) on an
input and returns its result files in the comparable form the tests use.  TEST INFRASTRUCTURE, like everything under oracle/.

The binary exists only where the reference tree was present when `make -C oracle ref` ran (the build container); it travels to
the GPU box with the repository snapshot.  available() says whether it can be used."""
import os
import shutil
import subprocess
import tempfile

_HERE = os.path.dirname(os.path.abspath(__file__))
BINARY = os.path.join(_HERE, "_ref", "run_pattern_matching_beta")
BINARY_APPROX = os.path.join(_HERE, "_ref", "run_pattern_matching_beta_2")  # the driver of approximate matching
BINARY_FUZZY = os.path.join(_HERE, "_ref", "run_pattern_matching")  # the driver of the run_fuzzy path
BINARY_EDGE_LIST = os.path.join(_HERE, "_ref", "edge_list_dump")  # the reference's own edge list reader
BINARY_RMAT = os.path.join(_HERE, "_ref", "rmat_edge_dump")  # the reference's own R-MAT generator + hash
REFERENCE = "/root/reference"

_TREE = ("all_ranks_active_vertices", "all_ranks_active_vertices_count", "all_ranks_active_edges",
         "all_ranks_active_edges_count", "all_ranks_messages", "all_ranks_subgraphs", "all_ranks_vertex_data")


def build():
    """compiles the binary if the reference tree is here; returns its path or None"""
    if os.path.isdir(os.path.join(REFERENCE, "src")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref", "REFERENCE=" + REFERENCE])
    return BINARY if os.path.exists(BINARY) else None


def available():
    return os.path.exists(BINARY) and os.access(BINARY, os.X_OK)


def approx_available():
    return os.path.exists(BINARY_APPROX) and os.access(BINARY_APPROX, os.X_OK)


def fuzzy_available():
    return os.path.exists(BINARY_FUZZY) and os.access(BINARY_FUZZY, os.X_OK)


def edge_list_dump(files, undirected):
    """What the reference's own parallel_edge_list_reader.hpp iterates over the listed "source target [weight]" files — the
    edges src/ingest_edge_list.cpp hands the graph constructor: (max vertex id, has edge data, [(source, target), ...])."""
    p = subprocess.run([BINARY_EDGE_LIST, "1" if undirected else "0"] + list(files), capture_output=True, text=True, timeout=300)
    if p.returncode != 0:
        raise RuntimeError("edge_list_dump failed: " + p.stderr[-500:])
    lines = [l for l in p.stdout.splitlines() if not l.startswith("Ingesting from")]
    maxv, has_data, n = (int(x) for x in lines[0].split())
    edges = [tuple(int(x) for x in l.split()) for l in lines[1:]]
    assert len(edges) == n
    return maxv, bool(has_data), edges


def rmat_edge_dump(scale, rank, ranks, max_edges=None):
    """The directed pairs the reference's own rmat_edge_generator.hpp yields for generating rank `rank` of `ranks`
    (src/generate_rmat.cpp:202-205): numpy (n, 2) uint64 — every generated edge (u, v) is followed by (v, u)."""
    import numpy as np
    cmd = [BINARY_RMAT, str(scale), str(rank), str(ranks)] + ([str(max_edges)] if max_edges is not None else [])
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if p.returncode != 0:
        raise RuntimeError("rmat_edge_dump failed: " + p.stderr[-500:])
    head, _, body = p.stdout.partition("\n")
    maxv, n = (int(x) for x in head.split())
    pairs = np.array(body.split(), dtype=np.uint64).reshape(-1, 2)
    assert len(pairs) == n and maxv == (1 << scale) - 1
    return pairs


def reference_hash_nbits(values, n):
    """detail::hash_nbits(x, n) of the reference's own include/havoqgt/detail/hash.hpp for every x"""
    p = subprocess.run([BINARY_RMAT, "hash", str(n)] + [str(int(v)) for v in values], capture_output=True, text=True, timeout=60)
    if p.returncode != 0:
        raise RuntimeError("rmat_edge_dump hash failed: " + p.stderr[-500:])
    return [int(x) for x in p.stdout.split()]


def write_slot_file(path, n_vertices, src, dst):
    """the text graph the stand-in's distributed_db reads: vertex count, then one directed slot per line"""
    with open(path, "w") as f:
        f.write("%d\n" % n_vertices)
        if hasattr(src, "dtype"):  # numpy arrays: large graphs
            import numpy as np
            both = np.stack([np.asarray(src, dtype=np.uint64), np.asarray(dst, dtype=np.uint64)], axis=1)
            for at in range(0, len(both), 1 << 22):
                np.savetxt(f, both[at:at + (1 << 22)], fmt="%d")
        else:
            f.write("".join("%d %d\n" % st for st in zip(src, dst)))


def make_result_tree(out, ps=0):
    """the directories the driver expects to exist under its -o argument"""
    for d in _TREE:
        os.makedirs(os.path.join(out, str(ps), d), exist_ok=True)


def launch(graph_path, pattern_dir, out_dir, vertex_data_base=None):
    """starts the driver on a slot file that is already on disk; returns the Popen (stdout captured as text)"""
    make_result_tree(out_dir)
    cmd = [BINARY, "-i", graph_path, "-p", pattern_dir, "-o", out_dir]
    if vertex_data_base:
        cmd += ["-v", vertex_data_base]
    return subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


def search_seconds(stdout):
    """the driver's own clock around one template's do/while (beta.cpp:543, 1352-1357): graph load and label
    construction are outside it"""
    import re
    m = re.search(r"Fuzzy Pattern Matching Time \| Pattern \[0\] : ([0-9.eE+-]+)", stdout)
    if not m:
        raise RuntimeError("the reference driver printed no pattern time")
    return float(m.group(1))


def parsed_pattern(stdout):
    """What the reference driver prints of the template it parsed (beta.cpp:446-468, 770-790):
    ([(vertex, offset, label, degree)], [neighbour list per vertex], diameter, {constraint: {"walk", "args"}})."""
    import re
    verts = [(int(a), int(b), int(c), int(d)) for a, b, c, d in
             re.findall(r"^(\d+) : off-set (\d+) vertex_data (\d+) vertex_degree (\d+)$", stdout, flags=re.M)]
    nbrs = [[int(x) for x in l.split(",") if x.strip()] for l in re.findall(r"^ neighbours : (.*)$", stdout, flags=re.M)]
    diameter = int(re.search(r"^diameter : (\d+)$", stdout, flags=re.M).group(1))
    cons = {}
    for pl, walk in re.findall(r"^Token Passing \[(\d+)\] \| Pattern Vertices : (.*)$", stdout, flags=re.M):
        cons.setdefault(int(pl), {})["walk"] = [int(x) for x in walk.split(",") if x.strip()]
    for pl, args in re.findall(r"^Token Passing \[(\d+)\] \| Arguments : (.*)$", stdout, flags=re.M):
        cons.setdefault(int(pl), {})["args"] = [int(x) for x in args.split()]
    return verts, nbrs, diameter, cons


def template_read_intact(stdout, spec):
    """False when the reference itself mis-read the template: ::graph::generate_vertex_list (graph.hpp:244-270) reads
    edge_list[l] one element PAST THE END on its last round (undefined behaviour); when the stale heap value there happens to
    equal the last template vertex's id, that vertex gains phantom neighbours read from beyond the edge array (seen on 12-vertex
    graphs with the README tree template: "6 : ... vertex_degree 2 / neighbours : 5, 6,").  Such a run searched a different
    template than the files describe and is not compared."""
    _, nbrs, _, _ = parsed_pattern(stdout)
    both = sorted(set((a, b) for a, b in spec["edges"]) | set((b, a) for a, b in spec["edges"]))
    return [(v, u) for v, row in enumerate(nbrs) for u in row] == both


def tds_at_constraint_4(spec):
    """The driver runs template-driven search from constraint index 4 on, whatever the pattern files say
    (beta.cpp:725-730).  A template whose enumeration constraint sits earlier gets copies of its FIRST constraint
    (a check that has already passed: no further pruning, same final sets) inserted before it, so that the reference
    does the same work as an engine told to start template-driven search at the template's own index.
    Returns (spec, index of the enumeration constraint or -1)."""
    cons = list(spec.get("constraints", []))
    at = next((i for i, c in enumerate(cons) if c.get("tds")), -1)
    if at < 0 or at >= 4:
        return spec, at
    pad = [{k: v for k, v in cons[0].items() if k != "tds"} for _ in range(4 - at)]
    return dict(spec, constraints=cons[:at] + pad + cons[at:]), 4


def run(n_vertices, src, dst, pattern_dir, labels=None, workdir=None, timeout=600):
    """pattern_dir: the directory that holds the pattern set (<pattern_dir>/0/pattern_*).  labels: None = the reference's own
    degree labels (vertex_data_db_degree.hpp), else one value per vertex, passed through its -v metadata loader.
    Returns dict(rows, iterations, vertices, edges, subgraphs, stdout) — the same keys as tests.cases.run_summary."""
    own = workdir is None
    work = tempfile.mkdtemp(prefix="pmref_") if own else workdir
    try:
        graph = os.path.join(work, "graph.slots")
        write_slot_file(graph, n_vertices, src, dst)
        out = os.path.join(work, "out")
        for d in _TREE:
            os.makedirs(os.path.join(out, "0", d), exist_ok=True)
        cmd = [BINARY, "-i", graph, "-p", pattern_dir, "-o", out]
        if labels is not None:
            vdir = os.path.join(work, "vertex_data")
            os.makedirs(vdir)
            with open(os.path.join(vdir, "labels_0"), "w") as f:
                f.write("".join("%d %d\n" % (v, int(l)) for v, l in enumerate(labels)))
            cmd += ["-v", os.path.join(vdir, "labels")]
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        if p.returncode != 0:
            raise RuntimeError("reference driver failed (%d): %s" % (p.returncode, p.stderr[-2000:]))
        res = parse_result_tree(out)
        res["stdout"] = p.stdout
        return res
    finally:
        if own:
            shutil.rmtree(work, ignore_errors=True)


def parse_result_tree(out, ps=0, rank=0):
    """result files of one rank (src/run_pattern_matching_beta.cpp:504-535, 1094-1138, 1375-1425) -> comparable view"""
    base = os.path.join(out, str(ps))

    def lines(rel):
        p = os.path.join(base, rel)
        return [l for l in open(p).read().splitlines() if l.strip()] if os.path.exists(p) else []

    # count files: "itr, LP|TP, index, count" per row, vertices and edges in the same row order
    vc = [[t.strip() for t in l.split(",")] for l in lines("all_ranks_active_vertices_count/active_vertices_%d" % rank)]
    ec = [[t.strip() for t in l.split(",")] for l in lines("all_ranks_active_edges_count/active_edges_%d" % rank)]
    assert [r[:3] for r in vc] == [r[:3] for r in ec], "count files disagree on their rows"
    rows = [(int(a[0]), a[1], int(a[2]), int(a[3]), int(b[3])) for a, b in zip(vc, ec)]
    # vertices: "rank, vertex, pattern index, label, bitset"; edges: "rank, vertex, neighbour"
    vertices = sorted((int(t[1]), int(t[4].strip(), 2)) for t in (l.split(",") for l in lines("all_ranks_active_vertices/active_vertices_%d" % rank)))
    edges = sorted((int(t[1]), int(t[2])) for t in (l.split(",") for l in lines("all_ranks_active_edges/active_edges_%d" % rank)))
    subgraphs = {}
    sdir = os.path.join(base, "all_ranks_subgraphs")
    for name in sorted(os.listdir(sdir)) if os.path.isdir(sdir) else []:
        parts = name.split("_")
        if len(parts) == 3 and parts[0] == "subgraphs" and int(parts[2]) == rank:
            # "[rank], v0, v1, ..., v_last, [v_last]" per completed walk (tds_batch_1.hpp:684-693)
            subgraphs[int(parts[1])] = sorted(tuple(int(x) for x in l.replace(",", " ").split() if not x.startswith("["))
                                              for l in lines("all_ranks_subgraphs/" + name))
    itr = [l for l in lines("result_iteration")]
    return dict(rows=rows, iterations=len(itr), vertices=vertices, edges=edges, subgraphs=subgraphs)


def run_fuzzy(n_vertices, src, dst, pattern_dir, labels, timeout=600):
    """The run_fuzzy path: the reference's src/run_pattern_matching.cpp with
    label_propagation_pattern_matching_bsp.hpp and token_passing_pattern_matching.hpp.  Its arguments are positional
    (run_pattern_matching.cpp:62-72: graph, vertex data base name, pattern directory, vertex rank output, backup graph,
    result directory, use-degree flag) and it reads the constraint list from <pattern_dir>/0/pattern — the same format as
    pattern_nlc (pattern_util::read_pattern_list), so that file is copied.  labels: one value per vertex, or None for the
    degree labels.  Returns dict(rows, iterations, vertices): rows (itr, "LP"|"TP", index, vertices, 0) — the driver has no
    edge counts —, vertices sorted (vertex, template vertex index)."""
    work = tempfile.mkdtemp(prefix="pmreff_")
    try:
        graph = os.path.join(work, "graph.slots")
        write_slot_file(graph, n_vertices, src, dst)
        pdir = os.path.join(work, "pattern")
        shutil.copytree(pattern_dir, pdir)
        shutil.copy(os.path.join(pdir, "0", "pattern_nlc"), os.path.join(pdir, "0", "pattern"))
        out = os.path.join(work, "out")
        make_result_tree(out)
        vbase = os.path.join(work, "vertex_data", "labels")
        os.makedirs(os.path.dirname(vbase))
        if labels is not None:
            with open(vbase + "_0", "w") as f:
                f.write("".join("%d %d\n" % (v, int(l)) for v, l in enumerate(labels)))
        cmd = [BINARY_FUZZY, graph, vbase, pdir, os.path.join(work, "vertex_rank"), "", out, "1" if labels is None else "0"]
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        if p.returncode != 0:
            raise RuntimeError("reference driver (run_fuzzy path) failed (%d): %s" % (p.returncode, p.stderr[-2000:]))
        base = os.path.join(out, "0")
        rows = []
        for l in open(os.path.join(base, "all_ranks_active_vertices_count", "active_vertices_0")).read().splitlines():
            t = [x.strip() for x in l.split(",")]
            if len(t) >= 4:
                rows.append((int(t[0]), t[1], int(t[2]), int(t[3]), 0))
        vertices = sorted((int(t[1]), int(t[2])) for t in (l.split(",") for l in
                          open(os.path.join(base, "all_ranks_active_vertices", "active_vertices_0")).read().splitlines() if l.strip()))
        itr = [l for l in open(os.path.join(base, "result_itr")).read().splitlines() if l.strip()]
        return dict(rows=rows, iterations=len(itr), vertices=vertices, stdout=p.stdout)
    finally:
        shutil.rmtree(work, ignore_errors=True)


def run_approx_first_lcc(n_vertices, src, dst, pattern_dir, labels, spec, timeout=600):
    """Approximate matching: the reference's src/run_pattern_matching_beta_2.cpp over
    approximate_pattern_matching/{pattern_graph,local_constraint_checking}.hpp.  Its non-local side (generated constraints,
    tds_batch_5) is not what this repository builds, so only the count rows of the FIRST local constraint checking call are
    returned: [(0, "LP", k, vertices, edges)] for k < diameter.  The driver indexes its constraint list unconditionally
    (beta_2.cpp:524-525), so one constraint in ITS file format (pattern_non_local_constraints: "vertices : enumeration :
    aggregation : cyclic : tds : lcc") is written: a two-edge walk of the template, which only runs after the rows wanted."""
    work = tempfile.mkdtemp(prefix="pmrefa_")
    try:
        graph = os.path.join(work, "graph.slots")
        write_slot_file(graph, n_vertices, src, dst)
        pdir = os.path.join(work, "pattern")
        shutil.copytree(pattern_dir, pdir)
        a, b = spec["edges"][0]
        third = next((x if y == b else y for x, y in spec["edges"][1:] if b in (x, y) and {x, y} != {a, b}), a)
        walk = " ".join(str(v) for v in (a, b, third))
        with open(os.path.join(pdir, "0", "pattern_non_local_constraints"), "w") as f:
            f.write("%s : %s : 0 0 0 : 0 : 0 : 0\n" % (walk, walk))
        out = os.path.join(work, "out")
        make_result_tree(out)
        vbase = os.path.join(work, "vertex_data", "labels")
        os.makedirs(os.path.dirname(vbase))
        with open(vbase + "_0", "w") as f:
            f.write("".join("%d %d\n" % (v, int(l)) for v, l in enumerate(labels)))
        p = subprocess.run([BINARY_APPROX, "-i", graph, "-v", vbase, "-p", pdir, "-o", out], capture_output=True, text=True,
                           timeout=timeout)
        vfile = os.path.join(out, "0", "all_ranks_active_vertices_count", "active_vertices_0")
        efile = os.path.join(out, "0", "all_ranks_active_edges_count", "active_edges_0")
        if not os.path.exists(vfile) or not os.path.getsize(vfile):
            raise RuntimeError("reference driver (approximate matching) wrote no rows (%d): %s" % (p.returncode, p.stderr[-2000:]))
        rows = []
        for lv, le in zip(open(vfile).read().splitlines(), open(efile).read().splitlines()):
            tv, te = [x.strip() for x in lv.split(",")], [x.strip() for x in le.split(",")]
            if tv[1] != "LP" or int(tv[0]) != 0 or (rows and int(tv[2]) <= rows[-1][2]):
                break
            rows.append((0, "LP", int(tv[2]), int(tv[3]), int(te[3])))
        return rows
    finally:
        shutil.rmtree(work, ignore_errors=True)
