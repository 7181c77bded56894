"""ctypes wrapper around oracle/libpm_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package
(fuzzypatternmatching_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libpm_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("pm_oracle.cpp", "pm_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


class OrcOptions(C.Structure):
    _fields_ = [("n_ranks", C.c_int), ("tds_from_pl", C.c_int), ("max_iterations", C.c_int),
                ("lcc_only", C.c_int), ("threads", C.c_int), ("keep_subgraphs", C.c_int),
                ("delegate_threshold", C.c_uint64)]


class OrcRow(C.Structure):
    _fields_ = [("itr", C.c_uint64), ("kind", C.c_int32), ("index", C.c_int32),
                ("n_vertices", C.c_uint64), ("n_edges", C.c_uint64), ("seconds", C.c_double)]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
        L.orc_rmat_stream.argtypes = [u64, u64, u64, vp]
        L.orc_hash_nbits.argtypes = [u64, i32]
        L.orc_hash_nbits.restype = u64
        L.orc_graph_from_slots.argtypes = [u64, u64, vp, vp]
        L.orc_graph_from_slots.restype = vp
        L.orc_graph_rmat.argtypes = [u64, u64, i32]
        L.orc_graph_rmat.restype = vp
        L.orc_graph_free.argtypes = [vp]
        for f in ("orc_graph_num_vertices", "orc_graph_num_slots_multi", "orc_graph_num_slots"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = u64
        for f in ("orc_graph_rowptr", "orc_graph_col", "orc_graph_degree"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = vp
        L.orc_labels_degree_log2.argtypes = [vp, vp]
        L.orc_pattern_load.argtypes = [C.c_char_p]
        L.orc_pattern_load.restype = vp
        L.orc_pattern_free.argtypes = [vp]
        L.orc_pattern_error.argtypes = [vp]
        L.orc_pattern_error.restype = C.c_char_p
        for f in ("orc_pattern_num_vertices", "orc_pattern_num_edges", "orc_pattern_diameter",
                  "orc_pattern_num_constraints"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = i32
        L.orc_run_pattern.argtypes = [vp, vp, vp, C.POINTER(OrcOptions)]
        L.orc_run_pattern.restype = vp
        L.orc_run_fuzzy.argtypes = [vp, vp, vp, C.POINTER(OrcOptions)]
        L.orc_run_fuzzy.restype = vp
        L.orc_run_free.argtypes = [vp]
        L.orc_run_error.argtypes = [vp]
        L.orc_run_error.restype = C.c_char_p
        for f in ("orc_run_num_rows", "orc_run_iterations", "orc_run_num_active_edges",
                  "orc_run_cumulative_path_count", "orc_run_edges_processed"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = u64
        L.orc_run_rows.argtypes = [vp]
        L.orc_run_rows.restype = C.POINTER(OrcRow)
        L.orc_run_search_seconds.argtypes = [vp]
        L.orc_run_search_seconds.restype = C.c_double
        for f in ("orc_run_template_vertices", "orc_run_in_map", "orc_run_hazards"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = vp
        L.orc_run_active_edges.argtypes = [vp, vp]
        L.orc_run_num_subgraphs.argtypes = [vp, i32]
        L.orc_run_num_subgraphs.restype = u64
        L.orc_run_subgraph_width.argtypes = [vp, i32]
        L.orc_run_subgraph_width.restype = i32
        L.orc_run_subgraphs.argtypes = [vp, i32]
        L.orc_run_subgraphs.restype = vp
        L.orc_run_write_results.argtypes = [vp, vp, vp, vp, C.c_char_p]
        L.orc_run_write_results.restype = i32
        _LIB = L
    return _LIB


def _view(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()


def rmat_stream(scale, rank, n_edges):
    out = np.empty(2 * n_edges, dtype=np.uint64)
    lib().orc_rmat_stream(scale, rank, n_edges, out.ctypes.data)
    return out.reshape(-1, 2)


def hash_nbits(x, n):
    return int(lib().orc_hash_nbits(int(x), int(n)))


class Graph:
    def __init__(self, handle):
        self.h = handle
        L = lib()
        self.V = int(L.orc_graph_num_vertices(handle))
        self.n_slots_multi = int(L.orc_graph_num_slots_multi(handle))
        self.n_slots = int(L.orc_graph_num_slots(handle))

    @classmethod
    def from_slots(cls, n_vertices, src, dst):
        src = np.ascontiguousarray(src, dtype=np.uint64)
        dst = np.ascontiguousarray(dst, dtype=np.uint64)
        return cls(lib().orc_graph_from_slots(n_vertices, len(src), src.ctypes.data, dst.ctypes.data))

    @classmethod
    def from_undirected(cls, n_vertices, edges):
        """edges: iterable of generated (u,v); each yields slots (u,v) and (v,u)
        like the reference's undirected edge iterator."""
        e = np.asarray(list(edges), dtype=np.uint64).reshape(-1, 2)
        src = np.empty(2 * len(e), dtype=np.uint64)
        dst = np.empty(2 * len(e), dtype=np.uint64)
        src[0::2], dst[0::2] = e[:, 0], e[:, 1]
        src[1::2], dst[1::2] = e[:, 1], e[:, 0]
        return cls.from_slots(n_vertices, src, dst)

    @classmethod
    def rmat(cls, scale, gen_ranks, threads=0):
        return cls(lib().orc_graph_rmat(scale, gen_ranks, threads))

    @property
    def rowptr(self):
        return _view(lib().orc_graph_rowptr(self.h), self.V + 1, np.uint64)

    @property
    def col(self):
        return _view(lib().orc_graph_col(self.h), self.n_slots, np.uint32)

    @property
    def degree(self):
        return _view(lib().orc_graph_degree(self.h), self.V, np.uint64)

    def labels_degree_log2(self):
        out = np.empty(self.V, dtype=np.uint64)
        lib().orc_labels_degree_log2(self.h, out.ctypes.data)
        return out

    def __del__(self):
        try:
            lib().orc_graph_free(self.h)
        except Exception:
            pass


class Pattern:
    def __init__(self, directory):
        L = lib()
        self.dir = directory
        self.h = L.orc_pattern_load(directory.encode())
        err = L.orc_pattern_error(self.h)
        if err:
            raise ValueError(err.decode())
        self.n_vertices = L.orc_pattern_num_vertices(self.h)
        self.n_edges = L.orc_pattern_num_edges(self.h)
        self.diameter = L.orc_pattern_diameter(self.h)
        self.n_constraints = L.orc_pattern_num_constraints(self.h)

    def __del__(self):
        try:
            lib().orc_pattern_free(self.h)
        except Exception:
            pass


class Run:
    """Result of the reference outer loop (beta.cpp:544-1351) on the CPU oracle."""

    def __init__(self, graph, labels, pattern, n_ranks=1, tds_from_pl=4, max_iterations=0,
                 lcc_only=False, threads=0, keep_subgraphs=True, delegate_threshold=0, fuzzy=False):
        """fuzzy=True: the run_fuzzy_pattern_matching path (unique-label LCC + cycle token passing over
        the unpruned adjacency) instead of run_pattern_matching_beta's."""
        L = lib()
        self.graph, self.pattern = graph, pattern
        self.labels = np.ascontiguousarray(labels, dtype=np.uint64)
        opt = OrcOptions(n_ranks, tds_from_pl, max_iterations, int(lcc_only), threads,
                         int(keep_subgraphs), delegate_threshold)
        self.h = (L.orc_run_fuzzy if fuzzy else L.orc_run_pattern)(graph.h, self.labels.ctypes.data, pattern.h, C.byref(opt))
        err = L.orc_run_error(self.h)
        if err:
            raise ValueError(err.decode())
        n = int(L.orc_run_num_rows(self.h))
        rows = L.orc_run_rows(self.h)
        self.rows = [(int(rows[i].itr), "LP" if rows[i].kind == 0 else "TP", int(rows[i].index),
                      int(rows[i].n_vertices), int(rows[i].n_edges)) for i in range(n)]
        self.row_seconds = [float(rows[i].seconds) for i in range(n)]
        self.iterations = int(L.orc_run_iterations(self.h))
        self.search_seconds = float(L.orc_run_search_seconds(self.h))
        V = graph.V
        self.template_vertices = _view(L.orc_run_template_vertices(self.h), V, np.uint16)
        self.in_map = _view(L.orc_run_in_map(self.h), V, np.uint8)
        ne = int(L.orc_run_num_active_edges(self.h))
        pairs = np.empty(2 * ne, dtype=np.uint64)
        if ne:
            L.orc_run_active_edges(self.h, pairs.ctypes.data)
        self.active_edges = pairs.reshape(-1, 2)
        self.hazards = _view(L.orc_run_hazards(self.h), 8, np.uint64)
        self.path_count = int(L.orc_run_cumulative_path_count(self.h))
        self.edges_processed = int(L.orc_run_edges_processed(self.h))
        self.subgraphs = []
        for pl in range(pattern.n_constraints):
            w = L.orc_run_subgraph_width(self.h, pl)
            k = int(L.orc_run_num_subgraphs(self.h, pl))
            if w and k:
                self.subgraphs.append(_view(L.orc_run_subgraphs(self.h, pl), k * w, np.uint32).reshape(k, w))
            else:
                self.subgraphs.append(np.zeros((0, max(w, 1)), dtype=np.uint32))

    def active_vertices(self):
        """sorted (vertex, T_arr) of the final vertex_state_map"""
        idx = np.nonzero(self.in_map)[0]
        return idx.astype(np.uint64), self.template_vertices[idx]

    def write_results(self, outdir):
        rc = lib().orc_run_write_results(self.h, self.graph.h, self.labels.ctypes.data,
                                         self.pattern.h, outdir.encode())
        if rc != 0:
            raise IOError("result tree under %s is incomplete (the reference never mkdirs)" % outdir)

    def __del__(self):
        try:
            lib().orc_run_free(self.h)
        except Exception:
            pass


def make_result_tree(outdir, ps=0):
    """Creates the directory skeleton the reference expects to pre-exist
    (examples/results/, beta.cpp:504-535)."""
    for d in ("all_ranks_active_vertices", "all_ranks_active_vertices_count", "all_ranks_active_edges",
              "all_ranks_active_edges_count", "all_ranks_messages", "all_ranks_subgraphs",
              "all_ranks_vertex_data"):
        os.makedirs(os.path.join(outdir, str(ps), d), exist_ok=True)
