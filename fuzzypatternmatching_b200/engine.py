"""Host-side mirror of the reference driver's interface for the pruning path.

The names follow /root/reference/src/run_pattern_matching_beta.cpp: the driver
loads a graph (-i/-b), builds degree labels unless -v is given, reads the
pattern directory (-p) and then alternates
`label_propagation_pattern_matching_bsp` (LCC, beta.cpp:577-583) with
`token_passing_pattern_matching` (NLCC, beta.cpp:879-909) until nothing is
removed, writing the result tree under -o.  Every method forwards to one entry
point of the C ABI (include/pmgpu.h); nothing is computed in Python.
"""
import ctypes as C

import numpy as np

from . import _lib


class PmError(RuntimeError):
    pass


class Engine:
    def __init__(self, device=0):
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = self._lib.pm_create(C.byref(h), device)
        if rc != 0:
            raise PmError("pm_create failed (%d): no usable CUDA device; this engine has no CPU fallback" % rc)
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise PmError("%s (status %d)" % (self._lib.pm_last_error(self._h).decode(), rc))

    # ---- several GPUs: one Engine (process) per GPU, owner(v) = v mod n_ranks ----------
    @staticmethod
    def comm_unique_id():
        """128 opaque bytes; rank 0 creates them and every rank passes the same bytes to comm_init."""
        buf = C.create_string_buffer(_lib.PM_COMM_ID_BYTES)
        rc = _lib.load().pm_comm_unique_id(buf)
        if rc != 0:
            raise PmError("pm_comm_unique_id failed (%d)" % rc)
        return buf.raw

    def comm_init(self, rank, n_ranks, unique_id):
        self.rank, self.n_ranks = rank, n_ranks
        self._chk(self._lib.pm_comm_init(self._h, rank, n_ranks, unique_id))

    # ---- graph (-i / -b) ------------------------------------------------------
    def graph_from_slots(self, n_vertices, src, dst):
        src = np.ascontiguousarray(src, dtype=np.uint32)
        dst = np.ascontiguousarray(dst, dtype=np.uint32)
        assert src.shape == dst.shape
        self._chk(self._lib.pm_graph_from_slots(self._h, n_vertices, src.size, src.ctypes.data, dst.ctypes.data))

    def graph_from_undirected(self, n_vertices, edges):
        e = np.asarray(edges, dtype=np.uint32).reshape(-1, 2)
        src = np.empty(2 * len(e), dtype=np.uint32)
        dst = np.empty(2 * len(e), dtype=np.uint32)
        src[0::2], dst[0::2] = e[:, 0], e[:, 1]
        src[1::2], dst[1::2] = e[:, 1], e[:, 0]
        self.graph_from_slots(n_vertices, src, dst)

    def graph_from_csr(self, rowptr, col, degree_multi, n_vertices=None):
        """Host CSR (distinct neighbours, ascending) + multigraph degrees -> device store.
        Several ranks: the rows of the vertices this rank owns (local row i = vertex i * n_ranks + rank)
        and the global vertex count."""
        rowptr = np.ascontiguousarray(rowptr, dtype=np.uint64)
        col = np.ascontiguousarray(col, dtype=np.uint32)
        degree_multi = np.ascontiguousarray(degree_multi, dtype=np.uint64)
        n = len(rowptr) - 1 if n_vertices is None else int(n_vertices)
        self._chk(self._lib.pm_graph_from_csr(self._h, n, rowptr.ctypes.data, col.ctypes.data,
                                              degree_multi.ctypes.data))

    def graph_rmat(self, scale, gen_ranks):
        """generate_rmat -s <scale> on <gen_ranks> ranks, on the GPU."""
        self._chk(self._lib.pm_graph_rmat(self._h, scale, gen_ranks))

    def graph_set_delegate_threshold(self, threshold):
        """hubs = vertices with multigraph out-degree >= threshold; returns their number (collective)"""
        self._chk(self._lib.pm_graph_set_delegate_threshold(self._h, int(threshold)))
        n = C.c_uint64(0)
        self._chk(self._lib.pm_graph_num_delegates(self._h, C.byref(n)))
        return int(n.value)

    def graph_info(self):
        gi = _lib.GraphInfo()
        self._chk(self._lib.pm_graph_info(self._h, C.byref(gi)))
        return {n: int(getattr(gi, n)) for n, _ in gi._fields_}

    def graph_degree(self):
        out = np.empty(self.graph_info()["n_local"], dtype=np.uint64)
        self._chk(self._lib.pm_graph_get_degree(self._h, out.ctypes.data))
        return out

    def graph_csr(self):
        gi = self.graph_info()
        rowptr = np.empty(gi["n_local"] + 1, dtype=np.uint64)
        col = np.empty(max(gi["n_slots"], 1), dtype=np.uint32)
        self._chk(self._lib.pm_graph_get_csr(self._h, rowptr.ctypes.data, col.ctypes.data))
        return rowptr, col[:gi["n_slots"]]

    # ---- labels (-v) ------------------------------------------------------------
    def labels_degree_log2(self):
        self._chk(self._lib.pm_labels_degree_log2(self._h))

    def labels_set(self, labels):
        labels = np.ascontiguousarray(labels, dtype=np.uint64)
        self._chk(self._lib.pm_labels_set(self._h, labels.ctypes.data))

    def labels_from_files(self, base):
        """-v <base>: "vertex label" files named <base>* (vertex_data_db.hpp:139-262)"""
        self._chk(self._lib.pm_labels_from_files(self._h, base.encode()))

    def labels_get(self):
        out = np.empty(self.graph_info()["n_vertices"], dtype=np.uint64)
        self._chk(self._lib.pm_labels_get(self._h, out.ctypes.data))
        return out

    # ---- pattern (-p) -------------------------------------------------------------
    def pattern_load_dir(self, directory):
        self._chk(self._lib.pm_pattern_load_dir(self._h, directory.encode()))
        pi = _lib.PatternInfo()
        self._chk(self._lib.pm_pattern_info(self._h, C.byref(pi)))
        self.pattern = {n: int(getattr(pi, n)) for n, _ in pi._fields_}
        return self.pattern

    # ---- the two operators of the path ------------------------------------------------
    def state_reset(self):
        self._chk(self._lib.pm_state_reset(self._h))

    def label_propagation_pattern_matching_bsp(self, global_init_step, global_not_finished=False):
        """One LCC call = `diameter` supersteps.  Returns (global_not_finished, [(n_vertices, n_edges, s)])."""
        nf = C.c_int(int(global_not_finished))
        counts = (_lib.Counts * self.pattern["diameter"])()
        self._chk(self._lib.pm_lcc(self._h, int(global_init_step), C.byref(nf), counts))
        return bool(nf.value), [(int(c.n_vertices), int(c.n_edges), float(c.seconds)) for c in counts]

    def token_passing_pattern_matching(self, pl, tds=False):
        """One NLCC constraint incl. the driver's token_source_map post-processing.
        Returns (pattern_found, token_source_deleted, (n_vertices, n_edges, s))."""
        found, deleted, cnt = C.c_int(0), C.c_int(0), _lib.Counts()
        self._chk(self._lib.pm_nlcc(self._h, pl, _lib.PM_NLCC_TDS if tds else _lib.PM_NLCC_NEM1,
                                    C.byref(found), C.byref(deleted), C.byref(cnt)))
        return bool(found.value), bool(deleted.value), (int(cnt.n_vertices), int(cnt.n_edges), float(cnt.seconds))

    # ---- the driver loop ----------------------------------------------------------------
    def run(self, tds_from_pl=4, max_iterations=0, lcc_only=False, keep_subgraphs=True):
        opt = _lib.RunOptions(tds_from_pl, max_iterations, int(lcc_only), int(keep_subgraphs))
        s = _lib.RunSummary()
        self._chk(self._lib.pm_run(self._h, C.byref(opt), C.byref(s)))
        self.summary = {n: getattr(s, n) for n, _ in s._fields_}
        return self.summary

    def run_fuzzy(self, max_iterations=0):
        """The run_fuzzy_pattern_matching path (run_fuzzy_pattern_matching.cpp / run_pattern_matching.cpp):
        unique-label LCC + cycle token passing over the unpruned adjacency."""
        opt = _lib.RunOptions(-1, max_iterations, 0, 0)
        s = _lib.RunSummary()
        self._chk(self._lib.pm_run_fuzzy(self._h, C.byref(opt), C.byref(s)))
        self.summary = {n: getattr(s, n) for n, _ in s._fields_}
        return self.summary

    def rows(self):
        n = int(self.summary["n_rows"])
        buf = (_lib.Row * max(n, 1))()
        self._chk(self._lib.pm_get_rows(self._h, buf))
        return [(int(r.itr), "LP" if r.kind == 0 else "TP", int(r.index), int(r.n_vertices), int(r.n_edges))
                for r in buf[:n]]

    def rows_timed(self):
        """rows() with the device milliseconds of each row appended"""
        return [r + (round(t * 1e3, 3),) for r, t in zip(self.rows(), self.row_seconds())]

    def row_seconds(self):
        n = int(self.summary["n_rows"])
        buf = (_lib.Row * max(n, 1))()
        self._chk(self._lib.pm_get_rows(self._h, buf))
        return [float(r.seconds) for r in buf[:n]]

    def active_vertices(self, n=None):
        n = int(self.summary["n_active_vertices"]) if n is None else n
        v = np.empty(max(n, 1), dtype=np.uint64)
        b = np.empty(max(n, 1), dtype=np.uint16)
        self._chk(self._lib.pm_get_active_vertices(self._h, v.ctypes.data, b.ctypes.data))
        return v[:n], b[:n]

    def active_edges(self, n=None):
        n = int(self.summary["n_active_edges"]) if n is None else n
        p = np.empty(2 * max(n, 1), dtype=np.uint64)
        self._chk(self._lib.pm_get_active_edges(self._h, p.ctypes.data))
        return p[:2 * n].reshape(-1, 2)

    def subgraphs(self, pl):
        cnt, w = C.c_uint64(0), C.c_int(0)
        self._chk(self._lib.pm_get_subgraph_count(self._h, pl, C.byref(cnt), C.byref(w)))
        if cnt.value == 0 or w.value == 0:
            return np.zeros((0, max(w.value, 1)), dtype=np.uint32)
        out = np.empty(cnt.value * w.value, dtype=np.uint32)
        self._chk(self._lib.pm_get_subgraphs(self._h, pl, out.ctypes.data))
        return out.reshape(-1, w.value)

    def subgraph_count(self, pl):
        cnt, w = C.c_uint64(0), C.c_int(0)
        self._chk(self._lib.pm_get_subgraph_count(self._h, pl, C.byref(cnt), C.byref(w)))
        return int(cnt.value)

    def write_results(self, outdir, ps=0):
        self._chk(self._lib.pm_write_results_ps(self._h, outdir.encode(), ps))

    def kernel_stats(self, bin=0):
        ks = _lib.KernelStats()
        self._chk(self._lib.pm_get_kernel_stats(self._h, bin, C.byref(ks)))
        return dict(launches=int(ks.launches), ms=float(ks.ms), slots=int(ks.slots), vertices=int(ks.vertices))

    def kernel_launches(self):
        return int(self._lib.pm_kernel_launches(self._h))


def read_vertex_data(base, n_vertices, labels=None):
    """Host-only -v reader (pm_io_read_vertex_data).  Returns (labels, number of pairs read)."""
    lib = _lib.load()
    out = np.zeros(n_vertices, dtype=np.uint64) if labels is None else np.ascontiguousarray(labels, dtype=np.uint64).copy()
    n = C.c_uint64(0)
    err = C.create_string_buffer(512)
    rc = lib.pm_io_read_vertex_data(base.encode(), n_vertices, out.ctypes.data, C.byref(n), err, len(err))
    if rc != 0:
        raise ValueError(err.value.decode() or "pm_io_read_vertex_data failed (%d)" % rc)
    return out, int(n.value)


def check_edge_data(base, n_vertices):
    """Host-only validation of -e files (pm_io_check_edge_data).  Returns the number of records."""
    lib = _lib.load()
    n = C.c_uint64(0)
    err = C.create_string_buffer(512)
    rc = lib.pm_io_check_edge_data(base.encode(), n_vertices, C.byref(n), err, len(err))
    if rc != 0:
        raise ValueError(err.value.decode() or "pm_io_check_edge_data failed (%d)" % rc)
    return int(n.value)


def read_edge_lists(files, undirected=False):
    """Host-only edge-list reader of ingest_edge_list (pm_io_read_edge_lists).  Returns (n_vertices, src, dst)."""
    lib = _lib.load()
    arr = (C.c_char_p * len(files))(*[f.encode() for f in files])
    nv, ns = C.c_uint64(0), C.c_uint64(0)
    err = C.create_string_buffer(512)
    rc = lib.pm_io_read_edge_lists(arr, len(files), int(undirected), C.byref(nv), C.byref(ns), None, None, err, len(err))
    if rc != 0:
        raise ValueError(err.value.decode() or "pm_io_read_edge_lists failed (%d)" % rc)
    src = np.empty(max(ns.value, 1), dtype=np.uint32)
    dst = np.empty(max(ns.value, 1), dtype=np.uint32)
    rc = lib.pm_io_read_edge_lists(arr, len(files), int(undirected), C.byref(nv), C.byref(ns), src.ctypes.data,
                                   dst.ctypes.data, err, len(err))
    if rc != 0:
        raise ValueError(err.value.decode())
    return int(nv.value), src[:ns.value], dst[:ns.value]


def pattern_check_dir(directory, max_constraints=64):
    """Host-only check of a pattern directory (pm_pattern_check_dir: no context, no GPU).  Returns the pattern's
    sizes and its constraints, raises ValueError with the reader's reason if the directory is rejected."""
    lib = _lib.load()
    info = _lib.PatternInfo()
    cons = (_lib.ConstraintInfo * max_constraints)()
    err = C.create_string_buffer(512)
    rc = lib.pm_pattern_check_dir(directory.encode(), C.byref(info), cons, max_constraints, err, len(err))
    if rc != 0:
        raise ValueError(err.value.decode() or "pm_pattern_check_dir failed (%d)" % rc)
    out = {n: int(getattr(info, n)) for n, _ in info._fields_}
    out["constraints"] = [{n: int(getattr(cons[i], n)) for n, _ in cons[i]._fields_}
                          for i in range(min(out["n_constraints"], max_constraints))]
    return out


def run_pattern_matching_beta(graph, pattern_dir, output_dir=None, labels=None, device=0, **run_kw):
    """The reference driver's flow (-i/-p/-o) in one call.
    graph: ("rmat", scale, gen_ranks) or (n_vertices, src, dst) directed slots."""
    eng = Engine(device)
    if graph[0] == "rmat":
        eng.graph_rmat(graph[1], graph[2])
    else:
        eng.graph_from_slots(*graph)
    if labels is None:
        eng.labels_degree_log2()
    else:
        eng.labels_set(labels)
    eng.pattern_load_dir(pattern_dir)
    eng.run(**run_kw)
    if output_dir is not None:
        eng.write_results(output_dir)
    return eng
