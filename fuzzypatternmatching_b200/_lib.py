"""ctypes binding of the C ABI in include/pmgpu.h.

This is the binding a reference-side maintainer would write for any FFI; the
package itself contains no compute.  If libpmgpu.so is missing the import
fails loudly — there is no CPU fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PMGPU_LIB: another build of the same library (kernel tuning experiments); the default is the in-tree build
LIB_PATH = os.environ.get("PMGPU_LIB") or os.path.join(_HERE, "libpmgpu.so")

PM_NLCC_NEM1, PM_NLCC_TDS = 0, 1
PM_COMM_ID_BYTES = 128


class GraphInfo(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_vertices", "n_local", "n_slots_multi", "n_slots",
                                          "n_slots_padded", "max_degree", "device_bytes")]


class PatternInfo(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("n_vertices", "n_edges", "diameter", "n_constraints")]


class ConstraintInfo(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("walk_length", "valid_cycle", "interleave_lcc", "order_independent")]


class Counts(C.Structure):
    _fields_ = [("n_vertices", C.c_uint64), ("n_edges", C.c_uint64), ("seconds", C.c_double)]


class RunOptions(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("tds_from_pl", "max_iterations", "lcc_only", "keep_subgraphs")]


class RunSummary(C.Structure):
    _fields_ = [("iterations", C.c_uint64), ("search_seconds", C.c_double), ("device_seconds", C.c_double),
                ("n_rows", C.c_uint64), ("n_active_vertices", C.c_uint64), ("n_active_edges", C.c_uint64),
                ("path_count", C.c_uint64), ("edges_processed", C.c_uint64), ("algorithmic_bytes", C.c_uint64)]


class KernelStats(C.Structure):
    _fields_ = [("launches", C.c_uint64), ("ms", C.c_double), ("slots", C.c_uint64), ("vertices", C.c_uint64)]


class Row(C.Structure):
    _fields_ = [("itr", C.c_uint64), ("kind", C.c_int32), ("index", C.c_int32), ("n_vertices", C.c_uint64),
                ("n_edges", C.c_uint64), ("seconds", C.c_double)]


# every symbol include/pmgpu.h declares: name -> (restype, argtypes)
_vp, _u64, _i = C.c_void_p, C.c_uint64, C.c_int
SYMBOLS = {
    "pm_create": (_i, [C.POINTER(_vp), _i]),
    "pm_destroy": (None, [_vp]),
    "pm_last_error": (C.c_char_p, [_vp]),
    "pm_kernel_launches": (_u64, [_vp]),
    "pm_comm_unique_id": (_i, [C.c_char_p]),
    "pm_comm_init": (_i, [_vp, _i, _i, C.c_char_p]),
    "pm_graph_from_slots": (_i, [_vp, _u64, _u64, _vp, _vp]),
    "pm_graph_rmat": (_i, [_vp, _u64, _u64]),
    "pm_graph_from_csr": (_i, [_vp, _u64, _vp, _vp, _vp]),
    "pm_get_kernel_stats": (_i, [_vp, _i, C.POINTER(KernelStats)]),
    "pm_graph_info": (_i, [_vp, C.POINTER(GraphInfo)]),
    "pm_graph_set_delegate_threshold": (_i, [_vp, _u64]),
    "pm_graph_num_delegates": (_i, [_vp, C.POINTER(_u64)]),
    "pm_graph_get_degree": (_i, [_vp, _vp]),
    "pm_graph_get_csr": (_i, [_vp, _vp, _vp]),
    "pm_labels_degree_log2": (_i, [_vp]),
    "pm_labels_set": (_i, [_vp, _vp]),
    "pm_labels_get": (_i, [_vp, _vp]),
    "pm_pattern_load_dir": (_i, [_vp, C.c_char_p]),
    "pm_pattern_info": (_i, [_vp, C.POINTER(PatternInfo)]),
    "pm_pattern_constraint_info": (_i, [_vp, _i, C.POINTER(ConstraintInfo)]),
    "pm_pattern_check_dir": (_i, [C.c_char_p, C.POINTER(PatternInfo), C.POINTER(ConstraintInfo), _i, C.c_char_p, C.c_size_t]),
    "pm_end_iteration": (_i, [_vp, C.c_double]),
    "pm_state_reset": (_i, [_vp]),
    "pm_lcc": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(Counts)]),
    "pm_nlcc": (_i, [_vp, _i, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(Counts)]),
    "pm_run": (_i, [_vp, C.POINTER(RunOptions), C.POINTER(RunSummary)]),
    "pm_run_fuzzy": (_i, [_vp, C.POINTER(RunOptions), C.POINTER(RunSummary)]),
    "pm_get_rows": (_i, [_vp, C.POINTER(Row)]),
    "pm_get_active_vertices": (_i, [_vp, _vp, _vp]),
    "pm_get_active_edges": (_i, [_vp, _vp]),
    "pm_get_subgraph_count": (_i, [_vp, _i, C.POINTER(_u64), C.POINTER(_i)]),
    "pm_get_subgraphs": (_i, [_vp, _i, _vp]),
    "pm_write_results": (_i, [_vp, C.c_char_p]),
    "pm_write_results_ps": (_i, [_vp, C.c_char_p, _i]),
    "pm_labels_from_files": (_i, [_vp, C.c_char_p]),
    "pm_io_read_vertex_data": (_i, [C.c_char_p, _u64, _vp, C.POINTER(_u64), C.c_char_p, C.c_size_t]),
    "pm_io_check_edge_data": (_i, [C.c_char_p, _u64, C.POINTER(_u64), C.c_char_p, C.c_size_t]),
    "pm_io_read_edge_lists": (_i, [C.POINTER(C.c_char_p), _i, _i, C.POINTER(_u64), C.POINTER(_u64), _vp, _vp,
                                   C.c_char_p, C.c_size_t]),
}

_LIB = None


def load():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libpmgpu.so is not built (%s); run `python -c 'import __graft_entry__ as g; "
                              "g.build()'` — there is no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export it
            fn.restype, fn.argtypes = res, args
        _LIB = lib
    return _LIB
