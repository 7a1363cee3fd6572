"""Loader of librhj.so, the C ABI declared in include/rhj.h.

The CUDA library is the product: there is no Python or CPU fallback.  If the shared object is
missing, or no sm_100 device is usable, importing callers get a RuntimeError -- never a silent
slow path.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RHJ_LIB") or os.path.join(_HERE, "librhj.so")  # RHJ_LIB: tuning builds only

c_u64 = ctypes.c_uint64
c_u64p = ctypes.POINTER(ctypes.c_uint64)
c_vp = ctypes.c_void_p


class PlanInfo(ctypes.Structure):
    """struct rhj_plan_info (include/rhj.h)."""
    _fields_ = [("bits_total", ctypes.c_uint32), ("bits_pass1", ctypes.c_uint32), ("bits_pass2", ctypes.c_uint32),
                ("build_is_S", ctypes.c_uint32), ("n_partitions", ctypes.c_uint32), ("n_items", ctypes.c_uint32),
                ("kernel_launches", ctypes.c_uint32), ("optimistic_pass1", ctypes.c_uint32)]


class ShardPlan(ctypes.Structure):
    """struct rhj_shard_plan (include/rhj.h)."""
    _fields_ = [("world", ctypes.c_uint32), ("rank_bits", ctypes.c_uint32), ("bits_total", ctypes.c_uint32),
                ("bits_pass1", ctypes.c_uint32), ("bits_pass2", ctypes.c_uint32), ("build_is_S", ctypes.c_uint32)]


class PipeCfg(ctypes.Structure):
    """struct rhj_pipe_cfg (include/rhj.h)."""
    _fields_ = [("world", ctypes.c_uint32), ("rank", ctypes.c_uint32), ("chunks", ctypes.c_uint32),
                ("ship_ctas", ctypes.c_uint32), ("wire_bytes", ctypes.c_uint32), ("reserved0", ctypes.c_uint32),
                ("nR_local_max", ctypes.c_uint64), ("nS_local_max", ctypes.c_uint64),
                ("sym", ctypes.c_void_p * 16)]


class QRelation(ctypes.Structure):
    _fields_ = [("columns", ctypes.POINTER(ctypes.c_void_p)), ("num_tuples", ctypes.c_uint64), ("num_columns", ctypes.c_uint64)]


class QFilter(ctypes.Structure):
    _fields_ = [("binding", ctypes.c_uint32), ("column", ctypes.c_uint32), ("op", ctypes.c_int32), ("reserved0", ctypes.c_uint32),
                ("constant", ctypes.c_uint64)]


class QJoin(ctypes.Structure):
    _fields_ = [("binding1", ctypes.c_uint32), ("column1", ctypes.c_uint32), ("binding2", ctypes.c_uint32),
                ("column2", ctypes.c_uint32)]


class QProj(ctypes.Structure):
    _fields_ = [("binding", ctypes.c_uint32), ("column", ctypes.c_uint32)]


class QueryDesc(ctypes.Structure):
    """struct rhj_query_desc (include/rhj.h)."""
    _fields_ = [("n_bindings", ctypes.c_uint32), ("n_filters", ctypes.c_uint32), ("n_joins", ctypes.c_uint32),
                ("n_projs", ctypes.c_uint32), ("reorder_joins", ctypes.c_uint32), ("reserved0", ctypes.c_uint32),
                ("bindings", ctypes.POINTER(QRelation)), ("filters", ctypes.POINTER(QFilter)),
                ("joins", ctypes.POINTER(QJoin)), ("projs", ctypes.POINTER(QProj))]


class QueryStats(ctypes.Structure):
    """struct rhj_query_stats (include/rhj.h)."""
    _fields_ = [(n, ctypes.c_uint64) for n in ("h2d_bytes", "d2h_bytes", "kernel_launches", "joins", "join_input_tuples",
                                               "join_output_pairs", "result_rows", "joins_reordered")]


# name -> (restype, argtypes): every symbol include/rhj.h declares
SIGNATURES = {
    "rhj_version": (ctypes.c_char_p, []),
    "rhj_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(c_vp)]),
    "rhj_destroy": (ctypes.c_int, [c_vp]),
    "rhj_last_error": (ctypes.c_char_p, [c_vp]),
    "rhj_reserve": (ctypes.c_int, [c_vp, c_u64, c_u64]),
    "rhj_workspace_bytes": (c_u64, [c_vp]),
    "rhj_join_device": (ctypes.c_int, [c_vp, c_vp, c_u64, c_vp, c_u64, c_vp, c_u64, c_u64p, ctypes.c_int, c_vp]),
    "rhj_join_count_device": (ctypes.c_int, [c_vp, c_vp, c_u64, c_vp, c_u64, c_u64p, c_vp]),
    "rhj_join_write_device": (ctypes.c_int, [c_vp, c_vp, c_u64, c_vp]),
    "rhj_join_host": (ctypes.c_int, [c_vp, c_vp, c_u64, c_vp, c_u64, ctypes.POINTER(c_vp), c_u64p]),
    "rhj_pairs_to_pages": (c_vp, [c_vp, c_u64, c_u64p]),
    "rhj_histogram_device": (ctypes.c_int, [c_vp, c_vp, c_u64, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp]),
    "rhj_partition_device": (ctypes.c_int, [c_vp, c_vp, c_u64, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp, c_vp]),
    "rhj_filter_u64_device": (ctypes.c_int, [c_vp, c_vp, c_vp, c_u64, ctypes.c_int, c_u64, c_vp, c_u64p, c_vp]),
    "rhj_gather_tuples_device": (ctypes.c_int, [c_vp, c_vp, c_vp, c_u64, c_vp, c_vp]),
    "rhj_gather_sum_u64_device": (ctypes.c_int, [c_vp, c_vp, c_vp, c_u64, c_u64p, c_vp]),
    "rhj_pairs_digest_device": (ctypes.c_int, [c_vp, c_vp, c_u64, c_u64p, c_u64p, c_vp]),
    "rhj_intermediate_expand_host": (ctypes.c_int, [c_vp, c_vp, c_u64, c_vp, c_u64, ctypes.c_int, c_vp, ctypes.c_uint32,
                                                    c_vp, c_u64p]),
    "rhj_intermediate_filter_host": (ctypes.c_int, [c_vp, c_vp, c_vp, c_u64, c_vp, c_u64, c_vp, ctypes.c_uint32, c_vp,
                                                    c_u64p]),
    "rhj_column_device": (ctypes.c_int, [c_vp, c_vp, c_u64, ctypes.POINTER(c_vp), c_u64p]),
    "rhj_column_cache_clear": (ctypes.c_int, []),
    "rhj_unique_rowids_device": (ctypes.c_int, [c_vp, c_vp, c_u64, c_u64, c_vp, c_u64p, c_vp]),
    "rhj_query_execute": (ctypes.c_int, [c_vp, ctypes.POINTER(QueryDesc), c_u64p, ctypes.POINTER(ctypes.c_int),
                                          ctypes.POINTER(QueryStats)]),
    "rhj_shuffle_partition_device": (ctypes.c_int, [c_vp, c_vp, c_u64, ctypes.c_int, c_vp, c_u64p, c_vp]),
    "rhj_shard_plan_make": (ctypes.c_int, [c_u64, c_u64, ctypes.c_int, ctypes.POINTER(ShardPlan)]),
    "rhj_shardx_begin": (ctypes.c_int, [c_vp, ctypes.POINTER(ShardPlan), c_vp]),
    "rhj_shardx_pass1_device": (ctypes.c_int, [c_vp, ctypes.POINTER(ShardPlan), ctypes.c_int, c_vp, c_u64, c_vp, c_vp, c_vp]),
    "rhj_shardx_layout_device": (ctypes.c_int, [c_vp, ctypes.POINTER(ShardPlan), ctypes.c_int, ctypes.c_int, c_vp, c_u64p,
                                                c_u64p, c_u64p, c_u64p, c_u64p, c_vp]),
    "rhj_shardx_pass2_device": (ctypes.c_int, [c_vp, ctypes.POINTER(ShardPlan), ctypes.c_int, c_vp, c_u64, c_vp]),
    "rhj_shardx_join_device": (ctypes.c_int, [c_vp, ctypes.POINTER(ShardPlan), c_vp, c_u64, c_u64p, c_vp]),
    "rhj_pipe_sym_bytes": (c_u64, [ctypes.POINTER(ShardPlan), ctypes.POINTER(PipeCfg)]),
    "rhj_pipe_open": (ctypes.c_int, [c_vp, ctypes.POINTER(ShardPlan), ctypes.POINTER(PipeCfg)]),
    "rhj_pipe_begin": (ctypes.c_int, [c_vp, c_u64, c_vp]),
    "rhj_pipe_pass1_device": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, c_vp, c_u64, c_vp]),
    "rhj_pipe_ship_device": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, c_vp]),
    "rhj_pipe_pass2_device": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, c_vp]),
    "rhj_pipe_post_device": (ctypes.c_int, [c_vp, c_vp]),
    "rhj_pipe_join_device": (ctypes.c_int, [c_vp, c_vp, c_u64, c_u64p, ctypes.POINTER(ctypes.c_uint32), c_vp]),
    "rhj_last_plan": (ctypes.c_int, [c_vp, ctypes.POINTER(PlanInfo)]),
    "rhj_set_profiling": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "rhj_last_phase_ms": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_float)]),
}

_lib = None


def load():
    """dlopen librhj.so and bind every declared entry point (no GPU needed for this)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C radixhashjoin_b200/csrc` -- there is no CPU fallback")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
