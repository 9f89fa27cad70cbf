"""Loader for the C-ABI shared library (include/grample_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no
fallback of any kind: if the library is missing, importing the package's compute classes
raises, and without a CUDA device every compute call returns an error.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgrample_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "grample_b200.h")

F64, F32, TABLE, HYBRID, TABLE_BITS = 0, 1, 2, 3, 4
MAX_ABS, MEAN_ABS, HELLINGER, JS = 0, 1, 2, 3
CHAINS_HISTORY = 1
CHAINS_PER_COLOUR = 2
CHAINS_RAO_BLACKWELL = 4
RB_UNIT = 2.0 ** -24  # one Rao-Blackwell bin unit (gb_chains_group_counts under CHAINS_RAO_BLACKWELL)
NEIGHBOR_VAR_MAX = 12
MAX_CARD = 64


class GrampleError(RuntimeError):
    """Raised for every non-zero C-ABI return (mirrors Go's `error` results)."""


_lib = None

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p

_SIGS = {
    "gb_last_error": (C.c_char_p, []),
    "gb_version": (C.c_int, []),
    "gb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "gb_model_create": (C.c_int, [C.c_int32, _i32p, _i32p, C.c_int32, _i32p, _i32p, _i64p, _f64p, C.c_int, C.POINTER(_vp)]),
    "gb_model_load_uai": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(_vp)]),
    "gb_model_destroy": (None, [_vp]),
    "gb_model_hybrid_mask": (C.c_int, [_vp, _i32p]),
    "gb_model_n_vars": (C.c_int, [_vp, _i32p]),
    "gb_model_n_funcs": (C.c_int, [_vp, _i32p]),
    "gb_model_total_card": (C.c_int, [_vp, _i32p]),
    "gb_model_cards": (C.c_int, [_vp, _i32p]),
    "gb_model_fixed": (C.c_int, [_vp, _i32p]),
    "gb_model_collapsed": (C.c_int, [_vp, _i32p]),
    "gb_model_func_arity": (C.c_int, [_vp, C.c_int32, _i32p]),
    "gb_model_func_scope": (C.c_int, [_vp, C.c_int32, _i32p]),
    "gb_model_func_table_size": (C.c_int, [_vp, C.c_int32, _i64p]),
    "gb_model_func_log_table": (C.c_int, [_vp, C.c_int32, _f64p]),
    "gb_model_blanket_size": (C.c_int, [_vp, C.c_int32, _i32p]),
    "gb_model_function_count": (C.c_int, [_vp, C.c_int32, _i32p]),
    "gb_model_schedule": (C.c_int, [_vp, _i32p, _i32p, _i32p, _i32p]),
    "gb_model_table_mode": (C.c_int, [_vp, _i32p, _i64p]),
    "gb_model_bits_mode": (C.c_int, [_vp, _i32p]),
    "gb_model_thresholds": (C.c_int, [_vp, C.c_int32, _i32p, C.POINTER(C.c_uint32)]),
    "gb_model_collapse": (C.c_int, [_vp, C.c_int32, C.c_uint64, _i32p, _f64p, C.POINTER(_vp)]),
    "gb_conditional": (C.c_int, [_vp, C.c_int, C.c_int32, _i32p, _i32p, _f64p]),
    "gb_model_sample": (C.c_int, [_vp, C.c_int, C.c_int32, C.c_int, C.c_uint64, C.c_uint64, _i32p, _i32p]),
    "gb_chains_create": (C.c_int, [C.c_int32, C.POINTER(_vp), _i32p, C.c_uint64, C.c_uint64, C.c_int, C.c_uint32, C.c_int, C.POINTER(_vp)]),
    "gb_chains_add_group": (C.c_int, [_vp, _vp, C.c_int32, C.c_uint64]),
    "gb_chains_destroy": (None, [_vp]),
    "gb_chains_n_groups": (C.c_int, [_vp, _i32p]),
    "gb_chains_n_chains": (C.c_int, [_vp, _i64p]),
    "gb_chains_sweep": (C.c_int, [_vp, C.c_int64, C.c_int]),
    "gb_chains_sweep_timed": (C.c_int, [_vp, C.c_int64, C.c_int, C.POINTER(C.c_float)]),
    "gb_chains_launch_count": (C.c_int, [_vp, _i64p]),
    "gb_chains_scan": (C.c_int, [_vp, C.c_int64, C.c_int]),
    "gb_chains_burnin": (C.c_int, [_vp, C.c_int64]),
    "gb_chains_advance": (C.c_int, [_vp, C.c_int32]),
    "gb_chains_total_samples": (C.c_int, [_vp, _i64p]),
    "gb_chains_group_sweep": (C.c_int, [_vp, C.c_int32, C.c_int64, C.c_int]),
    "gb_chains_group_advance": (C.c_int, [_vp, C.c_int32, C.c_int32]),
    "gb_chains_group_info": (C.c_int, [_vp, C.c_int32, _i32p, _i64p, C.POINTER(_vp)]),
    "gb_chains_synchronize": (C.c_int, [_vp]),
    "gb_chains_merged_marginals": (C.c_int, [_vp, _f64p, _i32p]),
    "gb_chains_merge_begin": (C.c_int, [_vp, _f64p, _i32p]),
    "gb_chains_merge_end": (C.c_int, [_vp, _i64p, _i64p]),
    "gb_chains_merge_timing": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "gb_chains_global_totals": (C.c_int, [_vp, _i64p, _i64p]),
    "gb_comm_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "gb_comm_init_rank": (C.c_int, [C.POINTER(C.c_uint8), C.c_int32, C.c_int32, C.c_int, C.POINTER(_vp)]),
    "gb_comm_info": (C.c_int, [_vp, _i32p, _i32p, C.POINTER(C.c_int)]),
    "gb_comm_destroy": (None, [_vp]),
    "gb_shard": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_uint64), _i32p]),
    "gb_chains_attach_comm": (C.c_int, [_vp, _vp]),
    "gb_fleet_create": (C.c_int, [C.c_int32, C.POINTER(C.c_int), C.POINTER(_vp)]),
    "gb_fleet_destroy": (None, [_vp]),
    "gb_fleet_size": (C.c_int, [_vp, _i32p]),
    "gb_fleet_attach": (C.c_int, [_vp, C.c_int32, _vp]),
    "gb_fleet_sweep": (C.c_int, [_vp, C.c_int64, C.c_int]),
    "gb_fleet_advance": (C.c_int, [_vp, C.c_int32]),
    "gb_fleet_synchronize": (C.c_int, [_vp]),
    "gb_fleet_merged_marginals": (C.c_int, [_vp, _f64p, _i32p]),
    "gb_fleet_merge_begin": (C.c_int, [_vp, _f64p, _i32p]),
    "gb_fleet_merge_end": (C.c_int, [_vp, _i64p, _i64p]),
    "gb_fleet_convergence": (C.c_int, [_vp, C.c_int, _f64p, _f64p]),
    "gb_fleet_adapt": (C.c_int, [_vp, C.POINTER(_vp), C.c_int32, C.c_int32, C.c_int, C.c_int32, C.c_int32, C.c_uint64, _i32p, _i32p]),
    "gb_chains_merge_partial_dev": (C.c_int, [_vp, C.POINTER(_vp), _i64p]),
    "gb_chains_merge_finalize": (C.c_int, [_vp, _f64p, _i32p]),
    "gb_chains_convergence": (C.c_int, [_vp, C.c_int, _f64p, _f64p]),
    "gb_chains_convergence_partial_dev": (C.c_int, [_vp, C.c_int, _f64p, C.POINTER(_vp), _i64p]),
    "gb_convergence_finalize": (C.c_int, [_vp, _f64p, C.c_int32, C.c_int64, _i32p, _f64p]),
    "gb_chains_adapt": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, C.c_int, C.c_int32, C.c_int32, C.c_uint64, _i32p, _i32p]),
    "gb_chains_adapt_scores": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, _f64p, C.c_int64, C.c_int32, C.c_uint64, C.c_uint64,
                                         _i32p, _i32p]),
    "gb_chains_get_state": (C.c_int, [_vp, C.c_int32, _i32p]),
    "gb_chains_set_state": (C.c_int, [_vp, C.c_int32, _i32p]),
    "gb_chains_group_counts": (C.c_int, [_vp, C.c_int32, C.POINTER(C.c_uint64)]),
    "gb_chains_group_history": (C.c_int, [_vp, C.c_int32, C.POINTER(C.c_uint16)]),
    "gb_error_suite": (C.c_int, [C.c_int32, _i32p, _i32p, _f64p, _i32p, _f64p, _f64p]),
    "gb_mar_load": (C.c_int, [C.c_char_p, _i32p, _i32p, _i32p, _f64p]),
}


def exported_symbols():
    """Names of every entry point the binding expects (== what include/grample_b200.h declares)."""
    return sorted(_SIGS)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GrampleError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "grample_b200 has no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise GrampleError(lib().gb_last_error().decode())
