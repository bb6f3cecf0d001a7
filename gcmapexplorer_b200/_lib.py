"""ctypes binding of libgcxnorm.so (include/gcx_norm.h).

There is no CPU fallback: if the CUDA library is missing or no B200 is visible, importing the library
handle or creating a context raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C gcmapexplorer_b200/csrc``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GCX_LIB") or os.path.join(_HERE, "libgcxnorm.so")
NCLASS = 8
KERNEL_CLASSES = ("matvec", "apply_sym", "vc_apply", "row_stats", "diag_reduce", "diag_apply", "vector_step", "other")


class GcxError(RuntimeError):
    pass


_lib = None

_c_f32p = C.POINTER(C.c_float)
_c_f64p = C.POINTER(C.c_double)
_c_i32p = C.POINTER(C.c_int32)
_c_i64p = C.POINTER(C.c_int64)
_c_u8p = C.POINTER(C.c_uint8)
_ctx = C.c_void_p

_PROTOS = {
    "gcx_last_error": (C.c_char_p, []),
    "gcx_version": (C.c_int, []),
    "gcx_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "gcx_ctx_create": (C.c_int, [C.c_int, C.POINTER(_ctx)]),
    "gcx_ctx_destroy": (C.c_int, [_ctx]),
    "gcx_sync": (C.c_int, [_ctx]),
    "gcx_device_info": (C.c_int, [_ctx, C.POINTER(C.c_int), _c_i64p, _c_i64p]),
    "gcx_host_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p)]),
    "gcx_host_free": (C.c_int, [C.c_void_p]),
    "gcx_map_create": (C.c_int, [_ctx, C.c_int64, C.c_int64, C.c_int64]),
    "gcx_map_upload": (C.c_int, [_ctx, C.c_void_p, C.c_int64]),
    "gcx_map_download": (C.c_int, [_ctx, C.c_int, C.c_void_p, C.c_int64]),
    "gcx_map_upload_rows": (C.c_int, [_ctx, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int]),
    "gcx_map_upload_finish": (C.c_int, [_ctx]),
    "gcx_map_upload_file": (C.c_int, [_ctx, C.c_char_p, C.c_int64, C.c_int]),
    "gcx_map_download_file": (C.c_int, [_ctx, C.c_int, C.c_char_p, C.c_int64]),
    "gcx_map_download_async": (C.c_int, [_ctx, C.c_int, C.c_void_p, C.c_int64]),
    "gcx_download_wait": (C.c_int, [_ctx]),
    "gcx_map_synth": (C.c_int, [_ctx, C.c_uint64, C.c_double, C.c_double, C.c_double]),
    "gcx_map_clamp": (C.c_int, [_ctx, C.c_int, C.c_float, C.c_int, C.c_float]),
    "gcx_map_scale": (C.c_int, [_ctx, C.c_float]),
    "gcx_map_swap": (C.c_int, [_ctx]),
    "gcx_map_checksum": (C.c_int, [_ctx, C.c_int, C.POINTER(C.c_uint64)]),
    "gcx_row_stats": (C.c_int, [_ctx, _c_f64p, _c_i32p]),
    "gcx_row_negative_flags": (C.c_int, [_ctx, _c_i32p]),
    "gcx_map_adopt_device": (C.c_int, [_ctx, C.c_void_p, C.c_int64]),
    "gcx_map_export_device": (C.c_int, [_ctx, C.c_int, C.c_void_p, C.c_int64]),
    "gcx_set_mask": (C.c_int, [_ctx, _c_u8p]),
    "gcx_ic_run": (C.c_int, [_ctx, C.c_double, C.c_int, _c_f64p, C.POINTER(C.c_int), _c_f64p, C.POINTER(C.c_int)]),
    "gcx_kr_run": (C.c_int, [_ctx, C.c_double, C.c_double, C.c_double, _c_f64p, C.POINTER(C.c_int), C.POINTER(C.c_int), _c_f64p]),
    "gcx_apply_sym_scale": (C.c_int, [_ctx, _c_f64p, C.c_int, C.c_int, _c_f32p, _c_f32p]),
    "gcx_apply_norm_vectors": (C.c_int, [_ctx, _c_f64p, _c_f64p, C.c_int, _c_f32p, _c_f32p]),
    "gcx_apply_sym_vc": (C.c_int, [_ctx, _c_u8p, C.c_int, C.c_int, _c_f32p, _c_f32p, _c_f32p, _c_f32p]),
    "gcx_vc_run": (C.c_int, [_ctx, C.c_int, C.c_int, _c_f32p, _c_f32p, _c_f32p]),
    "gcx_diag_stats": (C.c_int, [_ctx, C.c_int, _c_f64p]),
    "gcx_diag_apply": (C.c_int, [_ctx, _c_f64p, C.c_int, C.c_int, _c_f32p, _c_f32p]),
    "gcx_mcfs_batch": (C.c_int, [C.POINTER(_ctx), C.c_int, C.c_int, C.c_int, _c_f32p, _c_f32p]),
    "gcx_diag_stats_pooled": (C.c_int, [C.POINTER(_ctx), C.c_int, C.c_int, _c_f64p, C.c_int64]),
    "gcx_diag_avg_read": (C.c_int, [_ctx, _c_f64p]),
    "gcx_compact_f64": (C.c_int, [_ctx, _c_i32p, C.c_int64, _c_f64p]),
    "gcx_downsample": (C.c_int, [_ctx, C.c_int, C.c_int, _c_f32p, C.c_int64, _c_f32p, _c_f32p]),
    "gcx_transition_run": (C.c_int, [_ctx, C.c_int, _c_f32p, _c_f32p]),
    "gcx_corr_run": (C.c_int, [_ctx, C.c_int, C.c_double, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float]),
    "gcx_corr_matrix_f64": (C.c_int, [_ctx, _c_f64p, C.c_int64, C.c_int, C.c_double, C.c_int, _c_f64p]),
    "gcx_debug_i8gemm": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "gcx_corr_last_route": (C.c_int, [_ctx, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gcx_stationary_run": (C.c_int, [_ctx, C.c_double, C.c_int, _c_f64p, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "gcx_set_option": (C.c_int, [_ctx, C.c_char_p, C.c_double]),
    "gcx_partition_rows": (C.c_int, [C.c_int64, C.c_int, C.c_int, _c_i64p, _c_i64p]),
    "gcx_dist_unique_id": (C.c_int, [C.c_void_p]),
    "gcx_dist_init": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_void_p]),
    "gcx_profile": (C.c_int, [_ctx, C.c_int]),
    "gcx_profile_read": (C.c_int, [_ctx, C.c_int, _c_i64p, _c_f64p]),
    "gcx_timer_start": (C.c_int, [_ctx]),
    "gcx_timer_stop": (C.c_int, [_ctx, _c_f64p]),
    "gcx_map_invalidate": (C.c_int, [_ctx]),
    "gcx_debug_sym_cycles": (C.c_int, [_ctx, _c_i64p, C.c_int, C.POINTER(C.c_int)]),
    "gcx_debug_sym_stamps": (C.c_int, [_ctx, _c_i64p, C.c_int, C.POINTER(C.c_int)]),
    "gcx_debug_sym_plan": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int, _c_i64p, C.POINTER(C.c_int32), C.c_int, _c_i64p]),
    "gcx_pass_bytes": (C.c_int, [_ctx, _c_i64p, C.POINTER(C.c_int)]),
    "gcx_last_run_ms": (C.c_int, [_ctx, _c_f64p]),
    "gcx_last_run_products": (C.c_int, [_ctx, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gcx_launch_count": (C.c_int, [_ctx, _c_i64p]),
    "gcx_bench_matvec": (C.c_int, [_ctx, C.c_int, C.c_int, _c_f64p]),
}

EXPORTS = tuple(_PROTOS)


def lib():
    """The loaded library (loads on first use; raises GcxError if it is not built)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise GcxError(f"{LIB_PATH} is not built (run __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise GcxError(lib().gcx_last_error().decode("utf-8", "replace"))


def device_count():
    n = C.c_int(0)
    check(lib().gcx_device_count(C.byref(n)))
    return n.value
