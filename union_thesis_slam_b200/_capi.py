"""ctypes binding of libtsdf_b200.so (include/tsdf_b200.h).

There is no CPU fallback: if the CUDA library is missing or cannot be loaded, importing the
product classes raises immediately.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# TSDF_B200_LIB selects another build of the same library (A/B measurements of compile-time variants)
LIB_PATH = os.environ.get("TSDF_B200_LIB") or os.path.join(_HERE, "libtsdf_b200.so")

TSDF_OK = 0
STATUS = {0: "TSDF_OK", -1: "TSDF_ERR_BAD_ARG", -2: "TSDF_ERR_NO_DEVICE", -3: "TSDF_ERR_CUDA",
          -4: "TSDF_ERR_OOM", -5: "TSDF_ERR_HEAP_EXHAUSTED", -6: "TSDF_ERR_TABLE_FULL",
          -7: "TSDF_ERR_CANDIDATES"}

# every symbol include/tsdf_b200.h declares: (name, restype, argtypes)
_p, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
_pp = C.POINTER(C.c_void_p)
SYMBOLS = [
    ("tsdf_last_error", C.c_char_p, []),
    ("tsdf_device_info", _i, [_i, _p, _p, _p, _p]),
    ("tsdf_grid_create", _i, [_p, _p, _d, _d, _i, _i, _i, _pp]),
    ("tsdf_grid_destroy", _i, [_p]),
    ("tsdf_grid_integrate", _i, [_p, _p, _p, _i, _i, _p, _p, _d]),
    ("tsdf_grid_integrate_u16", _i, [_p, _p, _p, _i, _i, _p, _p, _d, _d, _i]),
    ("tsdf_grid_integrate_device", _i, [_p, _p, _p, _i, _i, _p, _p, _d]),
    ("tsdf_grid_integrate_device_u16", _i, [_p, _p, _p, _i, _i, _p, _p, _d, _d, _i]),
    ("tsdf_grid_frame_source", _i, [_p, _p]),
    ("tsdf_grid_download", _i, [_p, _p, _p, _p]),
    ("tsdf_grid_download_device", _i, [_p, _p, _p, _p]),
    ("tsdf_grid_local_x", _i, [_p, _p, _p]),
    ("tsdf_grid_stats", _i, [_p, _p]),
    ("tsdf_grid_stats_post", _i, [_p, _p]),
    ("tsdf_grid_stats_wait", _i, [_p, _i64, _p]),
    ("tsdf_grid_sync", _i, [_p]),
    ("tsdf_grid_last_kernel_ms", _i, [_p, _p]),
    ("tsdf_grid_set_profiling", _i, [_p, _i]),
    ("tsdf_grid_last_stage_ms", _i, [_p, _p, _p]),
    ("tsdf_grid_stream", _i, [_p, _pp]),
    ("tsdf_grid_flush", _i, [_p]),
    ("tsdf_grid_set_batch", _i, [_p, _i]),
    ("vh_create", _i, [_p, _p, _d, _d, _i64, _i64, _i, _i, _i, _pp]),
    ("vh_destroy", _i, [_p]),
    ("vh_integrate", _i, [_p, _p, _p, _i, _i, _p, _p, _d]),
    ("vh_integrate_u16", _i, [_p, _p, _p, _i, _i, _p, _p, _d, _d, _i]),
    ("vh_integrate_device", _i, [_p, _p, _p, _i, _i, _p, _p, _d]),
    ("vh_integrate_device_u16", _i, [_p, _p, _p, _i, _i, _p, _p, _d, _d, _i]),
    ("vh_frame_source", _i, [_p, _p]),
    ("tsdf_fanout_create", _i, [_i, _i, _i, _i64, _i, _pp]),
    ("tsdf_fanout_export", _i, [_p, _p]),
    ("tsdf_fanout_connect", _i, [_p, _p]),
    ("tsdf_fanout_push", _i, [_p, _i64, _p, _i64, _p, _i64, _i64]),
    ("tsdf_fanout_acquire", _i, [_p, _i64, _i, _i64, _i64, _pp, _pp, _pp]),
    ("tsdf_fanout_release", _i, [_p, _i64]),
    ("tsdf_fanout_integrate", _i, [_p, _p, _p, _i64, _i, _p, _p, _i, _i, _i, _p, _p, _d, _d, _i]),
    ("tsdf_fanout_drain", _i, [_p]),
    ("tsdf_fanout_set_timeout", _i, [_p, _d]),
    ("tsdf_fanout_destroy", _i, [_p]),
    ("vh_stats", _i, [_p, _p]),
    ("vh_num_blocks", _i, [_p, _p]),
    ("vh_export_blocks_device", _i, [_p, _i64, _p, _p, _p]),
    ("vh_scatter_blocks_device", _i, [_p, _p, _p, _i64, _i, _p, _p, _p]),
    ("vh_voxel_query", _i, [_p, _p, _i64, _i, _p, _p]),
    ("vh_count_voxels", _i, [_p, _p]),
    ("vh_stats_post", _i, [_p, _p]),
    ("vh_stats_wait", _i, [_p, _i64, _p]),
    ("vh_export_blocks", _i, [_p, _i64, _p, _p, _p]),
    ("vh_to_dense", _i, [_p, _p, _p, _p]),
    ("vh_to_dense_device", _i, [_p, _p, _p, _p]),
    ("vh_remove_blocks", _i, [_p, _p, _i64, _p]),
    ("vh_sync", _i, [_p]),
    ("vh_last_kernel_ms", _i, [_p, _p]),
    ("vh_stream", _i, [_p, _pp]),
    ("vbt_create", _i, [_i64, _i, _d, _i, _pp]),
    ("vbt_destroy", _i, [_p]),
    ("vbt_apply", _i, [_p, _p, _p, _p, _i64, _p]),
    ("vbt_lookup", _i, [_p, _p, _i64, _p, _p]),
    ("vbt_double_table_size", _i, [_p]),
    ("vbt_stats", _i, [_p, _p]),
    ("vbt_dump", _i, [_p, _i64, _p, _p, _p]),
    ("vbt_integrate_frame", _i, [_p, _p, _p, _d, _d, _p, _p, _i, _i, _p, _p, _p]),
    ("vbt_to_dense", _i, [_p, _p, _p, _p]),
    ("tsdf_mesh_extract", _i, [_p, _p, _p, _d, _p, _i, _pp, _p, _p]),
    ("tsdf_mesh_download", _i, [_p, _p, _p, _p, _p]),
    ("tsdf_mesh_destroy", _i, [_p]),
]

_lib = None


class TSDFError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("%s: %s" % (STATUS.get(status, status), message))
        self.status = status


def lib():
    """Loads the CUDA library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libtsdf_b200.so is not built (%s). Run `python -m union_thesis_slam_b200.build`; "
                "there is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        missing = []
        for name, res, args in SYMBOLS:
            try:
                fn = getattr(L, name)
            except AttributeError:
                missing.append(name)
                continue
            fn.restype = res
            fn.argtypes = args
        if missing:  # the header and the library disagree: refuse to run half a product
            raise RuntimeError("libtsdf_b200.so does not export %s; rebuild it" % ", ".join(missing))
        _lib = L
    return _lib


def check(status):
    if status != TSDF_OK:
        raise TSDFError(status, (lib().tsdf_last_error() or b"").decode())


def ptr(a):
    """void* of a numpy array, a torch tensor (data_ptr) or None."""
    if a is None or isinstance(a, int):  # an int is a raw address already
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data  # plain int: accepted for c_void_p parameters, cheaper than data_as()
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError("cannot take a pointer of %r" % type(a))


def device_info(device=0):
    sm, mem, maj, mnr = C.c_int(), C.c_int64(), C.c_int(), C.c_int()
    check(lib().tsdf_device_info(device, C.byref(sm), C.byref(mem), C.byref(maj), C.byref(mnr)))
    return dict(sm_count=sm.value, total_mem=mem.value, cc=(maj.value, mnr.value))
