"""CPU oracle for the TSDF integration hot path -- TEST INFRASTRUCTURE ONLY.

Thin ctypes layer over ``oracle/tsdf_oracle.c`` (the C/OpenMP restatement of
/root/reference/grid_fusion.py:170-212,228-232,260-314,363-368 and
hash_fusion.py:103-145 + data_structures/voxel.py:19-49).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; the product package never does.

Parity status: pinned against the unmodified reference by tests/test_oracle_golden.py
(fixtures produced by tests/golden/make_golden.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc, OpenMP)."""
    src = os.path.join(_HERE, "tsdf_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        i64, dbl, i32 = C.c_int64, C.c_double, C.c_int
        p = C.c_void_p
        L.oracle_fold_color.argtypes = [p, p, i64]
        L.oracle_fold_color.restype = None
        L.oracle_dense_integrate.argtypes = [p, p, p, p, i64, i64, i64, p, dbl, dbl,
                                             p, p, i32, i32, p, p, dbl, i32]
        L.oracle_dense_integrate.restype = i64
        L.oracle_hashvox_integrate.argtypes = [p, p, p, p, p, i64, i64, i64, p, dbl, dbl,
                                               p, p, i32, i32, p, p, i32, p]
        L.oracle_hashvox_integrate.restype = i64
        L.oracle_hash_function.argtypes = [i64, i64, i64, i64]
        L.oracle_hash_function.restype = i64
        L.oracle_max_threads.restype = i32
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def fold_color(color_im):
    """grid_fusion.py:228-232."""
    rgb = np.ascontiguousarray(color_im, dtype=np.uint8)
    out = np.empty(rgb.shape[:2], np.float32)
    lib().oracle_fold_color(_ptr(rgb), _ptr(out), out.size)
    return out


def intrinsics_f64(cam_intr):
    """grid_fusion.py:190-192: intrinsics pass through float32 before use."""
    k = np.asarray(cam_intr).astype(np.float32)
    return np.array([k[0, 0], k[1, 1], k[0, 2], k[1, 2]], dtype=np.float64)


def volume_geometry(vol_bnds, voxel_size):
    """grid_fusion.py:33-44 (same as hash_fusion.py:42-53).  Returns (dim int64[3], origin f32[3]);
    like the reference it snaps vol_bnds[:,1] in the caller's array."""
    vol_bnds = np.asarray(vol_bnds)
    assert vol_bnds.shape == (3, 2), "[!] `vol_bnds` should be of shape (3, 2)."
    dim = np.ceil((vol_bnds[:, 1] - vol_bnds[:, 0]) / voxel_size).copy(order="C").astype(int)
    vol_bnds[:, 1] = vol_bnds[:, 0] + dim * voxel_size
    origin = vol_bnds[:, 0].copy(order="C").astype(np.float32)
    return dim.astype(np.int64), origin


class _Base:
    def __init__(self, vol_bnds, voxel_size, xs=None, nthreads=0):
        self._vol_bnds = np.asarray(vol_bnds)
        self._voxel_size = float(voxel_size)
        self._trunc_margin = 5 * self._voxel_size
        self._vol_dim, self._vol_origin = volume_geometry(self._vol_bnds, self._voxel_size)
        self.xs = (np.arange(self._vol_dim[0], dtype=np.int64) if xs is None
                   else np.ascontiguousarray(xs, dtype=np.int64))
        self.local_shape = (len(self.xs), int(self._vol_dim[1]), int(self._vol_dim[2]))
        self.nthreads = int(nthreads)
        self.updates_per_frame = []

    def _frame(self, color_im, depth_im, cam_intr, cam_pose):
        cfold = fold_color(color_im)
        depth = np.ascontiguousarray(depth_im, dtype=np.float64)  # f32 input upcasts, :278-279
        intr = intrinsics_f64(cam_intr)
        tinv = np.ascontiguousarray(np.linalg.inv(np.asarray(cam_pose, dtype=np.float64)))  # :265
        return cfold, depth, intr, tinv


class OracleTSDFVolume(_Base):
    """Dense grid oracle with the reference call surface (TSDFVolume CPU branch).

    ``xs`` restricts the oracle to a subset of x-sheets (slabs / cyclic sheets); the local
    planes then have shape [len(xs), Y, Z].
    """

    def __init__(self, vol_bnds, voxel_size, xs=None, nthreads=0):
        super().__init__(vol_bnds, voxel_size, xs, nthreads)
        self._tsdf_vol_cpu = np.ones(self.local_shape, np.float32)
        self._weight_vol_cpu = np.zeros(self.local_shape, np.float32)
        self._color_vol_cpu = np.zeros(self.local_shape, np.float32)

    def integrate(self, color_im, depth_im, cam_intr, cam_pose, obs_weight=1.):
        cfold, depth, intr, tinv = self._frame(color_im, depth_im, cam_intr, cam_pose)
        h, w = depth.shape
        n = lib().oracle_dense_integrate(
            _ptr(self._tsdf_vol_cpu), _ptr(self._weight_vol_cpu), _ptr(self._color_vol_cpu),
            _ptr(self.xs), len(self.xs), self.local_shape[1], self.local_shape[2],
            _ptr(self._vol_origin), self._voxel_size, self._trunc_margin,
            _ptr(cfold), _ptr(depth), h, w, _ptr(intr), _ptr(tinv),
            float(obs_weight), self.nthreads)
        self.updates_per_frame.append(int(n))
        return int(n)

    def get_volume(self):
        return self._tsdf_vol_cpu, self._color_vol_cpu


class OracleHashVox(_Base):
    """Reference HashTable viewed as a set of per-voxel float64 entries.

    ``entries`` (bool volume) is the entry set; ``get_volume`` follows hash_fusion.py:442-463.
    ``obs_weight`` is accepted and ignored, exactly like hash_fusion.py:141,145.
    """

    def __init__(self, vol_bnds, voxel_size, xs=None, nthreads=0):
        super().__init__(vol_bnds, voxel_size, xs, nthreads)
        self.sdf = np.ones(self.local_shape, np.float64)
        self.weight = np.zeros(self.local_shape, np.float64)
        self.color = np.zeros(self.local_shape, np.float64)
        self.entries = np.zeros(self.local_shape, np.uint8)
        self.created_per_frame = []

    def integrate(self, color_im, depth_im, cam_intr, cam_pose, obs_weight=1.):
        cfold, depth, intr, tinv = self._frame(color_im, depth_im, cam_intr, cam_pose)
        h, w = depth.shape
        made = C.c_int64(0)
        n = lib().oracle_hashvox_integrate(
            _ptr(self.sdf), _ptr(self.weight), _ptr(self.color), _ptr(self.entries),
            _ptr(self.xs), len(self.xs), self.local_shape[1], self.local_shape[2],
            _ptr(self._vol_origin), self._voxel_size, self._trunc_margin,
            _ptr(cfold), _ptr(depth), h, w, _ptr(intr), _ptr(tinv),
            self.nthreads, C.byref(made))
        self.updates_per_frame.append(int(n))
        self.created_per_frame.append(int(made.value))
        return int(n)

    def get_volume(self):
        tsdf = np.ones(self.local_shape, np.float32)
        color = np.zeros(self.local_shape, np.float32)
        m = self.entries.astype(bool)
        tsdf[m] = self.sdf[m]
        color[m] = self.color[m]
        return tsdf, color

    def block_set(self, block=8):
        """{voxel // block : voxel has an entry} as a sorted (n,3) int64 array (global x)."""
        idx = np.argwhere(self.entries)
        if idx.size == 0:
            return np.zeros((0, 3), np.int64)
        idx[:, 0] = self.xs[idx[:, 0]]
        return np.unique(idx // block, axis=0)


def hash_function(coord, table_size):
    """hash_fusion.py:182-190."""
    return int(lib().oracle_hash_function(int(coord[0]), int(coord[1]), int(coord[2]), int(table_size)))


def block_set_from_weight(weight, xs=None, block=8):
    """Blocks (voxel//block) holding at least one voxel with w>0."""
    idx = np.argwhere(weight > 0)
    if idx.size == 0:
        return np.zeros((0, 3), np.int64)
    if xs is not None:
        idx[:, 0] = np.asarray(xs)[idx[:, 0]]
    return np.unique(idx // block, axis=0)
