"""Dense-grid TSDF fusion on B200 -- drop-in for the reference's ``grid_fusion`` module.

``TSDFVolume`` keeps the reference's constructor / ``integrate`` / ``get_volume`` surface
(reference grid_fusion.py:19-360) but every voxel is touched by hand-written sm_100a kernels
behind the C ABI in include/tsdf_b200.h (csrc/dense_kernels.cu).  Results follow the reference
CPU path (numba branch, grid_fusion.py:260-314) bit for bit; there is no CPU fallback.

Multi-GPU: one process per GPU.  With an initialised ``torch.distributed`` group (or explicit
``rank`` / ``world_size``) each process owns the brick-thick x-sheets dealt to it in snake order (dist.sheet_owner);
``integrate`` needs no communication and ``get_volume`` gathers over NCCL.
"""
import ctypes as C

import numpy as np

from . import _capi, dist

BRICK = 8


class TSDFVolume:
    """Volumetric TSDF fusion of RGB-D images (reference grid_fusion.py:19)."""

    def __init__(self, vol_bnds, voxel_size, use_gpu=True, rank=None, world_size=None, device=None,
                 group=None):
        """vol_bnds: (3, 2) xyz min/max in metres; voxel_size in metres (grid_fusion.py:22-57).

        ``use_gpu`` is accepted for signature compatibility; this implementation only has the
        CUDA path and raises if it cannot be used."""
        print("Initializing voxel grids ... ")
        vol_bnds = np.asarray(vol_bnds)
        assert vol_bnds.shape == (3, 2), "[!] `vol_bnds` should be of shape (3, 2)."

        self._vol_bnds = vol_bnds
        self._voxel_size = float(voxel_size)
        self._trunc_margin = 5 * self._voxel_size
        self._color_const = 256 * 256
        # grid_fusion.py:41-44, including the in-place snap of the caller's upper bounds
        self._vol_dim = np.ceil((self._vol_bnds[:, 1] - self._vol_bnds[:, 0]) / self._voxel_size).copy(
            order='C').astype(int)
        self._vol_bnds[:, 1] = self._vol_bnds[:, 0] + self._vol_dim * self._voxel_size
        self._vol_origin = self._vol_bnds[:, 0].copy(order='C').astype(np.float32)
        print("Voxel volume size: {} x {} x {} - # points: {:,}".format(
            self._vol_dim[0], self._vol_dim[1], self._vol_dim[2],
            self._vol_dim[0] * self._vol_dim[1] * self._vol_dim[2]))

        self.gpu_mode = 1
        self.rank, self.world_size = dist.rank_world(rank, world_size)
        self.device = dist.local_device(device)
        self._group = group
        self._lib = _capi.lib()
        self._handle = C.c_void_p()
        dim = np.ascontiguousarray(self._vol_dim, dtype=np.int64)
        _capi.check(self._lib.tsdf_grid_create(
            _capi.ptr(dim), _capi.ptr(self._vol_origin), self._voxel_size, self._trunc_margin,
            self.device, self.rank, self.world_size, C.byref(self._handle)))
        n = C.c_int64()
        _capi.check(self._lib.tsdf_grid_local_x(self._handle, C.byref(n), None))
        self._local_x = np.empty(n.value, np.int64)
        _capi.check(self._lib.tsdf_grid_local_x(self._handle, None, _capi.ptr(self._local_x)))
        self._host_cache = None  # (tsdf, weight, color) host volumes, invalidated by integrate
        # per-call scratch for the small matrices: their pointers are taken once (ctypes pointer
        # construction costs more than the copy)
        self._k_buf = np.zeros((3, 3), np.float64)
        self._tinv_buf = np.zeros((4, 4), np.float64)
        self._k_ptr = self._k_buf.ctypes.data
        self._tinv_ptr = self._tinv_buf.ctypes.data

    # -- hot path ------------------------------------------------------------------------------
    def integrate(self, color_im, depth_im, cam_intr, cam_pose, obs_weight=1.):
        """Integrate an RGB-D frame (reference grid_fusion.py:214-314).

        color_im (H,W,3) uint8-valued RGB, depth_im (H,W) metres (0 = invalid), cam_intr (3,3),
        cam_pose (4,4) camera->world."""
        im_h, im_w = depth_im.shape
        rgb = np.ascontiguousarray(color_im, dtype=np.uint8)
        depth = np.ascontiguousarray(depth_im, dtype=np.float64)  # float32 depth upcasts, as in the reference
        self._k_buf[...] = cam_intr
        # the reference inverts the pose on the host in float64 (grid_fusion.py:265); so do we
        self._tinv_buf[...] = np.linalg.inv(np.asarray(cam_pose, dtype=np.float64))
        _capi.check(self._lib.tsdf_grid_integrate(
            self._handle, _capi.ptr(rgb), _capi.ptr(depth), im_h, im_w, self._k_ptr, self._tinv_ptr,
            float(obs_weight)))
        self._host_cache = None

    def integrate_raw(self, color_im, depth_raw, cam_intr, cam_pose, obs_weight=1., depth_scale=1000.,
                      invalid_raw=65535):
        """Frame ingest: ``depth_raw`` is the sensor's uint16 image (millimetres for the bundled
        data); the division by ``depth_scale`` and the ``== invalid`` masking of the reference
        demos (grid_demo1.py:81-83) happen on the device, bit-identically."""
        im_h, im_w = depth_raw.shape
        rgb = np.ascontiguousarray(color_im, dtype=np.uint8)
        raw = np.ascontiguousarray(depth_raw, dtype=np.uint16)
        self._k_buf[...] = cam_intr
        self._tinv_buf[...] = np.linalg.inv(np.asarray(cam_pose, dtype=np.float64))
        _capi.check(self._lib.tsdf_grid_integrate_u16(
            self._handle, _capi.ptr(rgb), _capi.ptr(raw), im_h, im_w, self._k_ptr, self._tinv_ptr,
            float(obs_weight), float(depth_scale), int(invalid_raw)))
        self._host_cache = None

    def integrate_device(self, d_rgb, d_depth, im_h, im_w, cam_intr, tinv, obs_weight=1., producer_stream=None):
        """Frame already resident in HBM (torch uint8 [H,W,3] / float64 [H,W] tensors or raw
        device pointers); ``tinv`` is the host-computed inverse pose.  ``producer_stream`` (a raw
        cudaStream_t) names the stream that is still writing the buffers (see dist.FrameFanout)."""
        self._k_buf[...] = cam_intr
        self._tinv_buf[...] = tinv
        if producer_stream is not None:
            _capi.check(self._lib.tsdf_grid_frame_source(self._handle, producer_stream))
        _capi.check(self._lib.tsdf_grid_integrate_device(
            self._handle, _capi.ptr(d_rgb), _capi.ptr(d_depth), im_h, im_w, self._k_ptr, self._tinv_ptr,
            float(obs_weight)))
        self._host_cache = None

    def integrate_device_raw(self, d_rgb, d_depth_raw, im_h, im_w, cam_intr, tinv, obs_weight=1., depth_scale=1000.,
                             invalid_raw=65535, producer_stream=None):
        """integrate_raw for a frame resident in HBM (uint8 colour, uint16 sensor depth)."""
        self._k_buf[...] = cam_intr
        self._tinv_buf[...] = tinv
        if producer_stream is not None:
            _capi.check(self._lib.tsdf_grid_frame_source(self._handle, producer_stream))
        _capi.check(self._lib.tsdf_grid_integrate_device_u16(
            self._handle, _capi.ptr(d_rgb), _capi.ptr(d_depth_raw), im_h, im_w, self._k_ptr, self._tinv_ptr,
            float(obs_weight), float(depth_scale), int(invalid_raw)))
        self._host_cache = None

    # -- read-out ------------------------------------------------------------------------------
    def _download_local(self):
        shape = (len(self._local_x), int(self._vol_dim[1]), int(self._vol_dim[2]))
        t = np.empty(shape, np.float32)
        w = np.empty(shape, np.float32)
        c = np.empty(shape, np.float32)
        _capi.check(self._lib.tsdf_grid_download(self._handle, _capi.ptr(t), _capi.ptr(w), _capi.ptr(c)))
        return t, w, c

    def _materialise(self):
        if self._host_cache is not None:
            return self._host_cache
        if self.world_size == 1:
            self._host_cache = self._download_local()
            return self._host_cache
        # one NCCL all-gather per plane over NVLink; the sheets interleave back into [X][Y][Z]
        import torch
        dev = torch.device("cuda", self.device)
        shape = (len(self._local_x), int(self._vol_dim[1]), int(self._vol_dim[2]))
        planes = []
        for which in range(3):
            loc = torch.empty(shape, dtype=torch.float32, device=dev)
            args = [None, None, None]
            args[which] = _capi.ptr(loc)
            _capi.check(self._lib.tsdf_grid_download_device(self._handle, *args))
            full = dist.gather_x_sheets(loc, int(self._vol_dim[0]), self.rank, self.world_size, self._group)
            planes.append(full.cpu().numpy())
            del loc, full
        self._host_cache = tuple(planes)
        return self._host_cache

    def get_volume(self):
        """(tsdf, color) float32 [X,Y,Z] host volumes (reference grid_fusion.py:316-320)."""
        t, _, c = self._materialise()
        return t, c

    @property
    def _tsdf_vol_cpu(self):
        return self._materialise()[0]

    @property
    def _weight_vol_cpu(self):
        return self._materialise()[1]

    @property
    def _color_vol_cpu(self):
        return self._materialise()[2]

    def local_volume(self):
        """This rank's sheets only: (tsdf, weight, color, global_x_indices)."""
        return self._download_local() + (self._local_x.copy(),)

    def stats(self):
        c = np.zeros(4, np.int64)
        _capi.check(self._lib.tsdf_grid_stats(self._handle, _capi.ptr(c)))
        return dict(updated_last=int(c[0]), updated_total=int(c[1]), bricks_visited_last=int(c[2]),
                    frames=int(c[3]))

    def stats_post(self):
        """Non-blocking stats: enqueues a snapshot of the counters behind everything issued so far and
        returns a ticket for stats_wait().  Reading frame k-1's ticket after integrating frame k keeps
        the device busy (tsdf_grid_stats_post)."""
        t = C.c_int64()
        _capi.check(self._lib.tsdf_grid_stats_post(self._handle, C.byref(t)))
        return t.value

    def stats_wait(self, ticket):
        """The snapshot of a stats_post() ticket (blocks until that snapshot has landed, not until the device is
        idle).  ``updated_total`` is -1 in a snapshot: later frames may already be adding to the running total;
        ``stats()`` reports it."""
        c = np.zeros(4, np.int64)
        _capi.check(self._lib.tsdf_grid_stats_wait(self._handle, int(ticket), _capi.ptr(c)))
        return dict(updated_last=int(c[0]), updated_total=int(c[1]), bricks_visited_last=int(c[2]),
                    frames=int(c[3]))

    def sync(self):
        _capi.check(self._lib.tsdf_grid_sync(self._handle))

    def set_batch(self, frames):
        """Temporal blocking (tsdf_grid_set_batch): up to ``frames`` (1..4) consecutive integrate
        calls are applied to every brick in one visit -- read once, updated by each frame in call
        order, written once.  Results are bit-identical to frame-by-frame integration; every
        read-out (get_volume, stats, sync, flush) first applies whatever is pending."""
        _capi.check(self._lib.tsdf_grid_set_batch(self._handle, int(frames)))

    def flush(self):
        """Device-side join: the stream of cuda_stream() waits for every kernel issued so far
        (the host does not block).  Record timing events after this."""
        _capi.check(self._lib.tsdf_grid_flush(self._handle))

    def last_kernel_ms(self):
        ms = C.c_float()
        _capi.check(self._lib.tsdf_grid_last_kernel_ms(self._handle, C.byref(ms)))
        return ms.value

    def set_profiling(self, on=True, frame_only=False):
        """Timing events are off by default (they cost API calls per frame).  ``on``: stage
        timing for last_stage_ms() (the frame's integrate kernels then share one stream);
        ``frame_only``: just the whole-frame time for last_kernel_ms(), streams as in production."""
        mode = 0 if not on else (2 if frame_only else 1)
        _capi.check(self._lib.tsdf_grid_set_profiling(self._handle, mode))

    def last_stage_ms(self):
        """Device milliseconds of the last frame's stages: [prepare+cull, integrate] with
        profiling on, [whole frame] otherwise."""
        ms = (C.c_float * 4)()
        n = C.c_int()
        _capi.check(self._lib.tsdf_grid_last_stage_ms(self._handle, ms, C.byref(n)))
        return [ms[i] for i in range(n.value)]

    def cuda_stream(self):
        """Raw cudaStream_t (int) the kernels run on; wrap with torch.cuda.ExternalStream."""
        s = C.c_void_p()
        _capi.check(self._lib.tsdf_grid_stream(self._handle, C.byref(s)))
        return s.value or 0

    def _device_volume(self, dev):
        """(tsdf, color) float32 [X,Y,Z] torch tensors on the device (all ranks' sheets when sharded)."""
        import torch
        shape = (len(self._local_x), int(self._vol_dim[1]), int(self._vol_dim[2]))
        out = []
        for which in (0, 2):
            loc = torch.empty(shape, dtype=torch.float32, device=dev)
            args = [None, None, None]
            args[which] = _capi.ptr(loc)
            _capi.check(self._lib.tsdf_grid_download_device(self._handle, *args))
            out.append(dist.gather_x_sheets(loc, int(self._vol_dim[0]), self.rank, self.world_size, self._group))
        return out[0], out[1]

    def get_point_cloud(self):
        """Surface points with colours, (n, 6) [x y z r g b] (reference grid_fusion.py:322-341)."""
        from . import meshing
        return meshing.point_cloud(self)

    def get_mesh(self):
        """Marching-cubes mesh: verts, faces, norms, colors (reference grid_fusion.py:343-360)."""
        from . import meshing
        return meshing.mesh(self)

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle:
            self._lib.tsdf_grid_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def rigid_transform(xyz, transform):
    """Apply a 4x4 transform to an (N, 3) point array (reference grid_fusion.py:363-368)."""
    homog = np.concatenate([xyz, np.ones((len(xyz), 1), dtype=np.float32)], axis=1)
    return (transform @ homog.T).T[:, :3]


def get_view_frustum(depth_im, cam_intr, cam_pose):
    """World-space corners of the camera frustum out to max(depth) (reference grid_fusion.py:371-383):
    the apex plus the four image corners back-projected at the largest depth; returns (3, 5)."""
    im_h, im_w = depth_im.shape[:2]
    far = np.max(depth_im)
    zs = np.array([0, far, far, far, far])
    corners = np.array([
        (np.array([0, 0, 0, im_w, im_w]) - cam_intr[0, 2]) * zs / cam_intr[0, 0],
        (np.array([0, 0, im_h, 0, im_h]) - cam_intr[1, 2]) * zs / cam_intr[1, 1],
        zs,
    ])
    return rigid_transform(corners.T, cam_pose).T


def _write_ply(filename, header_props, rows_fmt, table, faces=None):
    with open(filename, 'w') as fh:
        fh.write("ply\nformat ascii 1.0\nelement vertex %d\n" % len(table))
        for prop in header_props:
            fh.write("property %s\n" % prop)
        if faces is not None:
            fh.write("element face %d\nproperty list uchar int vertex_index\n" % len(faces))
        fh.write("end_header\n")
        for row in table:
            fh.write(rows_fmt % tuple(row))
        if faces is not None:
            for f in faces:
                fh.write("3 %d %d %d\n" % (f[0], f[1], f[2]))


def meshwrite(filename, verts, faces, norms, colors):
    """ASCII .ply mesh in the reference's column order (reference grid_fusion.py:386-420)."""
    table = [tuple(verts[i]) + tuple(norms[i]) + tuple(int(c) for c in colors[i]) for i in range(len(verts))]
    _write_ply(filename,
               ["float x", "float y", "float z", "float nx", "float ny", "float nz",
                "uchar red", "uchar green", "uchar blue"],
               "%f %f %f %f %f %f %d %d %d\n", table, faces)


def pcwrite(filename, xyzrgb):
    """ASCII .ply point cloud (reference grid_fusion.py:423-446)."""
    rgb = xyzrgb[:, 3:].astype(np.uint8)
    table = [tuple(xyzrgb[i, :3]) + tuple(int(c) for c in rgb[i]) for i in range(len(xyzrgb))]
    _write_ply(filename,
               ["float x", "float y", "float z", "uchar red", "uchar green", "uchar blue"],
               "%f %f %f %d %d %d\n", table)
