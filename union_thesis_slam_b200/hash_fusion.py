"""Voxel-block hash map TSDF fusion on B200 -- drop-in for the reference's ``hash_fusion`` module
on the integrate / get_volume path.

The reference ``HashTable`` (hash_fusion.py:29-507) is a Python list of 5-slot buckets holding one
entry per *voxel*.  Here the map stores 8^3-voxel blocks in HBM, found through a lock-free
open-addressing table keyed on packed 64-bit block coordinates (same three primes as
hash_fusion.py:24-26), with block storage from an atomic free-list heap; per frame a ray walk
emits candidate blocks, new keys are inserted with atomicCAS, visible blocks are compacted and
each is integrated by one CTA (csrc/hash_kernels.cu).  A block exists iff one of its voxels was
updated, so the exported block set equals {v // 8 : v is an entry of the reference map}.

Voxel arithmetic is the dense rule (float32 state, reference grid_fusion.py:207-211,302-314):
dense and hash outputs are then identical, and differ from the reference hash map's float64
running mean (voxel.py:32-37) by ~1e-7 in tsdf only.  ``obs_weight`` is honoured (the reference
silently drops it, hash_fusion.py:141,145; pass 1.0 to compare).
"""
import ctypes as C

import numpy as np

from . import _capi, dist

P1 = 73856093
P2 = 19349669
P3 = 83492791
BLOCK = 8


class HashTable:
    """The data structure for voxel storage and retrieval: hash tables (reference hash_fusion.py:29)."""

    def __init__(self, vol_bounds, voxel_size, map_size=1000000, load_factor=0.75, use_gpu=False,
                 heap_blocks=None, rank=None, world_size=None, device=None, group=None, per_voxel=False):
        """``map_size`` is the number of table slots (reference: number of buckets,
        hash_fusion.py:36).  ``heap_blocks`` (extension) bounds the number of 8^3 blocks this rank
        can hold; default: every block of the bounded volume, capped at map_size.

        ``per_voxel=True`` (extension) selects the reference's own data structure for ``integrate`` /
        ``get_volume`` as well: one entry per voxel in the 5-slot-bucket container with float64
        running means, inserted in the reference's order, so the occupancy metrics (entries,
        non-empty buckets, collisions, resizes) and the volumes equal the reference's exactly.  It
        runs on the device but sequentially (the layout is order-dependent): a parity and
        comparison mode, orders of magnitude slower than the default block map."""
        vol_bounds = np.asarray(vol_bounds)
        assert vol_bounds.shape == (3, 2), "[!] `vol_bounds` should be of shape (3, 2)."
        self._vol_bounds = vol_bounds
        self._voxel_size = voxel_size
        self._trunc_margin = 5 * self._voxel_size
        self._color_const = 256 * 256
        self._load_factor = 0.75  # the reference ignores its argument too (hash_fusion.py:38)
        self._bucket_size = 1
        # hash_fusion.py:50-57, including the in-place snap of the caller's upper bounds
        self._vol_dim = np.ceil((self._vol_bounds[:, 1] - self._vol_bounds[:, 0]) / self._voxel_size).copy(
            order='C').astype(int)
        self._vol_bounds[:, 1] = self._vol_bounds[:, 0] + self._vol_dim * self._voxel_size
        self._vol_origin = self._vol_bounds[:, 0].copy(order='C').astype(np.float32)
        print("Voxel volume size: {} x {} x {} - # points: {:,}".format(
            self._vol_dim[0], self._vol_dim[1], self._vol_dim[2],
            self._vol_dim[0] * self._vol_dim[1] * self._vol_dim[2]))

        self.rank, self.world_size = dist.rank_world(rank, world_size)
        self.device = dist.local_device(device)
        self._group = group
        nblocks_total = int(np.prod((self._vol_dim + BLOCK - 1) // BLOCK))
        if heap_blocks is None:
            heap_blocks = min(nblocks_total, int(map_size))
        self._lib = _capi.lib()
        self._handle = C.c_void_p()
        dim = np.ascontiguousarray(self._vol_dim, dtype=np.int64)
        _capi.check(self._lib.vh_create(
            _capi.ptr(dim), _capi.ptr(self._vol_origin), float(self._voxel_size), float(self._trunc_margin),
            int(map_size), int(heap_blocks), self.device, self.rank, self.world_size, C.byref(self._handle)))
        self._table_size = self.stats()["table_slots"]
        self._map_size = int(map_size)
        # per-call scratch for the small matrices: their pointers are taken once (ctypes pointer
        # construction costs more than the copy)
        self._k_buf = np.zeros((3, 3), np.float64)
        self._tinv_buf = np.zeros((4, 4), np.float64)
        self._k_ptr = self._k_buf.ctypes.data
        self._tinv_ptr = self._tinv_buf.ctypes.data
        self._voxels = None  # per-voxel container with the reference's bucket layout, created on first use
        self._per_voxel = bool(per_voxel)
        if self._per_voxel and self.world_size != 1:
            raise ValueError("per_voxel=True is a single-process mode")

    # -- hot path ------------------------------------------------------------------------------
    def integrate(self, color_im, depth_im, cam_intr, cam_pose, obs_weight=1.):
        """Integrate an RGB-D image into the hash map (reference hash_fusion.py:103-145)."""
        if self._per_voxel:  # the reference's per-voxel loop; obs_weight is dropped like hash_fusion.py:141,145
            self.voxel_table.integrate_frame(self._vol_dim, self._vol_origin, self._voxel_size, self._trunc_margin,
                                             color_im, depth_im, cam_intr, cam_pose)
            return
        im_h, im_w = depth_im.shape
        rgb = np.ascontiguousarray(color_im, dtype=np.uint8)
        depth = np.ascontiguousarray(depth_im, dtype=np.float64)
        self._k_buf[...] = cam_intr
        self._tinv_buf[...] = np.linalg.inv(np.asarray(cam_pose, dtype=np.float64))  # hash_fusion.py:112
        _capi.check(self._lib.vh_integrate(
            self._handle, _capi.ptr(rgb), _capi.ptr(depth), im_h, im_w, self._k_ptr, self._tinv_ptr,
            float(obs_weight)))

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
        _capi.check(self._lib.vh_integrate_u16(
            self._handle, _capi.ptr(rgb), _capi.ptr(raw), im_h, im_w, self._k_ptr, self._tinv_ptr,
            float(obs_weight), float(depth_scale), int(invalid_raw)))

    def integrate_device(self, d_rgb, d_depth, im_h, im_w, cam_intr, tinv, obs_weight=1., producer_stream=None):
        """Frame resident in HBM; ``producer_stream``: see grid_fusion.TSDFVolume.integrate_device."""
        self._k_buf[...] = cam_intr
        self._tinv_buf[...] = tinv
        if producer_stream is not None:
            _capi.check(self._lib.vh_frame_source(self._handle, producer_stream))
        _capi.check(self._lib.vh_integrate_device(
            self._handle, _capi.ptr(d_rgb), _capi.ptr(d_depth), im_h, im_w, self._k_ptr, self._tinv_ptr,
            float(obs_weight)))

    def integrate_device_raw(self, d_rgb, d_depth_raw, im_h, im_w, cam_intr, tinv, obs_weight=1., depth_scale=1000.,
                             invalid_raw=65535, producer_stream=None):
        self._k_buf[...] = cam_intr
        self._tinv_buf[...] = tinv
        if producer_stream is not None:
            _capi.check(self._lib.vh_frame_source(self._handle, producer_stream))
        _capi.check(self._lib.vh_integrate_device_u16(
            self._handle, _capi.ptr(d_rgb), _capi.ptr(d_depth_raw), im_h, im_w, self._k_ptr, self._tinv_ptr,
            float(obs_weight), float(depth_scale), int(invalid_raw)))

    # -- read-out ------------------------------------------------------------------------------
    def get_volume_with_weight(self):
        """(tsdf, weight, color) dense float32 [X,Y,Z]; with world_size > 1 the union over ranks."""
        shape = tuple(int(d) for d in self._vol_dim)
        if self.world_size == 1:
            t = np.empty(shape, np.float32)
            w = np.empty(shape, np.float32)
            c = np.empty(shape, np.float32)
            _capi.check(self._lib.vh_to_dense(self._handle, _capi.ptr(t), _capi.ptr(w), _capi.ptr(c)))
            return t, w, c
        import torch
        dev = torch.device("cuda", self.device)
        t, w, c = self._gathered_device_volume(dev, want_weight=True)
        return t.cpu().numpy(), w.cpu().numpy(), c.cpu().numpy()

    def export_blocks_device(self, dev):
        """This rank's blocks as device tensors: keys [n,3] int32, payload [n, 3*512] float32."""
        import torch
        n = C.c_int64()
        _capi.check(self._lib.vh_export_blocks_device(self._handle, 0, C.byref(n), None, None))
        keys = torch.empty((max(n.value, 1), 3), dtype=torch.int32, device=dev)
        payload = torch.empty((max(n.value, 1), 3 * BLOCK ** 3), dtype=torch.float32, device=dev)
        if n.value:
            _capi.check(self._lib.vh_export_blocks_device(self._handle, n.value, C.byref(n), _capi.ptr(keys),
                                                          _capi.ptr(payload)))
        return keys[:n.value], payload[:n.value]

    def _gathered_device_volume(self, dev, want_weight=False):
        """Read-out at world > 1: exchange BLOCK LISTS (dist.gather_block_lists) and scatter the union into dense device
        volumes on every rank -- bit-exact by construction, blocks are disjoint across ranks."""
        import torch
        keys, payload = self.export_blocks_device(dev)
        keys, payload = dist.gather_block_lists(keys, payload, self._group)
        shape = tuple(int(d) for d in self._vol_dim)
        t = torch.empty(shape, dtype=torch.float32, device=dev)
        c = torch.empty(shape, dtype=torch.float32, device=dev)
        w = torch.empty(shape, dtype=torch.float32, device=dev) if want_weight else None
        _capi.check(self._lib.vh_scatter_blocks_device(self._handle, _capi.ptr(keys), _capi.ptr(payload), keys.shape[0], 1,
                                                       _capi.ptr(t), _capi.ptr(w), _capi.ptr(c)))
        return t, w, c

    def get_volume(self):
        """tsdf (default 1) and colour (default 0) volumes (reference hash_fusion.py:442-463)."""
        if self._per_voxel:
            return self.voxel_table.to_dense(self._vol_dim)
        t, _, c = self.get_volume_with_weight()
        return t, c

    def export_blocks(self):
        """This rank's blocks: keys (n,3) int32 (voxel // 8), payload (n,3,8,8,8) float32
        [tsdf|weight|colour][x][y][z]."""
        n = C.c_int64()
        _capi.check(self._lib.vh_export_blocks(self._handle, 0, C.byref(n), None, None))
        keys = np.empty((n.value, 3), np.int32)
        payload = np.empty((n.value, 3, BLOCK, BLOCK, BLOCK), np.float32)
        _capi.check(self._lib.vh_export_blocks(self._handle, n.value, C.byref(n), _capi.ptr(keys),
                                               _capi.ptr(payload)))
        return keys[:n.value], payload[:n.value]

    def remove_blocks(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.int32).reshape(-1, 3)
        removed = C.c_int64()
        _capi.check(self._lib.vh_remove_blocks(self._handle, _capi.ptr(keys), len(keys), C.byref(removed)))
        return removed.value

    # -- per-voxel API (reference hash_fusion.py:192-440) -------------------------------------------
    # The reference keeps one entry per voxel in 5-slot buckets with offset chains, fills it from
    # integrate() and exposes it directly (its tests/hash_map_test.py drives add / get / remove by hand).
    # Here integrate() stores 8^3 blocks in the lock-free block table, and the per-voxel calls see ONE
    # map made of two stores:
    #   * voxels written by integrate(): answered from the block table (vh_voxel_query) -- a voxel has an
    #     entry iff its weight is > 0, exactly when the reference would have created one;
    #   * entries added by hand (add_hash_entry / add_voxel): the reference-layout bucket container
    #     (bucket_table.VoxelBucketTable, csrc/bucket_kernels.cu), created by the first such call.
    # Look-ups never create the container, so they do not change what the occupancy metrics report.
    @property
    def voxel_table(self):
        if self._voxels is None:
            from . import bucket_table
            self._voxels = bucket_table.VoxelBucketTable(self._map_size, 5, self._load_factor, self.device)
        return self._voxels

    def _block_query(self, coords, remove=False):
        """(found bool[n], values float32[n,3] = sdf, weight, colour) of voxel coordinates in the block map."""
        q = np.ascontiguousarray(np.asarray(coords).reshape(-1, 3), dtype=np.int32)
        out = np.empty((len(q), 3), np.float32)
        found = np.zeros(len(q), np.uint8)
        _capi.check(self._lib.vh_voxel_query(self._handle, _capi.ptr(q), len(q), 1 if remove else 0,
                                              _capi.ptr(out), _capi.ptr(found)))
        return found.astype(bool), out

    def query_voxels(self, coords):
        """Batch form of get_voxel for voxels written by integrate(): (found, sdf, weight, colour) arrays."""
        found, out = self._block_query(coords)
        return found, out[:, 0], out[:, 1], out[:, 2]

    def add_voxel(self, voxel, world_coord):
        self.voxel_table.add_voxel(voxel, world_coord)

    def add_hash_entry(self, hash_entry):
        return self.voxel_table.add_hash_entry(hash_entry)

    def get_hash_entry(self, world_coord):
        """The entry of a voxel coordinate or None (reference hash_fusion.py:285-310).  Entries are
        snapshots: the map itself lives on the device."""
        if self._voxels is not None or self._per_voxel:
            e = self.voxel_table.get_hash_entry(world_coord)
            if e is not None or self._per_voxel:
                return e
        found, out = self._block_query(world_coord)
        if not found[0]:
            return None
        from . import bucket_table
        pos = [int(v) for v in np.asarray(world_coord).reshape(-1)[:3]]
        return bucket_table.HashEntry(pos, None, bucket_table.FrozenVoxel(float(out[0, 0]), float(out[0, 2]), float(out[0, 1])))

    def get_voxel(self, world_coord):
        """The voxel of a coordinate, or None when it has no entry (the reference raises AttributeError
        on None here, hash_fusion.py:312-320; returning None is the usable form of the same answer)."""
        e = self.get_hash_entry(world_coord)
        return None if e is None else e.get_voxel()

    def remove(self, world_coord):
        """Removes the voxel's entry (reference hash_fusion.py:322-329): a hand-added container entry if
        there is one, else the voxel written by integrate() goes back to the no-entry state."""
        if self._voxels is not None or self._per_voxel:
            if self._per_voxel or self.voxel_table.get_hash_entry(world_coord) is not None:
                self.voxel_table.remove(world_coord)
                return
        self._block_query(world_coord, remove=True)

    def remove_hash_entry(self, hash_entry):
        return self.voxel_table.remove_hash_entry(hash_entry)

    def double_table_size(self):
        self.voxel_table.double_table_size()

    def get_num_collisions(self):
        return self.voxel_table.get_num_collisions() if (self._voxels is not None or self._per_voxel) else 0

    def count_voxels(self):
        """Voxels with an entry among those written by integrate() (weight > 0), counted on the device; with
        world_size > 1 the sum over the ranks' disjoint block sets."""
        n = C.c_int64()
        _capi.check(self._lib.vh_count_voxels(self._handle, C.byref(n)))
        if self.world_size > 1:
            import torch
            import torch.distributed as tdist
            t = torch.tensor([n.value], dtype=torch.int64, device=torch.device("cuda", self.device))
            tdist.all_reduce(t, group=self._group)
            return int(t.item())
        return n.value

    # -- occupancy metrics (reference hash_fusion.py:147-180) ------------------------------------
    _STAT_NAMES = ["blocks", "table_keys", "table_slots", "heap_capacity", "heap_free", "updated_last",
                   "updated_total", "blocks_visited_last", "max_probe", "keys_inserted_last", "frames",
                   "candidates_last"]

    def stats(self):
        s = np.zeros(12, np.int64)
        _capi.check(self._lib.vh_stats(self._handle, _capi.ptr(s)))
        return {k: int(v) for k, v in zip(self._STAT_NAMES, s)}

    def stats_post(self):
        """Non-blocking stats (vh_stats_post): returns a ticket for stats_wait()."""
        t = C.c_int64()
        _capi.check(self._lib.vh_stats_post(self._handle, C.byref(t)))
        return t.value

    def stats_wait(self, ticket):
        s = np.zeros(12, np.int64)
        _capi.check(self._lib.vh_stats_wait(self._handle, int(ticket), _capi.ptr(s)))
        return {k: int(v) for k, v in zip(self._STAT_NAMES, s)}

    def count_num_hash_entries(self):
        """Number of voxel entries, the reference's unit (hash_fusion.py:147-154): voxels written by
        integrate() that have weight > 0 plus entries added by hand.  ``stats()["blocks"]`` is the
        number of 8^3 blocks that hold them."""
        if self._per_voxel:
            return self.voxel_table.count_num_hash_entries()
        n = self.count_voxels()
        if self._voxels is not None:
            n += self._voxels.count_num_hash_entries()
        return n

    def get_num_non_empty_bucket(self):
        if self._voxels is not None:
            return self._voxels.get_num_non_empty_bucket()
        return self.stats()["table_keys"]

    get_num_non_empty_buckets = get_num_non_empty_bucket

    def get_load_factor(self):
        if self._voxels is not None:
            return self._voxels.get_load_factor()
        s = self.stats()
        return s["table_keys"] / s["table_slots"]

    def needs_resize(self):
        return self.get_load_factor() >= self._load_factor

    def hash_function(self, world_coord):
        """H(x,y,z) = (x*p1 xor y*p2 xor z*p3) mod n (reference hash_fusion.py:182-190); n is the
        bucket count of the per-voxel container once that is in use, else the block table's slots."""
        if self._voxels is not None:
            return self._voxels.hash_function(world_coord)
        x, y, z = (int(v) for v in world_coord[:3])
        return ((x * P1) ^ (y * P2) ^ (z * P3)) % self._table_size

    def sync(self):
        _capi.check(self._lib.vh_sync(self._handle))

    def last_kernel_ms(self):
        ms = C.c_float()
        _capi.check(self._lib.vh_last_kernel_ms(self._handle, C.byref(ms)))
        return ms.value

    def cuda_stream(self):
        """Raw cudaStream_t (int) of the integrate kernels; wrap with torch.cuda.ExternalStream."""
        s = C.c_void_p()
        _capi.check(self._lib.vh_stream(self._handle, C.byref(s)))
        return s.value or 0

    def _device_volume(self, dev):
        """(tsdf, color) float32 [X,Y,Z] torch tensors on the device (union over ranks when sharded)."""
        import torch
        if self._per_voxel:
            t, c = self.voxel_table.to_dense(self._vol_dim)
            return torch.from_numpy(t).to(dev), torch.from_numpy(c).to(dev)
        shape = tuple(int(d) for d in self._vol_dim)
        t = torch.empty(shape, dtype=torch.float32, device=dev)
        c = torch.empty(shape, dtype=torch.float32, device=dev)
        if self.world_size == 1:
            _capi.check(self._lib.vh_to_dense_device(self._handle, _capi.ptr(t), None, _capi.ptr(c)))
            return t, c
        t, _, c = self._gathered_device_volume(dev)
        return t, c

    def get_mesh(self):
        """Marching-cubes mesh of the scattered volume (reference hash_fusion.py:465-485)."""
        from . import meshing
        return meshing.mesh(self)

    def get_point_cloud(self):
        """Surface points with colours (reference hash_fusion.py:487-507)."""
        from . import meshing
        return meshing.point_cloud(self)

    def close(self):
        if getattr(self, "_voxels", None) is not None:
            self._voxels.close()
            self._voxels = None
        if getattr(self, "_handle", None) is not None and self._handle:
            self._lib.vh_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
