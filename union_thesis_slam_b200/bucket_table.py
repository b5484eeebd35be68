"""Per-voxel hash container with the reference's slot layout, resident on the device.

Mirrors the container half of the reference ``HashTable`` (hash_fusion.py:147-180, 192-440) and
its value types ``HashEntry`` (data_structures/hash_entry.py) and ``Voxel``
(data_structures/voxel.py): buckets of five slots, the last one heading the bucket's collision
chain, chains threaded through ``(bucket, slot)`` offsets into other buckets' free slots,
resize by re-insertion at 0.75 non-empty buckets.  The tables live in HBM and every operation
runs in csrc/bucket_kernels.cu behind the ``vbt_*`` C ABI; mutating operations execute strictly
in call order on the device because the reference's layout is order-dependent (where an entry
lands depends on everything inserted before it).  Return values -- the ``(bucket, slot)`` pairs
of ``add_hash_entry``, ``1`` / ``0`` of ``remove_hash_entry`` -- are the reference's.

``get_hash_entry`` returns a snapshot (position, offset, voxel) of the stored entry, not a live
object: mutate through the table (``integrate_voxels``, ``remove``), not through the snapshot.
There is no CPU fallback: without the CUDA library and a device the constructor raises.
"""
import ctypes as C
import math

import numpy as np

from . import _capi

OP_ADD, OP_GET, OP_REMOVE, OP_INTEGRATE = 0, 1, 2, 3
P1 = 73856093
P2 = 19349669
P3 = 83492791


class Voxel:
    """Value object of data_structures/voxel.py:6-17 (sdf = 1, colour = 0, weight = 0)."""

    def __init__(self, sdf=1, color=0, weight=0):
        self._sdf = sdf
        self._color = color
        self._weight = weight

    def get_sdf(self):
        return self._sdf

    def get_color(self):
        return self._color

    def get_weight(self):
        return self._weight


class HashEntry:
    """Value object of data_structures/hash_entry.py:6-45: (position, offset, voxel)."""

    def __init__(self, position, offset, voxel):
        self._position = position
        self._offset = offset
        self._voxel = voxel

    def get_position(self):
        return self._position

    def get_voxel(self):
        return self._voxel

    def set_offset(self, pointer):
        self._offset = pointer

    def get_offset(self):
        return self._offset

    def is_empty_offset(self):
        return self._offset is None

    def equals(self, hash_entry):
        return np.array_equal(self._position, hash_entry.get_position())

    def match_position(self, position):
        return np.array_equal(self._position, position)


class _LiveVoxel(Voxel):
    """Voxel of an entry returned by ``VoxelBucketTable.get_hash_entry``: ``integrate`` writes through to the
    device table (the reference returns the live object, hash_fusion.py:145), then refreshes this snapshot."""

    def __init__(self, table, position, sdf, color, weight):
        super().__init__(sdf, color, weight)
        self._table, self._position = table, position

    def integrate(self, new_dist, new_color, obs_weight=1.):
        if obs_weight != 1.:
            raise ValueError("the device container integrates with obs_weight 1, like the reference's own loop "
                             "(hash_fusion.py:141,145)")
        self._table.integrate_voxels([self._position], [new_dist], [new_color])
        _, vox = self._table.lookup([self._position], with_voxels=True)
        self._sdf, self._color, self._weight = float(vox[0, 0]), float(vox[0, 1]), float(vox[0, 2])


class _LiveHashEntry(HashEntry):
    def integrate_voxel(self, new_dist, new_color, obs_weight=1.):
        """hash_entry.py:18-21, executed on the device table."""
        self._voxel.integrate(new_dist, new_color, obs_weight)


class FrozenVoxel(Voxel):
    """Snapshot of a voxel that ``HashTable.integrate`` wrote into the block map: read-only.  The reference hands
    out the live object; here the map lives on the device, so mutation has to go through ``HashTable.integrate`` or
    ``HashTable.remove`` -- calling ``integrate`` on the snapshot raises instead of silently changing nothing."""

    def integrate(self, new_dist, new_color, obs_weight=1.):
        raise TypeError("read-only snapshot of a voxel of the device block map: update the map with "
                        "HashTable.integrate(); HashTable.remove(coord) deletes the entry")


class Bucket:
    """Host-side value type of data_structures/bucket.py:7-143: a fixed list of ``bucket_size``
    entry slots, the last one reserved for the head of the bucket's collision chain.  The device
    container stores the same thing as five consecutive table rows; this class exists so that code
    written against the reference's ``Bucket`` (its tests/bucket_test.py) keeps working."""

    def __init__(self, bucket_size=5):
        self._bucket_size = bucket_size
        self._bucket_list = [None] * self._bucket_size

    def get_ith_entry(self, i):
        try:
            return self._bucket_list[i]
        except IndexError:
            print("bucket.get_ith_entry: index out of range")

    def set_ith_entry(self, hash_entry, i):
        self._bucket_list[i] = hash_entry

    def remove_ith_entry(self, i):
        self._bucket_list[i] = None

    def is_empty_entry(self, i):
        return self._bucket_list[i] is None

    def get_num_entry_stored(self):
        return sum(1 for e in self._bucket_list if e is not None)

    def is_full(self):
        return self.get_num_entry_stored() == self._bucket_size

    def is_full_for_alien_entries(self):
        return all(e is not None for e in self._bucket_list[:-1])

    def is_empty(self):
        return self.get_num_entry_stored() == 0

    def is_overflow(self):
        last = self._bucket_list[-1]
        return last is not None and not last.is_empty_offset()

    def _first_free(self, upto):
        for i in range(upto):
            if self._bucket_list[i] is None:
                return i
        return -1

    def add_hash_entry(self, hash_entry):
        """Index of the slot the entry went into, -1 if the bucket is full."""
        i = self._first_free(self._bucket_size)
        if i >= 0:
            self._bucket_list[i] = hash_entry
        return i

    def add_alien_hash_entry(self, hash_entry):
        i = self._first_free(self._bucket_size - 1)  # the last slot only takes the bucket's own list head
        if i >= 0:
            self._bucket_list[i] = hash_entry
        return i

    def get_hash_entry(self, world_coord):
        for entry in self._bucket_list:
            if entry is not None and entry.match_position(world_coord):
                return entry
        return None

    def remove_hash_entry(self, world_coord):
        """Index of the slot that was cleared, -1 if no entry has this position."""
        for i, entry in enumerate(self._bucket_list):
            if entry is not None and entry.match_position(world_coord):
                self._bucket_list[i] = None
                return i
        return -1

    def contains(self, hash_entry):
        return any(e is not None and e.equals(hash_entry) for e in self._bucket_list)


def _payload(voxel):
    """(sdf, colour, weight) float64 of a Voxel; NaN sdf marks an entry stored without one."""
    if voxel is None:
        return (math.nan, 0.0, 0.0)
    f = [voxel.get_sdf(), voxel.get_color(), voxel.get_weight()]
    return tuple(math.nan if v is None else float(v) for v in f)


class VoxelBucketTable:
    def __init__(self, map_size=1000000, bucket_size=5, load_factor=0.75, device=0):
        self._lib = _capi.lib()
        self._handle = C.c_void_p()
        self._bucket_size = int(bucket_size)
        self._load_factor = 0.75 if load_factor is None else float(load_factor)
        _capi.check(self._lib.vbt_create(int(map_size), self._bucket_size, self._load_factor, int(device),
                                         C.byref(self._handle)))

    # -- batches ---------------------------------------------------------------------------------
    def apply(self, ops, positions, args=None):
        """Executes the operations in order on the device; returns (n, 2) int64 results
        (see ``vbt_apply`` in include/tsdf_b200.h)."""
        ops = np.ascontiguousarray(ops, dtype=np.int32).reshape(-1)
        pos = np.ascontiguousarray(positions, dtype=np.int64).reshape(-1, 3)
        if len(ops) != len(pos):
            raise ValueError("ops and positions differ in length")
        if args is not None:
            args = np.ascontiguousarray(args, dtype=np.float64).reshape(len(ops), 5)
        out = np.full((len(ops), 2), -1, np.int64)
        _capi.check(self._lib.vbt_apply(self._handle, _capi.ptr(ops), _capi.ptr(pos), _capi.ptr(args), len(ops),
                                        _capi.ptr(out)))
        return out

    def lookup(self, positions, with_voxels=False):
        """Parallel read-only ``get_hash_entry`` of a batch: (n, 2) (bucket, slot) or (-1, -1)."""
        pos = np.ascontiguousarray(positions, dtype=np.int64).reshape(-1, 3)
        out = np.full((len(pos), 2), -1, np.int64)
        vox = np.zeros((len(pos), 3), np.float64) if with_voxels else None
        _capi.check(self._lib.vbt_lookup(self._handle, _capi.ptr(pos), len(pos), _capi.ptr(out), _capi.ptr(vox)))
        return (out, vox) if with_voxels else out

    def integrate_voxels(self, positions, dists, colors):
        """The reference's per-voxel loop (hash_fusion.py:135-145) for a batch of (voxel, dist,
        folded colour): look up, create + insert on a miss, float64 running mean
        (data_structures/voxel.py:19-49) either way.  Returns where each voxel is stored."""
        pos = np.ascontiguousarray(positions, dtype=np.int64).reshape(-1, 3)
        args = np.zeros((len(pos), 5), np.float64)
        args[:, 0] = dists
        args[:, 1] = colors
        return self.apply(np.full(len(pos), OP_INTEGRATE, np.int32), pos, args)

    def integrate_frame(self, vol_dim, vol_origin, voxel_size, trunc_margin, color_im, depth_im, cam_intr, cam_pose):
        """The reference's ``HashTable.integrate`` as a whole (hash_fusion.py:103-145): the voxels the
        frame updates, in ``vox_coords`` order, through the per-voxel loop.  Returns their number."""
        im_h, im_w = depth_im.shape
        dim = np.ascontiguousarray(vol_dim, dtype=np.int64)
        origin = np.ascontiguousarray(vol_origin, dtype=np.float32)
        rgb = np.ascontiguousarray(color_im, dtype=np.uint8)
        depth = np.ascontiguousarray(depth_im, dtype=np.float64)
        K = np.ascontiguousarray(cam_intr, dtype=np.float64)
        tinv = np.ascontiguousarray(np.linalg.inv(np.asarray(cam_pose, dtype=np.float64)))  # hash_fusion.py:112
        n = C.c_int64()
        _capi.check(self._lib.vbt_integrate_frame(
            self._handle, _capi.ptr(dim), _capi.ptr(origin), float(voxel_size), float(trunc_margin), _capi.ptr(rgb),
            _capi.ptr(depth), im_h, im_w, _capi.ptr(K), _capi.ptr(tinv), C.byref(n)))
        return n.value

    def to_dense(self, vol_dim):
        """(tsdf, colour) float32 [X,Y,Z] volumes of the stored voxels (hash_fusion.py:442-463)."""
        dim = np.ascontiguousarray(vol_dim, dtype=np.int64)
        shape = tuple(int(d) for d in dim)
        t = np.empty(shape, np.float32)
        c = np.empty(shape, np.float32)
        _capi.check(self._lib.vbt_to_dense(self._handle, _capi.ptr(dim), _capi.ptr(t), _capi.ptr(c)))
        return t, c

    # -- reference container methods ---------------------------------------------------------------
    def hash_function(self, world_coord):
        """H(x,y,z) = (x*p1 xor y*p2 xor z*p3) mod n (hash_fusion.py:182-190)."""
        x, y, z = (int(v) for v in world_coord[:3])
        return ((x * P1) ^ (y * P2) ^ (z * P3)) % self.stats()["table_size"]

    def add_hash_entry(self, hash_entry):
        """Returns (ith_bucket, ith_entry) where the entry is stored, (-1, -1) if that failed
        (hash_fusion.py:199-224)."""
        if hash_entry is None:
            return -1, -1
        args = np.zeros((1, 5), np.float64)
        args[0, 2:] = _payload(hash_entry.get_voxel())
        r = self.apply([OP_ADD], [hash_entry.get_position()], args)[0]
        return int(r[0]), int(r[1])

    def add_voxel(self, voxel, world_coord):
        self.add_hash_entry(HashEntry(world_coord, None, voxel))

    def add_positions(self, positions, voxels=None):
        """Batch ``add_hash_entry``: (n, 2) slots, in order."""
        pos = np.ascontiguousarray(positions, dtype=np.int64).reshape(-1, 3)
        args = np.zeros((len(pos), 5), np.float64)
        args[:, 2] = math.nan
        if voxels is not None:
            args[:, 2:] = [_payload(v) for v in voxels]
        return self.apply(np.full(len(pos), OP_ADD, np.int32), pos, args)

    def get_hash_entry(self, world_coord):
        """The stored entry or None (hash_fusion.py:285-310): a snapshot whose ``integrate_voxel`` /
        ``get_voxel().integrate`` write through to the device table."""
        slots, vox = self.lookup([world_coord], with_voxels=True)
        b, e = int(slots[0, 0]), int(slots[0, 1])
        if b < 0:
            return None
        pos = tuple(int(v) for v in world_coord[:3])
        voxel = None if math.isnan(vox[0, 0]) else _LiveVoxel(self, pos, vox[0, 0], vox[0, 1], vox[0, 2])
        entry = _LiveHashEntry(pos, None, voxel)
        entry.slot = (b, e)
        return entry

    def get_voxel(self, world_coord):
        return self.get_hash_entry(world_coord).get_voxel()

    def remove_hash_entry(self, hash_entry):
        """1 if removed, 0 if not found (hash_fusion.py:331-384)."""
        return int(self.apply([OP_REMOVE], [hash_entry.get_position()])[0, 0])

    def remove(self, world_coord):
        self.remove_hash_entry(HashEntry(world_coord, None, None))

    def double_table_size(self):
        before = self.stats()["table_size"]
        print("Resizing hash table from {} to {}".format(before, before * 2))
        _capi.check(self._lib.vbt_double_table_size(self._handle))
        print("Resize finished.")

    # -- occupancy metrics (hash_fusion.py:147-180) ---------------------------------------------------
    def stats(self):
        s = np.zeros(6, np.int64)
        _capi.check(self._lib.vbt_stats(self._handle, _capi.ptr(s)))
        names = ["table_size", "non_empty_buckets", "entries", "collisions", "resizes", "bucket_size"]
        return {k: int(v) for k, v in zip(names, s)}

    def count_num_hash_entries(self):
        return self.stats()["entries"]

    def get_num_non_empty_bucket(self):
        return self.stats()["non_empty_buckets"]

    get_num_non_empty_buckets = get_num_non_empty_bucket

    def get_load_factor(self):
        s = self.stats()
        return s["non_empty_buckets"] / s["table_size"]

    def needs_resize(self):
        return self.get_load_factor() >= self._load_factor

    def get_num_collisions(self):
        return self.stats()["collisions"]

    # -- read-out --------------------------------------------------------------------------------
    def layout(self, with_voxels=False):
        """(n, 7) int64 rows {bucket, slot, x, y, z, offset bucket, offset slot} in slot order
        (offset -1, -1 = empty) and optionally the (n, 3) voxel payloads."""
        n = C.c_int64()
        _capi.check(self._lib.vbt_dump(self._handle, 0, C.byref(n), None, None))
        rows = np.zeros((n.value, 7), np.int64)
        vox = np.zeros((n.value, 3), np.float64) if with_voxels else None
        if n.value:
            _capi.check(self._lib.vbt_dump(self._handle, n.value, C.byref(n), _capi.ptr(rows), _capi.ptr(vox)))
        return (rows, vox) if with_voxels else rows

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle:
            self._lib.vbt_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
