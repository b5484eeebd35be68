"""Single-box multi-GPU plumbing: one process per GPU, torch.distributed over NCCL.

Neither map needs a per-frame collective (every rank receives the whole frame and updates only
the bricks / blocks it owns).  The only exchange is the read-out gather for get_volume().
The gather logic works on whatever backend the process group uses, so the world_size-2 CPU
tests run it over gloo.
"""
import os

import numpy as np

BRICK = 8


def rank_world(rank=None, world_size=None):
    """(rank, world) from the arguments, else the initialised process group, else (0, 1)."""
    if rank is not None and world_size is not None:
        return int(rank), int(world_size)
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def local_device(device=None):
    if device is not None:
        return int(device)
    return int(os.environ.get("LOCAL_RANK", "0"))


def sheet_owner(sheet, world):
    """Owner of brick-thick x-sheet(s): the sheets are dealt in snake order (0 1 .. W-1 W-1 .. 1 0 0 1 ..), which
    cancels the linear trend of a frustum's load along x that plain round-robin turns into a 7-8 % per-frame
    imbalance at 8 ranks.  Mirrors sheet_of() in csrc/launch.h."""
    m = np.asarray(sheet, dtype=np.int64) % (2 * world)
    return np.where(m < world, m, 2 * world - 1 - m)


def owned_x(dim_x, rank, world):
    """Global x indices (voxels) of the brick-thick x-sheets rank owns, ascending."""
    x = np.arange(int(dim_x), dtype=np.int64)
    return x[sheet_owner(x // BRICK, world) == rank]


def block_owner(keys, world):
    """Owner rank of 8^3 blocks (keys (n,3) int): a multiplicative mix of the packed key, so the
    owner is decorrelated from the table-slot hash ('shard hash != slot hash').
    Must match owner_of() in csrc/hash_kernels.cu."""
    k = np.asarray(keys, dtype=np.uint64)
    packed = (k[..., 0] << np.uint64(42)) | (k[..., 1] << np.uint64(21)) | k[..., 2]
    with np.errstate(over="ignore"):
        h = packed * np.uint64(0x9E3779B97F4A7C15)
    return ((h >> np.uint64(33)) % np.uint64(world)).astype(np.int64)


def gather_x_sheets(local, dim_x, rank, world, group=None):
    """All-gather per-rank [n_x_local, Y, Z] tensors and interleave the x-sheets back into the
    global [X, Y, Z] order.  `local` is a torch tensor on the group's device type."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    counts = [len(owned_x(dim_x, r, world)) for r in range(world)]
    n_max = max(counts)
    yz = tuple(local.shape[1:])
    padded = torch.zeros((n_max,) + yz, dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = torch.empty((world, n_max) + yz, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(world * n_max, *yz), padded, group=group)
    full = torch.empty((int(dim_x),) + yz, dtype=local.dtype, device=local.device)
    for r in range(world):
        xs = torch.as_tensor(owned_x(dim_x, r, world), device=local.device)
        full[xs] = out[r, :counts[r]]
    return full


def gather_block_lists(keys, payload, group=None):
    """Read-out gather of the hash map: every rank contributes its block list -- keys [n, 3] int32,
    payload [n, 3*512] float32 -- and receives the concatenation over ranks.  Counts are exchanged first, the lists
    are padded to the longest and all-gathered (two collectives of the blocks' own size instead of three all-reduces
    of dense volumes), then compacted.  Works on whatever backend the group uses (NCCL tensors on the GPU, gloo in the
    CPU tests)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = torch.tensor([keys.shape[0]], dtype=torch.int64, device=keys.device)
    counts = torch.zeros(world, dtype=torch.int64, device=keys.device)
    dist.all_gather_into_tensor(counts, n, group=group)
    counts = [int(c) for c in counts.tolist()]
    n_max = max(max(counts), 1)
    k_pad = torch.zeros((n_max, 3), dtype=keys.dtype, device=keys.device)
    p_pad = torch.zeros((n_max, payload.shape[1]), dtype=payload.dtype, device=payload.device)
    k_pad[:keys.shape[0]] = keys
    p_pad[:payload.shape[0]] = payload
    k_all = torch.empty((world * n_max, 3), dtype=keys.dtype, device=keys.device)
    p_all = torch.empty((world * n_max, payload.shape[1]), dtype=payload.dtype, device=payload.device)
    dist.all_gather_into_tensor(k_all, k_pad, group=group)
    dist.all_gather_into_tensor(p_all, p_pad, group=group)
    sel = torch.cat([torch.arange(r * n_max, r * n_max + counts[r], device=keys.device) for r in range(world)])
    return k_all[sel].contiguous(), p_all[sel].contiguous()


def merge_disjoint_volumes(tsdf, weight, color, group=None):
    """Hash read-out: every rank scattered only its own blocks into dense volumes with the
    defaults (1, 0, 0) elsewhere.  Blocks are disjoint and a voxel that was never updated holds
    exactly (1, 0, 0) whether or not its block exists, so zeroing the tsdf of w == 0 voxels,
    summing over ranks (x + 0 + ... + 0 is exact) and restoring tsdf = 1 where the summed weight
    is 0 reconstructs the union bit for bit."""
    import torch
    import torch.distributed as dist
    t = torch.where(weight > 0, tsdf, torch.zeros_like(tsdf))
    w = weight.clone()
    c = color.clone()
    for x in (t, w, c):
        dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
    return torch.where(w > 0, t, torch.ones_like(t)), w, c


class FrameFanout:
    """Frame fan-out for a map sharded over the GPUs of one box (BASELINE north_star (3): "each frame's depth,
    colour and pose broadcast to all ranks"; the reference reads every frame in its one process,
    grid_fusion_demos/grid_demo1.py:76-87).

    Without it every rank uploads every frame: N copies of the same image over the host's PCIe complex per step,
    which is what bounds the end-to-end rate from 4 ranks on.  With it frame i crosses PCIe ONCE, on rank
    ``i % world`` (the uploads rotate, so every GPU's link carries 1/N of the stream), and reaches the other
    ranks over NVLink; each rank then integrates its own shard from the resident copy.  This is the path's one
    per-frame exchange step; nothing else about the maps changes (no collective on the voxel data).

    transport "peer" (default on a GPU): the native fan-out of csrc/capi_fanout.cu -- the source rank stages the
    frame in a ring slot of its arena, every peer's copy engine pulls it from there into its own slot (CUDA IPC
    mappings, exchanged here by one all-gather at construction); "staged" / "pulled" are sequence-number flags
    waited for on the device.
    transport "collective": one torch.distributed.broadcast per frame on this object's stream -- works on any
    backend (the gloo CPU tests use it); on NCCL its kernel has to find room beside the persistent integrate
    grids and was measured slower than every rank uploading for itself (profiles/r2_fanout.md).

    Ordering is by stream, never by host wait: the frame arrives on the fan-out's stream, the map's preparation
    kernel is ordered after it (``producer_stream`` of ``integrate_device``), and the stream is ordered after that
    kernel before the slot may be overwritten.

    ``target`` is a grid_fusion.TSDFVolume or hash_fusion.HashTable of this rank.  Every rank calls ``integrate``
    for every frame with the frame's intrinsics and pose (200 bytes of control data the tracker hands to all
    ranks); only the source rank's ``color_im`` / ``depth_im`` are read, the others may pass None.  The source
    rank's host arrays are read by the copy engine after the call returns when they are pinned: they must stay
    unchanged until that rank's next upload (which waits for the copy first) or ``drain()`` returns.
    """

    def __init__(self, target, im_h, im_w, raw_depth=False, depth_scale=1000., invalid_raw=65535, group=None,
                 slots=None, device=None, transport=None, rank=None, world_size=None):
        import torch
        self.target = target
        # the ranks that share the frames: the map's own unless the caller fans out over another group
        self.rank = int(target.rank if rank is None else rank)
        self.world = int(target.world_size if world_size is None else world_size)
        self.group = group
        self.h, self.w = int(im_h), int(im_w)
        self.raw = bool(raw_depth)
        self.depth_scale, self.invalid_raw = float(depth_scale), int(invalid_raw)
        self.depth_dtype = np.uint16 if self.raw else np.float64
        npix = self.h * self.w
        self.depth_bytes = npix * np.dtype(self.depth_dtype).itemsize
        self.rgb_bytes = npix * 3
        self.rgb_off = (self.depth_bytes + 255) // 256 * 256
        self.nbytes = self.rgb_off + self.rgb_bytes
        self.device = torch.device("cuda", target.device) if device is None else torch.device(device)
        self.on_gpu = self.device.type == "cuda"
        self.transport = transport or ("peer" if self.on_gpu else "collective")
        # ring slots per rank: with at least `world` of them the uploads of a whole rotation can be in flight
        self.slots = min(8, max(3, self.world)) if slots is None else max(1, int(slots))
        self.frames = 0
        self._handle = None
        if self.transport == "peer":
            if not self.on_gpu:
                raise ValueError("the peer transport moves frames between GPUs")
            from . import _capi
            import ctypes as C
            self._capi, self._lib = _capi, _capi.lib()
            h = C.c_void_p()
            _capi.check(self._lib.tsdf_fanout_create(self.device.index, self.rank, self.world, self.nbytes, self.slots,
                                                     C.byref(h)))
            self._handle = h
            if self.world > 1:
                import torch.distributed as dist
                mine = np.zeros(64, np.uint8)
                _capi.check(self._lib.tsdf_fanout_export(h, _capi.ptr(mine)))
                every = torch.empty(self.world * 64, dtype=torch.uint8, device=self.device)
                dist.all_gather_into_tensor(every, torch.from_numpy(mine).to(self.device), group=group)
                _capi.check(self._lib.tsdf_fanout_connect(h, _capi.ptr(np.ascontiguousarray(every.cpu().numpy()))))
            self._out = (C.c_void_p(), C.c_void_p(), C.c_void_p())
            self._refs = tuple(C.byref(x) for x in self._out)
            # the product's own maps take the whole step in one native call; anything else goes through the
            # target's integrate_device[_raw] methods
            from . import grid_fusion, hash_fusion
            self._native = None
            if isinstance(target, grid_fusion.TSDFVolume):
                self._native = (target._handle, None)
            elif isinstance(target, hash_fusion.HashTable) and not getattr(target, "_per_voxel", False):
                self._native = (None, target._handle)
            self._k = np.zeros(9, np.float64)
            self._k_ptr = self._k.ctypes.data
        elif self.transport == "collective":
            self.ring = [torch.empty(self.nbytes, dtype=torch.uint8, device=self.device) for _ in range(self.slots)]
            self.stream = torch.cuda.Stream(device=self.device) if self.on_gpu else None
            self._uploaded = None  # event behind this rank's last host->device copy
        else:
            raise ValueError("unknown transport %r" % (transport,))

    def source_of(self, index):
        """Rank that uploads frame ``index`` (rotating, so the PCIe load is shared)."""
        return int(index) % self.world

    def payload_bytes(self):
        """Bytes one frame moves: once over PCIe on its source rank, once over NVLink to each other rank."""
        return self.depth_bytes + self.rgb_bytes

    def _images(self, color_im, depth_im):
        if color_im is None or depth_im is None:
            raise ValueError("rank %d is the source of frame %d and needs its images" % (self.rank, self.frames))
        depth = np.ascontiguousarray(depth_im, dtype=self.depth_dtype)
        rgb = np.ascontiguousarray(color_im, dtype=np.uint8)
        if depth.shape != (self.h, self.w) or rgb.shape != (self.h, self.w, 3):
            raise ValueError("frame of %r / %r does not match the fan-out's %dx%d" % (rgb.shape, depth.shape, self.w, self.h))
        return rgb, depth

    def _collective(self, src, color_im, depth_im):
        import torch
        buf = self.ring[self.frames % self.slots]
        ctx = torch.cuda.stream(self.stream) if self.on_gpu else _NoContext()
        with ctx:
            if self.rank == src:
                rgb, depth = self._images(color_im, depth_im)
                if self._uploaded is not None:  # at most one borrowed host frame in flight
                    self._uploaded.synchronize()
                buf[:self.depth_bytes].copy_(torch.from_numpy(depth.reshape(-1).view(np.uint8)), non_blocking=True)
                buf[self.rgb_off:].copy_(torch.from_numpy(rgb.reshape(-1)), non_blocking=True)
                if self.on_gpu:
                    self._uploaded = torch.cuda.Event()
                    self._uploaded.record(self.stream)
            if self.world > 1:
                import torch.distributed as dist
                g_src = dist.get_global_rank(self.group, src) if self.group is not None else src
                dist.broadcast(buf, src=g_src, group=self.group)
        return buf[:self.depth_bytes], buf[self.rgb_off:], (self.stream.cuda_stream if self.on_gpu else None)

    def _peer(self, src, color_im, depth_im):
        lib, h, seq = self._lib, self._handle, self.frames
        if self.rank == src:
            rgb, depth = self._images(color_im, depth_im)
            self._capi.check(lib.tsdf_fanout_push(h, seq, depth.ctypes.data, self.depth_bytes, rgb.ctypes.data,
                                                  self.rgb_bytes, self.rgb_off))
        self._capi.check(lib.tsdf_fanout_acquire(h, seq, src, self.nbytes, self.rgb_off, *self._refs))
        d_depth, d_rgb, stream = self._out
        return d_depth.value, d_rgb.value, stream.value

    def integrate(self, color_im, depth_im, cam_intr, cam_pose, obs_weight=1., src=None):
        """Integrate one frame on every rank; ``src`` (default: rotating) is the rank that holds its images."""
        src = self.source_of(self.frames) if src is None else int(src)
        if not 0 <= src < self.world:
            raise ValueError("source rank %d outside the world of %d" % (src, self.world))
        tinv = np.linalg.inv(np.asarray(cam_pose, dtype=np.float64))  # grid_fusion.py:265
        if self._handle is not None and self._native is not None:
            if self.rank == src:
                rgb, depth = self._images(color_im, depth_im)
                p_rgb, p_depth = rgb.ctypes.data, depth.ctypes.data
            else:
                p_rgb = p_depth = None
            self._k.reshape(3, 3)[...] = cam_intr
            self._capi.check(self._lib.tsdf_fanout_integrate(
                self._handle, self._native[0], self._native[1], self.frames, src, p_depth, p_rgb, self.h, self.w,
                int(self.raw), self._k_ptr, tinv.ctypes.data, float(obs_weight), self.depth_scale, self.invalid_raw))
            if hasattr(self.target, "_host_cache"):
                self.target._host_cache = None  # TSDFVolume's read-out cache
            self.frames += 1
            return
        d_depth, d_rgb, producer = (self._peer if self.transport == "peer" else self._collective)(src, color_im, depth_im)
        if self.raw:
            self.target.integrate_device_raw(d_rgb, d_depth, self.h, self.w, cam_intr, tinv, obs_weight,
                                             self.depth_scale, self.invalid_raw, producer_stream=producer)
        else:
            self.target.integrate_device(d_rgb, d_depth, self.h, self.w, cam_intr, tinv, obs_weight,
                                         producer_stream=producer)
        if self._handle is not None:  # the preparation kernel is queued: the peers may learn when it is through
            self._capi.check(self._lib.tsdf_fanout_release(self._handle, self.frames))
        self.frames += 1

    def drain(self):
        """Host waits until every upload and transfer issued so far has completed (raises if a peer failed to
        deliver or consume a frame within the time-out)."""
        if self._handle is not None:
            self._capi.check(self._lib.tsdf_fanout_drain(self._handle))
        elif self.on_gpu:
            self.stream.synchronize()
            self._uploaded = None

    def close(self):
        """Collective at world > 1: the peers write into this rank's arena until every rank is through."""
        if self._handle is None:
            return
        self.drain()
        self.target.sync()  # the map's preparation kernels read the arena
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier(group=self.group)
        self._lib.tsdf_fanout_destroy(self._handle)
        self._handle = None

    def __del__(self):
        # never a collective from a finaliser: an unclosed fan-out leaks its arena instead of risking a peer's write
        # into freed memory
        pass


class _NoContext:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False
