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
