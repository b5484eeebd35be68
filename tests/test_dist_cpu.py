"""Multi-rank host logic on CPU: world_size-2 gloo process groups (-m "not gpu").

The kernels need no per-frame collective; what is distributed is (1) the ownership maps and
(2) the read-out gather/merge, which are exercised here over gloo with the CPU oracle standing
in for the per-rank shards (the oracle accepts the same x-sheet subsets).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import frames
from oracle import oracle
from union_thesis_slam_b200 import dist as tdist


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return dict(ret)


def test_owned_x_is_a_partition_with_brick_thick_snake_dealt_sheets():
    for dim_x, world in [(405, 2), (405, 8), (1000, 4), (7, 3), (64, 8), (100, 1)]:
        seen = np.zeros(dim_x, int)
        for r in range(world):
            xs = tdist.owned_x(dim_x, r, world)
            m = (xs // 8) % (2 * world)   # snake: 0 1 .. W-1 W-1 .. 1 0
            assert np.all(np.where(m < world, m, 2 * world - 1 - m) == r) and np.all(np.diff(xs) > 0)
            seen[xs] += 1
        assert np.all(seen == 1)


def test_snake_dealing_cancels_a_linear_load_trend():
    """What the snake order is for: a load that grows linearly along x (a frustum's cross-section) splits evenly."""
    world, sheets = 8, 128
    load = 10.0 + np.arange(sheets)
    owner = tdist.sheet_owner(np.arange(sheets), world)
    per_rank = np.bincount(owner, weights=load, minlength=world)
    assert per_rank.max() / per_rank.mean() < 1.001
    round_robin = np.bincount(np.arange(sheets) % world, weights=load, minlength=world)
    assert round_robin.max() / round_robin.mean() > 1.04


def test_block_owner_is_balanced_and_decorrelated_from_slot_hash():
    g = np.stack(np.meshgrid(np.arange(51), np.arange(33), np.arange(37), indexing="ij"), -1).reshape(-1, 3)
    for world in (2, 4, 8):
        own = tdist.block_owner(g, world)
        counts = np.bincount(own, minlength=world)
        assert counts.min() > 0.9 * len(g) / world and counts.max() < 1.1 * len(g) / world
        # with owner = slot_hash % P the keys of one rank use 1/P of a 2^k table
        slot = ((g[:, 0] * 73856093) ^ (g[:, 1] * 19349669) ^ (g[:, 2] * 83492791)) % 65536
        mine = slot[own == 0]
        assert len(np.unique(mine % world)) == world


def _gather_job(rank, world):
    z = np.load(os.path.join(frames.GOLDEN_DIR, "small_dense_12.npz"))
    X = z["tsdf"].shape[0]
    xs = tdist.owned_x(X, rank, world)
    vol = oracle.OracleTSDFVolume(z["vol_bnds"].copy(), float(z["voxel_size"]), xs=xs)
    for i in range(len(z["obs"])):
        vol.integrate(z["color"][i], z["depth"][i], z["K"], z["pose"][i], obs_weight=float(z["obs"][i]))
    out = []
    for plane, ref in ((vol._tsdf_vol_cpu, z["tsdf"]), (vol._weight_vol_cpu, z["weight"]),
                       (vol._color_vol_cpu, z["color_vol"])):
        full = tdist.gather_x_sheets(torch.from_numpy(plane), X, rank, world).numpy()
        out.append(bool(np.array_equal(full, ref)))
    assert tdist.rank_world() == (rank, world)
    return out


def test_x_sheet_gather_reconstructs_the_reference_volume_world2():
    res = _spawn(_gather_job, 2)
    assert res == {0: [True, True, True], 1: [True, True, True]}


def test_x_sheet_gather_world3_uneven_sheets():
    res = _spawn(_gather_job, 3)  # 33 rows -> 5 sheets: ranks own 2/2/1 sheets, the last one partial
    assert all(all(v) for v in res.values()) and len(res) == 3


def _merge_job(rank, world):
    z = np.load(os.path.join(frames.GOLDEN_DIR, "small_dense_11.npz"))
    t, w, c = z["tsdf"], z["weight"], z["color_vol"]
    idx = np.stack(np.meshgrid(*[np.arange(s) for s in t.shape], indexing="ij"), -1)
    mine = tdist.block_owner(idx // 8, world) == rank
    lt = np.where(mine, t, 1.0).astype(np.float32)
    lw = np.where(mine, w, 0.0).astype(np.float32)
    lc = np.where(mine, c, 0.0).astype(np.float32)
    mt, mw, mc = tdist.merge_disjoint_volumes(torch.from_numpy(lt), torch.from_numpy(lw), torch.from_numpy(lc))
    return [bool(np.array_equal(mt.numpy(), t)), bool(np.array_equal(mw.numpy(), w)), bool(np.array_equal(mc.numpy(), c))]


def test_hash_shard_merge_is_bit_exact_world2():
    res = _spawn(_merge_job, 2)
    assert res == {0: [True, True, True], 1: [True, True, True]}


def _block_list_job(rank, world):
    """Hash read-out by block lists: each rank holds the 8^3 blocks it owns (keys + payload), gather_block_lists
    returns the union on every rank, and scattering it reconstructs the reference volume."""
    z = np.load(os.path.join(frames.GOLDEN_DIR, "small_dense_11.npz"))
    t, w, c = z["tsdf"], z["weight"], z["color_vol"]
    nb = [(s + 7) // 8 for s in t.shape]
    pad = [n * 8 for n in nb]

    def padded(a, fill):
        out = np.full(pad, fill, np.float32)
        out[:a.shape[0], :a.shape[1], :a.shape[2]] = a
        return out

    pt, pw, pc = padded(t, 1.0), padded(w, 0.0), padded(c, 0.0)
    keys = np.stack(np.meshgrid(*[np.arange(n) for n in nb], indexing="ij"), -1).reshape(-1, 3)
    has = np.array([pw[k[0] * 8:k[0] * 8 + 8, k[1] * 8:k[1] * 8 + 8, k[2] * 8:k[2] * 8 + 8].max() > 0 for k in keys])
    keys = keys[has]                                      # a block exists iff one of its voxels was updated
    mine = keys[tdist.block_owner(keys, world) == rank]
    if rank == world - 1:
        mine = mine[:0] if world > 2 else mine            # world 3: one rank with an empty list
    payload = np.stack([np.concatenate([p[k[0] * 8:k[0] * 8 + 8, k[1] * 8:k[1] * 8 + 8, k[2] * 8:k[2] * 8 + 8].reshape(-1)
                                        for p in (pt, pw, pc)]) for k in mine]) if len(mine) else np.zeros((0, 1536), np.float32)
    gk, gp = tdist.gather_block_lists(torch.from_numpy(mine.astype(np.int32)), torch.from_numpy(payload.astype(np.float32)))
    gk, gp = gk.numpy(), gp.numpy()
    ot, ow, oc = np.ones(pad, np.float32), np.zeros(pad, np.float32), np.zeros(pad, np.float32)
    for k, p in zip(gk, gp):
        sl = (slice(k[0] * 8, k[0] * 8 + 8), slice(k[1] * 8, k[1] * 8 + 8), slice(k[2] * 8, k[2] * 8 + 8))
        ot[sl], ow[sl], oc[sl] = p[:512].reshape(8, 8, 8), p[512:1024].reshape(8, 8, 8), p[1024:].reshape(8, 8, 8)
    cut = tuple(slice(0, s) for s in t.shape)
    return [len(gk), bool(np.array_equal(ot[cut], t)), bool(np.array_equal(ow[cut], w)), bool(np.array_equal(oc[cut], c)),
            int(has.sum())]


def test_block_list_gather_reconstructs_the_volume_world2():
    res = _spawn(_block_list_job, 2)
    for r in (0, 1):
        n, ok_t, ok_w, ok_c, total = res[r]
        assert n == total and ok_t and ok_w and ok_c


def test_block_list_gather_world3_with_an_empty_rank():
    res = _spawn(_block_list_job, 3)
    lens = {res[r][0] for r in range(3)}
    assert len(lens) == 1 and all(res[r][0] < res[r][4] for r in range(3))  # the last rank's blocks are missing everywhere


class _RecordingMap:
    """Stands in for a rank's TSDFVolume / HashTable: keeps what the fan-out hands to integrate_device."""

    def __init__(self, rank, world):
        self.rank, self.world_size, self.device = rank, world, 0
        self.got = []

    def integrate_device(self, d_rgb, d_depth, im_h, im_w, cam_intr, tinv, obs_weight=1., producer_stream=None):
        assert producer_stream is None  # host memory: nothing to order
        self.got.append((d_rgb.numpy().copy().reshape(im_h, im_w, 3), d_depth.numpy().copy().view(np.float64).reshape(im_h, im_w),
                         np.array(tinv), float(obs_weight)))

    def integrate_device_raw(self, d_rgb, d_raw, im_h, im_w, cam_intr, tinv, obs_weight=1., depth_scale=1000.,
                             invalid_raw=65535, producer_stream=None):
        self.got.append((d_rgb.numpy().copy().reshape(im_h, im_w, 3), d_raw.numpy().copy().view(np.uint16).reshape(im_h, im_w),
                         np.array(tinv), (float(depth_scale), int(invalid_raw))))


def _fanout_frames(n, h=6, w=10):
    rng = np.random.default_rng(11)
    out = []
    for i in range(n):
        pose = np.eye(4)
        pose[:3, 3] = rng.normal(size=3)
        out.append((rng.integers(0, 256, (h, w, 3), dtype=np.uint8), rng.uniform(0.0, 4.0, (h, w)),
                    rng.integers(0, 65536, (h, w)).astype(np.uint16), pose))
    return out


def _fanout_worker(rank, world):
    frs = _fanout_frames(5)
    K = np.eye(3)
    res = {}
    for raw in (False, True):
        m = _RecordingMap(rank, world)
        fan = tdist.FrameFanout(m, 6, 10, raw_depth=raw, depth_scale=500., invalid_raw=7, device="cpu")
        assert fan.payload_bytes() == 6 * 10 * ((2 if raw else 8) + 3)
        for i, (c, d, r, p) in enumerate(frs):
            mine = fan.source_of(i) == rank
            # only the frame's source rank holds its images
            fan.integrate(c if mine else None, (r if raw else d) if mine else None, K, p, obs_weight=1.0 + i)
        fan.drain()
        res[raw] = m.got
    with pytest.raises(ValueError):  # the source rank must supply the images
        tdist.FrameFanout(_RecordingMap(rank, world), 6, 10, device="cpu").integrate(None, None, K, np.eye(4), src=rank)
    return res


def test_frame_fanout_delivers_every_frame_to_every_rank_world2():
    frs = _fanout_frames(5)
    out = _spawn(_fanout_worker, world=2)
    for rank in (0, 1):
        for raw in (False, True):
            got = out[rank][raw]
            assert len(got) == len(frs)
            for i, (c, d, r, p) in enumerate(frs):
                assert np.array_equal(got[i][0], c)
                assert np.array_equal(got[i][1], r if raw else d)
                assert np.array_equal(got[i][2], np.linalg.inv(p))
                assert got[i][3] == ((500.0, 7) if raw else 1.0 + i)


def test_frame_fanout_rotates_the_upload_rank():
    fan = tdist.FrameFanout(_RecordingMap(1, 4), 4, 4, device="cpu")
    assert [fan.source_of(i) for i in range(9)] == [0, 1, 2, 3, 0, 1, 2, 3, 0]
    with pytest.raises(ValueError):
        fan.integrate(None, None, np.eye(3), np.eye(4), src=4)


def test_frame_fanout_rejects_what_it_cannot_do():
    with pytest.raises(ValueError):   # frames move between GPUs on the peer transport
        tdist.FrameFanout(_RecordingMap(0, 1), 4, 4, device="cpu", transport="peer")
    with pytest.raises(ValueError):
        tdist.FrameFanout(_RecordingMap(0, 1), 4, 4, device="cpu", transport="carrier-pigeon")
    fan = tdist.FrameFanout(_RecordingMap(0, 1), 4, 6, device="cpu")
    assert fan.transport == "collective" and fan.slots == 3 and fan.payload_bytes() == 4 * 6 * 11
    with pytest.raises(ValueError):   # wrong image size
        fan.integrate(np.zeros((4, 4, 3), np.uint8), np.zeros((4, 4)), np.eye(3), np.eye(4))
    # a one-rank fan-out is the plain hand-over of a staged frame
    fan.integrate(np.full((4, 6, 3), 7, np.uint8), np.full((4, 6), 1.5), np.eye(3), np.eye(4), obs_weight=2.0)
    rgb, depth, tinv, w = fan.target.got[0]
    assert (rgb == 7).all() and (depth == 1.5).all() and np.array_equal(tinv, np.eye(4)) and w == 2.0
    fan.close()
