"""GPU parity of the dense brick grid against the oracle and the reference goldens (-m gpu).

Tolerances (BASELINE.json north_star): weights bit-exact, |tsdf| <= 1e-5, colour +-1 per channel.
The kernels are written to be bit-exact on all three planes, so the strict tests also assert
array_equal / sha256 equality with the reference.
"""
import glob
import os

import numpy as np
import pytest

import frames
import scenes
from oracle import oracle

pytestmark = pytest.mark.gpu

TSDF_TOL = 1e-5


def _unpack(c):
    c = c.astype(np.int64)
    return np.stack([c >> 16, (c >> 8) & 255, c & 255], -1)


def assert_parity(got, want, strict=True):
    (t, w, c), (tr, wr, cr) = got, want
    assert t.shape == tr.shape
    assert np.array_equal(w, wr), "weights differ in %d voxels" % int((w != wr).sum())
    assert float(np.abs(t.astype(np.float64) - tr).max()) <= TSDF_TOL
    assert int(np.abs(_unpack(c) - _unpack(cr)).max()) <= 1
    if strict:
        assert np.array_equal(t, tr), "tsdf not bit-exact in %d voxels" % int((t != tr).sum())
        assert np.array_equal(c, cr), "colour not bit-exact in %d voxels" % int((c != cr).sum())


def _gpu_volume(bnds, vs, **kw):
    from union_thesis_slam_b200 import grid_fusion
    return grid_fusion.TSDFVolume(np.array(bnds, dtype=np.float64), vs, **kw)


def _state(vol):
    t, c = vol.get_volume()
    return t, vol._weight_vol_cpu, c


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(frames.GOLDEN_DIR, "small_dense_*.npz"))))
def test_small_golden_cases_match_reference(path):
    z = np.load(path)
    vol = _gpu_volume(z["vol_bnds"], float(z["voxel_size"]))
    for i in range(len(z["obs"])):
        vol.integrate(z["color"][i], z["depth"][i], z["K"], z["pose"][i], obs_weight=float(z["obs"][i]))
    assert tuple(vol._vol_dim) == z["tsdf"].shape
    assert_parity(_state(vol), (z["tsdf"], z["weight"], z["color_vol"]))
    assert vol.stats()["updated_total"] > 0


def test_config1_ten_frames_sha256_equals_reference():
    """Bundled sequence, 2 cm grid (BASELINE.json configs[0]): sha256 of all three planes equals
    the unmodified reference after 1, 2 and 10 frames; counts equal the reference's pstats dumps."""
    G = frames.golden()
    K = frames.intrinsics()
    vol = _gpu_volume(scenes.CONFIG1_BOUNDS, scenes.CONFIG1_VOXEL)
    assert list(vol._vol_dim) == [405, 264, 289]
    sum_w = 0
    for n, i in enumerate(range(10), 1):
        c, d, p = frames.load_frame(i)
        vol.integrate(c, d, K, p, obs_weight=1.)
        sum_w += vol.stats()["updated_last"]
        if str(n) in G["dense"]:
            g = G["dense"][str(n)]
            t, w, col = _state(vol)
            assert frames.sha(w) == g["weight"]
            assert frames.sha(t) == g["tsdf"]
            assert frames.sha(col) == g["color"]
            assert int((w > 0).sum()) == g["n_w_pos"]
            assert sum_w == g["sum_w"]
    assert sum_w == 3557017 and vol.stats()["updated_total"] == 3557017


@pytest.mark.parametrize("seed,dims,vs,kw", [
    (101, (40, 40, 40), 0.02, {}),
    (102, (7, 9, 5), 0.04, {}),                       # smaller than one brick in z
    (103, (65, 17, 33), 0.015, dict(obs_weights=(0.25, 1.75, 3.0))),
    (104, (50, 48, 52), 0.02, dict(inside=True, nframes=5)),
    (105, (31, 31, 31), 0.01, dict(depth_dtype=np.float32, quantise=False)),
    (106, (96, 20, 24), 0.03, dict(hw=(120, 160), nframes=4)),
])
def test_random_volumes_match_oracle(seed, dims, vs, kw):
    case = scenes.small_case(seed, dims=dims, voxel_size=vs, **kw)
    vol = _gpu_volume(case["vol_bnds"].copy(), vs)
    orc = oracle.OracleTSDFVolume(case["vol_bnds"].copy(), vs)
    for c, d, p, obs in case["frames"]:
        vol.integrate(c, d, case["K"], p, obs_weight=obs)
        n = orc.integrate(c, d, case["K"], p, obs_weight=obs)
        assert vol.stats()["updated_last"] == n
    assert_parity(_state(vol), (orc._tsdf_vol_cpu, orc._weight_vol_cpu, orc._color_vol_cpu))


def test_synthetic_room_matches_oracle():
    """Config-4-shaped workload at a size the oracle finishes in seconds (3 m x 2.5 m x 2 m, 1 cm)."""
    from union_thesis_slam_b200 import synth
    room = synth.room_bounds((3.0, 2.5, 2.0))
    fr, K = synth.room_stream(4, room, seed=3)
    vol = _gpu_volume(room.copy(), 0.01)
    orc = oracle.OracleTSDFVolume(room.copy(), 0.01)
    for c, d, p in fr:
        vol.integrate(c, d, K, p)
        orc.integrate(c, d, K, p)
    assert orc.updates_per_frame[0] > 100000
    assert_parity(_state(vol), (orc._tsdf_vol_cpu, orc._weight_vol_cpu, orc._color_vol_cpu))


def test_x_sheet_shards_union_equals_single():
    """Logical ranks run one after another on one GPU (no spin-waiting between ranks): the union
    of the brick-thick cyclic x-sheets equals the single-rank result bit for bit."""
    case = scenes.small_case(107, dims=(70, 24, 30), voxel_size=0.02, nframes=3)
    single = _gpu_volume(case["vol_bnds"].copy(), 0.02)
    for c, d, p, obs in case["frames"]:
        single.integrate(c, d, case["K"], p, obs_weight=obs)
    ref = _state(single)
    for world in (2, 3, 8):
        t = np.full_like(ref[0], np.nan); w = np.full_like(t, np.nan); col = np.full_like(t, np.nan)
        seen = np.zeros(ref[0].shape[0], bool)
        for r in range(world):
            vol = _gpu_volume(case["vol_bnds"].copy(), 0.02, rank=r, world_size=world)
            for c, d, p, obs in case["frames"]:
                vol.integrate(c, d, case["K"], p, obs_weight=obs)
            lt, lw, lc, xs = vol.local_volume()
            assert not seen[xs].any()
            seen[xs] = True
            t[xs], w[xs], col[xs] = lt, lw, lc
            vol.close()
        assert seen.all()
        assert_parity((t, w, col), ref)


def test_device_resident_frames_and_pinned_host_buffers():
    import torch
    case = scenes.small_case(108, dims=(33, 33, 33), voxel_size=0.02, nframes=2)
    a = _gpu_volume(case["vol_bnds"].copy(), 0.02)
    b = _gpu_volume(case["vol_bnds"].copy(), 0.02)
    p = _gpu_volume(case["vol_bnds"].copy(), 0.02)
    for c, d, pose, obs in case["frames"]:
        a.integrate(c, d, case["K"], pose, obs_weight=obs)
        dc = torch.from_numpy(c).cuda()
        dd = torch.from_numpy(d).cuda()
        torch.cuda.synchronize()
        b.integrate_device(dc, dd, d.shape[0], d.shape[1], case["K"], np.linalg.inv(pose), obs)
        b.sync()
        pc = torch.from_numpy(c).pin_memory()
        pd = torch.from_numpy(d).pin_memory()
        p.integrate(pc.numpy(), pd.numpy(), case["K"], pose, obs_weight=obs)
    assert_parity(_state(b), _state(a))
    assert_parity(_state(p), _state(a))


def test_frame_order_and_buffer_reuse():
    """The caller may overwrite its frame buffers as soon as integrate() returns."""
    case = scenes.small_case(109, dims=(30, 30, 30), voxel_size=0.02, nframes=6)
    vol = _gpu_volume(case["vol_bnds"].copy(), 0.02)
    orc = oracle.OracleTSDFVolume(case["vol_bnds"].copy(), 0.02)
    cbuf = np.empty_like(case["frames"][0][0]); dbuf = np.empty_like(case["frames"][0][1])
    for c, d, p, obs in case["frames"]:
        cbuf[...] = c; dbuf[...] = d
        vol.integrate(cbuf, dbuf, case["K"], p, obs_weight=obs)
        cbuf[...] = 0; dbuf[...] = 0
        orc.integrate(c, d, case["K"], p, obs_weight=obs)
    assert_parity(_state(vol), (orc._tsdf_vol_cpu, orc._weight_vol_cpu, orc._color_vol_cpu))


def test_bad_arguments_raise():
    from union_thesis_slam_b200 import _capi
    case = scenes.small_case(110, dims=(16, 16, 16), voxel_size=0.02, nframes=1)
    vol = _gpu_volume(case["vol_bnds"].copy(), 0.02)
    c, d, p, _ = case["frames"][0]
    with pytest.raises(_capi.TSDFError):
        vol.integrate(c, d, case["K"], p, obs_weight=0.0)
    with pytest.raises((_capi.TSDFError, np.linalg.LinAlgError)):
        vol.integrate(c, d, case["K"], np.full((4, 4), np.nan), obs_weight=1.0)
    with pytest.raises(AssertionError):
        _gpu_volume(np.zeros((2, 3)), 0.02)
    with pytest.raises(_capi.TSDFError):
        _gpu_volume(case["vol_bnds"].copy(), 0.02, rank=3, world_size=2)


def test_degenerate_frames_touch_nothing_and_odd_image_sizes():
    """All-invalid depth, a camera looking away from the volume, and an image whose sides are not
    multiples of the 8-pixel tiles."""
    case = scenes.small_case(111, dims=(30, 22, 26), voxel_size=0.02, nframes=3, hw=(50, 70))
    vol = _gpu_volume(case["vol_bnds"].copy(), 0.02)
    orc = oracle.OracleTSDFVolume(case["vol_bnds"].copy(), 0.02)
    c, d, p, _ = case["frames"][0]
    vol.integrate(c, np.zeros_like(d), case["K"], p)
    assert vol.stats()["updated_last"] == 0 and vol.stats()["bricks_visited_last"] == 0
    away = p.copy()
    away[:3, :3] = away[:3, :3] @ np.diag([1.0, -1.0, -1.0])  # rotate 180 degrees about the camera x axis
    vol.integrate(c, d, case["K"], away)
    orc.integrate(c, d, case["K"], away)
    assert vol.stats()["updated_last"] == orc.updates_per_frame[-1]
    for c, d, p, obs in case["frames"]:
        vol.integrate(c, d, case["K"], p, obs_weight=obs)
        n = orc.integrate(c, d, case["K"], p, obs_weight=obs)
        assert vol.stats()["updated_last"] == n
    assert_parity(_state(vol), (orc._tsdf_vol_cpu, orc._weight_vol_cpu, orc._color_vol_cpu))


def test_camera_inside_volume_near_plane_voxels():
    """Voxels a few millimetres in front of / behind the camera plane: Z <= 0 rejects, tiny Z gives huge
    pixel coordinates (the fixed-point fast path must hand those to the exact path or reject them)."""
    case = scenes.small_case(112, dims=(48, 40, 44), voxel_size=0.01, nframes=4, inside=True, hw=(96, 128))
    vol = _gpu_volume(case["vol_bnds"].copy(), 0.01)
    orc = oracle.OracleTSDFVolume(case["vol_bnds"].copy(), 0.01)
    for c, d, p, obs in case["frames"]:
        # put the camera centre exactly on a lattice plane so that some voxels have Z == 0 or ~1e-17
        p = p.copy()
        p[:3, 3] = np.float32(case["vol_bnds"][:, 0]) + 0.01 * np.round((p[:3, 3] - case["vol_bnds"][:, 0]) / 0.01)
        p[:3, :3] = np.eye(3)
        vol.integrate(c, d, case["K"], p, obs_weight=obs)
        n = orc.integrate(c, d, case["K"], p, obs_weight=obs)
        assert vol.stats()["updated_last"] == n
    assert_parity(_state(vol), (orc._tsdf_vol_cpu, orc._weight_vol_cpu, orc._color_vol_cpu))


def test_pipelined_stats_snapshots_match_per_frame_counts():
    """stats_post()/stats_wait(): every frame's counters read without idling the device -- the snapshot
    of frame k taken while later frames are already queued equals the oracle's per-frame update count."""
    from union_thesis_slam_b200 import _capi
    case = scenes.small_case(108, dims=(50, 48, 52), voxel_size=0.02, nframes=20)
    vol = _gpu_volume(case["vol_bnds"].copy(), 0.02)
    orc = oracle.OracleTSDFVolume(case["vol_bnds"].copy(), 0.02)
    tickets, want = [], []
    for c, d, p, obs in case["frames"]:
        vol.integrate(c, d, case["K"], p, obs_weight=obs)
        tickets.append(vol.stats_post())
        want.append(orc.integrate(c, d, case["K"], p, obs_weight=obs))
    got = [vol.stats_wait(t) for t in tickets[-16:]]
    assert [g["updated_last"] for g in got] == want[-16:]
    assert [g["frames"] for g in got] == list(range(5, 21))
    with pytest.raises(_capi.TSDFError):
        vol.stats_wait(tickets[0])  # only the last 16 snapshots are kept
    assert vol.stats()["updated_total"] == sum(want)
    assert_parity(_state(vol), (orc._tsdf_vol_cpu, orc._weight_vol_cpu, orc._color_vol_cpu))
