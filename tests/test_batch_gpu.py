"""Temporal blocking parity (-m gpu): integrating K frames per brick visit (TSDFVolume.set_batch) must
give the same bits as frame-by-frame integration -- the rule has no cross-voxel term and every voxel
applies its frames in call order (reference call order: grid_demo1.py:76-87)."""
import numpy as np
import pytest

import frames
import scenes
from oracle import oracle

pytestmark = pytest.mark.gpu


def _state(vol):
    t, c = vol.get_volume()
    return t, vol._weight_vol_cpu, c


def _assert_state_equal(got, want, what):
    for name, a, b in zip(("tsdf", "weight", "colour"), got, want):
        assert np.array_equal(a, b), "%s: %s differs in %d voxels" % (what, name, int((a != b).sum()))


@pytest.mark.parametrize("batch", [2, 3, 4])
@pytest.mark.parametrize("seed,dims,vs,kw", [
    (201, (40, 40, 40), 0.02, dict(nframes=9)),                                   # 9 frames: partial last batch
    (202, (65, 17, 33), 0.015, dict(obs_weights=(0.25, 1.75, 3.0), nframes=7)),    # non-unit weights inside a batch
    (203, (50, 48, 52), 0.02, dict(inside=True, nframes=6)),
    (204, (31, 31, 29), 0.01, dict(depth_dtype=np.float32, quantise=False, nframes=5)),
])
def test_batched_integration_equals_oracle(batch, seed, dims, vs, kw):
    from union_thesis_slam_b200 import grid_fusion
    case = scenes.small_case(seed, dims=dims, voxel_size=vs, **kw)
    vol = grid_fusion.TSDFVolume(case["vol_bnds"].copy(), vs)
    vol.set_batch(batch)
    orc = oracle.OracleTSDFVolume(case["vol_bnds"].copy(), vs)
    total = 0
    for c, d, p, obs in case["frames"]:
        vol.integrate(c, d, case["K"], p, obs_weight=obs)
        total += orc.integrate(c, d, case["K"], p, obs_weight=obs)
    st = vol.stats()  # applies the pending partial batch
    assert st["updated_total"] == total and st["frames"] == len(case["frames"])
    _assert_state_equal(_state(vol), (orc._tsdf_vol_cpu, orc._weight_vol_cpu, orc._color_vol_cpu), "batch %d" % batch)


def test_batched_config1_sha256_equals_reference():
    """Bundled frames 0..9 at 2 cm in batches of 4 (4 + 4 + 2): sha256 of all three planes equals the
    unmodified reference after 10 frames; read-outs in the middle of a batch see every frame so far."""
    from union_thesis_slam_b200 import grid_fusion
    G = frames.golden()
    K = frames.intrinsics()
    vol = grid_fusion.TSDFVolume(np.array(scenes.CONFIG1_BOUNDS), scenes.CONFIG1_VOXEL)
    vol.set_batch(4)
    for n, i in enumerate(range(10), 1):
        c, d, p = frames.load_frame(i)
        vol.integrate(c, d, K, p, obs_weight=1.)
        if n in (1, 2):  # mid-batch read-out
            g = G["dense"][str(n)]
            t, w, col = _state(vol)
            assert (frames.sha(w), frames.sha(t), frames.sha(col)) == (g["weight"], g["tsdf"], g["color"])
    g = G["dense"]["10"]
    t, w, col = _state(vol)
    assert (frames.sha(w), frames.sha(t), frames.sha(col)) == (g["weight"], g["tsdf"], g["color"])
    assert vol.stats()["updated_total"] == 3557017


def test_batched_room_long_history_and_shards():
    """Synthetic room, 40 frames in batches of 4 on two logical ranks (cyclic x-sheets): union of the
    shards equals the oracle; weights reach ~40, so the converged-colour shortcut runs inside batches."""
    from union_thesis_slam_b200 import grid_fusion, synth
    room = synth.room_bounds((3.0, 2.5, 2.0))
    fr, K = synth.room_stream(10, room, seed=3)
    orc = oracle.OracleTSDFVolume(room.copy(), 0.01)
    vols = [grid_fusion.TSDFVolume(room.copy(), 0.01, rank=r, world_size=2) for r in range(2)]
    for v in vols:
        v.set_batch(4)
    for n in range(40):
        c, d, p = fr[n % 10]
        for v in vols:
            v.integrate(c, d, K, p)
        orc.integrate(c, d, K, p)
    want = (orc._tsdf_vol_cpu, orc._weight_vol_cpu, orc._color_vol_cpu)
    for v in vols:
        xs = np.asarray(v._local_x)
        t, w, c = v._download_local()
        _assert_state_equal((t, w, c), tuple(a[xs] for a in want), "rank %d" % v.rank)


def test_batch_flushes_on_image_size_change():
    from union_thesis_slam_b200 import grid_fusion
    a = scenes.small_case(205, dims=(40, 36, 44), voxel_size=0.02, nframes=3)
    b = scenes.small_case(205, dims=(40, 36, 44), voxel_size=0.02, nframes=3, hw=(120, 160))
    vol = grid_fusion.TSDFVolume(a["vol_bnds"].copy(), 0.02)
    vol.set_batch(4)
    orc = oracle.OracleTSDFVolume(a["vol_bnds"].copy(), 0.02)
    seq = [(a, 0), (a, 1), (b, 0), (b, 1), (a, 2), (b, 2)]
    for case, i in seq:
        c, d, p, obs = case["frames"][i]
        vol.integrate(c, d, case["K"], p, obs_weight=obs)
        orc.integrate(c, d, case["K"], p, obs_weight=obs)
    _assert_state_equal(_state(vol), (orc._tsdf_vol_cpu, orc._weight_vol_cpu, orc._color_vol_cpu), "mixed sizes")
