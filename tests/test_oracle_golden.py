"""Pins the CPU oracle (oracle/tsdf_oracle.c) to the unmodified reference.

Goldens were produced by tests/golden/make_golden.py running /root/reference/grid_fusion.py and
hash_fusion.py (numba CPU path) in the build container.  CPU-only tests (-m "not gpu").
"""
import glob
import os

import numpy as np
import pytest

import frames
import scenes
from oracle import oracle

G = frames.golden()


def _load_small(path):
    z = np.load(path)
    fr = [(z["color"][i], z["depth"][i], z["pose"][i], float(z["obs"][i])) for i in range(len(z["obs"]))]
    return z, fr


@pytest.fixture(scope="module")
def dense10():
    """C oracle over fixture frames 0..9 at config 1, with checkpoints after 1, 2 and 10 frames."""
    K = frames.intrinsics()
    vol = oracle.OracleTSDFVolume(np.array(scenes.CONFIG1_BOUNDS), scenes.CONFIG1_VOXEL)
    snaps = {}
    for n, i in enumerate(range(10), 1):
        c, d, p = frames.load_frame(i)
        vol.integrate(c, d, K, p, obs_weight=1.)
        if str(n) in G["dense"]:
            snaps[n] = (frames.sha(vol._tsdf_vol_cpu), frames.sha(vol._weight_vol_cpu),
                        frames.sha(vol._color_vol_cpu), int((vol._weight_vol_cpu > 0).sum()),
                        int(vol._weight_vol_cpu.astype(np.float64).sum()),
                        len(oracle.block_set_from_weight(vol._weight_vol_cpu)))
    return vol, snaps


def test_config1_geometry():
    bnds = np.array(scenes.CONFIG1_BOUNDS)
    dim, origin = oracle.volume_geometry(bnds, scenes.CONFIG1_VOXEL)
    assert list(dim) == G["config1_dim"] == [405, 264, 289]
    assert origin.dtype == np.float32
    # the reference snaps the upper bounds in the caller's array (grid_fusion.py:43)
    assert np.allclose(bnds[:, 1], bnds[:, 0] + dim * scenes.CONFIG1_VOXEL)


@pytest.mark.parametrize("n", [1, 2, 10])
def test_dense_oracle_matches_reference_sha(dense10, n):
    _, snaps = dense10
    g = G["dense"][str(n)]
    t, w, c, npos, sumw, nblk = snaps[n]
    assert (t, w, c) == (g["tsdf"], g["weight"], g["color"])
    assert (npos, sumw, nblk) == (g["n_w_pos"], g["sum_w"], g["n_blocks"])


def test_dense_oracle_matches_pstats_counts(dense10):
    """Counts frozen in the reference's committed cProfile dumps."""
    vol, snaps = dense10
    pc = G["pstats_counts"]
    assert snaps[1][3] == pc["voxel_integrate_frame0"] == 354383
    assert snaps[10][3] == pc["add_hash_entry_10frames"] == 389590
    assert snaps[10][4] == pc["get_hash_entry_10frames"] == 3557017
    assert sum(vol.updates_per_frame) == 3557017


def test_hashvox_oracle_matches_reference_sha():
    K = frames.intrinsics()
    hv = oracle.OracleHashVox(np.array(scenes.CONFIG1_BOUNDS), scenes.CONFIG1_VOXEL)
    for n, i in enumerate(range(3), 1):
        c, d, p = frames.load_frame(i)
        hv.integrate(c, d, K, p, obs_weight=1.)
        g = G["hash"][str(n)]
        t, col = hv.get_volume()
        assert frames.sha(t) == g["tsdf"] and frames.sha(col) == g["color"]
        assert int(hv.entries.sum()) == g["entries"]
        assert len(hv.block_set()) == g["n_blocks"]
    # after frame 0 the reference's two maps agree exactly; later the f64 running
    # mean of voxel.py drifts from the dense f32 state by ~1e-7
    assert G["hash"]["1"]["tsdf"] == G["dense"]["1"]["tsdf"]
    assert G["hash"]["2"]["tsdf"] != G["dense"]["2"]["tsdf"]
    assert G["hash"]["2"]["color"] == G["dense"]["2"]["color"]


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(frames.GOLDEN_DIR, "small_dense_*.npz"))))
def test_small_dense_cases(path):
    z, fr = _load_small(path)
    vol = oracle.OracleTSDFVolume(z["vol_bnds"].copy(), float(z["voxel_size"]))
    for c, d, p, obs in fr:
        vol.integrate(c, d, z["K"], p, obs_weight=obs)
    assert np.array_equal(vol._tsdf_vol_cpu, z["tsdf"])
    assert np.array_equal(vol._weight_vol_cpu, z["weight"])
    assert np.array_equal(vol._color_vol_cpu, z["color_vol"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(frames.GOLDEN_DIR, "small_hash_*.npz"))))
def test_small_hash_cases(path):
    z, fr = _load_small(path)
    hv = oracle.OracleHashVox(z["vol_bnds"].copy(), float(z["voxel_size"]))
    for c, d, p, obs in fr:
        hv.integrate(c, d, z["K"], p, obs_weight=obs)
    t, col = hv.get_volume()
    assert np.array_equal(t, z["tsdf"]) and np.array_equal(col, z["color_vol"])
    assert np.array_equal(np.argwhere(hv.entries), z["entries"])


def test_small_scene_generator_is_frozen():
    """The committed npz inputs are what scenes.small_case() still generates."""
    for spec in scenes.GOLDEN_SMALL_DENSE:
        case = scenes.small_case(**spec)
        z = np.load(os.path.join(frames.GOLDEN_DIR, "small_dense_%d.npz" % spec["seed"]))
        assert np.array_equal(case["vol_bnds"], z["vol_bnds"])
        assert np.array_equal(case["frames"][0][1], z["depth"][0])
        assert np.array_equal(case["frames"][-1][2], z["pose"][-1])


def test_hash_function_kats():
    for k in G["hash_kats"]:
        assert oracle.hash_function(k["coord"], k["table_size"]) == k["hash"]
    assert oracle.hash_function([333, 234, 241], 100000) == 53356
    assert oracle.hash_function([1, 2, 3], 1000000) == 163698
    assert oracle.hash_function([404, 263, 288], 1000000) == 187815


def test_slab_and_cyclic_shards_union_equals_single():
    """Multi-rank correctness oracle: per-voxel independence => union(shards) == single run."""
    z, fr = _load_small(os.path.join(frames.GOLDEN_DIR, "small_dense_12.npz"))
    X = z["tsdf"].shape[0]
    for xs_parts in ([np.arange(0, X // 2), np.arange(X // 2, X)],
                     [np.arange(X)[(np.arange(X) // 8) % 3 == r] for r in range(3)]):
        t = np.empty_like(z["tsdf"]); w = np.empty_like(t); col = np.empty_like(t)
        for xs in xs_parts:
            vol = oracle.OracleTSDFVolume(z["vol_bnds"].copy(), float(z["voxel_size"]), xs=xs)
            for c, d, p, obs in fr:
                vol.integrate(c, d, z["K"], p, obs_weight=obs)
            t[xs], w[xs], col[xs] = vol._tsdf_vol_cpu, vol._weight_vol_cpu, vol._color_vol_cpu
        assert np.array_equal(t, z["tsdf"]) and np.array_equal(w, z["weight"]) and np.array_equal(col, z["color_vol"])


def test_oracle_first_100_frames_equal_reference_and_survey_appendix_b():
    """C oracle over the first 100 bundled frames against the reference's digests (golden_long.json), which the
    generating script also checked against SURVEY.md Appendix B (rows 50 and 100)."""
    G = frames.golden_long()
    K = frames.intrinsics()
    orc = oracle.OracleTSDFVolume(np.array(scenes.CONFIG1_BOUNDS), scenes.CONFIG1_VOXEL)
    for n, i in enumerate(range(100), 1):
        c, d, p = frames.load_frame(i)
        orc.integrate(c, d, K, p, obs_weight=1.)
        if str(n) in G["dense"]:
            g = G["dense"][str(n)]
            assert (frames.sha(orc._weight_vol_cpu), frames.sha(orc._tsdf_vol_cpu), frames.sha(orc._color_vol_cpu)) == (
                g["weight"], g["tsdf"], g["color"]), "after %d frames" % n
    assert G["dense"]["50"]["tsdf"].startswith("2e9bb509beb8d38d") and G["dense"]["100"]["color"].startswith("fe43593d4f315064")
