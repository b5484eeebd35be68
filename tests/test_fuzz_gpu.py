"""Seeded fuzz of both maps against the CPU oracle (-m gpu): many small random volumes with
varying grid sizes, voxel sizes, image sizes, camera placement, obs_weight sequences (integer,
float32-representable and not), depth dtypes.  Everything must be bit-exact."""
import numpy as np
import pytest

import scenes
from oracle import oracle

pytestmark = pytest.mark.gpu


def _spec(seed):
    rng = np.random.default_rng(seed)
    dims = tuple(int(v) for v in rng.integers(5, 60, 3))
    vs = float(rng.choice([0.005, 0.01, 0.0173, 0.02, 0.03, 0.05]))
    hw = (int(rng.integers(20, 130)), int(rng.integers(20, 170)))
    weights = [(1.0,), (1.0, 2.0, 1.0), (0.5, 0.25, 4.0), (0.3, 0.7, 1.1), (1.0, 0.1)][int(rng.integers(0, 5))]
    return dict(seed=seed, dims=dims, voxel_size=vs, hw=hw, nframes=int(rng.integers(2, 6)), obs_weights=weights,
                inside=bool(rng.integers(0, 2)), depth_dtype=[np.float64, np.float32][int(rng.integers(0, 2))],
                quantise=bool(rng.integers(0, 2)))


@pytest.mark.parametrize("block", range(4))
def test_dense_fuzz(block):
    from union_thesis_slam_b200 import grid_fusion
    for seed in range(1000 + 8 * block, 1000 + 8 * block + 8):
        case = scenes.small_case(**_spec(seed))
        vol = grid_fusion.TSDFVolume(case["vol_bnds"].copy(), case["voxel_size"])
        orc = oracle.OracleTSDFVolume(case["vol_bnds"].copy(), case["voxel_size"])
        for c, d, p, obs in case["frames"]:
            vol.integrate(c, d, case["K"], p, obs_weight=obs)
            n = orc.integrate(c, d, case["K"], p, obs_weight=obs)
            assert vol.stats()["updated_last"] == n, "seed %d" % seed
        t, col = vol.get_volume()
        assert np.array_equal(vol._weight_vol_cpu, orc._weight_vol_cpu), "seed %d" % seed
        assert np.array_equal(t, orc._tsdf_vol_cpu), "seed %d" % seed
        assert np.array_equal(col, orc._color_vol_cpu), "seed %d" % seed
        vol.close()


@pytest.mark.parametrize("block", range(2))
def test_hash_fuzz(block):
    from union_thesis_slam_b200 import hash_fusion
    for seed in range(2000 + 6 * block, 2000 + 6 * block + 6):
        spec = _spec(seed)
        case = scenes.small_case(**spec)
        ht = hash_fusion.HashTable(case["vol_bnds"].copy(), case["voxel_size"], map_size=1 << 14)
        orc = oracle.OracleTSDFVolume(case["vol_bnds"].copy(), case["voxel_size"])
        for c, d, p, obs in case["frames"]:
            ht.integrate(c, d, case["K"], p, obs_weight=obs)
            n = orc.integrate(c, d, case["K"], p, obs_weight=obs)
            assert ht.stats()["updated_last"] == n, "seed %d" % seed
        t, w, col = ht.get_volume_with_weight()
        assert np.array_equal(w, orc._weight_vol_cpu) and np.array_equal(t, orc._tsdf_vol_cpu), "seed %d" % seed
        assert np.array_equal(col, orc._color_vol_cpu), "seed %d" % seed
        keys, _ = ht.export_blocks()
        got = np.unique(keys.astype(np.int64), axis=0) if len(keys) else np.zeros((0, 3), np.int64)
        assert np.array_equal(got, oracle.block_set_from_weight(orc._weight_vol_cpu)), "seed %d" % seed
        ht.close()
