"""1280 x 720 parity (-m gpu): the hash bench's own image size (BASELINE.json configs[4] shape) on a
room small enough for the per-voxel hash oracle: 8 x 8 x 3 m at 2 cm (400 x 400 x 150 voxels),
8 frames.  Dense grid and hash map against the oracle: weights and block set exact, tsdf <= 1e-5
(the kernels are bit-exact against the dense rule; the reference hash map keeps float64 running
means, hence the tolerance), colour +-1 per channel."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


def _unpack(c):
    c = c.astype(np.int64)
    return np.stack([c >> 16, (c >> 8) & 255, c & 255], -1)


def test_hd_stream_dense_and_hash_match_oracles():
    from union_thesis_slam_b200 import grid_fusion, hash_fusion, synth
    room = synth.room_bounds((8.0, 8.0, 3.0))
    fr, K = synth.room_stream(8, room, width=1280, height=720, seed=5)
    assert fr[0][1].shape == (720, 1280)
    vol = grid_fusion.TSDFVolume(room.copy(), 0.02)
    ht = hash_fusion.HashTable(room.copy(), 0.02, map_size=1 << 17)
    orc = oracle.OracleTSDFVolume(room.copy(), 0.02)
    orh = oracle.OracleHashVox(room.copy(), 0.02)
    for c, d, p in fr:
        vol.integrate(c, d, K, p)
        ht.integrate(c, d, K, p)
        n = orc.integrate(c, d, K, p)
        orh.integrate(c, d, K, p)
        assert vol.stats()["updated_last"] == n == ht.stats()["updated_last"]
    t, c_ = vol.get_volume()
    w = vol._weight_vol_cpu
    assert np.array_equal(w, orc._weight_vol_cpu)
    assert np.array_equal(t, orc._tsdf_vol_cpu) and np.array_equal(c_, orc._color_vol_cpu)
    ht_t, ht_w, ht_c = ht.get_volume_with_weight()
    assert np.array_equal(ht_w, w) and np.array_equal(ht_t, t) and np.array_equal(ht_c, c_), "hash != dense"
    # against the reference hash map restated per voxel (float64 running means)
    rt, rc = orh.get_volume()
    assert np.array_equal(orh.entries.astype(bool), w > 0), "entry set differs"
    assert float(np.abs(ht_t.astype(np.float64) - rt).max()) <= 1e-5
    assert int(np.abs(_unpack(ht_c) - _unpack(rc)).max()) <= 1
    keys, _ = ht.export_blocks()
    got = np.unique(keys.astype(np.int64), axis=0)
    assert np.array_equal(got, oracle.block_set_from_weight(orc._weight_vol_cpu)), "allocated block set differs"
