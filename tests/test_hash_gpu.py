"""GPU parity of the voxel-block hash map (-m gpu).

Bars (BASELINE.json north_star): allocated block set bit-exact as a set, voxel weights bit-exact,
|tsdf| <= 1e-5 against the reference hash map (float64 running mean, voxel.py:32-37), colour +-1.
Against the dense GPU grid the hash output must be identical (same arithmetic).
"""
import glob
import os

import numpy as np
import pytest

import frames
import scenes
from oracle import oracle

pytestmark = pytest.mark.gpu


def _table(bnds, vs, **kw):
    from union_thesis_slam_b200 import hash_fusion
    kw.setdefault("map_size", 1 << 16)
    return hash_fusion.HashTable(np.array(bnds, dtype=np.float64), vs, **kw)


def _dense(bnds, vs, **kw):
    from union_thesis_slam_b200 import grid_fusion
    return grid_fusion.TSDFVolume(np.array(bnds, dtype=np.float64), vs, **kw)


def _unpack(c):
    c = c.astype(np.int64)
    return np.stack([c >> 16, (c >> 8) & 255, c & 255], -1)


def _block_set(keys):
    return np.unique(np.asarray(keys, dtype=np.int64).reshape(-1, 3), axis=0)


def check_against_hashvox(ht, hv, strict_color=True):
    """ht: GPU HashTable, hv: OracleHashVox (the reference map as a voxel set)."""
    t, w, c = ht.get_volume_with_weight()
    rt, rc = hv.get_volume()
    assert np.array_equal(w, hv.weight.astype(np.float32)), "weights differ"
    assert float(np.abs(t.astype(np.float64) - rt).max()) <= 1e-5
    assert int(np.abs(_unpack(c) - _unpack(rc)).max()) <= 1
    if strict_color:
        assert np.array_equal(c, rc)
    keys, payload = ht.export_blocks()
    assert np.array_equal(_block_set(keys), hv.block_set()), "allocated block set differs from the reference's"
    assert len(keys) == len(_block_set(keys)) == ht.stats()["blocks"], "duplicate or missing blocks in the export"
    return t, w, c


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(frames.GOLDEN_DIR, "small_hash_*.npz"))))
def test_small_golden_cases_match_reference(path):
    z = np.load(path)
    ht = _table(z["vol_bnds"], float(z["voxel_size"]))
    for i in range(len(z["obs"])):
        ht.integrate(z["color"][i], z["depth"][i], z["K"], z["pose"][i], obs_weight=1.0)
    t, w, c = ht.get_volume_with_weight()
    assert float(np.abs(t.astype(np.float64) - z["tsdf"]).max()) <= 1e-5
    assert np.array_equal(c, z["color_vol"])
    # entry set of the reference container, block-wise
    assert np.array_equal(np.argwhere(w > 0), z["entries"])
    keys, _ = ht.export_blocks()
    assert np.array_equal(_block_set(keys), np.unique(z["entries"] // 8, axis=0))


def test_config2_three_frames():
    """Bundled sequence into the hash map at 2 cm (BASELINE.json configs[1]); dense-vs-hash compared."""
    G = frames.golden()
    K = frames.intrinsics()
    ht = _table(scenes.CONFIG1_BOUNDS, scenes.CONFIG1_VOXEL, map_size=1 << 17)
    dn = _dense(scenes.CONFIG1_BOUNDS, scenes.CONFIG1_VOXEL)
    hv = oracle.OracleHashVox(np.array(scenes.CONFIG1_BOUNDS), scenes.CONFIG1_VOXEL)
    for n, i in enumerate(range(3), 1):
        c, d, p = frames.load_frame(i)
        ht.integrate(c, d, K, p, obs_weight=1.)
        dn.integrate(c, d, K, p, obs_weight=1.)
        hv.integrate(c, d, K, p, obs_weight=1.)
        g = G["hash"][str(n)]
        st = ht.stats()
        assert st["blocks"] == g["n_blocks"]
        assert st["updated_last"] == hv.updates_per_frame[-1]
        if n == 1:
            t, c_ = ht.get_volume()
            # after frame 0 the reference's hash and dense maps are identical
            assert frames.sha(t) == g["tsdf"] and frames.sha(c_) == g["color"]
    t, w, c_ = check_against_hashvox(ht, hv)
    assert int((w > 0).sum()) == G["hash"]["3"]["entries"]
    assert frames.sha(c_) == G["hash"]["3"]["color"]
    # dense-vs-hash: identical arithmetic => identical volumes
    dt, dc = dn.get_volume()
    assert np.array_equal(t, dt) and np.array_equal(c_, dc) and np.array_equal(w, dn._weight_vol_cpu)
    st = ht.stats()
    assert st["table_keys"] >= st["blocks"] and st["heap_free"] == st["heap_capacity"] - st["blocks"]


def test_config2_ten_frames_equals_dense_reference_sha256():
    """Ten bundled frames into the hash map at 2 cm: since hash and dense share their arithmetic, the
    scattered volumes must hash to the DENSE reference digests; the block count is the reference's."""
    G = frames.golden()
    K = frames.intrinsics()
    ht = _table(scenes.CONFIG1_BOUNDS, scenes.CONFIG1_VOXEL, map_size=1 << 17)
    total = 0
    for n, i in enumerate(range(10), 1):
        c, d, p = frames.load_frame(i)
        ht.integrate(c, d, K, p, obs_weight=1.)
        total += ht.stats()["updated_last"]
    g = G["dense"]["10"]
    t, w, c_ = ht.get_volume_with_weight()
    assert frames.sha(w) == g["weight"] and frames.sha(t) == g["tsdf"] and frames.sha(c_) == g["color"]
    assert total == g["sum_w"] == 3557017
    st = ht.stats()
    assert st["blocks"] == g["n_blocks"] == 1511 and st["updated_total"] == 3557017


@pytest.mark.parametrize("seed,dims,vs,kw", [
    (201, (40, 40, 40), 0.02, {}),
    (202, (7, 9, 5), 0.04, {}),
    (203, (65, 17, 33), 0.015, dict(nframes=4)),
    (204, (50, 48, 52), 0.02, dict(inside=True, nframes=5)),
    (205, (96, 20, 24), 0.03, dict(hw=(120, 160), nframes=4)),
])
def test_random_volumes_match_reference_map(seed, dims, vs, kw):
    case = scenes.small_case(seed, dims=dims, voxel_size=vs, **kw)
    ht = _table(case["vol_bnds"].copy(), vs)
    hv = oracle.OracleHashVox(case["vol_bnds"].copy(), vs)
    dn = _dense(case["vol_bnds"].copy(), vs)
    for c, d, p, _ in case["frames"]:
        ht.integrate(c, d, case["K"], p, obs_weight=1.)
        dn.integrate(c, d, case["K"], p, obs_weight=1.)
        n = hv.integrate(c, d, case["K"], p)
        assert ht.stats()["updated_last"] == n
    t, w, c_ = check_against_hashvox(ht, hv)
    dt, dc = dn.get_volume()
    assert np.array_equal(t, dt) and np.array_equal(c_, dc) and np.array_equal(w, dn._weight_vol_cpu)


def test_synthetic_room_hash_equals_dense_and_oracle():
    from union_thesis_slam_b200 import synth
    room = synth.room_bounds((3.0, 2.5, 2.0))
    fr, K = synth.room_stream(4, room, seed=5)
    ht = _table(room.copy(), 0.02)
    hv = oracle.OracleHashVox(room.copy(), 0.02)
    for c, d, p in fr:
        ht.integrate(c, d, K, p)
        hv.integrate(c, d, K, p)
    check_against_hashvox(ht, hv)


def test_obs_weight_follows_the_dense_rule():
    case = scenes.small_case(206, dims=(33, 30, 35), voxel_size=0.02, nframes=4, obs_weights=(0.5, 2.0, 1.0, 0.3))
    ht = _table(case["vol_bnds"].copy(), 0.02)
    orc = oracle.OracleTSDFVolume(case["vol_bnds"].copy(), 0.02)
    for c, d, p, obs in case["frames"]:
        ht.integrate(c, d, case["K"], p, obs_weight=obs)
        orc.integrate(c, d, case["K"], p, obs_weight=obs)
    t, w, c_ = ht.get_volume_with_weight()
    assert np.array_equal(w, orc._weight_vol_cpu) and np.array_equal(t, orc._tsdf_vol_cpu)
    assert np.array_equal(c_, orc._color_vol_cpu)


def test_export_payload_matches_dense_scatter():
    case = scenes.small_case(207, dims=(30, 26, 41), voxel_size=0.02, nframes=2)
    ht = _table(case["vol_bnds"].copy(), 0.02)
    for c, d, p, _ in case["frames"]:
        ht.integrate(c, d, case["K"], p)
    t, w, c_ = ht.get_volume_with_weight()
    keys, payload = ht.export_blocks()
    assert len(keys) > 0
    X, Y, Z = t.shape
    for k, blk in zip(keys[:50], payload[:50]):
        x0, y0, z0 = (int(v) * 8 for v in k)
        sx, sy, sz = min(8, X - x0), min(8, Y - y0), min(8, Z - z0)
        assert np.array_equal(blk[0, :sx, :sy, :sz], t[x0:x0 + sx, y0:y0 + sy, z0:z0 + sz])
        assert np.array_equal(blk[1, :sx, :sy, :sz], w[x0:x0 + sx, y0:y0 + sy, z0:z0 + sz])
        assert np.array_equal(blk[2, :sx, :sy, :sz], c_[x0:x0 + sx, y0:y0 + sy, z0:z0 + sz])
        assert blk[1].max() > 0, "a block without any updated voxel owns storage"


def test_hash_shards_partition_the_block_set():
    """Logical ranks run one after another on one GPU: blocks are owned by exactly one rank, the
    owner matches dist.block_owner, and the merged volumes equal the single-rank result."""
    from union_thesis_slam_b200 import dist
    case = scenes.small_case(208, dims=(64, 40, 48), voxel_size=0.02, nframes=3)
    single = _table(case["vol_bnds"].copy(), 0.02)
    for c, d, p, _ in case["frames"]:
        single.integrate(c, d, case["K"], p)
    st, sw, sc = single.get_volume_with_weight()
    skeys = _block_set(single.export_blocks()[0])
    for world in (2, 3, 8):
        allkeys = []
        t = np.ones_like(st); w = np.zeros_like(sw); col = np.zeros_like(sc)
        for r in range(world):
            ht = _table(case["vol_bnds"].copy(), 0.02, rank=r, world_size=world)
            for c, d, p, _ in case["frames"]:
                ht.integrate(c, d, case["K"], p)
            keys, _ = ht.export_blocks()
            assert np.all(dist.block_owner(keys, world) == r)
            allkeys.append(keys)
            import ctypes as C
            from union_thesis_slam_b200 import _capi
            lt = np.empty_like(st); lw = np.empty_like(sw); lc = np.empty_like(sc)
            _capi.check(ht._lib.vh_to_dense(ht._handle, _capi.ptr(lt), _capi.ptr(lw), _capi.ptr(lc)))
            m = lw > 0
            assert not (m & (w > 0)).any(), "two ranks updated the same voxel"
            t[m], w[m], col[m] = lt[m], lw[m], lc[m]
            ht.close()
        cat = np.concatenate(allkeys)
        assert len(cat) == len(_block_set(cat)), "a block is owned by two ranks"
        assert np.array_equal(_block_set(cat), skeys)
        assert np.array_equal(t, st) and np.array_equal(w, sw) and np.array_equal(col, sc)
        counts = [len(k) for k in allkeys]
        assert max(counts) <= 2.0 * (sum(counts) / world) + 8, "owner hash is badly unbalanced: %s" % counts


def test_heap_exhaustion_and_table_full_are_reported():
    from union_thesis_slam_b200 import _capi
    case = scenes.small_case(209, dims=(48, 48, 48), voxel_size=0.02, nframes=1)
    c, d, p, _ = case["frames"][0]
    ht = _table(case["vol_bnds"].copy(), 0.02, heap_blocks=5)
    ht.integrate(c, d, case["K"], p)
    with pytest.raises(_capi.TSDFError) as ei:
        ht.stats()
    assert ei.value.status == -5
    # a big volume at a fine voxel size, tiny table: more keys than slots
    case = scenes.small_case(210, dims=(160, 160, 160), voxel_size=0.01, nframes=1, hw=(96, 128))
    c, d, p, _ = case["frames"][0]
    ht = _table(case["vol_bnds"].copy(), 0.01, map_size=1024, heap_blocks=8000)
    ht.integrate(c, d, case["K"], p)
    with pytest.raises(_capi.TSDFError) as ei:
        ht.sync()
    assert ei.value.status == -6


def test_remove_blocks_returns_storage_to_the_heap():
    case = scenes.small_case(211, dims=(40, 40, 40), voxel_size=0.02, nframes=2)
    ht = _table(case["vol_bnds"].copy(), 0.02)
    c, d, p, _ = case["frames"][0]
    ht.integrate(c, d, case["K"], p)
    keys, _ = ht.export_blocks()
    n0 = ht.stats()["blocks"]
    assert n0 == len(keys) > 4
    victims = np.concatenate([keys[:3], keys[:2], [[9999, 0, 0]]])  # duplicates and an unknown key
    assert ht.remove_blocks(victims) == 3
    s = ht.stats()
    assert s["blocks"] == n0 - 3 and s["heap_free"] == s["heap_capacity"] - s["blocks"]
    left = _block_set(ht.export_blocks()[0])
    assert len(left) == n0 - 3
    # the same frame again re-creates the removed blocks from scratch (weight 1) next to weight-2 blocks
    ht.integrate(c, d, case["K"], p)
    assert ht.stats()["blocks"] == n0
    _, w, _ = ht.get_volume_with_weight()
    assert set(np.unique(w)) <= {0.0, 1.0, 2.0} and (w == 1.0).any() and (w == 2.0).any()


def test_hash_function_matches_reference_kats():
    G = frames.golden()
    case = scenes.small_case(212, dims=(16, 16, 16), voxel_size=0.02, nframes=1)
    ht = _table(case["vol_bnds"].copy(), 0.02, map_size=65536)
    assert ht._table_size == 65536
    for k in G["hash_kats"]:
        if k["table_size"] == 65536:
            assert ht.hash_function(k["coord"]) == k["hash"]
    assert ht.hash_function([0, 0, 0]) == 0
    assert ht.get_load_factor() == 0.0 and ht.count_num_hash_entries() == 0


def test_allocation_front_ends_agree(monkeypatch):
    """TSDF_HASH_ALLOC=walk (cone walk per pixel tile, then refine) and =grid (refine tests the
    bounded super-block grid itself) must allocate the same blocks and produce the same voxels."""
    from union_thesis_slam_b200 import hash_fusion
    case = scenes.small_case(seed=77, dims=(70, 45, 52), voxel_size=0.02, hw=(96, 128), nframes=4)
    out = {}
    for mode in ("walk", "grid"):
        monkeypatch.setenv("TSDF_HASH_ALLOC", mode)
        ht = hash_fusion.HashTable(case["vol_bnds"].copy(), case["voxel_size"], map_size=1 << 14)
        for c, d, p, obs in case["frames"]:
            ht.integrate(c, d, case["K"], p, obs_weight=obs)
        keys, payload = ht.export_blocks()
        order = np.lexsort(keys.T[::-1])
        out[mode] = (keys[order], payload[order], ht.stats()["updated_total"])
        ht.close()
    assert out["walk"][2] == out["grid"][2] > 0
    assert np.array_equal(out["walk"][0], out["grid"][0]) and np.array_equal(out["walk"][1], out["grid"][1])


def test_integrate_and_per_voxel_calls_see_one_map():
    """Reference hash_fusion.py:285-320, 147-154: a voxel inserted by HashTable.integrate is found by
    get_hash_entry / get_voxel, and count_num_hash_entries counts voxel entries.  Frame 0 of the
    bundled sequence at 2 cm: the reference's own count is 354 383 (its committed pstats dump)."""
    from union_thesis_slam_b200 import hash_fusion
    K = frames.intrinsics()
    c, d, p = frames.load_frame(0)
    ht = hash_fusion.HashTable(np.array(scenes.CONFIG1_BOUNDS), scenes.CONFIG1_VOXEL, map_size=1 << 16)
    orc = oracle.OracleHashVox(np.array(scenes.CONFIG1_BOUNDS), scenes.CONFIG1_VOXEL)
    ht.integrate(c, d, K, p)
    orc.integrate(c, d, K, p)
    lf0, ne0 = ht.get_load_factor(), ht.get_num_non_empty_bucket()
    assert ht.count_num_hash_entries() == 354383 == int(orc.entries.sum())
    rng = np.random.default_rng(9)
    hit = np.argwhere(orc.entries)
    miss = np.argwhere(orc.entries == 0)
    hit = hit[rng.choice(len(hit), 1000, replace=False)]
    miss = miss[rng.choice(len(miss), 1000, replace=False)]
    found, sdf, wgt, col = ht.query_voxels(np.concatenate([hit, miss]))
    assert found[:1000].all() and not found[1000:].any()
    i = tuple(hit.T)
    assert np.array_equal(wgt[:1000], orc.weight[i].astype(np.float32))
    assert float(np.abs(sdf[:1000].astype(np.float64) - orc.sdf[i]).max()) <= 1e-5
    assert np.array_equal(col[:1000], orc.color[i].astype(np.float32))
    # object API, one voxel at a time (snapshots of the device map)
    for v in hit[:5]:
        e = ht.get_hash_entry(v)
        assert e is not None and list(e.get_position()) == [int(x) for x in v]
        vox = ht.get_voxel(v)
        assert vox.get_weight() == orc.weight[tuple(v)] and abs(vox.get_sdf() - orc.sdf[tuple(v)]) <= 1e-5
        assert vox.get_color() == orc.color[tuple(v)]
    for v in miss[:5]:
        assert ht.get_hash_entry(v) is None and ht.get_voxel(v) is None
    assert ht.get_hash_entry([-3, 0, 0]) is None and ht.get_hash_entry([10 ** 6, 0, 0]) is None
    # look-ups neither create the hand-fed container nor change what the metrics report
    assert ht._voxels is None and (ht.get_load_factor(), ht.get_num_non_empty_bucket()) == (lf0, ne0)
    # remove (hash_fusion.py:322-329) on a voxel written by integrate
    ht.remove(hit[0])
    assert ht.get_hash_entry(hit[0]) is None and ht.count_num_hash_entries() == 354382
    t, w, _ = ht.get_volume_with_weight()
    assert t[tuple(hit[0])] == 1.0 and w[tuple(hit[0])] == 0.0
    # hand-added entries live beside the integrated ones and are counted with them
    from union_thesis_slam_b200 import bucket_table
    ht.add_hash_entry(bucket_table.HashEntry([1, 2, 3], None, bucket_table.Voxel()))
    assert ht.get_hash_entry([1, 2, 3]) is not None and ht.count_num_hash_entries() == 354383
    assert ht.get_hash_entry(hit[1]) is not None
