"""Device-resident per-voxel container (csrc/bucket_kernels.cu, vbt_* C ABI) -- GPU parity tests.

Shaped like the reference's tests/hash_map_test.py (same bounds literal, coordinate lists, op
patterns and assertions, driven through the drop-in ``HashTable``), plus replays of the fixtures
recorded from the unmodified reference container (tests/golden/bucket_golden.json): every return
value and the final slot layout -- buckets, slots, chain offsets -- must match, and the float64
voxel payloads of the per-voxel integrate loop must be bit-identical."""
import json
import os

import numpy as np
import pytest

from oracle import bucket_oracle as bo

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "bucket_golden.json")) as fh:
    GOLD = json.load(fh)
BOUNDS = GOLD["bounds"]


def _table(map_size):
    from union_thesis_slam_b200 import hash_fusion
    return hash_fusion.HashTable(np.array(BOUNDS), 0.02, map_size, False)


def _replay(vt, case):
    """Feeds the case's op stream to the device in maximal batches (resizes split them)."""
    res, batch = [], []

    def flush():
        if not batch:
            return
        ops = np.array([b[0] for b in batch], np.int32)
        pos = np.array([b[1:4] for b in batch], np.int64)
        args = np.zeros((len(batch), 5))
        args[:, 2] = np.nan
        for i, b in enumerate(batch):
            if b[0] == bo.OP_INTEGRATE:
                args[i, 0], args[i, 1] = b[4], b[5]
        out = vt.apply(ops, pos, args)
        res.extend([int(r[0]), int(r[1])] for r in out)
        batch.clear()

    for op in case["ops"]:
        if op[0] == "resize":
            flush()
            vt.double_table_size()
            res.append([vt.stats()["table_size"], -1])
        else:
            batch.append(op)
    flush()
    return res


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_device_container_replays_reference(case):
    from union_thesis_slam_b200 import bucket_table
    vt = bucket_table.VoxelBucketTable(case["map_size"])
    assert _replay(vt, case) == case["results"]
    rows, vox = vt.layout(with_voxels=True)
    assert rows.tolist() == case["layout"]
    st, ref = vt.stats(), case["stats"]
    assert (st["table_size"], st["entries"], st["non_empty_buckets"], st["collisions"]) == (
        ref["table_size"], ref["entries"], ref["non_empty"], ref["collisions"])
    assert vt.get_load_factor() == ref["load_factor"]
    if "voxels_hex" in case:
        assert [[float(x).hex() for x in v] for v in vox] == case["voxels_hex"]
    vt.close()


def test_one_op_at_a_time_equals_batch():
    """The reference API is one call per entry; it must land entries exactly where a batch does."""
    from union_thesis_slam_b200 import bucket_table
    case = next(c for c in GOLD["cases"] if c["name"] == "random_seed12_size16")
    vt = bucket_table.VoxelBucketTable(case["map_size"])
    res = []
    for op in case["ops"]:
        e = bucket_table.HashEntry(op[1:4], None, None)
        if op[0] == bo.OP_ADD:
            res.append(list(vt.add_hash_entry(e)))
        elif op[0] == bo.OP_GET:
            got = vt.get_hash_entry(op[1:4])
            res.append([-1, -1] if got is None else list(got.slot))
        else:
            res.append([vt.remove_hash_entry(e), -1])
    assert res == case["results"]
    assert vt.layout().tolist() == case["layout"]
    vt.close()


# ---- the reference's tests/hash_map_test.py, through the drop-in HashTable --------------------
def _coords(name):
    case = next(c for c in GOLD["cases"] if c["name"] == name)
    return [op[1:4] for op in case["ops"] if op[0] == bo.OP_ADD]


def test_resize_maintain_num_entries():
    from union_thesis_slam_b200.bucket_table import HashEntry
    hash_table = _table(10)
    world_coords = _coords("test_resize_maintain_num_entries")
    for position in world_coords:
        hash_table.add_hash_entry(HashEntry(position, None, None))
    hash_table.double_table_size()
    assert hash_table.count_num_hash_entries() == 40
    for position in world_coords:
        assert hash_table.get_hash_entry(position) is not None
    hash_table.close()


def test_add_until_full_size_10():
    from union_thesis_slam_b200.bucket_table import HashEntry
    hash_table = _table(10)
    for position in _coords("test_add_until_full_size_10"):
        hash_table.hash_function(position)
        hash_table.add_hash_entry(HashEntry(position, None, None))
    assert hash_table.count_num_hash_entries() == 40
    hash_table.close()


def test_remove_all_entries_full_size_10():
    from union_thesis_slam_b200.bucket_table import HashEntry
    hash_table = _table(10)
    world_coords = _coords("test_add")
    for position in world_coords:
        hash_table.add_hash_entry(HashEntry(position, None, None))
    assert hash_table.count_num_hash_entries() == 40
    for position in world_coords:
        hash_table.remove_hash_entry(HashEntry(position, None, None))
    assert hash_table.get_num_non_empty_buckets() == 0
    assert hash_table.count_num_hash_entries() == 0
    hash_table.close()


def test_add_until_full_size_1000():
    hash_map = _table(1000)
    rng = np.random.default_rng(0)
    pos = np.stack([rng.integers(0, 500, 4500), rng.integers(0, 500, 4500), np.arange(4500)], 1)
    hash_map.voxel_table.add_positions(pos)
    assert hash_map.count_num_hash_entries() == 4500
    hash_map.close()


def test_add_same_hash_value():
    hash_map = _table(1000)
    slots = hash_map.voxel_table.add_positions(np.zeros((4000, 3), np.int64))
    assert hash_map.count_num_hash_entries() == 4000          # identical entries are all stored
    assert len({tuple(s) for s in slots.tolist()}) == 4000     # each in its own slot
    orc = bo.OracleBucketTable(1000)
    assert [list(orc.add([0, 0, 0])) for _ in range(4000)] == slots.tolist()
    hash_map.close()


def test_general():
    """hash_map_test.py:89-121 at full size (40 000 adds, 20 000 removals, look-ups, 20 000 adds),
    checked against the oracle op by op."""
    n = 10000
    hash_table = _table(n)
    vt = hash_table.voxel_table
    rng = np.random.default_rng(1)
    coords = np.stack([rng.integers(0, 500, 4 * n), rng.integers(0, 500, 4 * n), np.arange(4 * n)], 1)
    orc = bo.OracleBucketTable(n)
    assert vt.add_positions(coords).tolist() == [list(orc.add(c)) for c in coords]
    assert hash_table.count_num_hash_entries() == 4 * n
    drop = rng.permutation(4 * n)[:2 * n]
    out = vt.apply(np.full(2 * n, bo.OP_REMOVE, np.int32), coords[drop])
    assert out[:, 0].tolist() == [orc.remove(c) for c in coords[drop]] and out[:, 0].sum() == 2 * n
    assert hash_table.count_num_hash_entries() == 2 * n
    keep = np.setdiff1d(np.arange(4 * n), drop)
    found = vt.lookup(coords[keep])
    assert (found[:, 0] >= 0).all() and found.tolist() == [list(orc.find(c)) for c in coords[keep]]
    assert (vt.lookup(coords[drop])[:, 0] == -1).all()
    more = np.stack([rng.integers(0, 500, 2 * n), rng.integers(0, 500, 2 * n) + 501, np.arange(2 * n)], 1)
    assert vt.add_positions(more).tolist() == [list(orc.add(c)) for c in more]
    assert hash_table.count_num_hash_entries() == 4 * n
    assert vt.layout().tolist() == [list(r) for r in orc.layout()]
    st = vt.stats()
    assert (st["table_size"], st["non_empty_buckets"], st["collisions"]) == (orc.S, orc.nonempty, orc.get_num_collisions())
    hash_table.close()


def test_integrate_voxels_matches_oracle_running_mean():
    """Per-voxel loop of hash_fusion.py:135-145 + voxel.py:19-49 on a random stream with resizes."""
    from union_thesis_slam_b200 import bucket_table
    rng = np.random.default_rng(5)
    n = 6000
    pos = rng.integers(0, 12, (n, 3))
    dist = np.minimum(1.0, rng.uniform(-1.0, 1.5, n))
    col = (rng.integers(0, 256, n) * 65536 + rng.integers(0, 256, n) * 256 + rng.integers(0, 256, n)).astype(np.float64)
    vt = bucket_table.VoxelBucketTable(64)
    orc = bo.OracleBucketTable(64)
    got = vt.integrate_voxels(pos, dist, col)
    assert got.tolist() == [list(orc.integrate(p, d, c)) for p, d, c in zip(pos, dist, col)]
    rows, vox = vt.layout(with_voxels=True)
    assert rows.tolist() == [list(r) for r in orc.layout()]
    want = np.array([orc.vox[r[0] * orc.B + r[1]] for r in rows])
    assert np.array_equal(vox, want)
    assert vt.stats()["resizes"] == orc.resizes and orc.resizes >= 2
    v = vt.get_voxel(pos[0])
    assert (v.get_sdf(), v.get_color(), v.get_weight()) == tuple(orc.get_voxel(pos[0]))
    vt.close()


def test_bad_arguments_raise():
    from union_thesis_slam_b200 import _capi, bucket_table
    with pytest.raises(_capi.TSDFError):
        bucket_table.VoxelBucketTable(0)
    vt = bucket_table.VoxelBucketTable(8)
    with pytest.raises(_capi.TSDFError):
        vt.apply([7], [[0, 0, 0]])
    with pytest.raises(_capi.TSDFError):
        vt.apply([bo.OP_ADD], [[1 << 40, 0, 0]])
    assert vt.get_hash_entry([1, 2, 3]) is None and vt.remove_hash_entry(bucket_table.HashEntry([1, 2, 3], None, None)) == 0
    vt.close()


# ---- HashTable(per_voxel=True): the reference's whole integrate path on its own data structure ----
def _feed_layout_oracle(orc_table, hashvox, before_weight):
    """Replays one frame's insertions into the container oracle: the voxels the frame updated, in
    vox_coords order (C order of the volume), looked up and inserted on a miss (hash_fusion.py:135-145)."""
    for v in np.argwhere(hashvox.weight != before_weight):
        if orc_table.find(v)[0] < 0:
            orc_table.add(v)


def test_per_voxel_mode_reproduces_reference_hash_table_config1():
    """Bundled frames 0..2 at config 1 (405 x 264 x 289 @ 2 cm, 1 000 000 buckets): entry count,
    non-empty buckets, collisions and the sha256 of both get_volume() planes are the numbers the
    UNMODIFIED reference HashTable produced (tests/golden/golden.json, made by make_golden.py)."""
    import frames
    import scenes
    from union_thesis_slam_b200 import hash_fusion
    gold = frames.golden()["hash"]
    K = frames.intrinsics()
    ht = hash_fusion.HashTable(np.array(scenes.CONFIG1_BOUNDS), scenes.CONFIG1_VOXEL, map_size=1000000, per_voxel=True)
    for n in range(3):
        c, d, p = frames.load_frame(n)
        ht.integrate(c, d, K, p)
        g = gold[str(n + 1)]
        assert ht.count_num_hash_entries() == g["entries"]
        assert ht.get_num_non_empty_bucket() == g["non_empty_buckets"]
        assert ht.get_num_collisions() == g["collisions"]
        assert ht.voxel_table.stats()["table_size"] == g["table_size"]
        t, col = ht.get_volume()
        assert frames.sha(t) == g["tsdf"] and frames.sha(col) == g["color"], "frame %d" % n
    ht.close()


@pytest.mark.parametrize("seed", [21, 22])
def test_per_voxel_mode_small_reference_fixtures(seed):
    """Full outputs of the reference HashTable on small cases (map_size 20 000, so the table resizes)."""
    from union_thesis_slam_b200 import hash_fusion
    z = np.load(os.path.join(HERE, "golden", "small_hash_%d.npz" % seed))
    ht = hash_fusion.HashTable(z["vol_bnds"].copy(), float(z["voxel_size"]), map_size=20000, per_voxel=True)
    for c, d, p, obs in zip(z["color"], z["depth"], z["pose"], z["obs"]):
        ht.integrate(c, d, z["K"], p, obs_weight=float(obs))
    t, col = ht.get_volume()
    assert np.array_equal(t, z["tsdf"]) and np.array_equal(col, z["color_vol"])
    rows = ht.voxel_table.layout()
    assert np.array_equal(np.array(sorted(rows[:, 2:5].tolist()), np.int64), z["entries"])
    st = ht.voxel_table.stats()
    assert st["table_size"] == int(z["table_size"]) and st["non_empty_buckets"] == int(z["non_empty"])
    ht.close()


def test_per_voxel_mode_matches_oracles_on_random_case():
    """Slot layout against the container oracle, float64 state against the per-voxel hash oracle."""
    import scenes
    from oracle import oracle
    from union_thesis_slam_b200 import hash_fusion
    case = scenes.small_case(seed=91, dims=(33, 41, 27), voxel_size=0.02, hw=(60, 80), nframes=4)
    ht = hash_fusion.HashTable(case["vol_bnds"].copy(), case["voxel_size"], map_size=3000, per_voxel=True)
    oh = oracle.OracleHashVox(case["vol_bnds"].copy(), case["voxel_size"])
    ot = bo.OracleBucketTable(3000)
    for c, d, p, obs in case["frames"]:
        before = oh.weight.copy()
        n = oh.integrate(c, d, case["K"], p, obs_weight=obs)
        _feed_layout_oracle(ot, oh, before)
        got = ht.voxel_table.integrate_frame(ht._vol_dim, ht._vol_origin, ht._voxel_size, ht._trunc_margin, c, d, case["K"], p)
        assert got == n
    t, col = ht.get_volume()
    et, ec = oh.get_volume()
    assert np.array_equal(t, et) and np.array_equal(col, ec)
    rows, vox = ht.voxel_table.layout(with_voxels=True)
    assert rows.tolist() == [list(r) for r in ot.layout()]
    idx = tuple(rows[:, 2:5].T)
    assert np.array_equal(vox[:, 0], oh.sdf[idx]) and np.array_equal(vox[:, 2], oh.weight[idx])
    assert np.array_equal(vox[:, 1], oh.color[idx])
    assert ht.voxel_table.stats()["resizes"] == ot.resizes >= 1
    ht.close()


def test_returned_entries_write_through_or_refuse():
    """Reference hash_fusion.py:145: `target_entry.integrate_voxel(...)` mutates the stored voxel.  Entries of the
    device container write through; snapshots of voxels in the block map refuse mutation instead of dropping it."""
    import scenes
    from union_thesis_slam_b200 import bucket_table, hash_fusion
    vt = bucket_table.VoxelBucketTable(64)
    orc = bo.OracleBucketTable(64)
    c1, c2 = float(7 * 65536 + 8 * 256 + 9), float(100 * 65536 + 50 * 256 + 25)
    vt.integrate_voxels([[3, 4, 5]], [0.25], [c1])
    orc.integrate([3, 4, 5], 0.25, c1)
    e = vt.get_hash_entry([3, 4, 5])
    e.integrate_voxel(-0.5, c2)
    b, s_ = orc.integrate([3, 4, 5], -0.5, c2)
    want = orc.vox[orc._i(b, s_)]
    got = vt.get_hash_entry([3, 4, 5]).get_voxel()
    assert (got.get_sdf(), got.get_color(), got.get_weight()) == tuple(want)
    assert (e.get_voxel().get_sdf(), e.get_voxel().get_weight()) == (want[0], want[2])   # the handed-out object follows
    case = scenes.small_case(7, dims=(24, 24, 24), voxel_size=0.02, nframes=1)
    ht = hash_fusion.HashTable(case["vol_bnds"].copy(), 0.02, map_size=4096)
    c, d, p, _ = case["frames"][0]
    ht.integrate(c, d, case["K"], p)
    _, w, _ = ht.get_volume_with_weight()
    v = np.argwhere(w > 0)[0]
    with pytest.raises(TypeError):
        ht.get_voxel(v).integrate(0.1, 0.0)
