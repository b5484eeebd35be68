"""Generate tests/golden/ fixtures by running the UNMODIFIED reference in this container.

    python tests/golden/make_golden.py            # ~5 min on 8 cores

Needs /root/reference (build container only).  Writes:
  frames/                 raw bytes of bundled frames {0..9, 500, 999} + camera-intrinsics.txt
                          (test inputs; decoded at test time exactly as grid_demo1.py:80-84)
  golden.json             sha256/counts of reference TSDFVolume after 1 and 10 frames (config 1),
                          of reference HashTable after frames 0..2, decoded-input digests, hash KATs
  small_dense_<seed>.npz  inputs + full reference outputs for tests/scenes.GOLDEN_SMALL_DENSE
  small_hash_<seed>.npz   same for the reference HashTable
The script also cross-checks the C oracle against every value it writes and refuses to write
fixtures the oracle disagrees with.
"""
import hashlib
import json
import os
import shutil
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle, ref_runner  # noqa: E402
import scenes  # noqa: E402

FRAME_IDS = list(range(10)) + [500, 999]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    grid_fusion, hash_fusion = ref_runner.load()
    gold = {"generated_with": "reference grid_fusion.py / hash_fusion.py, numba CPU path",
            "config1_bounds": scenes.CONFIG1_BOUNDS, "config1_voxel": scenes.CONFIG1_VOXEL}

    # -- inputs ---------------------------------------------------------------------------
    fdir = os.path.join(HERE, "frames")
    os.makedirs(fdir, exist_ok=True)
    src = os.path.join(ref_runner.REFERENCE_ROOT, "data")
    shutil.copyfile(os.path.join(src, "camera-intrinsics.txt"), os.path.join(fdir, "camera-intrinsics.txt"))
    gold["inputs"] = {}
    for i in FRAME_IDS:
        for ext in ("color.jpg", "depth.png", "pose.txt"):
            n = "frame-%06d.%s" % (i, ext)
            shutil.copyfile(os.path.join(src, n), os.path.join(fdir, n))
        c, d, p = ref_runner.load_bundled_frame(i)
        gold["inputs"][str(i)] = {"color": sha(c), "depth": sha(d), "pose": sha(p)}
    K = np.loadtxt(os.path.join(src, "camera-intrinsics.txt"), delimiter=" ")

    # -- config 1, dense reference vs C oracle ---------------------------------------------
    t0 = time.time()
    ref = grid_fusion.TSDFVolume(np.array(scenes.CONFIG1_BOUNDS), scenes.CONFIG1_VOXEL)
    orc = oracle.OracleTSDFVolume(np.array(scenes.CONFIG1_BOUNDS), scenes.CONFIG1_VOXEL)
    gold["config1_dim"] = [int(x) for x in ref._vol_dim]
    gold["dense"] = {}
    for n, i in enumerate(range(10), 1):
        c, d, p = ref_runner.load_bundled_frame(i)
        ref.integrate(c, d, K, p, obs_weight=1.)
        orc.integrate(c, d, K, p, obs_weight=1.)
        if n in (1, 2, 10):
            rec = {"tsdf": sha(ref._tsdf_vol_cpu), "weight": sha(ref._weight_vol_cpu),
                   "color": sha(ref._color_vol_cpu),
                   "n_w_pos": int((ref._weight_vol_cpu > 0).sum()),
                   "sum_w": int(ref._weight_vol_cpu.astype(np.float64).sum()),
                   "n_blocks": int(len(oracle.block_set_from_weight(ref._weight_vol_cpu)))}
            assert rec["tsdf"] == sha(orc._tsdf_vol_cpu), "oracle tsdf != reference at %d" % n
            assert rec["weight"] == sha(orc._weight_vol_cpu), "oracle weight != reference at %d" % n
            assert rec["color"] == sha(orc._color_vol_cpu), "oracle colour != reference at %d" % n
            gold["dense"][str(n)] = rec
            print("dense after", n, rec, "%.0fs" % (time.time() - t0), flush=True)
    del ref

    # -- config 2, reference HashTable (per-voxel map) frames 0..2 ---------------------------
    ht = hash_fusion.HashTable(np.array(scenes.CONFIG1_BOUNDS), scenes.CONFIG1_VOXEL, map_size=1000000)
    oh = oracle.OracleHashVox(np.array(scenes.CONFIG1_BOUNDS), scenes.CONFIG1_VOXEL)
    gold["hash"] = {}
    for n, i in enumerate(range(3), 1):
        c, d, p = ref_runner.load_bundled_frame(i)
        ht.integrate(c, d, K, p, obs_weight=1.)
        oh.integrate(c, d, K, p, obs_weight=1.)
        tv, cv = ht.get_volume()
        otv, ocv = oh.get_volume()
        rec = {"tsdf": sha(tv), "color": sha(cv), "entries": int(ht.count_num_hash_entries()),
               "non_empty_buckets": int(ht.get_num_non_empty_bucket()),
               "collisions": int(ht.get_num_collisions()), "table_size": int(ht._table_size),
               "n_blocks": int(len(oh.block_set()))}
        assert rec["tsdf"] == sha(otv) and rec["color"] == sha(ocv), "hash oracle != reference at %d" % n
        assert rec["entries"] == int(oh.entries.sum())
        gold["hash"][str(n)] = rec
        print("hash after", n, rec, "%.0fs" % (time.time() - t0), flush=True)
    del ht

    # -- small random volumes -------------------------------------------------------------
    for spec in scenes.GOLDEN_SMALL_DENSE:
        case = scenes.small_case(**spec)
        ref = grid_fusion.TSDFVolume(case["vol_bnds"].copy(), case["voxel_size"])
        orc = oracle.OracleTSDFVolume(case["vol_bnds"].copy(), case["voxel_size"])
        assert tuple(ref._vol_dim) == case["dims"], (ref._vol_dim, case["dims"])
        for (c, d, p, obs) in case["frames"]:
            ref.integrate(c, d, case["K"], p, obs_weight=obs)
            orc.integrate(c, d, case["K"], p, obs_weight=obs)
        assert np.array_equal(ref._tsdf_vol_cpu, orc._tsdf_vol_cpu)
        assert np.array_equal(ref._weight_vol_cpu, orc._weight_vol_cpu)
        assert np.array_equal(ref._color_vol_cpu, orc._color_vol_cpu)
        nupd = int((ref._weight_vol_cpu > 0).sum())
        assert nupd > 500, "degenerate golden case %s: %d voxels touched" % (spec, nupd)
        np.savez_compressed(
            os.path.join(HERE, "small_dense_%d.npz" % spec["seed"]),
            vol_bnds=case["vol_bnds"], voxel_size=case["voxel_size"], K=case["K"],
            color=np.stack([f[0] for f in case["frames"]]),
            depth=np.stack([f[1] for f in case["frames"]]),
            pose=np.stack([f[2] for f in case["frames"]]),
            obs=np.array([f[3] for f in case["frames"]]),
            tsdf=ref._tsdf_vol_cpu, weight=ref._weight_vol_cpu, color_vol=ref._color_vol_cpu)
        print("small dense", spec["seed"], case["dims"], "touched", nupd, flush=True)

    for spec in scenes.GOLDEN_SMALL_HASH:
        case = scenes.small_case(**spec)
        ht = hash_fusion.HashTable(case["vol_bnds"].copy(), case["voxel_size"], map_size=20000)
        oh = oracle.OracleHashVox(case["vol_bnds"].copy(), case["voxel_size"])
        for (c, d, p, obs) in case["frames"]:
            ht.integrate(c, d, case["K"], p, obs_weight=obs)
            oh.integrate(c, d, case["K"], p, obs_weight=obs)
        tv, cv = ht.get_volume()
        otv, ocv = oh.get_volume()
        assert np.array_equal(tv, otv) and np.array_equal(cv, ocv)
        # entry positions straight from the reference container
        pos = []
        for b in ht._hash_table:
            if b is None:
                continue
            for j in range(ht._bucket_size):
                e = b.get_ith_entry(j)
                if e is not None:
                    pos.append([int(v) for v in e.get_position()])
        pos = np.array(sorted(pos), dtype=np.int64)
        assert np.array_equal(pos, np.argwhere(oh.entries))
        np.savez_compressed(
            os.path.join(HERE, "small_hash_%d.npz" % spec["seed"]),
            vol_bnds=case["vol_bnds"], voxel_size=case["voxel_size"], K=case["K"],
            color=np.stack([f[0] for f in case["frames"]]),
            depth=np.stack([f[1] for f in case["frames"]]),
            pose=np.stack([f[2] for f in case["frames"]]),
            obs=np.array([f[3] for f in case["frames"]]),
            tsdf=tv, color_vol=cv, entries=pos,
            table_size=ht._table_size, non_empty=ht.get_num_non_empty_bucket())
        print("small hash", spec["seed"], case["dims"], "entries", len(pos), flush=True)

    # -- hash_function known answers (hash_fusion.py:182-190) ------------------------------
    kats = []
    for coord, size in [([333, 234, 241], 100000), ([342, 234, 241], 100000), ([332, 234, 242], 100000),
                        ([1, 2, 3], 1000000), ([404, 263, 288], 1000000), ([0, 0, 0], 1000000),
                        ([50, 33, 36], 65536), ([123456, 7, 99999], 2000000)]:
        t = hash_fusion.HashTable.__new__(hash_fusion.HashTable)
        t._table_size = size
        h = int(t.hash_function(np.array(coord)))
        assert h == oracle.hash_function(coord, size)
        kats.append({"coord": coord, "table_size": size, "hash": h})
    gold["hash_kats"] = kats

    # -- counts frozen in the reference's committed pstats dumps ---------------------------
    gold["pstats_counts"] = {"voxel_integrate_frame0": 354383, "add_hash_entry_10frames": 389590,
                             "get_hash_entry_10frames": 3557017}
    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(gold, fh, indent=1, sort_keys=True)
    print("wrote golden.json in %.0fs" % (time.time() - t0))


if __name__ == "__main__":
    main()
