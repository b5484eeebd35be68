"""Tiny end-to-end run of both maps for compute-sanitizer (memcheck / racecheck)."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import scenes
from union_thesis_slam_b200 import grid_fusion, hash_fusion
case = scenes.small_case(7, dims=(24, 20, 28), voxel_size=0.02, nframes=2, hw=(24, 32))
vol = grid_fusion.TSDFVolume(case["vol_bnds"].copy(), 0.02)
ht = hash_fusion.HashTable(case["vol_bnds"].copy(), 0.02, map_size=1024, heap_blocks=64)
for c, d, p, obs in case["frames"]:
    vol.integrate(c, d, case["K"], p, obs_weight=obs)
    ht.integrate(c, d, case["K"], p, obs_weight=obs)
t, c = vol.get_volume(); ht_t, ht_c = ht.get_volume()
keys, payload = ht.export_blocks()
print("removed", ht.remove_blocks(keys[:3]), "of", len(keys), "equal", np.array_equal(t, ht_t))
v, f, n, col = vol.get_mesh()
print("mesh", len(v), len(f))
