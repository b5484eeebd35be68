"""Fixed cost of a frame: integrate an all-invalid depth image (no visible bricks)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from union_thesis_slam_b200 import grid_fusion, synth
room = synth.room_bounds((10.0, 10.0, 4.0))
frames, K = synth.room_stream(2, room, seed=0)
for world in (1, 8):
    vol = grid_fusion.TSDFVolume(room.copy(), 0.01, rank=0, world_size=world)
    c, d, p = frames[0]
    dc = torch.from_numpy(c).cuda(); dd = torch.zeros_like(torch.from_numpy(d)).cuda(); ti = np.linalg.inv(p)
    dreal = torch.from_numpy(d).cuda()
    vol.set_profiling(True)
    for name, depth in (("empty", dd), ("real", dreal)):
        ts = []
        for i in range(20):
            vol.integrate_device(dc, depth, 480, 640, K, ti, 1.0)
            ts.append(vol.last_stage_ms())
        ts = np.array(ts[5:]) * 1e3
        print("world %d %s frame: prepare+cull %.1f us, integrate %.1f us, updates %d" % (
            world, name, ts[:, 0].mean(), ts[:, 1].mean(), vol.stats()["updated_last"]))
    vol.close()
