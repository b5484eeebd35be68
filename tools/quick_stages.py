"""Stage times of the dense frame at world 1 / 8 (development aid): [prepare+cull, integrate] with every frame
serialised on one stream, next to the pipelined per-frame time."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from union_thesis_slam_b200 import grid_fusion, synth
room = synth.room_bounds((10.0, 10.0, 4.0))
frames, K = synth.room_stream(32, room, seed=0)
dev = [(torch.from_numpy(c).cuda(), torch.from_numpy(d).cuda(), np.linalg.inv(p)) for c, d, p in frames]
for world, rank in ((1, 0), (8, 3)):
    sys.stdout = sys.stderr
    vol = grid_fusion.TSDFVolume(room.copy(), 0.01, rank=rank, world_size=world)
    sys.stdout = sys.__stdout__
    for i in range(5):
        vol.integrate_device(dev[i][0], dev[i][1], 480, 640, K, dev[i][2], 1.0)
    vol.set_profiling(True)
    ts, us = [], []
    for i in range(5, 25):
        vol.integrate_device(dev[i][0], dev[i][1], 480, 640, K, dev[i][2], 1.0)
        ts.append(vol.last_stage_ms()); us.append(vol.stats()["updated_last"])
    ts = np.array(ts) * 1e3
    print("world %d rank %d serialised: prepare+cull %.1f us, integrate (2 launches back to back) %.1f us, %.3g updates/frame -> %.2f ns per 1000 updates" % (
        world, rank, ts[:, 0].mean(), ts[:, 1].mean(), np.mean(us), 1e3 * ts[:, 1].mean() / np.mean(us) * 1e3 / 1e3))
    vol.close()
