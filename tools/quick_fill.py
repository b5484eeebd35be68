"""Per-frame fill ratio (updated voxels / visited brick slots) and kernel time, dense, optional shard."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from union_thesis_slam_b200 import grid_fusion, synth
world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
room = synth.room_bounds((10.0, 10.0, 4.0))
frames, K = synth.room_stream(32, room, seed=0)
vol = grid_fusion.TSDFVolume(room.copy(), 0.01, rank=0, world_size=world)
dev = [(torch.from_numpy(c).cuda(), torch.from_numpy(d).cuda(), np.linalg.inv(p)) for c, d, p in frames]
vol.set_profiling(True)
rows = []
for rep in range(2):
    for i, (dc, dd, ti) in enumerate(dev):
        vol.integrate_device(dc, dd, 480, 640, K, ti, 1.0)
        st = vol.stats(); ms = vol.last_stage_ms()
        if rep == 1:
            rows.append((i, 1e3 * ms[1], st["updated_last"], st["bricks_visited_last"]))
for i, us, u, b in rows:
    print("frame %2d: %7.1f us  updates %9d  bricks %7d  fill %.2f  ns/kslot %.2f" % (i, us, u, b, u / max(b * 512, 1), 1e3 * us / max(b * 512 / 1e3, 1)))
tot_u = sum(r[2] for r in rows); tot_b = sum(r[3] for r in rows)
print("overall fill %.3f" % (tot_u / (tot_b * 512.0)))
