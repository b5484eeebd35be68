"""Quick device-side timing of the dense path on the synthetic room (development aid)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from union_thesis_slam_b200 import grid_fusion, synth

vs = float(sys.argv[1]) if len(sys.argv) > 1 else 0.01
nfr = int(sys.argv[2]) if len(sys.argv) > 2 else 16
room = synth.room_bounds((10.0, 10.0, 4.0))
frames, K = synth.room_stream(nfr, room, seed=0)
vol = grid_fusion.TSDFVolume(room.copy(), vs)
dev = [(torch.from_numpy(c).cuda(), torch.from_numpy(d).cuda(), np.linalg.inv(p)) for c, d, p in frames]
torch.cuda.synchronize()
vol.set_profiling(True, frame_only=True)
H, W = frames[0][1].shape
REPS = int(sys.argv[3]) if len(sys.argv) > 3 else 3
for rep in range(REPS):
    tot_ms = 0.0; tot_u = 0
    t0 = time.perf_counter()
    for dc, dd, ti in dev:
        vol.integrate_device(dc, dd, H, W, K, ti, 1.0)
        ms = vol.last_kernel_ms(); st = vol.stats()
        tot_ms += ms; tot_u += st["updated_last"]
        if rep == REPS - 1:
            by = 24.0 * st["updated_last"] + 12.0 * H * W
            print("frame: %.3f ms  updated %d  bricks %d  -> %.0f GB/s algorithmic (%.1f%% of 6551)" % (
                ms, st["updated_last"], st["bricks_visited_last"], by / ms / 1e6, by / ms / 1e6 / 65.51))
    print("rep %d: %.3f ms/frame kernel, %.3e voxel-updates/s, wall %.3f s" % (
        rep, tot_ms / nfr, tot_u / (tot_ms / 1e3), time.perf_counter() - t0))
