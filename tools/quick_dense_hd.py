"""Dense path with 1280x720 frames in the 40x40x6 m room at 2 cm (same stream as the hash bench)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from union_thesis_slam_b200 import grid_fusion, synth
W, H = (1280, 720) if len(sys.argv) < 2 else (int(sys.argv[1]), int(sys.argv[2]))
room = synth.room_bounds((40.0, 40.0, 6.0))
frames, K = synth.room_stream(4, room, width=W, height=H, seed=1)
vol = grid_fusion.TSDFVolume(room.copy(), 0.02)
dev = [(torch.from_numpy(c).cuda(), torch.from_numpy(d).cuda(), np.linalg.inv(p)) for c, d, p in frames]
torch.cuda.synchronize()
vol.set_profiling(True, frame_only=True)
REPS = 30
for rep in range(REPS):
    for dc, dd, ti in dev:
        vol.integrate_device(dc, dd, H, W, K, ti, 1.0)
        if rep in (2, REPS - 1):
            ms = vol.last_kernel_ms(); st = vol.stats()
            by = 24.0 * st["updated_last"] + 12.0 * H * W
            print("rep %d frame: %.3f ms updated %d bricks %d -> %.0f GB/s (%.1f%%)" % (rep, ms, st["updated_last"], st["bricks_visited_last"], by / ms / 1e6, by / ms / 1e6 / 65.51))
