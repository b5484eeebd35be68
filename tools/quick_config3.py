"""BASELINE configs[2]-shaped timing: bundled fixture frames replayed into the 5 mm grid (1.97e9 voxels)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import frames as fx, scenes
from union_thesis_slam_b200 import grid_fusion
vs = float(sys.argv[1]) if len(sys.argv) > 1 else 0.005
K = fx.intrinsics()
vol = grid_fusion.TSDFVolume(np.array(scenes.CONFIG1_BOUNDS), vs)
fr = [fx.load_frame(i) for i in range(10)]
dev = [(torch.from_numpy(c).cuda(), torch.from_numpy(d).cuda(), np.linalg.inv(p)) for c, d, p in fr]
torch.cuda.synchronize()
H, W = fr[0][1].shape
stream = torch.cuda.ExternalStream(vol.cuda_stream())
for rep in range(6):
    vol.sync(); u0 = vol.stats()["updated_total"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for dc, dd, ti in dev:
        vol.integrate_device(dc, dd, H, W, K, ti, 1.0)
    vol.flush(); e1.record(stream); vol.sync()
    ms = e0.elapsed_time(e1) / len(dev); u = (vol.stats()["updated_total"] - u0) / len(dev)
    by = 24.0 * u + 12.0 * H * W
    print("rep %d: %.1f us/frame, %.3e updates/frame, %.3e voxel-updates/s, %.0f GB/s (%.1f%% of 6551)" % (
        rep, 1e3 * ms, u, u / (ms * 1e-3), by / ms / 1e6, by / ms / 1e6 / 65.51))
vol.set_profiling(True)
for dc, dd, ti in dev[:3]:
    vol.integrate_device(dc, dd, H, W, K, ti, 1.0)
    st = vol.stats(); print("stages [prepare+cull, integrate] us:", [round(1e3 * x, 1) for x in vol.last_stage_ms()], "bricks", st["bricks_visited_last"], "updated", st["updated_last"])
