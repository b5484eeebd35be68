"""Per-frame floor of the dense pipeline (development aid): frames whose depth image is all-invalid (nothing visible, the
integrate kernels exit at once) and real frames, pipelined (events around N frames) at world 1 and 8."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from union_thesis_slam_b200 import grid_fusion, synth
room = synth.room_bounds((10.0, 10.0, 4.0))
frames, K = synth.room_stream(32, room, seed=0)
dev = [(torch.from_numpy(c).cuda(), torch.from_numpy(d).cuda(), np.linalg.inv(p)) for c, d, p in frames]
zero = torch.zeros_like(dev[0][1])
for world in (1, 8):
    sys.stdout = sys.stderr
    vol = grid_fusion.TSDFVolume(room.copy(), 0.01, rank=3 % world, world_size=world)
    sys.stdout = sys.__stdout__
    stream = torch.cuda.ExternalStream(vol.cuda_stream())
    for name in ("empty", "real"):
        for rep in range(2):
            vol.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(stream)
            n = 200
            for i in range(n):
                dc, dd, ti = dev[i % 32]
                vol.integrate_device(dc, zero if name == "empty" else dd, 480, 640, K, ti, 1.0)
            t_issue = time.perf_counter() - t0
            vol.flush(); e1.record(stream); vol.sync()
        print("world %d %s frames: %.1f us/frame on the device, host issue %.1f us/frame" % (
            world, name, 1e3 * e0.elapsed_time(e1) / n, 1e6 * t_issue / n))
    vol.close()
