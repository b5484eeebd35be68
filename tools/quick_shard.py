"""Per-rank cost of the dense path when the room is sharded over `world` ranks (run on one GPU)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from union_thesis_slam_b200 import grid_fusion, synth
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nfr = 32
room = synth.room_bounds((10.0, 10.0, 4.0))
frames, K = synth.room_stream(nfr, room, seed=0)
vol = grid_fusion.TSDFVolume(room.copy(), 0.01, rank=0, world_size=world)
dev = [(torch.from_numpy(c).cuda(), torch.from_numpy(d).cuda(), np.linalg.inv(p)) for c, d, p in frames]
H, W = frames[0][1].shape
stream = torch.cuda.ExternalStream(vol.cuda_stream())
for rep in range(4):
    vol.sync(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(stream)
    for dc, dd, ti in dev:
        vol.integrate_device(dc, dd, H, W, K, ti, 1.0)
    t_issue = time.perf_counter() - t0
    e1.record(stream); vol.sync()
    print("rep %d: device %.1f us/frame, host issue %.1f us/frame" % (rep, 1e3 * e0.elapsed_time(e1) / nfr, 1e6 * t_issue / nfr))
