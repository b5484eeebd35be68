"""A/B of the dense integrate variants on the bench's 1 cm room (development aid): device time of the bench's own
regime (5 warm-up frames + 20 timed, first pass over the trajectory) and of the steady state (third pass, 32 frames).
usage: python tools/quick_batch.py BATCH [world rank]     (TSDF_DENSE_V1=1 selects the round-1 kernel, batch 1 only)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from union_thesis_slam_b200 import grid_fusion, synth

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
world = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
room = synth.room_bounds((10.0, 10.0, 4.0))
frames, K = synth.room_stream(32, room, seed=0)
H, W = frames[0][1].shape
dev = [(torch.from_numpy(c).cuda(), torch.from_numpy(d).cuda(), np.linalg.inv(p)) for c, d, p in frames]
sys.stdout = sys.stderr
vol = grid_fusion.TSDFVolume(room.copy(), 0.01, rank=rank, world_size=world)
sys.stdout = sys.__stdout__
vol.set_batch(batch)
stream = torch.cuda.ExternalStream(vol.cuda_stream())

def run(i0, n):
    u0 = vol.stats()["updated_total"]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(i0, i0 + n):
        dc, dd, ti = dev[i % 32]
        vol.integrate_device(dc, dd, H, W, K, ti, 1.0)
    vol.flush()
    e1.record(stream)
    vol.sync()
    ms = e0.elapsed_time(e1)
    upd = vol.stats()["updated_total"] - u0
    gbs = (24.0 * upd + 12.0 * H * W * n) / (ms * 1e-3) / 1e9
    return ms / n, upd / (ms * 1e-3), gbs

for i in range(5):
    dc, dd, ti = dev[i]
    vol.integrate_device(dc, dd, H, W, K, ti, 1.0)
vol.sync()
tag = "v2" if (os.environ.get("TSDF_DENSE_V2") or batch > 1) else "v1"
r = run(5, 20)
print("%s batch %d world %d: bench regime   %.4f ms/frame  %.4g upd/s  %.0f GB/s (%.3f of 6551)" % (tag, batch, world, r[0], r[1], r[2], r[2] / 6551))
run(25, 7 + 32)
r = run(0, 32)
print("%s batch %d world %d: third pass     %.4f ms/frame  %.4g upd/s  %.0f GB/s (%.3f of 6551)" % (tag, batch, world, r[0], r[1], r[2], r[2] / 6551))
for _ in range(6):
    run(0, 32)
r = run(0, 32)
print("%s batch %d world %d: tenth pass     %.4f ms/frame  %.4g upd/s  %.0f GB/s (%.3f of 6551)" % (tag, batch, world, r[0], r[1], r[2], r[2] / 6551))
