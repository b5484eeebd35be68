"""Host cost of the fan-out step on ONE GPU (development aid): a rank-of-N shard of the bench's room driven through
dist.FrameFanout with a one-rank proxy, so every frame is staged by this process (the source rank's step) and the
device time is a shard's.  usage: python tools/quick_fan.py [steps [world rank [u16|f64 [generic]]]]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from union_thesis_slam_b200 import grid_fusion, synth, dist as tdist

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 3
raw = len(sys.argv) > 4 and sys.argv[4] == "u16"
room = synth.room_bounds((10.0, 10.0, 4.0))
frames, K = synth.room_stream(32, room, seed=0)
sys.stdout = sys.stderr
vol = grid_fusion.TSDFVolume(room.copy(), 0.01, rank=rank, world_size=world)
sys.stdout = sys.__stdout__


class OneRank:
    rank, world_size, device = 0, 1, vol.device
    integrate_device = staticmethod(vol.integrate_device)
    integrate_device_raw = staticmethod(vol.integrate_device_raw)
    sync = staticmethod(vol.sync)


h = []
for c, d, p in frames:
    dd = np.round(d * 1000.0).astype(np.uint16) if raw else d
    pc, pd = torch.from_numpy(c).pin_memory(), torch.from_numpy(dd).pin_memory()
    h.append((pc.numpy(), pd.numpy(), p, pc, pd))
fan = tdist.FrameFanout(OneRank, 480, 640, raw_depth=raw, slots=8)
if not (len(sys.argv) > 5 and sys.argv[5] == "generic"):
    fan._native, fan.target = (vol._handle, None), vol  # the one-call native step, as for the product's own classes
for i in range(8):
    fan.integrate(h[i][0], h[i][1], K, h[i][2])
vol.sync()
for depth in (1, 3, 6):
    t_int = t_post = t_wait = 0.0
    vol.sync()
    t0 = time.perf_counter()
    posted = []
    for i in range(steps):
        a = time.perf_counter()
        fan.integrate(h[i % 32][0], h[i % 32][1], K, h[i % 32][2])
        b = time.perf_counter()
        posted.append(vol.stats_post())
        c = time.perf_counter()
        if len(posted) > depth:
            vol.stats_wait(posted.pop(0))
        d = time.perf_counter()
        t_int += b - a; t_post += c - b; t_wait += d - c
    for t in posted:
        vol.stats_wait(t)
    vol.sync()
    tot = time.perf_counter() - t0
    print("fan-out %s, read-back depth %d: %7.1f us/step   host: integrate %6.1f  post %5.1f  wait %6.1f" % (
        "u16" if raw else "f64", depth, 1e6 * tot / steps, 1e6 * t_int / steps, 1e6 * t_post / steps, 1e6 * t_wait / steps))
fan.close()
