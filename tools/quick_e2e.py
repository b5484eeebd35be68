"""Where does the end-to-end step go?  Host time per call (integrate / stats_post / stats_wait) of the pipelined
loop on the bench's 1 cm room (development aid).  usage: python tools/quick_e2e.py [steps [world rank]]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from union_thesis_slam_b200 import grid_fusion, synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
world = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
room = synth.room_bounds((10.0, 10.0, 4.0))
frames, K = synth.room_stream(32, room, seed=0)
sys.stdout = sys.stderr
vol = grid_fusion.TSDFVolume(room.copy(), 0.01, rank=rank, world_size=world)
sys.stdout = sys.__stdout__
h = []
for c, d, p in frames:
    pc, pd = torch.from_numpy(c).pin_memory(), torch.from_numpy(d).pin_memory()
    raw = np.where(d > 0, np.round(d * 1000.0), 65535).astype(np.uint16)
    pr = torch.from_numpy(raw).pin_memory()
    h.append((pc.numpy(), pd.numpy(), p, pc, pd, pr.numpy(), pr))
dev = [(torch.from_numpy(c).cuda(), torch.from_numpy(d).cuda(), np.linalg.inv(p)) for c, d, p in frames]
for i in range(8):
    vol.integrate(h[i][0], h[i][1], K, h[i][2])
vol.sync()
for mode in ("host frames, pipelined read", "host frames, no read", "host u16 frames, pipelined read", "host u16 frames, no read",
             "resident frames, pipelined read", "resident frames, no read"):
    t_int = t_post = t_wait = 0.0
    vol.sync()
    t0 = time.perf_counter()
    prev = None
    for i in range(steps):
        a = time.perf_counter()
        if mode.startswith("host u16"):
            vol.integrate_raw(h[i % 32][0], h[i % 32][5], K, h[i % 32][2])
        elif mode.startswith("host"):
            vol.integrate(h[i % 32][0], h[i % 32][1], K, h[i % 32][2])
        else:
            vol.integrate_device(dev[i % 32][0], dev[i % 32][1], 480, 640, K, dev[i % 32][2], 1.0)
        b = time.perf_counter()
        if "pipelined" in mode:
            t = vol.stats_post()
            c = time.perf_counter()
            if prev is not None:
                vol.stats_wait(prev)
            prev = t
            d = time.perf_counter()
        else:
            c = d = b
        t_int += b - a; t_post += c - b; t_wait += d - c
    vol.sync()
    tot = time.perf_counter() - t0
    print("%-34s %7.1f us/step   host: integrate %6.1f  post %5.1f  wait %6.1f" % (
        mode, 1e6 * tot / steps, 1e6 * t_int / steps, 1e6 * t_post / steps, 1e6 * t_wait / steps))
