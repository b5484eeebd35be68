"""A/B of the two-stream checkerboard split of the dense integrate (TSDF_DENSE_CLASSES=1|2):
device time per frame of rank 0 of `world` on the bench stream, one GPU."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from union_thesis_slam_b200 import grid_fusion, synth
nfr = 32
room = synth.room_bounds((10.0, 10.0, 4.0))
frames, K = synth.room_stream(nfr, room, seed=0)
dev = [(torch.from_numpy(c).cuda(), torch.from_numpy(d).cuda(), np.linalg.inv(p)) for c, d, p in frames]
H, W = frames[0][1].shape
worlds = [int(a) for a in sys.argv[1:]] or [1, 8]
for world in worlds:
    for classes, ctas in ((2, 14), (1, 14), (2, 14)):
        os.environ["TSDF_DENSE_CLASSES"] = str(classes)
        os.environ["TSDF_DENSE_CTAS_PER_SM"] = str(ctas)
        vol = grid_fusion.TSDFVolume(room.copy(), 0.01, rank=0, world_size=world)
        stream = torch.cuda.ExternalStream(vol.cuda_stream())
        res = []
        for rep in range(5):
            vol.sync(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); e0.record(stream)
            for it in range(4):
                for dc, dd, ti in dev:
                    vol.integrate_device(dc, dd, H, W, K, ti, 1.0)
            t_issue = time.perf_counter() - t0
            vol.flush(); e1.record(stream); vol.sync()
            res.append((1e3 * e0.elapsed_time(e1) / (4 * nfr), 1e6 * t_issue / (4 * nfr)))
        st = vol.stats()
        print("world %d classes %d ctas/SM %d: device us/frame %s | host issue us/frame %.1f | updated_total %d" % (
            world, classes, ctas, " ".join("%.1f" % r[0] for r in res), res[-1][1], st["updated_total"]), flush=True)
        vol.close()
