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
    vol.flush(); e1.record(stream); vol.sync()
    print("rep %d: device %.1f us/frame, host issue %.1f us/frame" % (rep, 1e3 * e0.elapsed_time(e1) / nfr, 1e6 * t_issue / nfr))
# kernel-only time of the shard's integrate kernel (serialised, per-frame sync)
vol.set_profiling(True)
ks = []; us = []
for dc, dd, ti in dev:
    vol.integrate_device(dc, dd, H, W, K, ti, 1.0)
    ks.append(vol.last_stage_ms()[-1]); us.append(vol.stats()["updated_last"])
print("integrate kernel alone: %.1f us/frame mean; updates/frame mean %.2e (rank 0 of %d)" % (1e3 * np.mean(ks), np.mean(us), world))
# balance over ranks: updates per frame for every rank
allu = []
for r in range(world):
    v2 = grid_fusion.TSDFVolume(room.copy(), 0.01, rank=r, world_size=world)
    u = []
    for dc, dd, ti in dev:
        v2.integrate_device(dc, dd, H, W, K, ti, 1.0); u.append(v2.stats()["updated_last"])
    allu.append(u); v2.close()
allu = np.array(allu, float)
print("imbalance: mean over frames of max_rank/mean_rank = %.3f" % np.mean(allu.max(0) / np.maximum(allu.mean(0), 1)))
print("ideal speed-up with this partition: %.2f of %d" % (allu.sum() / allu.max(0).sum(), world))
ks = np.array(ks); us = np.array(us, float)
A = np.stack([np.ones_like(us), us], 1); coef = np.linalg.lstsq(A, ks, rcond=None)[0]
print("fit kernel_ms = %.1f us + %.2f ns per 1000 updates" % (1e3 * coef[0], 1e9 * coef[1]))
for k_, u_ in sorted(zip(ks, us))[::4]:
    print("  %.1f us  %.2e updates" % (1e3 * k_, u_))
