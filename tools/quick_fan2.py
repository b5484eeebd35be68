"""Fan-out dynamics at a shard's frame time on TWO GPUs (development aid): each process holds one rank-of-8 shard of
the bench's room and the frames rotate between the two processes through dist.FrameFanout.
usage: torchrun --nproc-per-node 2 tools/quick_fan2.py [u16]"""
import os, sys, time
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, ".")
from union_thesis_slam_b200 import grid_fusion, synth, dist as tdist

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
raw = len(sys.argv) > 1 and sys.argv[1] == "u16"
room = synth.room_bounds((10.0, 10.0, 4.0))
frames, K = synth.room_stream(32, room, seed=0)
sys.stdout = open(os.devnull, "w")
vol = grid_fusion.TSDFVolume(room.copy(), 0.01, rank=3 + rank, world_size=8, device=local)
sys.stdout = sys.__stdout__
h = []
for c, d, p in frames:
    dd = np.round(d * 1000.0).astype(np.uint16) if raw else d
    pc, pd = torch.from_numpy(c).pin_memory(), torch.from_numpy(dd).pin_memory()
    h.append((pc.numpy(), pd.numpy(), p, pc, pd))
fan = tdist.FrameFanout(vol, 480, 640, raw_depth=raw, rank=rank, world_size=world)
flag = torch.zeros(1, device="cuda")


def run(steps, depth, first):
    vol.sync()
    dist.all_reduce(flag)
    torch.cuda.synchronize()
    t_int = t_post = t_wait = 0.0
    t0 = time.perf_counter()
    posted = []
    for i in range(first, first + steps):
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
    fan.drain()
    t1 = time.perf_counter()
    dist.all_reduce(flag)
    torch.cuda.synchronize()
    tot = time.perf_counter() - t0
    if rank == 0:
        print("%s steps %4d depth %d: %7.1f us/step (%.1f before the closing barrier)  host: integrate %6.1f  post %5.1f  wait %6.1f" % (
            "u16" if raw else "f64", steps, depth, 1e6 * tot / steps, 1e6 * (t1 - t0) / steps, 1e6 * t_int / steps,
            1e6 * t_post / steps, 1e6 * t_wait / steps), flush=True)


for i in range(8):
    fan.integrate(h[i][0], h[i][1], K, h[i][2])
for steps in (20, 20, 400):
    for depth in (3, 6):
        run(steps, depth, 8)
fan.close()
dist.destroy_process_group()
