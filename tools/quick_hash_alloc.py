"""A/B of the hash map's allocation front end (TSDF_HASH_ALLOC=walk|grid) on the bench stream
(BASELINE configs[4] shape): wall ms/frame of rank 0 of `world`, one GPU."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from union_thesis_slam_b200 import hash_fusion, synth
ROOM = tuple(float(x) for x in os.environ.get("HASH_ROOM", "40,40,6").split(","))
room = synth.room_bounds(ROOM)
# HASH_BOUNDS: bounds of the MAP (>= the room the frames are rendered in): a map bounded much larger than what is
# ever observed is where the bounded-grid scan loses to the cone walk
BOUNDS = synth.room_bounds(tuple(float(x) for x in os.environ.get("HASH_BOUNDS", os.environ.get("HASH_ROOM", "40,40,6")).split(",")))
frames, K = synth.room_stream(bench.N_FRAMES, room, width=bench.HASH_W, height=bench.HASH_H, seed=1)
dev = [(torch.from_numpy(c).cuda(), torch.from_numpy(d).cuda(), np.linalg.inv(p)) for c, d, p in frames]
nblocks = min(int(np.prod(np.ceil(np.array(ROOM) / bench.HASH_VOXEL / 8))), 4000000)
ref_blocks = {}
for world in [int(a) for a in sys.argv[1:]] or [1, 8]:
    for mode in ("walk", "grid", "auto", "walk", "grid"):
        os.environ["TSDF_HASH_ALLOC"] = mode
        per_rank = int(nblocks / world * 1.25) + 1024
        ht = hash_fusion.HashTable(BOUNDS.copy(), bench.HASH_VOXEL, map_size=4 * per_rank, heap_blocks=per_rank, rank=0, world_size=world)
        for dc, dd, ti in dev:
            ht.integrate_device(dc, dd, bench.HASH_H, bench.HASH_W, K, ti, 1.0)
        ht.sync()
        res = []
        for rep in range(3):
            t0 = time.perf_counter()
            for it in range(2):
                for dc, dd, ti in dev:
                    ht.integrate_device(dc, dd, bench.HASH_H, bench.HASH_W, K, ti, 1.0)
            ht.sync()
            res.append(1e3 * (time.perf_counter() - t0) / (2 * len(dev)))
        st = ht.stats()
        ref_blocks.setdefault(world, st["blocks"])
        print("world %d alloc %s: ms/frame %s | blocks %d (same as first mode: %s) table_keys %d updated_total %d" % (
            world, mode, " ".join("%.3f" % r for r in res), st["blocks"], st["blocks"] == ref_blocks[world],
            st["table_keys"], st["updated_total"]), flush=True)
        ht.close()
