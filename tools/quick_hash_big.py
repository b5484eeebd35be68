"""Hash map at BASELINE configs[4] scale (development aid)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from union_thesis_slam_b200 import hash_fusion, synth
size = (40.0, 40.0, 6.0); vs = 0.02; W, H = 1280, 720
nfr = int(sys.argv[1]) if len(sys.argv) > 1 else 8
room = synth.room_bounds(size)
frames, K = synth.room_stream(nfr, room, width=W, height=H, seed=1)
nblocks = int(np.prod(np.ceil(np.array(size) / vs / 8)))
ht = hash_fusion.HashTable(room.copy(), vs, map_size=4 * nblocks, heap_blocks=nblocks)
dev = [(torch.from_numpy(c).cuda(), torch.from_numpy(d).cuda(), np.linalg.inv(p)) for c, d, p in frames]
torch.cuda.synchronize()
for rep in range(3):
    for dc, dd, ti in dev:
        ht.integrate_device(dc, dd, H, W, K, ti, 1.0)
        ms = ht.last_kernel_ms(); st = ht.stats()
        if rep == 2:
            by = 24.0 * st["updated_last"] + 16.0 * (st["blocks_visited_last"] + st["keys_inserted_last"]) + 12.0 * H * W
            print("frame: %.3f ms updated %d visited %d cand %d newkeys %d keys %d blocks %d -> %.0f GB/s (%.1f%%)" % (
                ms, st["updated_last"], st["blocks_visited_last"], st["candidates_last"], st["keys_inserted_last"],
                st["table_keys"], st["blocks"], by / ms / 1e6, by / ms / 1e6 / 65.51))
