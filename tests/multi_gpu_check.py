"""Multi-GPU parity check, launched with torchrun on a box with >= 2 GPUs (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

Every rank integrates the same frames into its shard (cyclic x-sheets / owner-hashed blocks, no
per-frame collective); get_volume() gathers over NCCL and must equal the CPU oracle bit for bit.
A second pair of maps receives the frames through dist.FrameFanout (only the frame's source rank holds
the images; one NCCL broadcast per frame) and must give the same volumes.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import scenes  # noqa: E402
from oracle import oracle  # noqa: E402
from union_thesis_slam_b200 import dist as tdist, grid_fusion, hash_fusion  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    case = scenes.small_case(301, dims=(70, 40, 44), voxel_size=0.02, nframes=3)
    vol = grid_fusion.TSDFVolume(case["vol_bnds"].copy(), 0.02)      # rank/world from the process group
    ht = hash_fusion.HashTable(case["vol_bnds"].copy(), 0.02, map_size=1 << 14)
    assert (vol.rank, vol.world_size, vol.device) == (rank, world, local)
    orc = oracle.OracleTSDFVolume(case["vol_bnds"].copy(), 0.02)
    vol_f = grid_fusion.TSDFVolume(case["vol_bnds"].copy(), 0.02)
    ht_f = hash_fusion.HashTable(case["vol_bnds"].copy(), 0.02, map_size=1 << 14)
    im_h, im_w = case["frames"][0][1].shape
    fan_v = tdist.FrameFanout(vol_f, im_h, im_w, slots=2)
    fan_h = tdist.FrameFanout(ht_f, im_h, im_w, slots=2)
    for rep in range(3):  # 9 frames over a ring of 2 buffers: every buffer is reused while frames are in flight
        for c, d, p, obs in case["frames"]:
            vol.integrate(c, d, case["K"], p, obs_weight=obs)
            ht.integrate(c, d, case["K"], p, obs_weight=obs)
            orc.integrate(c, d, case["K"], p, obs_weight=obs)
            mine = fan_v.source_of(fan_v.frames) == rank  # only the source rank is given the images
            fan_v.integrate(c if mine else None, d if mine else None, case["K"], p, obs_weight=obs)
            fan_h.integrate(c if mine else None, d if mine else None, case["K"], p, obs_weight=obs)
    fan_v.drain(); fan_h.drain()
    # the same through one NCCL broadcast per frame (transport "collective")
    vol_c = grid_fusion.TSDFVolume(case["vol_bnds"].copy(), 0.02)
    fan_c = tdist.FrameFanout(vol_c, im_h, im_w, slots=2, transport="collective")
    for rep in range(3):
        for c, d, p, obs in case["frames"]:
            mine = fan_c.source_of(fan_c.frames) == rank
            fan_c.integrate(c if mine else None, d if mine else None, case["K"], p, obs_weight=obs)
    fan_c.drain()
    # one rank owns the camera (every frame from rank 0) and the ring has a single slot: the source may not
    # overwrite it before every peer has pulled, and no peer may pull before the upload is through
    vol_s = grid_fusion.TSDFVolume(case["vol_bnds"].copy(), 0.02)
    fan_s = tdist.FrameFanout(vol_s, im_h, im_w, slots=1)
    for rep in range(3):
        for c, d, p, obs in case["frames"]:
            fan_s.integrate(c if rank == 0 else None, d if rank == 0 else None, case["K"], p, obs_weight=obs, src=0)
    fan_s.drain()
    t, c_ = vol.get_volume()
    w = vol._weight_vol_cpu
    ok = (np.array_equal(t, orc._tsdf_vol_cpu) and np.array_equal(w, orc._weight_vol_cpu)
          and np.array_equal(c_, orc._color_vol_cpu))
    ht_t, ht_w, ht_c = ht.get_volume_with_weight()
    ok_h = np.array_equal(ht_t, t) and np.array_equal(ht_w, w) and np.array_equal(ht_c, c_)
    tf, cf = vol_f.get_volume()
    ok_f = np.array_equal(tf, t) and np.array_equal(cf, c_) and np.array_equal(vol_f._weight_vol_cpu, w)
    hf_t, hf_w, hf_c = ht_f.get_volume_with_weight()
    ok_f = ok_f and np.array_equal(hf_t, t) and np.array_equal(hf_w, w) and np.array_equal(hf_c, c_)
    tc, cc = vol_c.get_volume()
    ok_f = ok_f and np.array_equal(tc, t) and np.array_equal(cc, c_) and np.array_equal(vol_c._weight_vol_cpu, w)
    ts, cs = vol_s.get_volume()
    ok_f = ok_f and np.array_equal(ts, t) and np.array_equal(cs, c_) and np.array_equal(vol_s._weight_vol_cpu, w)
    fan_v.close(); fan_h.close(); fan_s.close()
    nblk = torch.tensor([ht.stats()["blocks"]], device="cuda")
    dist.all_reduce(nblk)
    want = len(oracle.block_set_from_weight(orc._weight_vol_cpu))
    # the per-voxel view over the sharded map: entry count == voxels the oracle updated
    ok_h = ok_h and ht.count_num_hash_entries() == int((orc._weight_vol_cpu > 0).sum())
    flag = torch.tensor([int(ok and ok_h and ok_f and int(nblk.item()) == want)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("multi_gpu_check world=%d dense=%s hash=%s fanout=%s blocks=%d/%d -> %s" % (
            world, ok, ok_h, ok_f, int(nblk.item()), want, "OK" if flag.item() == 1 else "FAIL"))
    dist.barrier()
    dist.destroy_process_group()
    return 0 if flag.item() == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
