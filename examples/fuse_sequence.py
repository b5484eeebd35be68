#!/usr/bin/env python
"""Fuse an RGB-D sequence into the dense grid and/or the voxel-block hash map and report time and
space of both -- the comparison the reference thesis repo exists for (reference README.md:3-5;
protocol of grid_fusion_demos/grid_demo1.py:44-100 and hash_fusion_demos/hash_demo1.py:92-138:
bounds pass over all frames, one integrate() per frame in order, "Average FPS", mesh + point cloud).

    python examples/fuse_sequence.py --data tests/golden/frames --frames 0-9 --voxel 0.02 --map both --ply out/

The sequence directory holds frame-%06d.{color.jpg,depth.png,pose.txt} and camera-intrinsics.txt
(7-Scenes layout).  Only the import line differs from code written against the reference modules.
"""
import argparse
import os
import sys
import time

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from union_thesis_slam_b200 import grid_fusion, hash_fusion  # noqa: E402  (reference: import grid_fusion, hash_fusion)


def frame_ids(spec):
    out = []
    for part in spec.split(","):
        if "-" in part:
            a, b = part.split("-")
            out.extend(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return out


def read_depth(data, i):
    depth_im = cv2.imread(os.path.join(data, "frame-%06d.depth.png" % i), -1).astype(float)
    depth_im /= 1000.  # depth is saved in 16-bit PNG in millimetres
    depth_im[depth_im == 65.535] = 0  # invalid depth of the 7-scenes data
    return depth_im


def sequence_bounds(data, ids, cam_intr):
    """Axis-aligned hull of every frame's view frustum (apex + image corners at the frame's largest
    depth), starting from the origin like the reference demos do."""
    hull = np.zeros((3, 2))
    for i in ids:
        pose = np.loadtxt(os.path.join(data, "frame-%06d.pose.txt" % i))
        corners = grid_fusion.get_view_frustum(read_depth(data, i), cam_intr, pose)
        hull[:, 0] = np.minimum(hull[:, 0], corners.min(axis=1))
        hull[:, 1] = np.maximum(hull[:, 1], corners.max(axis=1))
    return hull


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "frames"))
    ap.add_argument("--frames", default="0-9")
    ap.add_argument("--voxel", type=float, default=0.02)
    ap.add_argument("--map", default="both", choices=["grid", "hash", "both"])
    ap.add_argument("--map-size", type=int, default=2000000)
    ap.add_argument("--raw", action="store_true", help="upload the uint16 depth and convert on the device")
    ap.add_argument("--ply", default=None, help="directory for mesh.ply / pc.ply")
    ap.add_argument("--per-voxel-frames", type=int, default=0,
                    help="also run the first N frames through the reference's own per-voxel bucket table on the device "
                         "(HashTable(per_voxel=True)) and print its occupancy line of hash_demo1.py:74-82")
    ap.add_argument("--per-voxel-map-size", type=int, default=1000000)
    args = ap.parse_args()
    ids = frame_ids(args.frames)
    cam_intr = np.loadtxt(os.path.join(args.data, "camera-intrinsics.txt"), delimiter=' ')

    print("Estimating voxel volume bounds...")
    vol_bnds = sequence_bounds(args.data, ids, cam_intr)

    maps = {}
    if args.map in ("grid", "both"):
        print("Initializing voxel volume...")
        maps["grid"] = grid_fusion.TSDFVolume(vol_bnds.copy(), voxel_size=args.voxel)
    if args.map in ("hash", "both"):
        print("Initializing hash table...")
        maps["hash"] = hash_fusion.HashTable(vol_bnds.copy(), voxel_size=args.voxel, map_size=args.map_size)

    report = {}
    for name, m in maps.items():
        t0_elapse = time.time()
        t_int = 0.0
        for n, i in enumerate(ids):
            print("Fusing frame %d/%d" % (n + 1, len(ids)))
            color_image = cv2.cvtColor(cv2.imread(os.path.join(args.data, "frame-%06d.color.jpg" % i)), cv2.COLOR_BGR2RGB)
            cam_pose = np.loadtxt(os.path.join(args.data, "frame-%06d.pose.txt" % i))
            tic = time.perf_counter()
            if args.raw:
                raw = cv2.imread(os.path.join(args.data, "frame-%06d.depth.png" % i), -1)
                m.integrate_raw(color_image, raw, cam_intr, cam_pose, obs_weight=1.)
            else:
                m.integrate(color_image, read_depth(args.data, i), cam_intr, cam_pose, obs_weight=1.)
            t_int += time.perf_counter() - tic
        m.sync()
        fps = len(ids) / (time.time() - t0_elapse)
        print("Average FPS: {:.2f}".format(fps))
        st = m.stats()
        nvox = int(np.prod(m._vol_dim))
        if name == "grid":
            bricks = int(np.prod((m._vol_dim + 7) // 8))
            mem = bricks * 6144
            report[name] = dict(fps=fps, integrate_s=t_int, updates=st["updated_total"], bytes=mem,
                                note="%d bricks of 8^3 voxels (whole bounded volume, %d voxels)" % (bricks, nvox))
        else:
            mem = st["blocks"] * 6144 + st["table_slots"] * 16
            report[name] = dict(fps=fps, integrate_s=t_int, updates=st["updated_total"], bytes=mem,
                                note="%d blocks allocated of %d in bounds, %d keys in %d slots (load %.3f), longest probe %d"
                                % (st["blocks"], int(np.prod((m._vol_dim + 7) // 8)), st["table_keys"], st["table_slots"],
                                   st["table_keys"] / st["table_slots"], st["max_probe"]))
        if args.ply:
            os.makedirs(args.ply, exist_ok=True)
            print("Saving mesh to mesh.ply...")
            verts, faces, norms, colors = m.get_mesh()
            grid_fusion.meshwrite(os.path.join(args.ply, "mesh_%s.ply" % name), verts, faces, norms, colors)
            print("Saving point cloud to pc.ply...")
            grid_fusion.pcwrite(os.path.join(args.ply, "pc_%s.ply" % name), m.get_point_cloud())

    print("\n== grid vs hash: running time and space (%d frames, %.3g m voxels) ==" % (len(ids), args.voxel))
    for name, r in report.items():
        print("%-5s %8.2f frames/s end to end | integrate %.4f s | %d voxel updates | %8.1f MB voxel state | %s"
              % (name, r["fps"], r["integrate_s"], r["updates"], r["bytes"] / 1e6, r["note"]))
    if args.per_voxel_frames > 0:
        print("Initializing per-voxel hash table (reference layout)...")
        pv = hash_fusion.HashTable(vol_bnds.copy(), voxel_size=args.voxel, map_size=args.per_voxel_map_size, per_voxel=True)
        for n, i in enumerate(ids[:args.per_voxel_frames]):
            color_image = cv2.cvtColor(cv2.imread(os.path.join(args.data, "frame-%06d.color.jpg" % i)), cv2.COLOR_BGR2RGB)
            cam_pose = np.loadtxt(os.path.join(args.data, "frame-%06d.pose.txt" % i))
            tic = time.perf_counter()
            pv.integrate(color_image, read_depth(args.data, i), cam_intr, cam_pose, obs_weight=1.)
            occupied, collisions = pv.get_num_non_empty_bucket(), pv.get_num_collisions()
            print("Integrate frame {} in {} seconds.".format(n + 1, round(time.perf_counter() - tic, 2)))
            print("Frame {}: num occupied buckets: {}; load factor: {}; collisions: {}; collision rate: {}".format(
                n + 1, occupied, pv.get_load_factor(), collisions, collisions / max(occupied, 1)))
        report["per_voxel"] = dict(entries=pv.count_num_hash_entries(), non_empty=pv.get_num_non_empty_bucket(),
                                   collisions=pv.get_num_collisions(), table_size=pv.voxel_table.stats()["table_size"])
        pv.close()
    if len(maps) == 2:
        t, c = maps["grid"].get_volume()
        ht, hc = maps["hash"].get_volume()
        print("dense vs hash get_volume(): tsdf equal = %s, colour equal = %s" % (np.array_equal(t, ht), np.array_equal(c, hc)))
    return report


if __name__ == "__main__":
    main()
