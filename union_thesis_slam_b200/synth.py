"""Synthetic RGB-D streams for BASELINE.json configs 4 and 5 (no dataset can be downloaded).

An axis-aligned box room rendered analytically: per pixel the camera ray
``R @ [(u-cx)/fx, (v-cy)/fy, 1]`` is intersected with the six walls; because the ray is not
normalised its parameter at the nearest hit IS the z-depth the integration rule compares against
(reference grid_fusion.py:282).  Depth is quantised to whole millimetres and stored as
``mm / 1000.`` float64 like the sensor path of the reference demos (grid_demo1.py:81-83), a
deterministic hole pattern exercises the invalid-depth branch, colour is a procedural checker.

Host-side NumPy only: the frames feed both the CUDA path and the CPU oracle.
"""
import numpy as np


def intrinsics(width=640, height=480):
    """585-px focal length at 640x480 (the bundled sequence's camera), scaled with the width."""
    s = width / 640.0
    return np.array([[585.0 * s, 0.0, width / 2.0], [0.0, 585.0 * s, height / 2.0], [0.0, 0.0, 1.0]])


def look_at(eye, target, up=(0.0, 0.0, 1.0)):
    """Camera->world pose with +z forward, +x right, +y down."""
    eye = np.asarray(eye, float)
    fwd = np.asarray(target, float) - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.asarray(up, float))
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    pose = np.eye(4)
    pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = right, down, fwd, eye
    return pose


def trajectory(n, room, seed=0, margin=0.8):
    """Seeded smooth trajectory inside the room: Lissajous position, look-at target on a second
    Lissajous curve, so consecutive frames overlap like a hand-held scan."""
    rng = np.random.default_rng(seed)
    room = np.asarray(room, float)
    lo, hi = room[:, 0] + margin, room[:, 1] - margin
    ph = rng.uniform(0, 2 * np.pi, 6)
    fr = rng.uniform(0.6, 1.6, 6)
    poses = []
    for i in range(n):
        t = 2 * np.pi * i / max(n, 64)
        a = 0.5 + 0.5 * np.sin(fr[:3] * t + ph[:3])
        b = 0.5 + 0.5 * np.sin(fr[3:] * t * 1.7 + ph[3:])
        eye = lo + a * (hi - lo)
        tgt = room[:, 0] + b * (room[:, 1] - room[:, 0])
        if np.linalg.norm(tgt - eye) < 0.5:
            tgt = tgt + np.array([1.0, 0.3, 0.0])
        poses.append(look_at(eye, tgt))
    return poses


def render_room(pose, K, room, width=640, height=480, holes=True, noise=0, rng=None):
    """Returns (color uint8 [H,W,3], depth float64 [H,W] metres).  noise > 0 adds uniform integer
    sensor noise in [-noise, noise] to every colour channel (seeded by rng)."""
    room = np.asarray(room, float)
    v, u = np.mgrid[0:height, 0:width]
    rays = np.stack([(u - K[0, 2]) / K[0, 0], (v - K[1, 2]) / K[1, 1], np.ones_like(u, float)], -1)
    dirs = rays @ pose[:3, :3].T
    eye = pose[:3, 3]
    best = np.full((height, width), np.inf)
    wall = np.zeros((height, width), np.int64)
    with np.errstate(divide="ignore", invalid="ignore"):
        for ax in range(3):
            for side in range(2):
                s = (room[ax, side] - eye[ax]) / dirs[..., ax]
                ok = (s > 1e-9) & (s < best)
                best = np.where(ok, s, best)
                wall = np.where(ok, 2 * ax + side, wall)
    hit = eye + dirs * best[..., None]
    depth = np.round(best * 1000.0) / 1000.0  # whole millimetres, as mm/1000. float64
    depth[~np.isfinite(depth)] = 0.0
    depth[depth > 65.0] = 0.0
    if holes:
        depth[((u * 7 + v * 13) % 97) == 0] = 0.0  # sparse dropouts
        depth[(u // 40 + v // 40) % 11 == 5] = 0.0  # blocky dropouts
    cell = np.floor(hit / 0.25).astype(np.int64)
    chk = (cell.sum(-1) + wall) & 1
    base = np.array([[200, 60, 60], [60, 200, 60], [60, 60, 200], [200, 200, 60], [200, 60, 200], [60, 200, 200]])
    color = base[wall] + 40 * chk[..., None] + (cell[..., :3] % 8)
    if noise:
        color = color + (rng or np.random.default_rng(0)).integers(-noise, noise + 1, color.shape)
    return np.clip(color, 0, 255).astype(np.uint8), depth


def room_stream(n, room, width=640, height=480, seed=0, holes=True, noise=4):
    """n frames: list of (color, depth, pose) plus the intrinsics.  Colours carry +-noise of
    per-frame sensor noise so that consecutive observations of a voxel never agree exactly."""
    K = intrinsics(width, height)
    frames = []
    rng = np.random.default_rng(seed + 1000)
    for pose in trajectory(n, room, seed):
        c, d = render_room(pose, K, room, width, height, holes, noise, rng)
        frames.append((c, d, pose))
    return frames, K


def room_bounds(size):
    """vol_bnds (3,2) of a room with the given (x, y, z) extent, origin at 0."""
    return np.array([[0.0, size[0]], [0.0, size[1]], [0.0, size[2]]])
