"""Seeded small RGB-D scenes for parity tests (shared by make_golden.py and the GPU tests).

The scenes are not physically meaningful renders -- TSDF integration only needs a depth image,
a colour image, intrinsics and a pose -- but they exercise every branch of the per-voxel rule
(SURVEY Appendix A): voxels behind the camera (Z<=0), outside the image, zero-depth holes,
free space (dist clamped to 1), the truncation band, voxels beyond the band, non-orthonormal
poses, non-symmetric intrinsics, float32 depth input and obs_weight != 1.
"""
import numpy as np


def _rot(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def small_case(seed, dims=(37, 29, 41), voxel_size=0.02, hw=(48, 64), nframes=3,
               obs_weights=(1.0,), depth_dtype=np.float64, inside=False, quantise=True):
    """Returns dict(vol_bnds, voxel_size, K, frames=[(color u8, depth, pose, obs)...])."""
    rng = np.random.default_rng(seed)
    H, W = hw
    dims = np.asarray(dims, dtype=np.int64)
    lo = rng.uniform(-1.0, 1.0, 3)
    # upper bound placed strictly inside the last voxel so ceil() reproduces `dims`
    hi = lo + (dims - 0.37) * voxel_size
    vol_bnds = np.stack([lo, hi], axis=1)
    ext = dims * voxel_size
    centre = lo + 0.5 * ext
    f = 0.9 * W
    K = np.array([[f * rng.uniform(0.95, 1.05), 0, W / 2 + rng.uniform(-3, 3)],
                  [0, f * rng.uniform(0.95, 1.05), H / 2 + rng.uniform(-3, 3)],
                  [0, 0, 1.0]])
    frames = []
    vv, uu = np.mgrid[0:H, 0:W]
    for i in range(nframes):
        R = _rot(*rng.uniform(-0.35, 0.35, 3))
        if inside:
            t = centre + rng.uniform(-0.2, 0.2, 3) * ext
        else:
            t = centre - R[:, 2] * (0.55 * ext[2] + rng.uniform(0.1, 0.4)) + rng.uniform(-0.1, 0.1, 3)
        pose = np.eye(4)
        pose[:3, :3] = R + rng.normal(0, 2e-4, (3, 3))  # bundled poses are not orthonormal either
        pose[:3, 3] = t
        base = (0.2 if inside else 0.55) * ext[2] + rng.uniform(0.15, 0.45)
        depth = (base + 0.12 * ext[2] * np.sin(uu / W * rng.uniform(3, 9) + rng.uniform(0, 6))
                 * np.cos(vv / H * rng.uniform(3, 9) + rng.uniform(0, 6))
                 + 0.3 * ext[2] * (uu > W * rng.uniform(0.5, 0.9)))
        if quantise:
            depth = np.round(depth * 1000.0) / 1000.0
        holes = rng.random((H, W)) < 0.07
        depth[holes] = 0.0
        depth[rng.integers(0, H // 2):H // 2 + 4, rng.integers(0, W // 2):W // 2 + 5] = 0.0
        depth = depth.astype(depth_dtype)
        color = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        obs = float(obs_weights[i % len(obs_weights)])
        frames.append((color, depth, pose, obs))
    return dict(vol_bnds=vol_bnds, voxel_size=float(voxel_size), K=K, frames=frames,
                dims=tuple(int(d) for d in dims))


# The cases frozen into tests/golden/small_*.npz (reference outputs stored beside the inputs).
GOLDEN_SMALL_DENSE = [
    dict(seed=11, dims=(37, 29, 41), voxel_size=0.02, nframes=3, obs_weights=(1.0,)),
    dict(seed=12, dims=(33, 40, 27), voxel_size=0.0173, nframes=4, obs_weights=(0.3, 2.5, 0.7, 1.0)),
    dict(seed=13, dims=(24, 24, 53), voxel_size=0.05, nframes=3, obs_weights=(1.0,),
         depth_dtype=np.float32, quantise=False),
    dict(seed=14, dims=(41, 35, 30), voxel_size=0.03, nframes=3, obs_weights=(1.0, 0.7), inside=True),
]
GOLDEN_SMALL_HASH = [
    dict(seed=21, dims=(37, 29, 41), voxel_size=0.02, nframes=3, obs_weights=(1.0,)),
    dict(seed=22, dims=(26, 31, 24), voxel_size=0.03, nframes=3, obs_weights=(1.0,), inside=True),
]

# Config-1 bounds literal, tests/hash_map_test.py:11 of the reference (== the demo's bounds pass).
CONFIG1_BOUNDS = [[-4.22106438, 3.86798203], [-2.6663104, 2.60146141], [0., 5.76272371]]
CONFIG1_VOXEL = 0.02
