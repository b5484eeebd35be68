"""Device marching cubes (get_mesh / get_point_cloud) on the GPU (-m gpu).

No reference fixture pins mesh output (scikit-image is absent and the reference has no mesh test),
so the checks are structural: identical to the NumPy restatement that uses the same generated
table, watertight, vertices on zero crossings, colours = colour volume at round(vertex) exactly
as grid_fusion.py:329-339 unpacks them.
"""
import numpy as np
import pytest

import mc_reference
import scenes

pytestmark = pytest.mark.gpu


def _extract(tsdf, color=None, vs=1.0, origin=(0, 0, 0)):
    import torch
    from union_thesis_slam_b200 import meshing
    t = torch.from_numpy(np.ascontiguousarray(tsdf)).cuda()
    c = torch.from_numpy(np.ascontiguousarray(color)).cuda() if color is not None else None
    return meshing.extract(t, c, tsdf.shape, vs, np.asarray(origin, np.float32))


@pytest.mark.parametrize("seed,shape", [(0, (14, 12, 13)), (1, (9, 17, 11)), (2, (20, 8, 8))])
def test_device_mesh_equals_reference_pass(seed, shape):
    f = mc_reference.blobs(shape, seed)
    rv, rf = mc_reference.marching_cubes(f)
    v, faces, n, c = _extract(f)
    assert v.shape == rv.shape and faces.shape == rf.shape
    assert np.allclose(v, rv, atol=1e-6) and np.array_equal(faces, rf)
    assert mc_reference.is_closed_manifold(faces)
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-4)
    # gradient normals agree with the face winding (both point towards larger values)
    p = v[faces]
    fn = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    ok = np.linalg.norm(fn, axis=1) > 1e-4
    assert ((fn * n[faces[:, 0]]).sum(1)[ok] > 0).mean() > 0.95


def test_world_coordinates_and_colour_lookup():
    f = mc_reference.blobs((16, 16, 16), 5)
    rng = np.random.default_rng(0)
    bgr = rng.integers(0, 256, f.shape + (3,))
    col = (bgr[..., 0] * 65536 + bgr[..., 1] * 256 + bgr[..., 2]).astype(np.float32)
    vs, origin = 0.02, np.array([-1.0, 0.5, 2.0], np.float32)
    rv, _ = mc_reference.marching_cubes(f)
    v, faces, n, c = _extract(f, col, vs, origin)
    assert np.allclose(v, rv * np.float32(vs) + origin, atol=1e-6)
    ind = np.round(rv).astype(int)  # grid_fusion.py:329
    rgb = col[ind[:, 0], ind[:, 1], ind[:, 2]]
    b = np.floor(rgb / 65536); g = np.floor((rgb - b * 65536) / 256); r = rgb - b * 65536 - g * 256
    assert np.array_equal(c, np.stack([r, g, b], 1).astype(np.uint8))


def test_get_mesh_and_point_cloud_through_the_drop_in_classes(tmp_path):
    from union_thesis_slam_b200 import grid_fusion, hash_fusion, synth
    room = synth.room_bounds((2.0, 1.6, 1.2))
    fr, K = synth.room_stream(3, room, width=160, height=120, seed=2)
    vol = grid_fusion.TSDFVolume(room.copy(), 0.02)
    ht = hash_fusion.HashTable(room.copy(), 0.02, map_size=1 << 14)
    for c, d, p in fr:
        vol.integrate(c, d, K, p)
        ht.integrate(c, d, K, p)
    verts, faces, norms, colors = vol.get_mesh()
    assert len(verts) > 1000 and len(faces) > 1000 and colors.dtype == np.uint8
    assert faces.min() >= 0 and faces.max() < len(verts)
    lo, hi = vol._vol_bnds[:, 0] - 1e-3, vol._vol_bnds[:, 1] + 1e-3
    assert (verts >= lo).all() and (verts <= hi).all()
    # the surface is the room's walls: every vertex sits within ~one voxel of a wall
    d_wall = np.minimum(np.abs(verts - room[:, 0]), np.abs(verts - room[:, 1])).min(1)
    assert np.percentile(d_wall, 95) < 0.04
    pc = vol.get_point_cloud()
    assert pc.shape == (len(verts), 6) and np.allclose(pc[:, :3], verts)
    hv, hf, hn, hc = ht.get_mesh()
    assert np.array_equal(hv, verts) and np.array_equal(hf, faces) and np.array_equal(hc, colors)
    grid_fusion.meshwrite(str(tmp_path / "mesh.ply"), verts[:50], faces[:0], norms[:50], colors[:50])
    grid_fusion.pcwrite(str(tmp_path / "pc.ply"), pc[:50])
    assert open(tmp_path / "mesh.ply").readline().strip() == "ply"
