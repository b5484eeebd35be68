"""CPU-only checks of the C-ABI library and the host-side logic (-m "not gpu").

No compute is launched: the library must load, export every symbol include/tsdf_b200.h
declares, and fail loudly (never fall back) when no CUDA device is present.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "tsdf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:tsdf|vh|vbt)_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__
    __graft_entry__.build()
    from union_thesis_slam_b200 import _capi
    return _capi


def test_header_symbols_are_exported_and_bound(built):
    names = _declared_symbols()
    assert len(names) >= 25 and "tsdf_grid_integrate" in names and "vh_integrate" in names
    lib = C.CDLL(built.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "%s is declared in include/tsdf_b200.h but not exported" % n
    bound = {s[0] for s in built.SYMBOLS}
    assert bound == set(names), "ctypes table and header disagree: %s" % sorted(bound ^ set(names))


def test_library_is_sm100a_only(built):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", built.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def _has_gpu():
    import torch
    return torch.cuda.is_available()


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device failure path")
def test_no_device_fails_loudly(built):
    from union_thesis_slam_b200 import grid_fusion, hash_fusion
    with pytest.raises(built.TSDFError) as ei:
        grid_fusion.TSDFVolume(np.array([[0., 1.], [0., 1.], [0., 1.]]), 0.1)
    assert ei.value.status == -2 and "no CPU fallback" in str(ei.value)
    with pytest.raises(built.TSDFError):
        hash_fusion.HashTable(np.array([[0., 1.], [0., 1.], [0., 1.]]), 0.1)
    with pytest.raises(built.TSDFError):
        built.device_info(0)
    h = C.c_void_p()
    assert built.lib().tsdf_fanout_create(0, 0, 1, 1 << 20, 3, C.byref(h)) == -2 and not h.value


def test_fanout_arguments_are_checked_before_any_device_work(built):
    """Argument errors of the frame fan-out do not need a GPU to be reported."""
    lib, h = built.lib(), C.c_void_p()
    for device, rank, world, nbytes, slots in ((0, 0, 17, 1 << 20, 3), (0, 2, 2, 1 << 20, 3), (0, 0, 1, 1 << 20, 0),
                                               (0, 0, 1, 1 << 20, 9), (0, 0, 1, 0, 3)):
        assert lib.tsdf_fanout_create(device, rank, world, nbytes, slots, C.byref(h)) == -1
        assert not h.value and lib.tsdf_last_error()
    assert lib.tsdf_fanout_create(0, 0, 1, 1 << 20, 3, None) == -1
    assert lib.tsdf_fanout_push(None, 0, None, 0, None, 0, 0) == -1
    assert lib.tsdf_fanout_integrate(None, None, None, 0, 0, None, None, 4, 4, 0, None, None, 1.0, 1000.0, -1) == -1
    assert lib.tsdf_fanout_destroy(None) == 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "union_thesis_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "synth.py", "%s mentions the oracle" % f
    txt = open(os.path.join(pkg, "synth.py")).read()
    assert "import oracle" not in txt and "from oracle" not in txt


def test_ctor_mirrors_reference_quirks(built, capsys):
    """Dims banner, in-place snap of the caller's upper bounds, shape assert (grid_fusion.py:30-49)."""
    from union_thesis_slam_b200 import grid_fusion
    bnds = np.array([[-4.22106438, 3.86798203], [-2.6663104, 2.60146141], [0., 5.76272371]])
    with pytest.raises(AssertionError):
        grid_fusion.TSDFVolume(np.zeros((2, 3)), 0.02)
    try:
        grid_fusion.TSDFVolume(bnds, 0.02)
    except built.TSDFError:
        pass  # no device on the CPU box; everything up to the create call already ran
    out = capsys.readouterr().out
    assert "Voxel volume size: 405 x 264 x 289 - # points: 30,899,880" in out
    assert np.allclose(bnds[:, 1], bnds[:, 0] + np.array([405, 264, 289]) * 0.02)


def test_module_helpers_match_reference_semantics():
    from union_thesis_slam_b200 import grid_fusion
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(7, 3)).astype(np.float32)
    T = np.eye(4); T[:3, :3] = np.linalg.qr(rng.normal(size=(3, 3)))[0]; T[:3, 3] = [0.1, -2.0, 3.0]
    out = grid_fusion.rigid_transform(pts, T)
    assert out.shape == (7, 3) and np.allclose(out, pts @ T[:3, :3].T + T[:3, 3])
    depth = np.zeros((480, 640)); depth[100, 100] = 2.5
    K = np.array([[585., 0, 320], [0, 585., 240], [0, 0, 1]])
    fr = grid_fusion.get_view_frustum(depth, K, np.eye(4))
    assert fr.shape == (3, 5) and np.allclose(fr[:, 0], 0) and np.allclose(fr[2, 1:], 2.5)
    assert np.allclose(fr[0, 1], (0 - 320) * 2.5 / 585) and np.allclose(fr[1, 2], (480 - 240) * 2.5 / 585)


def test_ply_writers(tmp_path):
    from union_thesis_slam_b200 import grid_fusion
    v = np.array([[0., 0, 0], [1, 0, 0], [0, 1, 0]]); f = np.array([[0, 1, 2]])
    n = np.tile([0., 0, 1], (3, 1)); c = np.array([[255, 0, 0], [0, 255, 0], [0, 0, 255]], np.uint8)
    grid_fusion.meshwrite(str(tmp_path / "m.ply"), v, f, n, c)
    lines = open(tmp_path / "m.ply").read().splitlines()
    assert lines[:3] == ["ply", "format ascii 1.0", "element vertex 3"] and lines[-1] == "3 0 1 2"
    assert lines[lines.index("end_header") + 1] == "0.000000 0.000000 0.000000 0.000000 0.000000 1.000000 255 0 0"
    grid_fusion.pcwrite(str(tmp_path / "p.ply"), np.hstack([v, c.astype(float)]))
    lines = open(tmp_path / "p.ply").read().splitlines()
    assert "element face" not in "\n".join(lines) and lines[-1] == "0.000000 1.000000 0.000000 0 0 255"


def test_synthetic_room_render_is_z_depth():
    """Depth must be camera z-depth (the rule compares depth - z_cam, grid_fusion.py:282)."""
    from union_thesis_slam_b200 import synth
    room = synth.room_bounds((4.0, 3.0, 2.5))
    K = synth.intrinsics(160, 120)
    pose = synth.look_at([1.0, 1.0, 1.2], [3.5, 2.0, 1.0])
    color, depth = synth.render_room(pose, K, room, 160, 120, holes=False)
    assert color.shape == (120, 160, 3) and color.dtype == np.uint8 and depth.dtype == np.float64
    v, u = 60, 80
    ray = pose[:3, :3] @ np.array([(u - K[0, 2]) / K[0, 0], (v - K[1, 2]) / K[1, 1], 1.0])
    hit = pose[:3, 3] + ray * depth[v, u]
    on_wall = np.isclose(hit, room[:, 0], atol=2e-3) | np.isclose(hit, room[:, 1], atol=2e-3)
    assert on_wall.any() and np.all(hit > room[:, 0] - 2e-3) and np.all(hit < room[:, 1] + 2e-3)
    assert np.all(np.round(depth * 1000) / 1000 == depth)
    _, d2 = synth.render_room(pose, K, room, 160, 120, holes=True)
    assert (d2 == 0).sum() > 100
