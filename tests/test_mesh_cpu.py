"""Marching-cubes table and reference pass on CPU (-m "not gpu")."""
import os

import numpy as np

import mc_reference
from union_thesis_slam_b200 import mc_tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_table_matches_generator():
    t = mc_tables.build_table()
    rows = []
    for line in open(os.path.join(ROOT, "union_thesis_slam_b200", "csrc", "mc_table.inc")):
        if line.startswith("{"):
            rows.append([int(v) for v in line.strip().strip("{},").split(",")])
    assert np.array_equal(np.array(rows, np.int8), t)


def test_table_invariants():
    t = mc_tables.build_table()
    assert t.shape == (256, 16) and (t[0] == -1).all() and (t[255] == -1).all()
    for case in range(256):
        crossed = {i for i, (a, b) in enumerate(mc_tables.EDGES) if ((case >> a) & 1) != ((case >> b) & 1)}
        used = {int(e) for e in t[case] if e >= 0}
        assert used == crossed
        n = int((t[case] >= 0).sum())
        assert n % 3 == 0 and n // 3 <= mc_tables.MAX_TRIS
        # the complementary case cuts the same edges
        assert {int(e) for e in t[255 - case] if e >= 0} == crossed
    # single inside corner -> one triangle on its three edges
    assert sorted(int(e) for e in t[1] if e >= 0) == [0, 4, 8]


def test_reference_pass_is_watertight_and_oriented():
    for seed in range(4):
        f = mc_reference.blobs((14, 12, 13), seed)
        verts, faces = mc_reference.marching_cubes(f)
        assert len(faces) > 50
        assert mc_reference.is_closed_manifold(faces)
        # every vertex lies on a sign change: the interpolated field is ~0 there
        ind = np.floor(verts).astype(int)
        assert (ind >= 0).all() and (ind < np.array(f.shape)).all()
        # orientation: face normals point from inside (negative) to outside (positive)
        p = verts[faces]
        n = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
        c = p.mean(1)
        eps = 0.25 * n / np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-12)
        probe = lambda q: f[tuple(np.clip(np.rint(q).astype(int), 0, np.array(f.shape) - 1).T)]
        ok = np.linalg.norm(n, axis=1) > 1e-4
        assert (probe(c + 2 * eps)[ok] >= probe(c - 2 * eps)[ok]).mean() > 0.95
        # closed surface of a union of blobs: Euler characteristic 2 per connected component
        V, E, F = len(verts), len(faces) * 3 // 2, len(faces)
        assert (V - E + F) % 2 == 0 and V - E + F >= 2
