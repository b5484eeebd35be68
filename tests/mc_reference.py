"""NumPy/Python restatement of the marching-cubes pass (test helper, small volumes only).

Uses the generated case table (union_thesis_slam_b200/mc_tables.py) with the same conventions as
csrc/mesh_kernels.cu: level 0, inside = value < 0, one vertex per crossed grid edge owned by the
voxel at its lower end, vertices enumerated in voxel order then axis order, faces in cell order.
"""
import numpy as np

from union_thesis_slam_b200 import mc_tables


def marching_cubes(tsdf):
    X, Y, Z = tsdf.shape
    table = mc_tables.build_table()
    inside = tsdf < 0
    vid = {}
    verts = []
    for x in range(X):
        for y in range(Y):
            for z in range(Z):
                v0 = tsdf[x, y, z]
                for a, (dx, dy, dz) in enumerate(((1, 0, 0), (0, 1, 0), (0, 0, 1))):
                    x1, y1, z1 = x + dx, y + dy, z + dz
                    if x1 < X and y1 < Y and z1 < Z and inside[x, y, z] != inside[x1, y1, z1]:
                        v1 = tsdf[x1, y1, z1]
                        s = np.float32(0.0 - v0) / np.float32(v1 - v0)
                        p = np.array([x, y, z], np.float32)
                        p[a] += s
                        vid[(x, y, z, a)] = len(verts)
                        verts.append(p)
    faces = []
    for x in range(X - 1):
        for y in range(Y - 1):
            for z in range(Z - 1):
                case = 0
                for i in range(8):
                    if inside[x + (i & 1), y + ((i >> 1) & 1), z + ((i >> 2) & 1)]:
                        case |= 1 << i
                row = table[case]
                for k in range(0, 15, 3):
                    if row[k] < 0:
                        break
                    tri = []
                    for e in row[k:k + 3]:
                        a, b = mc_tables.EDGES[e]
                        axis = {1: 0, 2: 1, 4: 2}[a ^ b]
                        tri.append(vid[(x + (a & 1), y + ((a >> 1) & 1), z + ((a >> 2) & 1), axis)])
                    faces.append(tri)
    return (np.array(verts, np.float32).reshape(-1, 3), np.array(faces, np.int32).reshape(-1, 3))


def is_closed_manifold(faces):
    """Every undirected edge is used by exactly two triangles, once in each direction."""
    from collections import Counter
    directed = Counter()
    for a, b, c in faces:
        for e in ((a, b), (b, c), (c, a)):
            directed[e] += 1
    for (a, b), n in directed.items():
        if n != 1 or directed.get((b, a), 0) != 1:
            return False
    return True


def blobs(shape, seed, n=3):
    """Signed field: negative inside a union of spheres, positive outside, positive border."""
    rng = np.random.default_rng(seed)
    g = np.stack(np.meshgrid(*[np.arange(s) for s in shape], indexing="ij"), -1).astype(np.float32)
    f = np.full(shape, 1.0, np.float32)
    for _ in range(n):
        c = rng.uniform(0.3, 0.7, 3) * np.array(shape)
        r = rng.uniform(0.12, 0.25) * min(shape)
        f = np.minimum(f, (np.linalg.norm(g - c, axis=-1) - r) / 4.0)
    f += rng.uniform(-1e-3, 1e-3, shape).astype(np.float32)  # avoid exact zeros on the lattice
    f[0], f[-1], f[:, 0], f[:, -1], f[:, :, 0], f[:, :, -1] = 1, 1, 1, 1, 1, 1
    return np.clip(f, -1, 1).astype(np.float32)
