"""Golden fixtures for the per-voxel hash CONTAINER, from the UNMODIFIED reference.

    python tests/golden/make_bucket_golden.py

Needs /root/reference (build container only).  Runs the reference ``HashTable`` container methods
(hash_fusion.py:192-440) on (1) the coordinate lists and op patterns of the reference's own
tests/hash_map_test.py, (2) seeded random add / get / remove streams with collisions, duplicates,
chain removals and resizes, (3) the per-voxel integrate loop (hash_fusion.py:135-145) on a small
synthetic stream, and writes tests/golden/bucket_golden.json: for each case the op stream, every
return value and the final slot layout (and voxel payloads).  Refuses to write a case the Python
oracle (oracle/bucket_oracle.py) disagrees with.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import bucket_oracle as bo, ref_runner  # noqa: E402

BOUNDS = [[-4.22106438, 3.86798203], [-2.6663104, 2.60146141], [0., 5.76272371]]  # tests/hash_map_test.py:11

def reference_test_coords():
    """The literal coordinate lists of the reference's tests/hash_map_test.py, read from that file
    (test name -> list of [x, y, z]); they are the reference's fixtures, not ours to restate."""
    import ast
    path = os.path.join(ref_runner.REFERENCE_ROOT, "tests", "hash_map_test.py")
    out = {}
    for node in ast.walk(ast.parse(open(path).read())):
        if isinstance(node, ast.FunctionDef):
            for stmt in node.body:
                if (isinstance(stmt, ast.Assign) and len(stmt.targets) == 1 and isinstance(stmt.targets[0], ast.Name)
                        and stmt.targets[0].id == "world_coords" and isinstance(stmt.value, ast.List) and stmt.value.elts):
                    out[node.name] = ast.literal_eval(stmt.value)
    return out


def ref_layout(ht):
    out = []
    for b, bucket in enumerate(ht._hash_table):
        if bucket is None:
            continue
        for e in range(ht._bucket_size):
            ent = bucket.get_ith_entry(e)
            if ent is None:
                continue
            off = ent.get_offset() or (-1, -1)
            p = ent.get_position()
            out.append([b, e, int(p[0]), int(p[1]), int(p[2]), int(off[0]), int(off[1])])
    return out


def ref_find(ht, entry):
    """(bucket, slot) of a HashEntry object inside the reference table (identity search)."""
    if entry is None:
        return [-1, -1]
    for b, bucket in enumerate(ht._hash_table):
        if bucket is None:
            continue
        for e in range(ht._bucket_size):
            if bucket.get_ith_entry(e) is entry:
                return [b, e]
    raise AssertionError("entry returned by the reference is not in its table")


def run_case(hf, he, vx, name, map_size, ops):
    """ops: list of [op, x, y, z] (+ [dist, color] for OP_INTEGRATE) or ["resize"]."""
    with contextlib.redirect_stdout(io.StringIO()):
        ref = hf.HashTable(np.array(BOUNDS), 0.02, map_size, False)
    orc = bo.OracleBucketTable(map_size)
    results = []
    for op in ops:
        with contextlib.redirect_stdout(io.StringIO()):
            if op[0] == "resize":
                ref.double_table_size()
                orc.double_table_size()
                r = o = [ref._table_size, -1]
                o = [orc.S, -1]
            elif op[0] == bo.OP_ADD:
                r = [int(v) for v in ref.add_hash_entry(he.HashEntry(list(op[1:4]), None, None))]
                o = [int(v) for v in orc.add(op[1:4])]
            elif op[0] == bo.OP_GET:
                r = ref_find(ref, ref.get_hash_entry(list(op[1:4])))
                o = [int(v) for v in orc.find(op[1:4])]
            elif op[0] == bo.OP_REMOVE:
                r = [int(ref.remove_hash_entry(he.HashEntry(list(op[1:4]), None, None))), -1]
                o = [int(orc.remove(op[1:4])), -1]
            elif op[0] == bo.OP_INTEGRATE:
                # hash_fusion.py:135-145 with the loop's own variables
                coord = np.array(op[1:4])
                dist, color = np.float64(op[4]), np.float32(op[5])
                target = ref.get_hash_entry(coord)
                if target is None:
                    voxel = vx.Voxel()
                    voxel.integrate(dist, color)
                    r = [int(v) for v in ref.add_hash_entry(he.HashEntry(coord, None, voxel))]
                else:
                    target.integrate_voxel(dist, color)
                    r = ref_find(ref, target)
                o = [int(v) for v in orc.integrate(op[1:4], op[4], op[5])]
            else:
                raise ValueError(op)
        assert r == o, "%s: op %r: reference %r, oracle %r" % (name, op, r, o)
        results.append(r)
    lay = ref_layout(ref)
    assert lay == [list(t) for t in orc.layout()], "%s: layouts differ" % name
    stats = {"table_size": int(ref._table_size), "entries": int(ref.count_num_hash_entries()),
             "non_empty": int(ref.get_num_non_empty_bucket()), "collisions": int(ref.get_num_collisions()),
             "load_factor": float(ref.get_load_factor())}
    assert stats["entries"] == orc.count_num_hash_entries() and stats["non_empty"] == orc.nonempty
    assert stats["collisions"] == orc.get_num_collisions() and stats["table_size"] == orc.S
    case = {"name": name, "map_size": map_size, "ops": ops, "results": results, "layout": lay, "stats": stats}
    if any(op[0] == bo.OP_INTEGRATE for op in ops):
        vox = []
        for b, e, *_ in lay:
            v = ref._hash_table[b].get_ith_entry(e).get_voxel()
            ov = orc.vox[b * orc.B + e]
            rv = [float(v.get_sdf()), float(v.get_color()), float(v._weight)]
            assert rv == [float(t) for t in ov], "%s: voxel at (%d,%d): %r vs %r" % (name, b, e, rv, ov)
            vox.append([x.hex() for x in rv])
        case["voxels_hex"] = vox
    return case


def main():
    _, hf = ref_runner.load()
    from data_structures import hash_entry as he, voxel as vx
    A, G, R, I = bo.OP_ADD, bo.OP_GET, bo.OP_REMOVE, bo.OP_INTEGRATE
    cases = []
    # (1) the reference's own tests
    coords = reference_test_coords()
    COORDS_RESIZE = coords["test_resize_maintain_num_entries"]
    COORDS_FULL10 = coords["test_add_until_full_size_10"]
    COORDS_ADD = coords["test_add"]
    assert len(COORDS_RESIZE) == len(COORDS_FULL10) == len(COORDS_ADD) == 40
    assert coords["test_remove_all_entries_full_size_10"] == COORDS_ADD
    ops = [[A] + c for c in COORDS_RESIZE] + [["resize"]] + [[G] + c for c in COORDS_RESIZE]
    cases.append(run_case(hf, he, vx, "test_resize_maintain_num_entries", 10, ops))
    cases.append(run_case(hf, he, vx, "test_add_until_full_size_10", 10, [[A] + c for c in COORDS_FULL10]))
    cases.append(run_case(hf, he, vx, "test_add", 10, [[A] + c for c in COORDS_ADD]))
    cases.append(run_case(hf, he, vx, "test_remove_all_entries_full_size_10", 10,
                          [[A] + c for c in COORDS_ADD] + [[R] + c for c in COORDS_ADD]))
    cases.append(run_case(hf, he, vx, "test_add_same_hash_value_400", 100, [[A, 0, 0, 0]] * 400))
    rng = np.random.default_rng(7)
    ops = [[A, int(rng.integers(500)), int(rng.integers(500)), i] for i in range(450)]
    cases.append(run_case(hf, he, vx, "test_add_until_full_size_100", 100, ops))
    # test_general at 1/20 scale: add 4n, remove 2n at random, look the rest up, add 2n more
    n = 500
    coords = [[int(rng.integers(500)), int(rng.integers(500)), i] for i in range(4 * n)]
    ops = [[A] + c for c in coords]
    left = list(coords)
    for _ in range(2 * n):
        c = left.pop(int(rng.integers(len(left) - 1)))
        ops.append([R] + c)
    ops += [[G] + c for c in left]
    ops += [[A, int(rng.integers(500)), int(rng.integers(500)) + 501, i] for i in range(2 * n)]
    cases.append(run_case(hf, he, vx, "test_general_n500", n, ops))
    # (2) seeded random streams on small tables: heavy collisions, duplicates, misses, chain removals
    for seed, size, nops, span in ((11, 7, 300, 6), (12, 16, 600, 9), (13, 64, 1500, 14), (14, 5, 200, 4)):
        rng = np.random.default_rng(seed)
        ops = []
        for _ in range(nops):
            c = [int(v) for v in rng.integers(-span, span, 3)]
            ops.append([int(rng.choice([A, A, A, G, R, R])),] + c)
        cases.append(run_case(hf, he, vx, "random_seed%d_size%d" % (seed, size), size, ops))
    # (2b) crafted collisions: chain look-ups, removal from the middle / tail / head of a chain
    # (hash_fusion.py:350-384 cases 1b, 2a, 2b), chains passing through a neighbour's free slots
    def colliding(size, bucket, count, start=0):
        out, v = [], start
        t = bo.OracleBucketTable(size)
        while len(out) < count:
            p = [v % 97, (v // 97) % 89, v // (97 * 89)]
            if t.hash_function(p) == bucket:
                out.append(p)
            v += 1
        return out
    P = colliding(64, 17, 14)
    Q = colliding(64, 18, 7)
    ops = [[A] + p for p in P] + [[A] + q for q in Q] + [[G] + p for p in P + Q]
    ops += [[R] + P[7], [R] + P[13], [R] + P[4], [R] + P[0], [R] + Q[5]] + [[G] + p for p in P + Q]
    ops += [[A] + P[7], [A] + P[0], [R] + P[9], [R] + P[5], [R] + Q[0]] + [[G] + p for p in P + Q]
    cases.append(run_case(hf, he, vx, "crafted_chain_remove_size64", 64, ops))
    rng = np.random.default_rng(31)
    pool = colliding(32, 5, 12) + colliding(32, 6, 8) + colliding(32, 7, 6)
    ops = [[int(rng.choice([A, A, G, R])),] + pool[int(rng.integers(len(pool)))] for _ in range(700)]
    cases.append(run_case(hf, he, vx, "crafted_chain_random_size32", 32, ops))
    # (2c) a resize whose re-insertion crosses the load factor again (add_hash_entry -> double_table_size
    # -> add_hash_entry -> double_table_size, hash_fusion.py:207-208 inside :425-436)
    # Buckets 1..7 of a 10-bucket table hold five entries each, chosen so that both of their residues
    # mod 20 occur early; bucket 0 holds one.  That is 8 non-empty buckets (>= 7.5), so the next insert
    # doubles the table, and while the 36 entries are re-added the 20-bucket table reaches 15 non-empty
    # buckets before the last of them: the re-add of the next entry doubles again, inside the first resize.
    ops = []
    for b in range(1, 8):
        lo, hi = colliding(20, b, 3), colliding(20, b + 10, 2)
        ops += [[A] + p for p in (lo[0], hi[0], lo[1], hi[1], lo[2])]
    ops += [[A] + colliding(20, 0, 1)[0]]
    ops += [[A, int(rng.integers(50)), int(rng.integers(50)), 1000 + i] for i in range(12)]
    ops += [[G] + op[1:4] for op in ops[:48]]
    cases.append(run_case(hf, he, vx, "crafted_nested_resize_size10", 10, ops))
    assert cases[-1]["results"][36][0] >= 0 and cases[-1]["stats"]["table_size"] >= 40
    # (3) the per-voxel integrate loop on a synthetic stream (positions revisited, colours blended)
    rng = np.random.default_rng(21)
    ops = []
    for _ in range(900):
        c = [int(v) for v in rng.integers(0, 7, 3)]
        dist = float(np.minimum(1.0, rng.uniform(-1.0, 1.6)))
        color = float(int(rng.integers(0, 256)) * 65536 + int(rng.integers(0, 256)) * 256 + int(rng.integers(0, 256)))
        ops.append([I] + c + [dist, color])
    cases.append(run_case(hf, he, vx, "integrate_loop_seed21", 40, ops))
    out = os.path.join(HERE, "bucket_golden.json")
    with open(out, "w") as fh:
        json.dump({"generated_with": "reference hash_fusion.HashTable container methods (unmodified)",
                   "bounds": BOUNDS, "cases": cases}, fh, separators=(",", ":"))
    print("wrote %s: %d cases, %d ops, %.0f kB" % (out, len(cases), sum(len(c["ops"]) for c in cases),
                                                   os.path.getsize(out) / 1e3))


if __name__ == "__main__":
    main()
