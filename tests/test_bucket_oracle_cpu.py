"""The container oracle (oracle/bucket_oracle.py) against fixtures recorded from the UNMODIFIED
reference container (tests/golden/bucket_golden.json, made by tests/golden/make_bucket_golden.py):
every return value of every op and the final slot layout."""
import json
import os

import pytest

from oracle import bucket_oracle as bo

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "bucket_golden.json")) as fh:
    GOLD = json.load(fh)


def replay(table, case):
    res = []
    for op in case["ops"]:
        if op[0] == "resize":
            table.double_table_size()
            res.append([table.S, -1])
        elif op[0] == bo.OP_ADD:
            res.append(list(table.add(op[1:4])))
        elif op[0] == bo.OP_GET:
            res.append(list(table.find(op[1:4])))
        elif op[0] == bo.OP_REMOVE:
            res.append([table.remove(op[1:4]), -1])
        else:
            res.append(list(table.integrate(op[1:4], op[4], op[5])))
    return res


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_oracle_replays_reference(case):
    t = bo.OracleBucketTable(case["map_size"])
    assert replay(t, case) == case["results"]
    assert [list(r) for r in t.layout()] == case["layout"]
    st = case["stats"]
    assert (t.S, t.count_num_hash_entries(), t.nonempty, t.get_num_collisions()) == (
        st["table_size"], st["entries"], st["non_empty"], st["collisions"])
    if "voxels_hex" in case:
        got = [[float(x).hex() for x in t.vox[r[0] * t.B + r[1]]] for r in case["layout"]]
        assert got == case["voxels_hex"]


def test_reference_test_expectations():
    """The assertions of the reference's tests/hash_map_test.py hold for the recorded runs."""
    by = {c["name"]: c for c in GOLD["cases"]}
    c = by["test_resize_maintain_num_entries"]
    assert c["stats"]["entries"] == 40 and all(r[0] >= 0 for r in c["results"][41:])
    assert by["test_add_until_full_size_10"]["stats"]["entries"] == 40
    assert by["test_add"]["stats"]["entries"] == 40
    c = by["test_remove_all_entries_full_size_10"]
    assert c["stats"]["entries"] == 0 and c["stats"]["non_empty"] == 0
    assert by["test_add_same_hash_value_400"]["stats"]["entries"] == 400


def test_voxel_integrate_first_observation():
    # voxel.py:29-37 on a fresh voxel: sdf = dist, weight = 1, colour = the observed colour
    assert bo.voxel_integrate((1.0, 0.0, 0.0), 0.25, 65536.0 * 12 + 256 * 34 + 56) == (0.25, 65536.0 * 12 + 256 * 34 + 56, 1.0)
