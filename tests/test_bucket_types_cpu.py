"""Host-side value types of the container mirror (union_thesis_slam_b200/bucket_table.py: Bucket,
HashEntry, Voxel), exercised the way the reference's tests/bucket_test.py exercises its own."""
import numpy as np
import pytest

from union_thesis_slam_b200.bucket_table import Bucket, HashEntry, Voxel


@pytest.fixture
def entries():
    pos = [[1, 2, 3], [2, 0, 5], [3, 9, 4], [4, 4, 3], [5, 7, 3]]
    return [HashEntry(np.array(p), None, Voxel(None, None, None)) for p in pos]


def test_is_full(entries):
    bucket = Bucket(5)
    assert not bucket.is_full()
    for e in entries:
        bucket.add_hash_entry(e)
    assert bucket.is_full()


def test_is_overflow(entries):
    bucket = Bucket(5)
    assert not bucket.is_overflow()
    for e in entries:
        bucket.add_hash_entry(e)
    assert bucket.get_num_entry_stored() == 5 and not bucket.is_overflow()
    bucket.get_ith_entry(-1).set_offset((2, 3))
    assert bucket.is_overflow()                      # last entry has a non-empty offset
    bucket.remove_hash_entry(np.array([4, 4, 3]))
    assert bucket.get_num_entry_stored() == 4
    assert bucket.is_overflow()                      # overflow survives a bucket that is no longer full


def test_add(entries):
    bucket = Bucket(5)
    assert bucket.get_num_entry_stored() == 0
    for i, e in enumerate(entries):
        assert bucket.add_hash_entry(e) == i         # index where the entry is stored
    dup = HashEntry(np.array([1, 2, 3]), None, Voxel(None, None, None))
    assert bucket.add_hash_entry(dup) == -1          # full
    bucket.remove_hash_entry(np.array([1, 2, 3]))
    bucket.remove_hash_entry(np.array([3, 9, 4]))
    assert bucket.add_hash_entry(dup) == 0           # first free slot


def test_get_remove_contains(entries):
    bucket = Bucket(5)
    for e in entries:
        bucket.add_hash_entry(e)
    assert entries[1].equals(bucket.get_hash_entry(np.array([2, 0, 5])))
    assert bucket.get_hash_entry(np.array([1, 5, 2])) is None
    assert bucket.remove_hash_entry(np.array([0, 0, 1])) == -1
    assert bucket.remove_hash_entry(np.array([3, 9, 4])) == 2
    assert bucket.contains(HashEntry(np.array([1, 2, 3]), None, None))
    assert not bucket.contains(HashEntry(np.array([6, 7, 8]), None, None))
    assert not bucket.is_full_for_alien_entries() and bucket.add_alien_hash_entry(entries[2]) == 2
    assert bucket.is_full_for_alien_entries()


def test_voxel_defaults():
    v = Voxel()
    assert (v.get_sdf(), v.get_color(), v.get_weight()) == (1, 0, 0)   # data_structures/voxel.py:8
