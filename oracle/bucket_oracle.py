"""CPU oracle of the reference's per-voxel hash CONTAINER -- TEST INFRASTRUCTURE ONLY.

Flat-array restatement of the bucket / offset-chain container of the reference
(/root/reference/hash_fusion.py:147-180, 192-440; data_structures/bucket.py;
data_structures/hash_entry.py; data_structures/voxel.py:19-49).  The reference stores Python
objects (``Bucket`` lists of ``HashEntry``); here slot ``(b, e)`` of the table is index
``b * bucket_size + e`` of three parallel lists (position, offset, voxel payload), which is the
layout the device container (csrc/bucket_kernels.cu) uses.  Every method follows the reference
statement by statement, including its quirks (no de-duplication on insert, chain members living
in other buckets' slots 0..3, ``last_entry`` being the last non-empty slot of the home bucket,
resize re-adding entries through ``add_hash_entry`` and therefore resizing recursively).

Where the reference would raise on a dangling chain pointer (``None.match_position``), this
restatement raises ``DanglingChain``; the device container reports an error code instead.

Parity status: pinned against the unmodified reference by tests/test_bucket_oracle_cpu.py
(fixtures tests/golden/bucket_*.json produced by tests/golden/make_bucket_golden.py, which runs
the reference's own container on the op sequences of its tests/hash_map_test.py and on seeded
random sequences and records every return value and the final slot layout).
Only tests/ may import this module.
"""
import math

P1 = 73856093
P2 = 19349669
P3 = 83492791

OP_ADD, OP_GET, OP_REMOVE, OP_INTEGRATE = 0, 1, 2, 3


class DanglingChain(RuntimeError):
    pass


class OracleBucketTable:
    def __init__(self, table_size, bucket_size=5, load_factor=0.75):
        self.S = int(table_size)              # hash_fusion.py:36
        self.B = int(bucket_size)             # hash_fusion.py:39
        self.load_factor = load_factor        # hash_fusion.py:38 (the reference hard-codes 0.75)
        self.pos = [None] * (self.S * self.B)
        self.off = [None] * (self.S * self.B)
        self.vox = [None] * (self.S * self.B)  # (sdf, color, weight) or None
        self.exists = [False] * self.S        # _hash_table[b] is a Bucket object (not None)
        self.nonempty = 0                     # _num_non_empty_bucket
        self.resizes = 0

    # -- bucket helpers (data_structures/bucket.py) ---------------------------------------------
    def _i(self, b, e):
        return b * self.B + e

    def _count(self, b):                      # bucket.py:62-70
        return sum(1 for e in range(self.B) if self.pos[self._i(b, e)] is not None)

    def _bucket_none(self, b):
        # A Bucket object normally exists iff it holds an entry; the one exception (a home slot
        # refilled with a dangling chain pointer's None, remove case 1b) is why this is a flag.
        return not self.exists[b]

    def _first_free(self, b):                 # bucket.py:72-82 add_hash_entry
        for e in range(self.B):
            if self.pos[self._i(b, e)] is None:
                return e
        return -1

    def _full_for_alien(self, b):             # bucket.py:34-41
        return all(self.pos[self._i(b, e)] is not None for e in range(self.B - 1))

    def _put(self, b, e, entry):
        i = self._i(b, e)
        self.pos[i], self.off[i], self.vox[i] = entry

    def _take(self, b, e):
        i = self._i(b, e)
        return (self.pos[i], self.off[i], self.vox[i])

    def _clear(self, b, e):
        i = self._i(b, e)
        self.pos[i] = self.off[i] = self.vox[i] = None

    def _entry_at(self, b, e):                # hash_fusion.py:312-320 _get_hash_entry
        """Slot index of the entry, or None when the bucket does not exist / the slot is empty."""
        i = self._i(b, e)
        return i if self.pos[i] is not None else None

    # -- hash_fusion.py:182-190 ----------------------------------------------------------------
    def hash_function(self, p):
        return ((int(p[0]) * P1) ^ (int(p[1]) * P2) ^ (int(p[2]) * P3)) % self.S

    # -- hash_fusion.py:147-180 ----------------------------------------------------------------
    def count_num_hash_entries(self):
        return sum(1 for p in self.pos if p is not None)

    def needs_resize(self):
        return self.nonempty / self.S >= self.load_factor

    def get_num_collisions(self):
        return sum(1 for b in range(self.S) if self.exists[b] and self._count(b) > 1)

    # -- hash_fusion.py:199-224 ----------------------------------------------------------------
    def add(self, position, voxel=None, _entry=None):
        entry = _entry if _entry is not None else (tuple(int(v) for v in position), None, voxel)
        if self.needs_resize():
            self.double_table_size()
        h = self.hash_function(entry[0])
        if self._bucket_none(h):
            self.exists[h] = True
            self._put(h, 0, entry)
            self.nonempty += 1
            return h, 0
        e = self._first_free(h)
        if e < 0:                             # bucket.is_full()
            return self._add_to_linked_list(h, entry)
        self._put(h, e, entry)
        return h, e

    def _add_to_linked_list(self, h, entry):  # hash_fusion.py:226-239
        last = self._i(h, self.B - 1)
        if self.off[last] is None:
            return self._find_next_free_space_and_add_entry(h, self.B - 1, last, entry)
        b, e = self.off[last]
        nxt = self._entry_at(b, e)
        if nxt is None:
            raise DanglingChain()
        while self.off[nxt] is not None:
            b, e = self.off[nxt]
            nxt = self._entry_at(b, e)
            if nxt is None:
                raise DanglingChain()
        return self._find_next_free_space_and_add_entry(b, e, nxt, entry)

    def _find_next_free_space_and_add_entry(self, begin_b, begin_e, prev, entry):  # :241-276
        b_idx = begin_b
        if begin_e < self.B - 1:
            for i in range(begin_e, self.B - 1):
                if self.pos[self._i(b_idx, i)] is None:
                    self._put(b_idx, i, entry)
                    self.off[prev] = (begin_b, i)
                    return begin_b, i
        scanned = False
        while not scanned:
            b_idx = 0 if b_idx >= self.S - 1 else b_idx + 1
            if self._bucket_none(b_idx):
                self.exists[b_idx] = True
                self._put(b_idx, 0, entry)
                self.nonempty += 1
                self.off[prev] = (b_idx, 0)
                return b_idx, 0
            if self._full_for_alien(b_idx):
                scanned = b_idx == begin_b
            else:
                e = self._first_free(b_idx)
                self._put(b_idx, e, entry)
                self.off[prev] = (b_idx, e)
                return b_idx, e
        return -1, -1

    # -- hash_fusion.py:285-310 ----------------------------------------------------------------
    def find(self, position):
        """(bucket, slot) of the entry get_hash_entry returns, or (-1, -1)."""
        p = tuple(int(v) for v in position)
        h = self.hash_function(p)
        if self._bucket_none(h) or self._count(h) == 0:
            return -1, -1
        for e in range(self.B):               # bucket.get_hash_entry: first matching slot
            if self.pos[self._i(h, e)] == p:
                return h, e
        last = self._i(h, self.B - 1)
        if self.pos[last] is not None and self.off[last] is not None:   # bucket.is_overflow()
            b, e = self.off[last]
            while True:
                nxt = self._entry_at(b, e)
                if nxt is None:
                    raise DanglingChain()
                if self.pos[nxt] == p:
                    return b, e
                if self.off[nxt] is None:
                    break
                b, e = self.off[nxt]
        return -1, -1

    def get_voxel(self, position):
        b, e = self.find(position)
        return None if b < 0 else self.vox[self._i(b, e)]

    # -- hash_fusion.py:331-393 ----------------------------------------------------------------
    def _remove_at(self, b, e):               # _remove_hash_entry :386-393
        if not self._bucket_none(b):
            self._clear(b, e)
            if self._count(b) == 0:
                self.exists[b] = False
                self.nonempty -= 1

    def remove(self, position):
        p = tuple(int(v) for v in position)
        h = self.hash_function(p)
        if self._bucket_none(h) or self._count(h) == 0:
            return 0
        last_entry = None
        for i in range(self.B):               # case 1: the entry is in its home bucket
            s = self._i(h, i)
            if self.pos[s] is None:
                continue
            last_entry = s
            if self.pos[s] == p:
                if self.off[s] is None:       # 1a
                    self._clear(h, i)
                    if self._count(h) == 0:
                        self.exists[h] = False
                        self.nonempty -= 1
                    return 1
                pb, pe = self.off[s]          # 1b: pull the next chain element into this slot
                nxt = self._entry_at(pb, pe)
                moved = self._take(pb, pe) if nxt is not None else (None, None, None)
                self._remove_at(pb, pe)
                self._put(h, i, moved)
                return 1
        if last_entry is not None:            # case 2: the entry is further down the chain
            cur = last_entry
            while self.off[cur] is not None:
                pb, pe = self.off[cur]
                nxt = self._entry_at(pb, pe)
                if nxt is None:
                    raise DanglingChain()
                if self.pos[nxt] == p:
                    self.off[cur] = self.off[nxt]   # 2a: None, 2b: the element after it
                    self._remove_at(pb, pe)
                    return 1
                cur = nxt
        return 0

    # -- hash_fusion.py:414-437 ----------------------------------------------------------------
    def double_table_size(self):
        old_pos, old_vox = self.pos, self.vox
        old_n = self.S * self.B
        self.S *= 2
        self.resizes += 1
        self.pos = [None] * (self.S * self.B)
        self.off = [None] * (self.S * self.B)
        self.vox = [None] * (self.S * self.B)
        self.exists = [False] * self.S
        self.nonempty = 0
        for i in range(old_n):                # buckets in order, slots in order
            if old_pos[i] is not None:
                self.add(None, _entry=(old_pos[i], None, old_vox[i]))

    # -- hash_fusion.py:135-145 + data_structures/voxel.py:19-49 ---------------------------------
    def integrate(self, position, dist, color):
        """One iteration of the reference's per-voxel loop: look the voxel up; miss -> new Voxel,
        integrate, insert; hit -> integrate in place.  dist: float64; color: the folded colour of
        the pixel (an integer-valued float, hash_fusion.py:109-110).  Returns the slot (b, e)."""
        b, e = self.find(position)
        if b < 0:
            v = voxel_integrate((1.0, 0.0, 0.0), dist, color)
            return self.add(position, voxel=v)
        i = self._i(b, e)
        self.vox[i] = voxel_integrate(self.vox[i], dist, color)
        return b, e

    # -- read-out --------------------------------------------------------------------------------
    def layout(self):
        """[(bucket, slot, x, y, z, off_bucket, off_slot)] of every stored entry, in slot order
        (off = -1, -1 for an empty offset)."""
        out = []
        for i, p in enumerate(self.pos):
            if p is not None:
                o = self.off[i] if self.off[i] is not None else (-1, -1)
                out.append((i // self.B, i % self.B, p[0], p[1], p[2], o[0], o[1]))
        return out


def voxel_integrate(vox, new_dist, new_color, obs_weight=1.0):
    """data_structures/voxel.py:19-49 on (sdf, color, weight), all float64 (the colour operands
    are integers < 2^24, so the reference's mixed float32/float64 evaluation is exact up to the
    final division, which it performs in float64)."""
    d_old, c_old, w_old = float(vox[0]), float(vox[1]), float(vox[2])
    new_dist = float(new_dist)
    new_color = float(new_color)
    w_new = w_old + obs_weight
    d_new = (d_old * w_old + new_dist * obs_weight) / w_new
    old_b = math.floor(c_old / 65536.0)
    old_g = math.floor((c_old - old_b * 65536.0) / 256.0)
    old_r = c_old - old_b * 65536.0 - old_g * 256.0
    new_b = math.floor(new_color / 65536.0)
    new_g = math.floor((new_color - new_b * 65536.0) / 256.0)
    new_r = new_color - new_b * 65536.0 - new_g * 256.0

    def blend(o, n):
        q = (w_old * o + obs_weight * n) / w_new
        return min(255.0, float(round(q)))    # Python round == numpy round: half to even

    c_new = blend(old_b, new_b) * 65536.0 + blend(old_g, new_g) * 256.0 + blend(old_r, new_r)
    return (d_new, c_new, w_new)
