"""The frame fan-out's ordering rules, checked on an executable model (tests/fanout_model.py) under random
interleavings: for any source schedule and ring size every rank's preparation kernel reads the frame it is meant to
read, and nothing deadlocks (-m "not gpu").  The GPU tests run the real thing on the rotating schedule only."""
import random

import pytest

from fanout_model import Model


def _schedules(world, n, rng):
    yield [k % world for k in range(n)]                    # dist.FrameFanout's default: rotating uploads
    yield [0] * n                                          # one rank owns the camera
    yield [rng.randrange(world) for _ in range(n)]         # arbitrary
    yield [(k // 3) % world for k in range(n)]             # bursts from one rank


@pytest.mark.parametrize("world", [1, 2, 3, 4])
@pytest.mark.parametrize("slots", [1, 2, 3, 4])
def test_every_rank_prepares_the_right_frame_under_random_interleavings(world, slots):
    rng = random.Random(1000 * world + slots)
    n = 14
    for sources in _schedules(world, n, rng):
        for seed in range(12):
            for lead in (None, 1, 3):
                m = Model(world, slots, sources, seed)
                errors = m.run(host_lead=lead)
                assert not errors, (world, slots, sources, seed, lead, errors[:3])
                for R in m.ranks:
                    assert [k for k, _ in R.prepared] == list(range(n))


def test_the_model_catches_a_missing_guard():
    """Sanity of the checker itself: without the 'every peer has pulled' guard a source overwrites its slot under
    a slow peer, and the model reports it."""
    class NoGuard(Model):
        def runnable(self, R, op):
            return True if op[0] == "guard" else super().runnable(R, op)
    bad = 0
    for seed in range(40):
        bad += bool(NoGuard(2, 1, [0] * 10, seed).run())
    assert bad > 0
