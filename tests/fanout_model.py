"""Executable model of the frame fan-out's ordering rules (csrc/capi_fanout.cu + FrameIO::adopt /
release_upload_after_prep in csrc/capi_grid.cu), for the CPU tests: test infrastructure, not product code.

Every rank has three in-order device streams -- `up` (its uploads), `st` (arrivals) and `pre` (the map's
preparation kernel) -- and a host thread that issues the per-frame calls in program order without ever waiting.
The model executes the queued operations under a seeded random scheduler and checks the two properties the design
has to guarantee for ANY source schedule, ring size and interleaving:

  * the preparation kernel of frame k reads frame k on every rank (no slot is overwritten before its readers --
    the own map, or peers still pulling -- are through, and no reader runs before the data is there);
  * nothing deadlocks.

CUDA event semantics are modelled as they are: a wait captures the most recent record ISSUED (not executed) before
it, and is a no-op when there is none.
"""
import random


class Event:
    def __init__(self):
        self.issued = 0     # records issued so far (host side)
        self.done = 0       # highest record generation that has executed


class Rank:
    def __init__(self, r, world, slots):
        self.r = r
        self.q = {"up": [], "st": [], "pre": []}
        self.slot = [None] * slots                       # frame number resident in the own ring slot
        self.ready = [0] * slots                         # flags in this rank's arena
        self.pulled = [[0] * world for _ in range(slots)]
        self.staged = [Event() for _ in range(slots)]
        self.consumed = [Event() for _ in range(slots)]
        self.ext_ready, self.ext_consumed = Event(), Event()
        self.last_pushed = [-1] * slots                  # host-side bookkeeping of tsdf_fanout
        self.prepared = []                               # (frame, what the preparation kernel read)


class Model:
    def __init__(self, world, slots, sources, seed):
        self.world, self.slots, self.sources = world, slots, list(sources)
        self.rng = random.Random(seed)
        self.ranks = [Rank(r, world, slots) for r in range(world)]
        self.errors = []

    # ---- operations: (kind, payload...) queued on a stream; `ready_to_run` says whether the head may execute ----
    def _record(self, rank, stream, ev):
        ev.issued += 1
        rank.q[stream].append(("record", ev, ev.issued))

    def _wait(self, rank, stream, ev):
        if ev.issued:                                    # no record issued yet: cudaStreamWaitEvent is a no-op
            rank.q[stream].append(("wait", ev, ev.issued))

    def issue_frame(self, r, k):
        """The host calls of rank r for frame k, in the order of tsdf_fanout_integrate."""
        R, s, src = self.ranks[r], k % self.slots, self.sources[k]
        if src == r:                                                     # tsdf_fanout_push
            if self.world > 1 and R.last_pushed[s] >= 0:
                R.q["up"].append(("guard", s, R.last_pushed[s] + 1))
            if k >= self.slots:
                self._wait(R, "up", R.consumed[s])
            R.q["up"].append(("h2d", s, k))
            self._record(R, "up", R.staged[s])
            R.last_pushed[s] = k
            if self.world > 1:
                R.q["up"].append(("signal_ready", s, k + 1))
        if src == r:                                                     # tsdf_fanout_acquire
            self._wait(R, "st", R.staged[s])
        else:
            if self.world > 1 and R.last_pushed[s] >= 0:
                R.q["st"].append(("guard", s, R.last_pushed[s] + 1))
            R.last_pushed[s] = -1
            R.q["st"].append(("wait_ready", s, k + 1))
            R.q["st"].append(("pull", s, src, k))
            R.q["st"].append(("signal_pulled", s, src, k + 1))
        self._record(R, "st", R.ext_ready)                               # frame_source (FrameIO::adopt)
        self._wait(R, "pre", R.ext_ready)
        R.q["pre"].append(("prep", s, k))                                # the map's preparation kernel
        self._record(R, "pre", R.ext_consumed)                           # release_upload_after_prep
        self._wait(R, "st", R.ext_consumed)
        self._record(R, "st", R.consumed[s])                             # tsdf_fanout_release

    def runnable(self, R, op):
        kind = op[0]
        if kind == "wait":
            return op[1].done >= op[2]
        if kind == "guard":
            return all(R.pulled[op[1]][p] >= op[2] for p in range(self.world) if p != R.r)
        if kind == "wait_ready":
            return R.ready[op[1]] >= op[2]
        return True

    def execute(self, R, op):
        kind = op[0]
        if kind == "record":
            op[1].done = max(op[1].done, op[2])
        elif kind == "h2d":
            R.slot[op[1]] = op[2]
        elif kind == "signal_ready":
            for P in self.ranks:
                if P is not R:
                    P.ready[op[1]] = max(P.ready[op[1]], op[2])
        elif kind == "pull":
            got = self.ranks[op[2]].slot[op[1]]
            if got != op[3]:
                self.errors.append("rank %d pulled frame %r from rank %d's slot %d, wanted %d" % (R.r, got, op[2], op[1], op[3]))
            R.slot[op[1]] = got
        elif kind == "signal_pulled":
            P = self.ranks[op[2]]
            P.pulled[op[1]][R.r] = max(P.pulled[op[1]][R.r], op[3])
        elif kind == "prep":
            got = R.slot[op[1]]
            R.prepared.append((op[2], got))
            if got != op[2]:
                self.errors.append("rank %d prepared frame %d from a slot holding %r" % (R.r, op[2], got))

    def run(self, host_lead=None):
        """Interleaves host issue and device execution at random; host_lead bounds how many frames a rank's host
        may be ahead of its own preparation kernel (None: unbounded, the host never waits)."""
        n = len(self.sources)
        issued = [0] * self.world
        while True:
            moves = []
            for r, R in enumerate(self.ranks):
                if issued[r] < n and (host_lead is None or issued[r] - len(R.prepared) < host_lead):
                    moves.append(("issue", r))
                for name, q in R.q.items():
                    if q and self.runnable(R, q[0]):
                        moves.append(("exec", r, name))
            if not moves:
                break
            m = self.rng.choice(moves)
            if m[0] == "issue":
                self.issue_frame(m[1], issued[m[1]])
                issued[m[1]] += 1
            else:
                R = self.ranks[m[1]]
                self.execute(R, R.q[m[2]].pop(0))
        stuck = [(R.r, name, q[0]) for R in self.ranks for name, q in R.q.items() if q]
        if stuck or any(i < n for i in issued):
            self.errors.append("deadlock: %r" % (stuck[:4],))
        return self.errors
