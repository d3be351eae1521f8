"""Race detector for the multi-stream schedule of the distributed factorisation (no GPU, no process group).

DistributedGP._factor enqueues its work on three streams (bulk = current, high-priority panel chain,
communication) and orders them with events.  The gloo tests execute the schedule in program order, which
cannot expose a missing event.  Here the tile operations are replaced by a tracer that records, per
operation, the stream it was enqueued on and the resources it reads / writes (block columns of the slab, ring
slots of the panel buffers); record()/wait() build the happens-before relation (vector clocks), and every
pair of conflicting accesses issued on different streams must be ordered by it."""
import sys

import pytest
import torch

from conftest import ROOT

sys.path.insert(0, ROOT)


class _Stream:
    def __init__(self, tr, name):
        self.tr, self.name = tr, name

    def wait_stream(self, other):
        self.tr.wait(self, self.tr.record_on(other))


class _Ctx:
    def __init__(self, tr, s):
        self.tr, self.s = tr, s

    def __enter__(self):
        self.prev = self.tr.cur
        self.tr.cur = self.s

    def __exit__(self, *a):
        self.tr.cur = self.prev
        return False


class _Work:
    def wait(self):
        pass


class TraceOps:
    """Stands in for DeviceOps: computes nothing, logs (stream, reads, writes) of every operation."""
    device_type = "cuda"

    def __init__(self):
        self.streams = {}
        self.main = self._stream("main")
        self.cur = self.main
        self.ops = []          # (stream name, index on stream, clock, reads, writes, label)
        self.count = {}        # ops issued per stream
        self.clock = {}        # stream -> vector clock of its last op
        self.pending = {}      # stream -> clocks to merge into its next op (from waits)
        self.gp = None

    def _stream(self, name):
        if name not in self.streams:
            self.streams[name] = _Stream(self, name)
        return self.streams[name]

    # ---- stream plumbing
    def device(self):
        return torch.device("cpu")

    def zeros(self, *shape, dtype=torch.float64):
        return torch.zeros(*shape, dtype=dtype)

    def empty(self, *shape, dtype=torch.float64):
        return torch.zeros(*shape, dtype=dtype)

    def current_stream(self):
        return self.main

    def side_stream(self, high_priority=False):
        return self._stream("chain" if high_priority else "comm")

    def stream_ctx(self, s):
        return _Ctx(self, s)

    def record_on(self, s):
        return dict(self.clock.get(s.name, {}))

    def record(self):
        return self.record_on(self.cur)

    def wait(self, stream, event):
        if event is not None:
            self.pending.setdefault(stream.name, []).append(event)

    def log(self, label, reads=(), writes=()):
        s = self.cur.name
        clk = dict(self.clock.get(s, {}))
        for ev in self.pending.pop(s, []):
            for k, v in ev.items():
                clk[k] = max(clk.get(k, -1), v)
        idx = self.count.get(s, 0)
        self.count[s] = idx + 1
        clk[s] = idx
        self.clock[s] = clk
        self.ops.append((s, idx, clk, tuple(reads), tuple(writes), label))

    # ---- resources from tensor arguments
    def _col(self, t):
        gp = self.gp
        off = t.storage_offset() % gp._ld
        return ("col", off // gp.nbw)

    def _buf(self, t):
        for i, b in enumerate(self.gp._pbuf):
            if b.untyped_storage().data_ptr() == t.untyped_storage().data_ptr():
                return ("buf", i)
        raise AssertionError("not a panel buffer")

    # ---- traced tile operations
    def factor_panel(self, P, rows, w, ld, dinv, info, j0):
        self.log("factor+pack", reads=[], writes=[self._col(P), self._buf(dinv)])

    def update(self, C, ldc, A, B, ldp, M, N, K):
        self.log("update", reads=[self._buf(A)], writes=[self._col(C)])

    def update_batch(self, tasks):
        for t in tasks:
            self.update(*t)

    def bcast(self, t, src, owner):
        self.log("bcast", reads=[self._buf(t)] if owner else [], writes=[] if owner else [self._buf(t)])
        return _Work()


def _check(tr):
    last = {}
    races = []
    for s, idx, clk, reads, writes, label in tr.ops:
        for res in set(reads) | set(writes):
            for (s0, i0, l0, w0) in last.get(res, []):
                conflict = w0 or (res in writes)
                if conflict and s0 != s and clk.get(s0, -1) < i0:
                    races.append((res, l0, s0, i0, label, s, idx))
        for res in set(reads) | set(writes):
            last.setdefault(res, []).append((s, idx, label, res in writes))
    return races


@pytest.mark.parametrize("world,NB,depth", [(1, 9, None), (2, 14, None), (4, 23, None), (8, 40, None), (8, 40, 1),
                                            (3, 17, 5), (2, 9, 0), (4, 21, 16), (8, 19, 30)])
def test_factor_schedule_has_no_cross_stream_race(world, NB, depth):
    from stpy_b200.distributed import DistributedGP
    nbw = 128
    n = NB * nbw - 37
    for rank in range(world):
        tr = TraceOps()

        class K:
            params_dict = {}
        gp = DistributedGP(K(), s=0.1, nbw=nbw, ops=tr, lookahead=(depth != 0), depth=depth)
        gp.world, gp.rank = world, rank
        gp.depth = (min(world, 4) if depth is None else depth)
        tr.gp = gp
        lay = gp._alloc(n, 0)
        gp._bcast = lambda t, src, tr=tr, rank=rank: tr.bcast(t, src, src == rank)
        # the Gram writes every local column on the bulk stream before the factorisation ...
        tr.log("gram", writes=[("col", lay.slot(g)) for g in lay.local_blocks])
        gp._factor(lay, n + 1)
        # ... and the solves read every column and every panel buffer afterwards
        tr.log("solves", reads=[("col", lay.slot(g)) for g in lay.local_blocks] + [("buf", i) for i in range(len(gp._pbuf))])
        races = _check(tr)
        assert not races, races[:5]
        # every local column received every earlier panel exactly once, in order
        seen = {}
        for s, idx, clk, reads, writes, label in tr.ops:
            if label == "update":
                seen.setdefault(writes[0], []).append(s)
        for g in lay.local_blocks:
            assert len(seen.get(("col", lay.slot(g)), [])) == g, (g, seen.get(("col", lay.slot(g))))


def test_the_detector_sees_a_missing_event():
    """Sanity of the detector itself: drop the `arrived` wait of the bulk stream and a race must be reported."""
    from stpy_b200.distributed import DistributedGP
    tr = TraceOps()

    class K:
        params_dict = {}
    gp = DistributedGP(K(), s=0.1, nbw=128, ops=tr, depth=2)
    gp.world, gp.rank, gp.depth = 2, 1, 2
    tr.gp = gp
    lay = gp._alloc(128 * 12, 0)
    gp._bcast = lambda t, src: tr.bcast(t, src, src == 1)
    real_wait = tr.wait
    tr.wait = lambda stream, ev: None if stream is tr.main else real_wait(stream, ev)
    tr.log("gram", writes=[("col", lay.slot(g)) for g in lay.local_blocks])
    gp._factor(lay, 128 * 12 + 1)
    assert _check(tr)
