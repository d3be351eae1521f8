"""The multi-GPU schedule (stpy_b200/distributed.py) driven on CPU: world_size 2 and 3 over gloo,
with torch-CPU tile operations standing in for the CUDA library.  Checks the block-cyclic
ownership map, the look-ahead ordering, the panel packing / broadcast protocol, the augmented
y-row forward solve, the two-scalar evidence reduction and the pipelined backward solve against
the serial oracle."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

DB = 128


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class CpuOps:
    """torch-CPU restatement of the tile operations (test infrastructure only)."""
    device_type = "cpu"

    def device(self):
        return torch.device("cpu")

    def zeros(self, *shape, dtype=torch.float64):
        return torch.zeros(*shape, dtype=dtype)

    def empty(self, *shape, dtype=torch.float64):
        return torch.full(shape, float("nan"), dtype=dtype) if dtype == torch.float64 else torch.zeros(*shape, dtype=dtype)

    def gram_block(self, kernel_object, params_dict, x_cols, x_rows, out, ld, diag_add):
        K = kernel_object.fn(x_cols, x_rows)
        w = x_cols.shape[0]
        K[:w, :w] += diag_add * torch.eye(w, dtype=torch.float64)
        out.copy_(torch.tril(K) + torch.triu(torch.full_like(K, float("nan")), 1))  # only the lower part is defined

    def gram_rect(self, kernel_object, params_dict, x_cols, x_rows, out, ld):
        out.copy_(kernel_object.fn(x_cols, x_rows))

    def gram_diag(self, kernel_object, params_dict, xt):
        return torch.stack([kernel_object.fn(xt[i:i + 1], xt[i:i + 1]).reshape(()) for i in range(xt.shape[0])])

    def factor_panel(self, P, rows, w, ld, dinv, info, j0):
        top = torch.tril(P[:w, :w])
        top = top + torch.tril(top, -1).T
        Lf = torch.linalg.cholesky(top)
        P[:w, :w] = torch.tril(Lf) + torch.triu(P[:w, :w], 1)
        if rows > w:
            P[w:rows, :w] = torch.linalg.solve_triangular(Lf, P[w:rows, :w].T, upper=False).T
        d = dinv.view(-1, DB, DB)
        for k in range((w + DB - 1) // DB):
            b = min(DB, w - k * DB)
            d[k].zero_()
            d[k][:b, :b] = torch.linalg.inv(Lf[k * DB:k * DB + b, k * DB:k * DB + b])

    def update(self, C, ldc, A, B, ldp, M, N, K):
        upd = A[:M, :K] @ B[:N, :K].T
        mask = torch.tril(torch.ones(M, N, dtype=torch.bool))
        C[:M, :N] = torch.where(mask, C[:M, :N] - upd, C[:M, :N])

    def update_batch(self, tasks):
        for t in tasks:
            self.update(*t)

    def trsv_t(self, Lblk, w, ld, dinv, x):
        Lf = torch.tril(Lblk[:w, :w])
        x[:w] = torch.linalg.solve_triangular(Lf.T, x[:w].view(-1, 1), upper=True).view(-1)

    def gemv_t_sub(self, A, rows, w, ld, v, y):
        y[:w] -= A[:rows, :w].T @ v[:rows]

    def evidence_terms(self, Lblk, w, ld, zrow, out3):
        out3[0] = (zrow[:w] * zrow[:w]).sum()
        out3[1] = 2.0 * torch.log(torch.diagonal(Lblk[:w, :w])).sum()
        out3[2] = 0.0

    def pred_partials(self, V, nx, w, ld, zrow, out2):
        out2[0] = V[:nx, :w] @ zrow[:w]
        out2[1] = (V[:nx, :w] * V[:nx, :w]).sum(dim=1)

    def alpha_step(self, Lcol, ld, below, w, dinv, zrow, alpha_below, seg):
        seg[:w] = zrow[:w]
        if below > 0:
            self.gemv_t_sub(Lcol[w:], below, w, ld, alpha_below, seg)
        self.trsv_t(Lcol, w, ld, dinv, seg)

    def side_stream(self, high_priority=False):
        return None

    def stream_ctx(self, s):
        return _Null()

    def record(self):
        return None

    def wait(self, stream, event):
        pass

    def current_stream(self):
        return None


class FakeKernel:
    params_dict = {}

    def __init__(self, fn):
        self.fn = fn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, nbw, lookahead, q, depth=None):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import stpy_oracle as O
        from stpy_b200.distributed import DistributedGP
        x, y = O.make_data(n, 3, seed=1)
        kern = lambda a, b: O.matern_kernel(a, b, gamma=1.0, nu=2.5)
        gp = DistributedGP(FakeKernel(kern), s=0.1, nbw=nbw, ops=CpuOps(), lookahead=lookahead, depth=depth)
        gp.fit_gp(x, y)
        gp.check()
        lml = float(gp.log_marginal(0.7))
        xt, _ = O.make_data(37, 3, seed=2)
        ref = O.gp_cholesky(kern, x, y, 0.1, xt)
        ref_lml = float(O.lml_cholesky(kern, x, y, 0.1, 0.7))
        err_a = float((gp.A - ref["A"]).abs().max() / ref["A"].abs().max())
        mu, sd = gp.mean_std(xt)  # prediction rows appended to the augmented matrix
        err_p = max(float((mu - ref["mean"]).abs().max() / ref["mean"].abs().max()),
                    float((sd ** 2 - ref["std"] ** 2).abs().max() / (ref["std"] ** 2).abs().max()))
        assert abs(float(gp.log_marginal(0.7)) - ref_lml) < 1e-8  # unchanged by the extra rows
        q.put((rank, abs(lml - ref_lml), max(err_a, err_p), gp.lay.local_blocks))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,nbw,lookahead,depth", [(2, 700, 128, True, None), (3, 1100, 256, True, None),
                                                         (2, 513, 128, False, None), (2, 1500, 128, True, 1),
                                                         (3, 1700, 128, True, 5), (2, 900, 128, True, 40)])
def test_block_cyclic_cholesky_over_gloo(world, n, nbw, lookahead, depth):
    """depth = how far the panel chain may lead the bulk updates: default (= world: one chain column per rank
    and step), 1 (classic look-ahead), not a multiple of the world size (5 on 3 ranks: 1-2 chain columns per
    step) and larger than the number of block columns (everything on the chain stream)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, nbw, lookahead, q, depth)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(world))
    owned = []
    for rank, dl, ea, blocks in res:
        assert dl < 1e-8, (rank, dl)
        assert ea < 1e-9, (rank, ea)
        assert all(b % world == rank for b in blocks)
        owned += blocks
    assert sorted(owned) == list(range((n + nbw - 1) // nbw))


def test_layout_map():
    sys.path.insert(0, ROOT)
    from stpy_b200.distributed import BlockCyclicLayout
    lay = BlockCyclicLayout(n=1000, nbw=256, world=3, rank=1)
    assert lay.NB == 4 and lay.local_blocks == [1] and lay.nloc == 1
    assert lay.width(3) == 1000 - 768 and lay.col0(1) == 0 and lay.owner(3) == 0 and lay.slot(3) == 1
    lay0 = BlockCyclicLayout(n=1000, nbw=256, world=3, rank=0)
    assert lay0.local_blocks == [0, 3] and lay0.col0(3) == 256


def test_single_rank_schedule_matches_oracle():
    """world = 1 without a process group: the same schedule degenerates to a serial blocked Cholesky."""
    sys.path.insert(0, ROOT)
    from oracle import stpy_oracle as O
    from stpy_b200.distributed import DistributedGP
    x, y = O.make_data(450, 2, seed=3)
    kern = lambda a, b: O.se_kernel(a, b, gamma=0.5)
    gp = DistributedGP(FakeKernel(kern), s=0.1, nbw=128, ops=CpuOps())
    gp.fit_gp(x, y)
    assert abs(float(gp.log_marginal(1.0)) - float(O.lml_cholesky(kern, x, y, 0.1))) < 1e-8
    ref = O.gp_cholesky(kern, x, y, 0.1)
    assert float((gp.A - ref["A"]).abs().max() / ref["A"].abs().max()) < 1e-9


def test_panel_width_heuristic():
    from stpy_b200.distributed import DistributedGP
    assert DistributedGP.pick_nbw(65536, 1) == 1024 and DistributedGP.pick_nbw(65536, 2) == 1024
    assert DistributedGP.pick_nbw(65536, 4) == 512 and DistributedGP.pick_nbw(65536, 8) == 512
    assert DistributedGP.pick_nbw(5000, 8) == 128 and DistributedGP.pick_nbw(100, 2) == 128
    for n in (300, 5000, 20000, 65536):
        for w in (1, 2, 3, 8):
            b = DistributedGP.pick_nbw(n, w)
            assert b % 128 == 0 and 128 <= b <= 1024


def _sweep_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import stpy_oracle as O
        from stpy_b200 import sweep
        x, y = O.make_data(120, 2, seed=4)
        gammas = [0.3, 0.5, 0.8, 1.1, 1.6]
        seen = []

        def local_sweep(kernels, xx, yy, s, weight=1.0, **kw):  # stands in for the CUDA sweep of one rank
            seen.extend(kernels)
            return torch.stack([O.lml_cholesky(lambda a, b, g=g: O.se_kernel(a, b, gamma=g), xx, yy, s, weight).reshape(())
                                for g in kernels])
        sweep.lml_sweep = local_sweep
        sweep.L.device = lambda: torch.device("cpu")
        vals = sweep.lml_sweep_distributed(gammas, x, y, 0.1, weight=0.9)
        q.put((rank, vals.tolist(), seen))  # plain lists: a tensor would travel through shared memory
    finally:
        dist.destroy_process_group()


def test_distributed_sweep_deals_kernels_round_robin_and_gathers():
    """lml_sweep_distributed: replicas with no data-path collective; rank r scores kernels r, r+W, ... and one
    all-reduce over disjoint supports gathers the values.  The per-rank CUDA sweep is replaced by the oracle."""
    from oracle import stpy_oracle as O
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sweep_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, y = O.make_data(120, 2, seed=4)
    gammas = [0.3, 0.5, 0.8, 1.1, 1.6]
    ref = torch.stack([O.lml_cholesky(lambda a, b, g=g: O.se_kernel(a, b, gamma=g), x, y, 0.1, 0.9).reshape(())
                       for g in gammas])
    assert got[0][2] == [0.3, 0.8, 1.6] and got[1][2] == [0.5, 1.1]
    for _, vals, _ in got:
        assert len(vals) == 5 and float((torch.tensor(vals, dtype=torch.float64) - ref).abs().max()) < 1e-10


def _restart_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import numpy as np
        from stpy_b200.estimator import Estimator

        class Toy(Estimator):  # a multi-modal "evidence" on the host: which restarts ran where is what is tested
            def __init__(self):
                self.kernel_object = type("K", (), {"params_dict": {'0': {'gamma': 1.0}}})()
                self.s, self.x, self.y, self.evals = 0.1, None, None, 0

            def ucb(self, x):
                pass

            def lcb(self, x):
                pass

            def fit_gp(self, x, y):
                self.refit = True

            def _lml_value(self, kernel, X, weight):
                self.evals += 1
                g = X['0']['gamma']
                return (torch.sin(3.0 * g) + 0.1 * (g - 2.0) ** 2).reshape(1, 1)

        starts = iter([0.3, 1.4, 2.4, 3.6, 4.5, 0.9])
        toy = Toy()
        toy.optimize_params_general(params={'0': {'gamma': (lambda d: np.array([next(starts)]), 1, (0.05, 6.0))}},
                                    restarts=6, parallel=True)
        q.put((rank, toy.kernel_object.params_dict['0']['gamma'], len(toy.optimization_result['evidence']), toy.evals))
    finally:
        dist.destroy_process_group()


def test_restarts_are_dealt_over_ranks_and_gathered():
    """optimize_params_general(parallel=True): replicas, rank r runs restarts r, r+W, ...; one all-gather of the
    (value, point) pairs; every rank ends with the same best point and all six results."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_restart_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == got[1][1] and got[0][2] == got[1][2] == 6
    assert abs(got[0][1] - 1.598) < 0.02  # the global minimum of sin(3g) + 0.1 (g - 2)^2 on [0.05, 6]
    serial = sum(t[3] for t in got)
    assert min(t[3] for t in got) > 0 and max(t[3] for t in got) < serial  # both ranks did part of the work
