"""Batched model scoring: the log marginal likelihood of many kernels on one dataset.

This is the shape of the reference's model-selection tutorial and of
CategoricalMixture.fit_gp (stpy/continuous_processes/categorical_mixture.py:48-65),
which build one Gram matrix and one LU log-evidence per kernel in a Python loop.
Here one DMMA pass over the shared squared-distance tiles feeds every kernel's
epilogue (stpyb_gram_multi), and the independent Cholesky factorisations run on
several CUDA streams so that one factorisation's latency-bound diagonal blocks
overlap the others' trailing updates.
"""
import ctypes

import torch

from . import _lib as L
from .kernels import _Item, _prep, _MATERN_KIND, _f


def _isotropic_spec(k):
    """(kind, arg_scale, kappa, group) of a single isotropic SE / Matern KernelFunction."""
    if len(k._owners) != 1:
        raise NotImplementedError("sweep entries must be single kernels")
    p = k.params_dict['0']
    kappa = _f(p.get('kappa', k.kappa))
    gamma = _f(p.get('gamma', k.gamma))
    group = tuple(p.get('group', k.group))
    if k.optkernel == "squared_exponential":
        return L.K_SE, -0.5 / (gamma * gamma), kappa, group
    if k.optkernel == "matern":
        nu = _f(p.get('nu', k.v))
        if nu not in _MATERN_KIND:
            raise NotImplementedError("Matern nu=%s" % nu)
        return _MATERN_KIND[nu], 1.0 / gamma, kappa, group
    raise NotImplementedError("sweep supports isotropic squared_exponential and matern kernels")


def lml_sweep_distributed(kernels, x, y, s, weight=1.0, **kw):
    """The sweep over the ranks of the default process group: kernels are dealt round-robin
    (replicas, no data-path collective), the per-rank values are all-gathered."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return lml_sweep(kernels, x, y, s, weight=weight, **kw)
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = list(range(rank, len(kernels), world))
    vals = torch.full((len(kernels),), float("nan"), dtype=torch.float64)
    if mine:
        vals[mine] = lml_sweep([kernels[i] for i in mine], x, y, s, weight=weight, **kw)
    dev = L.device()
    buf = torch.nan_to_num(vals, nan=0.0).to(dev)
    dist.all_reduce(buf)  # disjoint supports: the sum is the gather
    return buf.cpu()


def lml_sweep(kernels, x, y, s, weight=1.0, batch=16, streams=8, outer_block=512):
    """Evidence 0.5 y^T K^-1 y + 0.5 w logdet K (gauss_procc.py:631-638) for every kernel; returns a
    CPU float64 tensor of len(kernels) values.  Raises LinAlgError if any Gram is not PD."""
    x_dev = L.to_device(x)
    y_dev = L.to_device(y).reshape(-1)
    n = x_dev.shape[0]
    specs = [_isotropic_spec(k) for k in kernels]
    groups = {sp[3] for sp in specs}
    if len(groups) != 1:
        raise NotImplementedError("all kernels of a sweep must act on the same input columns")
    xp, nrm, dpad = _prep(x_dev, _Item(L.K_LINEAR, list(groups.pop())))
    nk = len(specs)
    batch = max(1, min(batch, nk, 64))
    ld = L.pad_ld(n)
    nblk = (n + L.DB - 1) // L.DB
    dev = x_dev.device
    bufs = torch.empty((batch, n, ld), dtype=torch.float64, device=dev)
    dinv = torch.empty((batch, nblk, L.DB, L.DB), dtype=torch.float64, device=dev)
    zs = torch.empty((batch, ld), dtype=torch.float64, device=dev)[:, :n]  # rows 16-byte aligned for any n
    info = torch.zeros((nk,), dtype=torch.int32, device=dev)
    out = torch.zeros((nk, 3), dtype=torch.float64, device=dev)
    main = torch.cuda.current_stream()
    side = [torch.cuda.Stream() for _ in range(max(1, streams))]
    # the factorisations already overlap one another stream against stream; the library's own look-ahead
    # (one shared side stream per device) would only serialise their panels: opt out per call
    _sweep_batches(specs, nk, batch, n, ld, xp, nrm, dpad, s, weight, bufs, dinv, zs, info, out, y_dev, main, side,
                   int(outer_block) | L.POTRF_NO_LOOKAHEAD)
    host = out.cpu()
    bad = info.cpu().nonzero()
    if bad.numel() > 0:
        raise torch.linalg.LinAlgError("kernel %d of the sweep: Gram matrix not positive-definite" % int(bad[0]))
    return host[:, 2].clone()


def _sweep_batches(specs, nk, batch, n, ld, xp, nrm, dpad, s, weight, bufs, dinv, zs, info, out, y_dev, main, side,
                   outer_block):
    for lo in range(0, nk, batch):
        hi = min(nk, lo + batch)
        b = hi - lo
        L.call("stpyb_gram_multi", b, L.host_ints([sp[0] for sp in specs[lo:hi]]),
               L.host_doubles([sp[1] for sp in specs[lo:hi]]), L.host_doubles([sp[2] for sp in specs[lo:hi]]),
               L.ptr(xp), L.ptr(nrm), n, dpad, float(s) ** 2, L.ptr(bufs), ld, n * ld, L.stream_ptr())
        zs[:b].copy_(y_dev.unsqueeze(0).expand(b, n))
        ready = torch.cuda.Event()
        ready.record(main)
        for q in range(b):
            st = side[q % len(side)]
            st.wait_event(ready)
            with torch.cuda.stream(st):
                sp = L.stream_ptr()
                L.call("stpyb_potrf", L.ptr(bufs[q]), n, ld, L.ptr(dinv[q]), L.ptr(info[lo + q:]), outer_block, sp)
                L.call("stpyb_trsv", L.ptr(bufs[q]), n, ld, L.ptr(dinv[q]), L.ptr(zs[q]), 0, sp)
                L.call("stpyb_lml", L.ptr(bufs[q]), n, ld, L.ptr(zs[q]), float(weight), L.ptr(out[lo + q]), sp)
        for st in side:
            main.wait_stream(st)
