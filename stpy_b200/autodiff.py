"""Analytic derivatives of the GP evidence and of the Gram operator, wired into torch.autograd.

The reference differentiates 0.5 y^T K^-1 y + 0.5 w logdet K by letting autograd
back-propagate through torch.linalg.solve / slogdet and through every kernel builder's
exp / mm / cdist (gauss_procc.py:631-638, kernels.py:146-157, driven by estimator.py:156-171).
Here the forward pass is the device factorisation and the backward pass is the closed form
    dLML/dtheta = 0.5 tr( (w K^-1 - alpha alpha^T) dK/dtheta )
evaluated by stpyb_potri + stpyb_kernel_grad, a derivative pass over the COMPOSITE kernel:
squared_exponential, ard (plain and additive groups), matern / ard_matern (nu in 1/2, 3/2, 5/2),
the per-group kernels, polynomial and linear, combined by + and * in any left fold.  Gradients are
returned for every tensor hyper-parameter that requires grad (gamma, ard_gamma, gamma_per_group,
ard_per_group, kappa) and for the noise level s.

`gram_with_grad` is the same pass with an explicit cotangent: it makes KernelFunction.kernel(a, b, **kw)
autograd-transparent in the tensors of kw (the reference's operator-seam contract, kernels.py:136-159).
"""
import ctypes

import torch

from . import _lib as L

GC_MAX = 32   # columns per item in one descriptor (stpyb_kernel_grad)
GI_MAX = 8    # items per composite kernel


def _tensors(params_dict):
    for sub in params_dict.values():
        for v in sub.values():
            if torch.is_tensor(v):
                yield v


def needs_grad(params_dict, s=None):
    if torch.is_tensor(s) and s.requires_grad:
        return True
    return any(t.requires_grad for t in _tensors(params_dict))


class _ScaleGrads(torch.autograd.Function):
    """value with precomputed d value / d input_k: backward scales them by the incoming scalar."""

    @staticmethod
    def forward(ctx, value, *grads_and_inputs):
        k = len(grads_and_inputs) // 2
        ctx.save_for_backward(*grads_and_inputs[:k])
        return value.clone()

    @staticmethod
    def backward(ctx, gout):
        grads = ctx.saved_tensors
        g = gout.reshape(())
        return (None,) + tuple(None for _ in grads) + tuple(g * t for t in grads)


def _descriptor(items, sub_ops):
    if len(items) > GI_MAX:
        raise NotImplementedError("the derivative pass handles at most %d items (sub-kernels x groups)" % GI_MAX)
    for it in items:
        if len(it["cols"]) > GC_MAX:
            raise NotImplementedError("the derivative pass handles at most %d input columns per item" % GC_MAX)
    cols = [0] * (len(items) * GC_MAX)
    sc = [0.0] * (len(items) * GC_MAX)
    for q, it in enumerate(items):
        for c, (col, v) in enumerate(zip(it["cols"], it["sc"])):
            cols[q * GC_MAX + c] = col
            sc[q * GC_MAX + c] = v
    return (len(items), L.host_ints([it["kind"] for it in items]), L.host_ints([len(it["cols"]) for it in items]),
            L.host_ints([it["sub"] for it in items]), L.host_ints(cols), L.host_doubles(sc),
            L.host_doubles([it["arg_scale"] for it in items]), L.host_doubles([it["kappa"] for it in items]),
            L.host_doubles([it["p0"] for it in items]), len(sub_ops), L.host_ints(sub_ops))


def _wanted(t):
    return torch.is_tensor(t) and t.requires_grad


def _run_passes(items, sub_ops, xr, xc, mode, cmat, ldc, alpha, weight, need_trace):
    """Launch the passes that some requested gradient needs; returns (per-pass records, host results)."""
    desc = _descriptor(items, sub_ops)
    m, n, d = xr.shape[0], xc.shape[0], xr.shape[1]
    passes = []
    for q, it in enumerate(items):
        want_ls = it["ls_idx"] is not None and _wanted(it["ls_src"])
        want_k = _wanted(it["kappa_src"])
        if want_ls:
            passes += [(q, off) for off in range(0, len(it["cols"]), 16)]
        elif want_k:
            passes.append((q, 0))
    if not passes and need_trace:
        passes.append((0, 0))
    out = torch.empty((max(1, len(passes)), 18), dtype=torch.float64, device=xr.device)
    for r, (q, off) in enumerate(passes):
        L.call("stpyb_kernel_grad", L.ptr(xr), m, xr.stride(0), L.ptr(xc), n, xc.stride(0), d, *desc, q, off, mode,
               L.ptr(cmat), ldc, L.ptr(alpha), float(weight), L.ptr(out[r]), L.stream_ptr())
    return passes, out


def _assemble(items, passes, host, extra=None):
    """Scatter the pass sums into one gradient per tensor that requires grad -> (inputs, grads)."""
    acc = {}

    def add(t, idx, val):
        if not _wanted(t):
            return
        key = id(t)
        if key not in acc:
            acc[key] = (t, torch.zeros(t.numel(), dtype=torch.float64))
        acc[key][1][idx] += val

    seen_kappa = set()
    for r, (q, off) in enumerate(passes):
        it = items[q]
        if it["ls_idx"] is not None:
            for u in range(min(16, len(it["cols"]) - off)):
                c = off + u
                add(it["ls_src"], it["ls_idx"][c], float(host[r, u]) * (-2.0 / it["ls"][c]))
        if off == 0 and q not in seen_kappa:
            seen_kappa.add(q)
            add(it["kappa_src"], 0, float(host[r, 16]) * it["dkappa"])
    for t, idx, val in (extra or []):
        add(t, idx, val)
    inputs = [t for t, _ in acc.values()]
    grads = [g.reshape(t.shape).to(dtype=t.dtype, device=t.device) for t, g in acc.values()]
    return inputs, grads


def lml_with_grad(gp, kernel_object, params_dict, weight):
    items, sub_ops = kernel_object.grad_plan(params_dict)
    n = gp.n
    s_val = float(gp.s.detach()) if torch.is_tensor(gp.s) else float(gp.s)
    f = gp._factor_for(kernel_object, params_dict, s_val)
    L.call("stpyb_lml", L.ptr(f.buf), n, f.ld, L.ptr(f.z), float(weight), L.ptr(f.out3), L.stream_ptr())
    alpha = f.z.clone()
    L.call("stpyb_trsv", L.ptr(f.buf), n, f.ld, L.ptr(f.dinv), L.ptr(alpha), 1, L.stream_ptr())
    work, ldw = L.empty_matrix(n, n)
    kinv, ldk = L.empty_matrix(n, n)
    L.call("stpyb_potri", L.ptr(f.buf), n, f.ld, L.ptr(f.dinv), L.ptr(work), ldw, L.ptr(kinv), ldk, L.stream_ptr())
    del work
    x = gp._x_dev
    passes, out = _run_passes(items, sub_ops, x, x, 0, kinv, ldk, alpha, weight, need_trace=_wanted(gp.s))
    host = torch.cat([f.out3, out.reshape(-1)]).cpu()  # one read-back
    f.check()
    value = host[2].view(1, 1)
    res = host[3:].view(-1, 18)
    extra = [(gp.s, 0, 2.0 * s_val * float(res[0, 17]))] if _wanted(gp.s) else []
    inputs, grads = _assemble(items, passes, res, extra)
    if not inputs:
        return value
    return _ScaleGrads.apply(value, *grads, *inputs)


class _GramFn(torch.autograd.Function):
    """K = kernel(a, b; theta) with d<G, K>/dtheta evaluated by the derivative pass (mode 1)."""

    @staticmethod
    def forward(ctx, kernel_object, params_dict, a_dev, b_dev, symmetric, to_cpu, *inputs):
        out, ld = L.empty_matrix(b_dev.shape[0], a_dev.shape[0])
        kernel_object.gram_into(a_dev, b_dev, params_dict, out, ld, symmetric=symmetric)
        ctx.kernel_object, ctx.params_dict, ctx.a_dev, ctx.b_dev = kernel_object, params_dict, a_dev, b_dev
        ctx.inputs = inputs
        return out.cpu() if to_cpu else out

    @staticmethod
    def backward(ctx, gout):
        items, sub_ops = ctx.kernel_object.grad_plan(ctx.params_dict)
        g, ldg = L.empty_matrix(gout.shape[0], gout.shape[1])
        g.copy_(gout.to(device=g.device, dtype=torch.float64))
        passes, out = _run_passes(items, sub_ops, ctx.b_dev, ctx.a_dev, 1, g, ldg, None, 1.0, need_trace=False)
        host = out.cpu()
        inputs, grads = _assemble(items, passes, host)
        by_id = {id(t): gr for t, gr in zip(inputs, grads)}
        return (None,) * 6 + tuple(by_id.get(id(t)) for t in ctx.inputs)


def gram_with_grad(kernel_object, params_dict, a_dev, b_dev, symmetric, to_cpu):
    inputs = []
    for t in _tensors(params_dict):
        if t.requires_grad and all(t is not u for u in inputs):
            inputs.append(t)
    return _GramFn.apply(kernel_object, params_dict, a_dev, b_dev, symmetric, to_cpu, *inputs)
