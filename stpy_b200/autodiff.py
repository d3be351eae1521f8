"""Analytic value + gradient of the GP evidence, wired into torch.autograd.

The reference differentiates 0.5 y^T K^-1 y + 0.5 w logdet K by letting autograd
back-propagate through torch.linalg.solve / slogdet / exp / mm
(gauss_procc.py:631-638, driven by estimator.py:156-171).  Here the forward pass
is the device factorisation and the backward pass is the closed form
    dLML/dtheta = 0.5 tr( (w K^-1 - alpha alpha^T) dK/dtheta )
evaluated by stpyb_potri + stpyb_lml_grad_se.  Supported: a single
squared_exponential or ard sub-kernel (gradients w.r.t. gamma / ard_gamma, kappa)
and the noise level s.
"""
import torch

from . import _lib as L


def _tensors(params_dict):
    for sub in params_dict.values():
        for v in sub.values():
            if torch.is_tensor(v):
                yield v


def needs_grad(params_dict, s):
    if torch.is_tensor(s) and s.requires_grad:
        return True
    return any(t.requires_grad for t in _tensors(params_dict))


class _LmlFn(torch.autograd.Function):

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


def lml_with_grad(gp, kernel_object, params_dict, weight):
    if len(kernel_object._owners) != 1:
        raise NotImplementedError("analytic LML gradients cover a single squared_exponential / ard kernel")
    owner = kernel_object._owners[0]
    kw = params_dict['0'] if '0' in params_dict else {}
    if owner.optkernel not in ("squared_exponential", "ard") or kw.get('groups', getattr(owner, 'groups', None)):
        raise NotImplementedError("analytic LML gradients cover squared_exponential and (non-additive) ard kernels")
    item = owner._items(kw)[0]
    n = gp.n
    s_val = float(gp.s.detach()) if torch.is_tensor(gp.s) else float(gp.s)
    f = gp._factor_for(kernel_object, params_dict, s_val)
    dev = f.buf.device
    L.call("stpyb_lml", L.ptr(f.buf), n, f.ld, L.ptr(f.z), float(weight), L.ptr(f.out3), L.stream_ptr())
    alpha = f.z.clone()
    L.call("stpyb_trsv", L.ptr(f.buf), n, f.ld, L.ptr(f.dinv), L.ptr(alpha), 1, L.stream_ptr())
    work, ldw = L.empty_matrix(n, n)
    kinv, ldk = L.empty_matrix(n, n)
    L.call("stpyb_potri", L.ptr(f.buf), n, f.ld, L.ptr(f.dinv), L.ptr(work), ldw, L.ptr(kinv), ldk, L.stream_ptr())
    del work
    from .kernels import _prep
    xp, nrm, dpad = _prep(gp._x_dev, item)
    dg = len(item.cols)
    out = torch.empty((dg + 2,), dtype=torch.float64, device=dev)
    L.call("stpyb_lml_grad_se", L.ptr(kinv), ldk, L.ptr(alpha), L.ptr(xp), L.ptr(nrm), n, dpad, dg,
           item.arg_scale, item.kappa, float(weight), L.ptr(out), L.stream_ptr())
    host = torch.cat([f.out3, out]).cpu()  # one read-back
    f.check()
    value = host[2].view(1, 1)
    g = host[3:]
    inputs, grads = [], []

    def add(t, grad):
        if torch.is_tensor(t) and t.requires_grad:
            inputs.append(t)
            grads.append(grad.to(dtype=t.dtype).reshape(t.shape).to(t.device))

    if owner.optkernel == "ard":
        ard = kw.get('ard_gamma', owner.ard_gamma)
        if torch.is_tensor(ard) and ard.requires_grad:
            full = torch.zeros(ard.numel(), dtype=torch.float64)
            flat = ard.detach().reshape(-1).cpu().double()
            for pos, col in enumerate(kw.get('group', owner.group)):
                full[col] += g[pos] / flat[col]
            add(ard, full)
    else:
        gamma = kw.get('gamma', owner.gamma)
        if torch.is_tensor(gamma) and gamma.requires_grad:
            gv = float(gamma.detach().reshape(-1)[0])
            add(gamma, g[:dg].sum() / gv ** 3)
    add(kw.get('kappa', None), g[dg])
    add(gp.s, 2.0 * s_val * g[dg + 1])
    if not inputs:
        return value
    return _LmlFn.apply(value, *grads, *inputs)
