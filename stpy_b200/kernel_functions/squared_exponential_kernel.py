"""squared_exponential_kernel / squared_exponential_kernel_diag as free functions
(stpy/kernel_functions/squared_exponential_kernel.py:5-38; the diag twin is the one
stpy/kernels.py:8,174 imports).  kwargs must carry gamma, kappa and group."""
import torch

from .. import _lib as L
from ..kernels import _Item, _f, _prep
from .kernel_params import KernelParams


def squared_exponential_kernel(a, b, **kwargs):
    """kappa exp(-|b_j - a_i|^2 / (2 gamma^2)) on the columns `group`; shape (|b|, |a|)."""
    p = KernelParams(kwargs)
    p.assert_existence(["gamma", "kappa", "group"])
    gamma = _f(p.gamma)
    item = _Item(L.K_SE, list(p.group), arg_scale=-0.5 / (gamma * gamma), kappa=_f(p.kappa))
    a_dev, b_dev = L.to_device(a), L.to_device(b)
    ap, na, dpad = _prep(a_dev, item)
    bp, nb, _ = _prep(b_dev, item)
    out, ld = L.empty_matrix(b_dev.shape[0], a_dev.shape[0])
    L.call("stpyb_gram", item.kind, L.ptr(ap), L.ptr(na), a_dev.shape[0], L.ptr(bp), L.ptr(nb), b_dev.shape[0], dpad,
           item.arg_scale, item.kappa, 0.0, 0, L.OP_SET, 0.0, 0, L.ptr(out), ld, None, L.stream_ptr())
    return out if (torch.is_tensor(a) and a.is_cuda) else out.cpu()


def squared_exponential_kernel_diag(a, b, **kwargs):
    """The reference applies the map to (a - b)^2 ELEMENTWISE, without summing over the group
    (squared_exponential_kernel.py:29-38): shape (n, |group|), the kernel diagonal when |group| = 1.
    Each column is one 1-d kernel diagonal on the device (stpyb_gram_diag)."""
    p = KernelParams(kwargs)
    p.assert_existence(["gamma", "kappa", "group"])
    gamma = _f(p.gamma)
    a_dev, b_dev = L.to_device(a), L.to_device(b)
    n = a_dev.shape[0]
    cols = []
    for c in p.group:
        item = _Item(L.K_SE, [c], arg_scale=-0.5 / (gamma * gamma), kappa=_f(p.kappa))
        ap, na, dpad = _prep(a_dev, item)
        bp, nb, _ = _prep(b_dev, item)
        out = torch.empty((n,), dtype=torch.float64, device=a_dev.device)
        L.call("stpyb_gram_diag", item.kind, L.ptr(ap), L.ptr(na), L.ptr(bp), L.ptr(nb), n, dpad, item.arg_scale,
               item.kappa, 0.0, L.OP_SET, L.ptr(out), None, L.stream_ptr())
        cols.append(out)
    res = torch.stack(cols, dim=1)
    return res if (torch.is_tensor(a) and a.is_cuda) else res.cpu()
