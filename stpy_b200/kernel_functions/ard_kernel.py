"""ard_kernel / ard_kernel_diag as free functions (stpy/kernel_functions/ard_kernel.py:5-44).
kwargs must carry ard_gamma (indexed by input column), kappa and group."""
import torch

from .. import _lib as L
from ..kernels import _Item, _f, _prep, _vec
from .kernel_params import KernelParams


def ard_kernel(a, b, **kwargs):
    """kappa exp(-0.5 sum_c ((b_jc - a_ic) / ard_gamma_c)^2) over the columns `group`; shape (|b|, |a|)."""
    p = KernelParams(kwargs)
    p.assert_existence(["ard_gamma", "kappa", "group"])
    ard = _vec(p.ard_gamma)
    item = _Item(L.K_SE, list(p.group), scale=[1.0 / ard[g] for g in p.group], arg_scale=-0.5, kappa=_f(p.kappa))
    a_dev, b_dev = L.to_device(a), L.to_device(b)
    ap, na, dpad = _prep(a_dev, item)
    bp, nb, _ = _prep(b_dev, item)
    out, ld = L.empty_matrix(b_dev.shape[0], a_dev.shape[0])
    L.call("stpyb_gram", item.kind, L.ptr(ap), L.ptr(na), a_dev.shape[0], L.ptr(bp), L.ptr(nb), b_dev.shape[0], dpad,
           item.arg_scale, item.kappa, 0.0, 0, L.OP_SET, 0.0, 0, L.ptr(out), ld, None, L.stream_ptr())
    return out if (torch.is_tensor(a) and a.is_cuda) else out.cpu()


def ard_kernel_diag(a, b, **kwargs):
    """In the reference this twin evaluates the FULL (|b|, |a|) matrix, exactly like ard_kernel
    (ard_kernel.py:27-44); kept that way."""
    return ard_kernel(a, b, **kwargs)
