"""Round-2 probes of the HBM-bound stages on one B200: Gram kernels and single-RHS triangular solves.
Prints one JSON object per measurement (CUDA events, best of `reps` after warm-up).

    python tools/probe_r02.py gram [n]      # SE / Matern-5/2 Gram, lower-only and rectangular
    python tools/probe_r02.py trsv [n]      # forward / backward solve against a factor of order n
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stpy_b200 import _lib as L
from stpy_b200.kernels import KernelFunction as KF

F64 = torch.float64


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def gram(n):
    d = 8
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(n, d, dtype=F64, generator=g) * 2 - 1).cuda()
    xt = (torch.rand(4096, d, dtype=F64, generator=g) * 2 - 1).cuda()
    out, ld = L.empty_matrix(n, n)
    rect, ldr = L.empty_matrix(4096, n)
    for name, k in (("se", KF(kernel_name="squared_exponential", gamma=1.0, d=d)),
                    ("matern52", KF(kernel_name="matern", gamma=1.0, nu=2.5, d=d)),
                    ("ard_matern52", KF(kernel_name="ard_matern", ard_gamma=torch.ones(d, dtype=F64), nu=2.5, d=d)),
                    ("poly2", KF(kernel_name="polynomial", power=2, d=d))):
        pd = k.params_dict
        ms = timeit(lambda: k.gram_into(x, x, pd, out, ld, symmetric=True, lower_only=True, diag_add=0.01))
        by = 4.0 * n * n
        print(json.dumps({"probe": "gram_lower", "kernel": name, "n": n, "ms": ms, "GBps": by / ms / 1e6}))
        ms = timeit(lambda: k.gram_into(x, xt, pd, rect, ldr))
        by = 8.0 * n * 4096
        print(json.dumps({"probe": "gram_rect_4096xn", "kernel": name, "n": n, "ms": ms, "GBps": by / ms / 1e6}))


def trsv(n):
    buf, ld = L.empty_matrix(n, n)
    buf.copy_(torch.rand(n, n, dtype=F64, device="cuda") * 1e-3)
    buf.diagonal().add_(1.0)
    nblk = (n + L.DB - 1) // L.DB
    dinv = torch.eye(L.DB, dtype=F64, device="cuda").repeat(nblk, 1, 1).contiguous()
    x = torch.rand(n, dtype=F64, device="cuda")
    for tr in (0, 1):
        ms = timeit(lambda: L.call("stpyb_trsv", L.ptr(buf), n, ld, L.ptr(dinv), L.ptr(x), tr, L.stream_ptr()))
        by = 4.0 * n * n
        print(json.dumps({"probe": "trsv", "transposed": tr, "n": n, "ms": ms, "GBps": by / ms / 1e6}))


def c2parts(n):
    """Where the C2 step (ARD-SE, d=10: fit + value + gradient) spends its time."""
    from stpy_b200 import autodiff
    from stpy_b200.continuous_processes.gauss_procc import GaussianProcess
    d = 10
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(n, d, dtype=F64, generator=g) * 2 - 1).cuda()
    y = torch.sin(3 * x.sum(dim=1, keepdim=True))
    ard0 = torch.linspace(0.8, 1.6, d, dtype=F64)
    k = KF(kernel_name="ard", ard_gamma=ard0.clone(), d=d)
    gp = GaussianProcess(kernel=k, s=0.1)
    print(json.dumps({"probe": "c2_fit", "n": n, "ms": timeit(lambda: gp.fit_gp(x, y), reps=3, warm=1)}))
    f = gp._fit
    work, ldw = L.empty_matrix(n, n)
    kinv, ldk = L.empty_matrix(n, n)
    ms = timeit(lambda: L.call("stpyb_potri", L.ptr(f.buf), n, f.ld, L.ptr(f.dinv), L.ptr(work), ldw, L.ptr(kinv), ldk,
                               L.stream_ptr()), reps=3, warm=1)
    print(json.dumps({"probe": "c2_potri", "n": n, "ms": ms, "tflops": 2 * n ** 3 / 3 / ms / 1e9}))
    a = ard0.clone().requires_grad_(True)
    items, sub_ops = k.grad_plan({'0': {'ard_gamma': a, 'group': list(range(d))}})
    alpha = gp._A_dev

    def passes():
        autodiff._run_passes(items, sub_ops, x, x, 0, kinv, ldk, alpha, 1.0, need_trace=False)
    print(json.dumps({"probe": "c2_grad_pass", "n": n, "ms": timeit(passes, reps=3, warm=1)}))

    def full():
        aa = ard0.clone().requires_grad_(True)
        v = gp.log_marginal(k, {'0': {'ard_gamma': aa}}, 1.0)
        v.backward()
    print(json.dumps({"probe": "c2_value_and_grad", "n": n, "ms": timeit(full, reps=3, warm=1)}))


if __name__ == "__main__":
    mode = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
    L.load()
    {"gram": gram, "trsv": trsv, "c2parts": c2parts}[mode](n)
