"""Per-step GPU time of the distributed schedule on ONE rank (no communication) for small trailing
sizes, plus the cost of one panel factorisation (stpyb_potrf_panel) as a function of its height."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import stpy_oracle as O
from stpy_b200 import _lib as L
from stpy_b200.kernels import KernelFunction
from stpy_b200.distributed import DistributedGP
L.load()
k = KernelFunction(kernel_name="matern", gamma=1.0, nu=2.5, d=8)
x, y = O.make_data(16384, 8, seed=0)
dg = DistributedGP(k, s=0.1, nbw=512)
dg.fit_gp(x.cuda(), y.cuda())
dg.profile = True
dg.fit_gp(x.cuda(), y.cuda())
print("world=1 n=16384 per-step ms:", dg.phase_ms["step_ms"])
# panel factorisation alone
for rows in (65536, 32768, 16384, 8192, 4096, 2048, 1024, 512):
    P, ld = L.empty_matrix(rows, 512)
    X = torch.randn(rows, 64, dtype=torch.float64, device="cuda")
    top = X[:512] @ X[:512].T / 64 + 2.0 * torch.eye(512, dtype=torch.float64, device="cuda")
    dinv = torch.empty(4 * 128 * 128, dtype=torch.float64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    best = 1e9
    for rep in range(4):
        P.normal_(); P[:512] = top
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.call("stpyb_potrf_panel", L.ptr(P), rows, 512, ld, L.ptr(dinv), L.ptr(info), 0, None, 0, L.stream_ptr())
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("potrf_panel rows=%6d w=512: %.3f ms" % (rows, best))
