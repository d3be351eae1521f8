"""Cycle breakdown of potrf_diag_kernel's phases (debug instrumentation)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stpy_b200 import _lib as L
L.load()
n = 128
X = torch.randn(n, 64, dtype=torch.float64)
K = (X @ X.T / 64 + 0.5 * torch.eye(n, dtype=torch.float64))
for rep in range(3):
    Kd, ld = L.empty_matrix(n, n); Kd.copy_(K)
    Linv = torch.empty(128, 128, dtype=torch.float64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = torch.zeros(8, dtype=torch.int64, device="cuda")
    L.call("stpyb_potrf_diag_profile", L.ptr(Kd), ld, n, L.ptr(Linv), L.ptr(info), L.ptr(st), L.stream_ptr())
    torch.cuda.synchronize()
    s = st.cpu().tolist()
    print("total %d | load %d | factor %d (leaf %d, solve %d, update %d) | assemble %d | store %d" % (
        s[4] - s[0], s[1] - s[0], s[2] - s[1], s[5], s[6], s[7], s[3] - s[2], s[4] - s[3]))
Lref = torch.linalg.cholesky(K)
print("L err", float((torch.tril(Kd.cpu()) - Lref).abs().max()), "inv err", float((Linv.cpu() @ Lref - torch.eye(n, dtype=torch.float64)).abs().max()))
