"""fit_gp time vs outer block (K depth of the trailing update) for several n."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import stpy_oracle as O
from stpy_b200.kernels import KernelFunction as KF
from stpy_b200.continuous_processes.gauss_procc import GaussianProcess
for n in (1024, 2048, 4096, 8192, 16384, 32768):
    x, y = O.make_data(n, 8, seed=0)
    xd, yd = x.cuda(), y.cuda()
    k = KF(kernel_name="matern", gamma=1.0, nu=2.5, d=8)
    row = []
    for outer in (128, 256, 512, 1024):
        gp = GaussianProcess(kernel=k, s=0.1)
        gp.outer_block = outer
        gp.fit_gp(xd, yd); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter(); gp.fit_gp(xd, yd); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
        row.append("%d: %.2f ms" % (outer, best * 1e3))
    print("n=%6d  " % n + "   ".join(row), flush=True)
