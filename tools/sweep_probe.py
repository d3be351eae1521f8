"""Sweep throughput (C5: 64 kernels, n=8192) vs number of streams / outer block."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import stpy_oracle as O
from stpy_b200.kernels import KernelFunction as KF
from stpy_b200.sweep import lml_sweep
x, y = O.make_data(8192, 4, seed=0)
gam = np.logspace(-1, 0.5, 32)
ks = [KF(kernel_name="squared_exponential", gamma=float(g), d=4) for g in gam] + [KF(kernel_name="matern", gamma=float(g), nu=2.5, d=4) for g in gam]
xd, yd = x.cuda(), y.cuda()
lml_sweep(ks[:8], xd, yd, s=0.1)
for streams in (2, 4, 8, 16):
    for outer in (128, 256, 512):
        for batch in (16, 32):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            v = lml_sweep(ks, xd, yd, s=0.1, streams=streams, outer_block=outer, batch=batch)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            print("streams=%2d outer=%3d batch=%2d: %.3f s  (%.1f TFLOP/s)" % (streams, outer, batch, dt, 64 * 8192 ** 3 / 3 / dt / 1e12), flush=True)
