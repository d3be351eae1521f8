"""Single-GPU look-ahead (next panel on a high-priority side stream) on/off: fit_gp time per problem size."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import stpy_oracle as O
from stpy_b200 import _lib as L
from stpy_b200.kernels import KernelFunction
from stpy_b200.continuous_processes.gauss_procc import GaussianProcess

L.load()
for n in (4096, 8192, 16384, 32768):
    x, y = O.make_data(n, 8, seed=0)
    x, y = x.cuda(), y.cuda()
    k = KernelFunction(kernel_name="matern", gamma=1.0, nu=2.5, d=8)
    gp = GaussianProcess(kernel=k, s=0.1)
    out = []
    for min_n in (-1, 0):
        L.call("stpyb_set_lookahead_min_n", min_n, None)
        gp.fit_gp(x, y)
        reps = 5 if n <= 16384 else 3
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gp.fit_gp(x, y); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out.append((best, float(gp.log_marginal(k, {}, 1.0))))
    print("n=%6d  fit_gp: look-ahead off %.2f ms, on %.2f ms (%.1f %%), LML equal: %s" %
          (n, out[0][0], out[1][0], 100 * (out[0][0] / out[1][0] - 1), out[0][1] == out[1][1]), flush=True)
    del gp
    torch.cuda.empty_cache()
