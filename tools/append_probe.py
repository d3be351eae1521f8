"""Sequential design at scale: cost of GaussianProcess.add_data_point (bordered factor) vs a full refit."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import stpy_oracle as O
from stpy_b200.kernels import KernelFunction
from stpy_b200.continuous_processes.gauss_procc import GaussianProcess

def timed(fn, reps=1):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / reps * 1e3

for n in (4096, 16384, 32768):
    x, y = O.make_data(n + 300, 8, seed=0)
    x, y = x.cuda(), y.cuda()
    xt = x[:256]
    k = KernelFunction(kernel_name="matern", gamma=1.0, nu=2.5, d=8)
    gp = GaussianProcess(kernel=k, s=0.1)
    gp.fit_gp(x[:n], y[:n])
    t_fit = timed(lambda: gp.fit_gp(x[:n], y[:n]))
    gp.add_data_point(x[n:n + 1], y[n:n + 1])  # grows the buffer once
    state = {"i": n + 1}
    def one():
        i = state["i"]; gp.add_data_point(x[i:i + 1], y[i:i + 1]); state["i"] = i + 1
    t_one = timed(one, reps=200)
    i = state["i"]
    t_batch = timed(lambda: gp.add_data_point(x[i:i + 64], y[i:i + 64]))
    fresh = GaussianProcess(kernel=k, s=0.1); fresh.incremental = False
    fresh.fit_gp(gp.x, gp.y)
    ea = float((gp.A - fresh.A).norm() / fresh.A.norm())
    m1, s1 = gp.mean_std(xt); m2, s2 = fresh.mean_std(xt)
    em = float((m1 - m2).norm() / m2.norm())
    print("n=%d: refit %.1f ms | append 1 point %.2f ms (mean of 200) | append 64 points %.2f ms | "
          "alpha relerr vs refit %.1e, mean relerr %.1e" % (n, t_fit, t_one, t_batch, ea, em), flush=True)
    del gp, fresh
    torch.cuda.empty_cache()
