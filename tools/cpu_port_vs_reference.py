"""Build-container only (needs /root/reference): time the UNMODIFIED reference's fit_gp + log_marginal against the
oracle's as-written port on the same inputs, to show that the CPU arm bench.py times (the port -- the reference's
sources may not be redistributed, so they cannot travel to the GPU box) has the reference's cost profile.
Writes profiles/cpu_port_vs_reference_r02.txt."""
import os
import sys
import time
from unittest.mock import MagicMock

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
for _m in ("cvxpy matplotlib matplotlib.pyplot cvxpylayers cvxpylayers.torch pymanopt pymanopt.manifolds "
           "pymanopt.optimizers pymanopt.function torchmin autograd_minimize mosek").split():
    sys.modules[_m] = MagicMock()
sys.path.insert(0, os.environ.get("STPY_REFERENCE", "/root/reference"))
from stpy.kernels import KernelFunction  # noqa: E402
from stpy.continuous_processes.gauss_procc import GaussianProcess  # noqa: E402
from oracle import stpy_oracle as O  # noqa: E402

lines = ["# unmodified reference (import shim) vs oracle port, Matern-5/2, d=8, s=0.1, %d threads, best of 2"
         % torch.get_num_threads(),
         "# n  reference fit_gp+log_marginal [s]  port fit_gp_as_written+lml_as_written [s]  ratio  |lml diff|"]
for n in (1024, 2048, 4096):
    x, y = O.make_data(n, 8, seed=0)
    tr = tp = 1e30
    for _ in range(2):
        k = KernelFunction(kernel_name="matern", gamma=1.0, nu=2.5, d=8)
        gp = GaussianProcess(kernel=k, s=0.1)
        t0 = time.perf_counter()
        gp.fit_gp(x, y)
        v_ref = float(gp.log_marginal(k, {}, 1.0))
        tr = min(tr, time.perf_counter() - t0)
        kern = lambda a, b: O.matern_kernel(a, b, gamma=1.0, nu=2.5)
        t0 = time.perf_counter()
        O.fit_gp_as_written(kern, x, y, 0.1)
        v_port = float(O.lml_as_written(kern, x, y, 0.1, 1.0))
        tp = min(tp, time.perf_counter() - t0)
    lines.append("%d  %.3f  %.3f  %.2f  %.2e" % (n, tr, tp, tp / tr, abs(v_ref - v_port)))
    print(lines[-1], flush=True)
open(os.path.join(ROOT, "profiles", "cpu_port_vs_reference_r02.txt"), "w").write("\n".join(lines) + "\n")
