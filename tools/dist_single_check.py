"""Single-GPU sanity run of the distributed schedule (world = 1, DeviceOps) against GaussianProcess."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import stpy_oracle as O
from stpy_b200.kernels import KernelFunction
from stpy_b200.continuous_processes.gauss_procc import GaussianProcess
from stpy_b200.distributed import DistributedGP

for n, nbw in ((1000, 128), (3000, 256), (5001, 512)):
    x, y = O.make_data(n, 8, seed=0)
    k = KernelFunction(kernel_name="matern", gamma=1.0, nu=2.5, d=8)
    gp = GaussianProcess(kernel=k, s=0.1)
    gp.fit_gp(x.cuda(), y.cuda())
    ref = float(gp.log_marginal(k, {}, 1.0))
    dg = DistributedGP(k, s=0.1, nbw=nbw)
    dg.fit_gp(x.cuda(), y.cuda())
    val = float(dg.log_marginal(1.0))
    ea = float((dg.A - gp.A).abs().max() / gp.A.abs().max())
    print("n=%d nbw=%d lml diff %.3e alpha relerr %.3e" % (n, nbw, abs(val - ref), ea))
    assert abs(val - ref) < 1e-8 and ea < 1e-9
print("dist single ok")
