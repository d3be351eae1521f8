"""Small end-to-end pass over every kernel family at ragged, odd sizes with the two-stream look-ahead forced
on, checked against the CPU oracle (it found the odd-n alignment bug of the sweep's work vectors)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import stpy_oracle as O
from stpy_b200 import _lib as L
from stpy_b200.kernels import KernelFunction as KF
from stpy_b200.continuous_processes.gauss_procc import GaussianProcess
from stpy_b200.continuous_processes.kernelized_features import KernelizedFeatures
from stpy_b200.embeddings.embedding import RFFEmbedding
from stpy_b200.sweep import lml_sweep

L.load()
L.call("stpyb_set_lookahead_min_n", 0, None)   # exercise the two-stream schedule at these sizes
rel = lambda a, b: float((a.cpu() - b).norm() / b.norm())

x, y = O.make_data(333, 3, seed=1)
xt, _ = O.make_data(37, 3, seed=2)
ard = torch.tensor([0.7, 1.1, 0.9], dtype=torch.float64)
cases = [
    (KF(kernel_name="squared_exponential", gamma=0.6, d=3), lambda a, b: O.se_kernel(a, b, gamma=0.6)),
    (KF(kernel_name="matern", gamma=0.9, nu=2.5, d=3), lambda a, b: O.matern_kernel(a, b, gamma=0.9, nu=2.5)),
    (KF(kernel_name="ard_matern", ard_gamma=ard, nu=1.5, d=3), lambda a, b: O.ard_matern_kernel(a, b, ard, nu=1.5)),
    (KF(kernel_name="ard", ard_gamma=ard, d=3) + KF(kernel_name="polynomial", power=2, kappa=0.1, d=3),
     lambda a, b: O.ard_kernel(a, b, ard) + O.polynomial_kernel(a, b, degree=2, kappa=0.1)),
]
for k, ko in cases:
    gp = GaussianProcess(kernel=k, s=0.1)
    gp.outer_block = 128
    gp.fit_gp(x, y)
    mu, sd = gp.mean_std(xt)
    r = O.gp_cholesky(ko, x, y, 0.1, xt)
    assert rel(mu, r["mean"]) < 1e-10 and rel(sd ** 2, r["std"] ** 2) < 1e-10
    assert abs(float(gp.log_marginal(k, {}, 1.0)) - float(O.lml_cholesky(ko, x, y, 0.1))) < 1e-8
    mu_f, cov = gp.mean_std(xt[:9], full=True)
    gp.add_data_point(xt[:5], torch.zeros(5, 1, dtype=torch.float64))
    assert gp.n == 338
# gradient (potri + fused reduce)
g = torch.tensor([0.8, 1.2, 1.0], dtype=torch.float64, requires_grad=True)
k = KF(kernel_name="ard", ard_gamma=ard.clone(), d=3)
gp = GaussianProcess(kernel=k, s=0.1)
gp.fit_gp(x, y)
val = gp.log_marginal(k, {'0': {'ard_gamma': g}}, 1.0)
val.backward()
_, gref, _, _ = O.lml_grad_ard(x, y, 0.1, g.detach())
assert rel(g.grad, gref) < 1e-8
# random features + Bayesian linear regression
np.random.seed(3)
emb = RFFEmbedding(gamma=0.8, m=96, d=3)
kf = KernelizedFeatures(embedding=emb, m=96, s=0.1, lam=1.0, d=3)
kf.fit_gp(x, y)
mu, sd = kf.mean_std(xt)
th, mr, sr = O.blr_cholesky(O.rff_embed(x, emb.W), y, 0.1, 1.0, O.rff_embed(xt, emb.W))
assert rel(mu, mr) < 1e-9 and rel(sd, sr) < 1e-9
# sweep (gram_multi + multi-stream factorisations)
ks = [KF(kernel_name="squared_exponential", gamma=g_, d=3) for g_ in (0.5, 0.9)] + \
     [KF(kernel_name="matern", gamma=1.1, nu=2.5, d=3)]
v = lml_sweep(ks, x, y, s=0.1)
assert abs(float(v[0]) - float(O.lml_cholesky(lambda a, b: O.se_kernel(a, b, gamma=0.5), x, y, 0.1))) < 1e-8
torch.cuda.synchronize()
print("ragged check ok")
