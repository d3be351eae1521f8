"""BASELINE.json configurations at FULL size on one B200, checked through size-independent
properties (the CPU oracle cannot run at these sizes inside a test): residuals of the linear
systems that were solved, a finite-difference check of the analytic gradient, agreement of the
batched sweep with single evaluations.  ~40 s on a B200; skipped when the GPU is too small."""
import numpy as np
import pytest
import torch

from conftest import relerr

pytestmark = pytest.mark.gpu
F = torch.float64


def _need_gb(gb):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    free, total = torch.cuda.mem_get_info()
    if free < gb * 2 ** 30:
        pytest.skip("needs %d GB of free device memory" % gb)


def test_c3_matern_n65536_fit_lml_residuals():
    """C3: Matern-5/2, n = 65 536, d = 8.  K alpha = y checked block-row-wise without ever holding K."""
    _need_gb(60)
    from oracle import stpy_oracle as O
    from stpy_b200.continuous_processes.gauss_procc import GaussianProcess
    from stpy_b200.kernels import KernelFunction as KF
    n, d = 65536, 8
    x, y = O.make_data(n, d, seed=0)
    xd, yd = x.cuda(), y.cuda()
    k = KF(kernel_name="matern", gamma=1.0, nu=2.5, d=d)
    gp = GaussianProcess(kernel=k, s=0.1)
    gp.fit_gp(xd, yd)
    lml = float(gp.log_marginal(k, {}, 1.0))
    assert np.isfinite(lml)
    worst = 0.0
    for lo in range(0, n, 8192):
        rows = k.kernel(xd, xd[lo:lo + 8192])               # (8192, n) block row of the Gram matrix
        r = rows @ gp.A + 0.01 * gp.A[lo:lo + 8192] - yd[lo:lo + 8192]
        worst = max(worst, float(r.abs().max()))
        del rows
    assert worst / float(yd.abs().max()) < 1e-9
    # the quadratic form of the evidence equals y^T alpha; the remaining term is the log-determinant
    quad = float((yd * gp.A).sum())
    logdet = 2.0 * (lml - 0.5 * quad)
    assert abs(quad - float(gp._fit.out3[0])) < 1e-8 * abs(quad)
    assert np.isfinite(logdet)
    mu, sd = gp.mean_std(xd[:256])
    assert relerr(mu + 0.01 * gp.A[:256], yd[:256]) < 1e-9  # K* alpha at training inputs
    assert float(sd.min()) > 0.0 and float(sd.max()) < 0.1 + 1e-9  # posterior std at data <= noise level


def test_c2_ard_n16384_gradient_matches_finite_differences():
    _need_gb(12)
    from oracle import stpy_oracle as O
    from stpy_b200.continuous_processes.gauss_procc import GaussianProcess
    from stpy_b200.kernels import KernelFunction as KF
    n, d = 16384, 10
    x, y = O.make_data(n, d, seed=0)
    ard0 = torch.linspace(0.8, 1.6, d, dtype=F)
    k = KF(kernel_name="ard", ard_gamma=ard0.clone(), d=d)
    gp = GaussianProcess(kernel=k, s=0.1)
    gp.fit_gp(x.cuda(), y.cuda())
    a = ard0.clone().requires_grad_(True)
    v = gp.log_marginal(k, {'0': {'ard_gamma': a}}, 1.0)
    v.backward()
    torch.manual_seed(0)
    u = torch.randn(d, dtype=F)
    u /= u.norm()
    h = 1e-5
    vp = float(gp.log_marginal(k, {'0': {'ard_gamma': ard0 + h * u}}, 1.0))
    vm = float(gp.log_marginal(k, {'0': {'ard_gamma': ard0 - h * u}}, 1.0))
    fd = (vp - vm) / (2 * h)
    assert abs(float(a.grad @ u) - fd) < 1e-6 * abs(fd)
    assert abs(float(v.detach()) - float(gp.log_marginal(k, {}, 1.0))) < 1e-8  # same point as the fit


def test_c4_rff_n1e6_m8192_normal_equations():
    _need_gb(20)
    from oracle import stpy_oracle as O
    from stpy_b200.continuous_processes.kernelized_features import KernelizedFeatures
    from stpy_b200.embeddings.embedding import RFFEmbedding
    n, d, m = 10 ** 6, 16, 8192
    x, y = O.make_data(n, d, seed=0)
    xd, yd = x.cuda(), y.cuda()
    np.random.seed(0)
    emb = RFFEmbedding(gamma=1.0, m=m, d=d)
    kf = KernelizedFeatures(embedding=emb, m=m, s=0.1, lam=1.0, d=d)
    kf.fit_gp(xd, yd)
    theta = kf._theta
    lhs = torch.zeros(m, dtype=F, device="cuda")
    rhs = torch.zeros(m, dtype=F, device="cuda")
    for lo in range(0, n, 50000):
        phi, _ = emb.embed_device(xd[lo:lo + 50000])
        lhs += phi.T @ (phi @ theta)
        rhs += phi.T @ yd[lo:lo + 50000].reshape(-1)
    lhs += 0.01 * theta
    assert relerr(lhs, rhs) < 1e-10
    mu, sd = kf.mean_std(xd[:256])
    assert bool(torch.isfinite(mu).all()) and float(sd.min()) > 0.0


def test_c5_sweep_64_kernels_n8192():
    _need_gb(12)
    from oracle import stpy_oracle as O
    from stpy_b200.continuous_processes.gauss_procc import GaussianProcess
    from stpy_b200.kernels import KernelFunction as KF
    from stpy_b200.sweep import lml_sweep
    n, d = 8192, 4
    x, y = O.make_data(n, d, seed=0)
    gam = np.logspace(-1, 0.5, 32)
    ks = [KF(kernel_name="squared_exponential", gamma=float(g), d=d) for g in gam] + \
         [KF(kernel_name="matern", gamma=float(g), nu=2.5, d=d) for g in gam]
    vals = lml_sweep(ks, x.cuda(), y.cuda(), s=0.1)
    assert vals.shape == (64,) and bool(torch.isfinite(vals).all())
    for i in (0, 17, 40, 63):
        gp = GaussianProcess(kernel=ks[i], s=0.1)
        gp.fit_gp(x.cuda(), y.cuda())
        ref = float(gp.log_marginal(ks[i], {}, 1.0))
        assert abs(float(vals[i]) - ref) < 1e-8 * max(1.0, abs(ref))
