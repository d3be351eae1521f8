"""CPU restatement of stpy's Gaussian-process hot path.  TEST INFRASTRUCTURE ONLY.

This module is the parity checker for the CUDA path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import it; nothing under stpy_b200/ does.

Every function restates, in plain torch-CPU float64 / numpy / scipy (the same
third-party arithmetic the reference itself calls: torch 2.11.0, numpy 2.3,
scipy 1.18 in this image -- stpy pins no versions, setup.py:3-18), the
algorithm of the reference lines cited in its docstring.  Paths are relative to
the stpy repository root.

PINNING.  The reference ships no assertions, golden vectors or known-answer
tests for this path, so the oracle is pinned against
outputs of the unmodified reference run in the build container:
tests/golden/make_golden.py imports /root/reference, runs fit_gp / mean_std /
log_marginal (+backward) / RFFEmbedding.embed / KernelizedFeatures on seeded
inputs and stores them in tests/golden/*.npz; tests/test_oracle_golden.py checks
every function below against those fixtures.
"""
import math

import numpy as np
import torch

try:  # scipy is only needed for the isotropic Matern kernel (as in the reference)
    from scipy.spatial.distance import cdist as _cdist
except Exception:  # pragma: no cover
    _cdist = None

F64 = torch.float64


# --------------------------------------------------------------------------- kernels
def se_kernel(a, b, gamma=1.0, kappa=1.0, group=None):
    """stpy/kernels.py:368-398 squared_exponential_kernel; returns (|b|, |a|)."""
    if group is not None:
        a, b = a[:, group], b[:, group]
    normx = torch.sum(a ** 2, dim=1).view(-1, 1)
    normy = torch.sum(b ** 2, dim=1).view(-1, 1)
    product = torch.mm(b, torch.t(a))
    sqdist = -2 * product + torch.t(normx) + normy
    return kappa * torch.exp((-0.5 / (gamma * gamma)) * sqdist)


def ard_kernel(a, b, ard_gamma, kappa=1.0, group=None):
    """stpy/kernels.py:552-583 ard_kernel (inputs scaled by diag(1/gamma) via mm, then SE with gamma=1)."""
    if group is None:
        group = list(range(a.shape[1]))
    a, b = a[:, group], b[:, group]
    D = torch.diag(1. / (ard_gamma[group]))
    a = torch.mm(a, D)
    b = torch.mm(b, D)
    normx = torch.sum(a ** 2, dim=1).reshape(-1, 1)
    normy = torch.sum(b ** 2, dim=1).reshape(-1, 1)
    product = torch.mm(b, torch.t(a))
    sqdist = -2 * product + torch.t(normx) + normy
    return kappa * torch.exp(-0.5 * sqdist)


def ard_kernel_additive(a, b, ard_gamma, groups, kappa=1.0):
    """stpy/kernels.py:700-729 ard_kernel_additive: mean over groups of ard_kernel."""
    r = torch.zeros((b.shape[0], a.shape[0]), dtype=F64)
    for g in groups:
        r = r + ard_kernel(a, b, ard_gamma, kappa=kappa, group=g)
    return r / float(len(groups))


def se_per_group_kernel(a, b, gamma_per_group, groups, kappa=1.0):
    """stpy/kernels.py:667-698 squared_exponential_per_group_kernel_additive: the mean over groups of SE Grams
    with one lengthscale per group.  kappa enters twice, once inside each SE term (it rides in kwargs)
    and once on the sum -- restated as written."""
    r = torch.zeros(b.shape[0], a.shape[0], dtype=F64)
    for group_add, gamma in zip(groups, gamma_per_group):
        r = r + se_kernel(a, b, gamma=gamma, kappa=kappa, group=group_add)
    return kappa * r / float(len(groups))


def ard_per_group_kernel(a, b, ard_per_group, groups, kappa=1.0):
    """stpy/kernels.py:618-665 ard_per_group_kernel_additive: the lengthscale vector is consumed group by group."""
    r = torch.zeros(b.shape[0], a.shape[0], dtype=F64)
    at = 0
    for group_add in groups:
        gamma = ard_per_group[at:at + len(group_add)]
        at += len(group_add)
        D = torch.diag(1. / gamma)
        ax, bx = torch.mm(a[:, group_add], D), torch.mm(b[:, group_add], D)
        normx = torch.sum(ax ** 2, dim=1).reshape(-1, 1)
        normy = torch.sum(bx ** 2, dim=1).reshape(-1, 1)
        sqdist = -2 * torch.mm(bx, torch.t(ax)) + torch.t(normx) + normy
        r = r + torch.exp(-0.5 * sqdist)
    return kappa * (r / float(len(groups)))


def _matern_map(dists, nu):
    """stpy/kernels.py:844-859 / 954-962: nu in {0.5, 1.5, 2.5} closed forms, else the Bessel-function form."""
    exp = torch.exp if torch.is_tensor(dists) else np.exp
    if nu == 0.5:
        return exp(-dists)
    if nu == 1.5:
        K = dists * math.sqrt(3)
        return (1. + K) * exp(-K)
    if nu == 2.5:
        K = dists * math.sqrt(5)
        return (1. + K + K ** 2 / 3.0) * exp(-K)
    if torch.is_tensor(dists):
        raise NotImplementedError("general nu: the reference's torch branch is broken (kernels.py:964-970)")
    # kernels.py:852-859, numpy branch: zeros moved to eps, 2^(1-nu)/Gamma(nu) t^nu K_nu(t), t = sqrt(2 nu) r
    from scipy.special import kv
    K = np.array(dists, dtype=np.float64, copy=True)
    K[K == 0.0] += np.finfo(float).eps
    tmp = math.sqrt(2 * nu) * K
    out = np.full_like(K, (2 ** (1. - nu)) / math.gamma(nu))
    out *= tmp ** nu
    out *= kv(nu, tmp)
    return out


def matern_kernel(a, b, gamma=1.0, nu=2.5, kappa=1.0, group=None):
    """stpy/kernels.py:811-859 matern_kernel: scipy cdist on a/gamma, b/gamma (direct differences)."""
    if group is not None:
        a, b = a[:, group], b[:, group]
    dists = _cdist(a.numpy() / gamma, b.numpy() / gamma, metric='euclidean').T
    return kappa * torch.from_numpy(_matern_map(dists, nu))


def ard_matern_kernel(a, b, ard_gamma, nu=2.5, kappa=1.0, group=None):
    """stpy/kernels.py:917-970 ard_matern_kernel: torch.cdist (GEMM expansion above 25 rows)."""
    if group is None:
        group = list(range(a.shape[1]))
    D = torch.diag(1. / (ard_gamma[group]))
    a = torch.mm(a, D)[:, group]
    b = torch.mm(b, D)[:, group]
    dists = torch.cdist(a, b, p=2).T
    return kappa * _matern_map(dists, nu)


def polynomial_kernel(a, b, degree=2, kappa=1.0, group=None):
    """stpy/kernels.py:766-784 polynomial_kernel."""
    if group is not None:
        a, b = a[:, group], b[:, group]
    return kappa * (torch.mm(b, torch.t(a)) + 1) ** degree


def linear_kernel(a, b, kappa=1.0, offset=0.0, group=None):
    """stpy/kernels.py:300-320 linear_kernel."""
    if group is not None:
        a, b = a[:, group], b[:, group]
    return kappa * (b @ a.T) + offset


# --------------------------------------------------------------------------- GP, as written
def fit_gp_as_written(kernel, x, y, s):
    """stpy/continuous_processes/gauss_procc.py:136-177 + 336-378 (back_prop=True branch).

    Faithful to the reference's cost profile: dense Sigma^T Sigma product, a second Gram
    inside mean_std(x), and pivoted-QR least-squares solves with 1 and n right-hand sides.
    Returns (K, A)."""
    n = x.shape[0]
    Sigma = s * torch.eye(n, dtype=F64)
    K = kernel(x, x) + Sigma.T @ Sigma
    K_star = kernel(x, x)
    A = torch.linalg.lstsq(K, y)[0]
    B = torch.t(torch.linalg.lstsq(K, torch.t(K_star))[0])
    _ = torch.einsum('ij,ji->i', (B, torch.t(K_star)))
    return K, A


def mean_std_as_written(kernel, x, y, s, xtest, K=None):
    """gauss_procc.py:336-401: K* = kernel(x, xtest), lstsq solves, einsum variance; returns (mean, std)."""
    n = x.shape[0]
    if K is None:
        K = kernel(x, x) + (s * s) * torch.eye(n, dtype=F64)
    K_star = kernel(x, xtest)
    diag = torch.hstack([kernel(xtest[i, :].view(1, -1), xtest[i, :].view(1, -1)).view(1)
                         for i in range(xtest.shape[0])])
    A = torch.linalg.lstsq(K, y)[0]
    B = torch.t(torch.linalg.lstsq(K, torch.t(K_star))[0])
    ymean = torch.mm(K_star, A)
    var = diag.view(-1, 1) - torch.einsum('ij,ji->i', (B, torch.t(K_star))).view(-1, 1)
    return ymean, torch.sqrt(var)


def lml_as_written(kernel, x, y, s, weight=1.0):
    """gauss_procc.py:631-638 _log_marginal_squared: slogdet (LU) + solve (LU); (1,1) tensor."""
    n = x.shape[0]
    K = kernel(x, x) + torch.eye(n, dtype=F64) * s * s
    logdet = -0.5 * torch.slogdet(K)[1] * weight
    alpha = torch.linalg.solve(K, y)
    logprob = -0.5 * torch.mm(torch.t(y), alpha) + logdet
    return -logprob


# --------------------------------------------------------------------------- GP, Cholesky restatement
def lml_cholesky(kernel, x, y, s, weight=1.0):
    """stpy/estimator.py:32-40 Estimator.log_marginal (Cholesky + cholesky_solve)."""
    n = x.shape[0]
    K = kernel(x, x) + torch.eye(n, dtype=F64) * s * s
    L = torch.linalg.cholesky(K)
    logdet = -0.5 * 2 * torch.sum(torch.log(torch.diag(L))) * weight
    alpha = torch.cholesky_solve(y, L)
    logprob = -0.5 * torch.mm(torch.t(y), alpha) + logdet
    return -logprob


def gp_cholesky(kernel, x, y, s, xtest=None, full=False):
    """One-Cholesky restatement of fit_gp + mean_std (gauss_procc.py:163, 381, 391-399).

    Used where the as-written path is too slow / too large; validated against it in
    tests/test_oracle_golden.py.  Returns dict(A, L, mean, std | cov)."""
    n = x.shape[0]
    K = kernel(x, x) + (s * s) * torch.eye(n, dtype=F64)
    L = torch.linalg.cholesky(K)
    A = torch.cholesky_solve(y, L)
    out = {"A": A, "L": L}
    if xtest is not None:
        K_star = kernel(x, xtest)
        V = torch.linalg.solve_triangular(L, K_star.T, upper=False)  # (n, nt)
        out["mean"] = K_star @ A
        if full:
            out["cov"] = kernel(xtest, xtest) - V.T @ V
        else:
            diag = torch.stack([kernel(xtest[i:i + 1], xtest[i:i + 1]).reshape(()) for i in range(xtest.shape[0])])
            out["std"] = torch.sqrt(diag.view(-1, 1) - (V * V).sum(0).view(-1, 1))
    return out


def mixture_weights(kernels, x, y, s, init_weights=None):
    """categorical_mixture.py:36-71 CategoricalMixture.log_prob_normal / fit_gp.

    log p_j = -0.5 y^T K_j^-1 y - 0.5 logdet K_j - 0.5 n log(2 pi) with K_j = k_j(x, x) + s^2 I (the reference
    gets there through an LU factorisation and slogdet), posterior weights by log-sum-exp against the prior
    weights.  Returns (logprobs (k,), weights (k,))."""
    k = len(kernels)
    w0 = torch.ones(k, dtype=F64) / k if init_weights is None else init_weights
    n = x.shape[0]
    logp = torch.stack([(-lml_cholesky(kj, x, y, s, 1.0)).reshape(()) - 0.5 * n * math.log(2 * math.pi)
                        for kj in kernels])
    log_post = torch.log(w0) + logp
    return logp, torch.exp(log_post - torch.logsumexp(log_post, dim=0))


def mixture_mean_std(kernels, weights, x, y, s, xtest):
    """categorical_mixture.py:73-83: weighted mean, sqrt of the weighted variances."""
    mu = torch.zeros(xtest.shape[0], 1, dtype=F64)
    var = torch.zeros(xtest.shape[0], 1, dtype=F64)
    for kj, wj in zip(kernels, weights):
        r = gp_cholesky(kj, x, y, s, xtest)
        mu = mu + wj * r["mean"]
        var = var + wj * r["std"] ** 2
    return mu, torch.sqrt(var)


def lml_grad_ard(x, y, s, ard_gamma, kappa=1.0, weight=1.0):
    """Autograd gradient of lml_as_written w.r.t. ard_gamma, kappa and s, exactly as
    optimize_params differentiates it (estimator.py:156-171 -> gauss_procc.py:631-638)."""
    g = ard_gamma.clone().requires_grad_(True)
    kap = torch.tensor(float(kappa), dtype=F64, requires_grad=True)
    st = torch.tensor(float(s), dtype=F64, requires_grad=True)
    n = x.shape[0]
    K = ard_kernel(x, x, g, kappa=kap) + torch.eye(n, dtype=F64) * st * st
    logdet = -0.5 * torch.slogdet(K)[1] * weight
    alpha = torch.linalg.solve(K, y)
    val = -(-0.5 * torch.mm(torch.t(y), alpha) + logdet)
    val.backward()
    return val.detach(), g.grad.detach(), kap.grad.detach(), st.grad.detach()


# --------------------------------------------------------------------------- RFF + BLR
def rff_embed(x, W, b=None, kappa=1.0):
    """stpy/embeddings/embedding.py:225-241 RFFEmbedding.embed.

    Unbiased (b is None): rows 0..m/2 of W x^T through cos, rows m/2..m through sin; (n, m).
    Biased: sqrt(2/m) cos(W x^T + b), transposed back to (n, m).  (The reference's biased
    branch applies torch.t twice and so returns (m, n); callers that need its exact layout
    transpose this result.)"""
    m, d = W.shape[0], x.shape[1]
    if b is not None:
        z = np.sqrt(2. / m) * torch.cos(W[:, 0:d].mm(torch.t(x)) + b.view(m, 1))
    else:
        q = W[:, 0:d].mm(torch.t(x))
        z1 = np.sqrt(2. / float(m)) * torch.cos(q[0:int(m / 2), :])
        z2 = np.sqrt(2. / float(m)) * torch.sin(q[int(m / 2):m, :])
        z = torch.cat([z1, z2])
    return torch.t(z) * np.sqrt(kappa)


def qff_embed(x, W, weights, kappa=1.0):
    """stpy/embeddings/embedding.py:450-466 QuadratureEmbedding.embed (cosine=False): the same
    frequencies feed sqrt(w) cos and sqrt(w) sin; returns (n, 2*len(W))."""
    d = x.shape[1]
    q = torch.mm(W[:, 0:d], torch.t(x))
    sw = torch.sqrt(weights.view(-1, 1))
    z = torch.cat([sw * torch.cos(q), sw * torch.sin(q)])
    return torch.t(z) * np.sqrt(kappa)


def blr_as_written(Phi, y, s, lam, Phi_test):
    """kernelized_features.py:236-240, 256, 274-288 (primal): V = Q^T Q + s^2 lam I,
    invV = pinverse(V), theta = invV Q^T y, mean = Phi* theta, std = s sqrt(diag(Phi* invV Phi*^T))."""
    m = Phi.shape[1]
    V = Phi.T @ Phi + s ** 2 * lam * torch.eye(m, dtype=F64)
    invV = torch.pinverse(V)
    theta = invV @ Phi.T @ y
    mean = Phi_test @ theta
    std = torch.sqrt(s ** 2 * torch.einsum('ij,jk,ik->i', (Phi_test, invV, Phi_test)).view(-1, 1))
    return theta, mean, std


def blr_cholesky(Phi, y, s, lam, Phi_test):
    """Same quantities through one Cholesky of V (SPD since s^2 lam > 0)."""
    m = Phi.shape[1]
    V = Phi.T @ Phi + s ** 2 * lam * torch.eye(m, dtype=F64)
    L = torch.linalg.cholesky(V)
    theta = torch.cholesky_solve(Phi.T @ y, L)
    mean = Phi_test @ theta
    Vt = torch.linalg.solve_triangular(L, Phi_test.T, upper=False)
    std = torch.sqrt(s ** 2 * (Vt * Vt).sum(0).view(-1, 1))
    return theta, mean, std


# --------------------------------------------------------------------------- synthetic workloads
def make_data(n, d, seed=0, noise=0.1):
    """x ~ U[-1,1]^d with full 53-bit mantissas, y = sin(3 sum_j x_j) + noise."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, d, dtype=F64, generator=g) * 2 - 1
    y = torch.sin(3 * x.sum(dim=1, keepdim=True)) + noise * torch.randn(n, 1, dtype=F64, generator=g)
    return x, y


def fit_lml_flops(n, d):
    """Algorithmic flops of fit + LML: n^3/3 + 2 d n^2 + 4 n^2."""
    return n ** 3 / 3.0 + 2.0 * d * n ** 2 + 4.0 * n ** 2
