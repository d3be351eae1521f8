"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference's optional dependencies that are not installed here (cvxpy,
matplotlib, pymanopt, torchmin, ...) are imported at module top by stpy but never
touched by the squared-loss path, so they are stubbed with MagicMock before import
(SURVEY.md section 8c).  Nothing in the test-suite imports this file or the reference.
"""
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("STPY_REFERENCE", "/root/reference")

for _m in ("cvxpy matplotlib matplotlib.pyplot cvxpylayers cvxpylayers.torch pymanopt pymanopt.manifolds "
           "pymanopt.optimizers pymanopt.function torchmin autograd_minimize mosek").split():
    sys.modules[_m] = MagicMock()
sys.path.insert(0, REF)

from stpy.kernels import KernelFunction  # noqa: E402
from stpy.continuous_processes.gauss_procc import GaussianProcess  # noqa: E402
from stpy.estimator import Estimator  # noqa: E402
from stpy.embeddings.embedding import RFFEmbedding, HermiteEmbedding, QuadratureEmbedding  # noqa: E402
from stpy.continuous_processes.kernelized_features import KernelizedFeatures  # noqa: E402

F64 = torch.float64


def data(n, d, seed=0, noise=0.1):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, d, dtype=F64, generator=g) * 2 - 1
    y = torch.sin(3 * x.sum(dim=1, keepdim=True)) + noise * torch.randn(n, 1, dtype=F64, generator=g)
    return x, y


def save(name, **arrays):
    out = {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items()})


def gram_cases():
    a, _ = data(48, 3, seed=10)
    b, _ = data(24, 3, seed=11)
    ard = torch.tensor([0.7, 1.3, 0.9], dtype=F64)
    out = {"a": a, "b": b, "ard_gamma": ard}
    out["se"] = KernelFunction(kernel_name="squared_exponential", gamma=0.5, kappa=1.5, d=3).kernel(a, b)
    out["se_sym"] = KernelFunction(kernel_name="squared_exponential", gamma=0.5, kappa=1.5, d=3).kernel(a, a)
    out["se_group"] = KernelFunction(kernel_name="squared_exponential", gamma=0.8, d=3, group=[0, 2]).kernel(a, b)
    out["ard"] = KernelFunction(kernel_name="ard", ard_gamma=ard, kappa=0.8, d=3).kernel(a, b)
    out["ard_additive"] = KernelFunction(kernel_name="ard", ard_gamma=ard, d=3, groups=[[0], [1, 2]]).kernel(a, b)
    for nu, tag in ((0.5, "12"), (1.5, "32"), (2.5, "52")):
        out["matern" + tag] = KernelFunction(kernel_name="matern", gamma=0.9, nu=nu, kappa=1.2, d=3).kernel(a, b)
        out["matern%s_sym" % tag] = KernelFunction(kernel_name="matern", gamma=0.9, nu=nu, d=3).kernel(a, a)
        out["ard_matern" + tag] = KernelFunction(kernel_name="ard_matern", ard_gamma=ard, nu=nu, d=3).kernel(a, b)
        out["ard_matern%s_sym" % tag] = KernelFunction(kernel_name="ard_matern", ard_gamma=ard, nu=nu, d=3).kernel(a, a)
    out["poly2"] = KernelFunction(kernel_name="polynomial", power=2, kappa=0.5, d=3).kernel(a, b)
    out["poly3"] = KernelFunction(kernel_name="polynomial", power=3, d=3).kernel(a, b)
    out["linear"] = KernelFunction(kernel_name="linear", kappa=2.0, offset=0.3, d=3).kernel(a, b)
    k_sum = KernelFunction(kernel_name="ard", ard_gamma=ard, d=3) + KernelFunction(kernel_name="polynomial", power=2, d=3)
    out["sum_ard_poly"] = k_sum.kernel(a, b)
    k_mul = KernelFunction(kernel_name="squared_exponential", gamma=0.6, d=3) * \
        KernelFunction(kernel_name="matern", gamma=1.1, nu=2.5, d=3)
    out["mul_se_matern"] = k_mul.kernel(a, b)
    k3 = KernelFunction(kernel_name="squared_exponential", gamma=0.6, d=3) + \
        KernelFunction(kernel_name="linear", d=3)
    k3 = k3 * KernelFunction(kernel_name="ard", ard_gamma=ard, d=3)
    out["fold3"] = k3.kernel(a, b)
    # per-index override, as log_marginal passes it (kernels.py:138-151)
    out["se_override"] = KernelFunction(kernel_name="squared_exponential", gamma=0.5, d=3).kernel(
        a, b, **{'0': {'gamma': 0.9}})
    save("gram", **out)


def gp_case(name, kernel, n, d, nt, s, seed, override=None, full_n=0):
    x, y = data(n, d, seed=seed)
    xt, _ = data(nt, d, seed=seed + 1)
    gp = GaussianProcess(kernel=kernel, s=s)
    gp.fit_gp(x, y)
    mu, std = gp.mean_std(xt)
    out = {"x": x, "y": y, "xt": xt, "s": s, "A": gp.A, "mu": mu, "std": std,
           "lml": gp.log_marginal(kernel, {}, 1.0), "lml_w": gp.log_marginal(kernel, {}, 0.5),
           "lml_chol": Estimator.log_marginal(gp, kernel, {}, 1.0)}
    if override is not None:
        out["lml_override"] = gp.log_marginal(kernel, override, 1.0)
    if full_n:
        mu_f, cov = gp.mean_std(xt[:full_n], full=True)
        out["cov"] = cov
    save(name, **out)


def grad_case():
    n, d = 200, 4
    x, y = data(n, d, seed=30)
    ard0 = torch.tensor([0.8, 1.1, 1.4, 0.9], dtype=F64)
    kernel = KernelFunction(kernel_name="ard", ard_gamma=ard0.clone(), d=d)
    gp = GaussianProcess(kernel=kernel, s=0.1)
    gp.fit_gp(x, y)
    g = torch.tensor([0.6, 1.2, 1.0, 1.5], dtype=F64, requires_grad=True)
    kap = torch.tensor(1.3, dtype=F64, requires_grad=True)
    val = gp.log_marginal(kernel, {'0': {'ard_gamma': g, 'kappa': kap}}, 1.0)
    val.backward()
    out = {"x": x, "y": y, "s": 0.1, "ard_eval": g.detach(), "kappa_eval": kap.detach(), "lml": val.detach(),
           "grad_ard": g.grad, "grad_kappa": kap.grad}
    # isotropic SE: gradient w.r.t. gamma, with weight 0.7
    k2 = KernelFunction(kernel_name="squared_exponential", gamma=0.5, d=d)
    gp2 = GaussianProcess(kernel=k2, s=0.2)
    gp2.fit_gp(x, y)
    gam = torch.tensor(0.75, dtype=F64, requires_grad=True)
    v2 = gp2.log_marginal(k2, {'0': {'gamma': gam}}, 0.7)
    v2.backward()
    out.update({"se_gamma_eval": gam.detach(), "se_lml": v2.detach(), "se_grad_gamma": gam.grad, "se_s": 0.2,
                "se_weight": 0.7})
    save("gp_grad", **out)


def grad_composite_case():
    """Evidence value + autograd gradients of composite kernels through the unmodified reference."""
    sys.path.insert(0, HERE)
    import grad_specs
    out = {}
    for name, c in grad_specs.cases().items():
        x, y = data(c["n"], c["d"], seed=c["seed"])
        kernel = c["build"](KernelFunction)
        s = torch.tensor(c["s"], dtype=F64, requires_grad=True) if c.get("noise_grad") else c["s"]
        gp = GaussianProcess(kernel=kernel, s=s)
        gp.fit_gp(x, y)
        ov = c["override"]()
        lv = grad_specs.leaves(ov)
        for _, _, t in lv:
            t.requires_grad_(True)
        val = gp.log_marginal(kernel, ov, c["weight"])
        val.backward()
        out[name + "__lml"] = val.detach()
        for idx, pname, t in lv:
            out["%s__grad__%s__%s" % (name, idx, pname)] = t.grad
        if c.get("noise_grad"):
            out[name + "__grad_s"] = s.grad
    save("gp_grad_composite", **out)


def mkl_case():
    """MultipleKernelLearner with FIXED weights: the Gram stack (mkl_estimator.py:35-37), the combined Gram (:90)
    and mean / mean_std against it (:103-121, 165-173).  The weight program itself is a cvxpy / MOSEK call that
    cannot run here, so the reference object is put in the state its fit_gp leaves behind and the rest of its
    own code does the work."""
    sys.modules["stpy.regularization.regularizer"] = MagicMock()
    sys.modules["stpy.regularization.simplex_regularizer"] = MagicMock()
    from stpy.continuous_processes.mkl_estimator import MultipleKernelLearner
    n, d, nt = 180, 3, 25
    x, y = data(n, d, seed=90)
    xt, _ = data(nt, d, seed=91)
    kernels = [KernelFunction(kernel_name="squared_exponential", gamma=0.4, d=d),
               KernelFunction(kernel_name="squared_exponential", gamma=1.1, kappa=0.7, d=d),
               KernelFunction(kernel_name="matern", gamma=0.8, nu=2.5, d=d),
               KernelFunction(kernel_name="polynomial", power=2, kappa=0.2, d=d)]
    alphas = torch.tensor([0.15, 0.0, 0.6, 0.25], dtype=F64)
    mkl = MultipleKernelLearner(kernels, lam=1.5, s=0.2)
    mkl.x, mkl.y = x, y
    mkl.n, mkl.d = x.size()
    mkl.Ks = [k.kernel(x, x) for k in kernels]
    mkl.alphas = alphas
    mkl.K = torch.sum(torch.stack([a * K for a, K in zip(mkl.alphas, mkl.Ks)]), dim=0) + np.eye(n) * mkl.lam * mkl.s ** 2
    mkl.fitted = True
    mu, sd = mkl.mean_std(xt)
    K_star, K_ss = mkl.execute(xt)
    save("mkl", x=x, y=y, xt=xt, alphas=alphas, Ks=torch.stack(mkl.Ks), K=mkl.K, mean=mu, std=sd, K_star=K_star,
         K_star_star=K_ss, lam=1.5, s=0.2)


def matern_nu_case():
    """General-nu Matern (the Bessel-function branch of matern_kernel, kernels.py:852-859): Gram matrices with
    exact-zero distances on the diagonal, and one GP fit / prediction / evidence."""
    a, _ = data(90, 3, seed=97)
    b, _ = data(55, 3, seed=98)
    out = {"a": a, "b": b}
    for nu in (0.8, 3.3, 1.0):
        k = KernelFunction(kernel_name="matern", gamma=0.9, nu=nu, kappa=1.3, d=3)
        out["K_ab_%s" % nu] = k.kernel(a, b)
        out["K_aa_%s" % nu] = k.kernel(a, a)
    x, y = data(220, 3, seed=99)
    xt, _ = data(30, 3, seed=100)
    k = KernelFunction(kernel_name="matern", gamma=1.1, nu=1.8, d=3)
    gp = GaussianProcess(kernel=k, s=0.1)
    gp.fit_gp(x, y)
    mu, sd = gp.mean_std(xt)
    out.update({"x": x, "y": y, "xt": xt, "mean": mu, "std": sd, "A": gp.A, "lml": gp.log_marginal(k, {}, 1.0)})
    save("matern_nu", **out)


def nystrom_case():
    """NystromFeatures.fit_gp / embed (nystrom_fea.py:106-207): landmark subsample (seeded np.random.choice) +
    eigh of its Gram, and the `svd` variant (top-m eigenpairs of the full Gram).  Its mean_std uses the removed
    torch.solve, so only fit_gp / embed / outer_kernel are exercised."""
    from stpy.continuous_processes.nystrom_fea import NystromFeatures
    x, y = data(150, 2, seed=110)
    xt, _ = data(40, 2, seed=111)
    out = {"x": x, "y": y, "xt": xt}
    k = KernelFunction(kernel_name="ard_matern", ard_gamma=torch.tensor([0.7, 0.9], dtype=F64), nu=1.5, d=2)
    np.random.seed(7)
    ny = NystromFeatures(k, m=24, approx="uniform", s=0.3)
    ny.fit_gp(x, y)
    out.update({"uni_C": np.asarray(ny.C), "uni_phi_x": ny.embed(x), "uni_phi_t": ny.embed(xt), "uni_Z": ny.Z_})
    k2 = KernelFunction(kernel_name="matern", gamma=0.45, nu=0.5, d=2)  # scipy distances: exact zeros on the diagonal
    ns = NystromFeatures(k2, m=60, approx="svd", s=0.2)
    ns.fit_gp(x, y)
    out.update({"svd_eigs": ns.eigs, "svd_phi_t": ns.embed(xt), "svd_Z": ns.Z_, "svd_outer": ns.outer_kernel()})
    save("nystrom", **out)


def rff_case():
    n, d, m, nt = 160, 4, 64, 48
    x, y = data(n, d, seed=40)
    xt, _ = data(nt, d, seed=41)
    np.random.seed(7)
    emb = RFFEmbedding(gamma=0.8, m=m, d=d, kappa=1.3, kernel="squared_exponential", approx="rff")
    phi = emb.embed(x)
    kf = KernelizedFeatures(embedding=emb, m=m, s=0.1, lam=1.0, d=d)
    kf.fit_gp(x, y)
    mu, std = kf.mean_std(xt)
    theta = kf.theta_mean()
    np.random.seed(8)
    embb = RFFEmbedding(gamma=0.8, m=m, d=d, biased=True, kernel="squared_exponential", approx="rff")
    phib = embb.embed(x)  # reference quirk: (m, n)
    save("rff", x=x, y=y, xt=xt, W=emb.W, phi=phi, theta=theta, mu=mu, std=std, s=0.1, lam=1.0, kappa=1.3,
         Wb=embb.W, bb=embb.b, phib=phib, gamma=0.8)


def qff_case():
    """Quadrature Fourier features (the reference's tests/kernelized-features-test.py recipe)."""
    n, d, m, nt = 120, 2, 64, 40
    x, y = data(n, d, seed=50)
    xt, _ = data(nt, d, seed=51)
    emb = HermiteEmbedding(gamma=0.5, m=m, d=d, kappa=1.2)
    phi = emb.embed(x)
    kf = KernelizedFeatures(embedding=emb, m=emb.get_m(), s=0.1, lam=1.0, d=d)
    kf.fit_gp(x, y)
    mu, std = kf.mean_std(xt)
    embq = QuadratureEmbedding(gamma=0.7, m=32, d=d)
    save("qff", x=x, y=y, xt=xt, W=emb.W, weights=emb.weights, phi=phi, mu=mu, std=std, m=emb.get_m(),
         Wq=embq.W, weightsq=embq.weights, phiq=embq.embed(x), mq=embq.get_m())


def groups_case():
    """Additive per-group kernels (kernels.py:618-698): one lengthscale (vector) per column group."""
    a, _ = data(40, 3, seed=80)
    b, _ = data(17, 3, seed=81)
    groups = [[0], [1, 2]]
    gpg = torch.tensor([0.5, 0.9], dtype=F64)
    apg = torch.tensor([0.5, 0.9, 1.3], dtype=F64)
    k1 = KernelFunction(kernel_name="squared_exponential_per_group", groups=groups, d=3, kappa=1.3,
                        params={'gamma_per_group': gpg})
    k2 = KernelFunction(kernel_name="ard_per_group", groups=groups, d=3, kappa=1.3, params={'ard_per_group': apg})
    save("gram_groups", a=a, b=b, gamma_per_group=gpg, ard_per_group=apg, kappa=1.3,
         se_per_group=k1.kernel(a, b), se_per_group_sym=k1.kernel(a, a),
         ard_per_group_k=k2.kernel(a, b), ard_per_group_sym=k2.kernel(a, a))


def sequential_case():
    """add_data_point (gauss_procc.py:100-111): a fit, then points appended one by one and in a batch."""
    n, d, nt = 120, 2, 32
    x, y = data(n + 13, d, seed=60)
    xt, _ = data(nt, d, seed=61)
    kernel = KernelFunction(kernel_name="matern", gamma=0.8, nu=2.5, d=d)
    gp = GaussianProcess(kernel=kernel, s=0.1)
    gp.fit_gp(x[:n], y[:n])
    gp.add_data_point(x[n:n + 1], y[n:n + 1])
    mu1, std1 = gp.mean_std(xt)
    gp.add_data_point(x[n + 1:n + 2], y[n + 1:n + 2])
    gp.add_data_point(x[n + 2:], y[n + 2:])  # 11 points at once, crossing row 128
    mu, std = gp.mean_std(xt)
    save("gp_sequential", x=x, y=y, xt=xt, s=0.1, n0=n, mu1=mu1, std1=std1, A=gp.A, mu=mu, std=std,
         lml=gp.log_marginal(kernel, {}, 1.0))


def mixture_case():
    """CategoricalMixture.fit_gp / mean_std (categorical_mixture.py:36-83): evidence-weighted model average."""
    from stpy.continuous_processes.categorical_mixture import CategoricalMixture
    n, d, nt = 150, 2, 24
    x, y = data(n, d, seed=70)
    xt, _ = data(nt, d, seed=71)
    gps = [GaussianProcess(kernel=KernelFunction(kernel_name="squared_exponential", gamma=0.4, d=d), s=0.1),
           GaussianProcess(kernel=KernelFunction(kernel_name="squared_exponential", gamma=0.9, d=d), s=0.1),
           GaussianProcess(kernel=KernelFunction(kernel_name="matern", gamma=0.7, nu=2.5, d=d), s=0.1),
           GaussianProcess(kernel=KernelFunction(kernel_name="matern", gamma=1.5, nu=1.5, d=d), s=0.1),
           GaussianProcess(kernel=KernelFunction(kernel_name="linear", kappa=1.0, d=d), s=0.1)]
    mix = CategoricalMixture(gps, d=d)
    mix.fit_gp(x, y)
    logp = torch.tensor([mix.log_prob_normal(g.get_kernel(), y) for g in gps], dtype=F64)
    mu, std = mix.mean_std(xt)
    save("mixture", x=x, y=y, xt=xt, s=0.1, weights=mix.weights, logprobs=logp, mu=mu, std=std)


def main():
    torch.manual_seed(0)
    if len(sys.argv) > 1:  # regenerate only the named cases, e.g. `make_golden.py sequential mixture`
        for name in sys.argv[1:]:
            globals()[name + "_case"]()
        return
    gram_cases()
    gp_case("gp_se_small", KernelFunction(kernel_name="squared_exponential", gamma=0.5, kappa=1., d=2),
            n=300, d=2, nt=64, s=0.1, seed=20, override={'0': {'gamma': 0.7}}, full_n=16)
    gp_case("gp_c1", KernelFunction(kernel_name="squared_exponential", gamma=0.5, kappa=1., d=2),
            n=1024, d=2, nt=256, s=0.1, seed=0)
    gp_case("gp_ard", KernelFunction(kernel_name="ard", ard_gamma=torch.tensor([0.8, 1.0, 1.2, 1.6], dtype=F64), d=4),
            n=260, d=4, nt=40, s=0.1, seed=21,
            override={'0': {'ard_gamma': torch.tensor([1.0, 0.9, 1.5, 1.1], dtype=F64)}})
    gp_case("gp_matern52", KernelFunction(kernel_name="matern", gamma=1.0, nu=2.5, d=3), n=257, d=3, nt=33, s=0.1,
            seed=22)
    gp_case("gp_ard_matern32", KernelFunction(kernel_name="ard_matern", ard_gamma=torch.ones(3, dtype=F64), nu=1.5,
                                              d=3), n=200, d=3, nt=30, s=0.1, seed=23)
    k_sum = KernelFunction(kernel_name="ard", ard_gamma=torch.tensor([0.9, 1.2], dtype=F64), d=2) + \
        KernelFunction(kernel_name="polynomial", power=2, kappa=0.1, d=2)
    gp_case("gp_sum", k_sum, n=150, d=2, nt=20, s=0.2, seed=24)
    grad_case()
    grad_composite_case()
    mkl_case()
    matern_nu_case()
    nystrom_case()
    rff_case()
    qff_case()
    groups_case()
    sequential_case()
    mixture_case()


if __name__ == "__main__":
    main()
