"""C-ABI surface and host logic that need no GPU."""
import os
import re

import pytest
import torch

from conftest import ROOT
from stpy_b200 import _lib


def _declared():
    text = open(os.path.join(ROOT, "include", "stpyb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(stpyb_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 19
    lib = _lib.load()  # resolves every entry of SIGNATURES or raises
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.SIGNATURES, "ctypes signature missing for " + n
    assert sorted(_lib.SIGNATURES) == names
    assert lib.stpyb_version() >= 100


def test_signature_arity_matches_header():
    text = open(os.path.join(ROOT, "include", "stpyb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, args in re.findall(r"\bint\s+(stpyb_\w+)\s*\(([^)]*)\)", text):
        args = args.strip()
        count = 0 if args in ("", "void") else len(args.split(","))
        assert count == len(_lib.SIGNATURES[name]), name


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from stpy_b200.kernels import KernelFunction
    from stpy_b200.continuous_processes.gauss_procc import GaussianProcess
    x = torch.rand(8, 2, dtype=torch.float64)
    with pytest.raises(_lib.StpybError):
        KernelFunction(d=2).kernel(x, x)
    with pytest.raises(_lib.StpybError):
        GaussianProcess(d=2).fit_gp(x, x[:, :1])


def test_kernel_algebra_and_param_protocol():
    from stpy_b200.kernels import KernelFunction
    k1 = KernelFunction(kernel_name="ard", ard_gamma=torch.tensor([1., 2.], dtype=torch.float64), d=2)
    k2 = KernelFunction(kernel_name="polynomial", power=3, d=2, group=[0, 1])
    k3 = KernelFunction(kernel_name="squared_exponential", gamma=0.3, d=3, group=[0, 1, 2])
    k = k1 + k2
    assert k is k1 and k.operations == ["-", "+"] and k.kernel_items == 2
    assert k.optkernel_list == ["ard", "polynomial"]
    assert set(k.params_dict.keys()) == {"0", "1"}
    assert k.params_dict["1"]["degree"] == 3 and k.params_dict["0"]["group"] == [0, 1]
    k = k * k3
    assert k.operations == ["-", "+", "*"] and k.kernel_items == 3
    assert "kernel: polynomial" in k.description() and "gamma=0.3" in k.description()
    # override protocol of log_marginal: only overridden keys + group are passed (kernels.py:105-110)
    over = k.add_groups({"0": {"ard_gamma": torch.tensor([3., 4.])}})
    assert over["1"] == {"group": [0, 1]} and over["2"]["group"] == [0, 1, 2]
    items = k._owners[0]._items(over["0"])
    assert len(items) == 1 and items[0].scale == [1 / 3., 1 / 4.] and items[0].arg_scale == -0.5
    se = k._owners[2]._items({})
    assert abs(se[0].arg_scale + 0.5 / 0.09) < 1e-15 and se[0].kind == _lib.K_SE
    add = KernelFunction(kernel_name="ard", d=3, groups=[[0], [1, 2]])._items({})
    assert [it.cols for it in add] == [[0], [1, 2]] and add[0].kappa == 0.5
    with pytest.raises(AssertionError):
        KernelFunction(kernel_name="gibbs", d=1)
    gen = KernelFunction(kernel_name="matern", nu=0.7, d=1)._items({})  # general nu: the Bessel-function kind
    assert gen[0].kind == _lib.K_MATERN_NU and gen[0].kparams[0] == 0.7 and len(gen[0].kparams) == 6
    with pytest.raises(NotImplementedError):
        KernelFunction(kernel_name="ard_matern", nu=0.7, d=1)._items({})


def test_rff_sampler_uses_numpy_global_rng():
    import numpy as np
    from stpy_b200.embeddings.embedding import RFFEmbedding
    np.random.seed(3)
    e = RFFEmbedding(gamma=0.5, m=8, d=2, biased=True)
    np.random.seed(3)
    W = np.random.normal(size=(8, 2)) * (1. / 0.5)
    b = 2. * np.pi * np.random.uniform(size=(8))
    assert np.array_equal(e.W.numpy(), W) and np.array_equal(e.b.numpy(), b)
    with pytest.raises(AssertionError):
        RFFEmbedding(m=7, d=2)


def test_factor_key_is_the_resolved_kernel_not_the_parameter_tree():
    """An override that spells out the fitted values resolves to the same Gram launches -> same factor key."""
    from stpy_b200.continuous_processes.gauss_procc import GaussianProcess
    from stpy_b200.kernels import KernelFunction
    ard = torch.tensor([1., 2.], dtype=torch.float64)
    k = KernelFunction(kernel_name="ard", ard_gamma=ard.clone(), d=2)
    gp = GaussianProcess(kernel=k, s=0.1)
    base = gp._key(k, k.params_dict, 0.1)
    same = k.add_groups({'0': {'ard_gamma': ard.clone().requires_grad_(True)}})
    assert gp._key(k, same, 0.1) == base
    assert gp._key(k, k.add_groups({'0': {'ard_gamma': torch.tensor([1., 2.5], dtype=torch.float64)}}), 0.1) != base
    assert gp._key(k, k.params_dict, 0.2) != base
    gp._data_version += 1
    assert gp._key(k, k.params_dict, 0.1) != base


def test_snapshot_detects_parameter_change():
    from stpy_b200.continuous_processes.gauss_procc import _snapshot
    from stpy_b200.kernels import KernelFunction
    k = KernelFunction(kernel_name="ard", ard_gamma=torch.tensor([1., 2.], dtype=torch.float64), d=2)
    a = _snapshot(k.params_dict)
    k.params_dict["0"]["ard_gamma"][1] = 2.5
    assert _snapshot(k.params_dict) != a


def test_quadrature_nodes_and_weights_match_reference_fixture():
    """Host-side part of the quadrature embeddings: tensor grid, ordering and weights."""
    from conftest import load_golden
    from stpy_b200.embeddings.embedding import HermiteEmbedding, QuadratureEmbedding
    g = load_golden("qff")
    e = HermiteEmbedding(gamma=0.5, m=64, d=2, kappa=1.2)
    assert e.get_m() == int(g["m"]) == 50
    assert torch.equal(e.W, g["W"]) and torch.equal(e.weights, g["weights"])
    q = QuadratureEmbedding(gamma=0.7, m=32, d=2)
    assert q.get_m() == int(g["mq"])
    assert torch.allclose(q.W, g["Wq"], rtol=0, atol=0) and torch.allclose(q.weights, g["weightsq"], rtol=1e-15, atol=0)


def test_matern_nu_constants_match_their_definitions():
    """Host constants of the general-nu Matern map against scipy's Gamma functions, including the mu -> 0 limit
    where gam1 = (1/Gamma(1-mu) - 1/Gamma(1+mu)) / (2 mu) cancels."""
    import math
    from scipy.special import digamma, rgamma
    from stpy_b200.kernels import matern_nu_constants
    for nu in (0.8, 1.2, 3.3, 0.25, 7.45):
        c = matern_nu_constants(nu)
        mu = nu - int(nu + 0.5)
        assert abs(mu) <= 0.5 and c[0] == nu
        assert abs(c[1] - (rgamma(1 - mu) - rgamma(1 + mu)) / (2 * mu)) < 1e-13
        assert abs(c[2] - 0.5 * (rgamma(1 - mu) + rgamma(1 + mu))) < 1e-15
        assert abs(c[3] - rgamma(1 + mu)) < 1e-15 and abs(c[4] - rgamma(1 - mu)) < 1e-15
        assert abs(c[5] - 2.0 ** (1 - nu) / math.gamma(nu)) < 1e-15 * max(1.0, c[5])
    assert abs(matern_nu_constants(2.0)[1] - digamma(1.0)) < 1e-15          # the limit is -Euler's constant
    assert abs(matern_nu_constants(2.0 + 1e-9)[1] - digamma(1.0)) < 1e-9    # and is approached smoothly


def _emulated_pass(items, sub_ops, x, Wfull, q, off):
    """torch restatement of ONE stpyb_kernel_grad launch in mode 0 (test infrastructure): the fold with its
    pre / suf bookkeeping, per-item values and d value / d sq, summed against W over all (i, j)."""
    import math
    from stpy_b200 import _lib
    vals, dvals, us = [], [], []
    for it in items:
        cols = it["cols"]
        sc = torch.tensor(it["sc"], dtype=torch.float64)
        xs = x[:, cols]
        if it["kind"] <= _lib.K_MATERN52:
            u = (xs[:, None, :] - xs[None, :, :]) * sc
            s = (u * u).sum(-1)
            if it["kind"] == _lib.K_SE:
                v = torch.exp(it["arg_scale"] * s)
                h = it["arg_scale"] * v
            else:
                r = torch.sqrt(s)
                c = {_lib.K_MATERN12: 1.0, _lib.K_MATERN32: math.sqrt(3.0), _lib.K_MATERN52: math.sqrt(5.0)}[it["kind"]]
                t = c * r
                e = torch.exp(-t)
                if it["kind"] == _lib.K_MATERN12:
                    v, h = e, torch.where(r > 0, -0.5 * e / r.clamp_min(1e-300), torch.zeros_like(r))
                elif it["kind"] == _lib.K_MATERN32:
                    v, h = (1 + t) * e, -1.5 * e
                else:
                    v, h = (1 + t + t * t / 3.0) * e, -(5.0 / 6.0) * (1 + t) * e
        else:
            u = None
            s = xs @ xs.T
            v = (s + 1.0) ** it["p0"] if it["kind"] == _lib.K_POLY else s
            h = torch.zeros_like(s)
        vals.append(it["kappa"] * v + (it["p0"] if it["kind"] == _lib.K_LINEAR else 0.0))
        dvals.append((v, it["kappa"] * h))
        us.append(u)
    nsub = len(sub_ops)
    G = [sum(vals[i] for i, it in enumerate(items) if it["sub"] == p) for p in range(nsub)]
    p_star = items[q]["sub"]
    out_prev, pre = None, torch.ones_like(G[0])
    for p in range(nsub):
        if p == p_star:
            pre = out_prev if (p > 0 and sub_ops[p] == _lib.OP_MUL) else torch.ones_like(G[0])
        out_prev = G[p] if p == 0 else (out_prev * G[p] if sub_ops[p] == _lib.OP_MUL else out_prev + G[p])
    suf = torch.ones_like(G[0])
    for p in range(p_star + 1, nsub):
        if sub_ops[p] == _lib.OP_MUL:
            suf = suf * G[p]
    cw = 0.5 * Wfull * pre * suf  # 0.5 tr(W dK): the kernel halves the diagonal and counts the mirror
    out = torch.zeros(18, dtype=torch.float64)
    v, kh = dvals[q]
    out[16] = (cw * v).sum()
    if us[q] is not None:
        for k in range(min(16, len(items[q]["cols"]) - off)):
            out[k] = (cw * kh * us[q][:, :, off + k] ** 2).sum()
    out[17] = 0.5 * torch.diagonal(Wfull).sum()
    return out


def test_derivative_plan_and_host_assembly_reproduce_reference_gradients():
    """Host half of the evidence gradient (kernels.grad_plan + autodiff._assemble: which parameter entry every
    column's lengthscale comes from, the -2/l and d kappa factors, the fold bookkeeping) against the reference's
    autograd fixture, with the device pass replaced by a torch restatement of its definition."""
    import os
    import sys
    from conftest import ROOT, load_golden
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import grad_specs
    from oracle import stpy_oracle as O
    from stpy_b200 import autodiff
    from stpy_b200.kernels import KernelFunction
    g = load_golden("gp_grad_composite")
    for name, c in grad_specs.cases().items():
        x, y = O.make_data(c["n"], c["d"], seed=c["seed"])
        kernel = c["build"](KernelFunction)
        ov = c["override"]()
        lv = grad_specs.leaves(ov)
        for _, _, t in lv:
            t.requires_grad_(True)
        params = kernel.add_groups(dict(ov))
        items, sub_ops = kernel.grad_plan(params)
        autodiff._descriptor(items, sub_ops)  # sizes within the device limits
        # the oracle builds K at the evaluation point; W = w K^-1 - alpha alpha^T; the emulation supplies the sums
        s = c["s"]
        fl = lambda t: float(t.detach())
        Kfun = {
            "ard_matern52": lambda: O.ard_matern_kernel(x, x, ov['0']['ard_gamma'].detach(), nu=2.5, kappa=fl(ov['0']['kappa'])),
            "ard_matern32": lambda: O.ard_matern_kernel(x, x, ov['0']['ard_gamma'].detach(), nu=1.5),
            "sum_ard_ard": lambda: O.ard_kernel(x, x, ov['0']['ard_gamma'].detach(), group=[0, 1])
            + O.ard_kernel(x, x, ov['1']['ard_gamma'].detach(), kappa=fl(ov['1']['kappa']), group=[2, 3]),
            "prod_se_ardmatern": lambda: O.se_kernel(x, x, gamma=fl(ov['0']['gamma']))
            * O.ard_matern_kernel(x, x, ov['1']['ard_gamma'].detach(), nu=2.5),
            "additive_groups": lambda: O.ard_kernel_additive(x, x, ov['0']['ard_gamma'].detach(), [[0, 1], [2, 3]],
                                                             kappa=fl(ov['0']['kappa'])),
            "sum_ard_poly": lambda: O.ard_kernel(x, x, ov['0']['ard_gamma'].detach())
            + O.polynomial_kernel(x, x, degree=2, kappa=fl(ov['1']['kappa'])),
            "fold3_noise": lambda: (O.se_kernel(x, x, gamma=fl(ov['0']['gamma']))
                                    + O.ard_kernel(x, x, ov['1']['ard_gamma'].detach(), kappa=fl(ov['1']['kappa'])))
            * O.polynomial_kernel(x, x, degree=2, kappa=fl(ov['2']['kappa'])),
        }[name]
        Kn = Kfun() + s * s * torch.eye(c["n"], dtype=torch.float64)
        Kinv = torch.linalg.inv(Kn)
        alpha = Kinv @ y
        W = c["weight"] * Kinv - alpha @ alpha.T
        passes = []
        for q, it in enumerate(items):
            want_ls = it["ls_idx"] is not None and autodiff._wanted(it["ls_src"])
            if want_ls:
                passes += [(q, off) for off in range(0, len(it["cols"]), 16)]
            elif autodiff._wanted(it["kappa_src"]):
                passes.append((q, 0))
        host = torch.stack([_emulated_pass(items, sub_ops, x, W, q, off) for q, off in passes])
        inputs, grads = autodiff._assemble(items, passes, host)
        by_id = {id(t): gr for t, gr in zip(inputs, grads)}
        for idx, pname, t in lv:
            ref = torch.as_tensor(g["%s__grad__%s__%s" % (name, idx, pname)], dtype=torch.float64).reshape(t.shape)
            got = by_id[id(t)]
            assert float((got - ref).abs().max() / ref.abs().max()) < 1e-8, (name, idx, pname, got, ref)
        if c.get("noise_grad"):
            assert abs(2.0 * s * float(host[0, 17]) - float(g[name + "__grad_s"])) < 1e-8 * abs(float(g[name + "__grad_s"]))


def test_bessel_k_algorithm_matches_scipy():
    """The K_nu evaluation of csrc/gram.cu::bessel_k (Temme series for x < 2, Steed's continued fraction above,
    upward recurrence), restated line by line in Python with the host constants the device receives, against
    scipy.special.kv -- the function the reference calls (kernels.py:858)."""
    import math
    import numpy as np
    from scipy.special import kv
    from stpy_b200.kernels import matern_nu_constants

    def bessel_k(kp, x):
        nu = kp[0]
        nl = int(nu + 0.5)
        mu = nu - nl
        mu2, xi, xi2, EPS = mu * mu, 1.0 / x, 2.0 / x, 1e-16
        if x < 2.0:
            x2, pimu = 0.5 * x, math.pi * mu
            fact = 1.0 if abs(pimu) < EPS else pimu / math.sin(pimu)
            d = -math.log(x2)
            e = mu * d
            fact2 = 1.0 if abs(e) < EPS else math.sinh(e) / e
            ff = fact * (kp[1] * math.cosh(e) + kp[2] * fact2 * d)
            s, e = ff, math.exp(e)
            p, q, c, d, s1 = 0.5 * e / kp[3], 0.5 / (e * kp[4]), 1.0, x2 * x2, 0.5 * e / kp[3]
            for i in range(1, 500):
                ff = (i * ff + p + q) / (i * i - mu2)
                c *= d / i
                p /= (i - mu)
                q /= (i + mu)
                de = c * ff
                s += de
                s1 += c * (p - i * ff)
                if abs(de) < abs(s) * EPS:
                    break
            rkmu, rk1 = s, s1 * xi2
        else:
            b = 2.0 * (1.0 + x)
            d = 1.0 / b
            h = delh = d
            q1, q2, a1 = 0.0, 1.0, 0.25 - mu2
            q = c = a1
            a = -a1
            s = 1.0 + q * delh
            for i in range(2, 500):
                a -= 2 * (i - 1)
                c = -a * c / i
                qnew = (q1 - b * q2) / a
                q1, q2 = q2, qnew
                q += c * qnew
                b += 2.0
                d = 1.0 / (b + a * d)
                delh = (b * d - 1.0) * delh
                h += delh
                dels = q * delh
                s += dels
                if abs(dels / s) < EPS:
                    break
            h = a1 * h
            rkmu = math.sqrt(math.pi / (2.0 * x)) * math.exp(-x) / s
            rk1 = rkmu * (mu + x + 0.5 - h) * xi
        for i in range(1, nl + 1):
            rkmu, rk1 = rk1, (mu + i) * xi2 * rk1 + rkmu
        return rkmu

    worst = 0.0
    for nu in (0.25, 0.5, 0.8, 1.0, 1.8, 2.0, 3.3, 7.45):
        kp = matern_nu_constants(nu)
        for x in list(np.logspace(-15, 2.5, 200)) + [1.999999, 2.0, 2.000001]:
            ref = float(kv(nu, x))
            if ref == 0.0 or not math.isfinite(ref):
                continue
            worst = max(worst, abs(bessel_k(kp, float(x)) - ref) / abs(ref))
    assert worst < 2e-13, worst


def test_jacobi_round_robin_visits_every_pair_once_per_sweep():
    """Index logic of csrc/eig.cu::jacobi_round_kernel (the circle method): in round r, CTA 0 pairs (np-1, r mod
    np-1) and CTA i pairs ((r+i) mod np-1, (r-i) mod np-1); the np/2 pairs of a round are disjoint and the np-1
    rounds of a sweep cover all np (np-1) / 2 pairs."""
    for np_ in (2, 4, 6, 38, 514):
        m = np_ - 1
        seen = set()
        for r in range(np_ - 1):
            used = set()
            for i in range(np_ // 2):
                p, q = (m, r % m) if i == 0 else ((r + i) % m, (r - i + m) % m)
                assert p != q and p not in used and q not in used
                used.update((p, q))
                seen.add((min(p, q), max(p, q)))
            assert len(used) == np_
        assert len(seen) == np_ * (np_ - 1) // 2


def test_gram_tile_enumeration_covers_the_lower_part_once():
    """Index logic of csrc/gram.cu (launch_gram + the decode at the top of gram_tile_kernel), restated: a lower-only
    launch over an m x n block enumerates tile rows ti < tri_rows with ti + 1 tiles each (closed-form triangular
    decode), then full rows; every 64 x 64 tile that contains an element on or below the diagonal appears once."""
    import math
    for m, n in ((64, 64), (65, 65), (700, 700), (5000, 128), (130, 1000), (1, 1), (8192, 8192)):
        T = 64
        tiles_m, tiles_n = -(-m // T), -(-n // T)
        tri_rows = min(tiles_m, tiles_n)
        tri_count = tri_rows * (tri_rows + 1) // 2
        grid = tri_count + (tiles_m - tri_rows) * tiles_n
        seen = set()
        for bid in range(grid if grid <= 20000 else 0):
            if bid < tri_count:
                r = int((math.sqrt(8.0 * bid + 1.0) - 1.0) * 0.5)
                while (r + 1) * (r + 2) // 2 <= bid:
                    r += 1
                while r * (r + 1) // 2 > bid:
                    r -= 1
                ti, tj = r, bid - r * (r + 1) // 2
            else:
                l = bid - tri_count
                ti, tj = tri_rows + l // tiles_n, l % tiles_n
            assert (ti, tj) not in seen and ti < tiles_m and tj < tiles_n
            seen.add((ti, tj))
        if grid <= 20000:
            need = {(i // T, j // T) for i in range(0, m) for j in (min(i, n - 1),)}  # the diagonal-most element of each row
            need |= {(ti, tj) for ti in range(tiles_m) for tj in range(min(ti + 1, tiles_n))}
            assert need <= seen and len(seen) == grid
        else:  # large case: only the counts
            assert grid == tri_count + (tiles_m - tri_rows) * tiles_n and tri_count == 128 * 129 // 2
