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
