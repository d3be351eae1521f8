"""Embedding / RFFEmbedding: mirror of stpy/embeddings/embedding.py (the RFF part).

Frequencies and phases are drawn on the host with numpy's global RNG exactly as
the reference does (embedding.py:149-223), so a seeded run produces the same W
and b.  `embed` runs on the device: the projection X W^T is a DMMA contraction
and the cos / sin / scale map is applied to the accumulators before the store
(libstpyb: stpyb_rff_embed), replacing the reference's mm + cos + sin + cat +
transpose chain (embedding.py:225-241).
"""
import numpy as np
import torch
from scipy.stats import chi, norm

from .. import _lib as L
from ..kernels import _Item, _prep


class Embedding():
    """Base class: hyper-parameters of a finite-dimensional kernel approximation (embedding.py:53-118)."""

    def __init__(self, gamma=0.1, nu=0.5, m=100, d=1, diameter=1.0, groups=None, kappa=1.0,
                 kernel="squared_exponential", cosine=False, approx="rff", **kwargs):
        # attribute names are the reference's (embedding.py:57-72); `n` aliases nu there too
        self.__dict__.update(gamma=float(gamma), n=nu, nu=nu, m=int(m), d=int(d), kappa=kappa, cosine=cosine,
                             diameter=diameter, groups=groups, kernel=kernel, approx=approx, gradient_avail=0)
        if self.m & 1:
            raise AssertionError("Number of random features has to be even.")

    def sample(self):
        raise AttributeError("Only derived classes can call this method.")

    def embed(self, x):
        raise AttributeError("Only derived classes can call this method.")

    def get_m(self):
        return self.m


class RFFEmbedding(Embedding):
    """Random Fourier features (embedding.py:136-241)."""

    def __init__(self, biased=False, **kwargs):
        super().__init__(**kwargs)
        self.biased = biased
        self.sample()

    def sampler(self, size):
        if self.kernel == "squared_exponential":
            distribution = lambda size: np.random.normal(size=size) * (1. / self.gamma)
            inv_cum_dist = lambda x: norm.ppf(x) * (1. / self.gamma)
        elif self.kernel == "laplace":
            distribution = None
            inv_cum_dist = lambda x: (np.tan(np.pi * x - np.pi) / self.gamma)
        else:
            raise NotImplementedError("RFF sampler for kernel '%s' is not on the B200 path "
                                      "(it is broken in the reference too, SURVEY.md section 2 #15)" % self.kernel)
        if self.approx == "rff":
            if distribution is None:
                self.W = inv_cum_dist(np.random.uniform(size=size))
            else:
                self.W = distribution(size)
        elif self.approx == "orf":
            W0 = np.random.normal(size=size) * (1.)
            self.Q, _ = np.linalg.qr(W0)
            self.S = np.diag(chi.rvs(size[1], size=size[0]))
            self.W = np.dot(self.S, self.Q) / self.gamma ** 2
        else:
            raise NotImplementedError("approx='%s' is not available (helper missing in the reference)" % self.approx)
        return self.W

    def sample(self):
        self.W = self.sampler(size=(self.m, self.d))
        self.W = torch.from_numpy(self.W)
        self._Wp = None
        if self.biased == True:
            self.b = 2. * np.pi * np.random.uniform(size=(self.m))
            self.bs = self.b.reshape(self.m, 1)
            self.b = torch.from_numpy(self.b)
            self.bs = torch.from_numpy(self.bs)

    # ---------------------------------------------------------------- device path
    def _spec(self, d):
        """(Wp, bias_dev, featw_dev, mode, scale, dpad) for inputs with d columns."""
        key = (int(d), self.W.data_ptr())
        if getattr(self, "_Wp", None) is None or self._Wp[0] != key:
            W_dev = L.to_device(self.W)
            wp, _, dpad = _prep(W_dev, _Item(L.K_LINEAR, list(range(d))), want_norms=False)
            bias = L.to_device(self.b) if self.biased else None
            self._Wp = (key, wp, bias, dpad)
        _, wp, bias, dpad = self._Wp
        scale = float(np.sqrt(2. / float(self.m)) * np.sqrt(self.kappa))
        return wp, bias, None, (1 if self.biased else 0), scale, dpad

    def embed_device(self, x_dev, transposed=False):
        """Phi (n x m) -- or Phi^T (m x n) -- as a view of a padded device buffer; returns (view, ld)."""
        n, d = x_dev.shape
        wp, bias, featw, mode, scale, dpad = self._spec(d)
        xp, _, _ = _prep(x_dev, _Item(L.K_LINEAR, list(range(d))), want_norms=False)
        out, ld = L.empty_matrix(self.m, n) if transposed else L.empty_matrix(n, self.m)
        L.call("stpyb_rff_embed", L.ptr(xp), n, L.ptr(wp), self.m, dpad, L.ptr(bias), L.ptr(featw), mode, scale,
               int(transposed), L.ptr(out), ld, L.stream_ptr())
        return out, ld

    def embed(self, x):
        """(n, m) features.  As in the reference, the biased variant comes back transposed, (m, n)
        (embedding.py:232, 241 apply torch.t twice)."""
        x_dev = L.to_device(x)
        out, _ = self.embed_device(x_dev, transposed=bool(self.biased))
        return out if (torch.is_tensor(x) and x.is_cuda) else out.cpu()


def _cartesian(arrays):
    """Cartesian product with the first array varying slowest (the ordering of stpy's helper.cartesian,
    stpy/helpers/helper.py:27-58)."""
    grids = np.meshgrid(*[np.asarray(a) for a in arrays], indexing="ij")
    return np.stack([g.reshape(-1) for g in grids], axis=1)


class QuadratureEmbedding(Embedding):
    """Quadrature Fourier features on a tensor grid (embedding.py:248-466).  Nodes and weights are
    formed on the host exactly as in the reference; `embed` is the same device kernel as the RFF
    embedding, with the square-rooted quadrature weights as per-feature factors (stpyb_rff_embed)."""

    def __init__(self, scale=1.0, **kwargs):
        Embedding.__init__(self, **kwargs)
        self.scale = scale
        self.compute()

    def reorder_complexity(self, omegas, weights):
        order = np.argsort(np.abs(omegas))
        return omegas[order], weights[order]

    def transform(self):
        """Spectral density of the kernel (embedding.py:396-421)."""
        if self.kernel == "squared_exponential":
            return lambda omega: np.exp(-np.sum(omega ** 2, axis=1).reshape(-1, 1) / 2 * (self.gamma ** 2)) * \
                np.power((self.gamma / np.sqrt(2 * np.pi)), 1.) * np.power(np.pi / 2, 1.)
        if self.kernel == "laplace":
            return lambda omega: np.prod(1. / ((self.gamma ** 2) * (omega ** 2) + 1.), axis=1).reshape(-1, 1) * \
                np.power(self.gamma / 2., 1.)
        raise NotImplementedError("spectral density of '%s' is not on the B200 path" % self.kernel)

    def nodesAndWeights(self, q):
        """Gauss-Legendre nodes mapped to the half line through cot (embedding.py:423-448)."""
        # same floating-point operations, in the same order, as the reference: the nodes and weights
        # must come out bit-identical (tests/test_abi.py checks them against the reference's)
        nodes, gl_w = np.polynomial.legendre.leggauss(2 * q)
        angle = ((nodes[q:] + 1.) / 2.) * np.pi            # positive half of the rule -> (pi/2, pi)
        jacobian = 1. / (np.sin(angle) ** 2)               # |d cot / d angle|
        freq = self.scale / np.tan(angle)
        density = self.transform()(freq.reshape(-1, 1)).flatten()
        return freq, self.scale * jacobian * (2 * gl_w[q:]) * density

    def compute(self, complexity_reorder=True):
        """Tensor grid of nodes and product weights (embedding.py:364-394)."""
        if self.cosine == False:
            self.q = int(np.power(self.m // 2, 1. / self.d))
            self.m = self.q ** self.d
        else:
            self.q = int(np.power(self.m, 1. / self.d))
            self.m = self.q ** self.d
        (omegas, weights) = self.nodesAndWeights(self.q)
        if complexity_reorder == True:
            (omegas, weights) = self.reorder_complexity(omegas, weights)
        self.weights = np.prod(_cartesian([weights for _ in range(self.d)]), axis=1)
        self.W = _cartesian([omegas for _ in range(self.d)])
        if self.cosine == False:
            self.m = self.m * 2
        self.W = torch.from_numpy(self.W)
        self.weights = torch.from_numpy(self.weights)
        self._Wp = None

    def _spec(self, d):
        """cos and sin share the frequencies here: stack W twice and let the split-mode epilogue
        (first half cos, second half sin) apply sqrt(weights) per feature."""
        key = (int(d), self.W.data_ptr())
        if getattr(self, "_Wp", None) is None or self._Wp[0] != key:
            W = L.to_device(self.W)
            sw = torch.sqrt(L.to_device(self.weights))
            if not self.cosine:
                W = torch.cat([W, W], dim=0).contiguous()
                sw = torch.cat([sw, sw]).contiguous()
            wp, _, dpad = _prep(W, _Item(L.K_LINEAR, list(range(d))), want_norms=False)
            self._Wp = (key, wp, sw, dpad)
        _, wp, sw, dpad = self._Wp
        return wp, None, sw, (1 if self.cosine else 0), float(np.sqrt(self.kappa)), dpad

    def embed_device(self, x_dev, transposed=False):
        n, d = x_dev.shape
        wp, bias, featw, mode, scale, dpad = self._spec(d)
        xp, _, _ = _prep(x_dev, _Item(L.K_LINEAR, list(range(d))), want_norms=False)
        out, ld = L.empty_matrix(self.m, n) if transposed else L.empty_matrix(n, self.m)
        L.call("stpyb_rff_embed", L.ptr(xp), n, L.ptr(wp), self.m, dpad, L.ptr(bias), L.ptr(featw), mode, scale,
               int(transposed), L.ptr(out), ld, L.stream_ptr())
        return out, ld

    def embed(self, x):
        """(n, m) features sqrt(w_i) cos / sin(omega_i . x) sqrt(kappa) (embedding.py:450-466)."""
        out, _ = self.embed_device(L.to_device(x), transposed=False)
        return out if (torch.is_tensor(x) and x.is_cuda) else out.cpu()


class HermiteEmbedding(QuadratureEmbedding):
    """Gauss-Hermite quadrature Fourier features for the squared exponential kernel (embedding.py:573-602)."""

    def __init__(self, ones=False, cosine=False, **kwargs):
        self.ones = ones
        kwargs = dict(kwargs, cosine=cosine)
        QuadratureEmbedding.__init__(self, **kwargs)
        if self.kernel != "squared_exponential":
            raise AssertionError("Hermite Embedding is allowed only with Squared Exponential Kernel")

    def nodesAndWeights(self, q):
        (nodes, weights) = np.polynomial.hermite.hermgauss(2 * q)
        nodes = nodes[q:]
        weights = 2 * weights[q:]
        if self.ones == True:
            weights = np.ones(q)
        nodes = np.sqrt(2) * nodes / self.gamma
        weights = weights / np.sqrt(np.pi)
        return (nodes, weights)
