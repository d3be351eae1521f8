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
    """Random Fourier features for the squared exponential kernel (embedding.py:136-241).

    Host side: only the two draws matter for reproducing a seeded reference run -- the (m, d) standard
    normal block scaled by 1/gamma, then (biased variant) m uniform phases -- both from numpy's GLOBAL
    generator, in that order.  The reference's other samplers (laplace via an inverse CDF, orthogonal
    random features, halton, rejection sampling for the modified Matern) are outside the B200 path; two of
    them call helpers that do not exist in the reference itself."""

    def __init__(self, biased=False, **kwargs):
        super().__init__(**kwargs)
        self.biased = biased
        self.sample()

    def sampler(self, size):
        """Spectral sample of the kernel: frequencies of shape `size` (numpy array)."""
        if self.kernel != "squared_exponential" or self.approx != "rff":
            raise NotImplementedError("RFF frequencies are drawn for kernel='squared_exponential', approx='rff' "
                                      "only (got kernel=%r, approx=%r)" % (self.kernel, self.approx))
        inv_lengthscale = 1. / self.gamma
        return np.random.normal(size=size) * inv_lengthscale

    def sample(self):
        self.W = torch.from_numpy(self.sampler((self.m, self.d)))
        self._Wp = None  # device copy of W is rebuilt on the next embed
        if self.biased:
            phases = np.random.uniform(size=self.m)
            self.b = torch.from_numpy(2. * np.pi * phases)

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
    """Quadrature Fourier features on a tensor grid (embedding.py:248-466): the kernel's spectral integral
    is replaced by a 1-d quadrature rule per input dimension, and the d-dimensional rule is the tensor product.
    Nodes and weights are host-side numpy (a few hundred numbers); `embed` is the same device kernel as the
    RFF embedding with the square-rooted weights as per-feature factors (stpyb_rff_embed)."""

    def __init__(self, scale=1.0, **kwargs):
        Embedding.__init__(self, **kwargs)
        self.scale = scale
        self.compute()

    def spectral_density(self, omega):
        """Density the 1-d rule integrates against, for frequencies omega of shape (k, 1): squared exponential
        only (embedding.py:396-405, including its pi/2 normalisation)."""
        if self.kernel != "squared_exponential":
            raise NotImplementedError("spectral density of '%s' is not on the B200 path" % self.kernel)
        sq = np.sum(omega ** 2, axis=1).reshape(-1, 1)
        gauss = np.exp(-sq / 2 * (self.gamma ** 2))
        return gauss * (self.gamma / np.sqrt(2 * np.pi)) * (np.pi / 2)

    def transform(self):
        """The reference exposes the density as a callable (embedding.py:396)."""
        return self.spectral_density

    def nodesAndWeights(self, q):
        """q-point rule on the half line [0, inf): the upper half of a 2q-point Gauss-Legendre rule on
        (-1, 1), pushed through omega = scale * cot(angle), angle in (pi/2, pi), with the Jacobian
        1 / sin^2 and the symmetric half folded in by doubling the weights (embedding.py:423-448)."""
        nodes, gl_w = np.polynomial.legendre.leggauss(2 * q)
        angle = ((nodes[q:] + 1.) / 2.) * np.pi
        jacobian = 1. / (np.sin(angle) ** 2)
        freq = self.scale / np.tan(angle)
        density = self.spectral_density(freq.reshape(-1, 1)).flatten()
        return freq, self.scale * jacobian * (2 * gl_w[q:]) * density

    def compute(self, complexity_reorder=True):
        """Per-dimension rule -> tensor grid (embedding.py:364-394).  m is rounded down to the largest full
        grid: q points per dimension with q^d <= m/2 (cos and sin share a node) or q^d <= m (cosine only)."""
        budget = self.m if self.cosine else self.m // 2
        self.q = int(np.power(budget, 1. / self.d))
        nodes, weights = self.nodesAndWeights(self.q)
        if complexity_reorder:  # low frequencies first
            order = np.argsort(np.abs(nodes))
            nodes, weights = nodes[order], weights[order]
        self.W = torch.from_numpy(_cartesian([nodes] * self.d))
        self.weights = torch.from_numpy(np.prod(_cartesian([weights] * self.d), axis=1))
        self.m = self.q ** self.d * (1 if self.cosine else 2)
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
    """Gauss-Hermite quadrature Fourier features (embedding.py:573-602).  For the squared exponential kernel the
    spectral density IS the Hermite weight function, so the rule needs no density factor: with x_i, w_i the
    upper half of the 2q-point Gauss-Hermite rule, omega_i = sqrt(2) x_i / gamma and weight 2 w_i / sqrt(pi)."""

    def __init__(self, ones=False, cosine=False, **kwargs):
        self.ones = ones
        QuadratureEmbedding.__init__(self, **dict(kwargs, cosine=cosine))
        if self.kernel != "squared_exponential":
            raise AssertionError("Hermite Embedding is allowed only with Squared Exponential Kernel")

    def nodesAndWeights(self, q):
        x, w = np.polynomial.hermite.hermgauss(2 * q)
        upper = slice(q, None)
        weights = np.ones(q) if self.ones else 2 * w[upper]
        return np.sqrt(2) * x[upper] / self.gamma, weights / np.sqrt(np.pi)
