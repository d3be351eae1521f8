"""KernelizedFeatures: mirror of stpy/continuous_processes/kernelized_features.py (primal path).

Bayesian linear regression over a finite embedding Phi:
    V = Phi^T Phi + s^2 lam I,  theta = V^-1 Phi^T y,
    mean = Phi* theta,          std = s sqrt(diag(Phi* V^-1 Phi*^T))
(kernelized_features.py:236-240, 256, 274-288).  The reference materialises
Q = embed(x) (n x m), forms Q^T Q, takes an SVD pseudo-inverse and re-evaluates
invV @ Q.T @ y (2 m^2 n flops) on every prediction.  Here the normal equations
are accumulated by a chunked embed^T -> DMMA-SYRK stream that never stores Phi
(stpyb_rff_normal_eq), V is factored once by the blocked Cholesky, theta is two
triangular solves, and the predictive std is a TRSM + row norms.
"""
import numpy as np
import torch

from .. import _lib as L
from ..kernels import KernelFunction, _Item, _prep
from .gauss_procc import GaussianProcess


class KernelizedFeatures(GaussianProcess):

    chunk = 16384  # training rows embedded per SYRK pass
    distributed = False  # True: every rank accumulates the normal equations of its row shard, then all-reduce

    def __init__(self, embedding, m, s=0.001, lam=1., d=1, diameter=1.0, theta_norm=1.0, verbose=True, groups=None,
                 bounds=None, scale=1.0, kappa=1.0, poly=2, primal=True, beta_fun=None, bound=1):
        # public attributes of the reference class (kernelized_features.py:17-60), same names and defaults
        self.__dict__.update(s=s, lam=lam, primal=primal, x=None, y=None, mu=0.0, m=torch.from_numpy(np.array(m)),
                             fitted=False, data=False, d=d, n=0, bounds=bounds, groups=groups, diameter=diameter,
                             theta_norm=theta_norm, verbose=verbose, admits_first_order=True, embedding=embedding,
                             embedding_map=embedding, kappa=kappa, scale=scale, poly=poly, to_add=[], prior_mean=0,
                             dual=False, beta_fun=beta_fun, bound=bound, loss="squared")
        self.linear_kernel = KernelFunction(kernel_name="linear").linear_kernel
        self._V = None      # (m+1) x ld device buffer: L of V in the top-left m x m, [Phi^T y] in row m
        self._theta = None

    def __getstate__(self):
        st = dict(self.__dict__)
        for k in ("_V", "_theta", "_dinv", "_Vraw"):
            st[k] = None
        st["fitted"] = False
        return st

    def description(self):
        return "Custom Features object"

    def embed(self, x):
        return self.embedding.embed(x)

    def set_embedding(self, embed):
        self.embedding_map = embed

    def get_basis_size(self):
        return int(torch.sum(self.m))

    def set_basis_size(self, m):
        self.m = m

    def kernel(self, x, y):
        return self.linear_kernel(self.embed(x), self.embed(y))

    def beta(self, delta=0.1, norm=None):
        if self.beta_fun is None:
            return 2.0
        raise NotImplementedError("beta_fun variants are outside the B200 hot path")

    # ------------------------------------------------------------------ fit
    def add_data_point(self, x, y):
        if self.n == 0:
            self.fit_gp(x, y)
        else:
            # the reference applies Woodbury rank-one updates (out of scope); refit from the joined data
            self.fit_gp(torch.cat((self.x, x), dim=0), torch.cat((self.y, y), dim=0))

    def fit(self, x=None, y=None):
        self.fit_gp(self.x, self.y)

    def fit_gp(self, x, y):
        self.x = x
        self.y = y
        self.n = list(self.x.size())[0]
        self.d = list(self.x.size())[1]
        self.dual = False
        if not self.primal and self.n < int(self.get_basis_size()):
            raise NotImplementedError("the dual (n < m) formulation is outside the B200 hot path; use primal=True")
        self.data = True
        self.fitted = False
        self.precompute()
        return None

    def _phi_t_device(self, x_dev):
        """Phi^T (m x n) on the device for any embedding object (the reference's embedding seam)."""
        if hasattr(self.embedding, "embed_device"):
            return self.embedding.embed_device(x_dev, transposed=True)
        phi = L.to_device(self.embedding.embed(x_dev))
        if phi.shape[0] != x_dev.shape[0]:
            phi = phi.t()
        out, ld = L.empty_matrix(phi.shape[1], phi.shape[0])
        out.copy_(phi.t())
        return out, ld

    def precompute(self):
        """V = Phi^T Phi + s^2 lam I, its Cholesky factor and theta (kernelized_features.py:227-242, 256)."""
        if self.fitted or not self.data:
            return
        m = self.get_basis_size()
        x_dev = L.to_device(self.x)
        y_dev = L.to_device(self.y).reshape(-1)
        world, rank = 1, 0
        if self.distributed:
            import torch.distributed as dist
            if dist.is_initialized():
                world, rank = dist.get_world_size(), dist.get_rank()
                # rows shard naturally: V = sum over ranks of Phi_p^T Phi_p
                per = (x_dev.shape[0] + world - 1) // world
                x_dev = x_dev[rank * per:(rank + 1) * per]
                y_dev = y_dev[rank * per:(rank + 1) * per]
        n = x_dev.shape[0]
        Vfull, ldv = L.empty_matrix(m + 1, m + 1, zero=True)
        emb = self.embedding
        if n == 0:
            pass
        elif hasattr(emb, "_spec"):
            wp, bias, featw, mode, scale, dpad = emb._spec(x_dev.shape[1])
            xp, _, _ = _prep(x_dev, _Item(L.K_LINEAR, list(range(x_dev.shape[1]))), want_norms=False)
            chunk = int(min(self.chunk, n))
            scratch, lds = L.empty_matrix(m + 1, chunk)
            L.call("stpyb_rff_normal_eq", L.ptr(xp), L.ptr(y_dev), n, L.ptr(wp), m, dpad, L.ptr(bias), L.ptr(featw),
                   mode, scale, chunk, L.ptr(scratch), lds, L.ptr(Vfull), ldv, L.stream_ptr())
            del scratch
        else:
            for lo in range(0, n, self.chunk):
                hi = min(n, lo + self.chunk)
                pt, _ = self._phi_t_device(x_dev[lo:hi])
                aug, lda = L.empty_matrix(m + 1, hi - lo)
                aug[:m].copy_(pt)
                aug[m].copy_(y_dev[lo:hi])
                L.call("stpyb_gemm_nt", m + 1, m + 1, hi - lo, L.ptr(aug), lda, L.ptr(aug), lda, L.ptr(Vfull), ldv,
                       1.0, 1.0, 1, L.stream_ptr())
        if world > 1:
            import torch.distributed as dist
            # one all-reduce of the (m+1) x ld accumulator (NCCL over NVLink); the m x m factorisation
            # that follows is replicated -- it is 1/(3 n / m) of the SYRK work
            dist.all_reduce(Vfull._base if Vfull._base is not None else Vfull)
        self._Vraw_row = Vfull[m, :m].clone()                      # Phi^T y
        Vfull[:m, :m].diagonal().add_(float(self.s) ** 2 * float(self.lam))
        nblk = (m + L.DB - 1) // L.DB
        self._dinv = torch.empty((nblk, L.DB, L.DB), dtype=torch.float64, device=Vfull.device)
        info = torch.zeros((1,), dtype=torch.int32, device=Vfull.device)
        self._Vsym = None
        self._keep_V = Vfull[:m, :m].clone() if m <= 4096 else None  # for the public V / invV attributes
        L.call("stpyb_potrf", L.ptr(Vfull), m, ldv, L.ptr(self._dinv), L.ptr(info), 256, L.stream_ptr())
        theta = self._Vraw_row.clone()
        L.call("stpyb_potrs_vec", L.ptr(Vfull), m, ldv, L.ptr(self._dinv), L.ptr(theta), L.stream_ptr())
        if int(info.item()) != 0:
            raise torch.linalg.LinAlgError("linalg.cholesky: V = Phi^T Phi + s^2 lam I is not positive-definite "
                                           "(leading minor of order %d)" % int(info.item()))
        self._V, self._ldv, self._theta = Vfull, ldv, theta
        self.fitted = True

    def _user(self, t, like):
        return t if (torch.is_tensor(like) and like.is_cuda) else t.cpu()

    @property
    def Q(self):
        """embed(x): re-materialised on access (65 GB at n = 1e6, m = 8192)."""
        return self.embed(self.x)

    @property
    def V(self):
        if self._keep_V is None:
            raise RuntimeError("V is not kept for m > 4096")
        Vl = torch.tril(self._keep_V)
        return self._user(Vl + torch.tril(Vl, -1).t(), self.x)

    @property
    def invV(self):
        m = self.get_basis_size()
        work, ldw = L.empty_matrix(m, m)
        out, ldo = L.empty_matrix(m, m)
        L.call("stpyb_potri", L.ptr(self._V), m, self._ldv, L.ptr(self._dinv), L.ptr(work), ldw, L.ptr(out), ldo,
               L.stream_ptr())
        lo = torch.tril(out)
        return self._user(lo + torch.tril(lo, -1).t(), self.x)

    def get_invV(self):
        self.precompute()
        return self.invV

    def theta_mean(self, var=False, prior=False):
        self.precompute()
        m = self.get_basis_size()
        if self.fitted and not prior:
            theta = self._user(self._theta.view(-1, 1), self.x)
            Z = (float(self.s) ** 2 * self.invV) if var else None
        else:
            theta = 0 * torch.ones(size=(m, 1)).double()
            Z = None
        return theta if var is False else (theta, Z)

    def mean(self, xtest):
        return self.mean_std(xtest)[0]

    def mean_std(self, xtest):
        """mean = Phi* theta ; std = s sqrt(diag(Phi* V^-1 Phi*^T)) = s ||L^-1 Phi*^T||_col (kernelized_features.py:269-288)."""
        self.precompute()
        m = self.get_basis_size()
        xt = L.to_device(xtest)
        nt = xt.shape[0]
        if hasattr(self.embedding, "embed_device"):
            phi, ldp = self.embedding.embed_device(xt, transposed=False)
        else:
            e = L.to_device(self.embedding.embed(xt))
            phi, ldp = L.empty_matrix(nt, m)
            phi.copy_(e)
        mean = torch.empty((nt,), dtype=torch.float64, device=xt.device)
        L.call("stpyb_gemv_rows", L.ptr(phi), nt, m, ldp, L.ptr(self._theta), L.ptr(mean), L.stream_ptr())
        L.call("stpyb_trsm_rt", L.ptr(self._V), m, self._ldv, L.ptr(self._dinv), L.ptr(phi), nt, ldp, L.stream_ptr())
        ss = torch.empty((nt,), dtype=torch.float64, device=xt.device)
        L.call("stpyb_row_sumsq", L.ptr(phi), nt, m, ldp, None, 0, L.ptr(ss), L.stream_ptr())
        std = torch.sqrt(float(self.s) ** 2 * ss)
        return self._user(mean.view(-1, 1), xtest), self._user(std.view(-1, 1), xtest)

    def sample_theta(self, size=1, prior=False):
        """Posterior (or prior) samples of the weight vector: theta ~ N(theta_mean, s^2 V^-1)
        (kernelized_features.py:319-336).  With V = L L^T the factor L^-T is a square root of V^-1,
        so a sample is theta_mean + s L^-T eps -- one transposed triangular solve per draw instead of
        the reference's Cholesky of the explicit inverse.  Same distribution; the draw for a given
        eps differs because the square root of s^2 V^-1 is a different one.  eps comes from the
        CPU generator, as in the reference."""
        basis = self.get_basis_size()
        rv = torch.normal(mean=torch.zeros(basis, size, dtype=torch.float64), std=1.)
        self.precompute()
        if self.fitted and not prior:
            cols = []
            for j in range(size):
                e = L.to_device(rv[:, j]).clone()
                L.call("stpyb_trsv", L.ptr(self._V), basis, self._ldv, L.ptr(self._dinv), L.ptr(e), 1, L.stream_ptr())
                cols.append(self._theta + float(self.s) * e)
            theta = torch.stack(cols, dim=1)
            return self._user(theta, self.x)
        return float(np.sqrt(self.lam)) * rv + self.prior_mean

    def ucb(self, xtest, delta=0.1):
        mu, std = self.mean_std(xtest)
        return mu + np.sqrt(self.beta(delta=delta)) * std

    def lcb(self, xtest, delta=0.1):
        mu, std = self.mean_std(xtest)
        return mu - np.sqrt(self.beta(delta=delta)) * std

    def log_marginal(self, kernel, X, weight):
        raise NotImplementedError("KernelizedFeatures has no evidence on the B200 path")
