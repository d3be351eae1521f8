"""NystromFeatures: mirror of stpy/continuous_processes/nystrom_fea.py on the device path.

The Nystrom map represents the kernel through m landmark points x_s and the eigendecomposition of their
(weighted) Gram matrix K_s = V diag(D) V^T:  phi(q) = D^-1/2 V^T w k(x_s, q), so that phi(q) . phi(q') is the
rank-m approximation of k(q, q') (nystrom_fea.py:188-196; the `svd` variant, :116-136, takes the top m
eigenpairs of the full n x n Gram).  On this path

  * Gram matrices come from the fused device kernel (KernelFunction.gram_into);
  * torch.linalg.eigh / torch.lobpcg are replaced by the hand-written one-sided Jacobi eigensolver of
    libstpyb (csrc/eig.cu: stpyb_jacobi_init / _sweep / _eigenvalues), which computes the small eigenvalues of
    a positive semi-definite Gram matrix to high relative accuracy -- what the D^-1/2 scaling needs;
  * embed is one rectangular Gram + one DMMA product with the m x |x_s| map M^T = D^-1/2 V^T w;
  * fit_gp forms Z = Phi^T Phi + s^2 I with the same contraction kernel.

Landmarks are drawn with numpy's global generator exactly as the reference does (np.random.choice, :48-52),
so a seeded run picks the same points.  The leverage-score samplers of the reference construct
GaussianProcess(kernel_custom=...), an argument that class does not have, and its mean_std calls the removed
torch.solve; here the samplers raise NotImplementedError and mean_std is the
working Cholesky form of the same formulas (:209-221).  `positive_svd` (sklearn NMF) and `cover`
(scipy sqrtm) are outside the hot path.
"""
import numpy as np
import torch

from .. import _lib as L
from ..embeddings.embedding import Embedding


def eigh_device(A_dev, tol=None, max_sweeps=30):
    """Eigenvalues (ascending) and eigenvectors (columns) of a symmetric matrix on the device, like
    torch.linalg.eigh: one-sided Jacobi sweeps until a whole sweep rotates nothing.  tol: pairs with
    |H_p . H_q| <= tol |H_p| |H_q| count as orthogonal (default 2 n eps: below that the dot products are
    rounding noise and further sweeps only let H drift away from W A)."""
    n = int(A_dev.shape[0])
    if tol is None:
        tol = max(1e-15, 2.0 * n * 2.220446049250313e-16)
    np_ = n + (n & 1)
    ld = L.pad_ld(np_)
    A, lda = L.empty_matrix(n, n)  # aligned rows for the contraction kernel
    A.copy_(A_dev)
    H = torch.empty((np_, ld), dtype=torch.float64, device=A.device)
    W = torch.empty((np_, ld), dtype=torch.float64, device=A.device)
    L.call("stpyb_jacobi_init", L.ptr(A), lda, L.ptr(H), L.ptr(W), n, np_, ld, L.stream_ptr())
    rotated = torch.zeros(1, dtype=torch.int32, device=A.device)
    if np_ >= 2:
        for _ in range(max_sweeps):
            L.call("stpyb_jacobi_sweep", L.ptr(H), L.ptr(W), np_, ld, float(tol), L.ptr(rotated), L.stream_ptr())
            if int(rotated.item()) == 0:
                break
    # eigenvalues as Rayleigh quotients against a freshly formed H = W A (the rotated H carries the rounding of
    # every sweep; W stays orthogonal to working precision)
    L.call("stpyb_gemm_nt", np_, n, n, L.ptr(W), ld, L.ptr(A), lda, L.ptr(H), ld, 1.0, 0.0, 0, L.stream_ptr())
    if np_ != n:
        H[:, n:np_].zero_()
    lam = torch.empty(np_, dtype=torch.float64, device=A.device)
    L.call("stpyb_jacobi_eigenvalues", L.ptr(H), L.ptr(W), np_, ld, L.ptr(lam), L.stream_ptr())
    vec = W[:, :n]
    if np_ != n:  # drop the eigenpair of the padding row (its eigenvector is the padding axis itself)
        keep = torch.argsort(W[:, n].abs())[:n]
        lam, vec = lam[keep], vec[keep]
    order = torch.argsort(lam)
    return lam[order], vec[order].t().contiguous()


class NystromFeatures(Embedding):

    def __init__(self, kernel_object, m=100, approx="uniform", s=1., samples=100):
        self.fit = False
        self.m = m
        try:
            self.ms = int(torch.sum(m))
        except Exception:
            self.ms = m
        self.samples = samples
        self.kernel_object = kernel_object
        self.kernel = kernel_object.kernel
        self.approx = approx
        self.s = s

    def description(self):
        return "Nystrom\n" + "Appprox: " + self.approx

    def get_m(self):
        return self.ms

    def uniform_subsampling(self, x, y):
        N = x.size()[0]
        C = np.random.choice(N, int(self.ms))
        return C, torch.ones(self.ms)

    def subsample(self, x, y):
        if self.approx == "uniform":
            return self.uniform_subsampling(x, y)
        raise NotImplementedError("approx='%s': the reference's leverage-score samplers cannot run (they pass "
                                  "kernel_custom= to GaussianProcess, which has no such argument)" % self.approx)

    # ------------------------------------------------------------------ device pieces
    def _gram(self, a_dev, b_dev, symmetric=False):
        out, ld = L.empty_matrix(b_dev.shape[0], a_dev.shape[0])
        self.kernel_object.gram_into(a_dev, b_dev, self.kernel_object.params_dict, out, ld, symmetric=symmetric)
        return out, ld

    def _set_map(self, xs_dev, Mt):
        """Mt = M^T (m x |x_s|), padded for the contraction kernel."""
        self._xs_dev = xs_dev
        buf, ld = L.empty_matrix(Mt.shape[0], Mt.shape[1])
        buf.copy_(Mt)
        self._Mt, self._ldm = buf, ld
        self.M = buf.t().cpu()

    def embed_device(self, q_dev, transposed=False):
        """Phi (|q| x m) = k(q, x_s)^T M, or Phi^T (m x |q|); returns (view, ld)."""
        kt, ldk = self._gram(self._xs_dev, q_dev)  # (|q| x |x_s|): row j = k(q_j, x_s)
        nq, ns, m = q_dev.shape[0], self._xs_dev.shape[0], self._Mt.shape[0]
        if transposed:
            out, ld = L.empty_matrix(m, nq)
            L.call("stpyb_gemm_nt", m, nq, ns, L.ptr(self._Mt), self._ldm, L.ptr(kt), ldk, L.ptr(out), ld, 1.0, 0.0, 0,
                   L.stream_ptr())
        else:
            out, ld = L.empty_matrix(nq, m)
            L.call("stpyb_gemm_nt", nq, m, ns, L.ptr(kt), ldk, L.ptr(self._Mt), self._ldm, L.ptr(out), ld, 1.0, 0.0, 0,
                   L.stream_ptr())
        return out, ld

    def embed(self, q):
        out, _ = self.embed_device(L.to_device(q))
        return out if (torch.is_tensor(q) and q.is_cuda) else out.cpu()

    # ------------------------------------------------------------------ fit
    def fit_gp(self, x, y, eps=1e-14):
        self.x, self.y = x, y
        self.d, self.N = int(x.size()[1]), int(x.size()[0])
        assert (self.ms <= self.N)
        x_dev = L.to_device(x)
        if self.approx == "svd":  # top-m eigenpairs of the full Gram (nystrom_fea.py:116-136)
            K, _ = self._gram(x_dev, x_dev, symmetric=True)
            D, V = eigh_device(K)
            D, V = D[self.N - self.ms:], V[:, self.N - self.ms:]
            D = torch.where(D <= eps, torch.zeros_like(D), D)
            self.eigs = D.cpu()
            self._set_map(x_dev, (V * torch.sqrt(1. / D)).t())
            self.xs, self.C = x, []
        elif self.approx == "nothing":
            self.xs = x[0:self.ms, :]
            self._set_map(x_dev[:self.ms], torch.eye(self.ms, dtype=torch.float64, device=x_dev.device))
        elif self.approx in ("positive_svd", "cover"):
            raise NotImplementedError("approx='%s' (sklearn NMF / scipy sqrtm) is outside the B200 hot path" % self.approx)
        else:  # landmark subsample + eigendecomposition of its weighted Gram (nystrom_fea.py:185-196)
            self.C, self.weights = self.subsample(x, y)
            xs_dev = x_dev[torch.as_tensor(np.asarray(self.C), device=x_dev.device)]
            w = L.to_device(self.weights)
            self.Dweights = torch.diag(self.weights).double()
            Ks, _ = self._gram(xs_dev, xs_dev, symmetric=True)
            D, V = eigh_device((w.view(-1, 1) * Ks) * w.view(1, -1))
            Dinv = 1. / D
            Dinv = torch.sqrt(torch.where(Dinv <= 0, torch.zeros_like(Dinv), Dinv))
            self.eigs = D.cpu()
            self.xs = x[self.C, :]
            self._set_map(xs_dev, (Dinv.view(-1, 1) * V.t()) * w.view(1, -1))
        phit, ldp = self.embed_device(x_dev, transposed=True)  # Q = Phi^T (m x N)
        Z, ldz = L.empty_matrix(self.ms, self.ms)
        L.call("stpyb_gemm_nt", self.ms, self.ms, self.N, L.ptr(phit), ldp, L.ptr(phit), ldp, L.ptr(Z), ldz, 1.0, 0.0, 0,
               L.stream_ptr())
        Z.diagonal().add_(float(self.s) ** 2)
        self._Z, self._ldz, self._Qt, self._ldq = Z, ldz, phit, ldp
        self.Z_ = Z.cpu()
        self.K = self.Z_
        self.Q = phit.cpu()
        self.fit = True
        return None

    def _theta_factor(self):
        """Cholesky of Z and theta = Z^-1 Q y (nystrom_fea.py:214-216 with a solve that exists)."""
        m = self.ms
        Lz, ld = L.empty_matrix(m, m)
        Lz.copy_(self._Z)
        nblk = (m + L.DB - 1) // L.DB
        dinv = torch.empty((nblk, L.DB, L.DB), dtype=torch.float64, device=Lz.device)
        info = torch.zeros(1, dtype=torch.int32, device=Lz.device)
        L.call("stpyb_potrf", L.ptr(Lz), m, ld, L.ptr(dinv), L.ptr(info), 512, L.stream_ptr())
        rhs = torch.empty(m, dtype=torch.float64, device=Lz.device)
        L.call("stpyb_gemv_rows", L.ptr(self._Qt), m, self.N, self._ldq, L.ptr(L.to_device(self.y).reshape(-1)), L.ptr(rhs),
               L.stream_ptr())
        L.call("stpyb_potrs_vec", L.ptr(Lz), m, ld, L.ptr(dinv), L.ptr(rhs), L.stream_ptr())
        if int(info.item()) != 0:
            raise torch.linalg.LinAlgError("Nystrom: Phi^T Phi + s^2 I is not positive-definite")
        return Lz, ld, dinv, rhs

    def mean_std(self, xtest):
        if self.fit == False:
            raise AssertionError("First fit")
        Lz, ld, dinv, theta = self._theta_factor()
        phi, ldp = self.embed_device(L.to_device(xtest))
        nt = phi.shape[0]
        mean = torch.empty(nt, dtype=torch.float64, device=phi.device)
        L.call("stpyb_gemv_rows", L.ptr(phi), nt, self.ms, ldp, L.ptr(theta), L.ptr(mean), L.stream_ptr())
        L.call("stpyb_trsm_rt", L.ptr(Lz), self.ms, ld, L.ptr(dinv), L.ptr(phi), nt, ldp, L.stream_ptr())
        q = torch.empty(nt, dtype=torch.float64, device=phi.device)
        L.call("stpyb_row_sumsq", L.ptr(phi), nt, self.ms, ldp, None, 0, L.ptr(q), L.stream_ptr())
        std = torch.sqrt(float(self.s) ** 2 * q)
        on_dev = torch.is_tensor(xtest) and xtest.is_cuda
        return (mean.view(-1, 1), std.view(-1, 1)) if on_dev else (mean.view(-1, 1).cpu(), std.view(-1, 1).cpu())

    def outer_kernel(self):
        """Phi Phi^T + s^2 I, the rank-m approximation of the regularised Gram (nystrom_fea.py:223-233)."""
        out, ld = L.empty_matrix(self.N, self.N)
        phi, ldp = self.embed_device(L.to_device(self.x))
        L.call("stpyb_gemm_nt", self.N, self.N, self.ms, L.ptr(phi), ldp, L.ptr(phi), ldp, L.ptr(out), ld, 1.0, 0.0, 0,
               L.stream_ptr())
        out.diagonal().add_(float(self.s) ** 2)
        return out.cpu()
