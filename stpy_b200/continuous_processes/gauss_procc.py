"""GaussianProcess: drop-in for stpy/continuous_processes/gauss_procc.py (squared loss).

Same constructor, attributes (x, y, n, d, s, A, K, fitted, kernel_object, kernel,
back_prop, max_size) and methods (fit_gp, fit, add_data_point, load_data,
mean_std, mean, log_marginal, sample, get_kernel, ucb, lcb).  The arithmetic is
ONE fused Gram + ONE blocked Cholesky on the device and triangular solves
against it, instead of the reference's dense Sigma^T Sigma product, second
Gram, n 1x1 kernel calls and pivoted-QR lstsq solves
(gauss_procc.py:151-177, 336-401) or its two LU factorisations per evidence
evaluation (gauss_procc.py:631-638).

What is deliberately different (documented in DESIGN.md):
  * the factor L overwrites the lower triangle of the device Gram buffer; `K`
    is re-materialised on attribute access (it is 34 GB at n = 65 536);
  * fit_gp does not populate `B` (the reference's side effect of calling
    mean_std(x) on the training set, an n-RHS solve whose result it discards);
  * non-squared losses (huber / svr / unif: cvxpy programs) are out of scope.
"""
import numpy as np
import torch

from .. import _lib as L
from ..estimator import Estimator
from ..kernels import KernelFunction
from .. import autodiff


def _snapshot(params_dict):
    """Hashable copy of a kernel parameter tree (values, not references)."""
    out = []
    for key in sorted(params_dict.keys()):
        sub = params_dict[key]
        items = []
        for k in sorted(sub.keys()):
            v = sub[k]
            if torch.is_tensor(v):
                v = tuple(v.detach().reshape(-1).cpu().tolist())
            elif isinstance(v, np.ndarray):
                v = tuple(v.reshape(-1).tolist())
            elif isinstance(v, (list, tuple)):
                v = repr(v)
            elif callable(v):
                v = id(v)
            items.append((k, v))
        out.append((key, tuple(items)))
    return tuple(out)


class _Factor:
    """Device-resident Cholesky factor of K = k(x, x) + s^2 I and what hangs off it."""

    def __init__(self, n, cap=None):
        self.n = n                      # rows in use
        self.cap = max(int(cap or n), n)  # rows allocated (add_data_point appends in place up to cap)
        self.buf, self.ld = L.empty_matrix(self.cap, self.cap)  # lower triangle: L after potrf
        nblk = (self.cap + L.DB - 1) // L.DB
        self.dinv = torch.empty((nblk, L.DB, L.DB), dtype=torch.float64, device=self.buf.device)
        self.info = torch.zeros((1,), dtype=torch.int32, device=self.buf.device)
        self.info_base = 0  # row offset of the last (partial) factorisation, for the error message
        self.z = None       # L^-1 y
        self.out3 = torch.empty((3,), dtype=torch.float64, device=self.buf.device)
        self.key = None

    def check(self):
        info = int(self.info.item())
        if info != 0:
            raise torch.linalg.LinAlgError(
                "linalg.cholesky: The factorization could not be completed because the input is not "
                "positive-definite (the leading minor of order %d is not positive-definite)."
                % (info + self.info_base))


class GaussianProcess(Estimator):

    # device-state defaults (subclasses with their own constructors -- KernelizedFeatures, CategoricalMixture -- share them)
    _A = None
    _fit = None
    _scratch = None
    _x_dev = _y_dev = _x_src = _y_src = _A_dev = None
    _data_version = 0
    Sigma = None
    incremental = True  # add_data_point borders the device factor (O(n^2 k)) instead of refitting (O(n^3))
    outer_block = 1024  # K-depth of the trailing SYRK (stpyb_potrf): best or tied for every n (profiles/outer_block_probe_r01.txt)

    def __init__(self, gamma=1, s=0.001, kappa=1., kernel_name="squared_exponential", diameter=1.0,
                 groups=None, bounds=None, nu=1.5, kernel=None, d=1, power=2, lam=1., loss='squared',
                 huber_delta=1.35, hyper='classical', B=1., svr_eps=0.1):
        self.s = s
        self.d = d
        self.x = None
        self.y = None
        self.n = 0
        # public attributes callers of the reference read or set (gauss_procc.py:36-74), same names and defaults
        self.__dict__.update(mu=0.0, lam=lam, total_bound=B, prob=0.5, svr_eps=svr_eps, safe=False, fitted=False,
                             diameter=diameter, bounds=bounds, admits_first_order=False, back_prop=True, loss=loss,
                             huber_delta=huber_delta, hyper=hyper, prepared_log_marginal=False,
                             warm_start_solution=None, max_size=10000)
        self.Sigma = None
        self._A = None
        if kernel is not None:
            self.kernel_object = kernel
            self.kernel = kernel.kernel
            self.d = kernel.d
        else:
            self.kernel_object = KernelFunction(kernel_name=kernel_name, gamma=gamma, nu=nu, groups=groups,
                                                kappa=kappa, power=power, d=d)
            self.kernel = self.kernel_object.kernel
            self.gamma, self.v, self.groups, self.kappa = gamma, nu, groups, kappa
            self.custom, self.optkernel = kernel, kernel_name
        self._fit = None      # _Factor of the fitted model
        self._scratch = None  # _Factor reused by log_marginal evaluations at other hyper-parameters
        self._x_dev = None
        self._y_dev = None
        self._x_src = None    # the host-side tensors the device copies were made from
        self._y_src = None
        self._A_dev = None
        self._data_version = 0

    # ------------------------------------------------------------------ plumbing
    def __getstate__(self):
        """Pickle / copy: hyper-parameters, data and `fitted` travel; the device factor is a cache that the first
        prediction of the restored object rebuilds (_ensure_factor) -- a pickled reference GP stays fitted too
        (stpy/test_functions/swissfel_simulator.py:18-26 relies on it)."""
        st = dict(self.__dict__)
        for k in ("_fit", "_scratch", "_x_dev", "_y_dev", "_A_dev", "_x_src", "_y_src"):
            st[k] = None
        return st

    @property
    def A(self):
        """alpha = K^-1 y, (n, 1) (gauss_procc.py:376).  Built on demand for a model that is fitted but holds no
        device state yet (unpickled, or scored in a batch by CategoricalMixture)."""
        if self._A is None:
            self._ensure_factor()
        return self._A

    @A.setter
    def A(self, value):
        self._A = value

    def _ensure_factor(self):
        """Rebuild the device state of a model that is `fitted` but holds no factor (unpickled / copied)."""
        if self.fitted and self._fit is None and self.x is not None:
            self.fit_gp(self.x, self.y, Sigma=self.Sigma)

    def _sync_data(self):
        """Upload self.x / self.y if they are not the tensors the device copies were made from (load_data, or a
        plain assignment to gp.x / gp.y, replaces them: estimator.py:28-30); the reference's log_marginal always
        reads self.x and self.y (gauss_procc.py:631-638)."""
        if self.x is None:
            return
        if self._x_dev is None or self._x_src is not self.x or self._y_src is not self.y:
            self.n, self.d = int(self.x.shape[0]), int(self.x.shape[1])
            self._x_dev = L.to_device(self.x)
            self._y_dev = L.to_device(self.y).reshape(-1)
            self._x_src, self._y_src = self.x, self.y
            self._data_version += 1

    def load_data(self, d):
        self.x = d[0]
        self.y = d[1]
        self._sync_data()

    def _out(self, t):
        """Return device results on the device the user's data lives on."""
        if torch.is_tensor(self.x) and self.x.is_cuda:
            return t
        return t.cpu()

    def description(self):
        return self.kernel_object.description() + "\nlambda=" + str(self.s)

    def embed(self, x):
        return self.kernel_object.embed(x)

    def get_basis_size(self):
        return self.kernel_object.get_basis_size()

    def get_kernel(self):
        return self.K

    def residuals(self, x, y):
        return self.mean(x) - y

    @property
    def K(self):
        """K = k(x, x) + Sigma^T Sigma, re-materialised on access (the device buffer holds L)."""
        self._sync_data()
        if self._x_dev is None:
            return np.array([1.0])
        out, ld = L.empty_matrix(self.n, self.n)
        self.kernel_object.gram_into(self._x_dev, self._x_dev, self.kernel_object.params_dict, out, ld,
                                     symmetric=True, diag_add=self._noise_diag())
        self._add_sigma(out, ld, lower_only=False)
        return self._out(out)

    def _noise_diag(self):
        return float(self.s) ** 2 if self.Sigma is None else 0.0

    def _add_sigma(self, out, ld, lower_only):
        if self.Sigma is None:
            return
        St, lds = L.empty_matrix(self.n, self.n)
        St.copy_(L.to_device(self.Sigma).t())
        L.call("stpyb_gemm_nt", self.n, self.n, self.n, L.ptr(St), lds, L.ptr(St), lds, L.ptr(out), ld, 1.0, 1.0,
               int(lower_only), L.stream_ptr())

    # ------------------------------------------------------------------ fit
    def add_data_point(self, x, y, Sigma=None):
        """gauss_procc.py:100-111.  The reference concatenates and refits from scratch; with the default
        noise model (no custom Sigma) and unchanged hyper-parameters the refit equals bordering the
        factor the model already holds, which is what _append does."""
        self._ensure_factor()
        if (self.incremental and self.fitted and self._fit is not None and self.x is not None
                and self.Sigma is None and Sigma is None
                and self._fit.key == self._key(self.kernel_object, self.kernel_object.params_dict, float(self.s))):
            return self._append(x, y)
        if self.x is not None:
            self.x = torch.cat((self.x, x), dim=0)
            self.y = torch.cat((self.y, y), dim=0)
            if self.Sigma is not None:
                extra = torch.eye(x.size()[0], dtype=torch.double) * self.s if Sigma is None else Sigma
                self.Sigma = torch.block_diag(self.Sigma, extra)
            elif Sigma is not None:
                self.Sigma = torch.block_diag(torch.eye(self.n, dtype=torch.double) * self.s, Sigma)
        else:
            self.x = x
            self.y = y
            self.Sigma = Sigma
        self.fit_gp(self.x, self.y, Sigma=self.Sigma)

    def _append(self, x, y):
        """Bordered Cholesky: with K' = [[K, B^T], [B, C]] and K = L L^T,
        L' = [[L, 0], [B L^-T, chol(C - B L^-T L^-1 B^T)]].  The new rows restart at the last
        128-aligned row n0 <= n so that the inverted diagonal blocks the solves consume stay aligned."""
        f = self._fit
        n_old, k = self.n, int(x.shape[0])
        N = n_old + k
        self.x = torch.cat((self.x, x), dim=0)
        self.y = torch.cat((self.y, y), dim=0)
        self._x_dev = torch.cat((self._x_dev, L.to_device(x)), dim=0)
        self._y_dev = torch.cat((self._y_dev, L.to_device(y).reshape(-1)))
        self._x_src, self._y_src = self.x, self.y
        self._data_version += 1
        if N > f.cap:
            g = _Factor(N, cap=((N + max(1024, N // 8) + 1023) // 1024) * 1024)
            g.buf[:n_old, :n_old].copy_(f.buf[:n_old, :n_old])
            nb_old = (n_old + L.DB - 1) // L.DB
            g.dinv[:nb_old].copy_(f.dinv[:nb_old])
            self._fit = None
            self._fit = f = g
        n0 = (n_old // L.DB) * L.DB
        r = N - n0
        pd = self.kernel_object.params_dict
        xr = self._x_dev[n0:N]
        corner = f.buf[n0:N, n0:N]
        self.kernel_object.gram_into(xr, xr, pd, corner, f.ld, symmetric=True, lower_only=True,
                                     diag_add=float(self.s) ** 2)
        if n0 > 0:
            rows = f.buf[n0:N, 0:n0]
            self.kernel_object.gram_into(self._x_dev[:n0], xr, pd, rows, f.ld)
            L.call("stpyb_trsm_rt", L.ptr(f.buf), n0, f.ld, L.ptr(f.dinv), L.ptr(rows), r, f.ld, L.stream_ptr())
            L.call("stpyb_gemm_nt", r, r, n0, L.ptr(rows), f.ld, L.ptr(rows), f.ld, L.ptr(corner), f.ld, -1.0, 1.0,
                   1, L.stream_ptr())
        L.call("stpyb_potrf", L.ptr(corner), r, f.ld, L.ptr(f.dinv[n0 // L.DB:]), L.ptr(f.info),
               int(self.outer_block), L.stream_ptr())
        f.info_base = n0
        f.n = self.n = N
        f.z = self._y_dev.clone()
        L.call("stpyb_trsv", L.ptr(f.buf), N, f.ld, L.ptr(f.dinv), L.ptr(f.z), 0, L.stream_ptr())
        alpha = f.z.clone()
        L.call("stpyb_trsv", L.ptr(f.buf), N, f.ld, L.ptr(f.dinv), L.ptr(alpha), 1, L.stream_ptr())
        f.check()
        f.key = self._key(self.kernel_object, pd, float(self.s))
        self._A_dev = alpha
        self.A = self._out(alpha.view(-1, 1))

    def _key(self, kernel_object, params_dict, s):
        """Cache key of a factor: the data version, the noise level and WHAT the Gram launches would compute --
        the resolved items of every sub-kernel (kind, columns, scales, map constants), not the raw parameter
        tree: an override that spells out the fitted values (the optimiser's first evaluation, a gradient
        request at the current point) is recognised as the fitted factor instead of being refactorised."""
        try:
            resolved = []
            for i, owner in enumerate(kernel_object._owners):
                kw = params_dict[str(i)] if str(i) in params_dict else {}
                if owner.optkernel == "custom":
                    resolved.append(("custom", id(owner.kernel_function), _snapshot({'0': kw})))
                    continue
                resolved.append(tuple((it.kind, tuple(it.cols), tuple(it.scale or ()), it.divide, it.arg_scale, it.kappa,
                                       it.p0, it.refine, tuple(it.kparams or ())) for it in owner._items(kw)))
            what = (tuple(kernel_object.operations), tuple(resolved))
        except Exception:  # unusual parameter trees: fall back to the value snapshot
            what = _snapshot(params_dict)
        return (id(kernel_object), what, float(s), self._data_version)

    def fit(self, x=None, y=None):
        if x is not None:
            self.fit_gp(x, y)
        else:
            self.fit_gp(self.x, self.y)

    def lcb(self, xtest):
        mu, s = self.mean_std(xtest)
        return mu - 2 * s

    def ucb(self, xtest):
        mu, s = self.mean_std(xtest)
        return mu + 2 * s

    def fit_gp(self, x, y, Sigma=None, iterative=False, extrapoint=False):
        """Gram + s^2 I -> Cholesky -> alpha = K^-1 y   (gauss_procc.py:136-177, 367-378)."""
        if self.loss != "squared":
            raise NotImplementedError("only the squared loss is on the B200 path (SURVEY.md section 8a)")
        self.n, self.d = int(x.shape[0]), int(x.shape[1])
        self.Sigma = Sigma
        self.x = x
        self.y = y
        self._x_dev = L.to_device(x)
        self._y_dev = L.to_device(y).reshape(-1)
        self._x_src, self._y_src = x, y
        self._data_version += 1
        if self._fit is None or self._fit.cap < self.n or self._fit.cap > 2 * self.n + 2048:
            self._fit = None
            self._fit = _Factor(self.n)
        f = self._fit
        f.n = self.n
        self._factorize(f, self.kernel_object, self.kernel_object.params_dict, float(self.s))
        alpha = f.z.clone()
        L.call("stpyb_trsv", L.ptr(f.buf), self.n, f.ld, L.ptr(f.dinv), L.ptr(alpha), 1, L.stream_ptr())
        f.check()
        self._A_dev = alpha
        self.A = self._out(alpha.view(-1, 1))
        self.fitted = True
        return None

    def _factorize(self, f, kernel_object, params_dict, s):
        """Lower Gram (+ noise) into f.buf, factor in place, z = L^-1 y."""
        n = self.n
        kernel_object.gram_into(self._x_dev, self._x_dev, params_dict, f.buf, f.ld, symmetric=True,
                                lower_only=True, diag_add=(s * s if self.Sigma is None else 0.0))
        self._add_sigma(f.buf, f.ld, lower_only=True)
        L.call("stpyb_potrf", L.ptr(f.buf), n, f.ld, L.ptr(f.dinv), L.ptr(f.info), int(self.outer_block),
               L.stream_ptr())
        f.z = self._y_dev.clone()
        L.call("stpyb_trsv", L.ptr(f.buf), n, f.ld, L.ptr(f.dinv), L.ptr(f.z), 0, L.stream_ptr())
        f.info_base = 0
        f.key = self._key(kernel_object, params_dict, s)

    # ------------------------------------------------------------------ prediction
    def mean_std(self, xtest, full=False, reuse=False):
        nt = xtest.size()[0]
        if nt < self.max_size:
            return self.mean_std_sub(xtest, full=full, reuse=reuse)
        mus, stds = [], []
        for lo in range(0, nt, self.max_size):
            mu, std = self.mean_std_sub(xtest[lo:lo + self.max_size, :], reuse=True)
            mus.append(mu)
            stds.append(std)
        return torch.cat(mus, dim=0), torch.cat(stds, dim=0)

    def _prior(self, xt_dev, full):
        pd = self.kernel_object.params_dict
        nt = xt_dev.shape[0]
        if full:
            cov, ld = L.empty_matrix(nt, nt)
            self.kernel_object.gram_into(xt_dev, xt_dev, pd, cov, ld, symmetric=True)
            return cov, ld
        return self.kernel_object.diag_device(xt_dev, xt_dev, pd), None

    def mean_std_sub(self, xtest, full=False, reuse=False):
        """Posterior mean and std (or full covariance): gauss_procc.py:336-401."""
        to_user = (lambda t: t) if (torch.is_tensor(xtest) and xtest.is_cuda) else (lambda t: t.cpu())
        xt = L.to_device(xtest)
        nt = xt.shape[0]
        self._ensure_factor()
        if not self.fitted:
            second, _ = self._prior(xt, full)
            zero = torch.zeros((nt, 1), dtype=torch.float64, device=xt.device)
            return to_user(zero), to_user(second if full else torch.sqrt(second.view(-1, 1)))
        f, n = self._fit, self.n
        pd = self.kernel_object.params_dict
        kstar, ldk = L.empty_matrix(nt, n)
        self.kernel_object.gram_into(self._x_dev, xt, pd, kstar, ldk)  # (nt x n), as kernel(self.x, xtest)
        mean = torch.empty((nt,), dtype=torch.float64, device=xt.device)
        L.call("stpyb_gemv_rows", L.ptr(kstar), nt, n, ldk, L.ptr(self._A_dev), L.ptr(mean), L.stream_ptr())
        # V^T = K* L^-T in place; K* K^-1 K*^T = V^T V
        L.call("stpyb_trsm_rt", L.ptr(f.buf), n, f.ld, L.ptr(f.dinv), L.ptr(kstar), nt, ldk, L.stream_ptr())
        if not full:
            kss, _ = self._prior(xt, False)
            std = torch.empty((nt,), dtype=torch.float64, device=xt.device)
            L.call("stpyb_row_sumsq", L.ptr(kstar), nt, n, ldk, L.ptr(kss), 1, L.ptr(std), L.stream_ptr())
            return to_user(mean.view(-1, 1)), to_user(std.view(-1, 1))
        cov, ldc = self._prior(xt, True)
        L.call("stpyb_gemm_nt", nt, nt, n, L.ptr(kstar), ldk, L.ptr(kstar), ldk, L.ptr(cov), ldc, -1.0, 1.0, 0,
               L.stream_ptr())
        return to_user(mean.view(-1, 1)), to_user(cov)

    def mean(self, xtest):
        self._ensure_factor()
        xt = L.to_device(xtest)
        nt, n = xt.shape[0], self.n
        kstar, ldk = L.empty_matrix(nt, n)
        self.kernel_object.gram_into(self._x_dev, xt, self.kernel_object.params_dict, kstar, ldk)
        mean = torch.empty((nt,), dtype=torch.float64, device=xt.device)
        L.call("stpyb_gemv_rows", L.ptr(kstar), nt, n, ldk, L.ptr(self._A_dev), L.ptr(mean), L.stream_ptr())
        out = mean.view(-1, 1)
        return out if (torch.is_tensor(xtest) and xtest.is_cuda) else out.cpu()

    def sample(self, xtest, size=1, jitter=10e-8):
        """Posterior (or prior) path samples: gauss_procc.py:461-482."""
        nn = int(xtest.size()[0])
        on_dev = torch.is_tensor(xtest) and xtest.is_cuda
        if self.fitted:
            ymean, cov = self.mean_std(xtest, full=True)
            eps = 10e-10
        else:
            cov, _ = self._prior(L.to_device(xtest), True)
            ymean, eps = self.mu, jitter
        cov = L.to_device(cov)
        buf, ld = L.empty_matrix(nn, nn)
        buf.copy_(cov)
        buf.diagonal().add_(eps)
        nblk = (nn + L.DB - 1) // L.DB
        dinv = torch.empty((nblk, L.DB, L.DB), dtype=torch.float64, device=buf.device)
        info = torch.zeros((1,), dtype=torch.int32, device=buf.device)
        L.call("stpyb_potrf", L.ptr(buf), nn, ld, L.ptr(dinv), L.ptr(info), 128, L.stream_ptr())
        if int(info.item()) != 0:
            raise torch.linalg.LinAlgError("linalg.cholesky: the covariance is not positive-definite "
                                           "(leading minor of order %d)" % int(info.item()))
        buf.tril_()
        # same host RNG stream as the reference (torch.normal on the CPU generator)
        rv = torch.normal(mean=torch.zeros(nn, size, dtype=torch.float64), std=1.)
        rvt, ldr = L.empty_matrix(size, nn)
        rvt.copy_(rv.t().to(buf.device))
        fout, ldf = L.empty_matrix(nn, size)
        L.call("stpyb_gemm_nt", nn, size, nn, L.ptr(buf), ld, L.ptr(rvt), ldr, L.ptr(fout), ldf, 1.0, 0.0, 0,
               L.stream_ptr())
        fout = fout if on_dev else fout.cpu()
        return ymean + fout

    def sample_and_max(self, xtest, size=1):
        f = self.sample(xtest, size=size)
        self.temp = f
        val, index = torch.max(f, dim=0)
        return (xtest[index, :], val)

    # ------------------------------------------------------------------ evidence
    def log_marginal(self, kernel, X, weight):
        if self.loss != "squared":
            raise NotImplementedError("only the squared loss is on the B200 path")
        return self._log_marginal_squared(kernel, X, weight)

    def _lml_value(self, kernel, X, weight):
        return self._log_marginal_squared(kernel, X, weight)

    def _log_marginal_squared(self, kernel, X, weight):
        """0.5 y^T K^-1 y + 0.5 w logdet K with K = k(x, x; X) + s^2 I  (gauss_procc.py:631-638).

        Returns a (1, 1) float64 tensor.  If a tensor in X (or self.s) requires grad the
        result carries an analytic backward (stpy_b200/autodiff.py)."""
        if self.x is None:
            raise RuntimeError("log_marginal needs data: call fit_gp or load_data first")
        self._sync_data()
        kernel_object = kernel
        if len(X) > 0:
            params_dict = dict(X)
            kernel_object.add_groups(params_dict)
        else:
            params_dict = kernel_object.params_dict
        if autodiff.needs_grad(params_dict, self.s):
            return autodiff.lml_with_grad(self, kernel_object, params_dict, weight)
        f = self._factor_for(kernel_object, params_dict, float(self.s))
        L.call("stpyb_lml", L.ptr(f.buf), self.n, f.ld, L.ptr(f.z), float(weight), L.ptr(f.out3), L.stream_ptr())
        out = f.out3.cpu()  # synchronises: one 24-byte read-back
        f.check()
        val = out[2].view(1, 1)
        return val.to(self._x_dev.device) if (torch.is_tensor(self.x) and self.x.is_cuda) else val

    def optimize_params(self, type='bandwidth', restarts=10, regularizer=None, maxiter=1000, mingradnorm=1e-4,
                        verbose=False, optimizer="pytorch-minimize", scale=1., weight=1., save=False,
                        save_name='model.np', init_func=None, bounds=None, parallel=False, cores=None):
        """Evidence maximisation over lengthscales (and noise): gauss_procc.py:640-702.

        type in {"bandwidth", "bandwidth+noise"}; rotations / covariance manifolds / discrete group
        search need pymanopt and are outside the B200 hot path."""
        if regularizer is not None:
            raise NotImplementedError("spectral-norm / lasso regularisers are outside the B200 hot path")
        if type not in ("bandwidth", "bandwidth+noise"):
            raise AttributeError("This quick-optimization is not implemented.")
        params = {}
        for key, dict2 in self.kernel_object.params_dict.items():
            if 'gamma' in dict2.keys():
                params[key] = {'gamma': (init_func, 1, bounds)}
            elif 'ard_gamma' in dict2.keys():
                params[key] = {'ard_gamma': (init_func, len(dict2['group']), bounds)}
        if type == "bandwidth+noise":
            params['likelihood'] = {'sigma': ((lambda k: np.array([float(self.s)])) if init_func is not None else None,
                                              1, None)}
        return self.optimize_params_general(params=params, restarts=restarts, optimizer=optimizer, maxiter=maxiter,
                                            mingradnorm=mingradnorm, verbose=verbose, scale=scale, weight=weight,
                                            save=save, save_name=save_name, parallel=parallel, cores=cores)

    def _factor_for(self, kernel_object, params_dict, s):
        """The fitted factor if the hyper-parameters are the ones it was built with, else a scratch one."""
        key = self._key(kernel_object, params_dict, s)
        if self._fit is not None and self._fit.key == key:
            return self._fit
        if self._scratch is None or self._scratch.n != self.n:
            self._scratch = None
            self._scratch = _Factor(self.n)
        if self._scratch.key != key:
            self._factorize(self._scratch, kernel_object, params_dict, s)
        return self._scratch
