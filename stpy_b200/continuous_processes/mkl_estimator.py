"""MultipleKernelLearner: mirror of stpy/continuous_processes/mkl_estimator.py on the device path.

The reference builds one Gram matrix per kernel in a Python loop (mkl_estimator.py:35-37), finds simplex weights
alpha by handing  min_alpha y^T (sum_q alpha_q K_q + lam s^2 I)^-1 y  to cvxpy / MOSEK (`opt='closed'`,
mkl_estimator.py:60-64), combines K = sum_q alpha_q K_q + lam s^2 I (:90) and predicts with lstsq / solve against
that K (:108-121, 165-173).  Here

  * the Gram STACK is one device buffer: a single stpyb_gram_multi pass (shared squared-distance tiles, one
    epilogue per kernel) when every member is an isotropic squared-exponential / Matern kernel on the same
    columns, otherwise one fused Gram launch per member; lower triangles only;
  * the weight program is solved with its exact gradient d/d alpha_q = -beta^T K_q beta, beta = K(alpha)^-1 y:
    every evaluation is stpyb_stack_combine -> stpyb_potrf -> two stpyb_trsv -> stpyb_stack_quadform on the
    device; the k-variable simplex problem itself (k <= 64) is driven by scipy's SLSQP on the host.  cvxpy and
    MOSEK are not dependencies; the reference's regularisers (cvxpy expressions) are out of scope;
  * the fitted model IS a GaussianProcess with the composite kernel sum_q alpha_q k_q and noise s sqrt(lam), so
    mean / mean_std / log_marginal / sample are the inherited device paths (std_fixed, :165-173, is the same
    formula as GaussianProcess.mean_std).
"""
import copy

import numpy as np
import torch

from .. import _lib as L
from .. import sweep
from ..kernels import KernelFunction, _f  # noqa: F401
from .gauss_procc import GaussianProcess


def _scaled_copy(k, factor):
    """A single (non-composite) kernel with its amplitude multiplied by `factor`; the original is untouched."""
    if len(k._owners) != 1:
        raise NotImplementedError("MultipleKernelLearner members must be single kernels (no +/* composites)")
    c = copy.copy(k)
    c.kappa = float(factor) * _f(k.params_dict['0'].get('kappa', k.kappa))
    c.params = dict(k.params_dict['0'], kappa=c.kappa)
    c._owners = [c]
    c.kernel_function_list = [c._single_kernel]
    c.kernel_diag_function_list = [c._single_kernel_diag]
    c.optkernel_list = [c.optkernel]
    c.params_dict = {'0': c.params}
    c.kernel_items = 1
    c.operations = ["-"]
    return c


class MultipleKernelLearner(GaussianProcess):

    def __init__(self, kernel_objects, lam=1.0, s=0.01, opt='closed', regularizer=None):
        if regularizer is not None:
            raise NotImplementedError("cvxpy regularisers are outside the B200 hot path (SURVEY.md section 2, row 12)")
        if opt != 'closed':
            raise NotImplementedError("opt='%s' (an SDP for MOSEK) is outside the B200 hot path" % opt)
        d = max(int(k.d) for k in kernel_objects)
        GaussianProcess.__init__(self, kernel=kernel_objects[0], s=s, d=d)
        self.kernel_objects = kernel_objects
        self.no_models = len(kernel_objects)
        self.regularizer = regularizer
        self.s_member = s
        self.lam = lam
        self.opt = opt
        self.var = 'fixed'
        self.alphas = None
        self._stack = None

    # ------------------------------------------------------------------ Gram stack
    def _build_stack(self, x_dev):
        """(k, n, ld) device buffer, lower triangles of k_q(x, x) (no noise)."""
        n, k = x_dev.shape[0], self.no_models
        ld = L.pad_ld(n)
        stack = torch.empty((k, n, ld), dtype=torch.float64, device=x_dev.device)
        try:
            specs = [sweep._isotropic_spec(ko) for ko in self.kernel_objects]
            shared = len({sp[3] for sp in specs}) == 1
        except NotImplementedError:
            specs, shared = None, False
        if shared and k <= 64:
            from ..kernels import _Item, _prep
            xp, nrm, dpad = _prep(x_dev, _Item(L.K_LINEAR, list(specs[0][3])))
            L.call("stpyb_gram_multi", k, L.host_ints([sp[0] for sp in specs]), L.host_doubles([sp[1] for sp in specs]),
                   L.host_doubles([sp[2] for sp in specs]), L.ptr(xp), L.ptr(nrm), n, dpad, 0.0, L.ptr(stack), ld, n * ld,
                   L.stream_ptr())
        else:
            for q, ko in enumerate(self.kernel_objects):
                ko.gram_into(x_dev, x_dev, ko.params_dict, stack[q, :, :n], ld, symmetric=True, lower_only=True)
        return stack, ld

    @property
    def Ks(self):
        """The k Gram matrices as full symmetric (n, n) tensors (mkl_estimator.py:35-37), materialised on access."""
        if self._stack is None:
            return []
        stack, ld = self._stack
        n = stack.shape[1]
        out = []
        for q in range(self.no_models):
            lo = torch.tril(stack[q, :, :n])
            out.append(self._out(lo + torch.tril(lo, -1).t()))
        return out

    # ------------------------------------------------------------------ weights
    def _objective(self, stack, ld, y_dev, alpha):
        """f(alpha) = y^T K(alpha)^-1 y and its gradient -beta^T K_q beta, all on the device."""
        k, n = self.no_models, stack.shape[1]
        w = self._work
        L.call("stpyb_stack_combine", L.ptr(stack), k, L.host_doubles(alpha), n, ld, n * ld,
               float(self.lam) * float(self.s_member) ** 2, 1, L.ptr(w["A"]), w["ld"], L.stream_ptr())
        L.call("stpyb_potrf", L.ptr(w["A"]), n, w["ld"], L.ptr(w["dinv"]), L.ptr(w["info"]), int(self.outer_block),
               L.stream_ptr())
        w["z"].copy_(y_dev)
        L.call("stpyb_trsv", L.ptr(w["A"]), n, w["ld"], L.ptr(w["dinv"]), L.ptr(w["z"]), 0, L.stream_ptr())
        L.call("stpyb_lml", L.ptr(w["A"]), n, w["ld"], L.ptr(w["z"]), 0.0, L.ptr(w["out3"]), L.stream_ptr())
        L.call("stpyb_trsv", L.ptr(w["A"]), n, w["ld"], L.ptr(w["dinv"]), L.ptr(w["z"]), 1, L.stream_ptr())
        L.call("stpyb_stack_quadform", L.ptr(stack), k, n, ld, n * ld, 1, L.ptr(w["z"]), L.ptr(w["g"]), L.stream_ptr())
        host = torch.cat([w["out3"][:1], w["g"], w["info"].double()]).cpu()
        if int(host[-1]) != 0:
            raise torch.linalg.LinAlgError("MKL: the combined Gram matrix is not positive-definite")
        return float(host[0]), -host[1:1 + k].numpy()

    def solve_weights(self, stack, ld, y_dev, maxiter=200):
        """argmin over the simplex of y^T K(alpha)^-1 y (mkl_estimator.py:60-64, convex in alpha)."""
        from scipy.optimize import minimize
        k, n = self.no_models, stack.shape[1]
        A, lda = L.empty_matrix(n, n)
        nblk = (n + L.DB - 1) // L.DB
        dev = stack.device
        self._work = {"A": A, "ld": lda, "dinv": torch.empty((nblk, L.DB, L.DB), dtype=torch.float64, device=dev),
                      "info": torch.zeros(1, dtype=torch.int32, device=dev),
                      "z": torch.empty(n, dtype=torch.float64, device=dev),
                      "out3": torch.empty(3, dtype=torch.float64, device=dev),
                      "g": torch.empty(k, dtype=torch.float64, device=dev)}
        scale = [1.0]

        def fg(a):
            v, g = self._objective(stack, ld, y_dev, np.maximum(a, 0.0))
            return v / scale[0], g / scale[0]
        a0 = np.full(k, 1.0 / k)
        scale[0] = max(abs(fg(a0)[0]), 1e-300)  # SLSQP's tolerances are absolute: normalise the objective
        res = minimize(fg, a0, jac=True, method="SLSQP", bounds=[(0.0, 1.0)] * k,
                       constraints=[{"type": "eq", "fun": lambda a: np.sum(a) - 1.0, "jac": lambda a: np.ones(k)}],
                       options={"maxiter": maxiter, "ftol": 1e-12})
        a = np.clip(res.x, 0.0, 1.0)
        a[a < 1e-10] = 0.0
        self._work = None
        return a / a.sum()

    # ------------------------------------------------------------------ fit / predict
    def fit(self):
        self.fit_gp(self.x, self.y)

    def fit_gp(self, x, y, alphas=None):
        """Gram stack -> simplex weights (or the given ones) -> GP fit with K = sum_q alpha_q K_q + lam s^2 I."""
        x_dev = L.to_device(x)
        y_dev = L.to_device(y).reshape(-1)
        self._stack = None
        self._stack = self._build_stack(x_dev)
        if alphas is None:
            alphas = self.solve_weights(self._stack[0], self._stack[1], y_dev)
        self.alphas = torch.as_tensor(np.asarray(alphas, dtype=np.float64)).reshape(-1)
        combined = None
        for a, ko in zip(self.alphas.tolist(), self.kernel_objects):
            if a == 0.0:
                continue
            term = _scaled_copy(ko, a)
            combined = term if combined is None else combined + term
        self.kernel_object = combined
        self.kernel = combined.kernel
        self.s = float(self.s_member) * float(np.sqrt(self.lam))
        return GaussianProcess.fit_gp(self, x, y)

    def execute(self, xtest):
        """(K*, K**) of the combined kernel (mkl_estimator.py:93-101)."""
        K_star = self.kernel(self.x, xtest) if self.fitted else None
        return K_star, self.kernel(xtest, xtest)

    def std_fixed(self, xtest):
        return self.mean_std(xtest)[1]
