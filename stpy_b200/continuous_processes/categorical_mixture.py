"""CategoricalMixture: drop-in for stpy/continuous_processes/categorical_mixture.py.

A finite mixture of Gaussian processes weighted by their evidence.  The reference fits
every member (Gram + n-RHS lstsq), then calls get_kernel() and factorises each K a
second time on the host (scipy LU + numpy slogdet, categorical_mixture.py:36-46).

Here fit_gp is ONE batched pass when the members allow it (plain GaussianProcess members with a
single isotropic squared-exponential / Matern kernel on the same input columns and the same noise
level -- the model-selection shape of the reference's tutorial): the shared squared-distance tiles
feed every member's Gram epilogue (stpyb_gram_multi) and the independent Cholesky factorisations
run on several streams (stpy_b200/sweep.py::lml_sweep).  Only the evidences are needed for the
weights; a member's own factor and alpha are built on demand, the first time a prediction or a
sample asks for them.  Other member types are fitted one by one and their log-probability is read
off the factor the fit already holds (stpyb_lml: no second factorisation).
"""
import math

import numpy as np
import torch

from .. import _lib as L
from .. import sweep
from .gauss_procc import GaussianProcess


class CategoricalMixture(GaussianProcess):

    def __init__(self, processes, init_weights=None, d=1, bounds=None):
        self.k = len(processes)
        if init_weights is None:
            init_weights = torch.ones(size=(self.k, 1)).view(-1).double() * 1. / float(self.k)
        if len(processes) != init_weights.shape[0]:
            raise AssertionError("Not the same number")
        total = torch.sum(init_weights)
        if total > 1.:  # prior weights summing to more than one are normalised, as in the reference
            init_weights = init_weights / total
        self.__dict__.update(processes=processes, bounds=bounds, beta=2., d=d, x=None, y=None,
                             init_weights=init_weights, weights=init_weights)
        self.logprobs = None
        self.fitted = False
        self.batched = True  # score all members in one lml_sweep pass when they allow it

    def add_data_point(self, x, y):
        for model in self.processes:
            model.add_data_point(x, y)

    def log_prob_normal(self, K, y):
        """log N(y; 0, K) for an explicit covariance (categorical_mixture.py:36-46), on the device."""
        n = int(y.shape[0])
        buf, ld = L.empty_matrix(n, n)
        buf.copy_(L.to_device(K))
        nblk = (n + L.DB - 1) // L.DB
        dinv = torch.empty((nblk, L.DB, L.DB), dtype=torch.float64, device=buf.device)
        info = torch.zeros((1,), dtype=torch.int32, device=buf.device)
        L.call("stpyb_potrf", L.ptr(buf), n, ld, L.ptr(dinv), L.ptr(info), int(GaussianProcess.outer_block),
               L.stream_ptr())
        z = L.to_device(y).reshape(-1).clone()
        L.call("stpyb_trsv", L.ptr(buf), n, ld, L.ptr(dinv), L.ptr(z), 0, L.stream_ptr())
        out3 = torch.empty((3,), dtype=torch.float64, device=buf.device)
        L.call("stpyb_lml", L.ptr(buf), n, ld, L.ptr(z), 1.0, L.ptr(out3), L.stream_ptr())
        val = float(out3.cpu()[2])
        if int(info.item()) != 0:
            raise torch.linalg.LinAlgError("log_prob_normal: covariance is not positive-definite "
                                           "(leading minor of order %d)" % int(info.item()))
        return -val - 0.5 * n * math.log(2 * math.pi)

    def _member_logprob(self, GP, y):
        if GP.Sigma is not None:  # custom noise covariance: not the s^2 I the evidence path assumes
            return self.log_prob_normal(GP.get_kernel(), y)
        lml = float(GP.log_marginal(GP.kernel_object, {}, 1.0))  # reuses the factor of the fit
        return -lml - 0.5 * GP.n * math.log(2 * math.pi)

    def _batchable(self):
        """True when every member is a plain GP whose evidence lml_sweep can score in one pass."""
        first_s, first_group = None, None
        for GP in self.processes:
            if type(GP) is not GaussianProcess or GP.Sigma is not None or GP.loss != "squared":
                return False
            try:
                _, _, _, group = sweep._isotropic_spec(GP.kernel_object)
            except NotImplementedError:
                return False
            s = float(GP.s)
            if first_s is None:
                first_s, first_group = s, group
            elif s != first_s or group != first_group:
                return False
        return self.k > 1

    def fit_gp(self, x, y, iterative=False):
        """Posterior model weights by log-sum-exp of prior weight + evidence (categorical_mixture.py:48-71)."""
        self.x = x
        self.y = y
        n = int(x.shape[0])
        if self.batched and self._batchable():
            vals = sweep.lml_sweep([GP.kernel_object for GP in self.processes], x, y, float(self.processes[0].s))
            logprobs = -vals.double() - 0.5 * n * math.log(2 * math.pi)
            for GP in self.processes:  # the member's own factor / alpha are built when first needed
                GP.load_data((x, y))
                GP.n, GP.d = n, int(x.shape[1])
                GP.fitted, GP._fit, GP.A, GP._A_dev = True, None, None, None
        else:
            logprobs = torch.zeros(size=(self.k, 1)).view(-1).double()
            for j in range(self.k):
                GP = self.processes[j]
                GP.fit(x, y)
                logprobs[j] = self._member_logprob(GP, y)
        self.logprobs = logprobs
        log_init_prob = torch.log(self.init_weights)
        log_posterior = log_init_prob + logprobs
        log_evidence = torch.logsumexp(log_posterior, dim=0)
        self.weights = torch.exp(log_posterior - log_evidence)
        self.fitted = True
        return True

    def mean_std(self, xtest):
        """Mixture mean and sqrt of the weighted member variances (categorical_mixture.py:73-83).  Members
        whose posterior weight underflowed to exactly 0 contribute exactly nothing and are not fitted."""
        on_dev = torch.is_tensor(xtest) and xtest.is_cuda
        xt = L.to_device(xtest)
        mu = torch.zeros(size=(xt.size()[0], 1), dtype=torch.float64, device=xt.device)
        s = torch.zeros(size=(xt.size()[0], 1), dtype=torch.float64, device=xt.device)
        for j in range(self.k):
            w = float(self.weights[j])
            if w == 0.0 and self.processes[j]._fit is None:
                continue
            (a1, a2) = self.processes[j].mean_std(xt)  # builds the member's factor on first use
            mu = mu + w * a1
            s = s + w * a2 ** 2
        s = torch.sqrt(s)
        return (mu, s) if on_dev else (mu.cpu(), s.cpu())

    def sample(self, xtest, size=1, with_mask=False):
        """Draw a member by its weight, then a path from it (categorical_mixture.py:85-111)."""
        p = self.weights.flatten().numpy()
        mask, cols = [], []
        for _ in range(size):
            k = int(np.random.choice(np.arange(0, self.k, 1), p=p))
            mask.append(k)
            if self.fitted and not self.processes[k].fitted:
                self.processes[k].fit_gp(self.x, self.y)
            self.processes[k]._ensure_factor()
            cols.append(self.processes[k].sample(xtest, size=1))
        samples = torch.cat(cols, dim=1)
        return (samples, mask) if with_mask else samples
