"""Estimator base: mirror of stpy/estimator.py::Estimator (load_data, log_marginal).

`log_marginal` is the Cholesky variant of the evidence (stpy/estimator.py:32-40);
on this path it is the same device computation as
GaussianProcess._log_marginal_squared (both evaluate 0.5 y^T K^-1 y + 0.5 w logdet K),
so subclasses that hold a `_lml_value` implementation inherit it from here.
"""
from abc import ABC, abstractmethod


class Estimator(ABC):

    def fit(self):
        pass

    @abstractmethod
    def ucb(self, x):
        pass

    @abstractmethod
    def lcb(self, x):
        pass

    def load_data(self, d):
        self.x = d[0]
        self.y = d[1]

    def log_marginal(self, kernel, X, weight):
        return self._lml_value(kernel, X, weight)
