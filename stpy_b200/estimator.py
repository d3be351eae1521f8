"""Estimator base: mirror of stpy/estimator.py::Estimator (load_data, log_marginal,
optimize_params_general).

`log_marginal` is the Cholesky variant of the evidence (stpy/estimator.py:32-40); on this path
it is the same device computation as GaussianProcess._log_marginal_squared.

`optimize_params_general` keeps the reference's parameter-tree protocol
(stpy/estimator.py:42-257): params = {kernel index | 'likelihood': {name: (init_func, manifold,
bounds)}}, `restarts` independent runs, the best point written back into
kernel_object.params_dict (and self.s), back_prop switched off, model refitted.  The optimiser
itself is scipy's L-BFGS-B driven by the device-evaluated value and ANALYTIC gradient
(stpy_b200/autodiff.py) -- the role pymanopt / torchmin / autograd_minimize play in the
reference, none of which is a dependency here.  X, y and the scratch factor stay resident on
the device across evaluations.
"""
import pickle
from abc import ABC, abstractmethod

import numpy as np
import torch


def _dim_of(manifold):
    """Dimension of a search space given as an int, a pymanopt-like manifold (.dim) or a sequence."""
    if isinstance(manifold, (int, np.integer)):
        return int(manifold)
    if hasattr(manifold, "dim"):
        return int(manifold.dim)
    return int(len(manifold))


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

    def optimize_params_general(self, params={}, restarts=2, optimizer="pytorch-minimize", maxiter=1000,
                                mingradnorm=1e-4, regularizer_func=None, verbose=False, scale=1., weight=1.,
                                save=False, save_name='model.np', parallel=False, cores=None):
        if optimizer not in ("pytorch-minimize", "scipy", "pymanopt", "l-bfgs-b"):
            raise AssertionError("Optimizer not implemented.")
        from scipy.optimize import minimize
        names, dims, inits, bounds = [], [0], [], []
        for key, dict_params in params.items():
            for var_name, value in dict_params.items():
                init_value, manifold, bound = value
                names.append((key, var_name))
                dims.append(_dim_of(manifold))
                inits.append(init_value)
                bounds.append(bound)
        dims = np.cumsum(dims).astype(int)
        dim = int(dims[-1])
        s_saved = self.s

        def value_and_grad(xnp):
            leaves = []
            input_dict = {}
            for i, (key, var_name) in enumerate(names):
                t = torch.tensor(xnp[dims[i]:dims[i + 1]], dtype=torch.float64, requires_grad=True)
                leaves.append(t)
                if key == "likelihood":
                    self.s = t[0] if t.numel() == 1 else t
                else:
                    input_dict.setdefault(key, {})[var_name] = t if t.numel() > 1 else t[0]
            f = self.log_marginal(self.kernel_object, input_dict, weight)
            if regularizer_func is not None:
                f = f + regularizer_func(leaves)
            f = f.reshape(())
            grads = torch.autograd.grad(f, leaves, allow_unused=True)
            g = np.concatenate([(gi if gi is not None else torch.zeros_like(li)).detach().cpu().numpy().reshape(-1)
                                for gi, li in zip(grads, leaves)])
            return float(f.detach()), g

        box = None
        if any(b is not None for b in bounds):
            box = []
            for i, b in enumerate(bounds):
                width = int(dims[i + 1] - dims[i])
                if b is None:
                    box += [(1e-8, None)] * width
                elif len(b) == width and not np.isscalar(b[0]):
                    box += [tuple(bb) for bb in b]
                else:
                    box += [tuple(b)] * width
        else:
            box = [(1e-8, None)] * dim  # lengthscales / noise are positive

        # parallel=True under an initialised torch.distributed group: the restarts are independent replicas, so rank r
        # runs restarts r, r + W, ... on its own GPU (no data-path collective) and the (value, point) pairs are gathered
        # at the end -- the reference accepts parallel= / cores= but never reads them (estimator.py:46).  Every rank
        # draws ALL starting points from the same host RNG stream so that the set of restarts does not depend on W.
        import torch.distributed as dist
        spread = bool(parallel) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        world, rank = (dist.get_world_size(), dist.get_rank()) if spread else (1, 0)
        objective_values, objective_params = [], []
        for rep in range(restarts):
            if inits[0] is None:
                x_init = (torch.randn(size=(dim, 1)).double().view(-1) ** 2 * scale).numpy()
            else:
                x_init = np.asarray(inits[0](dim), dtype=np.float64).reshape(-1)
            x_init = np.maximum(x_init, 1e-6)
            if rep % world != rank:
                continue
            try:
                res = minimize(value_and_grad, x_init, jac=True, method='L-BFGS-B', bounds=box,
                               options={'maxiter': maxiter, 'gtol': mingradnorm, 'ftol': 1e-12, 'maxls': 30})
                objective_params.append(torch.from_numpy(res.x.copy()))
                objective_values.append(float(res.fun))
            except torch.linalg.LinAlgError:
                continue  # a restart that walks into a non-PD Gram is dropped
            if verbose:
                print("restart %d: evidence %.6f after %d iterations" % (rep, objective_values[-1], res.nit))
        self.s = s_saved
        if spread:
            gathered = [None] * world
            dist.all_gather_object(gathered, (objective_values, [p.numpy() for p in objective_params]))
            objective_values = [v for vals, _ in gathered for v in vals]
            objective_params = [torch.from_numpy(p) for _, pts in gathered for p in pts]
        if not objective_values:
            raise RuntimeError("every restart failed")
        if save:
            with open(save_name, 'wb') as f:
                pickle.dump({'params': objective_params, 'evidence': objective_values, 'repeats': restarts,
                             'dim': dims, 'param_names': params}, f)
        best = int(np.argmin(objective_values))
        for i, (key, var_name) in enumerate(names):
            val = objective_params[best][dims[i]:dims[i + 1]].clone()
            if key == "likelihood":
                self.s = float(val[0])
            else:
                self.kernel_object.params_dict[key][var_name] = val if val.numel() > 1 else float(val[0])
        self.optimization_result = {'evidence': objective_values, 'params': objective_params, 'best': best}
        self.back_prop = False
        self.fitted = False
        self.fit_gp(self.x, self.y)
        return True
