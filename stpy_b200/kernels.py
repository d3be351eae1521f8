"""KernelFunction: the host-side mirror of stpy/kernels.py::KernelFunction.

Same constructor, same `kernel(a, b, **params_by_index)` / `kernel_diag` /
`get_kernel` / `params_dict` / `+` and `*` protocol (stpy/kernels.py:12-165),
same (|b|, |a|) orientation of the returned Gram matrix (stpy/kernels.py:393-398).
All arithmetic runs in libstpyb.so (stpyb_gram_prep + stpyb_gram): one fused
pass per sub-kernel, accumulated in place for the +/* algebra, instead of the
reference's chain of n^2 torch temporaries.

In scope: squared_exponential, ard (plain and additive groups),
squared_exponential_per_group, ard_per_group, matern (nu in {0.5, 1.5, 2.5} in closed form, any
other nu > 0 through the modified Bessel function K_nu), ard_matern (nu in {0.5, 1.5, 2.5}),
polynomial, linear, and `kernel_function=` callables
(the reference's operator seam, stpy/kernels.py:16-31).
"""
import math

import numpy as np
import torch

from . import _lib as L

_MATERN_KIND = {0.5: L.K_MATERN12, 1.5: L.K_MATERN32, 2.5: L.K_MATERN52}


def _f(v):
    """Python float of a scalar parameter (float, numpy scalar or 0/1-element tensor)."""
    if torch.is_tensor(v):
        return float(v.detach().reshape(-1)[0])
    return float(np.asarray(v).reshape(-1)[0]) if not isinstance(v, (int, float)) else float(v)


def _vec(v):
    """Host list of a vector parameter (tensor / ndarray / list)."""
    if torch.is_tensor(v):
        return [float(z) for z in v.detach().reshape(-1).cpu().tolist()]
    return [float(z) for z in np.asarray(v, dtype=np.float64).reshape(-1).tolist()]


class _Item:
    """One fused Gram launch: column selection, input scaling and the scalar map."""
    __slots__ = ("kind", "cols", "scale", "divide", "arg_scale", "kappa", "p0", "refine", "kparams")

    def __init__(self, kind, cols, scale=None, divide=0, arg_scale=0.0, kappa=1.0, p0=0.0, refine=0, kparams=None):
        self.kind, self.cols, self.scale, self.divide = kind, [int(c) for c in cols], scale, divide
        self.arg_scale, self.kappa, self.p0, self.refine = float(arg_scale), float(kappa), float(p0), refine
        self.kparams = kparams  # host doubles for STPYB_K_MATERN_NU (matern_nu_constants), else None


def matern_nu_constants(nu):
    """Host-side constants of the general-nu Matern map (stpy/kernels.py:852-859 evaluates it with scipy's kv):
    {nu, gam1, gam2, 1/Gamma(1+mu), 1/Gamma(1-mu), 2^(1-nu)/Gamma(nu)}, mu = nu - round(nu) in [-1/2, 1/2].
    gam1 = (1/Gamma(1-mu) - 1/Gamma(1+mu)) / (2 mu) cancels for small mu, so the Gamma values are taken in
    extended precision (mpmath, a dependency of torch's sympy) when available; the fallback switches to the
    Taylor limit gam1 -> -digamma(1) - ..., accurate to 1e-16 for |mu| < 1e-3."""
    nu = float(nu)
    if not nu > 0.0:
        raise ValueError("Matern nu must be positive")
    mu = nu - int(nu + 0.5)
    try:
        import mpmath
        with mpmath.workdps(40):
            m = mpmath.mpf(mu)
            gp, gm = mpmath.rgamma(1 + m), mpmath.rgamma(1 - m)
            gam1 = (gm - gp) / (2 * m) if m != 0 else mpmath.digamma(1)
            vals = [nu, float(gam1), float((gm + gp) / 2), float(gp), float(gm),
                    float(mpmath.power(2, 1 - mpmath.mpf(nu)) * mpmath.rgamma(mpmath.mpf(nu)))]
    except ImportError:
        gp, gm = 1.0 / math.gamma(1.0 + mu), 1.0 / math.gamma(1.0 - mu)
        euler, c3 = 0.5772156649015329, -0.04200263503409524  # 1/Gamma(1+z) = 1 + euler z + ... + c3 z^3 + ...
        gam1 = (gm - gp) / (2.0 * mu) if abs(mu) > 1e-3 else -(euler + c3 * mu * mu)
        vals = [nu, gam1, 0.5 * (gm + gp), gp, gm, 2.0 ** (1.0 - nu) / math.gamma(nu)]
    return vals


def _prep(x_dev, item, want_norms=True):
    """stpyb_gram_prep: (n x dpad) scaled / selected / zero-padded copy of x and its row norms."""
    n = x_dev.shape[0]
    dg = len(item.cols)
    if dg == 0 or dg > L.MAX_DIM:
        raise ValueError("a sub-kernel may select between 1 and %d input columns, got %d" % (L.MAX_DIM, dg))
    if max(item.cols) >= x_dev.shape[1]:
        raise IndexError("group index out of range for input with %d columns" % x_dev.shape[1])
    dpad = ((dg + 3) // 4) * 4
    xp = torch.empty((n, dpad), dtype=torch.float64, device=x_dev.device)
    nrm = torch.empty((n,), dtype=torch.float64, device=x_dev.device) if want_norms else None
    scale = item.scale if item.scale is not None else []
    L.call("stpyb_gram_prep", L.ptr(x_dev), n, x_dev.stride(0), L.host_ints(item.cols), dg,
           L.host_doubles(scale), len(scale), int(item.divide), L.ptr(xp), dpad, L.ptr(nrm), L.stream_ptr())
    return xp, nrm, dpad


class KernelFunction:

    def __init__(self, kernel_function=None, kernel_name="squared_exponential",
                 freq=None, groups=None, d=1, gamma=1, ard_gamma=None, nu=1.5, kappa=1, map=None, power=2,
                 cov=None, params=None, group=None, offset=0.):
        self.offset = offset
        self.kappa = kappa
        self.d = d
        self.group = [i for i in range(d)] if group is None else group
        if kernel_function is not None:
            # operator seam: a user callable(a, b, **params) -> (|b|, |a|) tensor
            self.kernel_function = kernel_function
            self.optkernel = "custom"
            self.params = {'kappa': self.kappa} if params is None else params
            self.initial_params = self.params
        else:
            self.optkernel = kernel_name
            self.gamma = gamma
            if ard_gamma is None:
                self.ard_gamma = torch.ones(d, dtype=torch.float64)
            elif torch.is_tensor(ard_gamma):
                self.ard_gamma = ard_gamma
            else:
                # the reference wraps scalars / lists as a (1, k) tensor (stpy/kernels.py:39-42)
                self.ard_gamma = torch.tensor([ard_gamma], dtype=torch.float64)
            self.power = power
            self.v = nu
            self.initial_params = params if params is not None else {'kappa': kappa}
            self.cov = torch.eye(d, dtype=torch.float64) if cov is None else cov
            self.map = map
            self.groups = groups
            self.freq = freq
            self.add = False
            self.params = self._default_params()

        self._owners = [self]
        self.kernel_function_list = [self._single_kernel]
        self.kernel_diag_function_list = [self._single_kernel_diag]
        self.optkernel_list = [self.optkernel]
        self.params_dict = {'0': self.params}
        self.kernel_items = 1
        self.operations = ["-"]

    # ------------------------------------------------------------------ registry
    def _default_params(self):
        """Per-kernel parameter dictionary, keyed as stpy/kernels.py:167-261 does."""
        p = {**self.initial_params, 'kappa': self.kappa, 'group': self.group, 'offset': self.offset}
        k = self.optkernel
        has_groups = self.groups is not None
        if k == "squared_exponential":
            p['gamma'] = self.gamma
        elif k == "ard" and not has_groups:
            p['ard_gamma'] = self.ard_gamma
        elif k == "ard" and has_groups:
            p['ard_gamma'] = self.ard_gamma
            p['groups'] = self.groups
        elif k == "linear":
            pass
        elif k == "matern":
            p['gamma'] = self.gamma
            p['nu'] = self.v
        elif k == "ard_matern":
            p['ard_gamma'] = self.ard_gamma
            p['nu'] = self.v
        elif k == "polynomial" and not has_groups:
            p['degree'] = self.power
        elif k == "polynomial" and has_groups:
            p['degree'] = self.power
            p['groups'] = self.groups
        elif k in ("squared_exponential_per_group", "ard_per_group") and has_groups:
            p['groups'] = self.groups
        else:
            raise AssertionError("Kernel '%s' is not implemented on the B200 path "
                                 "(out of scope, see SURVEY.md section 2)." % k)
        return p

    def _items(self, kw):
        """Resolve one sub-kernel (defaults from self, overrides from kw) into fused launches."""
        get = lambda key, default: kw[key] if key in kw else default
        k = self.optkernel
        kappa = _f(get('kappa', self.kappa))
        group = list(get('group', self.group))
        if k == "squared_exponential":
            gamma = _f(get('gamma', self.gamma))
            return [_Item(L.K_SE, group, arg_scale=-0.5 / (gamma * gamma), kappa=kappa)]
        if k == "ard":
            ard = _vec(get('ard_gamma', self.ard_gamma))
            groups = get('groups', self.groups)
            if groups is None:
                return [_Item(L.K_SE, group, scale=[1.0 / ard[g] for g in group], arg_scale=-0.5, kappa=kappa)]
            items = []
            for ga in groups:  # ard_kernel_additive, stpy/kernels.py:700-729
                items.append(_Item(L.K_SE, [group[g] for g in ga], scale=[1.0 / ard[g] for g in ga],
                                   arg_scale=-0.5, kappa=kappa / float(len(groups))))
            return items
        if k == "squared_exponential_per_group":  # stpy/kernels.py:668-698 (kappa enters twice there)
            groups = get('groups', self.groups)
            if 'gamma_per_group' not in kw:
                raise AssertionError("This kernel requires 'gamma_per_group' initial parameters")
            gammas = _vec(kw['gamma_per_group'])
            return [_Item(L.K_SE, ga, arg_scale=-0.5 / (g * g), kappa=kappa * kappa / float(len(groups)))
                    for ga, g in zip(groups, gammas)]
        if k == "ard_per_group":  # stpy/kernels.py:620-666
            groups = get('groups', self.groups)
            if 'ard_per_group' not in kw:
                raise AssertionError("This kernel requires 'ard_per_group' initial parameters")
            ard = _vec(kw['ard_per_group'])
            items, at = [], 0
            for ga in groups:
                gam = ard[at:at + len(ga)]
                at += len(ga)
                items.append(_Item(L.K_SE, ga, scale=[1.0 / g for g in gam], arg_scale=-0.5,
                                   kappa=kappa / float(len(groups))))
            return items
        if k == "matern":  # scipy cdist semantics: direct differences, x / gamma
            gamma = _f(get('gamma', self.gamma))
            nu = _f(get('nu', self.v))
            if nu not in _MATERN_KIND:  # general nu: the Bessel-function branch (kernels.py:852-859)
                return [_Item(L.K_MATERN_NU, group, scale=[gamma], divide=1, kappa=kappa, refine=1,
                              kparams=matern_nu_constants(nu))]
            return [_Item(_MATERN_KIND[nu], group, scale=[gamma], divide=1, kappa=kappa, refine=1)]
        if k == "ard_matern":  # torch.cdist semantics: clamped expansion
            ard = _vec(get('ard_gamma', self.ard_gamma))
            nu = _f(get('nu', self.v))
            if nu not in _MATERN_KIND:
                # the reference's own general branch fails on torch tensors here (kernels.py:964-970: K.fill is
                # not a tensor method), so there is nothing to mirror
                raise NotImplementedError("ard_matern: only nu in {0.5, 1.5, 2.5} (the general branch is broken in the "
                                          "reference; use kernel_name='matern' for a general nu)")
            return [_Item(_MATERN_KIND[nu], group, scale=[1.0 / ard[g] for g in group], kappa=kappa, refine=0)]
        if k == "polynomial":
            degree = _f(get('degree', self.power))
            groups = get('groups', self.groups)
            if groups is None:
                return [_Item(L.K_POLY, group, kappa=kappa, p0=degree)]
            return [_Item(L.K_POLY, [group[g] for g in ga], kappa=kappa / float(len(groups)), p0=degree)
                    for ga in groups]
        if k == "linear":
            return [_Item(L.K_LINEAR, group, kappa=kappa, p0=_f(get('offset', self.offset)))]
        raise AssertionError("Kernel not implemented.")

    # ------------------------------------------------------------------ derivative plan
    def _grad_items(self, kw, sub):
        """This sub-kernel as derivative-pass items (stpyb_kernel_grad): the same maps as _items, always in the
        scaled form sq = sum_c ((x_ic - x_jc) / lengthscale_c)^2, each column remembering WHICH parameter entry
        its lengthscale comes from so that the host can scatter the sums back
        (what autograd does through kernels.py:390-398, 572-583, 700-729, 944-962)."""
        get = lambda key, default: kw[key] if key in kw else default
        k = self.optkernel
        kap_src = get('kappa', self.kappa)
        kappa = _f(kap_src)
        group = list(get('group', self.group))

        def item(kind, cols, ls_src, ls_idx, kappa_item, dkappa, arg_scale=-0.5, p0=0.0):
            vals = None if ls_src is None else (_vec(ls_src) if not isinstance(ls_src, float) else [ls_src])
            ls = [1.0] * len(cols) if vals is None else [vals[i] for i in ls_idx]
            return {"kind": kind, "cols": [int(c) for c in cols], "sc": [1.0 / v for v in ls], "ls": ls,
                    "ls_src": ls_src, "ls_idx": list(ls_idx) if vals is not None else None,
                    "arg_scale": arg_scale, "kappa": kappa_item, "dkappa": dkappa, "kappa_src": kap_src,
                    "p0": float(p0), "sub": sub}

        if k == "squared_exponential":
            gam = get('gamma', self.gamma)
            return [item(L.K_SE, group, gam, [0] * len(group), kappa, 1.0)]
        if k in ("ard", "ard_matern"):
            ard = get('ard_gamma', self.ard_gamma)
            kind = L.K_SE
            if k == "ard_matern":
                nu = _f(get('nu', self.v))
                if nu not in _MATERN_KIND:
                    raise NotImplementedError("Matern nu=%s has no analytic derivative on the B200 path" % nu)
                kind = _MATERN_KIND[nu]
            groups = get('groups', self.groups) if k == "ard" else None
            if groups is None:
                return [item(kind, group, ard, list(group), kappa, 1.0)]
            ng = float(len(groups))
            return [item(kind, [group[g] for g in ga], ard, list(ga), kappa / ng, 1.0 / ng) for ga in groups]
        if k == "matern":
            nu = _f(get('nu', self.v))
            if nu not in _MATERN_KIND:
                raise NotImplementedError("Matern nu=%s has no analytic derivative on the B200 path" % nu)
            gam = get('gamma', self.gamma)
            return [item(_MATERN_KIND[nu], group, gam, [0] * len(group), kappa, 1.0)]
        if k == "squared_exponential_per_group":  # kappa enters twice there (kernels.py:668-698)
            groups = get('groups', self.groups)
            gam = kw['gamma_per_group']
            ng = float(len(groups))
            return [item(L.K_SE, ga, gam, [i] * len(ga), kappa * kappa / ng, 2.0 * kappa / ng)
                    for i, ga in enumerate(groups)]
        if k == "ard_per_group":
            groups = get('groups', self.groups)
            ard = kw['ard_per_group']
            ng = float(len(groups))
            out, at = [], 0
            for ga in groups:
                out.append(item(L.K_SE, ga, ard, list(range(at, at + len(ga))), kappa / ng, 1.0 / ng))
                at += len(ga)
            return out
        if k == "polynomial":
            degree = _f(get('degree', self.power))
            groups = get('groups', self.groups)
            if groups is None:
                return [item(L.K_POLY, group, None, [], kappa, 1.0, 0.0, degree)]
            ng = float(len(groups))
            return [item(L.K_POLY, [group[g] for g in ga], None, [], kappa / ng, 1.0 / ng, 0.0, degree) for ga in groups]
        if k == "linear":
            return [item(L.K_LINEAR, group, None, [], kappa, 1.0, 0.0, _f(get('offset', self.offset)))]
        raise NotImplementedError("kernel '%s' has no analytic derivative on the B200 path (plug-in callables and "
                                  "out-of-scope kernels)" % k)

    def grad_plan(self, params_dict):
        """(items, sub_ops) of the composite kernel for the derivative pass."""
        items, sub_ops = [], []
        for i, owner in enumerate(self._owners):
            kw = params_dict[str(i)] if str(i) in params_dict else {}
            items += owner._grad_items(kw, i)
            sub_ops.append({"-": L.OP_SET, "+": L.OP_ADD, "*": L.OP_MUL}[self.operations[i]])
        return items, sub_ops

    # ------------------------------------------------------------------ algebra
    def __combine__(self, other):
        self._owners = self._owners + other._owners
        self.kernel_function_list = self.kernel_function_list + other.kernel_function_list
        self.kernel_diag_function_list = self.kernel_diag_function_list + other.kernel_diag_function_list
        self.optkernel_list = self.optkernel_list + other.optkernel_list
        self.operations = self.operations + other.operations[1:]
        for _, value in other.params_dict.items():
            self.params_dict[str(self.kernel_items)] = value
            self.kernel_items += 1

    def __add__(self, other):
        # mutates and returns self, as the reference does (stpy/kernels.py:84-94)
        self.__combine__(other)
        self.d += len(set(other.group) - set(self.group))
        self.operations.append("+")
        return self

    def __mul__(self, other):
        self.__combine__(other)
        self.operations.append("*")
        return self

    def description(self):
        desc = "Kernel description:"
        for index in range(self.kernel_items):
            desc += "\n\n\tkernel: " + self.optkernel_list[index]
            desc += "\n\toperation: " + self.operations[index]
            desc += "\n\t" + "\n\t".join(
                "{0}={1}".format(key, value) for key, value in self.params_dict[str(index)].items())
        return desc

    def add_groups(self, dict):
        """Complete a per-index override tree with every sub-kernel's column group (kernels.py:96-101)."""
        for index, own in self.params_dict.items():
            dict.setdefault(index, {})['group'] = own['group']
        return dict

    def get_param_refs(self):
        return self.params_dict

    def get_kernel(self):
        return self.kernel

    def embed(self, x):
        if self.optkernel == "linear":
            return x
        raise AttributeError("This type of kernel does not support a finite dimensional embedding")

    def get_basis_size(self):
        if self.optkernel == "linear":
            return self.d
        raise AttributeError("This type of kernel does not support a finite dimensional embedding")

    # ------------------------------------------------------------------ device path
    def _resolve_params(self, kwargs):
        if len(kwargs) > 0:
            params_dict = kwargs
            self.add_groups(params_dict)
            return params_dict
        return self.params_dict

    def gram_into(self, a_dev, b_dev, params_dict, out, ld, symmetric=False, lower_only=False, diag_add=0.0):
        """Fill `out` ((|b| x |a|) view of an (|b| x ld) buffer) with the composite Gram matrix.

        The left fold out = k0; out = out (+|*) k_i of stpy/kernels.py:146-157 is done in
        place by the kernels' accumulate mode; diag_add (s^2) rides on the last launch."""
        n, m = a_dev.shape[0], b_dev.shape[0]
        plan = []  # (op, items | callable, kw)
        for i, owner in enumerate(self._owners):
            kw = params_dict[str(i)] if str(i) in params_dict else {}
            op = {"-": L.OP_SET, "+": L.OP_ADD, "*": L.OP_MUL}[self.operations[i]]
            if owner.optkernel == "custom":
                plan.append((op, owner, kw))
            else:
                plan.append((op, owner._items(kw), kw))
        pending_diag = float(diag_add)
        for pi, (op, what, kw) in enumerate(plan):
            last_sub = pi == len(plan) - 1
            if isinstance(what, KernelFunction):  # plug-in callable: evaluated by the caller's code
                val = what.kernel_function(a_dev, b_dev, **kw).to(out.device, torch.float64)
                if op == L.OP_SET:
                    out.copy_(val)
                elif op == L.OP_ADD:
                    out.add_(val)
                else:
                    out.mul_(val)
                continue
            items = what
            if op == L.OP_MUL and len(items) > 1:
                # (additive sub-kernel) as a factor: needs its own buffer before the product
                tmp, ldt = L.empty_matrix(m, n)
                for q, it in enumerate(items):
                    self._launch(it, a_dev, b_dev, symmetric, tmp, ldt, L.OP_SET if q == 0 else L.OP_ADD, 0.0, False)
                out.mul_(tmp)
                continue
            for q, it in enumerate(items):
                this_op = op if q == 0 else L.OP_ADD
                last = last_sub and q == len(items) - 1
                self._launch(it, a_dev, b_dev, symmetric, out, ld, this_op, pending_diag if last else 0.0,
                             lower_only)
                if last:
                    pending_diag = 0.0
        if pending_diag != 0.0:
            out.diagonal().add_(pending_diag)
        return out

    @staticmethod
    def _launch(it, a_dev, b_dev, symmetric, out, ld, op, diag_add, lower_only):
        n, m = a_dev.shape[0], b_dev.shape[0]
        ap, na, dpad = _prep(a_dev, it)
        if symmetric:
            bp, nb = ap, na
        else:
            bp, nb, _ = _prep(b_dev, it)
        L.call("stpyb_gram", it.kind, L.ptr(ap), L.ptr(na), n, L.ptr(bp), L.ptr(nb), m, dpad, it.arg_scale,
               it.kappa, it.p0, int(it.refine), int(op), float(diag_add), int(bool(lower_only)), L.ptr(out), ld,
               L.host_doubles(it.kparams) if it.kparams else None, L.stream_ptr())

    def kernel(self, a, b, **kwargs):
        """Gram matrix K[j, i] = k(b_j, a_i), shape (|b|, |a|); returned on a's device."""
        params_dict = self._resolve_params(kwargs)
        a_dev, b_dev = L.to_device(a), L.to_device(b)
        symmetric = (a is b) or (a_dev.data_ptr() == b_dev.data_ptr() and a_dev.shape == b_dev.shape)
        on_dev = torch.is_tensor(a) and a.is_cuda
        from . import autodiff
        if autodiff.needs_grad(params_dict):
            # operator-seam contract (kernels.py:136-159): differentiable in the tensors of the parameter tree
            return autodiff.gram_with_grad(self, params_dict, a_dev, b_dev, symmetric, not on_dev)
        out, ld = L.empty_matrix(b_dev.shape[0], a_dev.shape[0])
        self.gram_into(a_dev, b_dev, params_dict, out, ld, symmetric=symmetric)
        return out if on_dev else out.cpu()

    def kernel_diag(self, a, b, **kwargs):
        """k(b_i, a_i) for paired rows (stpy/kernels.py:112-134), shape (n,)."""
        params_dict = self._resolve_params(kwargs)
        a_dev, b_dev = L.to_device(a), L.to_device(b)
        out = self.diag_device(a_dev, b_dev, params_dict)
        return out if (torch.is_tensor(a) and a.is_cuda) else out.cpu()

    def diag_device(self, a_dev, b_dev, params_dict):
        n = a_dev.shape[0]
        out = torch.empty((n,), dtype=torch.float64, device=a_dev.device)
        for i, owner in enumerate(self._owners):
            kw = params_dict[str(i)] if str(i) in params_dict else {}
            op = {"-": L.OP_SET, "+": L.OP_ADD, "*": L.OP_MUL}[self.operations[i]]
            if owner.optkernel == "custom":
                vals = torch.stack([owner.kernel_function(a_dev[j:j + 1], b_dev[j:j + 1], **kw).reshape(())
                                    for j in range(n)])
                out = vals if op == L.OP_SET else (out + vals if op == L.OP_ADD else out * vals)
                continue
            items = owner._items(kw)
            target = out
            if op == L.OP_MUL and len(items) > 1:
                target = torch.empty_like(out)
            for q, it in enumerate(items):
                ap, na, dpad = _prep(a_dev, it)
                bp, nb, _ = _prep(b_dev, it)
                this_op = (op if target is out else L.OP_SET) if q == 0 else L.OP_ADD
                L.call("stpyb_gram_diag", it.kind, L.ptr(ap), L.ptr(na), L.ptr(bp), L.ptr(nb), n, dpad,
                       it.arg_scale, it.kappa, it.p0, int(this_op), L.ptr(target),
                       L.host_doubles(it.kparams) if it.kparams else None, L.stream_ptr())
            if target is not out:
                out.mul_(target)
        return out

    # single sub-kernel callables kept for code that indexes kernel_function_list
    def _single_kernel(self, a, b, **kw):
        a_dev, b_dev = L.to_device(a), L.to_device(b)
        out, ld = L.empty_matrix(b_dev.shape[0], a_dev.shape[0])
        for q, it in enumerate(self._items(kw)):
            self._launch(it, a_dev, b_dev, False, out, ld, L.OP_SET if q == 0 else L.OP_ADD, 0.0, False)
        return out if (torch.is_tensor(a) and a.is_cuda) else out.cpu()

    def _single_kernel_diag(self, a, b, **kw):
        single = KernelFunction.__new__(KernelFunction)
        single.__dict__.update(self.__dict__)
        single._owners, single.operations = [self], ["-"]
        return single.kernel_diag(a, b, **({'0': kw} if kw else {}))

    # convenience names the reference exposes as bound methods
    def squared_exponential_kernel(self, a, b, **kw):
        return self._named("squared_exponential", a, b, kw)

    def ard_kernel(self, a, b, **kw):
        return self._named("ard", a, b, kw)

    def matern_kernel(self, a, b, **kw):
        return self._named("matern", a, b, kw)

    def ard_matern_kernel(self, a, b, **kw):
        return self._named("ard_matern", a, b, kw)

    def polynomial_kernel(self, a, b, **kw):
        return self._named("polynomial", a, b, kw)

    def linear_kernel(self, a, b, **kw):
        return self._named("linear", a, b, kw)

    def _named(self, name, a, b, kw):
        if getattr(self, "optkernel", None) == name and (name != "ard" or self.groups is None):
            return self._single_kernel(a, b, **kw)
        d = max(int(a.shape[1]), 1)
        tmp = KernelFunction(kernel_name=name, d=d, kappa=getattr(self, "kappa", 1.),
                             gamma=getattr(self, "gamma", 1), ard_gamma=getattr(self, "ard_gamma", None),
                             nu=getattr(self, "v", 1.5), power=getattr(self, "power", 2),
                             offset=getattr(self, "offset", 0.))
        return tmp._single_kernel(a, b, **kw)
