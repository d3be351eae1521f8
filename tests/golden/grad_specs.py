"""Composite-kernel evidence-gradient cases shared by make_golden.py (which runs them through the UNMODIFIED
reference and its autograd) and tests/test_gpu_parity.py (which runs them through stpy_b200).  Each case builds
its kernel from whichever KernelFunction class it is handed -- the two classes take the same constructor
arguments -- and names the hyper-parameter tensors the evidence is differentiated in
(recipes: /root/reference/tests/marginalized_likelihood_test.py:15-108, tests/kernels/ard_matern_kernel_test.py:13-20)."""
import torch

F64 = torch.float64


def _t(*v):
    return torch.tensor(list(v), dtype=F64)


def cases():
    """name -> dict(n, d, seed, s, weight, build(KF) -> kernel, override() -> {index: {param: tensor}})"""
    return {
        "ard_matern52": dict(n=150, d=3, seed=50, s=0.1, weight=1.0,
                             build=lambda KF: KF(kernel_name="ard_matern", ard_gamma=_t(1.0, 1.0, 1.0), nu=2.5, d=3),
                             override=lambda: {'0': {'ard_gamma': _t(0.7, 1.3, 1.1), 'kappa': torch.tensor(1.4, dtype=F64)}}),
        "ard_matern32": dict(n=140, d=3, seed=51, s=0.15, weight=0.8,
                             build=lambda KF: KF(kernel_name="ard_matern", ard_gamma=_t(1.0, 1.0, 1.0), nu=1.5, d=3),
                             override=lambda: {'0': {'ard_gamma': _t(1.2, 0.8, 1.5)}}),
        "sum_ard_ard": dict(n=150, d=4, seed=52, s=0.1, weight=1.0,
                            build=lambda KF: KF(kernel_name="ard", ard_gamma=_t(1.0, 1.0), d=2, group=[0, 1])
                            + KF(kernel_name="ard", ard_gamma=_t(1.0, 1.0, 1.0, 1.0), d=2, group=[2, 3]),
                            override=lambda: {'0': {'ard_gamma': _t(0.8, 1.1)},
                                              '1': {'ard_gamma': _t(1.0, 1.0, 1.3, 0.9), 'kappa': torch.tensor(0.7, dtype=F64)}}),
        "prod_se_ardmatern": dict(n=130, d=4, seed=53, s=0.2, weight=1.0,
                                  build=lambda KF: KF(kernel_name="squared_exponential", gamma=1.0, d=4)
                                  * KF(kernel_name="ard_matern", ard_gamma=_t(1.0, 1.0, 1.0, 1.0), nu=2.5, d=4),
                                  override=lambda: {'0': {'gamma': torch.tensor(1.6, dtype=F64)},
                                                    '1': {'ard_gamma': _t(1.5, 2.0, 1.2, 1.8)}}),
        "additive_groups": dict(n=150, d=4, seed=54, s=0.1, weight=1.0,
                                build=lambda KF: KF(kernel_name="ard", ard_gamma=_t(1.0, 1.0, 1.0, 1.0), d=4,
                                                    groups=[[0, 1], [2, 3]]),
                                override=lambda: {'0': {'ard_gamma': _t(0.9, 1.2, 0.7, 1.4),
                                                        'kappa': torch.tensor(1.2, dtype=F64),
                                                        'groups': [[0, 1], [2, 3]]}}),
        "sum_ard_poly": dict(n=150, d=4, seed=55, s=0.2, weight=1.0,
                             build=lambda KF: KF(kernel_name="ard", ard_gamma=_t(1.0, 1.0, 1.0, 1.0), d=4)
                             + KF(kernel_name="polynomial", power=2, kappa=0.1, d=4),
                             override=lambda: {'0': {'ard_gamma': _t(0.9, 1.2, 1.0, 1.5)},
                                               '1': {'kappa': torch.tensor(0.15, dtype=F64), 'degree': 2}}),
        "fold3_noise": dict(n=120, d=3, seed=56, s=0.25, weight=1.0, noise_grad=True,
                            build=lambda KF: (KF(kernel_name="squared_exponential", gamma=1.0, d=3)
                                              + KF(kernel_name="ard", ard_gamma=_t(1.0, 1.0, 1.0), d=3))
                            * KF(kernel_name="polynomial", power=2, kappa=0.5, d=3),
                            override=lambda: {'0': {'gamma': torch.tensor(0.9, dtype=F64)},
                                              '1': {'ard_gamma': _t(1.4, 0.8, 1.1), 'kappa': torch.tensor(0.6, dtype=F64)},
                                              '2': {'kappa': torch.tensor(0.4, dtype=F64), 'degree': 2}}),
    }


def leaves(override):
    """[(index, name, tensor)] of the float tensors in an override tree, in a fixed order."""
    out = []
    for idx in sorted(override.keys()):
        for name in sorted(override[idx].keys()):
            v = override[idx][name]
            if torch.is_tensor(v) and v.dtype == F64:
                out.append((idx, name, v))
    return out
