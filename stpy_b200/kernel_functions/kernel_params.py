"""KernelParams: keyword holder of the free-function kernels (stpy/kernel_functions/kernel_params.py:2-11):
every keyword becomes an attribute; assert_existence names the first required one that is absent."""


class KernelParams:

    def __init__(self, param_dict):
        vars(self).update(param_dict)

    def assert_existence(self, names):
        missing = [str(n) for n in names if n not in vars(self)]
        if missing:
            raise AttributeError("Missing attribute of the kernel %s" % missing[0])
