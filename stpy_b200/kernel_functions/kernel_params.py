"""KernelParams: keyword holder of the free-function kernels (stpy/kernel_functions/kernel_params.py:2-11)."""


class KernelParams():

    def __init__(self, param_dict):
        self.__dict__.update(param_dict)

    def assert_existence(self, names):
        for name in names:
            if name not in self.__dict__:
                raise AttributeError("Missing attribute of the kernel %s" % str(name))
