"""ctypes binding of libstpyb.so (include/stpyb.h) and small device helpers.

There is no CPU fallback: if the library is missing or CUDA is unavailable the
first compute call raises.  torch is used only as the device-memory container
and stream provider.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstpyb.so")

c_dp = ctypes.c_void_p
c_i64 = ctypes.c_longlong
c_int = ctypes.c_int
c_dbl = ctypes.c_double

# name -> argtypes; every symbol include/stpyb.h declares must appear here
SIGNATURES = {
    "stpyb_version": [],
    "stpyb_profile": [c_int],
    "stpyb_profile_read": [c_dp, c_dp],
    "stpyb_potrf_diag_profile": [c_dp, c_i64, c_int, c_dp, c_dp, c_dp, c_dp],
    "stpyb_gram_prep": [c_dp, c_i64, c_i64, c_dp, c_int, c_dp, c_int, c_int, c_dp, c_int, c_dp, c_dp],
    "stpyb_gram": [c_int, c_dp, c_dp, c_i64, c_dp, c_dp, c_i64, c_int, c_dbl, c_dbl, c_dbl, c_int, c_int,
                   c_dbl, c_int, c_dp, c_i64, c_dp, c_dp],
    "stpyb_gram_diag": [c_int, c_dp, c_dp, c_dp, c_dp, c_i64, c_int, c_dbl, c_dbl, c_dbl, c_int, c_dp, c_dp, c_dp],
    "stpyb_gram_multi": [c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_i64, c_int, c_dbl, c_dp, c_i64, c_i64, c_dp],
    "stpyb_potrf": [c_dp, c_i64, c_i64, c_dp, c_dp, c_int, c_dp],
    "stpyb_set_lookahead_min_n": [c_i64, c_dp],
    "stpyb_trsv": [c_dp, c_i64, c_i64, c_dp, c_dp, c_int, c_dp],
    "stpyb_potrs_vec": [c_dp, c_i64, c_i64, c_dp, c_dp, c_dp],
    "stpyb_trsm_rt": [c_dp, c_i64, c_i64, c_dp, c_dp, c_i64, c_i64, c_dp],
    "stpyb_lml": [c_dp, c_i64, c_i64, c_dp, c_dbl, c_dp, c_dp],
    "stpyb_row_sumsq": [c_dp, c_i64, c_i64, c_i64, c_dp, c_int, c_dp, c_dp],
    "stpyb_gemv_rows": [c_dp, c_i64, c_i64, c_i64, c_dp, c_dp, c_dp],
    "stpyb_gemm_nt": [c_int, c_int, c_int, c_dp, c_i64, c_dp, c_i64, c_dp, c_i64, c_dbl, c_dbl, c_int, c_dp],
    "stpyb_potri": [c_dp, c_i64, c_i64, c_dp, c_dp, c_i64, c_dp, c_i64, c_dp],
    "stpyb_kernel_grad": [c_dp, c_i64, c_i64, c_dp, c_i64, c_i64, c_int, c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                          c_dp, c_int, c_dp, c_int, c_int, c_int, c_dp, c_i64, c_dp, c_dbl, c_dp, c_dp],
    "stpyb_rff_embed": [c_dp, c_i64, c_dp, c_int, c_int, c_dp, c_dp, c_int, c_dbl, c_int, c_dp, c_i64, c_dp],
    "stpyb_rff_normal_eq": [c_dp, c_dp, c_i64, c_dp, c_int, c_int, c_dp, c_dp, c_int, c_dbl, c_i64, c_dp, c_i64,
                            c_dp, c_i64, c_dp],
    "stpyb_gemv_t_sub": [c_dp, c_i64, c_int, c_i64, c_dp, c_dp, c_dp],
    "stpyb_gemm_nt_batch": [c_int, c_dp, c_dp, c_int, c_dp, c_i64, c_dp, c_i64, c_dp, c_i64, c_dbl, c_dbl, c_int, c_dp,
                            c_dp, c_int],
    "stpyb_p2p_alloc": [c_i64, c_dp, c_dp],
    "stpyb_p2p_open": [c_dp, c_dp],
    "stpyb_p2p_close": [c_dp],
    "stpyb_p2p_free": [c_dp],
    "stpyb_dist_strip": [c_dp, c_i64, c_int, c_i64, c_dp, c_dp, c_dp, c_int, c_int, c_i64, c_dp, c_dp],
    "stpyb_dist_solve_publish": [c_dp, c_i64, c_int, c_dp, c_dp, c_dp, c_dp, c_int, c_int, c_i64, c_int, c_int, c_dp],
    "stpyb_memcpy_d2d": [c_dp, c_dp, c_i64, c_dp],
    "stpyb_p2p_alpha_publish": [c_dp, c_dp, c_int, c_int, c_i64, c_int, c_int, c_int, c_dp],
    "stpyb_p2p_wait_flags": [c_dp, c_int, c_int, c_int, c_i64, c_dp, c_dp],
    "stpyb_dist_alpha_step": [c_dp, c_i64, c_i64, c_int, c_dp, c_dp, c_dp, c_dp, c_dp],
    "stpyb_jacobi_init": [c_dp, c_i64, c_dp, c_dp, c_i64, c_i64, c_i64, c_dp],
    "stpyb_jacobi_sweep": [c_dp, c_dp, c_i64, c_i64, c_dbl, c_dp, c_dp],
    "stpyb_jacobi_eigenvalues": [c_dp, c_dp, c_i64, c_i64, c_dp, c_dp],
    "stpyb_stack_combine": [c_dp, c_int, c_dp, c_i64, c_i64, c_i64, c_dbl, c_int, c_dp, c_i64, c_dp],
    "stpyb_stack_quadform": [c_dp, c_int, c_i64, c_i64, c_i64, c_int, c_dp, c_dp, c_dp],
    "stpyb_potrf_panel": [c_dp, c_i64, c_int, c_i64, c_dp, c_dp, c_i64, c_dp, c_i64, c_dp],
}

# kernel kinds / ops, mirrored from include/stpyb.h
K_SE, K_MATERN12, K_MATERN32, K_MATERN52, K_POLY, K_LINEAR, K_MATERN_NU = range(7)
OP_SET, OP_ADD, OP_MUL = range(3)
MAX_DIM = 64
DB = 128
POTRF_NO_LOOKAHEAD = 0x40000000

_lib = None


class StpybError(RuntimeError):
    pass


def load():
    """Load libstpyb.so; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StpybError(
            "libstpyb.so not found at %s: build it with `python stpy_b200/csrc/build.py` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.argtypes = argtypes
        fn.restype = c_int
    _lib = lib
    return lib


def check(rc, what=""):
    if rc == 0:
        return
    if rc < 0:
        raise StpybError("%s: invalid argument #%d" % (what, -rc))
    if rc >= 1000:
        raise StpybError("%s: CUDA error %d" % (what, rc - 1000))
    raise StpybError("%s: error %d" % (what, rc))


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    check(rc, name)


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor (or NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def device():
    if not torch.cuda.is_available():
        raise StpybError("stpy_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def to_device(t):
    """float64, contiguous, on the current CUDA device."""
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    return t.detach().to(device=device(), dtype=torch.float64).contiguous()


def pad_ld(n, mult=16):
    """Leading dimension: rows start on 128-byte boundaries."""
    return ((int(n) + mult - 1) // mult) * mult


def empty_matrix(rows, cols, zero=False):
    """(rows x cols) float64 view into a (rows x ld) allocation; returns (view, ld)."""
    ld = pad_ld(cols)
    buf = (torch.zeros if zero else torch.empty)((int(rows), ld), dtype=torch.float64, device=device())
    return buf[:, : int(cols)], ld


def host_doubles(vals):
    arr = (ctypes.c_double * len(vals))(*[float(v) for v in vals])
    return arr


def host_ints(vals):
    arr = (ctypes.c_int * len(vals))(*[int(v) for v in vals])
    return arr
