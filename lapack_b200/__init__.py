"""lapack_b200 -- B200-native (sm_100a) implementation of LAPACK's blocked one-sided factorizations.

The product is the C-ABI shared library ``liblapack_b200.so`` (sources in ``csrc/``, headers in
``../include``): Fortran-77 symbols (``dgetrf_`` ...), LAPACKE entry points and a device-pointer API
(``lb200_*``).  This Python package is only the thin loader used by the tests and the benchmark: it binds
the C ABI with ctypes and passes raw pointers (numpy host arrays, torch device tensors).  There is no
Python or CPU implementation of any routine here -- if the CUDA library is missing, importing fails.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblapack_b200.so")
_lib = None

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)
LL = C.c_longlong
VP = C.c_void_p


class LibraryMissing(ImportError):
    pass


def lib() -> C.CDLL:
    """Load liblapack_b200.so (RTLD_GLOBAL so a separately loaded LAPACKE layer can bind to dgetrf_ ...)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not found: build it with `python -m lapack_b200.build` (nvcc, sm_100a). "
            "lapack_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    _declare(L)
    _lib = L
    return L


def _declare(L):
    i, d, ch = C.c_int, C.c_double, C.c_char
    L.lb200_version.restype = i
    L.lb200_launch_count.restype = C.c_ulonglong
    L.lb200_fp64_peak_tflops.restype = d
    L.lb200_fp64_peak_tflops.argtypes = [VP, i, i, i, i]
    L.lb200_set_gemm_config.argtypes = [i]
    L.lb200_set_trsm_inverse.argtypes = [i]
    L.lb200_set_gemm_splitk_balance.argtypes = [i]
    L.lb200_set_getrf_params.argtypes = [i, i, i]
    L.lb200_set_getrf_cluster_max.argtypes = [i]
    L.lb200_set_getrf_big_leaf.argtypes = [i]
    L.lb200_set_getrf_tall_rows.argtypes = [i]
    L.lb200_set_getrf_cluster_fat.argtypes = [i]
    L.lb200_set_getrf_thin.argtypes = [i, i]
    L.lb200_set_getrf_super.argtypes = [i]
    L.lb200_set_getrf_defer_left.argtypes = [i, i]
    L.lb200_set_batched_mode.argtypes = [i]
    L.lb200_set_laswp_bulk.argtypes = [i]
    L.lb200_set_geqrf_cluster_max.argtypes = [i]
    L.lb200_set_potrf_params.argtypes = [i, i]
    L.lb200_set_geqrf_params.argtypes = [i, i]
    L.lb200_dgemm.argtypes = [VP, ch, ch, i, i, i, d, VP, LL, VP, LL, d, VP, LL]
    L.lb200_dsyrk.argtypes = [VP, ch, ch, i, i, d, VP, LL, d, VP, LL]
    L.lb200_dtrsm.argtypes = [VP, ch, ch, ch, ch, i, i, d, VP, LL, VP, LL]
    L.lb200_dtrmm.argtypes = [VP, ch, ch, ch, ch, i, i, d, VP, LL, VP, LL]
    L.lb200_dlaswp.argtypes = [VP, i, VP, LL, i, i, VP, i]
    L.lb200_dgetrf.argtypes = [VP, i, i, VP, LL, VP, VP]
    L.lb200_dgetrf2.argtypes = [VP, i, i, VP, LL, VP, VP]
    L.lb200_dgetrs.argtypes = [VP, ch, i, i, VP, LL, VP, VP, LL]
    L.lb200_dpotrf.argtypes = [VP, ch, i, VP, LL, VP]
    L.lb200_dpotrf2.argtypes = [VP, ch, i, VP, LL, VP]
    L.lb200_dpotrs.argtypes = [VP, ch, i, i, VP, LL, VP, LL]
    L.lb200_dgeqrf.argtypes = [VP, i, i, VP, LL, VP]
    L.lb200_dgeqr2.argtypes = [VP, i, i, VP, LL, VP]
    L.lb200_dlarft.argtypes = [VP, i, i, VP, LL, VP, VP, LL]
    L.lb200_dlarfb.argtypes = [VP, ch, ch, i, i, i, VP, LL, VP, LL, VP, LL]
    L.lb200_dgetri.argtypes = [VP, i, VP, LL, VP, VP]
    L.lb200_dgeqrt.argtypes = [VP, i, i, i, VP, LL, VP, LL]
    L.lb200_dgemqrt.argtypes = [VP, ch, ch, i, i, i, i, VP, LL, VP, LL, VP, LL]
    L.lb200_dormqr.argtypes = [VP, ch, ch, i, i, i, VP, LL, VP, VP, LL]
    L.lb200_dorgqr.argtypes = [VP, i, i, i, VP, LL, VP]
    L.lb200_dgetrf_batched32.argtypes = [VP, LL, VP, VP, VP]
    L.lb200_dpotrf_batched32.argtypes = [VP, ch, LL, VP, VP]
    L.lb200_dlarnv_matrix.argtypes = [VP, C.POINTER(C.c_int * 4), LL, i, i, VP, LL]
    L.lb200_make_spd.argtypes = [VP, i, VP, LL, d]
    L.lb200_dlarnv_submatrix.argtypes = [VP, C.POINTER(C.c_int * 4), LL, LL, i, i, VP, LL]
    L.lb200_laswp_compose.argtypes = [VP, i, VP, VP, VP]
    L.lb200_gather_rows.argtypes = [VP, i, VP, VP, LL, i, VP, LL]
    L.lb200_scatter_rows.argtypes = [VP, i, VP, VP, LL, i, VP, LL]
    L.lb200_set_fewrhs_mode.argtypes = [i]
    L.lb200_set_l2_fetch_granularity.argtypes = [i]
    L.lb200_dlacpy.argtypes = [VP, ch, i, i, VP, LL, VP, LL]
    L.lb200_transpose.argtypes = [VP, i, i, VP, LL, VP, LL]
    L.lb200_set_xerbla_mode.argtypes = [i]
    L.lb200_last_xerbla.argtypes = [C.c_char_p, c_int_p]
    L.lb200_last_xerbla.restype = i
    for name in ("dgemm", "dsyrk", "dtrsm", "dtrmm", "dlaswp", "dgetrf", "dgetrf2", "dgetrs", "dpotrf", "dpotrf2",
                 "dpotrs", "dgeqrf", "dgeqr2", "dlarft", "dlarfb", "dgetrf_batched32", "dpotrf_batched32",
                 "dlarnv_matrix", "make_spd", "dlacpy", "transpose"):
        getattr(L, "lb200_" + name).restype = i


from . import f77, dev  # noqa: E402,F401  (thin ctypes front-ends over the C ABI)
