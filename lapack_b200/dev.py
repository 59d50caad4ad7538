"""ctypes bindings of the device-pointer C API (include/lapack_b200.h) for torch CUDA tensors.

torch is used for device memory and streams only.  Matrices are column-major: a `colmajor(m, n)` tensor has
shape (m, n) and strides (1, ld).  All calls are asynchronous on torch's current stream.
"""
from __future__ import annotations

import ctypes as C

from . import lib


def _torch():
    import torch
    return torch


def stream():
    return C.c_void_p(_torch().cuda.current_stream().cuda_stream)


def colmajor(m, n, ld=None, dtype=None, device="cuda"):
    torch = _torch()
    ld = max(1, m) if ld is None else ld
    buf = torch.empty((max(1, n), ld), dtype=dtype or torch.float64, device=device)
    return buf.t()[:m, :n]


def to_colmajor(x):
    """copy of a torch matrix in column-major storage"""
    out = colmajor(x.shape[0], x.shape[1], dtype=x.dtype, device=x.device)
    out.copy_(x)
    return out


def ld(a):
    assert a.dim() == 2 and (a.stride(0) == 1 or a.shape[0] <= 1), "column-major tensor expected"
    return max(1, a.stride(1)) if a.shape[1] > 1 else max(1, a.shape[0], a.stride(1))


def _chk(rc):
    if rc != 0:
        raise RuntimeError(f"lapack_b200 device call failed: rc={rc} (CUDA error {-rc - 1000})")


def _c(ch):
    return ch.encode()


def gemm(ta, tb, alpha, a, b, beta, c):
    m, n = c.shape
    k = a.shape[1] if ta.upper() == "N" else a.shape[0]
    _chk(lib().lb200_dgemm(stream(), _c(ta), _c(tb), m, n, k, alpha, a.data_ptr(), ld(a), b.data_ptr(), ld(b), beta,
                           c.data_ptr(), ld(c)))


def syrk(uplo, trans, alpha, a, beta, c):
    n = c.shape[0]
    k = a.shape[1] if trans.upper() == "N" else a.shape[0]
    _chk(lib().lb200_dsyrk(stream(), _c(uplo), _c(trans), n, k, alpha, a.data_ptr(), ld(a), beta, c.data_ptr(), ld(c)))


def trsm(side, uplo, trans, diag, alpha, a, b):
    m, n = b.shape
    _chk(lib().lb200_dtrsm(stream(), _c(side), _c(uplo), _c(trans), _c(diag), m, n, alpha, a.data_ptr(), ld(a),
                           b.data_ptr(), ld(b)))


def trmm(side, uplo, trans, diag, alpha, a, b):
    m, n = b.shape
    _chk(lib().lb200_dtrmm(stream(), _c(side), _c(uplo), _c(trans), _c(diag), m, n, alpha, a.data_ptr(), ld(a),
                           b.data_ptr(), ld(b)))


def laswp(a, k1, k2, ipiv, incx):
    _chk(lib().lb200_dlaswp(stream(), a.shape[1], a.data_ptr(), ld(a), k1, k2, ipiv.data_ptr(), incx))


def getrf(a, recursive=False):
    """in place; returns (ipiv int32 device tensor, info int32 device tensor of 1 element)"""
    torch = _torch()
    m, n = a.shape
    ipiv = torch.zeros(max(1, min(m, n)), dtype=torch.int32, device=a.device)
    info = torch.zeros(1, dtype=torch.int32, device=a.device)
    fn = lib().lb200_dgetrf2 if recursive else lib().lb200_dgetrf
    _chk(fn(stream(), m, n, a.data_ptr(), ld(a), ipiv.data_ptr(), info.data_ptr()))
    return ipiv[:min(m, n)], info


def getrs(trans, a, ipiv, b):
    _chk(lib().lb200_dgetrs(stream(), _c(trans), a.shape[0], b.shape[1], a.data_ptr(), ld(a), ipiv.data_ptr(),
                            b.data_ptr(), ld(b)))


def potrf(uplo, a, recursive=False):
    torch = _torch()
    info = torch.zeros(1, dtype=torch.int32, device=a.device)
    fn = lib().lb200_dpotrf2 if recursive else lib().lb200_dpotrf
    _chk(fn(stream(), _c(uplo), a.shape[0], a.data_ptr(), ld(a), info.data_ptr()))
    return info


def potrs(uplo, a, b):
    _chk(lib().lb200_dpotrs(stream(), _c(uplo), a.shape[0], b.shape[1], a.data_ptr(), ld(a), b.data_ptr(), ld(b)))


def geqrf(a, unblocked=False):
    torch = _torch()
    m, n = a.shape
    tau = torch.zeros(max(1, min(m, n)), dtype=torch.float64, device=a.device)
    fn = lib().lb200_dgeqr2 if unblocked else lib().lb200_dgeqrf
    _chk(fn(stream(), m, n, a.data_ptr(), ld(a), tau.data_ptr()))
    return tau[:min(m, n)]


def larft(v, tau):
    n, k = v.shape
    t = colmajor(k, k, device=v.device)
    t.zero_()
    _chk(lib().lb200_dlarft(stream(), n, k, v.data_ptr(), ld(v), tau.data_ptr(), t.data_ptr(), ld(t)))
    return t


def larfb(side, trans, v, t, c):
    m, n = c.shape
    k = t.shape[0]
    _chk(lib().lb200_dlarfb(stream(), _c(side), _c(trans), m, n, k, v.data_ptr(), ld(v), t.data_ptr(), ld(t),
                            c.data_ptr(), ld(c)))


def getri(a, ipiv):
    """a (DGETRF factors) := inv(A) in place; returns the device INFO word"""
    torch = _torch()
    info = torch.zeros(1, dtype=torch.int32, device=a.device)
    _chk(lib().lb200_dgetri(stream(), a.shape[0], a.data_ptr(), ld(a), ipiv.data_ptr(), info.data_ptr()))
    return info


def ormqr(side, trans, a, tau, c):
    m, n = c.shape
    _chk(lib().lb200_dormqr(stream(), _c(side), _c(trans), m, n, tau.shape[0], a.data_ptr(), ld(a), tau.data_ptr(),
                            c.data_ptr(), ld(c)))


def orgqr(a, tau):
    m, n = a.shape
    _chk(lib().lb200_dorgqr(stream(), m, n, tau.shape[0], a.data_ptr(), ld(a), tau.data_ptr()))


def getrf_batched32(a):
    """a: (batch, 32, 32) tensor whose [b] slices are column-major, i.e. a contiguous (batch, 32(col), 32(row)) buffer"""
    torch = _torch()
    batch = a.shape[0]
    ipiv = torch.zeros((batch, 32), dtype=torch.int32, device=a.device)
    info = torch.zeros(batch, dtype=torch.int32, device=a.device)
    _chk(lib().lb200_dgetrf_batched32(stream(), batch, a.data_ptr(), ipiv.data_ptr(), info.data_ptr()))
    return ipiv, info


def potrf_batched32(uplo, a):
    torch = _torch()
    batch = a.shape[0]
    info = torch.zeros(batch, dtype=torch.int32, device=a.device)
    _chk(lib().lb200_dpotrf_batched32(stream(), _c(uplo), batch, a.data_ptr(), info.data_ptr()))
    return info


def larnv_matrix(m, n, iseed=(1988, 1989, 1990, 1991), offset=0, ld_=None, device="cuda"):
    a = colmajor(m, n, ld_, device=device)
    seed = (C.c_int * 4)(*iseed)
    _chk(lib().lb200_dlarnv_matrix(stream(), C.byref(seed), offset, m, n, a.data_ptr(), ld(a)))
    return a


def larnv_submatrix(a, stream_offset, stream_ld, iseed=(1988, 1989, 1990, 1991)):
    """fill the column-major view `a` with the window of the global DLARNV(2) matrix: a(i,j) = draw (offset + j*stream_ld + i)"""
    seed = (C.c_int * 4)(*iseed)
    _chk(lib().lb200_dlarnv_submatrix(stream(), C.byref(seed), stream_offset, stream_ld, a.shape[0], a.shape[1], a.data_ptr(), ld(a)))
    return a


def laswp_compose(ipiv_rel):
    """(src_top, inv_top) int32 device tensors for one panel's interchanges (see include/lapack_b200.h)"""
    torch = _torch()
    np_ = ipiv_rel.shape[0]
    src = torch.empty(np_, dtype=torch.int32, device=ipiv_rel.device)
    inv = torch.empty(np_, dtype=torch.int32, device=ipiv_rel.device)
    _chk(lib().lb200_laswp_compose(stream(), np_, ipiv_rel.data_ptr(), src.data_ptr(), inv.data_ptr()))
    return src, inv


def gather_rows(a, idx, w):
    """w(t, :) = a(idx[t], :) for idx[t] >= 0; a, w column-major views, idx int32"""
    _chk(lib().lb200_gather_rows(stream(), idx.shape[0], idx.data_ptr(), a.data_ptr(), ld(a), a.shape[1], w.data_ptr(), ld(w)))


def scatter_rows(w, idx, a):
    """a(idx[t], :) = w(t, :) for idx[t] >= 0"""
    _chk(lib().lb200_scatter_rows(stream(), idx.shape[0], idx.data_ptr(), w.data_ptr(), ld(w), a.shape[1], a.data_ptr(), ld(a)))


def make_spd(a, shift):
    _chk(lib().lb200_make_spd(stream(), a.shape[0], a.data_ptr(), ld(a), float(shift)))
    return a


def transpose(a):
    m, n = a.shape
    b = colmajor(n, m, device=a.device)
    _chk(lib().lb200_transpose(stream(), m, n, a.data_ptr(), ld(a), b.data_ptr(), ld(b)))
    return b
