"""Stand-in for the reference's `Blas` backend.  TEST / BENCH INFRASTRUCTURE ONLY.

The reference's Blas backend (src/backends/blas.rs) is nothing but FFI calls into
an UN-VENDORED system BLAS through `cblas-sys = "0.1.4"` (Cargo.toml:19; provider
openblas-src 0.10.4 or blis-src 0.2.0, no Cargo.lock so no pinned version).  This
module issues the *identical* cblas calls -- same layout/trans/alpha/beta/ld
arguments as blas.rs:21,39-46,94-109,136-149,161 -- against the OpenBLAS 0.3.x that
scipy bundles (symbols `scipy_cblas_*`, LP64).  It is the second oracle for
matrix_mul / matrix_vector_mul / Blas.dot / nrm2 and the CPU timing baseline.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
RowMajor, NoTrans, Trans = 101, 111, 112

_blas = None
_loop = None


def _find_openblas() -> str:
    import scipy

    root = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
    hits = sorted(glob.glob(os.path.join(root, "libscipy_openblas*.so")))
    if not hits:
        raise RuntimeError("scipy's bundled OpenBLAS not found under " + root)
    return hits[0]


def blas() -> C.CDLL:
    global _blas
    if _blas is None:
        L = C.CDLL(_find_openblas())
        i, f, d, vp = C.c_int, C.c_float, C.c_double, C.c_void_p
        L.scipy_cblas_sdot.restype = f
        L.scipy_cblas_sdot.argtypes = [i, vp, i, vp, i]
        L.scipy_cblas_ddot.restype = d
        L.scipy_cblas_ddot.argtypes = [i, vp, i, vp, i]
        L.scipy_cblas_cdotu_sub.argtypes = [i, vp, i, vp, i, vp]
        L.scipy_cblas_zdotu_sub.argtypes = [i, vp, i, vp, i, vp]
        L.scipy_cblas_sgemm.argtypes = [i, i, i, i, i, i, f, vp, i, vp, i, f, vp, i]
        L.scipy_cblas_dgemm.argtypes = [i, i, i, i, i, i, d, vp, i, vp, i, d, vp, i]
        L.scipy_cblas_sgemv.argtypes = [i, i, i, i, f, vp, i, vp, i, f, vp, i]
        L.scipy_cblas_dgemv.argtypes = [i, i, i, i, d, vp, i, vp, i, d, vp, i]
        for n, r in (("snrm2", f), ("dnrm2", d), ("scnrm2", f), ("dznrm2", d)):
            fn = getattr(L, "scipy_cblas_" + n)
            fn.restype = r
            fn.argtypes = [i, vp, i]
        L.scipy_openblas_set_num_threads.argtypes = [i]
        L.scipy_openblas_get_num_threads.restype = i
        _blas = L
    return _blas


def loop_lib() -> C.CDLL:
    global _loop
    if _loop is None:
        so = os.path.join(_HERE, "_build", "libblas_loop.so")
        if not os.path.exists(so):
            subprocess.run(["make", "-C", _HERE, "-s"], check=True)
        L = C.CDLL(so)
        sz, vp, i = C.c_size_t, C.c_void_p, C.c_int
        L.blas_gemm_loop.restype = i
        L.blas_gemm_loop.argtypes = [vp, i, sz, vp, vp, vp, sz, sz, sz, i, i, i]
        L.blas_sdot_chunked.restype = C.c_double
        L.blas_sdot_chunked.argtypes = [vp, sz, vp, vp]
        _loop = L
    return _loop


def set_threads(n: int) -> None:
    blas().scipy_openblas_set_num_threads(int(n))


def _p(a: np.ndarray):
    return a.ctypes.data


def dot(a: np.ndarray, b: np.ndarray):
    """Blas::dot, blas.rs:15-23 (real) / :31-52 (complex, unconjugated)."""
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    L, n = blas(), a.size
    if a.dtype == np.float32:
        return np.float32(L.scipy_cblas_sdot(n, _p(a), 1, _p(b), 1))
    if a.dtype == np.float64:
        return np.float64(L.scipy_cblas_ddot(n, _p(a), 1, _p(b), 1))
    out = np.zeros(1, dtype=a.dtype)
    (L.scipy_cblas_cdotu_sub if a.dtype == np.complex64 else L.scipy_cblas_zdotu_sub)(
        n, _p(a), 1, _p(b), 1, _p(out))
    return out[0]


def norm(a: np.ndarray):
    """Blas::norm, blas.rs:156-162 (cblas_{s,d,sc,dz}nrm2)."""
    a = np.ascontiguousarray(a)
    name = {np.dtype(np.float32): "snrm2", np.dtype(np.float64): "dnrm2",
            np.dtype(np.complex64): "scnrm2", np.dtype(np.complex128): "dznrm2"}[a.dtype]
    r = getattr(blas(), "scipy_cblas_" + name)(a.size, _p(a), 1)
    return (np.float32 if a.dtype in (np.float32, np.complex64) else np.float64)(r)


def gemm(a, b, m, n, k, lda, ldb, ldc, a_trans, b_trans):
    """Blas::matrix_mul, blas.rs:94-109: RowMajor, alpha=1, beta=0.  Returns the m*ldc C buffer."""
    a, b = np.ascontiguousarray(a).ravel(), np.ascontiguousarray(b).ravel()
    c = np.zeros(m * ldc, dtype=a.dtype)
    f = blas().scipy_cblas_sgemm if a.dtype == np.float32 else blas().scipy_cblas_dgemm
    f(RowMajor, Trans if a_trans else NoTrans, Trans if b_trans else NoTrans, m, n, k, 1.0,
      _p(a), lda, _p(b), ldb, 0.0, _p(c), ldc)
    return c


def gemv(a, x, m, n, lda, a_trans):
    """Blas::matrix_vector_mul, blas.rs:136-149: cblas_gemv(RowMajor, op, M=n, N=m, ...)."""
    a, x = np.ascontiguousarray(a).ravel(), np.ascontiguousarray(x).ravel()
    y = np.zeros(m if a_trans else n, dtype=a.dtype)
    f = blas().scipy_cblas_sgemv if a.dtype == np.float32 else blas().scipy_cblas_dgemv
    f(RowMajor, Trans if a_trans else NoTrans, n, m, 1.0, _p(a), lda, _p(x), 1, 0.0, _p(y), 1)
    return y


def gemm_batched(a, b, c, batch, m, n, k, a_trans=False, b_trans=False, threads=1):
    """A C loop of `batch` cblas_gemm calls over contiguous objects, optionally split over pthreads.
    OpenBLAS's own threading is forced to 1 (tiny products never thread inside BLAS)."""
    assert a.dtype == b.dtype == c.dtype and a.flags.c_contiguous and c.flags.c_contiguous
    set_threads(1)
    is_d = a.dtype == np.float64
    fn = blas().scipy_cblas_dgemm if is_d else blas().scipy_cblas_sgemm
    rc = loop_lib().blas_gemm_loop(C.cast(fn, C.c_void_p), int(is_d), batch, _p(a), _p(b), _p(c),
                                   m, n, k, int(a_trans), int(b_trans), int(threads))
    if rc:
        raise MemoryError("blas_gemm_loop")
    return c
