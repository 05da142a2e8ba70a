"""ctypes/numpy face of the CPU oracle (oracle/slas_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Nothing under slas_b200/ does.

Each wrapper cites the reference function it restates (paths relative to
/root/reference); see slas_oracle.h for the full notes and the "parity unpinned"
list (real nrm2, dgemm, dgemv, ddot, zdotu, dznrm2 have no golden vector upstream).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")


def build(force: bool = False) -> None:
    """Compile the C restatement (gcc, seconds).  Building the checker is not using it."""
    so = os.path.join(_BUILD, "libslas_oracle.so")
    src = os.path.join(_HERE, "slas_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def _cpu_has_avx2() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            return " avx2" in f.read()
    except OSError:
        return False


_libs: dict[str, C.CDLL] = {}
_sz = C.c_size_t
_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)


def lib(fast: bool = False) -> C.CDLL:
    """fast=True picks the -O3 -mavx2 timing build when the host supports it."""
    name = "libslas_oracle_avx2.so" if (fast and _cpu_has_avx2()) else "libslas_oracle.so"
    if name not in _libs:
        build()
        L = C.CDLL(os.path.join(_BUILD, name))
        L.oracle_sdot.restype = C.c_float
        L.oracle_sdot.argtypes = [_sz, _fp, _fp, C.c_int]
        L.oracle_ddot.restype = C.c_double
        L.oracle_ddot.argtypes = [_sz, _dp, _dp, C.c_int]
        L.oracle_cdotu.argtypes = [_sz, _fp, _fp, _fp]
        L.oracle_zdotu.argtypes = [_sz, _dp, _dp, _dp]
        L.oracle_sewise.argtypes = [C.c_int, _sz, _fp, _fp, _fp]
        L.oracle_dewise.argtypes = [C.c_int, _sz, _dp, _dp, _dp]
        L.oracle_snorm.restype = C.c_float
        L.oracle_snorm.argtypes = [_sz, _fp]
        L.oracle_dnorm.restype = C.c_double
        L.oracle_dnorm.argtypes = [_sz, _dp]
        L.oracle_cnorm.restype = C.c_float
        L.oracle_cnorm.argtypes = [_sz, _fp]
        L.oracle_znorm.restype = C.c_double
        L.oracle_znorm.argtypes = [_sz, _dp]
        for n, p in (("oracle_snormalize", _fp), ("oracle_dnormalize", _dp),
                     ("oracle_cnormalize", _fp), ("oracle_znormalize", _dp)):
            getattr(L, n).argtypes = [_sz, p]
            getattr(L, n).restype = None
        L.oracle_transpose.argtypes = [_sz, _sz, C.c_void_p, C.c_void_p, _sz]
        L.oracle_transpose_inplace.argtypes = [_sz, _sz, C.c_void_p, _sz]
        for n, p in (("oracle_sgemm", _fp), ("oracle_sgemm_f64acc", _fp), ("oracle_dgemm", _dp)):
            getattr(L, n).argtypes = [p, p, p, _sz, _sz, _sz, _sz, _sz, _sz, C.c_int, C.c_int]
            getattr(L, n).restype = None
        for n, p in (("oracle_sgemv", _fp), ("oracle_dgemv", _dp)):
            getattr(L, n).argtypes = [p, p, p, _sz, _sz, _sz, C.c_int]
            getattr(L, n).restype = None
        L.oracle_tensor_index.restype = _sz
        L.oracle_tensor_index.argtypes = [_sz, C.POINTER(_sz), C.POINTER(_sz)]
        L.oracle_matrix_index.restype = _sz
        L.oracle_matrix_index.argtypes = [C.POINTER(_sz), C.c_int, _sz, _sz]
        L.oracle_matmul_args.argtypes = [C.POINTER(_sz), C.c_int, C.POINTER(_sz), C.c_int, C.POINTER(_sz)]
        for n, p in (("oracle_sgemm_batched", _fp), ("oracle_dgemm_batched", _dp)):
            getattr(L, n).argtypes = [_sz, p, p, p, _sz, _sz, _sz, C.c_int, C.c_int]
            getattr(L, n).restype = None
        _libs[name] = L
    return _libs[name]


def _c(a: np.ndarray, dt) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dt)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_fp if a.dtype == np.float32 else _dp)


def _real_dtype(a) -> np.dtype:
    dt = np.asarray(a).dtype
    if dt in (np.float32, np.complex64):
        return np.dtype(np.float32)
    return np.dtype(np.float64)


# Default lane counts of the reference on a default x86-64 build (SSE2): 4 f32 / 2 f64
# (src/simd_lanes.rs:24-37); 8/4 when built with AVX.
def default_lanes(dtype, avx: bool = False) -> int:
    base = 8 if avx else 4
    return base if np.dtype(dtype) == np.float32 else base // 2


def dot(a, b, lanes: int | None = None, fast: bool = False):
    """Rust::dot, src/backends/rust.rs:20-41."""
    a = np.asarray(a)
    dt = _real_dtype(a)
    a, b = _c(a, dt), _c(b, dt)
    assert a.shape == b.shape and a.ndim == 1
    lanes = default_lanes(dt) if lanes is None else lanes
    L = lib(fast)
    if dt == np.float32:
        return np.float32(L.oracle_sdot(a.size, _ptr(a), _ptr(b), lanes))
    return np.float64(L.oracle_ddot(a.size, _ptr(a), _ptr(b), lanes))


def dotu(a, b):
    """Blas::dot for Complex, cblas_{c,z}dotu_sub, src/backends/blas.rs:31-52."""
    a = np.asarray(a)
    cdt = np.complex64 if a.dtype == np.complex64 else np.complex128
    a, b = _c(a, cdt), _c(b, cdt)
    rdt = np.float32 if cdt == np.complex64 else np.float64
    av, bv = a.view(rdt), b.view(rdt)
    out = np.zeros(2, dtype=rdt)
    (lib().oracle_cdotu if rdt == np.float32 else lib().oracle_zdotu)(a.size, _ptr(av), _ptr(bv), _ptr(out))
    return cdt(complex(out[0], out[1]))


_OPS = {"add": 0, "sub": 1, "mul": 2, "div": 3}


def ewise(op: str, a, b, fast: bool = False):
    """Rust::{add,sub,mul,div}, src/backends/rust.rs:45-73."""
    a = np.asarray(a)
    dt = _real_dtype(a)
    a, b = _c(a, dt), _c(b, dt)
    c = np.empty_like(a)
    L = lib(fast)
    (L.oracle_sewise if dt == np.float32 else L.oracle_dewise)(_OPS[op], a.size, _ptr(a), _ptr(b), _ptr(c))
    return c


def norm(a, fast: bool = False):
    """Rust::norm (real: rust.rs:116-121; complex: rust.rs:129-139)."""
    a = np.ascontiguousarray(a)
    L = lib(fast)
    if a.dtype == np.complex64:
        return np.float32(L.oracle_cnorm(a.size, _ptr(a.view(np.float32))))
    if a.dtype == np.complex128:
        return np.float64(L.oracle_znorm(a.size, _ptr(a.view(np.float64))))
    if a.dtype == np.float32:
        return np.float32(L.oracle_snorm(a.size, _ptr(a)))
    a = _c(a, np.float64)
    return np.float64(L.oracle_dnorm(a.size, _ptr(a)))


def normalize(a, fast: bool = False):
    """Rust::normalize (rust.rs:123-126, 141-146); returns a new array."""
    a = np.array(a, copy=True, order="C")
    L = lib(fast)
    if a.dtype == np.complex64:
        L.oracle_cnormalize(a.size, _ptr(a.view(np.float32)))
    elif a.dtype == np.complex128:
        L.oracle_znormalize(a.size, _ptr(a.view(np.float64)))
    elif a.dtype == np.float32:
        L.oracle_snormalize(a.size, _ptr(a))
    else:
        a = a.astype(np.float64)
        L.oracle_dnormalize(a.size, _ptr(a))
    return a


def transpose(a, columns: int):
    """Rust::transpose, src/backends/rust.rs:162-177 (any Copy element type)."""
    a = np.ascontiguousarray(a)
    out = np.zeros_like(a)
    lib().oracle_transpose(a.size, a.itemsize, a.ctypes.data, out.ctypes.data, columns)
    return out


def gemm(a, b, m, n, k, lda, ldb, ldc, a_trans, b_trans, f64acc: bool = False):
    """Blas::matrix_mul, src/backends/blas.rs:67-111.  Returns the ldc-strided C buffer (m*ldc)."""
    a = np.asarray(a)
    dt = _real_dtype(a)
    a, b = _c(a, dt).ravel(), _c(b, dt).ravel()
    c = np.zeros(m * ldc, dtype=dt)
    L = lib()
    if dt == np.float32:
        f = L.oracle_sgemm_f64acc if f64acc else L.oracle_sgemm
    else:
        f = L.oracle_dgemm
    f(_ptr(a), _ptr(b), _ptr(c), m, n, k, lda, ldb, ldc, int(a_trans), int(b_trans))
    return c


def gemm_batched(a, b, batch, m, n, k, a_trans=False, b_trans=False, fast: bool = False):
    """A loop of single matrix_mul calls over [[T; LEN]; batch] (what a slas user writes today)."""
    a = np.asarray(a)
    dt = _real_dtype(a)
    a, b = _c(a, dt).ravel(), _c(b, dt).ravel()
    c = np.zeros(batch * m * n, dtype=dt)
    L = lib(fast)
    f = L.oracle_sgemm_batched if dt == np.float32 else L.oracle_dgemm_batched
    f(batch, _ptr(a), _ptr(b), _ptr(c), m, n, k, int(a_trans), int(b_trans))
    return c


def gemv(a, x, m, n, lda, a_trans):
    """Blas::matrix_vector_mul with the trait's (m = width, n = height), blas.rs:114-151."""
    a = np.asarray(a)
    dt = _real_dtype(a)
    a, x = _c(a, dt).ravel(), _c(x, dt).ravel()
    y = np.zeros(m if a_trans else n, dtype=dt)
    (lib().oracle_sgemv if dt == np.float32 else lib().oracle_dgemv)(
        _ptr(a), _ptr(x), _ptr(y), m, n, lda, int(a_trans))
    return y


def tensor_index(shape, index) -> int:
    """tensor_index, src/tensor.rs:202-219; raises IndexError where the reference asserts."""
    nd = len(shape)
    S = (_sz * nd)(*shape)
    I = (_sz * nd)(*index)
    r = lib().oracle_tensor_index(nd, S, I)
    if r == _sz(-1).value:
        raise IndexError(f"Index {list(index)} out of bounds {list(shape)}")
    return int(r)


def matrix_index(shape, is_trans: bool, r: int, c: int) -> int:
    """Matrix::index, src/tensor.rs:296-301 (shape = [axis0 (width), axis1 (rows)])."""
    S = (_sz * 2)(*shape)
    v = lib().oracle_matrix_index(S, int(is_trans), r, c)
    if v == _sz(-1).value:
        raise IndexError(f"Index ({r}, {c}) out of bounds {list(shape)} trans={is_trans}")
    return int(v)


def matmul_args(shape_a, a_trans, shape_b, b_trans):
    """(m, n, k, lda, ldb, ldc) as Matrix::matrix_mul_buffer derives them, src/tensor.rs:376-382."""
    A = (_sz * 2)(*shape_a)
    B = (_sz * 2)(*shape_b)
    out = (_sz * 6)()
    lib().oracle_matmul_args(A, int(a_trans), B, int(b_trans), out)
    return tuple(int(v) for v in out)
