"""One launch (after one warm-up) of each hot-path kernel at its BASELINE size, for ncu.
Usage:  python profiles/prof_targets.py [names...]     (default: all)
Never a source of bench numbers: anything timed under a profiler is not a measurement."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slas_b200 import _lib  # noqa: E402
from slas_b200.backends import Cuda, set_sgemm_path, synchronize  # noqa: E402
from slas_b200.static_vec import CudaVec  # noqa: E402

cuda = Cuda()


def uniform(n, dt, seed, lo=-1.0, hi=1.0):
    v = CudaVec(n, dt)
    _lib.call("slas_cuda_sfill_uniform" if dt == np.float32 else "slas_cuda_dfill_uniform", n, v.as_ptr(), seed, 0, lo, hi)
    return v


def gemm_batched(size, dt, log2):
    b = 1 << log2
    a, bb, c = uniform(b * size * size, dt, 1), uniform(b * size * size, dt, 2), CudaVec(b * size * size, dt)
    for _ in range(2):
        cuda.matrix_mul_batched(a, bb, c, size, size, size)
    synchronize()


def transpose_batched(size, dt, log2):
    b = 1 << log2
    a, c = uniform(b * size * size, dt, 1), CudaVec(b * size * size, dt)
    for _ in range(2):
        cuda.transpose_batched(a, c, size * size, size)
    synchronize()


def transpose_single(side, dt):
    a, c = CudaVec(side * side, dt).zero_(), CudaVec(side * side, dt)
    for _ in range(2):
        cuda.transpose(a, c, side)
    synchronize()


def small_odd():
    """3-D geometry sizes and a run-time specialised shape: the bulk-TMA staged thread-per-object kernels"""
    b = 1 << 24
    a, x, y = uniform(b * 9, np.float32, 1), uniform(b * 3, np.float32, 2), CudaVec(b * 3, np.float32)
    c, o = CudaVec(b * 9, np.float32), CudaVec(b, np.float32)
    for _ in range(2):
        cuda.matrix_mul_batched(a, a, c, 3, 3, 3)
        cuda.matrix_vector_mul_batched(a, x, y, 3, 3, False)
        cuda.transpose_batched(a, c, 9, 3)
        cuda.dot_batched(x, y, o, 3)
        cuda.normalize_batched(x, 3)
    b2 = 1 << 22
    p, q, r = uniform(b2 * 128, np.float32, 3), uniform(b2 * 64, np.float32, 4), CudaVec(b2 * 32, np.float32)
    for _ in range(2):
        cuda.matrix_mul_batched(p, q, r, 8, 4, 16)
    synchronize()


def vector_ops():
    n = 1 << 28
    a, b, c = uniform(n, np.float32, 3, 0, 1), uniform(n, np.float32, 4, 0, 1), CudaVec(n, np.float32)
    out = CudaVec(4, np.float32)
    for _ in range(2):
        _lib.call("slas_cuda_sdot", n, a.as_ptr(), b.as_ptr(), out.as_ptr())
        cuda.add(a, b, c)
        cuda.normalize(c)
    synchronize()


def vector_small():
    """BASELINE configs[0] as written: single dot / add calls at LEN 2^20 on distinct operands (balanced one-wave
    geometry, flagged-word combine, dependent launch)"""
    n, calls = 1 << 20, 24
    a, b, c = uniform(calls * n, np.float32, 3, 0, 1), uniform(calls * n, np.float32, 4, 0, 1), CudaVec(calls * n, np.float32)
    out = CudaVec(calls, np.float32)
    for i in range(calls):
        _lib.call("slas_cuda_sdot", n, a.as_ptr() + 4 * i * n, b.as_ptr() + 4 * i * n, out.as_ptr() + 4 * i)
    for i in range(calls):
        _lib.call("slas_cuda_sadd", n, a.as_ptr() + 4 * i * n, b.as_ptr() + 4 * i * n, c.as_ptr() + 4 * i * n)
    synchronize()


def gemm_large(path):
    n = 4096
    a, b, c = uniform(n * n, np.float32, 5), uniform(n * n, np.float32, 6), CudaVec(n * n, np.float32)
    set_sgemm_path(path)
    for _ in range(2):
        cuda.matrix_mul(a, b, c, n, n, n, n, n, n, False, False)
    synchronize()
    set_sgemm_path("auto")


def odd_normalize():
    """batched normalize of vectors that are not whole 16-byte words: LEN 3 (built-in constant) and LEN 30 (NVRTC)"""
    for length, log2 in ((3, 24), (30, 21)):
        b = 1 << log2
        x = uniform(b * length, np.float32, 2)
        for _ in range(2):
            cuda.normalize_batched(x, length)
    synchronize()


def odd_final():
    """the shipped small-object kernels: 3x3 f32 / f64 transposes (a thread per object, 4 KB stages), 3-vector dot / norm /
    normalize, LEN 30 normalize (NVRTC), batched normalize of 8192- and 65536-element vectors (CTA / cluster per vector)"""
    b = 1 << 24
    a, c = uniform(b * 9, np.float32, 1), CudaVec(b * 9, np.float32)
    x, y, o = uniform(b * 3, np.float32, 2), uniform(b * 3, np.float32, 3), CudaVec(b, np.float32)
    ad, cd = uniform(b // 2 * 9, np.float64, 4), CudaVec(b // 2 * 9, np.float64)
    x30 = uniform((1 << 21) * 30, np.float32, 5)
    m8, m64 = uniform((1 << 15) * 8192, np.float32, 6), uniform((1 << 12) * 65536, np.float32, 7)
    for _ in range(2):
        cuda.transpose_batched(a, c, 9, 3)
        cuda.transpose_batched(ad, cd, 9, 3)
        cuda.dot_batched(x, y, o, 3)
        cuda.norm_batched(x, o, 3)
        cuda.normalize_batched(x, 3)
        cuda.normalize_batched(x30, 30)
        cuda.normalize_batched(m8, 8192)
        cuda.normalize_batched(m64, 65536)
    synchronize()


TARGETS = {
    "odd_final": odd_final,
    "odd_normalize": odd_normalize,
    "gemm4_f32": lambda: gemm_batched(4, np.float32, 24),
    "gemm16_f32": lambda: gemm_batched(16, np.float32, 20),
    "gemm4_f64": lambda: gemm_batched(4, np.float64, 23),
    "gemm32_f32": lambda: gemm_batched(32, np.float32, 20),
    "gemm16_f64": lambda: gemm_batched(16, np.float64, 20),
    "gemm32_f64": lambda: gemm_batched(32, np.float64, 20),
    "tr16_f32": lambda: transpose_batched(16, np.float32, 20),
    "tr32_f32": lambda: transpose_batched(32, np.float32, 20),
    "tr32_f64": lambda: transpose_batched(32, np.float64, 20),
    "tr_u8": lambda: transpose_single(8192, np.uint8),
    "tr_i16": lambda: transpose_single(8192, np.int16),
    "small_odd": small_odd,
    "vector": vector_ops,
    "vector_small": vector_small,
    "gemm4096_ffma": lambda: gemm_large("ffma"),
    "gemm4096_tc": lambda: gemm_large("tcgen05"),
}

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(TARGETS)):
        TARGETS[name]()
        print("ran", name, flush=True)
