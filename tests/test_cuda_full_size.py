"""Parity at BASELINE.json's FULL sizes (`pytest -m gpu`).  The oracle cannot chew 2^24 products or
2^32 elements in seconds, so each case combines (i) objects / chunks sampled at random and checked
against the oracle exactly as in test_cuda_parity.py with (ii) size-independent properties computed on
the device: linearity of the product in B, transpose(transpose(A)) == A bit-for-bit, a sum of chunk
dots, Freivalds' probe for the 4096^2 product, and analytic values for the 2^32-element vectors."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle
from slas_b200 import _lib
from slas_b200.backends import Cuda, device_count, set_sgemm_path
from slas_b200.static_vec import CudaBatch, CudaVec

pytestmark = pytest.mark.gpu
cuda = Cuda()
TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-13}


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert device_count() > 0, "the gpu-marked tests need a CUDA device (there is no CPU fallback)"


def uniform(n, dt, seed, lo=-1.0, hi=1.0, offset=0):
    v = CudaVec(n, dt)
    _lib.call("slas_cuda_sfill_uniform" if np.dtype(dt) == np.float32 else "slas_cuda_dfill_uniform", n, v.as_ptr(), seed,
              offset, lo, hi)
    return v


def device_norm(v):
    return float(cuda.norm(v))


@pytest.mark.parametrize("size,dt,log2", [(4, np.float32, 24), (16, np.float32, 20), (32, np.float32, 20),
                                          (16, np.float64, 20), (32, np.float64, 20)])
def test_batched_matmul_full_batch(size, dt, log2):
    batch, e = 1 << log2, size * size
    a, b1, b2 = uniform(batch * e, dt, 11), uniform(batch * e, dt, 12), uniform(batch * e, dt, 13)
    c1, c2, c12 = CudaVec(batch * e, dt), CudaVec(batch * e, dt), CudaVec(batch * e, dt)
    cuda.matrix_mul_batched(a, b1, c1, size, size, size)
    # (i) sampled objects against the oracle
    rng = np.random.default_rng(size + log2)
    tol = TOL[np.dtype(dt)] * np.sqrt(size)
    for i in [0, batch - 1] + [int(x) for x in rng.integers(0, batch, 200)]:
        A, B, got = a.view(i * e, e).to_host(), b1.view(i * e, e).to_host(), c1.view(i * e, e).to_host()
        want = oracle.gemm(A, B, size, size, size, size, size, size, False, False, f64acc=(dt == np.float32))
        mag = (np.abs(A.reshape(size, size)).astype(np.float64) @ np.abs(B.reshape(size, size)).astype(np.float64)).ravel()
        ref = (A.reshape(size, size).astype(np.float64) @ B.reshape(size, size).astype(np.float64)).ravel()
        assert np.all(np.abs(got - ref) <= tol * mag + 1e-300), f"object {i}"
        assert np.max(np.abs(got - want)) <= 2 * tol * size
    # (ii) linearity over the WHOLE batch: A.(B1 + B2) == A.B1 + A.B2 up to rounding
    cuda.matrix_mul_batched(a, b2, c2, size, size, size)
    cuda.add(b1, b2, b2)            # b2 <- b1 + b2 (element-wise, bit-exact op)
    cuda.matrix_mul_batched(a, b2, c12, size, size, size)
    cuda.add(c1, c2, c2)            # c2 <- A.B1 + A.B2
    cuda.sub(c12, c2, c1)           # c1 <- difference
    rel = device_norm(c1) / device_norm(c12)
    assert rel <= 8 * np.finfo(dt).eps * np.sqrt(size), f"linearity residual {rel:.3e}"


@pytest.mark.parametrize("size,dt", [(16, np.float32), (32, np.float32), (16, np.float64), (32, np.float64), (4, np.float32)])
def test_batched_transpose_full_batch(size, dt):
    batch, e = 1 << 20, size * size
    a = uniform(batch * e, dt, 21)
    t, back = CudaVec(batch * e, dt), CudaVec(batch * e, dt)
    cuda.transpose_batched(a, t, e, size)
    rng = np.random.default_rng(size)
    for i in [0, batch - 1] + [int(x) for x in rng.integers(0, batch, 100)]:
        assert t.view(i * e, e).to_host().tobytes() == oracle.transpose(a.view(i * e, e).to_host(), size).tobytes()
    cuda.transpose_batched(t, back, e, size)     # idempotence: transposing twice is the identity, bit for bit
    cuda.sub(back, a, t)
    assert device_norm(t) == 0.0


def test_dot_add_len_2_28():
    n = 1 << 28
    a, b = uniform(n, np.float32, 31, 0.0, 1.0), uniform(n, np.float32, 32, -1.0, 1.0)
    whole = float(cuda.dot(a, b))
    # a checksum of checksums: the sum of 16 chunk dots equals the whole dot up to rounding
    chunk = n // 16
    parts = [float(cuda.dot(a.view(i * chunk, chunk), b.view(i * chunk, chunk))) for i in range(16)]
    scale = n / 4  # E|a*b| = 1/4
    assert abs(whole - sum(parts)) <= 1e-5 * np.sqrt(n) * scale
    # one chunk against the oracle (both lane variants of the reference) and an f64 truth
    ha, hb = a.view(0, 1 << 22).to_host(), b.view(0, 1 << 22).to_host()
    got = float(cuda.dot(a.view(0, 1 << 22), b.view(0, 1 << 22)))
    truth = float(np.dot(ha.astype(np.float64), hb.astype(np.float64)))
    bound = 1e-5 * np.sqrt(1 << 22) * float(np.sum(np.abs(ha.astype(np.float64) * hb)))
    assert abs(got - truth) <= bound
    for lanes in (4, 8):
        assert abs(got - float(oracle.dot(ha, hb, lanes=lanes))) <= 2 * bound
    # add: bit-exact on sampled windows of the full-size result
    c = CudaVec(n, np.float32)
    cuda.add(a, b, c)
    for off in (0, n // 3, n - (1 << 16)):
        w = 1 << 16
        assert np.array_equal(c.view(off, w).to_host(), a.view(off, w).to_host() + b.view(off, w).to_host())
    # norm / normalize at full size
    nrm = float(cuda.norm(a))
    assert abs(nrm - np.sqrt(n / 3)) <= 1e-3 * np.sqrt(n / 3)   # E[u^2] = 1/3
    before = a.view(12345, 4096).to_host()
    cuda.normalize(a)
    assert np.array_equal(a.view(12345, 4096).to_host(), before / np.float32(nrm))  # true division by the norm
    assert abs(float(cuda.norm(a)) - 1.0) <= 1e-5


def test_host_pointer_pipeline_crosses_chunk_boundaries():
    """The reference's calling convention is HOST pointers -- pageable ones (numpy arrays here, [T; LEN] / Vec there):
    large host operands go through the 3-slot H2D / kernel / D2H pipeline in chunks of ~64 MB, staged through the pinned
    bounce ring by the copy threads.  Every byte must land where the one-shot device path puts it, including ragged
    final chunks (bit-exact: same kernels, same per-object arithmetic)."""
    rng = np.random.default_rng(77)
    # element-wise: 3 operands x 4 B -> ~5.6M-element chunks; 20M + 5 elements = 4 chunks, ragged tail
    n = 20_000_005
    a, b = rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32) + 0.5
    c = np.empty_like(a)
    cuda.div(a, b, c)
    assert np.array_equal(c, a / b)
    # batched 4x4 product: 192 B per pair -> ~350K-pair chunks; 1.3M + 3 pairs = 4 chunks
    batch = 1_300_003
    A = rng.integers(-4, 5, batch * 16).astype(np.float32)
    B = rng.integers(-4, 5, batch * 16).astype(np.float32)
    C = np.empty(batch * 16, np.float32)
    cuda.matrix_mul_batched(A, B, C, 4, 4, 4)
    assert np.array_equal(C.reshape(batch, 4, 4), A.reshape(batch, 4, 4) @ B.reshape(batch, 4, 4))
    # batched transpose and batched dot through the same pipeline
    T16 = rng.random(300_000 * 256, dtype=np.float32)
    out = np.empty_like(T16)
    cuda.transpose_batched(T16, out, 256, 16)
    assert np.array_equal(out.reshape(-1, 16, 16), T16.reshape(-1, 16, 16).transpose(0, 2, 1))
    va = rng.integers(-3, 4, 1_000_003 * 32).astype(np.float32)
    vb = rng.integers(-3, 4, 1_000_003 * 32).astype(np.float32)
    d = np.empty(1_000_003, np.float32)
    cuda.dot_batched(va, vb, d, 32)
    assert np.array_equal(d, np.einsum("ij,ij->i", va.reshape(-1, 32), vb.reshape(-1, 32)))
    # a whole-vector reduction and an in-place normalize staged from the host
    x = rng.integers(-3, 4, 30_000_001).astype(np.float32)
    # every product and partial sum is an exact integer, the f64 grid reduction too: the result is the correctly
    # rounded f32 of the exact sum
    assert float(cuda.dot(x, x)) == float(np.float32(np.dot(x.astype(np.float64), x.astype(np.float64))))
    y = x[:5_000_003].copy()
    nrm = cuda.norm(y)
    cuda.normalize(y)
    assert np.array_equal(y, x[:5_000_003] / nrm)


@pytest.mark.parametrize("path", ["ffma", "tcgen05"])
def test_matmul_4096_freivalds(path):
    n = 4096
    a, b, c = uniform(n * n, np.float32, 41), uniform(n * n, np.float32, 42), CudaVec(n * n, np.float32)
    set_sgemm_path(path)
    try:
        cuda.matrix_mul(a, b, c, n, n, n, n, n, n, False, False)
    finally:
        set_sgemm_path("auto")
    A, B, Cm = (v.to_host().reshape(n, n).astype(np.float64) for v in (a, b, c))
    rng = np.random.default_rng(4)
    for _ in range(3):  # Freivalds: C x == A (B x) for random probes, in f64 on the host
        x = rng.standard_normal(n)
        lhs, rhs = Cm @ x, A @ (B @ x)
        scale = np.abs(A) @ (np.abs(B) @ np.abs(x))
        assert np.all(np.abs(lhs - rhs) <= 1e-5 * np.sqrt(n) * scale)
    # a few entries against the f64-accumulated oracle product
    for (i, j) in [(0, 0), (n - 1, n - 1), (17, 4000), (2048, 1)]:
        ref = float(A[i] @ B[:, j])
        mag = float(np.abs(A[i]) @ np.abs(B[:, j]))
        assert abs(Cm[i, j] - ref) <= 1e-5 * np.sqrt(n) * mag


def test_dot_normalize_2_32_single_gpu():
    """Config 5 on one rank: the whole 2^32-element pair (32 GiB) is resident in one B200's HBM."""
    n = 1 << 32
    a, b = uniform(n, np.float32, 4, 0.0, 1.0), uniform(n, np.float32, 5, 0.0, 1.0)
    d = float(cuda.dot_sharded(a, b))          # no communicator: the local partial is the sum
    assert abs(d / n - 0.25) < 2e-4             # E[u v] = 1/4; the sample mean over 2^32 terms is far tighter
    # the same bits as the sum over 4 slices cut by the library's own shard arithmetic, up to rounding
    from slas_b200 import sharding
    parts = []
    for r in range(4):
        lo, hi = sharding.shard_range(n, 4, 4, r)
        out = np.zeros(1, np.float64)
        _lib.call("slas_cuda_sdot_partial", hi - lo, a.as_ptr() + 4 * lo, b.as_ptr() + 4 * lo, out.ctypes.data)
        parts.append(float(out[0]))
    assert abs(sum(parts) - d) <= 1e-6 * d      # f64 partials: slice order does not matter at f32 resolution
    cuda.normalize_sharded(a)
    assert abs(float(cuda.norm_sharded(a)) - 1.0) < 1e-5
