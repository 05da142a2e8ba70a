"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Every op is called through the C ABI
(slas_b200.backends.Cuda -> libslas_cuda.so) and compared with the CPU oracle on the same seeded
inputs, against the reference's golden vectors, and -- at BASELINE.json's full sizes -- through
sampled objects plus size-independent properties.

Bars (BASELINE.json north_star): bit-exact for element-wise ops, transposes and anything that is
index/shape work; floating-point reductions within  f32: 1e-5 * sqrt(K),  f64: 1e-13 * sqrt(K)
relative (K = reduction length), taken relative to sum|a_i*b_i| so that cancelling inputs are
judged on the same scale the rounding errors live on."""
import numpy as np
import pytest

from oracle import oracle
from slas_b200 import tensor as T
from slas_b200.backends import Cuda, device_count, launch_count
from slas_b200.static_vec import CudaBatch, CudaVec, moo

pytestmark = pytest.mark.gpu

DT = {"f32": np.float32, "f64": np.float64, "c64": np.complex64}
TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-13}
cuda = Cuda()


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert device_count() > 0, "the gpu-marked tests need a CUDA device (there is no CPU fallback)"
    before = launch_count()
    yield
    assert launch_count() > before, "no kernel of libslas_cuda was launched"
    if MARGINS:
        print("\nobserved max error / bar per family (bar: f32 1e-5 * sqrt(K), f64 1e-13 * sqrt(K), relative to sum|a_i b_i|):")
        for k_, v_ in sorted(MARGINS.items()):
            print(f"  {v_:10.3e}  {k_}")


def _cplx(pairs, dt):
    return np.array([complex(r, i) for r, i in pairs], dtype=dt)


def _rand(rng, n, dt, lo=-1.0, hi=1.0):
    return (rng.random(n) * (hi - lo) + lo).astype(dt)


def _both(x):
    """The same operand as a host vector and as a device-resident CudaVec."""
    return [np.array(x, copy=True), CudaVec.from_host(x)]


def _host(x):
    return x.to_host() if isinstance(x, CudaVec) else x


# ---- golden vectors of the reference's own tests, through the reference-shaped API ---------------
def test_dot_golden(golden):
    for g in golden["dot"]:
        dt = DT[g["dtype"]]
        for a, b in zip(_both(np.array(g["a"], dt)), _both(np.array(g["b"], dt))):
            r = moo(dt, g["a"], on=Cuda).dot(moo(dt, g["b"], on=Cuda)) if isinstance(a, np.ndarray) else cuda.dot(a, b)
            if "abs_tol" in g:
                assert abs(float(r) - g["expect"]) < g["abs_tol"], g["cite"]
            else:
                assert float(r) == g["expect"], g["cite"]


def test_complex_golden(golden):
    for g in golden["dotu"]:
        a, b = _cplx(g["a"], np.complex64), _cplx(g["b"], np.complex64)
        assert complex(cuda.dot(a, b)) == complex(*g["expect"]), g["cite"]
        assert complex(cuda.dot(CudaVec.from_host(a), CudaVec.from_host(b))) == complex(*g["expect"])
    for g in golden["norm"]:
        a = _cplx(g["a"], np.complex64)
        assert cuda.norm(a) == np.float32(g["expect"]), g["cite"]
        assert cuda.norm(a.astype(np.complex128)) == pytest.approx(g["expect"], rel=1e-7)


def test_ewise_golden(golden):
    for g in golden["ewise"]:
        dt = DT[g["dtype"]]
        a, b = np.array(g["a"], dt), np.array(g["b"], dt)
        want = oracle.ewise(g["op"], a, b)
        c = np.zeros_like(a)
        getattr(cuda, g["op"])(a, b, c)
        assert np.array_equal(c, want), g["cite"]
        dc = CudaVec(a.size, dt)
        getattr(cuda, g["op"])(CudaVec.from_host(a), CudaVec.from_host(b), dc)
        assert np.array_equal(dc.to_host(), want), g["cite"]


def test_transpose_golden(golden):
    for g in golden["transpose"]:
        for dt in (np.float32, np.float64, np.complex128):
            a = np.array(g["a"], dtype=dt)
            want = np.array(g["expect"], dtype=dt)
            buf = np.zeros_like(a)
            cuda.transpose(a, buf, g["columns"])
            assert np.array_equal(buf, want), g["cite"]
            d = CudaVec.from_host(a)
            cuda.transpose_inplace(d, g["columns"])
            assert np.array_equal(d.to_host(), want), g["cite"]
    g = [t for t in golden["transpose"] if "debug" in t][0]
    m = T.Matrix(T.Tensor(np.array(g["a"], np.float32), [3, 2], cuda), True)
    t = m.deref_mut()  # materialise the lazy transpose on the GPU
    assert list(t.data.data) == g["expect"] and t.shape.slice() == [2, 3]


def test_matmul_golden(golden):
    for g in golden["matmul"]:
        for dt in (np.float32, np.float64):
            a, b = g["a"], g["b"]
            for da, db in zip(_both(np.array(a["data"], dt)), _both(np.array(b["data"], dt))):
                A = T.Matrix(T.Tensor(da, a["shape"], cuda), a["trans"])
                B = T.Matrix(T.Tensor(db, b["shape"], cuda), b["trans"])
                got = _host(A.matrix_mul(B))
                assert np.array_equal(got, np.array(g["expect"], dt)), g["cite"]


def test_gemv_golden(golden):
    for g in golden["gemv"]:
        for dt in (np.float32, np.float64):
            a = g["a"]
            for da, dx in zip(_both(np.array(a["data"], dt)), _both(np.array(g["x"], dt))):
                A = T.Matrix(T.Tensor(da, a["shape"], cuda), a["trans"])
                assert np.array_equal(_host(A.vector_mul(dx)), np.array(g["expect"], dt)), g["cite"]
    # gemv == gemm with n = 1 (tests/src/main.rs:355-394)
    a = np.arange(2, 14, dtype=np.float32)
    x = np.array([2, 3, 4, 5], np.float32)
    A = T.matrix(a, 4, 3, cuda).transpose()
    assert np.array_equal(A.vector_mul(x), A.matrix_mul(T.matrix(x, 4, 1, cuda)))


# ---- element-wise: bit-exact, host and device operands, ragged lengths, unaligned views ------------
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 3, 4, 5, 12, 1000, 4097, (1 << 20) + 3])
def test_ewise_bit_exact(dt, n):
    rng = np.random.default_rng(n)
    a, b = _rand(rng, n, dt, -4, 4), _rand(rng, n, dt, 0.25, 4)
    da, db, dc = CudaVec.from_host(a), CudaVec.from_host(b), CudaVec(n, dt)
    for op in ("add", "sub", "mul", "div"):
        want = oracle.ewise(op, a, b)
        getattr(cuda, op)(da, db, dc)
        assert np.array_equal(dc.to_host(), want), op
        c = np.empty_like(a)
        getattr(cuda, op)(a, b, c)
        assert np.array_equal(c, want), op
    if n > 8:  # device pointers that are not 16-byte aligned take the scalar kernel
        cuda.div(da.view(1, n - 1), db.view(1, n - 1), dc.view(1, n - 1))
        assert np.array_equal(dc.to_host()[1:], oracle.ewise("div", a[1:], b[1:]))


def test_ewise_special_values_bit_exact():
    a = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, 3.4e38, 1.0, 5e-324, 1.0], np.float32)
    b = np.array([0.0, 0.0, np.inf, 1.0, 1.0, 3.0, 3.4e38, 3.0, 2.0, 0.0], np.float32)
    with np.errstate(all="ignore"):
        for op in ("add", "sub", "mul", "div"):
            c = np.empty_like(a)
            getattr(cuda, op)(a, b, c)
            want = oracle.ewise(op, a, b)
            # every non-NaN result is bit-identical (denormals, signed zeros, infinities included); a NaN
            # result is a NaN in the same place.  NaN PAYLOADS are not part of the contract: the
            # reference's own payload depends on the host ISA and on operand order after codegen (x86 SSE
            # propagates the first operand's payload, the SM's FADD/FMUL return the canonical 0x7fffffff).
            nan = np.isnan(want)
            assert np.array_equal(np.isnan(c), nan), op
            assert c[~nan].tobytes() == want[~nan].tobytes(), op


def test_empty_vectors():
    e = np.zeros(0, np.float32)
    cuda.add(e, e, np.zeros(0, np.float32))
    assert cuda.dot(e, e) == 0.0 and cuda.norm(e) == 0.0  # the reference's empty sum
    cuda.normalize(np.zeros(0, np.float64))


def test_bad_arguments_of_the_newer_entry_points():
    """fused chains, the batched in-place transpose and the graph API reject what they cannot do and leave the library usable"""
    import ctypes as C
    from slas_b200 import _lib
    L = _lib.lib()
    a = np.arange(8, dtype=np.float32)
    d = CudaVec.from_host(a)
    out = np.zeros(3, np.float32)
    p = lambda x: x.ctypes.data  # noqa: E731
    assert L.slas_cuda_sadd_dot(8, p(a), d.as_ptr(), None, p(a), p(out)) == _lib.SLAS_ERR_INVALID      # host / device mixed
    assert L.slas_cuda_sadd_dot(8, p(a), p(a), None, p(a), None) == _lib.SLAS_ERR_INVALID             # no result pointer
    assert L.slas_cuda_sdot_norms(8, p(a), None, p(out)) == _lib.SLAS_ERR_INVALID                      # null operand
    assert L.slas_cuda_snormalize_into(8, p(a), p(a)) == _lib.SLAS_ERR_INVALID                         # aliasing: use normalize
    assert L.slas_cuda_transpose_inplace_batched(2, 4, 0, p(a), 2) == _lib.SLAS_ERR_INVALID           # element size 0 (any other size is a valid Copy type)
    assert L.slas_cuda_transpose_inplace_batched(2, 4, 4, p(a), 0) == _lib.SLAS_ERR_INVALID           # columns == 0
    assert L.slas_cuda_transpose_inplace_batched(0, 4, 4, p(a), 2) == _lib.SLAS_OK                     # empty batch: nothing to do
    g = C.c_void_p()
    assert L.slas_cuda_graph_end(C.byref(g)) != _lib.SLAS_OK                                           # no capture in progress
    assert L.slas_cuda_graph_launch(None) == _lib.SLAS_ERR_INVALID
    assert L.slas_cuda_graph_destroy(None) == _lib.SLAS_OK
    # ... and everything still works
    assert cuda.ewise_dot("add", a, a, a) == 2 * float(np.sum(a * a))
    assert cuda.dot_norms(a, a)[0] == 140.0
    h = a.copy()
    cuda.transpose_inplace_batched(h, 4, 2)
    assert list(h) == [0, 2, 1, 3, 4, 6, 5, 7]


def test_bad_arguments_fail_loudly_not_fatally():
    """Where the reference asserts or cblas rejects its arguments, the C ABI returns a status (the shim panics)."""
    import ctypes as C
    from slas_b200 import _lib
    L = _lib.lib()
    a = np.arange(6, dtype=np.float32)
    d = CudaVec.from_host(a)
    out = np.zeros(1, np.float32)
    p = lambda x: x.ctypes.data  # noqa: E731
    assert L.slas_cuda_sdot(6, None, p(a), p(out)) == _lib.SLAS_ERR_INVALID              # null operand
    assert L.slas_cuda_sdot(6, p(a), d.as_ptr(), p(out)) == _lib.SLAS_ERR_INVALID         # host and device operands mixed
    assert b"mix" in L.slas_cuda_last_error()
    assert L.slas_cuda_sdot(6, p(a), p(a), None) == _lib.SLAS_ERR_INVALID                  # no result pointer
    assert L.slas_cuda_transpose(6, 4, p(a), p(a), 0) == _lib.SLAS_ERR_INVALID             # columns == 0 (rust.rs:169 divides)
    assert L.slas_cuda_transpose(6, 0, p(a), p(out), 2) == _lib.SLAS_ERR_INVALID           # element size 0 (every other size is a valid Copy type)
    c = np.zeros(4, np.float32)
    assert L.slas_cuda_sgemm(p(a), p(a), p(c), 2, 2, 3, 2, 2, 2, 0, 0) == _lib.SLAS_ERR_INVALID   # lda < k
    assert L.slas_cuda_sgemm(p(a), p(a), p(c), 2, 2, 3, 3, 2, 1, 0, 0) == _lib.SLAS_ERR_INVALID   # ldc < n
    assert L.slas_cuda_sgemv(p(a), p(a), p(c), 3, 2, 2, 0) == _lib.SLAS_ERR_INVALID               # lda < width
    assert L.slas_cuda_set_tunable(b"no_such_knob", 1) == _lib.SLAS_ERR_INVALID
    assert L.slas_cuda_set_sgemm_path(7) == _lib.SLAS_ERR_INVALID
    # ... and the library keeps working afterwards
    assert cuda.dot(a, a) == 55.0
    flag = C.c_int(-1)
    _lib.call("slas_cuda_is_device_ptr", d.as_ptr(), C.byref(flag))
    assert flag.value == 1
    _lib.call("slas_cuda_is_device_ptr", p(a), C.byref(flag))
    assert flag.value == 0
    # k == 0: beta = 0 semantics, C becomes zeros (cblas with k = 0)
    z = np.full(4, 7.0, np.float32)
    cuda.matrix_mul(np.zeros(0, np.float32), np.zeros(0, np.float32), z, 2, 2, 0, 1, 2, 2, False, False)
    assert np.array_equal(z, np.zeros(4, np.float32))


def test_concurrent_callers_like_libtest():
    """libtest runs the reference's tests on several threads at once and a backend is a stateless ZST, so the
    library serialises enqueue + scratch internally (SURVEY 8b 'Threading'): 8 threads, mixed ops, exact answers."""
    import threading
    errors = []

    def worker(seed):
        try:
            rng = np.random.default_rng(seed)
            for _ in range(40):
                n = int(rng.integers(1, 5000))
                a = rng.integers(-8, 9, n).astype(np.float32)   # small integers: every result is exact
                b = rng.integers(-8, 9, n).astype(np.float32)
                assert float(cuda.dot(a, b)) == float(np.dot(a.astype(np.float64), b.astype(np.float64)))
                c = np.empty_like(a)
                cuda.add(a, b, c)
                assert np.array_equal(c, a + b)
                m = rng.integers(-3, 4, 16).astype(np.float32)
                got = T.matrix(m, 4, 4, cuda).matrix_mul(T.matrix(m, 4, 4, cuda))
                assert np.array_equal(got.reshape(4, 4), m.reshape(4, 4) @ m.reshape(4, 4))
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(s,)) for s in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]


# ---- dot / norm / normalize ------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 2, 7, 8, 100, 750, 4097, 65536 + 5, 1 << 20])
def test_dot_within_tolerance(dt, n):
    rng = np.random.default_rng(n + 1)
    a, b = _rand(rng, n, dt, 0, 1), _rand(rng, n, dt)
    truth = float(np.dot(a.astype(np.longdouble), b.astype(np.longdouble)))
    scale = float(np.sum(np.abs(a.astype(np.float64) * b.astype(np.float64)))) + 1e-300
    bound = TOL[np.dtype(dt)] * np.sqrt(n) * scale
    got_h = cuda.dot(a, b)
    got_d = cuda.dot(CudaVec.from_host(a), CudaVec.from_host(b))
    assert got_h == got_d  # deterministic: same bits from the host-staged and the device-resident path
    assert _margin(f"dot {np.dtype(dt).name} vs long-double truth", np.array([abs(float(got_d) - truth)]), np.array([bound])) <= 1.0
    for lanes in (oracle.default_lanes(dt), oracle.default_lanes(dt, avx=True)):  # the reference's two builds
        assert abs(float(got_d) - float(oracle.dot(a, b, lanes=lanes))) <= 2 * bound
    if n > 8:
        d = CudaVec.from_host(a)
        assert abs(float(cuda.dot(d.view(1, n - 1), CudaVec.from_host(b[1:]))) -
                   float(np.dot(a[1:].astype(np.float64), b[1:].astype(np.float64)))) <= bound


def test_dot_is_deterministic_run_to_run():
    rng = np.random.default_rng(5)
    a, b = CudaVec.from_host(_rand(rng, (1 << 22) + 17, np.float32)), CudaVec.from_host(_rand(rng, (1 << 22) + 17, np.float32))
    assert len({cuda.dot(a, b).tobytes() for _ in range(8)}) == 1


@pytest.mark.parametrize("dt", [np.complex64, np.complex128])
def test_complex_dotu_and_norm(dt):
    rng = np.random.default_rng(9)
    rdt = np.float32 if dt == np.complex64 else np.float64
    for n in (1, 5, 1000, 4099):
        a = (_rand(rng, n, rdt) + 1j * _rand(rng, n, rdt)).astype(dt)
        b = (_rand(rng, n, rdt) + 1j * _rand(rng, n, rdt)).astype(dt)
        truth = np.sum(a.astype(np.complex128) * b.astype(np.complex128))  # UNconjugated
        tol = TOL[np.dtype(rdt)] * np.sqrt(2 * n) * float(np.sum(np.abs(a) * np.abs(b)))
        got = cuda.dot(CudaVec.from_host(a), CudaVec.from_host(b))
        assert abs(complex(got) - complex(truth)) <= tol
        assert abs(complex(got) - complex(oracle.dotu(a, b))) <= 2 * tol
        nrm = cuda.norm(a)
        assert nrm.dtype == rdt  # NormOutput is the real type (src/backends/rust.rs:130)
        assert abs(float(nrm) - float(oracle.norm(a))) <= TOL[np.dtype(rdt)] * np.sqrt(2 * n) * float(nrm)
        h = a.copy()
        cuda.normalize(h)
        want = oracle.normalize(a)
        assert np.max(np.abs(h - want)) <= 4 * np.finfo(rdt).eps * np.max(np.abs(want))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 3, 1000, 4097, (1 << 20) + 1])
def test_norm_and_normalize(dt, n):
    rng = np.random.default_rng(n + 2)
    a = _rand(rng, n, dt)
    truth = float(np.sqrt(np.sum(a.astype(np.longdouble) ** 2)))
    rel = TOL[np.dtype(dt)] * np.sqrt(n)
    nrm = cuda.norm(CudaVec.from_host(a))
    assert _margin(f"norm {np.dtype(dt).name} vs f64 truth", np.array([abs(float(nrm) - truth)]), np.array([rel * truth])) <= 1.0
    assert abs(float(nrm) - float(oracle.norm(a))) <= 2 * rel * truth
    d = CudaVec.from_host(a)
    cuda.normalize(d)
    got = d.to_host()
    assert np.array_equal(got, a / nrm)  # normalize is a TRUE division by the backend's own norm: bit-exact
    want = oracle.normalize(a)
    assert np.max(np.abs(got - want)) <= (2 * rel + 2 * np.finfo(dt).eps) * np.max(np.abs(want))
    h = a.copy()
    cuda.normalize(h)  # in place through a host pointer
    assert np.array_equal(h, got)


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.complex64])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 1000, (1 << 18) + 7])
def test_views_at_odd_offsets(dt, n):
    """device views that start 1, 2 or 3 elements past a 16-byte boundary (static_slice at an odd offset): a shared
    misalignment is peeled (scalar head, vector body), different misalignments take the scalar loop -- same results"""
    rng = np.random.default_rng(n)
    pad = 4
    cplx = np.dtype(dt).kind == "c"
    mk = (lambda k: (_rand(rng, k, np.float32) + 1j * _rand(rng, k, np.float32)).astype(dt)) if cplx else (lambda k: _rand(rng, k, dt))
    A, B = mk(n + pad), mk(n + pad)
    dA, dB, dC = CudaVec.from_host(A), CudaVec.from_host(B), CudaVec(n + pad, dt)
    rel = (1e-5 if np.dtype(dt).itemsize in (4, 8) and dt != np.float64 else 1e-13) * np.sqrt(n)
    for oa, ob in ((1, 1), (2, 2), (3, 3), (0, 1), (2, 1)):
        a, b = A[oa:oa + n], B[ob:ob + n]
        got = cuda.dot(dA.view(oa, n), dB.view(ob, n))
        truth = np.sum(a.astype(np.complex128 if cplx else np.longdouble) * b)
        assert abs(complex(got) - complex(truth)) <= rel * float(np.sum(np.abs(a) * np.abs(b))) + 1e-300, (oa, ob)
        nrm = cuda.norm(dA.view(oa, n))
        assert abs(float(nrm) - float(np.sqrt(np.sum(np.abs(a.astype(np.complex128)) ** 2)))) <= 2 * rel * float(nrm) + 1e-300
        if not cplx:
            dC.zero_()
            cuda.add(dA.view(oa, n), dB.view(ob, n), dC.view(oa, n))
            c = dC.to_host()
            assert np.array_equal(c[oa:oa + n], a + b) and np.all(c[:oa] == 0) and np.all(c[oa + n:] == 0), (oa, ob)
            dN = CudaVec.from_host(A)
            cuda.normalize(dN.view(oa, n))
            h = dN.to_host()
            assert np.array_equal(h[oa:oa + n], a / nrm) and np.array_equal(h[:oa], A[:oa]) and np.array_equal(h[oa + n:], A[oa + n:])


def test_reductions_special_values_and_overflow_policy():
    """IEEE propagation through the reductions, the normalize of a zero vector, and overflow of the sum of squares:
    like the Rust backend's naive f32 sum (rust.rs:116-121) a norm overflows to infinity when squares or a thread's
    partial sum overflow f32; partial sums are COMBINED in f64, so where only the grand total would overflow the
    result can stay finite (the Blas backend's scaled nrm2 answer, blas.rs:156-162) -- never the other way round."""
    f = np.float32
    with np.errstate(all="ignore"):
        assert np.isinf(cuda.dot(np.array([np.inf, 1], f), np.array([1, 1], f)))
        assert np.isnan(cuda.dot(np.array([np.inf, -np.inf], f), np.array([1, 1], f)))
        assert np.isnan(cuda.dot(np.array([np.nan, 1], f), np.array([1, 1], f)))
        assert np.isnan(cuda.norm(np.array([1, np.nan, 2], f))) and np.isinf(cuda.norm(np.array([1, -np.inf], f)))
        z = np.zeros(5, f)
        cuda.normalize(z)  # 0 / 0, as the reference's `a[i] /= norm`
        assert np.all(np.isnan(z))
        one = np.array([3e38, 3e38, 0, 0], f)  # each square overflows f32 on its own: inf, as in the reference
        assert np.isinf(cuda.norm(one)) and np.isinf(oracle.norm(one))
        big = np.full(1 << 12, 1.5e19, f)  # squares are finite in f32 (2.25e38), four of them in one accumulator are not
        assert np.isinf(cuda.norm(big)) and np.isinf(oracle.norm(big))
        two = np.array([1.5e19, 1.5e19], f)  # 4.5e38 only in total: inf for the naive f32 sum, finite here (f64 combine)
        got = cuda.norm(two)
        assert np.isinf(oracle.norm(two)) and np.isfinite(got) and abs(float(got) - 1.5e19 * np.sqrt(2)) <= 1e-5 * 3e19


# ---- fused chains: same results as the op sequence they replace ---------------------------
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 3, 5, 1000, 4097, (1 << 20) + 3])
@pytest.mark.parametrize("op", ["add", "sub", "mul", "div"])
def test_fused_ewise_dot_matches_the_chain(dt, n, op):
    """`c = a op b` then `dot(c, d)`: the intermediate is the reference's bit-exact element-wise result
    (rust.rs:45-73), the dot is within the reduction bar AND carries the same bits as the unfused GPU chain"""
    rng = np.random.default_rng(n + len(op))
    a, b, d = _rand(rng, n, dt, -4, 4), _rand(rng, n, dt, 0.25, 4), _rand(rng, n, dt)
    want_c = oracle.ewise(op, a, b)
    truth = float(np.sum(want_c.astype(np.longdouble) * d.astype(np.longdouble)))
    mag = float(np.sum(np.abs(want_c.astype(np.longdouble) * d.astype(np.longdouble))))
    da, db, dd = CudaVec.from_host(a), CudaVec.from_host(b), CudaVec.from_host(d)
    tmp = CudaVec(n, dt)
    getattr(cuda, op)(da, db, tmp)
    unfused = cuda.dot(tmp, dd)
    dc = CudaVec(n, dt).zero_()
    got = cuda.ewise_dot(op, da, db, dd, dc)
    assert np.array_equal(dc.to_host(), want_c)
    assert got.tobytes() == unfused.tobytes()
    assert abs(float(got) - truth) <= TOL[np.dtype(dt)] * np.sqrt(n) * mag
    assert cuda.ewise_dot(op, da, db, dd).tobytes() == unfused.tobytes()  # intermediate never materialised
    hc = np.zeros_like(a)
    assert cuda.ewise_dot(op, a, b, d, hc).tobytes() == unfused.tobytes()  # host pointers
    assert np.array_equal(hc, want_c)
    if n > 8:  # unaligned device views: scalar path, within the bar
        g = cuda.ewise_dot(op, da.view(1, n - 1), db.view(1, n - 1), dd.view(1, n - 1))
        t1 = float(np.sum(want_c[1:].astype(np.longdouble) * d[1:].astype(np.longdouble)))
        assert abs(float(g) - t1) <= TOL[np.dtype(dt)] * np.sqrt(n) * mag


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 3, 1000, 4097, (1 << 20) + 1])
def test_fused_dot_norms_matches_the_three_ops(dt, n):
    rng = np.random.default_rng(n + 11)
    a, b = _rand(rng, n, dt), _rand(rng, n, dt)
    da, db = CudaVec.from_host(a), CudaVec.from_host(b)
    dot, na, nb = cuda.dot_norms(da, db)
    assert dot.tobytes() == cuda.dot(da, db).tobytes()
    assert na.tobytes() == cuda.norm(da).tobytes() and nb.tobytes() == cuda.norm(db).tobytes()
    rel = TOL[np.dtype(dt)] * np.sqrt(n)
    assert abs(float(na) - float(oracle.norm(a))) <= 2 * rel * float(oracle.norm(a))
    assert abs(float(dot) - float(oracle.dot(a, b))) <= 2 * rel * float(np.sum(np.abs(a.astype(np.float64) * b)))
    assert cuda.dot_norms(a, b)[0].tobytes() == dot.tobytes()  # host pointers


def test_fused_chains_of_empty_vectors():
    e = np.zeros(0, np.float32)
    assert cuda.ewise_dot("add", e, e, e) == 0.0
    assert cuda.dot_norms(e, e) == (0.0, 0.0, 0.0)
    cuda.normalize_into(e, np.zeros(0, np.float32))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 3, 1000, (1 << 20) + 1])
def test_normalize_into_is_normalize_of_a_copy(dt, n):
    a = _rand(np.random.default_rng(n + 5), n, dt)
    da = CudaVec.from_host(a)
    out = CudaVec(n, dt).zero_()
    cuda.normalize_into(da, out)
    ref = CudaVec.from_host(a)
    cuda.normalize(ref)
    assert np.array_equal(out.to_host(), ref.to_host())
    assert np.array_equal(da.to_host(), a)  # the source is untouched (the borrowed side of a StaticCowVec)
    h = np.zeros_like(a)
    cuda.normalize_into(a, h)
    assert np.array_equal(h, ref.to_host())


def test_static_cow_vec_on_the_device():
    """a borrowed StaticCowVec over device memory: reads never copy, normalize writes the owned side in one fused
    pass (normalize_into) and leaves the source untouched; other mutable uses copy device-to-device first"""
    from slas_b200.static_vec import StaticCowVec
    a = _rand(np.random.default_rng(9), 5000, np.float32)
    src = CudaVec.from_host(a)
    v = StaticCowVec(src)
    assert cuda.dot(v, src).tobytes() == cuda.dot(src, src).tobytes() and v.is_borrowed()
    cuda.normalize(v)
    ref = CudaVec.from_host(a)
    cuda.normalize(ref)
    assert v.is_owned() and np.array_equal(v.to_host(), ref.to_host()) and np.array_equal(src.to_host(), a)
    w = StaticCowVec(src)
    cuda.add(src, src, w)  # output operand: as_mut_ptr -> copy, then overwritten
    assert w.is_owned() and np.array_equal(w.to_host(), a + a) and np.array_equal(src.to_host(), a)
    h = StaticCowVec(a)  # host source
    cuda.normalize(h)
    assert h.is_owned() and np.array_equal(h.to_host(), ref.to_host()) and a[0] == src.to_host()[0]


def test_cuda_graph_records_and_replays_a_sequence():
    """a launch-bound chain of device-pointer ops recorded once and replayed: same bytes as the eager chain, also
    after the inputs change (the graph holds pointers, not values)"""
    from slas_b200 import Graph, _lib as L
    rng = np.random.default_rng(21)
    n, batch = 1 << 16, 2048
    a, b = _rand(rng, n, np.float32), _rand(rng, n, np.float32)
    A, B = _rand(rng, batch * 16, np.float32), _rand(rng, batch * 16, np.float32)
    da, db, dc, dres = CudaVec.from_host(a), CudaVec.from_host(b), CudaVec(n, np.float32), CudaVec(4, np.float32)
    dA, dB, dC, dT = CudaVec.from_host(A), CudaVec.from_host(B), CudaBatch(batch, 16, np.float32), CudaBatch(batch, 16, np.float32)

    def chain():
        cuda.add(da, db, dc)
        L.call("slas_cuda_sdot", n, dc.as_ptr(), db.as_ptr(), dres.as_ptr())          # device result pointer
        L.call("slas_cuda_snrm2", n, dc.as_ptr(), dres.as_ptr() + 4)
        cuda.matrix_mul_batched(dA, dB, dC, 4, 4, 4)
        cuda.transpose_batched(dC, dT, 16, 4)

    chain()  # eager: the expected bytes, and every scratch buffer at its final size
    want = (dc.to_host(), dres.to_host()[:2].copy(), dT.to_host().copy())
    for v in (dc, dres, dC, dT):
        v.zero_()
    n0 = launch_count()
    with Graph() as g:
        chain()
    assert np.all(dT.to_host() == 0), "ops inside the capture must be recorded, not executed"
    g.launch()
    got = (dc.to_host(), dres.to_host()[:2], dT.to_host())
    for w, x in zip(want, got):
        assert np.array_equal(w, x)
    da.copy_from_host(b)  # new inputs, same graph
    g.launch()
    assert np.array_equal(dc.to_host(), b + b)
    assert launch_count() > n0
    # a host result pointer cannot be captured: the call fails, the capture is closed, the library stays usable
    with pytest.raises(Exception):
        with Graph():
            cuda.dot(da, db)
    assert cuda.dot(da, db).tobytes() == cuda.dot(da, db).tobytes()


# ---- transpose: bit-exact for any Copy element --------------------------------------------------------
@pytest.mark.parametrize("dt", [np.float32, np.float64, np.complex128, np.int32, np.int16, np.uint8])
@pytest.mark.parametrize("rows,cols", [(1, 1), (1, 7), (7, 1), (2, 3), (4, 4), (16, 16), (32, 32), (5, 64), (33, 65), (100, 257), (1024, 48), (128, 64), (192, 320), (1024, 512)])
def test_transpose_bit_exact(dt, rows, cols):
    a = np.arange(rows * cols).astype(dt)  # int16 / uint8 wrap: still a distinguishing pattern per tile
    if dt == np.int32:
        a = a.view(np.float32)  # opaque 4-byte payloads, NaN patterns included
    if dt in (np.int16, np.uint8):
        a = (np.arange(rows * cols) * 2654435761 >> 7).astype(dt)
    want = oracle.transpose(a, cols)
    buf = CudaVec(a.size, a.dtype).zero_()
    cuda.transpose(CudaVec.from_host(a), buf, cols)
    assert buf.to_host().tobytes() == want.tobytes()
    h = np.zeros_like(a)
    cuda.transpose(a, h, cols)
    assert h.tobytes() == want.tobytes()
    d = CudaVec.from_host(a)
    cuda.transpose_inplace(d, cols)
    assert d.to_host().tobytes() == want.tobytes()


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.complex128, np.int16, np.uint8])
@pytest.mark.parametrize("rows,cols", [(4, 4), (2, 2), (8, 8), (12, 12), (16, 16), (20, 20), (33, 33), (48, 48), (60, 60), (64, 64),
                                       (100, 100), (128, 128), (1000, 1000), (1024, 1024), (2048, 64), (64, 2048), (300, 7)])
def test_transpose_inplace_kernels_bit_exact(dt, rows, cols):
    """thread-per-object, warp-per-object, diagonal tile-pair exchange (32- and 64-wide) and the temporary
    fallback for large non-square objects all give the reference's transpose_inplace (rust.rs:152-160)"""
    a = (np.arange(rows * cols) * 2654435761 >> 5).astype(dt)
    want = oracle.transpose(a, cols)
    d = CudaVec.from_host(a)
    cuda.transpose_inplace(d, cols)
    assert d.to_host().tobytes() == want.tobytes()
    h = a.copy()
    cuda.transpose_inplace(h, cols)  # host pointer
    assert h.tobytes() == want.tobytes()


@pytest.mark.parametrize("length,cols", [(11, 3), (1000, 7), (5000, 64), (70000, 257), (5, 9)])
def test_transpose_inplace_ragged_tail_is_zeroed(length, cols):
    """the reference transposes into a ZEROED temporary and copies all of it back (rust.rs:157-159): in place,
    the elements past rows*columns become 0 (out of place they keep the buffer's bytes)"""
    a = np.arange(1, length + 1, dtype=np.float32)
    rows = length // cols
    want = np.zeros(length, np.float32)
    for column in range(cols):
        for row in range(rows):
            want[cols * row + column] = a[rows * column + row]
    d = CudaVec.from_host(a)
    cuda.transpose_inplace(d, cols)
    assert np.array_equal(d.to_host(), want)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("rows,cols,batch", [(4, 4, 1001), (12, 12, 100), (16, 16, 257), (20, 20, 33), (32, 32, 65), (3, 5, 77),
                                             (64, 64, 9), (128, 128, 3), (100, 100, 5), (96, 32, 4)])
def test_transpose_inplace_batched_bit_exact(dt, rows, cols, batch):
    rng = np.random.default_rng(rows * cols + batch)
    e = rows * cols
    a = _rand(rng, batch * e, dt)
    want = np.concatenate([oracle.transpose(a[i * e:(i + 1) * e], cols) for i in range(batch)])
    d = CudaBatch.from_host(a, e)
    cuda.transpose_inplace_batched(d, e, cols)
    assert d.to_host().ravel().tobytes() == want.tobytes()
    h = a.copy()
    cuda.transpose_inplace_batched(h, e, cols)
    assert h.tobytes() == want.tobytes()


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10])
def test_transpose_benchmark_variants_bit_exact(variant):
    """every kernel behind the transpose_variant sweep knob (profiles/tune.py) gives the same bytes"""
    from slas_b200.backends import set_tunable
    try:
        set_tunable("transpose_variant", variant)
        for dt in (np.float32, np.float64):
            for n, batch in ((2, 999), (4, 1001), (8, 300), (16, 257), (32, 65), (48, 11), (64, 9), (128, 3)):
                e = n * n
                a = _rand(np.random.default_rng(n), batch * e, dt)
                want = np.concatenate([oracle.transpose(a[i * e:(i + 1) * e], n) for i in range(batch)])
                buf = CudaBatch(batch, e, dt)
                cuda.transpose_batched(CudaVec.from_host(a), buf, e, n)
                assert buf.to_host().ravel().tobytes() == want.tobytes(), (dt, n)
                d = CudaBatch.from_host(a, e)
                cuda.transpose_inplace_batched(d, e, n)
                assert d.to_host().ravel().tobytes() == want.tobytes(), (dt, n, "in place")
    finally:
        set_tunable("transpose_variant", 0)


@pytest.mark.parametrize("knob,size,dt,variants", [
    ("gemm4s_variant", 4, np.float32, (1, 3, 4, 5, 6, 7)), ("gemm16s_variant", 16, np.float32, (1, 2, 3)),
    ("gemm32s_variant", 32, np.float32, (1, 3)), ("gemm32s_variant", 64, np.float32, (1, 2)),
    ("gemm32d_variant", 32, np.float64, (1,))])
def test_small_gemm_benchmark_variants_keep_parity(knob, size, dt, variants):
    """the alternative tile configurations behind the sweep knobs meet the same bar as the defaults"""
    from slas_b200.backends import set_tunable
    batch = 257
    rng = np.random.default_rng(size)
    e = size * size
    A, B = _rand(rng, batch * e, dt), _rand(rng, batch * e, dt)
    tol = (1e-5 if dt == np.float32 else 1e-13) * np.sqrt(size)
    try:
        for v in variants:
            set_tunable(knob, v)
            for ta, tb in ((False, False), (True, False), (False, True), (True, True)):
                C = CudaBatch(batch, e, dt)
                cuda.matrix_mul_batched(CudaVec.from_host(A), CudaVec.from_host(B), C, size, size, size, ta, tb)
                got = C.to_host().reshape(batch, e)
                for i in (0, 1, batch // 2, batch - 1):
                    ref, mag = _gemm_truth(A[i * e:(i + 1) * e], B[i * e:(i + 1) * e], size, size, size, ta, tb)
                    assert np.max(np.abs(got[i] - ref) / np.maximum(mag, 1e-30)) <= tol, (v, ta, tb, i)
    finally:
        set_tunable(knob, 0)


@pytest.mark.parametrize("dt", [np.uint8, np.int16])
@pytest.mark.parametrize("rows,cols,batch", [(128, 128, 1), (128, 256, 3), (384, 128, 2), (64, 64, 5), (64, 128, 2), (64, 384, 2),
                                             (256, 128, 2), (512, 256, 3), (768, 128, 1), (1024, 2048, 1), (128, 192, 1), (192, 128, 3)])
def test_transpose_narrow_elements_tiled(dt, rows, cols, batch):
    """1- and 2-byte elements in whole 128-byte tiles: word loads, byte gathers, 256-bit sector stores"""
    e = rows * cols
    a = (np.arange(batch * e, dtype=np.int64) * 2654435761 >> 9).astype(dt)
    want = np.concatenate([oracle.transpose(a[i * e:(i + 1) * e], cols) for i in range(batch)])
    buf = CudaBatch(batch, e, dt)
    cuda.transpose_batched(CudaVec.from_host(a), buf, e, cols)
    assert buf.to_host().ravel().tobytes() == want.tobytes()
    # a base that is only element-aligned takes the element-granular kernels: same bytes
    big = CudaVec.from_host(np.concatenate([np.zeros(1, dt), a]))
    out = CudaVec((batch * e) + 1, dt).zero_()
    cuda.transpose_batched(big.view(1, batch * e), out.view(1, batch * e), e, cols)
    assert out.to_host()[1:].tobytes() == want.tobytes()


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.complex128, np.int16, np.uint8])
@pytest.mark.parametrize("rows,cols,batch", [(3, 3, 4096), (2, 3, 1000), (5, 5, 777), (3, 5, 4099), (9, 9, 513), (24, 8, 300), (100, 3, 260),
                                             (1, 7, 5000), (7, 1, 5000), (12, 12, 1024), (33, 31, 256), (100, 100, 260), (70, 50, 257),
                                             (32, 8, 300), (96, 96, 256), (16, 3, 1000), (64, 48, 256), (48, 5, 513)])
def test_transpose_many_small_objects_staged(dt, rows, cols, batch):
    """batches of small objects of any shape: bulk-TMA staged gather, out of place and in place, whole and ragged
    final chunks; bytes that do not make whole 16-byte words fall back to the per-warp kernel -- same result"""
    e = rows * cols
    a = (np.arange(batch * e, dtype=np.int64) * 2654435761 >> 11).astype(dt)
    want = np.concatenate([oracle.transpose(a[i * e:(i + 1) * e], cols) for i in range(batch)])
    buf = CudaBatch(batch + 1, e, dt)
    buf.copy_from_host(np.full((batch + 1) * e, 7, dt))
    cuda.transpose_batched(CudaVec.from_host(a), buf.view(0, batch * e), e, cols)
    got = buf.to_host().ravel()
    assert got[:batch * e].tobytes() == want.tobytes()
    assert np.all(got[batch * e:] == np.array(7, dt))  # nothing written past the last object
    d = CudaBatch.from_host(a, e)
    cuda.transpose_inplace_batched(d, e, cols)
    assert d.to_host().ravel().tobytes() == want.tobytes()


def test_transpose_ragged_len_leaves_tail_untouched():
    a = np.arange(11, dtype=np.float32)  # columns = 3 -> rows = 3, elements 9, 10 never written (rust.rs:168-175)
    buf = np.full(11, -7.0, np.float32)
    cuda.transpose(a, buf, 3)
    ref = np.full(11, -7.0, np.float32)
    for column in range(3):
        for row in range(3):
            ref[3 * row + column] = a[3 * column + row]
    assert np.array_equal(buf, ref)


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int16, np.uint8])
@pytest.mark.parametrize("rows,cols,batch", [(4, 4, 1000), (16, 16, 257), (32, 32, 65), (3, 5, 77), (8, 2, 1), (64, 64, 9)])
def test_transpose_batched_bit_exact(dt, rows, cols, batch):
    rng = np.random.default_rng(rows * cols + batch)
    a = _rand(rng, batch * rows * cols, dt, -100.0, 100.0)  # integer element types: distinct small values
    want = np.concatenate([oracle.transpose(a[i * rows * cols:(i + 1) * rows * cols], cols) for i in range(batch)])
    buf = CudaBatch(batch, rows * cols, dt)
    cuda.transpose_batched(CudaVec.from_host(a), buf, rows * cols, cols)
    assert buf.to_host().ravel().tobytes() == want.tobytes()
    h = np.zeros_like(a)
    cuda.transpose_batched(a, h, rows * cols, cols)
    assert h.tobytes() == want.tobytes()


# ---- matrix_mul ---------------------------------------------------------------------------------------
def _batched_gemm_truth(A, B, batch, m, n, k, ta, tb):
    """f64 products and sum|a||b| of EVERY object of a batch (the scale the rounding errors live on)"""
    Am = A.reshape(batch, k, m).transpose(0, 2, 1) if ta else A.reshape(batch, m, k)
    Bm = B.reshape(batch, n, k).transpose(0, 2, 1) if tb else B.reshape(batch, k, n)
    ref = np.matmul(Am.astype(np.float64), Bm.astype(np.float64))
    mag = np.matmul(np.abs(Am).astype(np.float64), np.abs(Bm).astype(np.float64))
    return ref.ravel(), mag.ravel()


# observed error / bar per kernel family, printed at the end of the module (`pytest -s` / -rP shows it)
MARGINS = {}


def _margin(family, err, bar):
    r = float(np.max(err / np.maximum(bar, 1e-300))) if np.size(err) else 0.0
    MARGINS[family] = max(MARGINS.get(family, 0.0), r)
    return r


def _gemm_truth(A, B, m, n, k, ta, tb):
    Am = A.reshape(k, m).T if ta else A.reshape(m, k)
    Bm = B.reshape(n, k).T if tb else B.reshape(k, n)
    ref = Am.astype(np.float64) @ Bm.astype(np.float64)
    mag = np.abs(Am).astype(np.float64) @ np.abs(Bm).astype(np.float64)
    return ref.ravel(), mag.ravel()


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("size", [4, 8, 16, 32, 64])
@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
def test_batched_small_gemm_vs_oracle(dt, size, ta, tb):
    m = n = k = size
    for batch in (1, 3, 1000 if size <= 16 else (300 if size == 32 else 67)):
        rng = np.random.default_rng(size * 7 + batch)
        A, B = _rand(rng, batch * m * k, dt), _rand(rng, batch * k * n, dt)
        want = oracle.gemm_batched(A, B, batch, m, n, k, ta, tb)
        dc = CudaBatch(batch, m * n, dt)
        cuda.matrix_mul_batched(CudaVec.from_host(A), CudaVec.from_host(B), dc, m, n, k, ta, tb)
        got = dc.to_host().ravel()
        tol = TOL[np.dtype(dt)] * np.sqrt(k)
        # the tight bar on EVERY element of EVERY object, against the f64 truth and against the oracle
        ref, mag = _batched_gemm_truth(A, B, batch, m, n, k, ta, tb)
        assert _margin(f"batched {size}^3 {np.dtype(dt).name} matrix_mul vs f64 truth", np.abs(got - ref), tol * mag) <= 1.0
        assert np.all(np.abs(got - want) <= 2 * tol * mag + 1e-300)
        h = np.zeros(batch * m * n, dt)
        cuda.matrix_mul_batched(A, B, h, m, n, k, ta, tb)  # host pointers: chunked H2D / kernel / D2H pipeline
        assert np.array_equal(h, got)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("m,n,k", [(2, 2, 3), (3, 3, 2), (8, 4, 16), (1, 5, 7), (16, 32, 8), (5, 1, 4), (12, 12, 12), (32, 8, 64)])
@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
def test_batched_rectangular_shapes_vs_oracle(dt, m, n, k, ta, tb):
    """Static shapes without a register-tiled specialisation (the reference's own 2x3 . 3x2 included) take the
    runtime-shape kernel behind the same bulk-TMA staging; tiny batches and odd totals take the plain kernel."""
    for batch in (1, 64, 1000, 1003):
        rng = np.random.default_rng(m * 100 + n * 10 + k + batch)
        A, B = _rand(rng, batch * m * k, dt), _rand(rng, batch * k * n, dt)
        want = oracle.gemm_batched(A, B, batch, m, n, k, ta, tb)
        dc = CudaBatch(batch, m * n, dt)
        cuda.matrix_mul_batched(CudaVec.from_host(A), CudaVec.from_host(B), dc, m, n, k, ta, tb)
        assert np.max(np.abs(dc.to_host().ravel() - want)) <= 2 * TOL[np.dtype(dt)] * np.sqrt(k) * k


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
def test_batched_tiny_shapes_thread_per_matrix(dt, ta, tb):
    """every (m, n, k) in 1..4 (2x3 . 3x2 of the reference's tests, 3x3, ...): the thread-per-matrix family, with
    whole and ragged final chunks; batches whose byte size is not a whole number of 16-byte words fall back"""
    from slas_b200.backends import set_tunable
    rng = np.random.default_rng(int(ta) * 2 + int(tb))
    for m in range(1, 5):
        for n in range(1, 5):
            for k in range(1, 5):
                for batch in (1024, 4100, 4098):
                    A, B = _rand(rng, batch * m * k, dt), _rand(rng, batch * k * n, dt)
                    want = oracle.gemm_batched(A, B, batch, m, n, k, ta, tb)
                    dc = CudaBatch(batch + 1, m * n, dt)
                    _lib_fill = np.full((batch + 1) * m * n, -3.0, dt)
                    dc.copy_from_host(_lib_fill)
                    cuda.matrix_mul_batched(CudaVec.from_host(A), CudaVec.from_host(B), dc.view(0, batch * m * n), m, n, k, ta, tb)
                    got = dc.to_host().ravel()
                    assert np.max(np.abs(got[:batch * m * n] - want)) <= 2 * TOL[np.dtype(dt)] * np.sqrt(k) * k, (m, n, k, batch)
                    assert np.all(got[batch * m * n:] == -3.0), (m, n, k, batch)  # nothing written past the last matrix
    # the same answers as the runtime-shape kernel it replaces (ascending-k FMA chain in both): bit-identical
    m, n, k, batch = 3, 3, 3, 4096
    A, B = _rand(rng, batch * m * k, dt), _rand(rng, batch * k * n, dt)
    c0, c1 = CudaBatch(batch, m * n, dt), CudaBatch(batch, m * n, dt)
    cuda.matrix_mul_batched(CudaVec.from_host(A), CudaVec.from_host(B), c0, m, n, k, ta, tb)
    set_tunable("gemm_tiny_variant", 1)
    try:
        cuda.matrix_mul_batched(CudaVec.from_host(A), CudaVec.from_host(B), c1, m, n, k, ta, tb)
    finally:
        set_tunable("gemm_tiny_variant", 0)
    assert np.array_equal(c0.to_host(), c1.to_host())


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("m,n,k", [(8, 4, 16), (16, 32, 8), (12, 12, 12), (32, 8, 64), (6, 6, 6), (24, 8, 12), (5, 8, 4), (8, 8, 16),
                                   (48, 16, 4), (5, 5, 5), (3, 7, 5), (7, 7, 7), (1, 6, 9), (9, 9, 9), (10, 10, 10), (5, 9, 13), (33, 3, 7),
                                   (15, 15, 15), (20, 20, 20), (24, 24, 24), (48, 48, 48), (40, 40, 40), (100, 100, 100), (70, 33, 50),
                                   (28, 28, 28), (36, 36, 36), (44, 20, 12)])
def test_batched_shapes_specialised_at_run_time(dt, m, n, k):
    """shapes without a built-in instantiation: the gemm_small kernel compiled for the shape with NVRTC on first use
    (gemm_jit.cu).  Without b_trans every C element is the same ascending-k FMA chain as in the runtime-shape kernel it
    replaces => identical bits (with b_trans the k-chunks are visited in a rotated order: same bar, different rounding);
    shapes outside the tiling rules (f32 6x6x6, 5x5x5: K is not a whole 16-byte vector) get the thread-per-matrix kernel
    specialised instead if a product fits one thread's registers, and scalar register tiles otherwise (9x9x9, 5x9x13)."""
    from slas_b200.backends import set_tunable
    rng = np.random.default_rng(m * 100 + n * 10 + k)
    for batch in ((64, 1000, 1003) if m * n * k <= 4096 else (64, 131)):  # medium sizes: also the one-layer-per-object kernel
        A, B = _rand(rng, batch * m * k, dt), _rand(rng, batch * k * n, dt)
        dA, dB = CudaVec.from_host(A), CudaVec.from_host(B)
        for ta in (False, True):
            for tb in (False, True):
                want = oracle.gemm_batched(A, B, batch, m, n, k, ta, tb)
                c0, c1 = CudaBatch(batch + 1, m * n, dt), CudaBatch(batch, m * n, dt)
                c0.copy_from_host(np.full((batch + 1) * m * n, -5.0, dt))
                cuda.matrix_mul_batched(dA, dB, c0.view(0, batch * m * n), m, n, k, ta, tb)
                got = c0.to_host().ravel()
                assert np.max(np.abs(got[:batch * m * n] - want)) <= 2 * TOL[np.dtype(dt)] * np.sqrt(k) * k, (batch, ta, tb)
                assert np.all(got[batch * m * n:] == -5.0)
                set_tunable("gemm_jit", 1)  # the runtime-shape kernel
                try:
                    cuda.matrix_mul_batched(dA, dB, c1, m, n, k, ta, tb)
                finally:
                    set_tunable("gemm_jit", 0)
                if not tb or n % (16 // np.dtype(dt).itemsize) or k % (16 // np.dtype(dt).itemsize):
                    assert np.array_equal(c1.to_host().ravel(), got[:batch * m * n]), (batch, ta, tb)
                else:
                    assert np.max(np.abs(c1.to_host().ravel() - want)) <= 2 * TOL[np.dtype(dt)] * np.sqrt(k) * k


def test_batched_view_api_and_integer_exactness():
    # small-integer entries: every product and sum is exact in f32, so == numpy exactly
    rng = np.random.default_rng(3)
    batch = 513
    for size in (4, 16, 32):
        A = rng.integers(-4, 5, batch * size * size).astype(np.float32)
        B = rng.integers(-4, 5, batch * size * size).astype(np.float32)
        a3, b3 = A.reshape(batch, size, size), B.reshape(batch, size, size)
        va = T.BatchedMatrix(CudaVec.from_host(A), size, size, cuda)
        vb = T.BatchedMatrix(CudaVec.from_host(B), size, size, cuda)
        assert np.array_equal(va.matrix_mul(vb).to_host().reshape(batch, size, size), a3 @ b3)
        assert np.array_equal(va.transpose().matrix_mul(vb).to_host().reshape(batch, size, size), a3.transpose(0, 2, 1) @ b3)
        assert np.array_equal(va.matrix_mul(vb.transpose()).to_host().reshape(batch, size, size), a3 @ b3.transpose(0, 2, 1))
        assert np.array_equal(va.transpose().matrix_mul(vb.transpose()).to_host().reshape(batch, size, size),
                              a3.transpose(0, 2, 1) @ b3.transpose(0, 2, 1))
        out = CudaBatch(batch, size * size, np.float32)
        va.materialize_transpose(out)
        assert np.array_equal(out.to_host().reshape(batch, size, size), a3.transpose(0, 2, 1))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (2, 2, 3), (5, 7, 3), (1, 9, 2), (17, 33, 65), (64, 64, 64), (100, 130, 70), (128, 128, 8), (257, 129, 31)])
@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
def test_single_gemm_any_shape(dt, m, n, k, ta, tb):
    rng = np.random.default_rng(m * n + k)
    pad = 3  # leading dimensions wider than the matrix: C keeps its bytes between rows (beta = 0 covers m x n only)
    lda, ldb, ldc = (m if ta else k) + pad, (k if tb else n) + pad, n + pad
    a_rows, b_rows = (k if ta else m), (n if tb else k)
    A, B = _rand(rng, a_rows * lda, dt), _rand(rng, b_rows * ldb, dt)
    C0 = np.full(m * ldc, 9.0, dt)
    want = oracle.gemm(A, B, m, n, k, lda, ldb, ldc, ta, tb, f64acc=(dt == np.float32))
    for dev in (False, True):
        c = CudaVec.from_host(C0) if dev else C0.copy()
        cuda.matrix_mul(CudaVec.from_host(A) if dev else A, CudaVec.from_host(B) if dev else B, c, m, n, k, lda, ldb, ldc, ta, tb)
        got = _host(c).reshape(m, ldc)
        assert np.all(got[:, n:] == 9.0)
        Av = A.reshape(a_rows, lda)[:, : (m if ta else k)]
        Bv = B.reshape(b_rows, ldb)[:, : (k if tb else n)]
        ref, mag = _gemm_truth(np.ascontiguousarray(Av).ravel(), np.ascontiguousarray(Bv).ravel(), m, n, k, ta, tb)
        assert _margin(f"single matrix_mul {np.dtype(dt).name} (any shape) vs f64 truth", np.abs(got[:, :n].ravel() - ref),
                       TOL[np.dtype(dt)] * np.sqrt(k) * mag + 1e-300) <= 1.0
        assert np.max(np.abs(got[:, :n] - want.reshape(m, ldc)[:, :n])) <= 2 * TOL[np.dtype(dt)] * np.sqrt(k) * k


@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
def test_large_dgemm(ta, tb):
    """f64 products from 2^30 multiply-adds up materialise a lazily transposed A first (the DFMA kernel's fast layout)"""
    rng = np.random.default_rng(31)
    m, n, k = 1024, 1088, 1024
    A, B = _rand(rng, m * k, np.float64), _rand(rng, k * n, np.float64)
    c = CudaVec(m * n, np.float64)
    cuda.matrix_mul(CudaVec.from_host(A), CudaVec.from_host(B), c, m, n, k, m if ta else k, k if tb else n, n, ta, tb)
    ref, mag = _gemm_truth(A, B, m, n, k, ta, tb)  # numpy f64 (pairwise / blocked sums): well inside the bar itself
    assert np.all(np.abs(c.to_host() - ref) <= 1e-13 * np.sqrt(k) * mag)


@pytest.mark.parametrize("size", [512, 1024])   # 1024^3 and up: operands are first put into the k-major / n-major form
@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
def test_large_gemm_ffma_path(size, ta, tb):
    rng = np.random.default_rng(12 + size)
    m, n, k = size, size + 64, size
    A, B = _rand(rng, m * k, np.float32), _rand(rng, k * n, np.float32)
    c = CudaVec(m * n, np.float32)
    from slas_b200.backends import set_sgemm_path
    set_sgemm_path("ffma")
    try:
        cuda.matrix_mul(CudaVec.from_host(A), CudaVec.from_host(B), c, m, n, k, m if ta else k, k if tb else n, n, ta, tb)
    finally:
        set_sgemm_path("auto")
    ref, mag = _gemm_truth(A, B, m, n, k, ta, tb)
    assert np.all(np.abs(c.to_host() - ref) <= 1e-5 * np.sqrt(k) * mag)


@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
def test_sgemm_3xtf32_strided_unaligned_operands(ta, tb):
    """leading dimensions wider than the matrices, not multiples of 4, and a C base that is not 16-byte aligned:
    the split pass and the epilogue fall back to scalar accesses; untouched C padding keeps its bytes"""
    from slas_b200.backends import set_sgemm_path
    m, n, k = 200, 300, 150
    lda, ldb, ldc = (m if ta else k) + 3, (k if tb else n) + 1, n + 5
    rng = np.random.default_rng(77)
    A = _rand(rng, (k if ta else m) * lda, np.float32)
    B = _rand(rng, (n if tb else k) * ldb, np.float32)
    cfull = CudaVec.from_host(np.full(1 + m * ldc, -7.0, np.float32))
    set_sgemm_path("tcgen05")
    try:
        cuda.matrix_mul(CudaVec.from_host(A), CudaVec.from_host(B), cfull.view(1, m * ldc), m, n, k, lda, ldb, ldc, ta, tb)
    finally:
        set_sgemm_path("auto")
    got = cfull.to_host()[1:].reshape(m, ldc)
    Am = A.reshape(k, lda)[:, :m].T if ta else A.reshape(m, lda)[:, :k]
    Bm = B.reshape(n, ldb)[:, :k].T if tb else B.reshape(k, ldb)[:, :n]
    ref = Am.astype(np.float64) @ Bm.astype(np.float64)
    mag = np.abs(Am).astype(np.float64) @ np.abs(Bm).astype(np.float64)
    assert np.all(np.abs(got[:, :n] - ref) <= 1e-5 * np.sqrt(k) * mag)
    assert np.all(got[:, n:] == -7.0) and cfull.to_host()[0] == -7.0


def test_sgemm_3xtf32_infinities_and_nans_follow_ieee():
    """an infinite operand must give the signed infinity of the fp32 chain, not the NaN a naive hi/lo split produces
    (0 * Inf in the cross terms); NaN and Inf * 0 stay NaN"""
    from slas_b200.backends import set_sgemm_path
    m, n, k = 128, 256, 64
    rng = np.random.default_rng(3)
    A = (rng.random(m * k) + 0.5).astype(np.float32)   # all positive
    B = (rng.random(k * n) + 0.5).astype(np.float32)
    A2, B2 = A.reshape(m, k).copy(), B.reshape(k, n).copy()
    A2[3, 5] = np.inf          # row 3 of C -> +inf
    A2[7, 1] = -np.inf         # row 7 -> -inf
    B2[9, 11] = np.inf         # column 11 -> +inf (rows 3 and 7 too: inf + inf, -inf + inf = nan)
    A2[20, 2] = np.nan         # row 20 -> nan
    A2[30, :] = 1.0            # exactly representable in tf32 (lo == 0) against the infinite B entry
    B2[0, 40] = 0.0
    A2[50, 0] = np.inf         # inf * 0 in (50, 40) -> nan
    set_sgemm_path("tcgen05")
    try:
        c = CudaVec(m * n, np.float32)
        cuda.matrix_mul(CudaVec.from_host(A2.ravel()), CudaVec.from_host(B2.ravel()), c, m, n, k, k, n, n, False, False)
    finally:
        set_sgemm_path("auto")
    got = c.to_host().reshape(m, n)
    with np.errstate(all="ignore"):
        want = A2.astype(np.float64) @ B2.astype(np.float64)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(np.isposinf(got), np.isposinf(want)) and np.array_equal(np.isneginf(got), np.isneginf(want))
    fin = np.isfinite(want)
    assert np.all(np.abs(got[fin] - want[fin]) <= 1e-5 * np.sqrt(k) * np.abs(want[fin]))


def test_large_sgemm_automatic_path_choice():
    """default ('auto'): tensor cores from 2^27 multiply-adds up when the shape fits the tiles, FFMA otherwise"""
    from slas_b200.backends import set_sgemm_path
    rng = np.random.default_rng(5)
    # 520 x 260 would be padded to 640 x 512 (2.4x the work): FFMA; 256^3 is below the 2^27 threshold
    for (m, n, k), launches in (((512, 512, 512), 3), ((500, 520, 530), 3), ((520, 260, 1024), 1), ((256, 256, 256), 1)):
        A, B = _rand(rng, m * k, np.float32), _rand(rng, k * n, np.float32)
        c = CudaVec(m * n, np.float32)
        n0 = launch_count()
        cuda.matrix_mul(CudaVec.from_host(A), CudaVec.from_host(B), c, m, n, k, k, n, n, False, False)
        assert launch_count() - n0 == launches, (m, n, k)
        ref, mag = _gemm_truth(A, B, m, n, k, False, False)
        assert np.all(np.abs(c.to_host() - ref) <= 1e-5 * np.sqrt(k) * mag)
    set_sgemm_path("ffma")
    try:
        A, B = _rand(rng, 512 * 512, np.float32), _rand(rng, 512 * 512, np.float32)
        c = CudaVec(512 * 512, np.float32)
        n0 = launch_count()
        cuda.matrix_mul(CudaVec.from_host(A), CudaVec.from_host(B), c, 512, 512, 512, 512, 512, 512, False, False)
        assert launch_count() - n0 == 1
    finally:
        set_sgemm_path("auto")


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("rows,cols", [(1, 1), (3, 2), (2, 3), (16, 16), (33, 100), (257, 64), (4, 4), (32, 32), (64, 16),
                                       (2, 128), (128, 8), (1, 4), (256, 64), (64, 128), (300, 1024), (1000, 516), (2048, 2048),
                                       (5, 24), (7, 40), (3, 256)])  # lane-per-row kernel: any row count, rows of whole 32-byte chunks
@pytest.mark.parametrize("ta", [False, True])
def test_gemv_vs_oracle(dt, rows, cols, ta):
    rng = np.random.default_rng(rows + cols)
    A, x = _rand(rng, rows * cols, dt), _rand(rng, rows if ta else cols, dt)
    want = oracle.gemv(A, x, cols, rows, cols, ta)  # trait: m = width, n = height
    tol = 2 * TOL[np.dtype(dt)] * np.sqrt(len(x)) * len(x)
    for dev in (False, True):
        y = CudaVec(want.size, dt) if dev else np.zeros(want.size, dt)
        cuda.matrix_vector_mul(CudaVec.from_host(A) if dev else A, CudaVec.from_host(x) if dev else x, y, cols, rows, cols, ta)
        assert np.max(np.abs(_host(y) - want)) <= tol
    batch = 37
    Ab, xb = _rand(rng, batch * rows * cols, dt), _rand(rng, batch * len(x), dt)
    yb = CudaBatch(batch, want.size, dt)
    cuda.matrix_vector_mul_batched(CudaVec.from_host(Ab), CudaVec.from_host(xb), yb, cols, rows, ta)
    got = yb.to_host()
    for i in range(batch):
        w = oracle.gemv(Ab[i * rows * cols:(i + 1) * rows * cols], xb[i * len(x):(i + 1) * len(x)], cols, rows, cols, ta)
        assert np.max(np.abs(got[i] - w)) <= tol


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("rows,cols", [(3, 3), (4, 3), (3, 4), (6, 6), (5, 5), (9, 9), (12, 12), (7, 3), (1, 5), (5, 1), (20, 6)])
@pytest.mark.parametrize("ta", [False, True])
def test_gemv_batched_small_odd_shapes(dt, rows, cols, ta):
    """3x3 . 3-vector and friends, large batches: routed through the shape-specialised batched GEMM kernels (B with one
    column); whole and ragged chunks, nothing written past the last result"""
    rng = np.random.default_rng(rows * 31 + cols)
    for batch in (64, 1000, 4099):
        xl, yl = (rows, cols) if ta else (cols, rows)
        A, x = _rand(rng, batch * rows * cols, dt), _rand(rng, batch * xl, dt)
        y = CudaBatch(batch + 1, yl, dt)
        y.copy_from_host(np.full((batch + 1) * yl, -9.0, dt))
        cuda.matrix_vector_mul_batched(CudaVec.from_host(A), CudaVec.from_host(x), y.view(0, batch * yl), cols, rows, ta)
        got = y.to_host().ravel()
        A3 = A.reshape(batch, rows, cols).astype(np.float64)
        ref = np.einsum("bji,bj->bi" if ta else "bij,bj->bi", A3, x.reshape(batch, xl).astype(np.float64)).ravel()
        mag = np.einsum("bji,bj->bi" if ta else "bij,bj->bi", np.abs(A3), np.abs(x.reshape(batch, xl)).astype(np.float64)).ravel()
        assert np.all(np.abs(got[:batch * yl] - ref) <= TOL[np.dtype(dt)] * np.sqrt(xl) * mag + 1e-300), (batch,)
        assert np.all(got[batch * yl:] == -9.0)


# ---- batched dot / norm / normalize --------------------------------------------------------------------
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("length,batch", [(1, 5), (3, 100), (16, 1000), (100, 333), (750, 64), (4096, 9), (20000, 5),
                                          (4, 3001), (8, 777), (24, 1025), (32, 33), (40, 999), (64, 257), (72, 100),  # a lane per vector (<= 256 bytes)
                                          (3, 5000), (5, 1000), (7, 4099), (9, 257), (30, 1025), (100, 300), (127, 256), (1, 4096),
                                          (6, 2000), (3, 4097), (50, 700), (127, 300), (150, 513), (201, 257),
                                          (1001, 37), (333, 100), (5001, 9), (8195, 5), (65, 300), (251, 64),  # odd lengths: per-vector peeling
                                          # ^ not whole 16-byte words: bulk-TMA staged, a thread per vector
                                          # warp / CTA per vector with ragged piece counts, more vectors than groups in
                                          # the grid, and the lengths either side of every kernel switch
                                          (1028, 77), (2052, 19), (4092, 3000), (8192, 11), (8196, 7), (12004, 1500),
                                          (32768, 5), (32772, 3), (65536, 3),
                                          # 8 KB .. 512 KB: a thread-block cluster per vector (1 / 2 / 4 / 8 CTAs, 4 / 8 / 16
                                          # vectors per thread), more vectors than resident clusters, either side of its range
                                          (2052, 700), (2560, 2000), (6000, 900), (49152, 300), (98304, 40), (131072, 3), (131076, 2)])
def test_batched_reductions(dt, length, batch):
    rng = np.random.default_rng(length + batch)
    a, b = _rand(rng, batch * length, dt), _rand(rng, batch * length, dt)
    a2, b2 = a.reshape(batch, length).astype(np.float64), b.reshape(batch, length).astype(np.float64)
    rel = TOL[np.dtype(dt)] * np.sqrt(length)
    out = CudaVec(batch, dt)
    cuda.dot_batched(CudaVec.from_host(a), CudaVec.from_host(b), out, length)
    assert _margin(f"batched dot {np.dtype(dt).name} vs f64 truth", np.abs(out.to_host() - np.sum(a2 * b2, axis=1)),
                   rel * np.sum(np.abs(a2 * b2), axis=1) + 1e-300) <= 1.0
    h = np.zeros(batch, dt)
    cuda.dot_batched(a, b, h, length)
    assert np.array_equal(h, out.to_host())
    cuda.norm_batched(CudaVec.from_host(a), out, length)
    norms = out.to_host()
    assert np.all(np.abs(norms - np.sqrt(np.sum(a2 * a2, axis=1))) <= rel * np.sqrt(np.sum(a2 * a2, axis=1)))
    d = CudaBatch.from_host(a, length)
    cuda.normalize_batched(d, length)
    assert np.array_equal(d.to_host(), a.reshape(batch, length) / norms[:, None])  # true division, bit-exact


# ---- the tensor-core path for one large f32 product: 3xTF32 on tcgen05 / TMEM -----------------------------
@pytest.mark.parametrize("m,n,k", [(128, 256, 32), (128, 256, 64), (256, 512, 96), (384, 256, 1024), (1024, 1024, 512),
                                   (100, 70, 50), (129, 257, 33), (300, 500, 700), (1000, 1000, 1000), (257, 64, 8)])
@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
def test_sgemm_3xtf32_tcgen05(m, n, k, ta, tb):
    from slas_b200.backends import set_sgemm_path
    rng = np.random.default_rng(m + n + k)
    A, B = _rand(rng, m * k, np.float32), _rand(rng, k * n, np.float32)
    c = CudaVec(m * n, np.float32).zero_()
    set_sgemm_path("tcgen05")
    try:
        n0 = launch_count()
        cuda.matrix_mul(CudaVec.from_host(A), CudaVec.from_host(B), c, m, n, k, m if ta else k, k if tb else n, n, ta, tb)
        assert launch_count() - n0 == 3, "expected split(A), split(B) and the tcgen05 kernel"
    finally:
        set_sgemm_path("auto")
    ref, mag = _gemm_truth(A, B, m, n, k, ta, tb)
    err = np.abs(c.to_host() - ref)
    assert _margin("sgemm 3xTF32 (tcgen05) vs f64 truth", err, 1e-5 * np.sqrt(k) * mag) <= 1.0, f"max err/mag {np.max(err / mag):.3e}"
    # 3xTF32 is an fp32-class result: far tighter than plain TF32 (2^-11 per operand)
    assert np.max(err / mag) < 2e-6
