"""Pins the CPU oracle (oracle/) against every golden vector the reference's own tests
hold for the hot path, and cross-checks the C restatement against
the `Blas` stand-in (identical cblas calls on scipy's OpenBLAS) and numpy."""
import numpy as np
import pytest

from oracle import blas_standin as blas
from oracle import oracle

DT = {"f32": np.float32, "f64": np.float64, "c64": np.complex64}


def _cplx(pairs, dt):
    return np.array([complex(r, i) for r, i in pairs], dtype=dt)


def test_dot_golden(golden):
    for g in golden["dot"]:
        dt = DT[g["dtype"]]
        a, b = np.array(g["a"], dtype=dt), np.array(g["b"], dtype=dt)
        results = []
        if g["backend"] in ("Rust", "default"):
            # default dispatch: LEN < 750 -> Rust (src/backends.rs:273-279)
            for lanes in (oracle.default_lanes(dt), oracle.default_lanes(dt, avx=True)):
                results.append(oracle.dot(a, b, lanes=lanes))
        if g["backend"] in ("Blas",):
            results.append(blas.dot(a, b))
        for r in results:
            if "abs_tol" in g:
                assert abs(float(r) - g["expect"]) < g["abs_tol"], g["cite"]
            else:
                assert float(r) == g["expect"], g["cite"]


def test_dotu_golden(golden):
    for g in golden["dotu"]:
        a, b = _cplx(g["a"], np.complex64), _cplx(g["b"], np.complex64)
        exp = complex(*g["expect"])
        assert complex(oracle.dotu(a, b)) == exp, g["cite"]
        assert complex(blas.dot(a, b)) == exp, g["cite"]


def test_complex_norm_golden(golden):
    for g in golden["norm"]:
        a = _cplx(g["a"], np.complex64)
        exp = np.float32(g["expect"])
        assert oracle.norm(a) == exp, g["cite"]       # Rust backend
        assert blas.norm(a) == exp, g["cite"]         # Blas backend


def test_ewise_golden(golden):
    ops = {"add": np.add, "sub": np.subtract, "mul": np.multiply, "div": np.divide}
    for g in golden["ewise"]:
        dt = DT[g["dtype"]]
        a, b = np.array(g["a"], dtype=dt), np.array(g["b"], dtype=dt)
        got = oracle.ewise(g["op"], a, b)
        assert got.dtype == dt
        assert np.array_equal(got, ops[g["op"]](a, b)), g["cite"]


def test_transpose_golden(golden):
    for g in golden["transpose"]:
        for dt in (np.float32, np.float64, np.int16, np.uint8):
            a = np.array(g["a"], dtype=dt)
            assert np.array_equal(oracle.transpose(a, g["columns"]), np.array(g["expect"], dtype=dt)), g["cite"]


def _matmul_case(g, dt, gemm):
    a, b = g["a"], g["b"]
    m, n, k, lda, ldb, ldc = oracle.matmul_args(a["shape"], a["trans"], b["shape"], b["trans"])
    c = gemm(np.array(a["data"], dtype=dt), np.array(b["data"], dtype=dt), m, n, k, lda, ldb, ldc,
             a["trans"], b["trans"])
    return c[: m * n], (m, n, k)


def test_matmul_golden(golden):
    for g in golden["matmul"]:
        for dt in (np.float32, np.float64):
            exp = np.array(g["expect"], dtype=dt)
            got, (m, n, _) = _matmul_case(g, dt, oracle.gemm)
            assert m * n == exp.size, g["cite"]
            assert np.array_equal(got, exp), g["cite"]
            got_b, _ = _matmul_case(g, dt, blas.gemm)
            assert np.array_equal(got_b, exp), g["cite"]
        got64, _ = _matmul_case(g, np.float32, lambda *a: oracle.gemm(*a, f64acc=True))
        assert np.array_equal(got64, np.array(g["expect"], dtype=np.float32)), g["cite"]


def test_gemv_golden(golden):
    for g in golden["gemv"]:
        a = g["a"]
        for dt in (np.float32, np.float64):
            A = np.array(a["data"], dtype=dt)
            x = np.array(g["x"], dtype=dt)
            # Matrix::vector_mul_buffer passes m = axis_len(0), n = axis_len(1), lda = axis_len(0)
            m, n, lda = a["shape"][0], a["shape"][1], a["shape"][0]
            exp = np.array(g["expect"], dtype=dt)
            assert np.array_equal(oracle.gemv(A, x, m, n, lda, a["trans"]), exp), g["cite"]
            assert np.array_equal(blas.gemv(A, x, m, n, lda, a["trans"]), exp), g["cite"]
            # numpy cross-check of the transcribed expectation
            M = A.reshape(n, m)
            ref = (M.T if a["trans"] else M) @ x
            assert np.array_equal(ref, exp), g["cite"]
            # gemv == gemm with n = 1 (tests/src/main.rs:355-394)
            mm, nn, kk, la, lb, lc = oracle.matmul_args(a["shape"], a["trans"], [1, x.size], False)
            assert np.array_equal(oracle.gemm(A, x, mm, nn, kk, la, lb, lc, a["trans"], False)[:mm], exp)


def test_index_golden(golden):
    for g in golden["index"]:
        kind = g["kind"]
        if kind == "tensor":
            assert oracle.tensor_index(g["shape"], g["index"]) == g["expect"]
        elif kind == "matrix_vs_tensor":
            r, c = g["matrix_rc"]
            assert oracle.tensor_index(g["shape"], g["tensor_index"]) == oracle.matrix_index(g["shape"], False, r, c)
        elif kind == "index_slice":
            vol = int(np.prod(g["shape"]))
            assert g["slice"] * (vol // g["shape"][-1]) == g["offset"]
            data = np.arange(vol, dtype=np.float32)[g["offset"]:]
            r, c = g["matrix_rc"]
            assert data[oracle.matrix_index(g["sub_shape"], False, r, c)] == g["expect_value"]
            r, c = g["transposed_rc"]
            assert data[oracle.matrix_index(g["sub_shape"], True, r, c)] == g["expect_transposed_value"]
        elif kind == "lazy_transpose":
            for (r0, c0), (r1, c1) in g["pairs"]:
                assert oracle.matrix_index(g["shape"], False, r0, c0) == oracle.matrix_index(g["shape"], True, r1, c1)
        elif kind == "out_of_bounds":
            with pytest.raises(IndexError):
                oracle.tensor_index(g["shape"], g["index"])


def test_matmul_args_rows_columns():
    # src/tensor.rs:376-382 + 495-510 on the six golden layouts
    assert oracle.matmul_args([3, 2], False, [2, 3], False) == (2, 2, 3, 3, 2, 2)
    assert oracle.matmul_args([3, 2], True, [2, 3], True) == (3, 3, 2, 3, 2, 3)
    assert oracle.matmul_args([2, 3], False, [2, 3], True) == (3, 3, 2, 2, 2, 3)


# ---- randomized cross-checks: C restatement vs Blas stand-in vs numpy (float64 truth) ----

@pytest.mark.parametrize("dt,tol", [(np.float32, 1e-5), (np.float64, 1e-13)])
@pytest.mark.parametrize("n", [1, 3, 7, 8, 100, 750, 4097])
def test_dot_oracle_vs_blas_and_truth(dt, tol, n):
    rng = np.random.default_rng(n)
    a, b = rng.random(n).astype(dt), (rng.random(n) * 2 - 1).astype(dt)
    truth = float(np.dot(a.astype(np.longdouble), b.astype(np.longdouble)))
    scale = float(np.sum(np.abs(a.astype(np.float64) * b.astype(np.float64)))) + 1e-300
    for lanes in (1, 2, 4, 8):
        assert abs(float(oracle.dot(a, b, lanes=lanes)) - truth) <= tol * np.sqrt(n) * scale
    assert abs(float(blas.dot(a, b)) - truth) <= tol * np.sqrt(n) * scale


@pytest.mark.parametrize("dt,tol", [(np.float32, 1e-5), (np.float64, 1e-13)])
def test_gemm_oracle_vs_blas(dt, tol):
    rng = np.random.default_rng(5)
    for (m, n, k) in [(4, 4, 4), (16, 16, 16), (5, 7, 3), (32, 32, 32), (1, 9, 2)]:
        for ta in (False, True):
            for tb in (False, True):
                A = (rng.random(m * k) * 2 - 1).astype(dt)
                B = (rng.random(k * n) * 2 - 1).astype(dt)
                lda, ldb = (m if ta else k), (k if tb else n)
                c0 = oracle.gemm(A, B, m, n, k, lda, ldb, n, ta, tb)
                c1 = blas.gemm(A, B, m, n, k, lda, ldb, n, ta, tb)
                Am = A.reshape(k, m).T if ta else A.reshape(m, k)
                Bm = B.reshape(n, k).T if tb else B.reshape(k, n)
                ref = (Am.astype(np.float64) @ Bm.astype(np.float64)).ravel()
                bound = tol * np.sqrt(k) * (np.abs(Am).astype(np.float64) @ np.abs(Bm).astype(np.float64)).ravel()
                assert np.all(np.abs(c0 - ref) <= bound + 1e-300)
                assert np.all(np.abs(c1 - ref) <= bound + 1e-300)


def test_norm_normalize_oracle():
    rng = np.random.default_rng(11)
    for dt in (np.float32, np.float64):
        a = (rng.random(1000) * 2 - 1).astype(dt)
        nrm = oracle.norm(a)
        assert abs(float(nrm) - float(np.linalg.norm(a.astype(np.float64)))) < 1e-4 * float(nrm)
        out = oracle.normalize(a)
        assert np.array_equal(out, a / nrm)          # true division by the oracle's own norm
        assert abs(float(blas.norm(a)) - float(nrm)) < 1e-5 * float(nrm)


def test_batched_loops_agree():
    rng = np.random.default_rng(3)
    batch, m = 64, 4
    A = (rng.random(batch * m * m) * 2 - 1).astype(np.float32)
    B = (rng.random(batch * m * m) * 2 - 1).astype(np.float32)
    c0 = oracle.gemm_batched(A, B, batch, m, m, m)
    c1 = blas.gemm_batched(A, B, np.zeros_like(c0), batch, m, m, m, threads=2)
    ref = np.einsum("bij,bjk->bik", A.reshape(batch, m, m).astype(np.float64), B.reshape(batch, m, m).astype(np.float64))
    assert np.allclose(c0, ref.ravel(), rtol=0, atol=1e-5)
    assert np.allclose(c1, ref.ravel(), rtol=0, atol=1e-5)
