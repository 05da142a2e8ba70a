"""Writes tests/golden/reference_vectors.json.

The reference (unic0rn9k/slas) is nightly-only Rust and this image has no
rustc/cargo, so the reference cannot be run to generate fixtures.  Every vector
below is TRANSCRIBED from a literal in the reference's own tests / doc-tests; the
`cite` field is the file:line (relative to the reference root) it comes from.
They pin conventions (layout, lazy-transpose flags, lda/ldb, m/n meaning of gemv,
unconjugated complex dot), not rounding: all expected values are exact in f32.

Run:  python tests/golden/make_golden.py
"""
import json
import os

r16 = [float(v) for v in range(1, 7)]          # moo![f32: 1..=6]


def mat(vals, m, k, trans=False):
    # .matrix::<B, M, K>() -> MatrixShape<M,K>, slice() == [K, M] (src/tensor.rs:112-114);
    # .transpose() only flips IS_TRANS (src/tensor.rs:512-514).
    return {"data": vals, "shape": [k, m], "trans": trans}


G = {
    "_about": "golden vectors transcribed from unic0rn9k/slas tests; see make_golden.py",
    "dot": [
        {"cite": "tests/src/main.rs:41-46", "backend": "Blas", "dtype": "f32",
         "a": [0, 1, 2, 3], "b": [0, -1, -2, 3], "expect": 4.0},
        {"cite": "tests/src/main.rs:49-56", "backend": "default", "dtype": "f32",
         "a": [0, 1, 2, 3], "b": [0, -1, -2, 3], "expect": 4.0},
        {"cite": "tests/src/main.rs:70-78", "backend": "Blas", "dtype": "f32",
         "a": [0, 1, 2, 3], "b": [1, 2, 3, 4], "expect": 20.0},
        {"cite": "tests/src/main.rs:81-95", "backend": "Blas", "dtype": "f32",
         "a": [1, 2, 3, 4], "b": [0, 1, 2, 3], "expect": 20.0},
        {"cite": "src/backends.rs:58-66", "backend": "Rust", "dtype": "f32",
         "a": [0, 1, 2, 3], "b": [1, 2, 3, 4], "expect": 20.0},
        {"cite": "tests/src/main.rs:146-148", "backend": "default", "dtype": "f32",
         "a": [1, 2, 3], "b": [1, 2, 3], "expect": 14.0},
        {"cite": "tests/src/main.rs:151-155", "backend": "Rust", "dtype": "f64",
         "a": [1, 2, 3, 4], "b": [1, 2, 3, 4], "expect": 30.0},
        {"cite": "tests/src/main.rs:156-159", "backend": "Rust", "dtype": "f64",
         "a": [1, 2, 3, 4, 5], "b": [1, 2, 3, 4, 5], "expect": 55.0},
        {"cite": "tests/src/main.rs:189-193", "backend": "default", "dtype": "f32",
         "a": [0, 1, 2, 3], "b": [0, -1, -2, 3], "expect": 4.0},
        {"cite": "src/backends/rust.rs:16-19", "backend": "Rust", "dtype": "f32",
         "a": [1, 2, 3], "b": [-1, 2, -1], "expect": 0.0},
        {"cite": "src/backends/blas.rs:11-14", "backend": "Blas", "dtype": "f32",
         "a": [1, 2, 3], "b": [-1, 2, -1], "expect": 0.0},
        {"cite": "src/backends.rs:267-272", "backend": "default", "dtype": "f32",
         "a": [0, 1, 2, 3], "b": [1.2, 1.2, 1.2, 1.2], "expect": 7.2, "abs_tol": 3e-6},
    ],
    "dot_len_mismatch_panics": {"cite": "tests/src/main.rs:58-67", "a_len": 5, "b_len": 4},
    "dotu": [
        {"cite": "tests/src/main.rs:196-201", "dtype": "c64",
         "a": [[1, 2]] * 5, "b": [[1, 2]] * 5, "expect": [-15.0, 20.0]},
    ],
    "norm": [
        {"cite": "tests/src/main.rs:133-143", "backend": "Blas+Rust", "dtype": "c64",
         "a": [[1.2, 2.3]] * 2, "expect": 3.6687872},
    ],
    "ewise": [
        # src/backends/rust.rs:97-110: a = moo![T: 1..13]; a.op(&a) == a[n] op a[n], both forms,
        # f32 and f64 (12 elements: SIMD body + scalar tail for LANES=8).
        {"cite": "src/backends/rust.rs:97-110", "op": op, "dtype": dt,
         "a": [float(v) for v in range(1, 13)], "b": [float(v) for v in range(1, 13)]}
        for op in ("add", "mul", "div", "sub") for dt in ("f32", "f64")
    ],
    "transpose": [
        {"cite": "tests/src/main.rs:443-453", "a": [1, 4, 2, 5, 3, 6], "columns": 3,
         "expect": [1, 2, 3, 4, 5, 6]},
        # Matrix::deref_mut on shape [3,2] + IS_TRANS: Rust.transpose_inplace(data, axis_len(1)=2)
        {"cite": "tests/src/main.rs:427-440", "a": r16, "columns": 2, "expect": [1, 4, 2, 5, 3, 6],
         "debug": "[\n   [1.0, 4.0],\n   [2.0, 5.0],\n   [3.0, 6.0],\n]"},
    ],
    "matmul": [
        {"cite": "tests/src/main.rs:289-297; src/lib.rs:135-140", "a": mat(r16, 2, 3), "b": mat(r16, 3, 2),
         "expect": [22, 28, 49, 64]},
        {"cite": "tests/src/main.rs:300-308", "a": mat(r16, 2, 3, True), "b": mat(r16, 3, 2, True),
         "expect": [9, 19, 29, 12, 26, 40, 15, 33, 51]},
        {"cite": "tests/src/main.rs:311-319", "a": mat(r16, 3, 2), "b": mat(r16, 3, 2, True),
         "expect": [5, 11, 17, 11, 25, 39, 17, 39, 61]},
        {"cite": "tests/src/main.rs:322-330", "a": mat(r16, 2, 3, True), "b": mat(r16, 2, 3),
         "expect": [17, 22, 27, 22, 29, 36, 27, 36, 45]},
        {"cite": "tests/src/main.rs:333-341", "a": mat(r16, 2, 3), "b": mat(r16, 2, 3, True),
         "expect": [14, 32, 32, 77]},
        {"cite": "tests/src/main.rs:344-352", "a": mat(r16, 3, 2, True), "b": mat(r16, 3, 2),
         "expect": [35, 44, 44, 56]},
        # dynamic shapes: reshape(&[3, 2]) = width 3, 2 rows (src/static_vec.rs:83-90)
        {"cite": "src/static_vec.rs:83-90", "a": {"data": [0, 1, 2, 3, 4, 5], "shape": [3, 2], "trans": False},
         "b": {"data": [0] * 6, "shape": [2, 3], "trans": False}, "expect": [0, 0, 0, 0]},
    ],
    "matmul_wrong_size_panics": {"cite": "tests/src/main.rs:455-469", "a_len": 5, "m": 2, "k": 3},
    "gemv": [
        {"cite": "src/lib.rs:138-141", "a": mat(r16, 3, 2), "x": [1, 2], "expect": [5, 11, 17]},
        # gemv == gemm with n = 1 (tests/src/main.rs:355-394); the reference asserts c == d only, the
        # numeric values here are the (exact) products, checked against numpy in test_oracle_golden.py
        {"cite": "tests/src/main.rs:355-366", "a": mat(r16, 2, 3), "x": [1, 2, 3], "expect": [14, 32]},
        {"cite": "tests/src/main.rs:369-380", "a": mat(r16, 3, 2, True), "x": [1, 2, 3], "expect": [22, 28]},
        {"cite": "tests/src/main.rs:383-394", "a": mat([float(v) for v in range(2, 14)], 4, 3, True),
         "x": [2, 3, 4, 5], "expect": [106, 120, 134]},
    ],
    "index": [
        {"cite": "src/lib.rs:166-168", "kind": "tensor", "shape": [3, 3, 3], "index": [0, 0, 1], "expect": 9},
        {"cite": "src/lib.rs:152-159", "kind": "matrix_vs_tensor", "shape": [3, 2],
         "tensor_index": [0, 1], "matrix_rc": [1, 0]},
        {"cite": "src/lib.rs:170-173", "kind": "index_slice", "shape": [3, 3, 3], "slice": 1,
         "offset": 9, "sub_shape": [3, 3], "matrix_rc": [0, 0], "expect_value": 9,
         "transposed_rc": [1, 0], "expect_transposed_value": 10},
        {"cite": "tests/src/main.rs:397-415", "kind": "lazy_transpose", "shape": [3, 2],
         "pairs": [[[1, 0], [0, 1]], [[0, 2], [2, 0]]]},
        {"cite": "tests/src/main.rs:479-485", "kind": "out_of_bounds", "shape": [3, 3, 3], "index": [3, 0, 0]},
        {"cite": "tests/src/main.rs:487-497", "kind": "index_slice_value", "shape": [3, 3], "slice": 1,
         "offset": 3, "expect_value": 3},
        {"cite": "tests/src/main.rs:537-548", "kind": "index_slice_oob", "shape": [3, 3, 3], "slice": 3},
    ],
    "shape": [
        {"cite": "tests/src/main.rs:566-581; src/tensor.rs:12-28", "m": 2, "k": 1, "slice": [1, 2],
         "axis_len": [1, 2], "volume": 2},
        {"cite": "src/tensor.rs:12-16", "m": 2, "k": 3, "slice": [3, 2], "axis_len": [3, 2], "volume": 6},
    ],
}

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")
    with open(out, "w") as f:
        json.dump(G, f, indent=1)
    print("wrote", out)
