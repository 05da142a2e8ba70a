"""Shaped views over a StaticVec: host-side mirror of src/tensor.rs for the hot path.

Everything here is index arithmetic and argument derivation -- the part of the reference that
must stay on the host and stay bit-identical:

* `Shape`: axis 0 is the FASTEST axis (width); `MatrixShape<M, K>` is M rows x K columns with
  `slice() == [K, M]`; a `(rows, cols)` tuple has axis_len(0) == cols (src/tensor.rs:8-115).
* `tensor_index`: sum of i_n * prod_{m<n} len_m with a bounds assert (src/tensor.rs:202-219).
* `Matrix`: lazy transpose as a flag that moves no data and is consumed as a_trans/b_trans by
  matrix_mul (src/tensor.rs:476-523); `matrix_mul_buffer` derives (m, n, k, lda, ldb, ldc)
  exactly as src/tensor.rs:376-391 and makes ONE backend call.

The arithmetic itself is the backend's (Cuda): `Matrix.matrix_mul` forwards to
`backend.matrix_mul`, the batched view to `backend.matrix_mul_batched`.
"""
from __future__ import annotations

import numpy as np

from .static_vec import CudaBatch, CudaVec, operand


# ---- Shape ------------------------------------------------------------------------------------------
class MatrixShape:
    """`MatrixShape<M, K>`: M rows, K columns (src/tensor.rs:86-115)."""

    def __init__(self, m: int, k: int):
        self.M, self.K = int(m), int(k)

    def axis_len(self, n: int) -> int:
        if n == 0:
            return self.K
        if n == 1:
            return self.M
        raise IndexError("Cannot get len of axis higher than 1, as a matrix only has 2 axies (rows and columns)")

    def volume(self) -> int:
        return self.M * self.K

    def slice(self):
        return [self.K, self.M]

    ndim = 2

    def __repr__(self):
        return f"MatrixShape<{self.M}, {self.K}>"


class _ListShape:
    """`[usize; NDIM]` (src/tensor.rs:22-40): axis n has length shape[n]."""

    def __init__(self, dims):
        self.dims = [int(d) for d in dims]
        self.ndim = len(self.dims)

    def axis_len(self, n: int) -> int:
        return self.dims[n]

    def volume(self) -> int:
        p = 1
        for d in self.dims:
            p *= d
        return p

    def slice(self):
        return self.dims

    def __repr__(self):
        return repr(self.dims)


class _TupleShape:
    """`(usize, usize)` as a shape: (rows, cols) (src/tensor.rs:58-78)."""

    ndim = 2

    def __init__(self, rows, cols):
        self.t = (int(rows), int(cols))

    def axis_len(self, n: int) -> int:
        if n == 0:
            return self.t[1]
        if n == 1:
            return self.t[0]
        raise IndexError(f"Index {n} out of range 2 (when trying to index into Shape)")

    def volume(self) -> int:
        return self.t[0] * self.t[1]

    def slice(self):
        raise NotImplementedError("unimplemented!() in the reference (src/tensor.rs:75-77)")

    def __repr__(self):
        return repr(self.t)


def as_shape(shape):
    if isinstance(shape, (MatrixShape, _ListShape, _TupleShape)):
        return shape
    if isinstance(shape, tuple):
        assert len(shape) == 2, "only (rows, cols) tuples are shapes"
        return _TupleShape(*shape)
    return _ListShape(shape)


def debug_shape(s, ndim=None) -> str:
    n = ndim if ndim is not None else s.ndim
    return ", ".join(str(s.axis_len(i)) for i in range(n))


def tensor_index(shape, index) -> int:
    """src/tensor.rs:202-219."""
    shape = as_shape(shape)
    index = list(index)
    assert len(index) == shape.ndim, f"index of {len(index)} axes into a {shape.ndim}-d tensor"
    total, product = 0, 1
    for n in range(shape.ndim):
        i, j = int(index[n]), shape.axis_len(n)
        assert 0 <= i < j, f"Index [{', '.join(map(str, index))}] out of bounds [{debug_shape(shape)}]"
        total += i * product
        product *= j
    return total


# ---- Tensor -----------------------------------------------------------------------------------------
def _host(data) -> np.ndarray:
    if isinstance(data, CudaVec):
        data.moo_ref()  # raises: device storage has no host accessors
    return data


class Tensor:
    """`Tensor<T, U, B, NDIM, LEN, S>` (src/tensor.rs:121-132): data bound to a backend + a shape."""

    def __init__(self, data, shape, backend):
        from .backends import WithStaticBackend
        self.data = data if isinstance(data, WithStaticBackend) else WithStaticBackend(data, backend)
        self.shape = as_shape(shape)

    @property
    def backend(self):
        return self.data.backend

    @property
    def LEN(self) -> int:
        return self.data.LEN

    @property
    def NDIM(self) -> int:
        return self.shape.ndim

    def _flat(self, i) -> int:
        if isinstance(i, tuple):  # (row, col) on a 2-d tensor -> [col, row] (src/tensor.rs:263-268)
            assert self.shape.ndim == 2 and len(i) == 2
            return tensor_index(self.shape, [i[1], i[0]])
        return tensor_index(self.shape, i)

    def __getitem__(self, i):
        return _host(self.data.data)[self._flat(i)]

    def __setitem__(self, i, value):
        _host(self.data.data)[self._flat(i)] = value

    def index_slice(self, i: int) -> "Tensor":
        """src/tensor.rs:321-352: drop the last axis, offset by i * (volume / axis_len(NDIM-1));
        bounds-checks i against axis 0 (sic) and keeps over-claiming LEN like the reference."""
        nd = self.shape.ndim
        assert nd > 1
        assert i < self.shape.axis_len(0)
        off = i * (self.shape.volume() // self.shape.axis_len(nd - 1))
        base = self.data.data
        if isinstance(base, CudaVec):
            sub = base.view(off, base.LEN - off)
        else:
            sub = base[off:]
        return Tensor(sub, self.shape.slice()[: nd - 1], self.backend)

    def matrix_view(self) -> "Matrix":
        assert self.shape.ndim == 2
        return Matrix(self)

    def __repr__(self):
        return f"Tensor(shape=[{debug_shape(self.shape)}], backend={self.backend!r})"


# ---- Matrix ----------------------------------------------------------------------------------------
class Matrix:
    """`Matrix<T, U, B, LEN, IS_TRANS, S>` (src/tensor.rs:476-483): a 2-d tensor plus the lazy
    transpose flag."""

    def __init__(self, tensor: Tensor, is_trans: bool = False):
        assert tensor.shape.ndim == 2
        self._t = tensor
        self.IS_TRANS = bool(is_trans)

    @property
    def backend(self):
        return self._t.backend

    @property
    def shape(self):
        return self._t.shape

    @property
    def data(self):
        return self._t.data

    @property
    def LEN(self) -> int:
        return self._t.LEN

    def rows(self) -> int:  # src/tensor.rs:495-501
        return self.shape.axis_len(0) if self.IS_TRANS else self.shape.axis_len(1)

    def columns(self) -> int:  # src/tensor.rs:504-510
        return self.shape.axis_len(1) if self.IS_TRANS else self.shape.axis_len(0)

    def transpose(self) -> "Matrix":  # src/tensor.rs:512-514: flag flip, zero data movement
        return Matrix(self._t, not self.IS_TRANS)

    as_transposed = transpose

    def deref(self) -> Tensor:  # src/tensor.rs:527-544
        if self.IS_TRANS:
            raise RuntimeError("Cannot deref lazily transposed matrix immutably. Try using &self.deref_mut() instead")
        return self._t

    def deref_mut(self) -> Tensor:
        """Materialise a lazy transpose (src/tensor.rs:546-564).  The reference hard-wires `Rust`
        here; the Cuda mirror runs the same transpose_inplace on the matrix's own backend.  Only
        for dynamic `[usize; 2]` shapes, and -- like the reference -- the flag stays set."""
        if self.IS_TRANS:
            assert isinstance(self.shape, _ListShape), "deref_mut exists for S = [usize; 2] only"
            self.backend.transpose_inplace(self.data.data, self.shape.axis_len(1))
            d = self.shape.dims
            d[0], d[1] = d[1], d[0]
        return self._t

    def _flat(self, i) -> int:  # src/tensor.rs:296-301
        r, c = i
        return tensor_index(self.shape, [r, c] if self.IS_TRANS else [c, r])

    def __getitem__(self, i):
        return _host(self.data.data)[self._flat(i)]

    def __setitem__(self, i, value):
        _host(self.data.data)[self._flat(i)] = value

    # -- products: src/tensor.rs:364-465 ----------------------------------------------------------
    def matmul_args(self, other: "Matrix"):
        m, k, n = self.rows(), other.rows(), other.columns()
        return m, n, k, self.shape.axis_len(0), other.shape.axis_len(0), n

    def matrix_mul_buffer(self, other: "Matrix", buffer) -> None:
        assert other.backend == self.backend, "both matrices must be bound to the same backend"
        m, n, k, lda, ldb, ldc = self.matmul_args(other)
        assert self.shape.volume() == self.LEN
        assert other.shape.volume() == other.LEN
        olen = operand(buffer)[1]
        assert m * n == olen, f"Matrix::matrix_mul_buffer expected buffer of {m * n} elements, found one of {olen}"
        self.backend.matrix_mul(self.data.data, other.data.data, buffer, m, n, k, lda, ldb, ldc,
                                self.IS_TRANS, other.IS_TRANS)

    def matrix_mul(self, other: "Matrix"):
        """Returns `[T; OLEN]` zero-initialised then filled (src/tensor.rs:441-454); device
        operands give a CudaVec."""
        m, n = self.rows(), other.columns()
        dt = operand(self.data.data)[2]
        if isinstance(self.data.data, CudaVec):
            buf = CudaVec(m * n, dt).zero_()
        else:
            buf = np.zeros(m * n, dtype=dt)
        self.matrix_mul_buffer(other, buf)
        return buf

    def vector_mul_buffer(self, other, buffer) -> None:
        olen = operand(buffer)[1]
        assert self.rows() == olen, f"Matrix::vector_mul_buffer expected buffer of {self.rows()} elements, found one of {olen}"
        assert operand(other)[1] == self.columns()
        self.backend.matrix_vector_mul(self.data.data, other, buffer, self.shape.axis_len(0), self.shape.axis_len(1),
                                       self.shape.axis_len(0), self.IS_TRANS)

    def vector_mul(self, other):
        dt = operand(self.data.data)[2]
        buf = CudaVec(self.rows(), dt).zero_() if isinstance(self.data.data, CudaVec) else np.zeros(self.rows(), dtype=dt)
        self.vector_mul_buffer(other, buf)
        return buf

    def __repr__(self):  # Debug goes through Matrix::index, so it works on lazy transposes (src/tensor.rs:177-192)
        rows = [[self[(r, c)] for c in range(self.columns())] for r in range(self.rows())]
        return "[" + ", ".join("[" + ", ".join(f"{float(v):.1f}" for v in row) + "]" for row in rows) + "]"


class BatchedMatrix:
    """Batched view: `batch` dense M x K matrices, contiguous -- the static shape selects the
    kernel template.  Lazy transpose as in `Matrix`."""

    def __init__(self, data, m: int, k: int, backend, is_trans: bool = False):
        self.data, self.M, self.K, self.backend, self.IS_TRANS = data, int(m), int(k), backend, bool(is_trans)
        total = operand(data)[1]
        assert total % (self.M * self.K) == 0, f"Cannot reshape vector of {total} elements as matrices of {self.M * self.K}"
        self.batch = total // (self.M * self.K)

    def rows(self):
        return self.K if self.IS_TRANS else self.M

    def columns(self):
        return self.M if self.IS_TRANS else self.K

    def transpose(self):
        return BatchedMatrix(self.data, self.M, self.K, self.backend, not self.IS_TRANS)

    def matrix_mul_buffer(self, other: "BatchedMatrix", buffer) -> None:
        assert other.batch == self.batch and other.backend == self.backend
        m, k, n = self.rows(), other.rows(), other.columns()
        assert self.columns() == k, f"inner dimensions differ ({self.columns()} vs {k})"
        self.backend.matrix_mul_batched(self.data, other.data, buffer, m, n, k, self.IS_TRANS, other.IS_TRANS)

    def matrix_mul(self, other: "BatchedMatrix"):
        m, n = self.rows(), other.columns()
        dt = operand(self.data)[2]
        if isinstance(self.data, CudaVec):
            buf = CudaBatch(self.batch, m * n, dt)
        else:
            buf = np.zeros(self.batch * m * n, dtype=dt)
        self.matrix_mul_buffer(other, buf)
        return buf

    def materialize_transpose(self, buffer) -> None:
        """Every object transposed out of place (batched Transpose::transpose)."""
        self.backend.transpose_batched(self.data, buffer, self.M * self.K, self.M)


# ---- constructors (src/static_vec.rs:62-118) ------------------------------------------------------------
def matrix(data, m: int, k: int, backend=None) -> Matrix:
    from .backends import Cuda
    length = operand(data)[1]
    assert m * k == length, f"Cannot reshape vector of {length} elements as matrix of {m * k}"
    return Matrix(Tensor(data, MatrixShape(m, k), backend or Cuda()))


def reshape(data, shape, backend=None):
    from .backends import Cuda
    shape = as_shape(shape)
    length = operand(data)[1]
    assert shape.volume() == length, f"Cannot reshape vector with lenght {length} as {shape!r}"
    return Tensor(data, shape, backend or Cuda())
