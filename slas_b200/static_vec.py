"""Storage layer of the Cuda backend: the host-side mirror of the reference's `StaticVec`
contract plus the device-resident implementors north_star asks for.

Reference: a `StaticVec<T, LEN>` gives a backend exactly one thing, `as_ptr()` to LEN contiguous
elements (src/static_vec.rs:121-134); `[T; LEN]`, `StaticVecUnion` and `StaticCowVec`
(src/lib.rs:329-456) all implement it over HOST memory.  Here:

* a C-contiguous 1-D numpy array plays `[T; LEN]` / `StaticVecUnion` (host pointer; the C ABI
  stages it);
* `CudaVec` is the new device-resident implementor: `as_ptr()` is a device pointer and the host
  accessors raise, exactly like `NullVec` (src/nullvec.rs:24-57) -- data only moves when the
  caller says so (`from_host` / `to_host`);
* `CudaBatch` is `[[T; LEN]; batch]` contiguous in HBM: the "batched view over many same-shape
  objects" whose static shape picks the kernel template.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_DTYPES = {np.dtype(np.float32): "s", np.dtype(np.float64): "d",
           np.dtype(np.complex64): "c", np.dtype(np.complex128): "z"}


def prefix(dtype) -> str:
    """BLAS-style type letter of the C ABI entry points."""
    try:
        return _DTYPES[np.dtype(dtype)]
    except KeyError:
        raise TypeError(f"slas backends are registered for f32, f64, Complex<f32>, Complex<f64> "
                        f"(src/backends/rust.rs:190-193), not {dtype}") from None


class CudaVec:
    """Device-resident `StaticVec<T, LEN>`.  Owns its allocation unless built as a view."""

    def __init__(self, length: int, dtype=np.float32, *, device: int | None = None, _ptr: int | None = None, _owner=None):
        self.dtype = np.dtype(dtype)  # any Copy element may be stored; arithmetic ops check the type
        self.LEN = int(length)
        self._owner = _owner
        if _ptr is None:
            p = C.c_void_p()
            if device is None:
                _lib.call("slas_cuda_malloc", C.byref(p), self.nbytes)
            else:  # several devices in one process (backends.init_devices): storage on a chosen one
                _lib.call("slas_cuda_malloc_on", int(device), C.byref(p), self.nbytes)
            self._ptr = p.value
            self._owns = True
        else:
            self._ptr = int(_ptr)
            self._owns = False

    # -- StaticVec contract -------------------------------------------------------------------
    def as_ptr(self) -> int:
        return self._ptr

    as_mut_ptr = as_ptr

    @property
    def nbytes(self) -> int:
        return self.LEN * self.dtype.itemsize

    def __len__(self) -> int:
        return self.LEN

    # host accessors must not be used on device storage (precedent: src/nullvec.rs:24-57)
    def moo_ref(self):
        raise RuntimeError("CudaVec lives in HBM: host access is not available (use to_host())")

    mut_moo_ref = get_unchecked = __getitem__ = __iter__ = moo_ref

    # -- explicit data movement -----------------------------------------------------------------
    @classmethod
    def from_host(cls, array, device: int | None = None) -> "CudaVec":
        a = np.ascontiguousarray(array)
        v = cls(a.size, a.dtype, device=device)
        _lib.call("slas_cuda_memcpy_h2d", v._ptr, a.ctypes.data, a.nbytes)
        return v

    def copy_from_host(self, array) -> None:
        a = np.ascontiguousarray(array, dtype=self.dtype)
        assert a.size == self.LEN, f"expected {self.LEN} elements, found {a.size}"
        _lib.call("slas_cuda_memcpy_h2d", self._ptr, a.ctypes.data, a.nbytes)

    def to_host(self) -> np.ndarray:
        out = np.empty(self.LEN, dtype=self.dtype)
        _lib.call("slas_cuda_memcpy_d2h", out.ctypes.data, self._ptr, self.nbytes)
        return out

    def zero_(self) -> "CudaVec":
        _lib.call("slas_cuda_memset", self._ptr, 0, self.nbytes)
        return self

    def view(self, offset: int, length: int) -> "CudaVec":
        """Non-owning sub-vector (static_slice_unchecked, src/static_vec.rs:160-175)."""
        assert 0 <= offset and offset + length <= self.LEN
        return CudaVec(length, self.dtype, _ptr=self._ptr + offset * self.dtype.itemsize, _owner=self)

    def free(self) -> None:
        if self._owns and self._ptr:
            _lib.call("slas_cuda_free", self._ptr)
            self._ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # -- the sugar of StaticVec (src/static_vec.rs:62-118, 211-222) ------------------------------
    def static_backend(self, backend=None):
        from .backends import Cuda, WithStaticBackend
        return WithStaticBackend(self, backend or Cuda())

    def matrix(self, m: int, k: int, backend=None):
        from .tensor import matrix
        return matrix(self, m, k, backend)

    def reshape(self, shape, backend=None):
        from .tensor import reshape
        return reshape(self, shape, backend)


class CudaBatch(CudaVec):
    """`[[T; LEN]; batch]` contiguous in HBM."""

    def __init__(self, batch: int, length: int, dtype=np.float32, *, _ptr=None, _owner=None):
        self.batch = int(batch)
        self.OBJ_LEN = int(length)
        super().__init__(self.batch * self.OBJ_LEN, dtype, _ptr=_ptr, _owner=_owner)

    @classmethod
    def from_host(cls, array, length: int | None = None) -> "CudaBatch":
        a = np.ascontiguousarray(array)
        if length is None:
            assert a.ndim >= 2, "give `length` for a flat array"
            length = int(np.prod(a.shape[1:]))
        assert a.size % length == 0
        v = cls(a.size // length, length, a.dtype)
        _lib.call("slas_cuda_memcpy_h2d", v._ptr, a.ctypes.data, a.nbytes)
        return v

    def to_host(self) -> np.ndarray:
        return super().to_host().reshape(self.batch, self.OBJ_LEN)

    def object(self, i: int) -> CudaVec:
        assert 0 <= i < self.batch
        return self.view(i * self.OBJ_LEN, self.OBJ_LEN)

    def shard(self, nranks: int, rank: int) -> "CudaBatch":
        """Contiguous batch-index slice owned by `rank` (no collective needed)."""
        lo, hi = self.batch * rank // nranks, self.batch * (rank + 1) // nranks
        return CudaBatch(hi - lo, self.OBJ_LEN, self.dtype,
                         _ptr=self._ptr + lo * self.OBJ_LEN * self.dtype.itemsize, _owner=self)


class StaticCowVec:
    """`StaticCowVec<'a, T, LEN>` (src/lib.rs:395-456; its StaticVec impl src/static_vec.rs:294-319): borrows its
    source -- a host array or a device-resident CudaVec -- until the first MUTABLE access (`as_mut_ptr` /
    `mut_moo_ref`), which copies it; `as_ptr` / `moo_ref` never copy."""

    def __init__(self, source, *, owned: bool = False):
        assert isinstance(source, (np.ndarray, CudaVec)), "a StaticCowVec wraps a host array or a CudaVec"
        self._data = source
        self._owned = bool(owned)

    @classmethod
    def from_owned(cls, data) -> "StaticCowVec":
        return cls(data, owned=True)

    def is_owned(self) -> bool:
        return self._owned

    def is_borrowed(self) -> bool:
        return not self._owned

    @property
    def dtype(self):
        return self._data.dtype

    @property
    def LEN(self) -> int:
        return self._data.LEN if isinstance(self._data, CudaVec) else self._data.size

    def __len__(self) -> int:
        return self.LEN

    def moo_ref(self):
        return self._data

    def mut_moo_ref(self):
        """deref_mut (src/lib.rs:446-455): `self.data.owned = *self.data.borrowed` on first use."""
        if not self._owned:
            if isinstance(self._data, CudaVec):
                copy = CudaVec(self._data.LEN, self._data.dtype)
                _lib.call("slas_cuda_memcpy_d2d", copy.as_ptr(), self._data.as_ptr(), self._data.nbytes)
                self._data = copy
            else:
                self._data = np.array(self._data, copy=True)
            self._owned = True
        return self._data

    def adopt(self, owned_data) -> None:
        """Become owned with `owned_data` as the storage (a fused op already produced the mutated copy)."""
        self._data = owned_data
        self._owned = True

    def as_ptr(self) -> int:
        return operand(self._data)[0]

    def as_mut_ptr(self) -> int:
        return operand(self.mut_moo_ref(), writable=True)[0]

    def to_host(self) -> np.ndarray:
        return self._data.to_host() if isinstance(self._data, CudaVec) else np.array(self._data, copy=True)


def operand(x, dtype=None, *, writable: bool = False, any_copy_type: bool = False):
    """(pointer, length, dtype, keepalive) of anything that satisfies the StaticVec contract.
    any_copy_type: accept any fixed-size element (Transpose is generic over T: Copy), not only the four
    number types the arithmetic backends are registered for."""
    if isinstance(x, StaticCowVec):  # a mutable access copies a borrowed vector first (static_vec.rs:304-318)
        return operand(x.mut_moo_ref() if writable else x.moo_ref(), dtype, writable=writable, any_copy_type=any_copy_type)
    if isinstance(x, CudaVec):
        if dtype is not None and x.dtype != np.dtype(dtype):
            raise TypeError(f"operand is {x.dtype}, expected {np.dtype(dtype)}")
        return x.as_ptr(), x.LEN, x.dtype, x
    if hasattr(x, "data") and hasattr(x, "backend") and not isinstance(x, np.ndarray):  # WithStaticBackend
        return operand(x.data, dtype, writable=writable, any_copy_type=any_copy_type)
    if isinstance(x, np.ndarray):
        if not x.flags.c_contiguous:
            raise ValueError("a StaticVec is LEN contiguous elements (src/static_vec.rs:121-126)")
        if writable and not x.flags.writeable:
            # `&T` implementors panic on as_mut_ptr (src/static_vec.rs:252-258)
            raise ValueError("cannot take a mutable pointer to a read-only vector")
        if dtype is not None and x.dtype != np.dtype(dtype):
            raise TypeError(f"operand is {x.dtype}, expected {np.dtype(dtype)}")
        if not any_copy_type:
            prefix(x.dtype)
        return x.ctypes.data, x.size, x.dtype, x
    if writable:
        raise TypeError("output operands must be numpy arrays or CudaVec (need a stable pointer)")
    a = np.ascontiguousarray(x, dtype=dtype if dtype is not None else np.float32)
    return a.ctypes.data, a.size, a.dtype, a


def moo(dtype, values, on=None):
    """`moo![T: ...]` / `moo![on B: T: ...]` (src/lib.rs:293-325): a host vector, optionally
    bound to a backend; `on=Cuda` with `device=True` semantics is spelt CudaVec.from_host."""
    if isinstance(values, range):
        values = list(values)
    a = np.array(values, dtype=dtype)
    assert a.ndim == 1
    if on is None:
        return a
    from .backends import WithStaticBackend
    return WithStaticBackend(a, on() if isinstance(on, type) else on)
