"""`slas_backend::Cuda` -- host-side mirror of the reference's backend interface for the hot path.

The reference's plugin boundary is `trait Backend<T>` + `mod operations` with eight operation
traits (src/backends.rs:94-183).  `Cuda` implements the same methods with the same names,
argument order and argument meaning as `Rust` (src/backends/rust.rs) and `Blas`
(src/backends/blas.rs), and forwards each one to exactly one C-ABI call of libslas_cuda.so
(include/slas_cuda.h).  No arithmetic happens in Python and there is no CPU fallback: without
the built library or without a GPU every method raises.

Operands are anything that satisfies the StaticVec contract (static_vec.operand): C-contiguous
numpy arrays (host pointers, staged by the library) or `CudaVec`/`CudaBatch` (device pointers).
Where the reference panics (length mismatch, out-of-bounds) these raise AssertionError.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .static_vec import CudaBatch, CudaVec, StaticCowVec, operand, prefix

_REAL = {np.dtype(np.float32): np.float32, np.dtype(np.float64): np.float64,
         np.dtype(np.complex64): np.float32, np.dtype(np.complex128): np.float64}


class Cuda:
    """A zero-sized backend, made per call site by `Cuda()` like `B::default()`
    (src/backends/rust.rs:2-3): all state is process-global inside the library."""

    __slots__ = ()

    def __eq__(self, other):
        return isinstance(other, Cuda)

    def __hash__(self):
        return hash("slas_backend::Cuda")

    def __repr__(self):
        return "slas_backend::Cuda"

    # -- DotProduct (src/backends.rs:122-126) ---------------------------------------------------
    def dot(self, a, b):
        pa, la, dt, ka = operand(a)
        pb, lb, _, kb = operand(b, dt)
        # both operands carry the same LEN at compile time in the reference (src/backends.rs:306-316)
        assert la == lb, f"dot: vectors of different lengths ({la} vs {lb})"
        t = prefix(dt)
        if t in "cz":
            out = np.zeros(2, dtype=_REAL[dt])
            _lib.call(f"slas_cuda_{t}dotu", la, pa, pb, out.ctypes.data)
            return dt.type(complex(out[0], out[1]))
        out = np.zeros(1, dtype=dt)
        _lib.call(f"slas_cuda_{t}dot", la, pa, pb, out.ctypes.data)
        return out[0]

    # -- Normalize (src/backends.rs:128-130) ------------------------------------------------------
    def norm(self, a):
        pa, la, dt, ka = operand(a)
        name = {"s": "snrm2", "d": "dnrm2", "c": "scnrm2", "z": "dznrm2"}[prefix(dt)]
        out = np.zeros(1, dtype=_REAL[dt])
        _lib.call(f"slas_cuda_{name}", la, pa, out.ctypes.data)
        return out[0]

    def normalize(self, a) -> None:
        if isinstance(a, StaticCowVec) and a.is_borrowed() and prefix(a.dtype) in "sd":
            # the copy-on-write copy (static_vec.rs:304-318) and the normalize in one pass: the owned side is written
            # directly with source / norm(source) instead of being copied first and scaled in place afterwards
            src = a.moo_ref()
            dst = CudaVec(a.LEN, a.dtype) if isinstance(src, CudaVec) else np.empty_like(src)
            self.normalize_into(src, dst)
            a.adopt(dst)
            return
        pa, la, dt, ka = operand(a, writable=True)
        _lib.call(f"slas_cuda_{prefix(dt)}normalize", la, pa)

    # -- Addition / Subtraction / Multiplication / Divition (src/backends.rs:156-182) -------------
    def _ewise(self, op, a, b, c) -> None:
        pa, la, dt, ka = operand(a)
        pb, lb, _, kb = operand(b, dt)
        pc, lc, _, kc = operand(c, dt, writable=True)
        assert la == lb == lc, f"{op}: vectors of different lengths ({la}, {lb}, {lc})"
        t = prefix(dt)
        if t not in "sd":
            raise TypeError(f"{op} is implemented for f32 and f64 (src/backends/rust.rs:185-188)")
        _lib.call(f"slas_cuda_{t}{op}", la, pa, pb, pc)

    def add(self, a, b, c) -> None:
        self._ewise("add", a, b, c)

    def sub(self, a, b, c) -> None:
        self._ewise("sub", a, b, c)

    def mul(self, a, b, c) -> None:
        self._ewise("mul", a, b, c)

    def div(self, a, b, c) -> None:
        self._ewise("div", a, b, c)

    # -- fused chains (include/slas_cuda.h "Fused chains"): one pass over HBM, same bits as the unfused ops -------
    def ewise_dot(self, op: str, a, b, d, c=None):
        """dot(a op b, d); the intermediate goes to `c` if given, else it is never materialised."""
        pa, la, dt, ka = operand(a)
        pb, lb, _, kb = operand(b, dt)
        pd, ld, _, kd = operand(d, dt)
        pc = 0
        if c is not None:
            pc, lc, _, kc = operand(c, dt, writable=True)
            assert lc == la, f"{op}_dot: vectors of different lengths"
        assert la == lb == ld, f"{op}_dot: vectors of different lengths ({la}, {lb}, {ld})"
        t = prefix(dt)
        if t not in "sd" or op not in ("add", "sub", "mul", "div"):
            raise TypeError(f"{op}_dot is implemented for add/sub/mul/div on f32 and f64")
        out = np.zeros(1, dtype=dt)
        _lib.call(f"slas_cuda_{t}{op}_dot", la, pa, pb, pc, pd, out.ctypes.data)
        return out[0]

    def dot_norms(self, a, b):
        """(dot(a, b), norm(a), norm(b)) from one read of both vectors."""
        pa, la, dt, ka = operand(a)
        pb, lb, _, kb = operand(b, dt)
        assert la == lb, f"dot_norms: vectors of different lengths ({la} vs {lb})"
        t = prefix(dt)
        if t not in "sd":
            raise TypeError("dot_norms is implemented for f32 and f64")
        out = np.zeros(3, dtype=dt)
        _lib.call(f"slas_cuda_{t}dot_norms", la, pa, pb, out.ctypes.data)
        return out[0], out[1], out[2]

    def normalize_into(self, a, out) -> None:
        """out = a / norm(a); `a` is left untouched (what normalising a borrowed StaticCowVec does, minus the copy)."""
        pa, la, dt, ka = operand(a)
        po, lo, _, ko = operand(out, dt, writable=True)
        assert la == lo, f"normalize_into: vectors of different lengths ({la} vs {lo})"
        t = prefix(dt)
        if t not in "sd":
            raise TypeError("normalize_into is implemented for f32 and f64")
        _lib.call(f"slas_cuda_{t}normalize_into", la, pa, po)

    # -- Transpose (src/backends.rs:152-154) --------------------------------------------------------
    def transpose(self, a, buffer, columns: int) -> None:
        # `impl<T: Copy> Transpose<T>` (src/backends/rust.rs:151): any element type, moved as opaque bytes
        pa, la, dt, ka = operand(a, any_copy_type=True)
        pb, lb, _, kb = operand(buffer, dt, writable=True, any_copy_type=True)
        assert la == lb, f"transpose: buffer of {lb} elements for a vector of {la}"
        _lib.call("slas_cuda_transpose", la, dt.itemsize, pa, pb, columns)

    def transpose_inplace(self, a, columns: int) -> None:
        pa, la, dt, ka = operand(a, writable=True, any_copy_type=True)
        _lib.call("slas_cuda_transpose_inplace", la, dt.itemsize, pa, columns)

    # -- MatrixMul (src/backends.rs:132-150) --------------------------------------------------------
    def matrix_mul(self, a, b, buffer, m, n, k, lda, ldb, ldc, a_trans, b_trans) -> None:
        pa, la, dt, ka = operand(a)
        pb, lb, _, kb = operand(b, dt)
        pc, lc, _, kc = operand(buffer, dt, writable=True)
        t = prefix(dt)
        if t not in "sd":
            raise TypeError("matrix_mul is implemented for f32 and f64 (src/backends/blas.rs:172-173)")
        _lib.call(f"slas_cuda_{t}gemm", pa, pb, pc, m, n, k, lda, ldb, ldc, int(bool(a_trans)), int(bool(b_trans)))

    def matrix_vector_mul(self, a, b, buffer, m, n, lda, a_trans) -> None:
        pa, la, dt, ka = operand(a)
        pb, lb, _, kb = operand(b, dt)
        pc, lc, _, kc = operand(buffer, dt, writable=True)
        t = prefix(dt)
        if t not in "sd":
            raise TypeError("matrix_vector_mul is implemented for f32 and f64 (src/backends/blas.rs:172-173)")
        _lib.call(f"slas_cuda_{t}gemv", pa, pb, pc, m, n, lda, int(bool(a_trans)))

    # -- batched views: the same ops over [[T; LEN]; batch] -----------------------------------------
    @staticmethod
    def _batch_of(x, length, any_copy_type=False):
        p, total, dt, keep = operand(x, any_copy_type=any_copy_type)
        assert length > 0 and total % length == 0, f"{total} elements are not whole objects of {length}"
        return p, total // length, dt, keep

    def matrix_mul_batched(self, a, b, buffer, m, n, k, a_trans=False, b_trans=False) -> None:
        pa, ba, dt, ka = self._batch_of(a, m * k)
        pb, bb, _, kb = self._batch_of(b, k * n)
        pc, lc, _, kc = operand(buffer, dt, writable=True)
        assert ba == bb and lc == ba * m * n, "matrix_mul_batched: operand batches differ"
        _lib.call(f"slas_cuda_{prefix(dt)}gemm_batched", ba, pa, pb, pc, m, n, k, int(bool(a_trans)), int(bool(b_trans)))

    def matrix_vector_mul_batched(self, a, x, buffer, m, n, a_trans=False) -> None:
        pa, ba, dt, ka = self._batch_of(a, m * n)
        px, bx, _, kx = self._batch_of(x, n if a_trans else m)
        py, ly, _, ky = operand(buffer, dt, writable=True)
        assert ba == bx and ly == ba * (m if a_trans else n), "matrix_vector_mul_batched: operand batches differ"
        _lib.call(f"slas_cuda_{prefix(dt)}gemv_batched", ba, pa, px, py, m, n, int(bool(a_trans)))

    def transpose_batched(self, a, buffer, length, columns) -> None:
        pa, ba, dt, ka = self._batch_of(a, length, any_copy_type=True)
        pb, lb, _, kb = operand(buffer, dt, writable=True, any_copy_type=True)
        assert lb == ba * length
        _lib.call("slas_cuda_transpose_batched", ba, length, dt.itemsize, pa, pb, columns)

    def transpose_inplace_batched(self, a, length, columns) -> None:
        pa, ba, dt, ka = self._batch_of(a, length, any_copy_type=True)
        operand(a, writable=True, any_copy_type=True)
        _lib.call("slas_cuda_transpose_inplace_batched", ba, length, dt.itemsize, pa, columns)

    def dot_batched(self, a, b, out, length) -> None:
        pa, ba, dt, ka = self._batch_of(a, length)
        pb, bb, _, kb = self._batch_of(b, length)
        po, lo, _, ko = operand(out, dt, writable=True)
        assert ba == bb == lo
        _lib.call(f"slas_cuda_{prefix(dt)}dot_batched", ba, length, pa, pb, po)

    def norm_batched(self, a, out, length) -> None:
        pa, ba, dt, ka = self._batch_of(a, length)
        po, lo, _, ko = operand(out, dt, writable=True)
        assert ba == lo
        _lib.call(f"slas_cuda_{prefix(dt)}nrm2_batched", ba, length, pa, po)

    def normalize_batched(self, a, length) -> None:
        pa, total, dt, ka = operand(a, writable=True)
        assert total % length == 0
        _lib.call(f"slas_cuda_{prefix(dt)}normalize_batched", total // length, length, pa)

    # -- sharded single-vector ops: this rank's slice + one allreduce --------------------
    def dot_sharded(self, a_local, b_local):
        pa, la, dt, ka = operand(a_local)
        pb, lb, _, kb = operand(b_local, dt)
        assert la == lb
        out = np.zeros(1, dtype=dt)
        _lib.call(f"slas_cuda_{prefix(dt)}dot_sharded", la, pa, pb, out.ctypes.data)
        return out[0]

    def norm_sharded(self, a_local):
        pa, la, dt, ka = operand(a_local)
        out = np.zeros(1, dtype=dt)
        _lib.call(f"slas_cuda_{prefix(dt)}nrm2_sharded", la, pa, out.ctypes.data)
        return out[0]

    def normalize_sharded(self, a_local) -> None:
        pa, la, dt, ka = operand(a_local, writable=True)
        _lib.call(f"slas_cuda_{prefix(dt)}normalize_sharded", la, pa)


    # -- one process, several GPUs: slice i device-resident on its own GPU, exchange inside the kernels ---------
    @staticmethod
    def _slices(vs, dt=None):
        ops = [operand(v, dt) for v in vs]
        dt = ops[0][2]
        lens = (C.c_size_t * len(ops))(*[o[1] for o in ops])
        ptrs = (C.c_void_p * len(ops))(*[o[0] for o in ops])
        return len(ops), lens, ptrs, dt

    def dot_multi(self, a_slices, b_slices):
        n, lens, pa, dt = self._slices(a_slices)
        nb, lens_b, pb, _ = self._slices(b_slices, dt)
        assert n == nb and list(lens) == list(lens_b), "dot_multi: slices of different lengths"
        out = np.zeros(1, dtype=dt)
        _lib.call(f"slas_cuda_{prefix(dt)}dot_multi", n, lens, pa, pb, out.ctypes.data)
        return out[0]

    def norm_multi(self, a_slices):
        n, lens, pa, dt = self._slices(a_slices)
        out = np.zeros(1, dtype=dt)
        _lib.call(f"slas_cuda_{prefix(dt)}nrm2_multi", n, lens, pa, out.ctypes.data)
        return out[0]

    def normalize_multi(self, a_slices) -> None:
        n, lens, pa, dt = self._slices(a_slices)
        _lib.call(f"slas_cuda_{prefix(dt)}normalize_multi", n, lens, pa)


class WithStaticBackend:
    """`WithStaticBackend<T, U, B, LEN>` (src/backends.rs:185-249, 306-331): data + backend."""

    def __init__(self, data, backend):
        self.data = data
        self.backend = backend

    @property
    def LEN(self):
        return operand(self.data)[1]

    def __len__(self):
        return self.LEN

    def _same_backend(self, other):
        assert isinstance(other, WithStaticBackend) and other.backend == self.backend, \
            "both operands must be bound to the same backend (src/backends.rs:306-316)"

    def dot(self, other):
        self._same_backend(other)
        return self.backend.dot(self.data, other.data)

    def norm(self):
        return self.backend.norm(self.data)

    def normalize(self):
        self.backend.normalize(self.data)

    def matrix(self, m, k):
        from .tensor import matrix
        return matrix(self.data, m, k, self.backend)

    def reshape(self, shape):
        from .tensor import reshape
        return reshape(self.data, shape, self.backend)


# ---- process-level helpers (plumbing, not ops) -----------------------------------------------------
def device_count() -> int:
    n = C.c_int(0)
    try:
        _lib.call("slas_cuda_device_count", C.byref(n))
    except _lib.SlasCudaError:
        return 0
    return n.value


def init_devices(n: int = 0) -> int:
    """Several GPUs inside this one process (include/slas_cuda.h, "several GPUs inside ONE process"): initialise
    devices 0..n-1 (0 = all visible).  Returns how many are active."""
    _lib.call("slas_cuda_init_devices", int(n))
    nd, px = C.c_int(), C.c_int()
    _lib.call("slas_cuda_devices", C.byref(nd), C.byref(px))
    return nd.value


def device_info() -> dict:
    sm, cc1, cc2 = C.c_int(), C.c_int(), C.c_int()
    l2, hbm = C.c_size_t(), C.c_size_t()
    _lib.call("slas_cuda_device_info", C.byref(sm), C.byref(l2), C.byref(hbm), C.byref(cc1), C.byref(cc2))
    return {"sm_count": sm.value, "l2_bytes": l2.value, "hbm_bytes": hbm.value, "cc": (cc1.value, cc2.value)}


def set_stream(cuda_stream: int | None) -> None:
    _lib.call("slas_cuda_set_stream", cuda_stream or None)


def synchronize() -> None:
    _lib.call("slas_cuda_synchronize")


def launch_count() -> int:
    return int(_lib.lib().slas_cuda_launch_count())


def set_tunable(name: str, value: int) -> None:
    """Benchmark knob (kernel variant selection); 0 is always the shipped default."""
    _lib.call("slas_cuda_set_tunable", name.encode(), int(value))


class Graph:
    """A recorded sequence of device-pointer ops (CUDA graph): `with Graph() as g: ...ops...`, then `g.launch()`.
    Ops inside the block are recorded, not executed; scalar results must go to device memory (`*_into` forms /
    C-ABI calls with a device result pointer)."""

    def __init__(self):
        self._exec = C.c_void_p()

    def __enter__(self):
        _lib.call("slas_cuda_graph_begin")
        return self

    def __exit__(self, exc_type, exc, tb):
        if exc_type is None:
            _lib.call("slas_cuda_graph_end", C.byref(self._exec))
        else:  # close the capture, keep the original exception
            tmp = C.c_void_p()
            _lib.lib().slas_cuda_graph_end(C.byref(tmp))
            if tmp.value:
                _lib.lib().slas_cuda_graph_destroy(tmp)
        return False

    def launch(self) -> None:
        _lib.call("slas_cuda_graph_launch", self._exec)

    def __del__(self):
        try:
            if self._exec.value:
                _lib.lib().slas_cuda_graph_destroy(self._exec)
        except Exception:
            pass


def set_sgemm_path(path: str) -> None:
    """'auto' (default: 3xTF32 on the tensor cores for large products whose shape fits its tiles, fp32 FMA
    otherwise), 'ffma' (always the ascending-k fp32 FMA chain) or 'tcgen05' (tensor cores wherever possible)."""
    _lib.call("slas_cuda_set_sgemm_path", {"ffma": 0, "tcgen05": 1, "auto": 2}[path])


__all__ = ["Cuda", "WithStaticBackend", "CudaVec", "CudaBatch", "device_count", "device_info", "set_stream",
           "synchronize", "launch_count", "set_sgemm_path", "Graph", "init_devices"]
