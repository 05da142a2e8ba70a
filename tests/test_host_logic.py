"""CPU-side checks (no GPU): the host mirror of the reference interface (index math, argument
derivation, lazy transposes, panics) against the oracle and the golden vectors, and the C-ABI
library: it loads, exports every symbol include/slas_cuda.h declares, and refuses to compute
without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import oracle
from slas_b200 import _lib
from slas_b200 import tensor as T
from slas_b200.backends import Cuda, WithStaticBackend, device_count
from slas_b200.static_vec import moo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAS_GPU = device_count() > 0


class RecordingBackend:
    """Stands where `Cuda` stands and records the one call the view layer makes."""

    def __init__(self):
        self.calls = []

    def __eq__(self, other):
        return isinstance(other, RecordingBackend)

    def matrix_mul(self, a, b, buffer, m, n, k, lda, ldb, ldc, ta, tb):
        self.calls.append(("matrix_mul", m, n, k, lda, ldb, ldc, bool(ta), bool(tb)))

    def matrix_vector_mul(self, a, b, buffer, m, n, lda, ta):
        self.calls.append(("matrix_vector_mul", m, n, lda, bool(ta)))

    def transpose_inplace(self, a, columns):
        self.calls.append(("transpose_inplace", columns))
        a[:] = oracle.transpose(a, columns)


# ---- the library and its header ----------------------------------------------------------------
def _header_symbols():
    text = open(os.path.join(ROOT, "include", "slas_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(slas_cuda_\w+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    syms = _header_symbols()
    assert len(syms) >= 60
    L = C.CDLL(_lib.LIB_PATH)
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, f"declared in include/slas_cuda.h but not exported: {missing}"


def test_python_binding_covers_the_header():
    assert sorted(_lib.SIGNATURES) == _header_symbols()
    _lib.lib()  # binds every signature; AttributeError if one is missing


@pytest.mark.skipif(HAS_GPU, reason="checks the no-device behaviour")
def test_no_cpu_fallback_without_a_device():
    a = np.arange(4, dtype=np.float32)
    with pytest.raises(_lib.SlasCudaError) as e:
        Cuda().dot(a, a)
    assert e.value.status == _lib.SLAS_ERR_CUDA
    with pytest.raises(_lib.SlasCudaError):
        Cuda().add(a, a, np.zeros_like(a))
    with pytest.raises(_lib.SlasCudaError):
        T.matrix(a, 2, 2).matrix_mul(T.matrix(a, 2, 2))


def test_product_code_never_touches_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "slas_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("// oracle", ""), f"{f} mentions the oracle"


def test_shard_range_partitions_exactly():
    L = _lib.lib()
    for length in (0, 1, 5, 1000, 2**20 + 3, 2**32):
        for elem in (4, 8):
            for n in (1, 2, 4, 8):
                prev = 0
                for r in range(n):
                    lo, hi = C.c_size_t(), C.c_size_t()
                    assert L.slas_cuda_shard_range(length, elem, n, r, C.byref(lo), C.byref(hi)) == 0
                    assert lo.value == prev and hi.value >= lo.value
                    if r < n - 1:
                        assert (hi.value * elem) % 16 == 0  # 16-byte aligned interior boundaries
                    prev = hi.value
                assert prev == length


# ---- Shape / index math against the golden vectors and the oracle -------------------------------
def test_shape_golden(golden):
    for g in golden["shape"]:
        s = T.MatrixShape(g["m"], g["k"])
        assert s.slice() == g["slice"]
        assert [s.axis_len(0), s.axis_len(1)] == g["axis_len"]
        assert s.volume() == g["volume"]
    assert T.as_shape((2, 3)).axis_len(0) == 3 and T.as_shape((2, 3)).axis_len(1) == 2


def test_index_golden(golden):
    for g in golden["index"]:
        kind = g["kind"]
        if kind == "tensor":
            assert T.tensor_index(g["shape"], g["index"]) == g["expect"]
        elif kind == "matrix_vs_tensor":
            vol = int(np.prod(g["shape"]))
            t = T.reshape(np.arange(vol, dtype=np.float32), g["shape"], RecordingBackend())
            assert t[g["tensor_index"]] == t.matrix_view()[tuple(g["matrix_rc"])]
        elif kind == "index_slice":
            vol = int(np.prod(g["shape"]))
            t = T.reshape(np.arange(vol, dtype=np.float32), g["shape"], RecordingBackend())
            s = t.index_slice(g["slice"]).matrix_view()
            assert s[tuple(g["matrix_rc"])] == g["expect_value"]
            assert s.transpose()[tuple(g["transposed_rc"])] == g["expect_transposed_value"]
        elif kind == "lazy_transpose":
            vol = int(np.prod(g["shape"]))
            m = T.reshape(np.arange(vol, dtype=np.float32), g["shape"], RecordingBackend()).matrix_view()
            for (r0, c0), (r1, c1) in g["pairs"]:
                assert m[(r0, c0)] == m.transpose()[(r1, c1)]
        elif kind == "out_of_bounds":
            with pytest.raises(AssertionError):
                T.tensor_index(g["shape"], g["index"])
        elif kind == "index_slice_value":
            vol = int(np.prod(g["shape"]))
            t = T.reshape(np.arange(vol, dtype=np.float32), g["shape"], RecordingBackend())
            assert t.index_slice(g["slice"])[[0]] == g["expect_value"]
        elif kind == "index_slice_oob":
            vol = int(np.prod(g["shape"]))
            t = T.reshape(np.arange(vol, dtype=np.float32), g["shape"], RecordingBackend())
            with pytest.raises(AssertionError):
                t.index_slice(g["slice"])


def test_tensor_index_matches_oracle_randomised():
    rng = np.random.default_rng(0)
    for _ in range(200):
        nd = int(rng.integers(1, 5))
        shape = [int(rng.integers(1, 7)) for _ in range(nd)]
        idx = [int(rng.integers(0, s)) for s in shape]
        assert T.tensor_index(shape, idx) == oracle.tensor_index(shape, idx)
        r_shape = shape[:2] if nd >= 2 else None
        if r_shape:
            for trans in (False, True):
                r, c = (idx[0], idx[1]) if trans else (idx[1], idx[0])
                m = T.Matrix(T.Tensor(np.zeros(r_shape[0] * r_shape[1], np.float32), r_shape, RecordingBackend()), trans)
                assert m._flat((r, c)) == oracle.matrix_index(r_shape, trans, r, c)


def test_matmul_args_match_oracle(golden):
    for g in golden["matmul"]:
        be = RecordingBackend()
        a, b = g["a"], g["b"]
        A = T.Matrix(T.Tensor(np.array(a["data"], np.float32), a["shape"], be), a["trans"])
        B = T.Matrix(T.Tensor(np.array(b["data"], np.float32), b["shape"], be), b["trans"])
        want = oracle.matmul_args(a["shape"], a["trans"], b["shape"], b["trans"])
        assert A.matmul_args(B) == want
        A.matrix_mul(B)
        assert be.calls == [("matrix_mul", *want, a["trans"], b["trans"])]


def test_vector_mul_passes_the_traits_swapped_m_n(golden):
    for g in golden["gemv"]:
        be = RecordingBackend()
        a = g["a"]
        A = T.Matrix(T.Tensor(np.array(a["data"], np.float32), a["shape"], be), a["trans"])
        A.vector_mul(np.array(g["x"], np.float32))
        # m = axis_len(0) (underlying width), n = axis_len(1) (underlying height), lda = width
        assert be.calls == [("matrix_vector_mul", a["shape"][0], a["shape"][1], a["shape"][0], a["trans"])]


def test_view_layer_panics(golden):
    g = golden["matmul_wrong_size_panics"]
    with pytest.raises(AssertionError, match="Cannot reshape vector"):
        T.matrix(np.zeros(g["a_len"], np.float32), g["m"], g["k"], RecordingBackend())
    be = RecordingBackend()
    A = T.matrix(np.zeros(6, np.float32), 2, 3, be)
    B = T.matrix(np.zeros(6, np.float32), 3, 2, be)
    with pytest.raises(AssertionError, match="expected buffer of 4 elements"):
        A.matrix_mul_buffer(B, np.zeros(5, np.float32))
    with pytest.raises(AssertionError):
        A.vector_mul_buffer(np.zeros(3, np.float32), np.zeros(3, np.float32))
    with pytest.raises(RuntimeError, match="Cannot deref lazily transposed"):
        A.transpose().deref()
    g = golden["dot_len_mismatch_panics"]
    with pytest.raises(AssertionError):
        Cuda().dot(np.zeros(g["a_len"], np.float32), np.zeros(g["b_len"], np.float32))
    with pytest.raises(AssertionError):
        WithStaticBackend(np.zeros(3, np.float32), Cuda()).dot(WithStaticBackend(np.zeros(3, np.float32), RecordingBackend()))


def test_materialised_transpose_uses_transpose_inplace(golden):
    g = [t for t in golden["transpose"] if "debug" in t][0]
    be = RecordingBackend()
    m = T.Matrix(T.Tensor(np.array(g["a"], np.float32), [3, 2], be), True)  # 1..=6 shaped [3, 2], transposed
    t = m.deref_mut()
    assert be.calls == [("transpose_inplace", 2)]  # columns = axis_len(1)
    assert t.shape.slice() == [2, 3]
    assert list(t.data.data) == g["expect"]


def test_moo_and_static_backend():
    v = moo(np.float32, range(0, 4), on=Cuda)
    assert isinstance(v, WithStaticBackend) and v.LEN == 4 and v.backend == Cuda()
    assert list(moo(np.float64, [1, 2, 3])) == [1.0, 2.0, 3.0]
    ro = np.arange(4, dtype=np.float32)
    ro.flags.writeable = False
    with pytest.raises(ValueError):
        Cuda().normalize(ro)  # `&T` implementors panic on as_mut_ptr (src/static_vec.rs:252-258)


def test_static_cow_vec_copies_on_first_mutable_access():
    """src/lib.rs:433-456 + src/static_vec.rs:294-319: as_ptr / moo_ref never copy; as_mut_ptr / mut_moo_ref copy a
    borrowed vector once, after which it is owned (the README's `moo` example, lib.rs:255-268)."""
    from slas_b200.static_vec import StaticCowVec, operand
    source = np.array([1.0, 2.0, 3.0], np.float32)
    v = StaticCowVec(source)
    assert v.is_borrowed() and not v.is_owned() and v.LEN == 3 and len(v) == 3
    assert v.as_ptr() == source.ctypes.data and v.moo_ref() is source
    assert operand(v)[0] == source.ctypes.data and v.is_borrowed()          # read access through the backend
    p = operand(v, writable=True)[0]                                         # mutable access: the copy happens here
    assert v.is_owned() and p != source.ctypes.data and p == v.as_mut_ptr() == v.as_ptr()
    v.mut_moo_ref()[0] = 9.0
    assert source[0] == 1.0 and v.moo_ref()[0] == 9.0                        # the source is untouched
    owned = StaticCowVec.from_owned(np.zeros(4, np.float64))
    q = owned.as_ptr()
    assert owned.is_owned() and owned.as_mut_ptr() == q                      # an owned vector is never copied
    ro = np.arange(3, dtype=np.float32)
    ro.flags.writeable = False
    w = StaticCowVec(ro)                                                     # borrowing an immutable source is the point
    assert operand(w, writable=True)[3].flags.writeable and ro[1] == 1.0


_NVRTC_CHECK = r"""
import sys
from cuda.bindings import nvrtc
text = open(sys.argv[1]).read()
src = text[text.index('R"SLASJIT(') + len('R"SLASJIT('):text.rindex(')SLASJIT"')]
exprs = sys.argv[2:]
for expr in exprs:
    err, prog = nvrtc.nvrtcCreateProgram(src.encode(), b"gemm_small_jit.cu", 0, [], [])
    assert err == nvrtc.nvrtcResult.NVRTC_SUCCESS
    nvrtc.nvrtcAddNameExpression(prog, expr.encode())
    opts = [b"--gpu-architecture=sm_100a", b"--std=c++17", b"-default-device"]
    err, = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    _, n = nvrtc.nvrtcGetProgramLogSize(prog)
    log = b" " * n
    nvrtc.nvrtcGetProgramLog(prog, log)
    assert err == nvrtc.nvrtcResult.NVRTC_SUCCESS, expr + "\n" + log.decode()[:2000]
    err, size = nvrtc.nvrtcGetCUBINSize(prog)
    assert err == nvrtc.nvrtcResult.NVRTC_SUCCESS and size > 0
    nvrtc.nvrtcDestroyProgram(prog)
print("ok", len(exprs))
"""


def test_run_time_specialised_kernel_text_compiles_for_sm_100a():
    """gemm_jit.cu and staged_jit.cu hand NVRTC kernel headers embedded as text by embed_jit_source.py at build time.  nvcc
    never sees gemm_rect_kernel and only some instantiations of the others, and a host include slipping into one of those
    headers would not break the build -- only silently send such shapes to the slower run-time-shape kernels on the GPU
    box.  So specialisations of each are compiled here (NVRTC needs no GPU) -- in a child process: importing the CUDA Python
    bindings into this one breaks a later `import torch`."""
    import importlib.util
    import subprocess
    import sys
    if importlib.util.find_spec("cuda") is None or importlib.util.find_spec("cuda.bindings") is None:
        pytest.skip("cuda-python is not installed")
    cases = {
        "gemm_small_jit_src.inc": [
            "slas::gemm_small_kernel<float, slas::SmallCfg<float, 8, 4, 16, 1, 4, 256, 32, 3, false>, false, true>",
            "slas::gemm_small_kernel<double, slas::SmallCfg<double, 12, 12, 12, 3, 6, 256, 32, 2, false>, true, false>",
            "slas::gemm_small_kernel<float, slas::SmallCfg<float, 20, 20, 20, 5, 4, 160, 16, 3, true>, false, false>",
            "slas::gemm_tiny_kernel<float, 5, 5, 5, 256>", "slas::gemm_tiny_kernel<double, 3, 7, 5, 128>",
            "slas::gemm_rect_kernel<float, 9, 9, 9, 3, 3, 28>", "slas::gemm_rect_kernel<double, 40, 40, 40, 5, 2, 1>"],
        # staged_jit.cu: async.cuh + divby.cuh + staged_objects_kernel.cuh (short-vector reductions, small transposes)
        "staged_jit_src.inc": [
            "slas::batched_reduce_staged_kernel<float, 2, 30u>", "slas::batched_reduce_staged_kernel<double, 0, 9u>",
            "slas::batched_reduce_staged_kernel<float, 1, 127u>",
            "slas::transpose_staged_kernel<uint64_t, 15u, 5u, 3u, true>", "slas::transpose_staged_kernel<uint4, 6u, 2u, 3u, true>"],
    }
    for name, exprs in cases.items():
        inc = os.path.join(ROOT, "slas_b200", "lib", "obj", name)
        assert os.path.exists(inc), "build the library first (__graft_entry__.build())"
        r = subprocess.run([sys.executable, "-c", _NVRTC_CHECK, inc] + exprs, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and r.stdout.startswith("ok"), name + "\n" + r.stdout + r.stderr


def test_c99_program_links_and_runs(tmp_path):
    """include/slas_cuda.h is C99-clean and libslas_cuda.so links from plain C (the reference is compiled code: a C / Rust
    / C++ host must be able to bind the boundary without a C++ compiler)"""
    import subprocess
    exe = tmp_path / "abi_smoke"
    src = os.path.join(ROOT, "tests", "c", "abi_smoke.c")
    libdir = os.path.join(ROOT, "slas_b200", "lib")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", str(exe),
                    "-L", libdir, "-lslas_cuda", f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "C_ABI_OK" in r.stdout, (r.returncode, r.stdout, r.stderr)
