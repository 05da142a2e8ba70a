"""GPU tests added in round 2 (run on the B200 box: `pytest -m gpu`): scratch ownership across the host
pipeline's streams, graphs that outlive a workspace growth, and the newer entry points.  Same bars as
tests/test_cuda_parity.py: bit-exact for data movement, f32 1e-5 * sqrt(K) / f64 1e-13 * sqrt(K) relative to
sum|a_i b_i| for reductions, against the oracle or an f64 / long-double truth."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle
from slas_b200 import _lib as L
from slas_b200.backends import Cuda, Graph, device_count, launch_count, set_sgemm_path, set_tunable
from slas_b200.static_vec import CudaBatch, CudaVec

pytestmark = pytest.mark.gpu
cuda = Cuda()


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert device_count() > 0, "the gpu-marked tests need a CUDA device (there is no CPU fallback)"
    before = launch_count()
    yield
    assert launch_count() > before, "no kernel of libslas_cuda was launched"


def _rand(rng, n, dt, lo=-1.0, hi=1.0):
    return (rng.random(n) * (hi - lo) + lo).astype(dt)


class Pinned:
    """numpy view of page-locked host memory from slas_cuda_malloc_host."""

    def __init__(self, n, dt):
        self.p = C.c_void_p()
        self.nbytes = n * np.dtype(dt).itemsize
        L.call("slas_cuda_malloc_host", C.byref(self.p), self.nbytes)
        ctype = {np.dtype(np.float32): C.c_float, np.dtype(np.float64): C.c_double, np.dtype(np.uint8): C.c_uint8}[np.dtype(dt)]
        self.a = np.ctypeslib.as_array((ctype * n).from_address(self.p.value))

    def free(self):
        del self.a
        L.call("slas_cuda_free_host", self.p.value)


def _gemm_truth(A, B, m, n, k):
    a = A.reshape(m, k).astype(np.float64)
    b = B.reshape(k, n).astype(np.float64)
    return (a @ b).ravel(), (np.abs(a) @ np.abs(b)).ravel()


# ---- consecutive chunks of the host pipeline run on different streams but share the workspace ----------
@pytest.mark.parametrize("dt,size,path", [(np.float32, 512, "tcgen05"), (np.float32, 1024, "ffma"), (np.float64, 1024, "auto")])
def test_host_pipeline_one_object_per_chunk_keeps_its_scratch(dt, size, path):
    """pinned host operands, one product per chunk (slot smaller than an object), every product uses the shared
    workspace (3xTF32 split / pre-transposed operands): chunk i + 1 must not overwrite it under chunk i's kernel"""
    rng = np.random.default_rng(31)
    batch, e = 6, size * size
    hA, hB, hC = Pinned(batch * e, dt), Pinned(batch * e, dt), Pinned(batch * e, dt)
    hA.a[:] = _rand(rng, batch * e, dt)
    hB.a[:] = _rand(rng, batch * e, dt)
    hC.a[:] = 0
    set_tunable("host_slot_mb", 1)
    set_sgemm_path(path)
    try:
        for rep in range(3):
            cuda.matrix_mul_batched(hA.a, hB.a, hC.a, size, size, size, True, False)  # a_trans: f64 path pre-transposes A
            for i in range(batch):
                ref, mag = _gemm_truth(np.ascontiguousarray(hA.a[i * e:(i + 1) * e].reshape(size, size).T).ravel(),
                                       hB.a[i * e:(i + 1) * e], size, size, size)
                tol = (1e-5 if dt == np.float32 else 1e-13) * np.sqrt(size)
                assert np.all(np.abs(hC.a[i * e:(i + 1) * e] - ref) <= tol * mag), (rep, i)
    finally:
        set_tunable("host_slot_mb", 0)
        set_sgemm_path("auto")
        for h in (hA, hB, hC):
            h.free()


# ---- the workspace a recorded graph points at must survive later growth --------------------------------
def test_graph_survives_a_later_workspace_growth_and_growth_inside_a_capture():
    rng = np.random.default_rng(32)
    n0, n1, n2 = 512, 1024, 1536
    A0, B0 = _rand(rng, n0 * n0, np.float32), _rand(rng, n0 * n0, np.float32)
    a0, b0, c0 = CudaVec.from_host(A0), CudaVec.from_host(B0), CudaVec(n0 * n0, np.float32)
    ref0, mag0 = _gemm_truth(A0, B0, n0, n0, n0)

    def small():
        cuda.matrix_mul(a0, b0, c0, n0, n0, n0, n0, n0, n0, False, False)  # tcgen05: hi/lo split in the workspace

    set_sgemm_path("tcgen05")
    try:
        small()  # eager once: workspace at the 512^3 size
        with Graph() as g0:
            small()
        # a bigger product grows the workspace while g0 is alive: the old block must be retired, not freed
        A1, B1 = _rand(rng, n1 * n1, np.float32), _rand(rng, n1 * n1, np.float32)
        a1, b1, c1 = CudaVec.from_host(A1), CudaVec.from_host(B1), CudaVec(n1 * n1, np.float32)
        cuda.matrix_mul(a1, b1, c1, n1, n1, n1, n1, n1, n1, False, False)
        ref1, mag1 = _gemm_truth(A1, B1, n1, n1, n1)
        assert np.all(np.abs(c1.to_host() - ref1) <= 1e-5 * np.sqrt(n1) * mag1)
        # scribble over whatever a freed-and-reused block would now hold, then replay
        junk = [CudaVec(1 << 22, np.float32).zero_() for _ in range(4)]
        c0.zero_()
        g0.launch()
        assert np.all(np.abs(c0.to_host() - ref0) <= 1e-5 * np.sqrt(n0) * mag0), "replay read a freed workspace"
        del junk
        # growth INSIDE a capture: allocated then, no synchronising call, the capture stays valid
        A2, B2 = _rand(rng, n2 * n2, np.float32), _rand(rng, n2 * n2, np.float32)
        a2, b2, c2 = CudaVec.from_host(A2), CudaVec.from_host(B2), CudaVec(n2 * n2, np.float32)
        with Graph() as g2:
            cuda.matrix_mul(a2, b2, c2, n2, n2, n2, n2, n2, n2, False, False)
        g2.launch()
        ref2, mag2 = _gemm_truth(A2, B2, n2, n2, n2)
        assert np.all(np.abs(c2.to_host() - ref2) <= 1e-5 * np.sqrt(n2) * mag2)
        c0.zero_()
        g0.launch()  # still fine after two growths
        assert np.all(np.abs(c0.to_host() - ref0) <= 1e-5 * np.sqrt(n0) * mag0)
    finally:
        set_sgemm_path("auto")


def test_graph_refuses_host_operands_and_stays_usable():
    a = _rand(np.random.default_rng(33), 4096, np.float32)
    da, dc = CudaVec.from_host(a), CudaVec(4096, np.float32)
    with pytest.raises(Exception, match="graph"):
        with Graph():
            cuda.add(a, a, np.empty_like(a))  # host operands need synchronising copies
    cuda.add(da, da, dc)
    assert np.array_equal(dc.to_host(), a + a)


# ---- the automatic switch to 3xTF32 at 2^27 multiply-adds, straddled with tiny-magnitude operands ------
@pytest.mark.parametrize("scale_log2", [0, -100, -120])
def test_sgemm_auto_threshold_straddled_with_tiny_operands(scale_log2):
    """just below the threshold the product is the fp32 FMA chain, at it the tensor-core 3xTF32 path: both sides stay
    inside the 1e-5 * sqrt(K) bar, also when A is so small that the lo parts of its split are denormal (the TF32 MMA
    flushes them: the compensation term of A is lost, an error of <= 2^-12 |a||b| per product)"""
    rng = np.random.default_rng(34)
    for (m, n, k), tensor_cores in (((512, 512, 511), False), ((512, 512, 512), True)):
        A = (_rand(rng, m * k, np.float32) * np.float32(2.0 ** scale_log2)).astype(np.float32)
        B = _rand(rng, k * n, np.float32)
        c = CudaVec(m * n, np.float32)
        n0 = launch_count()
        cuda.matrix_mul(CudaVec.from_host(A), CudaVec.from_host(B), c, m, n, k, k, n, n, False, False)
        assert (launch_count() - n0 >= 2) == tensor_cores  # the 3xTF32 path is the split pass + the tcgen05 kernel
        ref, mag = _gemm_truth(A, B, m, n, k)
        got = c.to_host().astype(np.float64)
        # results that are denormal in f32 carry an absolute rounding error of half the smallest denormal per operation
        floor = 2.0 ** -149 * k
        assert np.all(np.abs(got - ref) <= 1e-5 * np.sqrt(k) * mag + floor), (m, n, k, scale_log2)


# ---- programmatic dependent launch of the launch-bound vector ops: overlap must never reorder dependent ops --------
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("pdl", [0, 1, 2])
@pytest.mark.parametrize("graph", [False, True])
def test_back_to_back_small_ops_keep_stream_order(pdl, graph, dt):
    """short dots / adds issued back to back start while their predecessor is still in its tail (pdl = 2: the caller
    vouches that only this library uses the stream; the library's own stream gets that by default).  Read-after-write,
    write-after-read and write-after-write chains must behave exactly like the serial sequence, eagerly and replayed
    from a graph; independent calls must give the bits of the same calls issued one at a time."""
    rng = np.random.default_rng(41)
    n, reps = 1 << 16, 48
    es = np.dtype(dt).itemsize
    t = "s" if dt == np.float32 else "d"
    a = _rand(rng, reps * n, dt)
    b = _rand(rng, reps * n, dt)
    d = _rand(rng, n, dt)
    da, db, dd = CudaVec.from_host(a), CudaVec.from_host(b), CudaVec.from_host(d)
    c, c2 = CudaVec(n, dt), CudaVec(n, dt)
    out = CudaVec(4 * reps, dt).zero_()
    # serial truth from this backend, one call at a time with a sync after each
    want = np.zeros(4 * reps, dt)
    for i in range(reps):
        ai, bi = a[i * n:(i + 1) * n], b[i * n:(i + 1) * n]
        ci = ai + bi                                     # RAW: add -> dot of the sum
        want[4 * i] = cuda.dot(CudaVec.from_host(ci), dd)
        want[4 * i + 1] = cuda.dot(CudaVec.from_host(ai), CudaVec.from_host(bi))  # independent
        want[4 * i + 2] = cuda.dot(CudaVec.from_host(ci * bi), dd)                # WAW on c2, then RAW
        want[4 * i + 3] = cuda.norm(CudaVec.from_host(ci))

    def chain():
        for i in range(reps):
            va, vb = da.view(i * n, n), db.view(i * n, n)
            cuda.add(va, vb, c)                                                                     # WAR: the previous dot read c
            L.call(f"slas_cuda_{t}dot", n, c.as_ptr(), dd.as_ptr(), out.as_ptr() + es * 4 * i)           # RAW on c
            L.call(f"slas_cuda_{t}dot", n, va.as_ptr(), vb.as_ptr(), out.as_ptr() + es * (4 * i + 1))      # independent of everything before
            cuda.mul(c, vb, c2)                                                                     # RAW on c, WAW on c2
            L.call(f"slas_cuda_{t}dot", n, c2.as_ptr(), dd.as_ptr(), out.as_ptr() + es * (4 * i + 2))
            L.call(f"slas_cuda_{t}nrm2", n, c.as_ptr(), out.as_ptr() + es * (4 * i + 3))

    set_tunable("pdl", pdl)
    try:
        for rep in range(3):
            out.zero_()
            if graph:
                with Graph() as g:
                    chain()
                g.launch()
                g.launch()
            else:
                chain()
            assert np.array_equal(out.to_host(), want), (pdl, graph, rep, dt)
    finally:
        set_tunable("pdl", 0)


# ---- pageable host operands (what a slas StaticVec is) go through the pinned bounce ring ---------------------------
def test_pageable_and_pinned_host_operands_give_the_same_bytes():
    """the reference hands a backend pageable memory ([T; LEN] / Vec, static_vec.rs:126): copy threads stage it
    through a pinned ring; pinned operands are copied from in place.  Same kernels, same bytes -- also for ragged last
    chunks, INOUT operands (a transpose with a tail the reference leaves untouched) and the one-slot case."""
    rng = np.random.default_rng(51)
    batch, e = 700_003, 16
    A = _rand(rng, batch * e, np.float32)
    B = _rand(rng, batch * e, np.float32)
    hA, hB, hC = Pinned(batch * e, np.float32), Pinned(batch * e, np.float32), Pinned(batch * e, np.float32)
    hA.a[:], hB.a[:] = A, B
    try:
        for slot_mb in (0, 3):  # default slots, and small ones (many chunks, every ring slot reused many times)
            set_tunable("host_slot_mb", slot_mb)
            Cp = np.zeros(batch * e, np.float32)
            cuda.matrix_mul_batched(A, B, Cp, 4, 4, 4)
            hC.a[:] = 0
            cuda.matrix_mul_batched(hA.a, hB.a, hC.a, 4, 4, 4)
            assert np.array_equal(Cp, hC.a), slot_mb
            want = oracle.gemm_batched(A[:4096 * e], B[:4096 * e], 4096, 4, 4, 4)
            assert np.max(np.abs(Cp[:4096 * e] - want)) <= 1e-5 * 2 * 16
            # transpose with len % columns != 0: the tail of every object keeps the caller's bytes (INOUT through the ring)
            ln, cols, nb = 37, 5, 300_001
            src = _rand(rng, nb * ln, np.float32)
            dst = np.full(nb * ln, 7.0, np.float32)
            cuda.transpose_batched(src, dst, ln, cols)
            ref = np.full((nb, ln), 7.0, np.float32)
            rows = ln // cols
            ref[:, :rows * cols] = src.reshape(nb, ln)[:, :rows * cols].reshape(nb, cols, rows).transpose(0, 2, 1).reshape(nb, rows * cols)
            assert np.array_equal(dst.reshape(nb, ln), ref), slot_mb
            x = _rand(rng, 9_000_001, np.float32)
            y = np.empty_like(x)
            cuda.add(x, x, y)
            assert np.array_equal(y, x + x)
    finally:
        set_tunable("host_slot_mb", 0)
        for h in (hA, hB, hC):
            h.free()


# ---- Transpose<T: Copy> for any element size (rust.rs:151): 12-byte [f32; 3], 24-byte, 6-byte, 3-byte elements --------
@pytest.mark.parametrize("elem", [12, 24, 6, 3, 20, 48])
@pytest.mark.parametrize("rows,cols,batch", [(1, 1, 1), (3, 5, 1), (33, 65, 1), (100, 257, 1), (16, 16, 300), (7, 3, 1001), (64, 40, 9)])
def test_transpose_elements_of_any_size(elem, rows, cols, batch):
    rng = np.random.default_rng(elem * 1000 + rows)
    dt = np.dtype((np.void, elem))
    raw = rng.integers(0, 256, batch * rows * cols * elem, dtype=np.uint8)
    a = raw.view(dt)                                   # batch * rows * cols opaque elements
    n = rows * cols
    want = raw.reshape(batch, cols, rows, elem).transpose(0, 2, 1, 3).reshape(-1)   # input read as cols x rows, written rows x cols
    for src in (a, CudaVec.from_host(a)):
        out = np.zeros(batch * n, dt) if isinstance(src, np.ndarray) else CudaVec(batch * n, dt).zero_()
        if batch == 1:
            cuda.transpose(src, out, cols)
        else:
            cuda.transpose_batched(src, out, n, cols)
        got = (out if isinstance(out, np.ndarray) else out.to_host()).view(np.uint8)
        assert np.array_equal(got, want), (elem, rows, cols, batch)
        # in place (Matrix::deref_mut): through the zeroed temporary, like the reference
        ip = a.copy() if isinstance(src, np.ndarray) else CudaVec.from_host(a)
        if batch == 1:
            cuda.transpose_inplace(ip, cols)
        else:
            cuda.transpose_inplace_batched(ip, n, cols)
        got = (ip if isinstance(ip, np.ndarray) else ip.to_host()).view(np.uint8)
        assert np.array_equal(got, want), ("inplace", elem, rows, cols, batch)
    # ragged: len % columns != 0 -- the tail is never written (rust.rs:168-175)
    if batch == 1 and rows * cols > 2:
        ln = n + 2
        raw2 = rng.integers(0, 256, ln * elem, dtype=np.uint8)
        out = np.full(ln * elem, 0xAB, np.uint8).view(dt)
        cuda.transpose(raw2.view(dt), out, cols)
        r2 = ln // cols
        w2 = raw2[:cols * r2 * elem].reshape(cols, r2, elem).transpose(1, 0, 2).reshape(-1)
        assert np.array_equal(out.view(np.uint8)[:r2 * cols * elem], w2) and np.all(out.view(np.uint8)[r2 * cols * elem:] == 0xAB)


def test_host_register_makes_a_callers_array_a_pinned_operand():
    rng = np.random.default_rng(61)
    n = 6_000_011
    a, b = _rand(rng, n, np.float32), _rand(rng, n, np.float32)
    want = np.empty_like(a)
    cuda.add(a, b, want)                      # pageable: through the bounce ring
    c = np.zeros_like(a)
    for arr in (a, b, c):
        L.call("slas_cuda_host_register", arr.ctypes.data, arr.nbytes)
    try:
        cuda.add(a, b, c)                     # registered in place: direct DMA
        assert np.array_equal(c, want) and np.array_equal(c, a + b)
        assert float(cuda.dot(a, b)) == float(cuda.dot(CudaVec.from_host(a), CudaVec.from_host(b)))
    finally:
        for arr in (a, b, c):
            L.call("slas_cuda_host_unregister", arr.ctypes.data)


# ---- host-bound scalars: the kernel stores them into mapped pinned memory, the host spins on the word ---------------
@pytest.mark.parametrize("dt", [np.float32, np.float64, np.complex64, np.complex128])
def test_host_result_direct_store_matches_the_copy_path(dt):
    rng = np.random.default_rng(71)
    for n in (1, 5, 4097, (1 << 20) + 3):
        if np.dtype(dt).kind == "c":
            rdt = np.float32 if dt == np.complex64 else np.float64
            a = (_rand(rng, n, rdt) + 1j * _rand(rng, n, rdt)).astype(dt)
        else:
            a = _rand(rng, n, dt)
        da = CudaVec.from_host(a)
        direct = (cuda.dot(da, da), cuda.norm(da))
        set_tunable("host_result", 1)
        try:
            copied = (cuda.dot(da, da), cuda.norm(da))
        finally:
            set_tunable("host_result", 0)
        assert direct[0].tobytes() == copied[0].tobytes() and direct[1].tobytes() == copied[1].tobytes(), (dt, n)
    # results that ARE special values come back intact (the wait falls back to the stream when a word never changes)
    z = CudaVec.from_host(np.zeros(1000, np.float32))
    assert cuda.dot(z, z) == 0.0 and cuda.norm(z) == 0.0
    nanv = CudaVec.from_host(np.array([np.nan, 1.0], np.float32))
    assert np.isnan(cuda.dot(nanv, nanv))
    inf = CudaVec.from_host(np.array([np.inf, 1.0], np.float32))
    assert np.isposinf(cuda.dot(inf, inf))


# ---- normalize of a short vector in one launch: same bits as norm + true division ------------------------------------
@pytest.mark.parametrize("dt", [np.float32, np.float64, np.complex64])
@pytest.mark.parametrize("n", [1, 3, 4, 1000, 4097, 65536, (1 << 20), (1 << 20) + 3, 1_200_000, 1_300_000])
def test_single_launch_normalize_matches_the_two_pass_path(dt, n):
    rng = np.random.default_rng(n)
    if np.dtype(dt).kind == "c":
        a = (_rand(rng, n, np.float32) + 1j * _rand(rng, n, np.float32)).astype(dt)
    else:
        a = _rand(rng, n, dt)
    one, two = CudaVec.from_host(a), CudaVec.from_host(a)
    nrm = cuda.norm(one)
    cuda.normalize(one)
    set_tunable("normalize_fused", 1)
    try:
        cuda.normalize(two)
    finally:
        set_tunable("normalize_fused", 0)
    assert np.array_equal(one.to_host(), two.to_host()), (dt, n)
    rdt = np.float32 if dt != np.float64 else np.float64
    assert np.array_equal(one.to_host().view(rdt), a.view(rdt) / nrm), (dt, n)  # true division of every real scalar by the backend's own norm
    h = a.copy()
    cuda.normalize(h)                                                           # host operand: staged, same kernel
    assert np.array_equal(h, one.to_host())
    # back to back on distinct vectors, replayed from a graph
    if n >= 4096 and np.dtype(dt).kind != "c":
        vs = [CudaVec.from_host(a * (i + 1)) for i in range(6)]
        t = "s" if dt == np.float32 else "d"
        with Graph() as g:
            for v in vs:
                L.call(f"slas_cuda_{t}normalize", n, v.as_ptr())
        g.launch()
        for i, v in enumerate(vs):
            ref = CudaVec.from_host(a * (i + 1))
            set_tunable("normalize_fused", 1)
            try:
                cuda.normalize(ref)
            finally:
                set_tunable("normalize_fused", 0)
            assert np.array_equal(v.to_host(), ref.to_host()), (dt, n, i)


# ---- small objects with the shape as a compile-time constant (staged_objects_kernel.cuh: built in / NVRTC / run-time) ----
@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("length,batch", [(1, 4096), (2, 5000), (3, 70001), (5, 4099), (6, 3000), (7, 2048), (9, 3001), (13, 1024),
                                          (19, 2000), (30, 1500), (63, 700), (127, 600)])
def test_staged_reductions_same_bits_for_every_form_of_the_shape(dt, length, batch):
    """tunable staged_jit: 2 = the length is a run-time argument, 1 = built-in constants only, 0 = NVRTC too"""
    rng = np.random.default_rng(length * 7 + batch)
    a = rng.uniform(-1, 1, batch * length).astype(dt)
    b = rng.uniform(-1, 1, batch * length).astype(dt)
    da, db = CudaVec.from_host(a), CudaVec.from_host(b)
    got = {}
    try:
        for mode in (2, 1, 0):
            set_tunable("staged_jit", mode)
            o1, o2, x = CudaVec(batch, dt), CudaVec(batch, dt), CudaVec.from_host(a)
            cuda.dot_batched(da, db, o1, length)
            cuda.norm_batched(da, o2, length)
            cuda.normalize_batched(x, length)
            got[mode] = (o1.to_host(), o2.to_host(), x.to_host())
    finally:
        set_tunable("staged_jit", 0)
    for mode in (1, 0):
        for r, w in zip(got[mode], got[2]):
            assert np.array_equal(r.view(np.uint8), w.view(np.uint8)), (mode, length, batch)
    A, B = a.reshape(batch, length).astype(np.float64), b.reshape(batch, length).astype(np.float64)
    eps = np.finfo(dt).eps
    assert np.allclose(got[0][0], (A * B).sum(1), rtol=0, atol=4 * length * eps)
    nrm = np.sqrt((A * A).sum(1))
    assert np.allclose(got[0][1], nrm, rtol=4 * length * eps, atol=0)
    assert np.allclose(got[0][2].reshape(batch, length), A / nrm[:, None], rtol=8 * length * eps, atol=eps)
    # normalize divides by exactly the bits norm returns (norm and normalize share a kernel)
    assert np.array_equal(got[0][2].reshape(batch, length), a.reshape(batch, length) / got[0][1][:, None])


@pytest.mark.gpu
@pytest.mark.parametrize("elem", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("rows,cols,batch", [(3, 3, 70001), (2, 3, 4096), (3, 2, 1001), (5, 5, 3000), (3, 5, 2050), (6, 6, 999), (7, 9, 1024),
                                             (10, 10, 600), (30, 31, 300), (1, 7, 4096), (7, 1, 4096),
                                             # input rows of a multiple of 16 words: the padded stage, filled row by row
                                             (64, 40, 300), (32, 16, 1000), (16, 64, 700), (48, 32, 260), (128, 16, 257)])
def test_staged_transposes_same_bytes_for_every_form_of_the_shape(elem, rows, cols, batch):
    rng = np.random.default_rng(elem * 100 + rows * 10 + cols)
    n = rows * cols
    raw = rng.integers(0, 256, batch * n * elem, dtype=np.uint8)
    dt = np.dtype((np.void, elem))
    want = raw.reshape(batch, cols, rows, elem).transpose(0, 2, 1, 3).reshape(-1)
    src = CudaVec.from_host(raw.view(dt))
    try:
        for mode in (2, 1, 0):
            set_tunable("staged_jit", mode)
            out = CudaVec(batch * n, dt).zero_()
            cuda.transpose_batched(src, out, n, cols)
            assert np.array_equal(out.to_host().view(np.uint8), want), (mode, elem, rows, cols, batch)
            ip = CudaVec.from_host(raw.view(dt))
            cuda.transpose_inplace_batched(ip, n, cols)
            assert np.array_equal(ip.to_host().view(np.uint8), want), ("in place", mode, elem, rows, cols, batch)
    finally:
        set_tunable("staged_jit", 0)


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("length,batch", [(2, 200_000), (3, 150_001), (7, 50_000), (13, 40_000), (16, 40_000), (19, 30_000), (64, 10_000),
                                          (100, 5_000), (1000, 600), (5000, 64), (1 << 20, 1)])
def test_normalize_divides_with_the_bits_of_an_ieee_division_over_the_whole_exponent_range(dt, length, batch):
    """divby.cuh hoists the reciprocal refinement of div.rn out of the per-element loop; zero / subnormal / huge
    operands take the plain division.  Every element must equal numpy's correctly rounded a / norm."""
    rng = np.random.default_rng(length)
    n = batch * length
    f32 = dt == np.float32
    mant = rng.uniform(1, 2, n)
    expo = rng.integers(-75, 30, n) if f32 else rng.integers(-530, 250, n)
    a = (np.ldexp(mant, expo) * rng.choice([-1.0, 1.0], n)).astype(dt)
    # sprinkle zeros, subnormals, the smallest normals and values that make the norm subnormal / huge / infinite
    special = np.array([0.0, -0.0, 1e-45, -3e-42, 1.1754944e-38, -1.1754944e-38, 2e-38, 1e-30, 3e18, -2.5e19] if f32 else
                       [0.0, -0.0, 5e-324, -3e-310, 2.2250738585072014e-308, -2.2250738585072014e-308, 1e-300, 1e-160, 1e150, -3e153], dt)
    idx = rng.integers(0, n, max(1, n // 50))
    a[idx] = special[rng.integers(0, len(special), idx.size)]
    if batch > 8:
        rows = a.reshape(batch, length)
        rows[0] = dt(1e-30 if f32 else 1e-200)   # squares underflow: a norm that is a (sub)normal tiny number or zero
        rows[1] = dt(1e-44 if f32 else 1e-320)   # all subnormal: norm 0 -> division by zero (inf / nan like the reference)
        rows[2] = dt(3e19 if f32 else 1e160)     # squares overflow: norm = inf
        rows[3, :] = 0
        rows[4] = dt(2.0) ** (-50 if f32 else -490)
        rows[5] = dt(2.0) ** (59 if f32 else 499)
    d = CudaVec.from_host(a)
    norms = CudaVec(batch, dt)
    with np.errstate(all="ignore"):
        if batch == 1:
            nrm = np.array([cuda.norm(d)], dt)
            cuda.normalize(d)
        else:
            cuda.norm_batched(d, norms, length)
            cuda.normalize_batched(d, length)
            nrm = norms.to_host()
        got = d.to_host().reshape(batch, length)
        want = a.reshape(batch, length) / nrm[:, None]
    u = np.uint32 if f32 else np.uint64
    same = (got.view(u) == want.view(u)) | (np.isnan(got) & np.isnan(want))
    assert same.all(), (length, batch, np.argwhere(~same)[:5], got[~same][:5], want[~same][:5])


@pytest.mark.gpu
def test_a_shape_is_specialised_by_nvrtc_inside_a_graph_capture():
    """first sight of a shape while a capture is open: NVRTC compile + module load must not invalidate the capture
    (relaxed capture mode), and the recorded launch of the driver-API kernel must replay"""
    rng = np.random.default_rng(77)
    length, batch = 21, 3000                       # LEN 21 f64 vectors; 3 x 7 objects of 8-byte elements: not built in
    a = rng.uniform(-1, 1, batch * length)
    x, t = CudaVec.from_host(a), CudaVec(batch * length, np.float64).zero_()
    with Graph() as g:
        cuda.transpose_batched(x, t, length, 7)    # read as 7 x 3, written 3 x 7
        cuda.normalize_batched(x, length)
    want_t = a.reshape(batch, 7, 3).transpose(0, 2, 1).reshape(-1)
    for _ in range(2):
        x.copy_from_host(a)
        t.zero_()
        g.launch()
        assert np.array_equal(t.to_host(), want_t)
        rows = a.reshape(batch, length)
        got = x.to_host().reshape(batch, length)
        nrm = np.sqrt((rows * rows).sum(1))
        assert np.allclose(got, rows / nrm[:, None], rtol=1e-13, atol=0)


@pytest.mark.gpu
def test_concurrent_first_sight_of_shapes_from_several_threads():
    """libtest-style concurrency (SURVEY 8b 'Threading') on the run-time specialisation: several threads meet new shapes at
    once -- the same one and different ones -- and every result is exact"""
    import threading
    errors = []

    def worker(seed):
        try:
            rng = np.random.default_rng(seed)
            for length in (11 + seed % 3, 23, 45):                      # two shared lengths, one per-thread group
                batch = 2048
                a = rng.integers(-4, 5, batch * length).astype(np.float64)
                a[::length] = 3.0                                       # no zero vectors
                x = CudaVec.from_host(a)
                o = CudaVec(batch, np.float64)
                cuda.norm_batched(x, o, length)
                rows = a.reshape(batch, length)
                assert np.array_equal(o.to_host(), np.sqrt((rows * rows).sum(1)))      # exact: small integers, sqrt.rn
                cuda.normalize_batched(x, length)
                assert np.array_equal(x.to_host().reshape(batch, length), rows / o.to_host()[:, None])
            r, c = 3 + seed % 2, 5
            m = rng.integers(0, 1 << 30, 1024 * r * c).astype(np.int64)
            t = CudaVec(m.size, np.int64)
            cuda.transpose_batched(CudaVec.from_host(m), t, r * c, c)
            assert np.array_equal(t.to_host(), m.reshape(1024, c, r).transpose(0, 2, 1).reshape(-1))
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(s,)) for s in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]


@pytest.mark.gpu
def test_hoisted_division_matches_ieee_division_for_every_float_numerator():
    """csrc/divby.cuh, exhaustively: for each divisor, all 2^32 float bit patterns as numerators, both forms the kernels
    use (div, and fast where ok) against the device's own `x / n`; divisors span the exponent range, both edges of the
    fast range [2^-60, 2^60], subnormals, zero, infinities, NaN"""
    rng = np.random.default_rng(5)
    mant = rng.uniform(1, 2, 40)
    rnd = (np.ldexp(mant, rng.integers(-126, 127, 40)) * rng.choice([-1.0, 1.0], 40)).astype(np.float32)
    edge = np.array([1.0, -1.0, 3.0, 0.1, 7.0e-5, np.float32(2.0) ** -60, np.nextafter(np.float32(2.0) ** -60, np.float32(0)),
                     np.nextafter(np.float32(2.0) ** -60, np.float32(1)), np.float32(2.0) ** 60, np.nextafter(np.float32(2.0) ** 60, np.float32(np.inf)),
                     np.nextafter(np.float32(2.0) ** 60, np.float32(0)), 1.1754944e-38, 1e-45, 5e-40, 0.0, -0.0, np.inf, -np.inf, np.nan,
                     3.4028235e38, 1.9999999, 0.99999994, 1.0000001, 16777216.0], np.float32)
    divisors = np.ascontiguousarray(np.concatenate([edge, rnd]))
    bad = C.c_uint64(123)
    L.call("slas_cuda_bench_divby_mismatches", divisors.ctypes.data, divisors.size, C.byref(bad))
    assert bad.value == 0, f"{bad.value} quotients differ from x / n"


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int16])
@pytest.mark.parametrize("n,batch", [(66, 7), (68, 3), (96, 40), (100, 9), (130, 2), (200, 5), (256, 3), (1000, 1), (72, 300)])
def test_transpose_inplace_square_of_any_size(dt, n, batch):
    """square objects too big for the register / staged kernels: vector tiles with predicated edges when rows are whole
    16-byte vectors (96, 200 ...), element-wise tiles otherwise (66 f32, 130 ...)"""
    rng = np.random.default_rng(n + batch)
    a = rng.integers(-30000, 30000, batch * n * n).astype(dt)
    want = a.reshape(batch, n, n).transpose(0, 2, 1).reshape(-1)
    d = CudaVec.from_host(a)
    if batch == 1:
        cuda.transpose_inplace(d, n)
    else:
        cuda.transpose_inplace_batched(d, n * n, n)
    assert np.array_equal(d.to_host(), want)
