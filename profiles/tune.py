"""Variant sweep for the benchmark knobs (slas_cuda_set_tunable): CUDA-event timings on one B200.
Usage: python profiles/tune.py [group ...]   groups: misc rect gemv bdot transpose inplace crossover e2e dgemm fused
gemm32d gemm32s gemm16s gemm64 large tc (see GROUPS at the bottom); the r1*_tune_*.log files next to this script are
its output on a B200."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slas_b200 as sl  # noqa: E402
from slas_b200 import _lib  # noqa: E402
from slas_b200.backends import set_tunable  # noqa: E402
from slas_b200.static_vec import CudaVec  # noqa: E402

PEAK = 6539.5
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
sl.set_stream(stream.cuda_stream)
cuda = sl.Cuda()


def uniform(n, dt, seed):
    t = torch.empty(n, device=dev, dtype=torch.float32 if dt == np.float32 else torch.float64)
    _lib.call("slas_cuda_sfill_uniform" if dt == np.float32 else "slas_cuda_dfill_uniform", n, t.data_ptr(), seed, 0, -1.0, 1.0)
    return t, CudaVec(n, dt, _ptr=t.data_ptr(), _owner=t)


def timed(fn, steps=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def report(name, ms, nbytes=None, flop=None):
    row = {"name": name, "ms": round(ms, 4)}
    if nbytes:
        row["gbs"] = round(nbytes / ms / 1e6, 1)
        row["hbm_frac"] = round(nbytes / ms / 1e6 / PEAK, 4)
    if flop:
        row["tflops"] = round(flop / ms / 1e9, 2)
    print(json.dumps(row), flush=True)


def g_transpose():
    for size, dt in ((32, np.float64), (32, np.float32), (16, np.float64), (16, np.float32), (64, np.float32), (8, np.float32)):
        b = 1 << 20
        e = size * size
        (_, a), (_, c) = uniform(b * e, dt, 1), uniform(b * e, dt, 2)
        for v in (0, 1, 9, 10):
            set_tunable("transpose_variant", v)
            report(f"transpose {size}x{size} {np.dtype(dt).name} variant {v}", timed(lambda: cuda.transpose_batched(a, c, e, size)),
                   nbytes=2.0 * b * e * np.dtype(dt).itemsize)
        set_tunable("transpose_variant", 0)
    # 4x4 f32 (thread per matrix)
    b, e = 1 << 24, 16
    (_, a), (_, c) = uniform(b * e, np.float32, 1), uniform(b * e, np.float32, 2)
    for v in (0, 1, 7):
        set_tunable("transpose_variant", v)
        report(f"transpose 4x4 float32 2^24 variant {v}", timed(lambda: cuda.transpose_batched(a, c, e, 4)), nbytes=2.0 * b * e * 4)
    set_tunable("transpose_variant", 0)


def g_inplace():
    """transpose_inplace: thread / warp per object, diagonal tile-pair exchange, temporary fallback; narrow elements"""
    for size, dt, log2 in ((4, np.float32, 24), (16, np.float32, 20), (32, np.float32, 20), (32, np.float64, 20), (64, np.float32, 18),
                           (128, np.float32, 16)):
        b, e = 1 << log2, size * size
        (_, a) = uniform(b * e, dt, 1)
        for v in (0, 10):
            set_tunable("transpose_variant", v)
            report(f"transpose_inplace {size}x{size} {np.dtype(dt).name} batch 2^{log2} variant {v}",
                   timed(lambda: cuda.transpose_inplace_batched(a, e, size)), nbytes=2.0 * b * e * np.dtype(dt).itemsize)
        set_tunable("transpose_variant", 0)
    n = 8192
    (t, a) = uniform(n * n, np.float32, 1)
    report(f"transpose_inplace {n}x{n} float32 single", timed(lambda: cuda.transpose_inplace(a, n)), nbytes=8.0 * n * n)
    a8 = CudaVec(n * n // 2, np.float64, _ptr=t.data_ptr(), _owner=t)
    report(f"transpose_inplace {n // 2}x{n} float64 single (temporary)", timed(lambda: cuda.transpose_inplace(a8, n)), nbytes=8.0 * n * n)
    n8 = 5792  # 5792^2 f64 = 268 MB
    (t8, a64) = uniform(n8 * n8, np.float64, 2)
    report(f"transpose_inplace {n8}x{n8} float64 single", timed(lambda: cuda.transpose_inplace(a64, n8)), nbytes=16.0 * n8 * n8)
    (t2, c) = uniform(n * n, np.float32, 3)
    for es, dt in ((1, np.uint8), (2, np.int16)):
        cnt = n * n * 4 // es
        side = int(cnt ** 0.5) // 64 * 64
        va = CudaVec(side * side, dt, _ptr=t.data_ptr(), _owner=t)
        vc = CudaVec(side * side, dt, _ptr=t2.data_ptr(), _owner=t2)
        report(f"transpose {side}x{side} {np.dtype(dt).name} single", timed(lambda: cuda.transpose(va, vc, side)), nbytes=2.0 * side * side * es)
        b = 1 << 20
        vb, vd = CudaVec(b * 256, dt, _ptr=t.data_ptr(), _owner=t), CudaVec(b * 256, dt, _ptr=t2.data_ptr(), _owner=t2)
        report(f"transpose 16x16 {np.dtype(dt).name} batch 2^20", timed(lambda: cuda.transpose_batched(vb, vd, 256, 16)), nbytes=2.0 * b * 256 * es)


def g_fused():
    """fused chains against the op sequences they replace (bytes = what the FUSED form has to move)"""
    n = 1 << 28
    (_, a), (_, b), (_, c), (_, d) = uniform(n, np.float32, 1), uniform(n, np.float32, 2), uniform(n, np.float32, 3), uniform(n, np.float32, 4)
    report("add then dot (unfused)", timed(lambda: (cuda.add(a, b, c), cuda.dot(c, d))), nbytes=16.0 * n)
    report("add_dot fused, c written", timed(lambda: cuda.ewise_dot("add", a, b, d, c)), nbytes=16.0 * n)
    report("add_dot fused, c not materialised", timed(lambda: cuda.ewise_dot("add", a, b, d)), nbytes=12.0 * n)
    report("dot + norm + norm (unfused)", timed(lambda: (cuda.dot(a, b), cuda.norm(a), cuda.norm(b))), nbytes=8.0 * n)
    report("dot_norms fused", timed(lambda: cuda.dot_norms(a, b)), nbytes=8.0 * n)
    report("copy then normalize (unfused)", timed(lambda: (_lib.call("slas_cuda_memcpy_d2d", c.as_ptr(), a.as_ptr(), 4 * n), cuda.normalize(c))), nbytes=12.0 * n)
    report("normalize_into fused", timed(lambda: cuda.normalize_into(a, c)), nbytes=12.0 * n)


def g_gemm(knob, size, dt, variants, log2=20):
    b = 1 << log2
    e = size * size
    (_, a), (_, bb), (_, c) = uniform(b * e, dt, 1), uniform(b * e, dt, 2), uniform(b * e, dt, 3)
    for v in variants:
        set_tunable(knob, v)
        report(f"gemm {size}^3 {np.dtype(dt).name} {knob}={v}", timed(lambda: cuda.matrix_mul_batched(a, bb, c, size, size, size)),
               nbytes=3.0 * b * e * np.dtype(dt).itemsize, flop=2.0 * size ** 3 * b)
    set_tunable(knob, 0)


def g_large(path, knob, variants):
    n = 4096
    (_, a), (_, b), (_, c) = uniform(n * n, np.float32, 5), uniform(n * n, np.float32, 6), uniform(n * n, np.float32, 7)
    sl.set_sgemm_path(path)
    for v in variants:
        set_tunable(knob, v)
        report(f"sgemm 4096^3 {path} {knob}={v}", timed(lambda: cuda.matrix_mul(a, b, c, n, n, n, n, n, n, False, False), 10, 3),
               flop=2.0 * n ** 3)
    set_tunable(knob, 0)
    sl.set_sgemm_path("auto")


def g_crossover():
    """FFMA vs tcgen05 3xTF32 across sizes: where the automatic choice should switch"""
    for n in (512, 768, 1024, 1280, 1536, 2048, 3072, 4096):
        (_, a), (_, b), (_, c) = uniform(n * n, np.float32, 5), uniform(n * n, np.float32, 6), uniform(n * n, np.float32, 7)
        for path in ("ffma", "tcgen05"):
            sl.set_sgemm_path(path)
            report(f"sgemm {n}^3 {path}", timed(lambda: cuda.matrix_mul(a, b, c, n, n, n, n, n, n, False, False), 10, 3), flop=2.0 * n ** 3)
    sl.set_sgemm_path("auto")


def g_dgemm():
    """large f64 product, all four transposition flags (FP64 pipe bound: 26.6 TFLOP/s measured DFMA peak)"""
    for n in (1024, 2048, 4096):
        (_, a), (_, b), (_, c) = uniform(n * n, np.float64, 5), uniform(n * n, np.float64, 6), uniform(n * n, np.float64, 7)
        for ta in (False, True):
            for tb in (False, True):
                report(f"dgemm {n}^3 ta={int(ta)} tb={int(tb)}", timed(lambda: cuda.matrix_mul(a, b, c, n, n, n, n, n, n, ta, tb), 5, 2),
                       flop=2.0 * n ** 3)


def g_e2e():
    """host-pointer pipeline (pinned operands through the C ABI): slot size sweep on the headline workload"""
    import ctypes as C
    import time
    b = 1 << 24
    nbytes = b * 64
    hp = [C.c_void_p() for _ in range(3)]
    for p in hp:
        _lib.call("slas_cuda_malloc_host", C.byref(p), nbytes)
    hA, hB, hC = [np.ctypeslib.as_array((C.c_float * (b * 16)).from_address(p.value)) for p in hp]
    hA[:] = 0.5
    hB[:] = 0.25
    for mb in (0, 16, 32, 48, 64, 128, 192):
        set_tunable("host_slot_mb", mb)
        cuda.matrix_mul_batched(hA, hB, hC, 4, 4, 4)
        t0 = time.perf_counter()
        for _ in range(4):
            cuda.matrix_mul_batched(hA, hB, hC, 4, 4, 4)
        ms = (time.perf_counter() - t0) * 1e3 / 4
        report(f"e2e 4x4 f32 2^24 pinned host operands, slot {mb or 96} MB", ms, nbytes=3.0 * nbytes, flop=128.0 * b)
    set_tunable("host_slot_mb", 0)
    for p in hp:
        _lib.call("slas_cuda_free_host", p.value)


def g_pageable():
    """host-pointer pipeline with PAGEABLE operands (what a slas StaticVec is): copy-thread count (environment, read once
    per process: one subprocess per setting), ring slot size, streaming stores on / off; plus the bare copy pool rate"""
    import subprocess
    import time
    if os.environ.get("SLAS_TUNE_PAGEABLE_CHILD"):
        b = 1 << 24
        A, B = np.full(b * 16, 0.5, np.float32), np.full(b * 16, 0.25, np.float32)
        Cm = np.zeros(b * 16, np.float32)
        for nt in (0, 1):
            set_tunable("host_copy_nt", nt)
            for mb in (8, 24, 64):
                set_tunable("host_slot_mb", mb)
                cuda.matrix_mul_batched(A, B, Cm, 4, 4, 4)
                t0 = time.perf_counter()
                for _ in range(3):
                    cuda.matrix_mul_batched(A, B, Cm, 4, 4, 4)
                ms = (time.perf_counter() - t0) * 1e3 / 3
                report(f"e2e 4x4 f32 2^24 PAGEABLE host operands, {os.environ.get('SLAS_CUDA_HOST_COPY_THREADS')} copy threads, slot {mb} MB, "
                       f"{'cached' if nt else 'streaming'} stores", ms, nbytes=3.0 * b * 64, flop=128.0 * b)
        set_tunable("host_slot_mb", 0)
        set_tunable("host_copy_nt", 0)
        # a plain 1 GiB numpy copy on one thread, for scale
        t0 = time.perf_counter()
        Cm[:] = A
        report("numpy 1 GiB copy, one thread", (time.perf_counter() - t0) * 1e3, nbytes=2.0 * b * 64)
        return
    for threads in (4, 8, 12, 16):
        env = dict(os.environ, SLAS_TUNE_PAGEABLE_CHILD="1", SLAS_CUDA_HOST_COPY_THREADS=str(threads))
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "pageable"], env=env, capture_output=True, text=True)
        sys.stdout.write(r.stdout)
        if r.returncode:
            sys.stdout.write(r.stderr[-2000:])
        sys.stdout.flush()


def g_gemv():
    for size, dt, log2 in ((4, np.float32, 24), (16, np.float32, 20), (32, np.float32, 20), (16, np.float64, 20), (32, np.float64, 20),
                           (64, np.float32, 18)):
        b = 1 << log2
        es = np.dtype(dt).itemsize
        (_, a), (_, x), (_, y) = uniform(b * size * size, dt, 1), uniform(b * size, dt, 2), uniform(b * size, dt, 3)
        for ta in (False, True):
            for v in ((0, 3) if not ta else (0,)):   # 0 shipped (lanes share a row); 3: lane per row
                set_tunable("gemv_variant", v)
                report(f"gemv {size}x{size} {np.dtype(dt).name} batch 2^{log2} trans={ta} variant {v}",
                       timed(lambda: cuda.matrix_vector_mul_batched(a, x, y, size, size, ta)), nbytes=float(b) * (size * size + 2 * size) * es)
        set_tunable("gemv_variant", 0)


def g_bdot():
    for length, log2, dt in ((1 << 20, 8, np.float32), (1 << 16, 12, np.float32), (1 << 15, 13, np.float32), (1 << 14, 14, np.float32),
                             (4096, 16, np.float32), (2048, 17, np.float32), (2048, 16, np.float64), (1 << 14, 13, np.float64),
                             (256, 20, np.float32), (64, 22, np.float32), (16, 24, np.float32), (8, 24, np.float32), (16, 23, np.float64),
                             (1 << 20, 7, np.float64)):
        b = 1 << log2
        es = np.dtype(dt).itemsize
        (_, a), (_, bb), (_, o) = uniform(b * length, dt, 1), uniform(b * length, dt, 2), uniform(b, dt, 3)
        report(f"dot_batched len {length} x 2^{log2} {np.dtype(dt).name}", timed(lambda: cuda.dot_batched(a, bb, o, length)),
               nbytes=2.0 * b * length * es)
        report(f"norm_batched len {length} x 2^{log2} {np.dtype(dt).name}", timed(lambda: cuda.norm_batched(a, o, length)),
               nbytes=1.0 * b * length * es)
        report(f"normalize_batched len {length} x 2^{log2} {np.dtype(dt).name}", timed(lambda: cuda.normalize_batched(a, length)),
               nbytes=2.0 * b * length * es)


def g_rect():
    for (m, n, k, dt, log2) in ((2, 2, 3, np.float32, 24), (3, 3, 3, np.float32, 24), (3, 3, 3, np.float64, 24), (2, 2, 2, np.float32, 24),
                                (4, 4, 2, np.float32, 24), (4, 1, 4, np.float32, 24), (2, 4, 4, np.float64, 24),
                                (5, 5, 5, np.float32, 23), (6, 6, 6, np.float32, 22), (6, 6, 6, np.float64, 22), (7, 7, 7, np.float32, 22),
                                (9, 9, 9, np.float32, 21), (10, 10, 10, np.float32, 21), (15, 15, 15, np.float64, 19), (5, 9, 13, np.float32, 21),
                                (8, 4, 16, np.float32, 22), (16, 32, 8, np.float32, 20),
                                (12, 12, 12, np.float64, 20), (4, 4, 4, np.float32, 24)):
        b = 1 << log2
        es = np.dtype(dt).itemsize
        (_, a), (_, bb), (_, c) = uniform(b * m * k, dt, 1), uniform(b * k * n, dt, 2), uniform(b * m * n, dt, 3)
        report(f"gemm {m}x{k} . {k}x{n} {np.dtype(dt).name} batch 2^{log2}", timed(lambda: cuda.matrix_mul_batched(a, bb, c, m, n, k)),
               nbytes=float(b) * (m * k + k * n + m * n) * es, flop=2.0 * m * n * k * b)


def g_misc():
    # one large gemv / transpose (single objects), and launch-bound small calls replayed from a CUDA graph
    n = 8192
    (_, a), (_, x), (_, y) = uniform(n * n, np.float32, 1), uniform(n, np.float32, 2), uniform(n, np.float32, 3)
    for ta in (False, True):
        report(f"gemv {n}x{n} float32 single trans={ta}", timed(lambda: cuda.matrix_vector_mul(a, x, y, n, n, n, ta)), nbytes=4.0 * n * n)
    (_, c) = uniform(n * n, np.float32, 4)
    report(f"transpose {n}x{n} float32 single", timed(lambda: cuda.transpose(a, c, n)), nbytes=8.0 * n * n)
    (_, a8), (_, c8) = uniform(n * n // 2, np.float64, 1), uniform(n * n // 2, np.float64, 4)
    report(f"transpose {n}x{n // 2} float64 single", timed(lambda: cuda.transpose(a8, c8, n)), nbytes=8.0 * n * n)
    l20 = 1 << 20
    ta_, va = uniform(256 * l20, np.float32, 5)
    tb_, vb = uniform(256 * l20, np.float32, 6)
    out = torch.zeros(256, device=dev, dtype=torch.float32)
    L = _lib.lib()

    def loop():
        for i in range(256):
            L.slas_cuda_sdot(l20, ta_.data_ptr() + 4 * i * l20, tb_.data_ptr() + 4 * i * l20, out.data_ptr() + 4 * i)

    report("dot 2^20 x 256 calls, eager", timed(loop, 5, 2), nbytes=8.0 * 256 * l20)
    g = torch.cuda.CUDAGraph()
    loop()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=stream):
        loop()
    report("dot 2^20 x 256 calls, CUDA graph replay", timed(g.replay, 5, 2), nbytes=8.0 * 256 * l20)
    with sl.Graph() as own:  # the library's own capture API (slas_cuda_graph_begin / _end / _launch)
        loop()
    report("dot 2^20 x 256 calls, slas_cuda_graph replay", timed(own.launch, 5, 2), nbytes=8.0 * 256 * l20)


def g_small():
    """BASELINE configs[0]: single dot / add calls at LEN 2^16 .. 2^22, 256 distinct pairs, replayed from one CUDA graph
    (per-call time without the interpreter), over the small-vector geometry knobs: CTA size, CTAs per SM, tail protocol
    (flagged words polled by block 0 vs ticket + last block), against the round-1 tiled geometry (reduce_small = 1)."""
    import ctypes as C
    ncalls = 256
    for log2 in (20, 16, 22):
        n = 1 << log2
        per = max(n, 1 << 20)  # distinct operands, >= 2 GiB in total for the 2^20 case
        (ta, _), (tb, _), (tc, _) = uniform(ncalls * per, np.float32, 1), uniform(ncalls * per, np.float32, 2), uniform(ncalls * per, np.float32, 3)
        out = torch.zeros(ncalls, device=dev, dtype=torch.float32)

        def loop(op):
            _lib.call("slas_cuda_bench_call_loop", op, n, ta.data_ptr(), tb.data_ptr(), (out if op == 0 else tc).data_ptr(), ncalls, per)

        def graph_time(op):
            loop(op)
            with sl.Graph() as g:
                loop(op)
            ms = timed(g.launch, 10, 3)
            del g
            return ms / ncalls

        # "pdl": 1 = plain launches, 0 = dependent launch with the wait at the kernel top (a stream shared with other
        # libraries), 2 = the caller vouches for the stream: loads run ahead of the predecessor's tail (what the library's
        # own stream gets by default)
        variants = [("round-1 tiled geometry + ticket, plain launch", {"reduce_small": 1, "pdl": 1}),
                    ("balanced 512 x 2/SM, flagged words, plain launch", {"reduce_small_threads": 512, "pdl": 1}),
                    ("balanced 512 x 2/SM, flagged words, PDL wait-at-top", {"reduce_small_threads": 512}),
                    ("balanced 512 x 2/SM, flagged words, PDL early loads", {"reduce_small_threads": 512, "pdl": 2}),
                    ("balanced 512 x 1/SM, flagged words, PDL early loads", {"reduce_small_threads": 512, "reduce_small_cps": 1, "pdl": 2}),
                    ("balanced 1024 x 1/SM, flagged words, plain launch", {"reduce_small_threads": 1024, "pdl": 1}),
                    ("balanced 1024 x 1/SM, flagged words, PDL early loads", {"reduce_small_threads": 1024, "pdl": 2}),
                    ("balanced 1024 x 2/SM, flagged words, PDL early loads", {"reduce_small_threads": 1024, "reduce_small_cps": 2, "pdl": 2}),
                    ("default geometry, PDL early loads", {"pdl": 2}),
                    ("default geometry, plain launch", {"pdl": 1}),
                    ("balanced 256 x 4/SM, flagged words, PDL early loads", {"reduce_small_threads": 256, "pdl": 2}),
                    ("balanced 256 x 2/SM, flagged words, PDL early loads", {"reduce_small_threads": 256, "reduce_small_cps": 2, "pdl": 2}),
                    ("balanced 512 x 2/SM, ticket, PDL early loads", {"reduce_small_threads": 512, "reduce_small_tail": 1, "pdl": 2})]
        for name, knobs in variants:
            for k, v in knobs.items():
                set_tunable(k, v)
            ms = graph_time(0)
            ms_c = timed(lambda: loop(0), 5, 2) / ncalls
            # phases of one cold kernel
            set_tunable("reduce_phases", 1)
            ph = (C.c_uint64 * 5)()
            st = []
            for i in range(8):
                _lib.lib().slas_cuda_sdot(n, ta.data_ptr() + 4 * (i + 64) * per, tb.data_ptr() + 4 * (i + 64) * per, out.data_ptr())
                _lib.call("slas_cuda_get_reduce_phases", ph)
                st.append([int(ph[j]) - int(ph[0]) for j in range(5)])
            set_tunable("reduce_phases", 0)
            med = [int(np.median([x[j] for x in st])) for j in range(1, 5)]
            for k in knobs:
                set_tunable(k, 0)
            report(f"dot 2^{log2} f32 single call [{name}] graph {ms * 1e3:.2f} us/call, C loop {ms_c * 1e3:.2f} us/call, "
                   f"phases ns (loads, block sum, partials in, written) {med}", ms, nbytes=8.0 * n)
        for name, knobs in (("round-1 tiled 256 x 1024 vec", {"ewise_small": 1}), ("balanced 256 x 8/SM, plain launch", {"pdl": 1}),
                            ("balanced 256 x 8/SM, PDL wait-at-top", {}), ("balanced 256 x 8/SM, PDL early", {"pdl": 2}),
                            ("balanced 512 x 4/SM, PDL early", {"ewise_small": 2, "pdl": 2}), ("balanced 1024 x 2/SM, PDL early", {"ewise_small": 3, "pdl": 2})):
            for k, v in knobs.items():
                set_tunable(k, v)
            ms = graph_time(1)
            ms_c = timed(lambda: loop(1), 5, 2) / ncalls
            for k in knobs:
                set_tunable(k, 0)
            report(f"add 2^{log2} f32 single call [{name}] graph {ms * 1e3:.2f} us/call, C loop {ms_c * 1e3:.2f} us/call", ms, nbytes=12.0 * n)
        del ta, tb, tc, out
        torch.cuda.empty_cache()


def g_vec3():
    """3-vector dot / norm / normalize (bulk-TMA staged, a thread per vector): stage size (CTAs per SM)"""
    b = 1 << 24
    (_, x), (_, y) = uniform(b * 3, np.float32, 1), uniform(b * 3, np.float32, 2)
    (_, o) = uniform(b, np.float32, 3)
    for kb in (0, 4, 8, 12, 24, 32):
        set_tunable("bdot_stage_kb", kb)
        report(f"batched 3-vector f32 dot, 2^24, stage {kb or 16} KB", timed(lambda: cuda.dot_batched(x, y, o, 3)), nbytes=28.0 * b)
        report(f"batched 3-vector f32 norm, 2^24, stage {kb or 16} KB", timed(lambda: cuda.norm_batched(x, o, 3)), nbytes=16.0 * b)
        report(f"batched 3-vector f32 normalize, 2^24, stage {kb or 16} KB", timed(lambda: cuda.normalize_batched(x, 3)), nbytes=24.0 * b)
    set_tunable("bdot_stage_kb", 0)


def g_odd():
    """small objects that are not whole 16-byte words (bulk-TMA staged kernels): shape as a run-time argument (staged_jit 2)
    against shape as a compile-time constant (0: built in for LEN <= 8 / 3x3, NVRTC for the rest); same bits required"""
    for length, log2, dt in ((3, 24, np.float32), (3, 24, np.float64), (2, 24, np.float32), (5, 23, np.float32), (7, 23, np.float32),
                             (7, 22, np.float64), (9, 22, np.float32), (13, 22, np.float32), (30, 21, np.float32), (19, 21, np.float64),
                             (38, 21, np.float32), (127, 19, np.float32)):
        b, es = 1 << log2, np.dtype(dt).itemsize
        (ta, a), (_, bb), (to, o) = uniform(b * length, dt, 1), uniform(b * length, dt, 2), uniform(b, dt, 3)
        keep = ta.clone()
        for name, fn, nbytes in (("dot", lambda: cuda.dot_batched(a, bb, o, length), (2.0 * length + 1) * b * es),
                                 ("norm", lambda: cuda.norm_batched(a, o, length), (length + 1.0) * b * es),
                                 ("normalize", lambda: cuda.normalize_batched(a, length), 2.0 * length * b * es)):
            res = []
            for mode in (2, 0):
                set_tunable("staged_jit", mode)
                ta.copy_(keep)
                fn()
                res.append((ta if name == "normalize" else to).clone())
                ta.copy_(keep)
                report(f"batched LEN {length} {np.dtype(dt).name} {name}, 2^{log2}, shape {'run-time' if mode else 'compile-time'}",
                       timed(fn), nbytes=nbytes)
            assert torch.equal(res[0].view(torch.uint8), res[1].view(torch.uint8)), (name, length)
            ta.copy_(keep)
        del ta, a, bb, to, o, keep
        torch.cuda.empty_cache()
    for rows, cols, log2, dt in ((3, 3, 24, np.float32), (3, 3, 23, np.float64), (2, 3, 24, np.float32), (5, 5, 22, np.float32),
                                 (3, 5, 23, np.float32), (6, 6, 22, np.float32), (7, 7, 21, np.float32), (9, 9, 21, np.float32),
                                 (5, 3, 22, np.float64), (10, 10, 20, np.float32), (30, 30, 18, np.float32), (100, 100, 14, np.float32)):
        b, es, e = 1 << log2, np.dtype(dt).itemsize, rows * cols
        (ta, a), (tc, c) = uniform(b * e, dt, 1), uniform(b * e, dt, 2)
        res = []
        for mode in (2, 0):
            set_tunable("staged_jit", mode)
            cuda.transpose_batched(a, c, e, cols)
            res.append(tc.clone())
            report(f"transpose {rows}x{cols} {np.dtype(dt).name}, 2^{log2}, shape {'run-time' if mode else 'compile-time'}",
                   timed(lambda: cuda.transpose_batched(a, c, e, cols)), nbytes=2.0 * b * e * es)
        assert torch.equal(res[0].view(torch.uint8), res[1].view(torch.uint8)), (rows, cols)
        set_tunable("staged_jit", 0)
        keep = ta.clone()
        cuda.transpose_inplace_batched(a, e, cols)
        assert torch.equal(ta.view(torch.uint8), res[0].view(torch.uint8)), ("in place", rows, cols)
        ta.copy_(keep)
        report(f"transpose_inplace {rows}x{cols} {np.dtype(dt).name}, 2^{log2}, compile-time",
               timed(lambda: cuda.transpose_inplace_batched(a, e, cols)), nbytes=2.0 * b * e * es)
        del ta, a, tc, c, keep
        torch.cuda.empty_cache()
    set_tunable("staged_jit", 0)


def g_oddsweep():
    """stage size of the staged small-object kernels (CTAs per SM against bytes per bulk copy)"""
    b = 1 << 24
    (_, x), (_, y), (_, o) = uniform(b * 3, np.float32, 1), uniform(b * 3, np.float32, 2), uniform(b, np.float32, 3)
    set_tunable("bdot_gcap", 8192)
    for kb in (4, 8, 12, 16, 24, 32, 48):
        set_tunable("bdot_stage_kb", kb)
        report(f"batched 3-vector f32 dot, 2^24, stage {kb} KB", timed(lambda: cuda.dot_batched(x, y, o, 3)), nbytes=28.0 * b)
        report(f"batched 3-vector f32 norm, 2^24, stage {kb} KB", timed(lambda: cuda.norm_batched(x, o, 3)), nbytes=16.0 * b)
        report(f"batched 3-vector f32 normalize, 2^24, stage {kb} KB", timed(lambda: cuda.normalize_batched(x, 3)), nbytes=24.0 * b)
    del x, y, o
    b = 1 << 21
    (_, x), (_, y), (_, o) = uniform(b * 30, np.float32, 1), uniform(b * 30, np.float32, 2), uniform(b, np.float32, 3)
    for kb in (4, 8, 12, 16, 24, 32, 48):
        set_tunable("bdot_stage_kb", kb)
        report(f"batched 30-vector f32 dot, 2^21, stage {kb} KB", timed(lambda: cuda.dot_batched(x, y, o, 30)), nbytes=244.0 * b)
        report(f"batched 30-vector f32 normalize, 2^21, stage {kb} KB", timed(lambda: cuda.normalize_batched(x, 30)), nbytes=240.0 * b)
    set_tunable("bdot_stage_kb", 0)
    set_tunable("bdot_gcap", 0)
    del x, y, o
    torch.cuda.empty_cache()
    for rows, cols, log2, dt in ((3, 3, 24, np.float32), (3, 3, 23, np.float64), (9, 9, 21, np.float32), (30, 30, 18, np.float32)):
        b, es, e = 1 << log2, np.dtype(dt).itemsize, rows * cols
        (ta, a), (tc, c) = uniform(b * e, dt, 1), uniform(b * e, dt, 2)
        for mode in (2, 0):
            set_tunable("staged_jit", mode)
            for kb in (4, 8, 12, 16, 24, 32, 40):
                set_tunable("transpose_stage_kb", kb)
                report(f"transpose {rows}x{cols} {np.dtype(dt).name}, 2^{log2}, shape {'run-time' if mode else 'compile-time'}, stage {kb} KB",
                       timed(lambda: cuda.transpose_batched(a, c, e, cols)), nbytes=2.0 * b * e * es)
        del ta, a, tc, c
        torch.cuda.empty_cache()
    set_tunable("staged_jit", 0)
    set_tunable("transpose_stage_kb", 0)


def g_oddsweep2():
    """small stages x CTAs per SM for the compile-time-shape staged kernels"""
    cases = []
    b = 1 << 24
    (_, x), (_, y), (_, o) = uniform(b * 3, np.float32, 1), uniform(b * 3, np.float32, 2), uniform(b, np.float32, 3)
    (_, m), (_, c) = uniform(b * 9, np.float32, 4), uniform(b * 9, np.float32, 5)
    b7 = 1 << 23
    (_, x7) = uniform(b7 * 7, np.float32, 6)
    (_, md), (_, cd) = uniform(b7 * 9, np.float64, 4), uniform(b7 * 9, np.float64, 5)
    set_tunable("bdot_gcap", 8192)
    for cap in (4, 5, 6, 7):
        set_tunable("staged_cta_cap", cap)
        for kb in (2, 3, 4, 6, 8, 12, 16):
            set_tunable("bdot_stage_kb", kb)
            set_tunable("transpose_stage_kb", kb)
            tag = f"cap {cap}, stage {kb} KB"
            report(f"3-vector f32 dot, {tag}", timed(lambda: cuda.dot_batched(x, y, o, 3)), nbytes=28.0 * b)
            report(f"3-vector f32 norm, {tag}", timed(lambda: cuda.norm_batched(x, o, 3)), nbytes=16.0 * b)
            report(f"3-vector f32 normalize, {tag}", timed(lambda: cuda.normalize_batched(x, 3)), nbytes=24.0 * b)
            report(f"7-vector f32 normalize, {tag}", timed(lambda: cuda.normalize_batched(x7, 7)), nbytes=56.0 * b7)
            report(f"transpose 3x3 f32, {tag}", timed(lambda: cuda.transpose_batched(m, c, 9, 3)), nbytes=72.0 * b)
            report(f"transpose 3x3 f64, {tag}", timed(lambda: cuda.transpose_batched(md, cd, 9, 3)), nbytes=144.0 * b7)
    for k in ("bdot_gcap", "staged_cta_cap", "bdot_stage_kb", "transpose_stage_kb"):
        set_tunable(k, 0)


def g_oddsweep3():
    """stage size for the staged normalize (two more shared tiles than dot: CTAs per SM against bytes per bulk copy)"""
    for length, log2, dt in ((3, 24, np.float32), (7, 23, np.float32), (9, 22, np.float32), (13, 22, np.float32), (30, 21, np.float32),
                             (38, 21, np.float32), (127, 19, np.float32), (3, 23, np.float64), (7, 22, np.float64), (19, 21, np.float64)):
        b, es = 1 << log2, np.dtype(dt).itemsize
        (_, x) = uniform(b * length, dt, 1)
        for kb in (6, 8, 12, 16, 24):
            set_tunable("bdot_stage_kb", kb)
            report(f"batched LEN {length} {np.dtype(dt).name} normalize, 2^{log2}, stage {kb} KB",
                   timed(lambda: cuda.normalize_batched(x, length)), nbytes=2.0 * length * b * es)
        del x
        torch.cuda.empty_cache()
    set_tunable("bdot_stage_kb", 0)


def g_bnorm():
    """batched norm / normalize of medium vectors (8 KB .. 512 KB): CTA per vector (bnorm_cluster = 1) against a thread-block
    cluster per vector with the vector in registers (0), CTAs per SM swept"""
    for length, log2, dt in ((2048, 17, np.float32), (4096, 16, np.float32), (8192, 15, np.float32), (16384, 14, np.float32),
                             (32768, 13, np.float32), (65536, 12, np.float32), (131072, 11, np.float32), (20000, 14, np.float32),
                             (2048, 16, np.float64), (16384, 13, np.float64), (65536, 11, np.float64)):
        b, es = 1 << log2, np.dtype(dt).itemsize
        (_, a), (_, o) = uniform(b * length, dt, 1), uniform(b, dt, 3)
        for off, caps in ((1, (0,)), (0, (0, 3, 6))):
            set_tunable("bnorm_cluster", off)
            for cap in caps:
                set_tunable("bnorm_cluster_ctas", cap)
                tag = "CTA per vector" if off else f"cluster per vector, {cap or 'default'} CTAs per SM"
                report(f"norm_batched len {length} x 2^{log2} {np.dtype(dt).name} [{tag}]", timed(lambda: cuda.norm_batched(a, o, length)),
                       nbytes=1.0 * b * length * es)
                report(f"normalize_batched len {length} x 2^{log2} {np.dtype(dt).name} [{tag}]", timed(lambda: cuda.normalize_batched(a, length)),
                       nbytes=2.0 * b * length * es)
        del a, o
        torch.cuda.empty_cache()
    set_tunable("bnorm_cluster", 0)
    set_tunable("bnorm_cluster_ctas", 0)


def g_trsmall():
    """small transposes, a thread per object: stage size sweep over shapes"""
    for rows, cols, log2, dt in ((2, 3, 24, np.float32), (3, 3, 24, np.float32), (3, 5, 23, np.float32), (5, 5, 22, np.float32), (7, 7, 21, np.float32),
                                 (3, 7, 22, np.float32), (2, 3, 23, np.float64), (3, 3, 23, np.float64), (5, 3, 22, np.float64), (5, 5, 21, np.float64)):
        b, es, e = 1 << log2, np.dtype(dt).itemsize, rows * cols
        (_, a), (_, c) = uniform(b * e, dt, 1), uniform(b * e, dt, 2)
        for mode in (2, 0):   # 2: the run-time walk, 0: a thread per object where the shape qualifies
            set_tunable("staged_jit", mode)
            for kb in (0, 2, 3, 4, 5, 6, 8, 12, 16):
                set_tunable("transpose_stage_kb", kb)
                report(f"transpose {rows}x{cols} {np.dtype(dt).name} 2^{log2}, {'run-time walk' if mode else 'thread per object'}, stage {kb or 'default'} KB",
                       timed(lambda: cuda.transpose_batched(a, c, e, cols)), nbytes=2.0 * b * e * es)
        del a, c
        torch.cuda.empty_cache()
    set_tunable("staged_jit", 0)
    set_tunable("transpose_stage_kb", 0)


def g_oddsweep4():
    """stage size for the staged dot / norm over vector lengths"""
    for length, log2, dt in ((2, 24, np.float32), (3, 24, np.float32), (5, 23, np.float32), (7, 23, np.float32), (9, 22, np.float32),
                             (13, 22, np.float32), (30, 21, np.float32), (38, 21, np.float32), (127, 19, np.float32),
                             (3, 23, np.float64), (7, 22, np.float64), (19, 21, np.float64)):
        b, es = 1 << log2, np.dtype(dt).itemsize
        (_, x), (_, y), (_, o) = uniform(b * length, dt, 1), uniform(b * length, dt, 2), uniform(b, dt, 3)
        for kb in (4, 6, 8, 12, 16, 24):
            set_tunable("bdot_stage_kb", kb)
            report(f"batched LEN {length} {np.dtype(dt).name} dot, 2^{log2}, stage {kb} KB",
                   timed(lambda: cuda.dot_batched(x, y, o, length)), nbytes=(2.0 * length + 1) * b * es)
            report(f"batched LEN {length} {np.dtype(dt).name} norm, 2^{log2}, stage {kb} KB",
                   timed(lambda: cuda.norm_batched(x, o, length)), nbytes=(length + 1.0) * b * es)
        del x, y, o
        torch.cuda.empty_cache()
    set_tunable("bdot_stage_kb", 0)


def g_trmid():
    """batched transposes across object sizes and aspect ratios, out of place and in place (which kernel takes what)"""
    for rows, cols, dt in ((12, 12, np.float32), (20, 20, np.float32), (24, 8, np.float32), (8, 24, np.float32), (33, 31, np.float32),
                           (31, 33, np.float32), (64, 40, np.float32), (100, 3, np.float32), (3, 100, np.float32), (50, 50, np.float32),
                           (96, 96, np.float32), (128, 128, np.float32), (200, 200, np.float32), (12, 12, np.float64), (48, 5, np.float64),
                           (33, 31, np.float64), (100, 100, np.float64)):
        e, es = rows * cols, np.dtype(dt).itemsize
        b = max(256, (1 << 28) // (e * es))
        (_, a), (_, c) = uniform(b * e, dt, 1), uniform(b * e, dt, 2)
        report(f"transpose {rows}x{cols} {np.dtype(dt).name} batch {b}", timed(lambda: cuda.transpose_batched(a, c, e, cols)), nbytes=2.0 * b * e * es)
        report(f"transpose_inplace {rows}x{cols} {np.dtype(dt).name} batch {b}", timed(lambda: cuda.transpose_inplace_batched(a, e, cols)),
               nbytes=2.0 * b * e * es)
        del a, c
        torch.cuda.empty_cache()


def g_lanevsstaged():
    """vectors of whole 16-byte words up to 256 bytes: a lane per vector with 256-bit accesses (shipped) against the bulk-TMA
    staged thread-per-vector kernel (bdot_variant 4)"""
    for length, log2, dt in ((4, 24, np.float32), (8, 24, np.float32), (16, 24, np.float32), (32, 23, np.float32), (64, 22, np.float32),
                             (4, 24, np.float64), (8, 23, np.float64), (16, 23, np.float64), (32, 22, np.float64)):
        b, es = 1 << log2, np.dtype(dt).itemsize
        (_, x), (_, y), (_, o) = uniform(b * length, dt, 1), uniform(b * length, dt, 2), uniform(b, dt, 3)
        for v in (0, 4):
            set_tunable("bdot_variant", v)
            tag = "lane per vector" if v == 0 else "staged"
            report(f"batched LEN {length} {np.dtype(dt).name} dot, 2^{log2}, {tag}", timed(lambda: cuda.dot_batched(x, y, o, length)), nbytes=(2.0 * length + 1) * b * es)
            report(f"batched LEN {length} {np.dtype(dt).name} norm, 2^{log2}, {tag}", timed(lambda: cuda.norm_batched(x, o, length)), nbytes=(length + 1.0) * b * es)
            report(f"batched LEN {length} {np.dtype(dt).name} normalize, 2^{log2}, {tag}", timed(lambda: cuda.normalize_batched(x, length)), nbytes=2.0 * length * b * es)
        del x, y, o
        torch.cuda.empty_cache()
    set_tunable("bdot_variant", 0)


def g_tinyvsjit():
    """every dimension in 1..4: a thread per matrix (gemm_tiny, shipped) against whatever the run-time specialisation picks
    (gemm_tiny_variant 1: register tiles where N and K are whole 16-byte vectors, else scalar register tiles)"""
    for (m, n, k, dt, log2) in ((2, 2, 2, np.float32, 24), (4, 4, 2, np.float32, 24), (2, 4, 4, np.float32, 24), (4, 1, 4, np.float32, 24),
                                (4, 4, 4, np.float64, 23), (2, 4, 4, np.float64, 24), (4, 2, 4, np.float64, 24), (2, 2, 2, np.float64, 24),
                                (3, 3, 3, np.float32, 24), (4, 4, 3, np.float32, 24)):
        b, es = 1 << log2, np.dtype(dt).itemsize
        (_, a), (_, bb), (_, c) = uniform(b * m * k, dt, 1), uniform(b * k * n, dt, 2), uniform(b * m * n, dt, 3)
        for v in (0, 1):
            set_tunable("gemm_tiny_variant", v)
            report(f"gemm {m}x{k} . {k}x{n} {np.dtype(dt).name} batch 2^{log2} [{'thread per matrix' if v == 0 else 'run-time specialised'}]",
                   timed(lambda: cuda.matrix_mul_batched(a, bb, c, m, n, k)), nbytes=float(b) * (m * k + k * n + m * n) * es)
        del a, bb, c
        torch.cuda.empty_cache()
    set_tunable("gemm_tiny_variant", 0)


def g_gemm4d():
    """4x4 f64 batched products (the f64 sibling of the headline): stage depth / matrices per stage / bulk C store"""
    b = 1 << 23
    (_, a), (_, bb), (_, c) = uniform(b * 16, np.float64, 1), uniform(b * 16, np.float64, 2), uniform(b * 16, np.float64, 3)
    for v in (0, 1, 2, 3, 4):
        set_tunable("gemm4d_variant", v)
        report(f"gemm 4x4 . 4x4 float64 batch 2^23 variant {v}", timed(lambda: cuda.matrix_mul_batched(a, bb, c, 4, 4, 4)), nbytes=float(b) * 48 * 8)
    set_tunable("gemm4d_variant", 0)


def g_jitbulkc():
    """run-time specialised register-tile products: per-warp C stores against a bulk C store from a shared tile"""
    for (m, n, k, dt, log2) in ((12, 12, 12, np.float64, 20), (8, 4, 16, np.float32, 22), (16, 32, 8, np.float32, 20), (6, 6, 6, np.float64, 22),
                                (20, 20, 20, np.float32, 18), (24, 24, 24, np.float32, 18), (8, 8, 4, np.float32, 22), (2, 4, 4, np.float64, 24),
                                (12, 12, 12, np.float32, 20), (20, 20, 20, np.float64, 17)):
        b, es = 1 << log2, np.dtype(dt).itemsize
        (_, a), (_, bb), (_, c) = uniform(b * m * k, dt, 1), uniform(b * k * n, dt, 2), uniform(b * m * n, dt, 3)
        for v in (2, 1, 0):
            set_tunable("gemm_jit_bulk_c", v)
            report(f"gemm {m}x{k} . {k}x{n} {np.dtype(dt).name} batch 2^{log2} [{('shipped rule', 'bulk C store', 'per-warp C stores')[v]}]",
                   timed(lambda: cuda.matrix_mul_batched(a, bb, c, m, n, k)), nbytes=float(b) * (m * k + k * n + m * n) * es)
        del a, bb, c
        torch.cuda.empty_cache()
    set_tunable("gemm_jit_bulk_c", 0)


def g_shortk():
    """run-time specialised register tiles with K of one or two 16-byte vectors"""
    for (m, n, k, dt, log2) in ((8, 8, 4, np.float32, 22), (16, 16, 4, np.float32, 20), (8, 16, 8, np.float32, 21), (8, 8, 2, np.float64, 22),
                                (16, 8, 4, np.float64, 20), (12, 12, 4, np.float32, 21)):
        b, es = 1 << log2, np.dtype(dt).itemsize
        (_, a), (_, bb), (_, c) = uniform(b * m * k, dt, 1), uniform(b * k * n, dt, 2), uniform(b * m * n, dt, 3)
        report(f"gemm {m}x{k} . {k}x{n} {np.dtype(dt).name} batch 2^{log2}", timed(lambda: cuda.matrix_mul_batched(a, bb, c, m, n, k)),
               nbytes=float(b) * (m * k + k * n + m * n) * es)
        del a, bb, c
        torch.cuda.empty_cache()


GROUPS = {
    "shortk": g_shortk,
    "jitbulkc": g_jitbulkc,
    "gemm4d": g_gemm4d,
    "tinyvsjit": g_tinyvsjit,
    "lanevsstaged": g_lanevsstaged,
    "trmid": g_trmid,
    "oddsweep4": g_oddsweep4,
    "trsmall": g_trsmall,
    "bnorm": g_bnorm,
    "oddsweep3": g_oddsweep3,
    "oddsweep2": g_oddsweep2,
    "oddsweep": g_oddsweep,
    "odd": g_odd,
    "vec3": g_vec3,
    "small": g_small,
    "pageable": g_pageable,
    "misc": g_misc,
    "rect": g_rect,
    "gemv": g_gemv,
    "bdot": g_bdot,
    "transpose": g_transpose,
    "inplace": g_inplace,
    "crossover": g_crossover,
    "e2e": g_e2e,
    "dgemm": g_dgemm,
    "fused": g_fused,
    "gemm32d": lambda: g_gemm("gemm32d_variant", 32, np.float64, (0, 1)),
    "gemm32s": lambda: g_gemm("gemm32s_variant", 32, np.float32, (0, 1, 3, 4, 5, 0)),
    "gemm16s": lambda: g_gemm("gemm16s_variant", 16, np.float32, (0, 1, 2, 3, 4, 5, 6, 0)),
    "gemm64": lambda: (g_gemm("gemm32s_variant", 64, np.float32, (0,), 18), g_gemm("gemm32d_variant", 64, np.float64, (0,), 17),
                       g_gemm("gemm4s_variant", 8, np.float32, (0,), 22), g_gemm("gemm4s_variant", 8, np.float64, (0,), 22),
                       g_gemm("gemm4s_variant", 4, np.float64, (0,), 24)),
    "large": lambda: g_large("ffma", "gemm_large_variant", (0, 2)),  # 0 default, 2 no pre-transposition of the operands
    "tc": lambda: g_large("tcgen05", "tc_variant", (0, 1, 2, 3)),
}

if __name__ == "__main__":
    for g in (sys.argv[1:] or list(GROUPS)):
        GROUPS[g]()
