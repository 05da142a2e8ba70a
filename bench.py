#!/usr/bin/env python
"""bench.py -- the hot path of `slas_backend::Cuda` on N B200s, one JSON line on stdout.

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): batched static
4x4 f32 matrix_mul, batch 2^24 per GPU, synthetic U[-1,1) inputs generated in HBM.  A "step" is
one pass of the batched product over the whole batch.

  value     whole-job GFLOP/s (128 flop per matrix), operands resident in HBM, CUDA-event timed,
            barrier + synchronize on both sides, max over ranks.
  e2e       the same product through the reference-facing C ABI with HOST (pinned) operands:
            H2D of A and B and D2H of C inside the timed region.
  roofline  the batched small-GEMM kernel against the measured HBM copy bandwidth
            (MEASURED_PEAKS.json): algorithmic bytes = 192 B per matrix.
  cpu_baseline  the reference's own way of doing this on the host: a loop of cblas_sgemm calls
            (src/backends/blas.rs:94-109 behind Matrix::matrix_mul, src/tensor.rs:441-454), timed
            on this box's cores on a bounded sample.
  others    the remaining BASELINE configs measured the same way (dot/add GB/s, 16x16 / 32x32
            f32/f64 batched matmul + transpose, the single 4096^2 product, sharded dot).

`--impl reference` times only the CPU reference arm.  N > 1: launched by torchrun, one rank per GPU;
the batch shards by batch index with no data-path collective (weak scaling); the sharded dot in
`others` is the one op with a collective (one NCCL allreduce of an f64 partial).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_4X4 = 128
BYTES_PER_4X4 = 192


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=50)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    p.add_argument("--batch-log2", type=int, default=24)
    p.add_argument("--no-others", action="store_true", help="skip the secondary configs")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    return p.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


def ncu_traffic(kernel_prefix: str):
    """DRAM bytes (read + write) per launch of a kernel, from the newest committed `ncu --set full` summary
    (profiles/*_ncu_summary.csv, written by profiles/summarize.py).  None when no capture exists."""
    import csv
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_summary.csv")), reverse=True):
        with open(path) as f:
            rows = list(csv.reader(f))
        if not rows:
            continue
        hdr = rows[0]
        try:
            ir = next(i for i, h in enumerate(hdr) if h.startswith("dram__bytes_read.sum"))
            iw = next(i for i, h in enumerate(hdr) if h.startswith("dram__bytes_write.sum"))
        except StopIteration:
            continue
        scale = {"[Gbyte]": 1e9, "[Mbyte]": 1e6, "[Kbyte]": 1e3, "[byte]": 1.0}
        sr = next((v for k, v in scale.items() if k in hdr[ir]), 1.0)
        sw = next((v for k, v in scale.items() if k in hdr[iw]), 1.0)
        vals = [float(r[ir]) * sr + float(r[iw]) * sw for r in rows[1:] if r and r[0].startswith(kernel_prefix)]
        if vals:
            return {"bytes": sum(vals) / len(vals), "source": os.path.relpath(path, ROOT)}
    return None


# ---- clocks during the timed region --------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ---- CPU reference arm ------------------------------------------------------------------------------
def cpu_reference_4x4(sample_log2: int, steps: int, warmup: int):
    """Loop of cblas_sgemm calls over `2^sample_log2` contiguous 4x4 pairs: what a slas user runs today."""
    from oracle import blas_standin as blas
    n = 1 << sample_log2
    rng = np.random.default_rng(1)
    a = (rng.random(n * 16, dtype=np.float32) * 2 - 1)
    b = (rng.random(n * 16, dtype=np.float32) * 2 - 1)
    c = np.zeros(n * 16, np.float32)
    for _ in range(max(1, warmup)):
        blas.gemm_batched(a, b, c, n, 4, 4, 4, threads=1)
    t0 = time.perf_counter()
    for _ in range(steps):
        blas.gemm_batched(a, b, c, n, 4, 4, 4, threads=1)
    dt = (time.perf_counter() - t0) / steps
    cores = os.cpu_count() or 1
    t1 = time.perf_counter()
    blas.gemm_batched(a, b, c, n, 4, 4, 4, threads=cores)
    dt_all = time.perf_counter() - t1
    return {"gflops_1t": FLOP_PER_4X4 * n / dt / 1e9, "ms_per_step": dt * 1e3, "cores_available": cores,
            "gflops_all_cores": FLOP_PER_4X4 * n / dt_all / 1e9,
            "sample": f"2^{sample_log2} of the 4x4 f32 pairs per step (loop of cblas_sgemm, OpenBLAS from scipy, 1 thread: "
                      f"the reference is single-threaded; pthread split over {cores} cores reported as gflops_all_cores)"}


def cpu_reference_config0():
    """BASELINE configs[0]: f32 dot + add on LEN = 2^20, the reference's Rust backend (C restatement, the lane
    count an AVX2 build gives it) and its Blas backend (cblas_sdot) on this box's cores."""
    from oracle import blas_standin as blas
    from oracle import oracle
    n = 1 << 20
    rng = np.random.default_rng(0)
    a, b = rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32)

    def best(fn, reps=30):
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return float(np.median(ts))

    out = {"len": n, "cores_available": os.cpu_count() or 1}
    t = best(lambda: oracle.dot(a, b, lanes=8, fast=True))
    out["rust_backend_dot_gbs_1_thread"] = 8.0 * n / t / 1e9
    t = best(lambda: oracle.ewise("add", a, b, fast=True))
    out["rust_backend_add_gbs_1_thread"] = 12.0 * n / t / 1e9
    blas.set_threads(1)
    t = best(lambda: blas.dot(a, b))
    out["blas_backend_dot_gbs_1_thread"] = 8.0 * n / t / 1e9
    blas.set_threads(os.cpu_count() or 1)
    t = best(lambda: blas.dot(a, b))
    out["blas_backend_dot_gbs_all_cores"] = 8.0 * n / t / 1e9
    blas.set_threads(1)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_log2 = min(args.batch_log2, 20)
    steps = max(1, min(args.steps, 40))
    r = cpu_reference_4x4(sample_log2, steps, min(args.warmup, 3))
    line = {
        "impl": "reference", "metric": "batched matrix_mul GFLOP/s", "value": r["gflops_1t"], "unit": "GFLOP/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 3), "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"batched static 4x4 f32 matrix_mul, batch 2^{args.batch_log2} per GPU (configs[1])",
                   "reference_path": "Matrix::matrix_mul -> Blas::matrix_mul -> cblas_sgemm per object, host cores"},
        "cpu_baseline": {"value": r["gflops_1t"], "unit": "GFLOP/s", "cores": 1, "kind": "port", "sample": r["sample"],
                         "all_cores_value": r["gflops_all_cores"], "cores_available": r["cores_available"]},
        "e2e": {"value": r["gflops_1t"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- GPU arm ---------------------------------------------------------------------------------------------
def run_cuda(args):
    # stdout carries exactly one JSON line: NCCL's version/debug banner goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    import torch.distributed as dist

    import slas_b200 as sl
    from slas_b200 import _lib
    from slas_b200.static_vec import CudaVec

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)  # a real (non-default) stream shared by torch's events and the library
    torch.cuda.set_stream(stream)
    sl.set_stream(stream.cuda_stream)
    cuda = sl.Cuda()
    L = _lib.lib()
    peak_gbs, peak_src, peaks = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warmup):
        """W warm-ups, then exactly `steps` calls between barriers; CUDA events on the launching stream."""
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = sl.launch_count()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps, (sl.launch_count() - n0)

    def dev_uniform(n, dtype, seed, lo=-1.0, hi=1.0):
        t = torch.empty(n, device=dev, dtype=dtype)
        name = "slas_cuda_sfill_uniform" if dtype == torch.float32 else "slas_cuda_dfill_uniform"
        _lib.call(name, n, t.data_ptr(), seed, 0, lo, hi)
        return t

    def vec(t):  # torch owns the HBM; the C ABI sees a device pointer
        return CudaVec(t.numel(), np.float32 if t.dtype == torch.float32 else np.float64, _ptr=t.data_ptr(), _owner=t)

    # ---- headline: batched 4x4 f32, batch 2^k per GPU ----
    batch = 1 << args.batch_log2
    A = dev_uniform(batch * 16, torch.float32, 1 + rank)
    B = dev_uniform(batch * 16, torch.float32, 101 + rank)
    Cc = torch.empty(batch * 16, device=dev, dtype=torch.float32)
    vA, vB, vC = vec(A), vec(B), vec(Cc)

    def step():
        cuda.matrix_mul_batched(vA, vB, vC, 4, 4, 4)

    with ClockSampler(local_rank) as clk:
        ms, launches = timed(step, args.steps, args.warmup)
    clocks = clk.summary()
    value = FLOP_PER_4X4 * batch * world / (ms * 1e-3) / 1e9
    achieved_gbs = BYTES_PER_4X4 * batch / (ms * 1e-3) / 1e9  # one launch per step: kernel time == step time

    # ---- e2e: host (pinned) operands through the C ABI ----
    e2e = None
    if not args.no_e2e:
        nbytes = batch * 16 * 4
        hp = [C.c_void_p() for _ in range(3)]
        for p in hp:
            _lib.call("slas_cuda_malloc_host", C.byref(p), nbytes)
        hA, hB, hC = [np.ctypeslib.as_array((C.c_float * (batch * 16)).from_address(p.value)) for p in hp]
        _lib.call("slas_cuda_memcpy_d2h", hp[0].value, A.data_ptr(), nbytes)
        _lib.call("slas_cuda_memcpy_d2h", hp[1].value, B.data_ptr(), nbytes)
        e2e_steps = max(2, min(args.steps, 5))

        def e2e_step():
            cuda.matrix_mul_batched(hA, hB, hC, 4, 4, 4)  # returns when C is back in host memory

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
        # the PCIe roofline of that step, measured here: one pinned 1 GiB host->device copy on its own (the step
        # moves 2x the batch down and 1x up; both directions run concurrently, so H2D bounds it)
        h2d_gbs = None
        if rank == 0:
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            for _ in range(3):
                _lib.call("slas_cuda_memcpy_h2d", A.data_ptr(), hp[0].value, nbytes)
            h2d_gbs = 3 * nbytes / (time.perf_counter() - t1) / 1e9
        # spot-check the host result against the device-resident one
        _lib.call("slas_cuda_memcpy_d2h", hp[0].value, Cc.data_ptr(), 4096)
        assert np.array_equal(hA[:1024], hC[:1024]), "host-path result differs from the device-resident result"
        e2e = {"value": FLOP_PER_4X4 * batch * world / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s",
               "h2d_bytes_per_step": 2 * nbytes, "d2h_bytes_per_step": nbytes, "ms_per_step": e2e_ms, "steps": e2e_steps,
               "pcie_gbs": 3 * nbytes / (e2e_ms * 1e-3) / 1e9,
               "h2d_gbs_in_step": 2 * nbytes / (e2e_ms * 1e-3) / 1e9, "h2d_gbs_copy_alone": h2d_gbs,
               "api": "slas_cuda_sgemm_batched with pinned host pointers (3-slot H2D/kernel/D2H pipeline)"}
        for p in hp:
            _lib.call("slas_cuda_free_host", p.value)
        del hA, hB, hC

    # ---- others: the remaining BASELINE configs (rank-local numbers unless stated) ----
    others = []
    fma_peak = None
    if not args.no_others:
        del A, B, Cc, vA, vB, vC
        torch.cuda.empty_cache()
        ksteps = max(3, min(args.steps, 20))

        # FP32 / FP64 FMA ceilings measured on this box (register-only 8x8 outer-product loop): the secondary figure
        # BASELINE asks for next to the HBM fraction of the batched products
        fma_peak = {}
        for is_d, key in ((0, "f32"), (1, "f64")):
            t = C.c_double()
            _lib.call("slas_cuda_bench_fma_peak", is_d, C.byref(t))
            fma_peak[key] = t.value

        def record(name, ms_, flop=None, bytes_=None, bound="hbm", extra=None, fma=None):
            row = {"name": name, "ms": ms_}
            if flop is not None:
                row["gflops"] = flop / (ms_ * 1e-3) / 1e9
                if fma:
                    row["fma_pipe_frac_of_measured_peak"] = row["gflops"] / 1e3 / fma_peak[fma]
                    row["fma_pipe_frac_of_nominal"] = row["gflops"] / 1e3 / (74.5 if fma == "f32" else 37.2)
            if bytes_ is not None:
                row["gbs"] = bytes_ / (ms_ * 1e-3) / 1e9
                if bound == "hbm":
                    row["hbm_frac"] = row["gbs"] / peak_gbs
            if extra:
                row.update(extra)
            others.append(row)

        # config 3: batched 16x16 / 32x32, f32 / f64, matmul + transpose, batch 2^20
        for n_, tdt, name in ((16, torch.float32, "f32"), (32, torch.float32, "f32"), (16, torch.float64, "f64"),
                              (32, torch.float64, "f64")):
            bsz = 1 << 20
            es = 4 if tdt == torch.float32 else 8
            a = dev_uniform(bsz * n_ * n_, tdt, 2)
            b = dev_uniform(bsz * n_ * n_, tdt, 3)
            c = torch.empty_like(a)
            va, vb, vc = vec(a), vec(b), vec(c)
            ms_, _ = timed(lambda: cuda.matrix_mul_batched(va, vb, vc, n_, n_, n_), ksteps, 3)
            record(f"batched {n_}x{n_} {name} matrix_mul, batch 2^20", ms_, flop=2.0 * n_ ** 3 * bsz, bytes_=3.0 * n_ * n_ * es * bsz,
                   fma=name)
            ms_, _ = timed(lambda: cuda.transpose_batched(va, vc, n_ * n_, n_), ksteps, 3)
            record(f"batched {n_}x{n_} {name} transpose, batch 2^20", ms_, bytes_=2.0 * n_ * n_ * es * bsz)
            del a, b, c, va, vb, vc
        # the callers either side of the path: other static shapes, matrix_vector_mul, in-place transposes
        for (m_, n_, k_, tdt, name, log2, how) in ((3, 3, 3, torch.float32, "f32", 24, "thread per matrix"),
                                                   (2, 2, 3, torch.float32, "f32", 24, "thread per matrix"),
                                                   (8, 4, 16, torch.float32, "f32", 22, "specialised at run time (NVRTC)"),
                                                   (12, 12, 12, torch.float64, "f64", 20, "specialised at run time (NVRTC)")):
            bsz = 1 << log2
            es = 4 if tdt == torch.float32 else 8
            a, b = dev_uniform(bsz * m_ * k_, tdt, 2), dev_uniform(bsz * k_ * n_, tdt, 3)
            c = torch.empty(bsz * m_ * n_, device=dev, dtype=tdt)
            va, vb, vc = vec(a), vec(b), vec(c)
            ms_, _ = timed(lambda: cuda.matrix_mul_batched(va, vb, vc, m_, n_, k_), ksteps, 3)
            record(f"batched {m_}x{k_} . {k_}x{n_} {name} matrix_mul, batch 2^{log2} ({how})", ms_, flop=2.0 * m_ * n_ * k_ * bsz,
                   bytes_=float(m_ * k_ + k_ * n_ + m_ * n_) * es * bsz)
            del a, b, c, va, vb, vc
        for n_, tdt, name in ((32, torch.float32, "f32"), (32, torch.float64, "f64")):
            bsz = 1 << 20
            es = 4 if tdt == torch.float32 else 8
            a, x = dev_uniform(bsz * n_ * n_, tdt, 2), dev_uniform(bsz * n_, tdt, 3)
            y = torch.empty_like(x)
            va, vx, vy = vec(a), vec(x), vec(y)
            for ta_ in (False, True):
                ms_, _ = timed(lambda: cuda.matrix_vector_mul_batched(va, vx, vy, n_, n_, ta_), ksteps, 3)
                record(f"batched {n_}x{n_} {name} matrix_vector_mul, a_trans={ta_}, batch 2^20", ms_, flop=2.0 * n_ * n_ * bsz,
                       bytes_=float(n_ * n_ + 2 * n_) * es * bsz)
            ms_, _ = timed(lambda: cuda.transpose_inplace_batched(va, n_ * n_, n_), ksteps, 3)
            record(f"batched {n_}x{n_} {name} transpose_inplace (Matrix::deref_mut per object), batch 2^20", ms_, bytes_=2.0 * n_ * n_ * es * bsz)
            del a, x, y, va, vx, vy
        # 3-D geometry sizes: 3x3 matrices and 3-vectors (not whole 16-byte words: bulk-TMA staged thread-per-object kernels)
        bsz = 1 << 24
        a, x = dev_uniform(bsz * 9, torch.float32, 2), dev_uniform(bsz * 3, torch.float32, 3)
        y = torch.empty_like(x)
        o = torch.empty(bsz, device=dev, dtype=torch.float32)
        va, vx, vy, vo = vec(a), vec(x), vec(y), vec(o)
        ms_, _ = timed(lambda: cuda.matrix_vector_mul_batched(va, vx, vy, 3, 3, False), ksteps, 3)
        record("batched 3x3 f32 matrix_vector_mul (3x3 . 3-vector), batch 2^24", ms_, flop=18.0 * bsz, bytes_=60.0 * bsz)
        at = torch.empty_like(a)
        vat = vec(at)
        ms_, _ = timed(lambda: cuda.transpose_batched(va, vat, 9, 3), ksteps, 3)
        record("batched 3x3 f32 transpose, batch 2^24", ms_, bytes_=72.0 * bsz)
        ms_, _ = timed(lambda: cuda.dot_batched(vx, vy, vo, 3), ksteps, 3)
        record("batched 3-vector f32 dot, batch 2^24", ms_, bytes_=24.0 * bsz)
        ms_, _ = timed(lambda: cuda.normalize_batched(vx, 3), ksteps, 3)
        record("batched 3-vector f32 normalize, batch 2^24", ms_, bytes_=24.0 * bsz)
        del a, x, y, o, at, va, vx, vy, vo, vat
        # config 1: dot + add, f32.  LEN = 2^20 is L2-resident and launch-bound as one call, so report
        # (i) one 2^28-element call, (ii) 1024 distinct 2^20 pairs in one batched launch, (iii) call latency.
        n28 = 1 << 28
        a, b = dev_uniform(n28, torch.float32, 4, 0.0, 1.0), dev_uniform(n28, torch.float32, 5, 0.0, 1.0)
        c = torch.empty_like(a)
        out = torch.zeros(1024, device=dev, dtype=torch.float32)
        va, vb, vc, vout = vec(a), vec(b), vec(c), vec(out)
        ms_, _ = timed(lambda: _lib.call("slas_cuda_sdot", n28, a.data_ptr(), b.data_ptr(), out.data_ptr()), ksteps, 3)
        record("dot f32 LEN=2^28 (one call)", ms_, flop=2.0 * n28, bytes_=8.0 * n28)
        ms_, _ = timed(lambda: cuda.add(va, vb, vc), ksteps, 3)
        record("add f32 LEN=2^28 (one call)", ms_, flop=1.0 * n28, bytes_=12.0 * n28)
        ms_, _ = timed(lambda: _lib.call("slas_cuda_snrm2", n28, a.data_ptr(), out.data_ptr()), ksteps, 3)
        record("norm f32 LEN=2^28 (one call)", ms_, bytes_=4.0 * n28)
        ms_, _ = timed(lambda: cuda.normalize(va), ksteps, 3)
        record("normalize f32 LEN=2^28 (one call; 12 B/elem: read, read, write)", ms_, bytes_=12.0 * n28)
        ms_, _ = timed(lambda: _lib.call("slas_cuda_sadd_dot", n28, a.data_ptr(), b.data_ptr(), None, c.data_ptr(), out.data_ptr()), ksteps, 3)
        record("fused add -> dot f32 LEN=2^28 (intermediate never materialised, 12 B/elem instead of 20)", ms_, bytes_=12.0 * n28)
        ms_, _ = timed(lambda: _lib.call("slas_cuda_sdot_norms", n28, a.data_ptr(), b.data_ptr(), out.data_ptr()), ksteps, 3)
        record("fused dot + norm + norm f32 LEN=2^28 (8 B/elem instead of 16)", ms_, bytes_=8.0 * n28)
        l20 = 1 << 20
        # (ii) of SURVEY 8d config 1: 1024 DISTINCT 2^20-element pairs (8 GiB >> L2) in one batched launch
        a4 = dev_uniform(1024 * l20, torch.float32, 8, 0.0, 1.0)
        b4 = dev_uniform(1024 * l20, torch.float32, 9, 0.0, 1.0)
        ms_, _ = timed(lambda: cuda.dot_batched(vec(a4), vec(b4), vout, l20), ksteps, 3)
        record("dot f32 LEN=2^20 x 1024 distinct pairs (one batched launch, 8 GiB)", ms_, flop=2.0 * 1024 * l20, bytes_=8.0 * 1024 * l20)
        del a4, b4

        def dot_loop():
            for i in range(256):
                L.slas_cuda_sdot(l20, a.data_ptr() + 4 * i * l20, b.data_ptr() + 4 * i * l20, out.data_ptr() + 4 * i)

        ms_, _ = timed(dot_loop, 3, 1)
        record("dot f32 LEN=2^20, 256 distinct pairs, one call each (launch-bound)", ms_, bytes_=8.0 * 256 * l20,
               extra={"us_per_call": ms_ * 1e3 / 256})

        def add_loop():
            for i in range(256):
                L.slas_cuda_sadd(l20, a.data_ptr() + 4 * i * l20, b.data_ptr() + 4 * i * l20, c.data_ptr() + 4 * i * l20)

        ms_, _ = timed(add_loop, 3, 1)
        record("add f32 LEN=2^20, 256 distinct pairs, one call each (launch-bound)", ms_, bytes_=12.0 * 256 * l20,
               extra={"us_per_call": ms_ * 1e3 / 256})
        del a, b, c, out, va, vb, vc, vout
        torch.cuda.empty_cache()
        # config 5 (vector half): dot / normalize over 2^32 f32 elements IN TOTAL, sharded by contiguous slice,
        # each rank's slice resident in its own HBM, ONE NCCL allreduce of the f64 partial per op
        from slas_b200 import sharding
        total = 1 << 32
        lo_, hi_ = sharding.shard_range(total, 4, world, rank)
        local = hi_ - lo_
        a = torch.empty(local, device=dev, dtype=torch.float32)
        b = torch.empty(local, device=dev, dtype=torch.float32)
        res = np.zeros(1, np.float32)
        dres = torch.zeros(4, device=dev, dtype=torch.float32)
        # the collective two ways: fused into the reduction kernel over NVLink peer memory (p2p, the default)
        # and as a separate ncclAllReduce behind it
        for fused in ((True, False) if world > 1 else (True,)):
            how = sharding.init_comm(dist, fused=fused) if world > 1 else "single"
            _lib.call("slas_cuda_sfill_uniform", local, a.data_ptr(), 4, lo_, 0.0, 1.0)
            _lib.call("slas_cuda_sfill_uniform", local, b.data_ptr(), 5, lo_, 0.0, 1.0)
            ms_, _ = timed(lambda: L.slas_cuda_sdot_sharded(local, a.data_ptr(), b.data_ptr(), res.ctypes.data), 5, 2)
            # E[u*v] = 1/4 for independent U[0,1): the analytic check SURVEY 8d asks for at this size
            assert abs(float(res[0]) / total - 0.25) < 1e-3, f"sharded dot {res[0]} is not ~ 2^32/4"
            record(f"sharded dot f32, 2^32 elements over {world} GPU(s), collective: {how}, scalar returned to host", ms_,
                   flop=2.0 * total, bytes_=8.0 * total, bound="hbm-aggregate",
                   extra={"aggregate_hbm_frac": 8.0 * total / (ms_ * 1e-3) / 1e9 / (peak_gbs * world), "result": float(res[0])})
            ms_, _ = timed(lambda: L.slas_cuda_sdot_sharded(local, a.data_ptr(), b.data_ptr(), dres.data_ptr()), 5, 2)
            record(f"sharded dot f32, 2^32 elements over {world} GPU(s), collective: {how}, result left in HBM", ms_,
                   flop=2.0 * total, bytes_=8.0 * total, bound="hbm-aggregate",
                   extra={"aggregate_hbm_frac": 8.0 * total / (ms_ * 1e-3) / 1e9 / (peak_gbs * world)})
            ms_, _ = timed(lambda: L.slas_cuda_snormalize_sharded(local, a.data_ptr()), 5, 2)
            record(f"sharded normalize f32, 2^32 elements over {world} GPU(s), collective: {how}", ms_, bytes_=12.0 * total,
                   bound="hbm-aggregate", extra={"aggregate_hbm_frac": 12.0 * total / (ms_ * 1e-3) / 1e9 / (peak_gbs * world)})
            L.slas_cuda_snrm2_sharded(local, a.data_ptr(), res.ctypes.data)
            assert abs(float(res[0]) - 1.0) < 1e-3, f"norm after sharded normalize is {res[0]}"
            if world > 1:
                sharding.destroy_comm(dist)
        del dres
        del a, b
        torch.cuda.empty_cache()
        # config 4: single 4096^2 f32 product, FFMA path (and tcgen05 3xTF32 when built)
        if rank == 0:
            n_ = 4096
            a, b = dev_uniform(n_ * n_, torch.float32, 6), dev_uniform(n_ * n_, torch.float32, 7)
            c = torch.empty_like(a)
            va, vb, vc = vec(a), vec(b), vec(c)
            flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)

            def gemm():
                cuda.matrix_mul(va, vb, vc, n_, n_, n_, n_, n_, n_, False, False)

            sl.set_sgemm_path("ffma")
            for _ in range(2):
                gemm()
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                flush.zero_()  # > L2: evict the operands between timed iterations
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); gemm(); e1.record(stream)
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            record("single 4096^2 f32 matrix_mul, FFMA path (set_sgemm_path('ffma'))", float(np.median(ts)), flop=2.0 * n_ ** 3, bound="fp32", fma="f32")
            try:
                sl.set_sgemm_path("tcgen05")
                for _ in range(2):
                    gemm()
                torch.cuda.synchronize()
                n_before = sl.launch_count()
                ts = []
                for _ in range(5):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream); gemm(); e1.record(stream)
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                record("single 4096^2 f32 matrix_mul, tcgen05 3xTF32 path (the default at this size)", float(np.median(ts)), flop=2.0 * n_ ** 3, bound="tensor",
                       extra={"launches_per_call": (sl.launch_count() - n_before) / 5})
            finally:
                sl.set_sgemm_path("auto")
            del a, b, c, va, vb, vc, flush
        if world > 1:
            dist.barrier()

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_reference_4x4(min(args.batch_log2, 22), 3, 1)
        cpu = {"value": r["gflops_1t"], "unit": "GFLOP/s", "cores": 1, "kind": "port", "sample": r["sample"],
               "all_cores_value": r["gflops_all_cores"], "cores_available": r["cores_available"]}
        if not args.no_others:
            cpu["config0"] = cpu_reference_config0()

    if rank == 0:
        traffic = ncu_traffic("void gemm_small_kernel<float, SmallCfg<float, 4, 4, 4")
        line = {
            "metric": "batched matrix_mul GFLOP/s", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"batched static 4x4 f32 matrix_mul, batch 2^{args.batch_log2} per GPU (BASELINE configs[1])",
                       "layout": "[[f32; 16]; batch] contiguous row-major, no transposes", "parallelism": f"batch-index shards x{world}, no collective",
                       "l2": "working set 3 GiB per GPU >> 126 MB L2 (no flush needed)"},
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": peak_gbs, "unit": "GB/s", "frac": achieved_gbs / peak_gbs,
                         "traffic": (traffic or {}).get("bytes") if batch == 1 << 24 else None,
                         "traffic_source": (traffic or {}).get("source"),
                         "peak_source": peak_src, "kernel": "gemm_small_kernel<float, 4x4x4>",
                         "algorithmic_bytes_per_launch": BYTES_PER_4X4 * batch},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "fma_peak_tflops_measured": fma_peak if not args.no_others else None, "others": others,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
