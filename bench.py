#!/usr/bin/env python
"""bench.py -- the hot path of `slas_backend::Cuda` on N B200s, one JSON line on stdout.

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): batched static
4x4 f32 matrix_mul, batch 2^24 per GPU, synthetic U[-1,1) inputs generated in HBM.  A "step" is
one pass of the batched product over the whole batch.

  value     whole-job GFLOP/s (128 flop per matrix), operands resident in HBM, CUDA-event timed,
            barrier + synchronize on both sides, max over ranks.
  e2e       the same product through the reference-facing C ABI with HOST operands (pinned, and the
            pageable memory a slas StaticVec really is: `pageable`): H2D of A and B and D2H of C inside
            the timed region.
  roofline  the batched small-GEMM kernel against the measured HBM copy bandwidth
            (MEASURED_PEAKS.json): algorithmic bytes = 192 B per matrix.  `roofline.configs` carries the
            rest of the BASELINE metric, one entry per config with {ms, achieved, peak, frac, unit}:
            dot / add GB/s (2^28 single call, 2^20 x 1024 batched, 2^20 single calls eager / C loop /
            graph replay), batched 16x16 / 32x32 f32 / f64 matmul + transpose, the single 4096^2 product on the
            FFMA and the tcgen05 3xTF32 path (against FMA / TF32 peaks measured on this box), and config 5:
            sharded dot / normalize over 2^32 elements + batched 16x16 weak / strong scaling.
  cpu_baseline  the reference's own way of doing this on the host: a loop of cblas_sgemm calls
            (src/backends/blas.rs:94-109 behind Matrix::matrix_mul, src/tensor.rs:441-454), timed
            on this box's cores on a bounded sample.

`--impl reference` times only the CPU reference arm (same config, steps, warm-up and sample).  N > 1:
launched by torchrun, one rank per GPU; the batch shards by batch index with no data-path collective (weak
scaling); the sharded dot / normalize of config 5 are the ops with an exchange step, and this run ASSERTS their
parity across ranks (bit-identical results on every rank, equal to the rank-ordered sum of the ranks' f64
partials, an independent f64 truth of every slice, sampled slices of the normalize bit-exact) -- a failure
exits non-zero.
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
CPU_SAMPLE_LOG2 = 22  # pairs per CPU step: the same bounded sample in the cpu_baseline leg and the reference arm


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
    p.add_argument("--no-numa-bind", action="store_true",
                   help="N > 1: do not bind a rank to the CPUs NVML reports as local to its GPU")
    return p.parse_args()


def workload_config(batch_log2: int, world: int) -> dict:
    """The `config` object both arms print (identical for the same command line)."""
    return {"workload": f"batched static 4x4 f32 matrix_mul, batch 2^{batch_log2} per GPU (BASELINE configs[1])",
            "layout": "[[f32; 16]; batch] contiguous row-major, no transposes",
            "parallelism": f"batch-index shards x{world}, no collective",
            "l2": "working set 3 GiB per GPU >> 126 MB L2 (no flush needed)"}


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
    """Loop of cblas_sgemm calls over `2^sample_log2` contiguous 4x4 pairs: what a slas user runs today
    (Matrix::matrix_mul -> Blas::matrix_mul -> cblas_sgemm per object, src/tensor.rs:441-454,
    src/backends/blas.rs:94-109).  W untimed warm-up steps, then exactly `steps` timed ones."""
    from oracle import blas_standin as blas
    n = 1 << sample_log2
    rng = np.random.default_rng(1)
    a = (rng.random(n * 16, dtype=np.float32) * 2 - 1)
    b = (rng.random(n * 16, dtype=np.float32) * 2 - 1)
    c = np.zeros(n * 16, np.float32)
    for _ in range(warmup):
        blas.gemm_batched(a, b, c, n, 4, 4, 4, threads=1)
    t0 = time.perf_counter()
    for _ in range(steps):
        blas.gemm_batched(a, b, c, n, 4, 4, 4, threads=1)
    dt = (time.perf_counter() - t0) / steps
    cores = os.cpu_count() or 1
    blas.gemm_batched(a, b, c, n, 4, 4, 4, threads=cores)
    t1 = time.perf_counter()
    for _ in range(3):
        blas.gemm_batched(a, b, c, n, 4, 4, 4, threads=cores)
    dt_all = (time.perf_counter() - t1) / 3
    return {"gflops_1t": FLOP_PER_4X4 * n / dt / 1e9, "ms_per_step": dt * 1e3, "cores_available": cores,
            "gflops_all_cores": FLOP_PER_4X4 * n / dt_all / 1e9,
            "sample": f"2^{sample_log2} of the 4x4 f32 pairs per step (loop of cblas_sgemm, OpenBLAS from scipy; the reference's "
                      f"loop is serial and BLAS does not thread a 4x4 product: 1 busy core; a pthread split over {cores} cores, "
                      f"which the reference does not have, is reported as all_cores_value)"}


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
    r = cpu_reference_4x4(min(args.batch_log2, CPU_SAMPLE_LOG2), args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "batched matrix_mul GFLOP/s", "value": r["gflops_1t"], "unit": "GFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.batch_log2, args.gpus),
        "cpu_baseline": {"value": r["gflops_1t"], "unit": "GFLOP/s", "cores": 1, "kind": "port", "sample": r["sample"],
                         "all_cores_value": r["gflops_all_cores"], "cores_available": r["cores_available"],
                         "reference_path": "Matrix::matrix_mul -> Blas::matrix_mul -> cblas_sgemm per object, host cores"},
        "e2e": {"value": r["gflops_1t"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- GPU arm ---------------------------------------------------------------------------------------------
def run_cuda(args):
    # stdout carries exactly one JSON line: everything else any library prints (NCCL's version banner ...) goes to
    # stderr -- file descriptor 1 is pointed at stderr for the run and the line is written to the saved descriptor
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist

    import slas_b200 as sl
    from oracle import oracle
    from slas_b200 import _lib, sharding
    from slas_b200.static_vec import CudaVec

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_note = ("not bound (one rank)" if world == 1 else "not bound (--no-numa-bind)" if args.no_numa_bind
                 else bind_to_gpu_local_cpus(local_rank))
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")  # a barrier that keeps the GPUs idle (the one-process rows below)
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

    def gather_f64(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world == 1:
            return [t.cpu().numpy()]
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        return [x.cpu().numpy() for x in g]

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

    def check_sampled_products(a, b, c, size, nobj, what, samples=2048, seed=0):
        """Sampled objects of a batched product against the oracle (outside every timed region): 1e-5 * sqrt(K)
        relative to sum|a||b| per element (north_star's bar).  Returns the worst observed error / bar."""
        g = torch.Generator(device="cpu").manual_seed(seed)
        idx = torch.randint(0, nobj, (min(samples, nobj),), generator=g).to(dev)
        e = size * size
        ha = a.view(nobj, e).index_select(0, idx).cpu().numpy().ravel()
        hb = b.view(nobj, e).index_select(0, idx).cpu().numpy().ravel()
        hc = c.view(nobj, e).index_select(0, idx).cpu().numpy().ravel()
        want = oracle.gemm_batched(ha, hb, idx.numel(), size, size, size)
        mag = np.einsum("bik,bkj->bij", np.abs(ha.reshape(-1, size, size)).astype(np.float64),
                        np.abs(hb.reshape(-1, size, size)).astype(np.float64)).ravel()
        tol = (1e-5 if ha.dtype == np.float32 else 1e-13) * np.sqrt(size) * mag
        err = np.abs(hc.astype(np.float64) - want.astype(np.float64))
        ratio = float(np.max(err / np.maximum(tol, 1e-300)))
        assert ratio <= 1.0, f"{what}: sampled objects differ from the oracle (error / bar = {ratio})"
        return ratio

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
    headline_parity = check_sampled_products(A, B, Cc, 4, batch, "headline 4x4 product", samples=8192, seed=rank)

    # ---- e2e: host operands through the C ABI ----
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

        def time_host_steps(fa, fb, fc):
            def e2e_step():
                cuda.matrix_mul_batched(fa, fb, fc, 4, 4, 4)  # returns when C is back in host memory
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            barrier()
            return max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps

        e2e_ms = time_host_steps(hA, hB, hC)
        # spot-check the host result against the device-resident one
        ref_c = Cc[:1 << 16].cpu().numpy()
        assert np.array_equal(ref_c, hC[:1 << 16]), "host-path result differs from the device-resident result"
        # the memory a slas StaticVec really is (static_vec.rs:126: [T; LEN] / Vec): pageable
        pA, pB = np.array(hA, copy=True), np.array(hB, copy=True)
        pC = np.zeros_like(pA)
        pageable_ms = time_host_steps(pA, pB, pC)
        assert np.array_equal(pC, hC), "pageable-operand result differs from the pinned-operand result"
        # the same pageable arrays page-locked IN PLACE once (slas_cuda_host_register): what a program does with a
        # long-lived Vec it wants the DMA engines to read directly
        t_reg = time.perf_counter()
        for arr in (pA, pB, pC):
            _lib.call("slas_cuda_host_register", arr.ctypes.data, arr.nbytes)
        t_reg = time.perf_counter() - t_reg
        pC[:] = 0
        registered_ms = time_host_steps(pA, pB, pC)
        assert np.array_equal(pC, hC), "registered-operand result differs from the pinned-operand result"
        for arr in (pA, pB, pC):
            _lib.call("slas_cuda_host_unregister", arr.ctypes.data)
        del pA, pB, pC
        # the PCIe roofline of that step, measured here: pinned copies on their own (torch pinned buffers + cudaMemcpyAsync,
        # nothing of the library), one direction and both at once, all ranks at the same time (N > 1: the host-side
        # ceiling the e2e number runs into)
        for p in hp[:2]:
            _lib.call("slas_cuda_free_host", p.value)
        del hA, hB
        pin_in = torch.empty(batch * 16, dtype=torch.float32, pin_memory=True)
        pin_out = torch.empty(batch * 16, dtype=torch.float32, pin_memory=True)
        s2 = torch.cuda.Stream(device=dev)

        def copy_time(h2d: bool, d2h: bool, reps=3):
            barrier()
            t1 = time.perf_counter()
            for _ in range(reps):
                if d2h:
                    with torch.cuda.stream(s2):
                        pin_out.copy_(Cc, non_blocking=True)
                if h2d:
                    A.copy_(pin_in, non_blocking=True)
            barrier()
            return (time.perf_counter() - t1) / reps

        copy_time(True, True, 1)
        h2d_gbs = nbytes / copy_time(True, False) / 1e9
        d2h_gbs = nbytes / copy_time(False, True) / 1e9
        bidir_s = copy_time(True, True)
        del pin_in, pin_out
        e2e = {"value": FLOP_PER_4X4 * batch * world / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s",
               "h2d_bytes_per_step": 2 * nbytes, "d2h_bytes_per_step": nbytes, "ms_per_step": e2e_ms, "steps": e2e_steps,
               "pcie_gbs": 3 * nbytes / (e2e_ms * 1e-3) / 1e9,
               "h2d_gbs_in_step": 2 * nbytes / (e2e_ms * 1e-3) / 1e9,
               "pageable": {"value": FLOP_PER_4X4 * batch * world / (pageable_ms * 1e-3) / 1e9, "ms_per_step": pageable_ms,
                            "frac_of_pinned": e2e_ms / pageable_ms,
                            "how": "pageable numpy operands: copy threads stage them through a pinned bounce ring (3 passes over host "
                                   "memory per byte instead of 1: host-DRAM bound)"},
               "registered_in_place": {"value": FLOP_PER_4X4 * batch * world / (registered_ms * 1e-3) / 1e9, "ms_per_step": registered_ms,
                                       "frac_of_pinned": e2e_ms / registered_ms, "one_time_register_s": t_reg,
                                       "how": "the same pageable arrays after slas_cuda_host_register (page-locked in place, once)"},
               "copy_only": {"ranks_copying_at_once": world, "h2d_gbs_per_gpu": h2d_gbs, "d2h_gbs_per_gpu": d2h_gbs,
                             "bidir_gbs_each_way_per_gpu": nbytes / bidir_s / 1e9,
                             "step_floor_ms": max(2 * nbytes / (h2d_gbs * 1e9), nbytes / (d2h_gbs * 1e9)) * 1e3,
                             "what": "pinned 1 GiB copies alone, every rank at the same time (barrier-bracketed): the link / "
                                     "host-memory ceiling of the e2e step, which moves 2 GiB down and 1 GiB up per GPU"},
               "api": "slas_cuda_sgemm_batched with host pointers (3-slot H2D/kernel/D2H pipeline)"}
        e2e["frac_of_copy_floor"] = e2e["copy_only"]["step_floor_ms"] / e2e_ms
        # all links together: what the step moves per second over the whole box against what bare copies reach
        e2e["aggregate_gbs_in_step"] = 3 * nbytes * world / (e2e_ms * 1e-3) / 1e9
        e2e["copy_only"]["aggregate_gbs"] = {"h2d_alone": h2d_gbs * world, "d2h_alone": d2h_gbs * world,
                                             "both_directions_at_once": 2 * nbytes / bidir_s / 1e9 * world}
        # with several ranks the host's memory system, not a link, is the limit: bare copies in both directions at once
        # reach less than the two directions alone add up to, and the floor above (directions alone) overstates what a
        # step can do.  >= 1.0 here means the step moves bytes at least as fast as bare bidirectional copies do.  (On ONE
        # GPU the step's 2 : 1 mix of directions is bound by the down direction -- frac_of_copy_floor is the figure there --
        # and cannot reach the 1 : 1 aggregate: ~0.8.)
        e2e["vs_bare_copies_both_directions"] = e2e["aggregate_gbs_in_step"] / e2e["copy_only"]["aggregate_gbs"]["both_directions_at_once"]
        e2e["cpu_affinity"] = numa_note
        _lib.call("slas_cuda_free_host", hp[2].value)
        del hC

    # ---- the rest of the metric: roofline.configs ----
    configs = {}
    fma_peak = None
    tf32_peak = None
    if not args.no_others:
        del A, B, Cc, vA, vB, vC
        torch.cuda.empty_cache()
        ksteps = max(3, min(args.steps, 20))

        # FP32 / FP64 FMA and TF32 tensor ceilings measured on this box, with the SM clock each one ran at
        fma_peak = {}
        for is_d, key in ((0, "f32"), (1, "f64")):
            t = C.c_double()
            with ClockSampler(local_rank) as ck:
                _lib.call("slas_cuda_bench_fma_peak", is_d, C.byref(t))
            fma_peak[key] = t.value
            fma_peak[key + "_sm_mhz"] = ck.summary()["sm_mhz"]
        t = C.c_double()
        with ClockSampler(local_rank) as ck:
            _lib.call("slas_cuda_bench_tf32_peak", C.byref(t))
        tf32_peak = {"tflops": t.value, "sm_mhz": ck.summary()["sm_mhz"],
                     "how": "tcgen05.mma.kind::tf32 M128 N256 K8 back to back from resident smem tiles, one CTA per SM"}
        NOMINAL = {"f32": 74.5, "f64": 37.2}

        def rec(key, ms_, what, bytes_=None, flop=None, bound="hbm", fma=None, extra=None, ranks=1):
            """One entry of roofline.configs.  bound hbm: achieved GB/s against the measured copy bandwidth
            (x ranks for whole-job rows); fp32 / fp64: TFLOP/s against the FMA peak measured here; tensor: executed
            TF32 TFLOP/s against the measured tcgen05 peak."""
            row = {"what": what, "ms": ms_, "bound": bound}
            if bound == "hbm":
                row.update(achieved=bytes_ / (ms_ * 1e-3) / 1e9, peak=peak_gbs * ranks, unit="GB/s")
            elif bound in ("fp32", "fp64"):
                k = "f32" if bound == "fp32" else "f64"
                row.update(achieved=flop / (ms_ * 1e-3) / 1e12, peak=fma_peak[k], unit="TFLOP/s",
                           frac_of_nominal=flop / (ms_ * 1e-3) / 1e12 / NOMINAL[k])
            elif bound == "tensor":
                row.update(achieved=3.0 * flop / (ms_ * 1e-3) / 1e12, peak=tf32_peak["tflops"], unit="TFLOP/s executed (3 MMAs per product)",
                           useful_tflops=flop / (ms_ * 1e-3) / 1e12)
            row["frac"] = row["achieved"] / row["peak"]
            if flop is not None and bound == "hbm":
                row["gflops"] = flop / (ms_ * 1e-3) / 1e9
                if fma:
                    row["fma_pipe_frac_of_measured_peak"] = row["gflops"] / 1e3 / fma_peak[fma]
            if extra:
                row.update(extra)
            configs[key] = row

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
            par = check_sampled_products(a, b, c, n_, bsz, f"batched {n_}x{n_} {name}", samples=256)
            rec(f"mm{n_}_{name}", ms_, f"batched {n_}x{n_} {name} matrix_mul, batch 2^20", flop=2.0 * n_ ** 3 * bsz,
                bytes_=3.0 * n_ * n_ * es * bsz, fma=name, extra={"parity_error_over_bar": par})
            ms_, _ = timed(lambda: cuda.transpose_batched(va, vc, n_ * n_, n_), ksteps, 3)
            rec(f"tr{n_}_{name}", ms_, f"batched {n_}x{n_} {name} transpose, batch 2^20", bytes_=2.0 * n_ * n_ * es * bsz)
            del a, b, c, va, vb, vc
        # the callers either side of the path: other static shapes, matrix_vector_mul, in-place transposes
        for (m_, n_, k_, tdt, name, log2, how) in ((3, 3, 3, torch.float32, "f32", 24, "thread per matrix"),
                                                   (8, 4, 16, torch.float32, "f32", 22, "specialised at run time (NVRTC)"),
                                                   (12, 12, 12, torch.float64, "f64", 20, "specialised at run time (NVRTC)")):
            bsz = 1 << log2
            es = 4 if tdt == torch.float32 else 8
            a, b = dev_uniform(bsz * m_ * k_, tdt, 2), dev_uniform(bsz * k_ * n_, tdt, 3)
            c = torch.empty(bsz * m_ * n_, device=dev, dtype=tdt)
            va, vb, vc = vec(a), vec(b), vec(c)
            ms_, _ = timed(lambda: cuda.matrix_mul_batched(va, vb, vc, m_, n_, k_), ksteps, 3)
            rec(f"mm{m_}x{k_}x{n_}_{name}", ms_, f"batched {m_}x{k_} . {k_}x{n_} {name} matrix_mul, batch 2^{log2} ({how})",
                flop=2.0 * m_ * n_ * k_ * bsz, bytes_=float(m_ * k_ + k_ * n_ + m_ * n_) * es * bsz)
            del a, b, c, va, vb, vc
        for n_, tdt, name in ((32, torch.float32, "f32"), (32, torch.float64, "f64")):
            bsz = 1 << 20
            es = 4 if tdt == torch.float32 else 8
            a, x = dev_uniform(bsz * n_ * n_, tdt, 2), dev_uniform(bsz * n_, tdt, 3)
            y = torch.empty_like(x)
            va, vx, vy = vec(a), vec(x), vec(y)
            for ta_ in (False, True):
                ms_, _ = timed(lambda: cuda.matrix_vector_mul_batched(va, vx, vy, n_, n_, ta_), ksteps, 3)
                rec(f"gemv{n_}_{name}_{'t' if ta_ else 'n'}", ms_, f"batched {n_}x{n_} {name} matrix_vector_mul, a_trans={ta_}, batch 2^20",
                    flop=2.0 * n_ * n_ * bsz, bytes_=float(n_ * n_ + 2 * n_) * es * bsz)
            ms_, _ = timed(lambda: cuda.transpose_inplace_batched(va, n_ * n_, n_), ksteps, 3)
            rec(f"tr{n_}_{name}_inplace", ms_, f"batched {n_}x{n_} {name} transpose_inplace (Matrix::deref_mut per object), batch 2^20",
                bytes_=2.0 * n_ * n_ * es * bsz)
            del a, x, y, va, vx, vy
        # 3-D geometry sizes: 3x3 matrices and 3-vectors (not whole 16-byte words: bulk-TMA staged thread-per-object kernels)
        bsz = 1 << 24
        a, x = dev_uniform(bsz * 9, torch.float32, 2), dev_uniform(bsz * 3, torch.float32, 3)
        y = torch.empty_like(x)
        o = torch.empty(bsz, device=dev, dtype=torch.float32)
        va, vx, vy, vo = vec(a), vec(x), vec(y), vec(o)
        ms_, _ = timed(lambda: cuda.matrix_vector_mul_batched(va, vx, vy, 3, 3, False), ksteps, 3)
        rec("gemv3_f32", ms_, "batched 3x3 f32 matrix_vector_mul (3x3 . 3-vector), batch 2^24", flop=18.0 * bsz, bytes_=60.0 * bsz)
        at = torch.empty_like(a)
        vat = vec(at)
        ms_, _ = timed(lambda: cuda.transpose_batched(va, vat, 9, 3), ksteps, 3)
        rec("tr3_f32", ms_, "batched 3x3 f32 transpose, batch 2^24", bytes_=72.0 * bsz)
        ms_, _ = timed(lambda: cuda.dot_batched(vx, vy, vo, 3), ksteps, 3)
        rec("dot3_f32", ms_, "batched 3-vector f32 dot, batch 2^24", bytes_=28.0 * bsz)
        ms_, _ = timed(lambda: cuda.normalize_batched(vx, 3), ksteps, 3)
        rec("normalize3_f32", ms_, "batched 3-vector f32 normalize, batch 2^24", bytes_=24.0 * bsz)
        del a, x, y, o, at, va, vx, vy, vo, vat
        # config 1: dot + add, f32.  LEN = 2^20 is L2-resident and launch-bound as one call, so report
        # (i) one 2^28-element call, (ii) 1024 distinct 2^20 pairs in one batched launch, (iii) single calls on 256
        # distinct pairs (2 GiB >> L2): eager from Python, the same loop from C, and replayed from a CUDA graph.
        n28 = 1 << 28
        a, b = dev_uniform(n28, torch.float32, 4, 0.0, 1.0), dev_uniform(n28, torch.float32, 5, 0.0, 1.0)
        c = torch.empty_like(a)
        out = torch.zeros(1024, device=dev, dtype=torch.float32)
        va, vb, vc, vout = vec(a), vec(b), vec(c), vec(out)
        ms_, _ = timed(lambda: _lib.call("slas_cuda_sdot", n28, a.data_ptr(), b.data_ptr(), out.data_ptr()), ksteps, 3)
        rec("dot_2^28", ms_, "dot f32 LEN=2^28 (one call)", flop=2.0 * n28, bytes_=8.0 * n28)
        ms_, _ = timed(lambda: cuda.add(va, vb, vc), ksteps, 3)
        rec("add_2^28", ms_, "add f32 LEN=2^28 (one call)", flop=1.0 * n28, bytes_=12.0 * n28)
        ms_, _ = timed(lambda: _lib.call("slas_cuda_snrm2", n28, a.data_ptr(), out.data_ptr()), ksteps, 3)
        rec("norm_2^28", ms_, "norm f32 LEN=2^28 (one call)", bytes_=4.0 * n28)
        ms_, _ = timed(lambda: cuda.normalize(va), ksteps, 3)
        rec("normalize_2^28", ms_, "normalize f32 LEN=2^28 (one call; 12 B/elem: read, read, write)", bytes_=12.0 * n28)
        ms_, _ = timed(lambda: _lib.call("slas_cuda_sadd_dot", n28, a.data_ptr(), b.data_ptr(), None, c.data_ptr(), out.data_ptr()), ksteps, 3)
        rec("add_dot_2^28", ms_, "fused add -> dot f32 LEN=2^28 (intermediate never materialised, 12 B/elem instead of 20)", bytes_=12.0 * n28)
        ms_, _ = timed(lambda: _lib.call("slas_cuda_sdot_norms", n28, a.data_ptr(), b.data_ptr(), out.data_ptr()), ksteps, 3)
        rec("dot_norms_2^28", ms_, "fused dot + norm + norm f32 LEN=2^28 (8 B/elem instead of 16)", bytes_=8.0 * n28)
        l20 = 1 << 20
        a4 = dev_uniform(1024 * l20, torch.float32, 8, 0.0, 1.0)
        b4 = dev_uniform(1024 * l20, torch.float32, 9, 0.0, 1.0)
        ms_, _ = timed(lambda: cuda.dot_batched(vec(a4), vec(b4), vout, l20), ksteps, 3)
        rec("dot_2^20x1024", ms_, "dot f32 LEN=2^20 x 1024 distinct pairs (one batched launch, 8 GiB)", flop=2.0 * 1024 * l20, bytes_=8.0 * 1024 * l20)
        del a4, b4

        # Back-to-back short calls overlap (programmatic dependent launch) when nothing but this library enqueues on the
        # stream between them: automatic on the library's own stream; on a stream shared with another library (here:
        # torch's, for the events) the caller vouches for it with the tunable "pdl" = 2.  Both are reported.
        ncalls = 256
        for opname, opid, bytes_per in (("dot", 0, 8.0), ("add", 1, 12.0)):
            dst = out if opid == 0 else c

            def py_loop():
                for i in range(ncalls):
                    if opid == 0:
                        L.slas_cuda_sdot(l20, a.data_ptr() + 4 * i * l20, b.data_ptr() + 4 * i * l20, out.data_ptr() + 4 * i)
                    else:
                        L.slas_cuda_sadd(l20, a.data_ptr() + 4 * i * l20, b.data_ptr() + 4 * i * l20, c.data_ptr() + 4 * i * l20)

            def c_loop():
                _lib.call("slas_cuda_bench_call_loop", opid, l20, a.data_ptr(), b.data_ptr(), dst.data_ptr(), ncalls, l20)

            ms_py0, _ = timed(py_loop, 3, 1)
            ms_c0, _ = timed(c_loop, 5, 2)
            with sl.Graph() as g:
                c_loop()
            ms_g0, _ = timed(g.launch, 10, 3)
            del g
            _lib.call("slas_cuda_set_tunable", b"pdl", 2)
            ms_py, _ = timed(py_loop, 3, 1)
            ms_c, _ = timed(c_loop, 5, 2)
            with sl.Graph() as g:
                c_loop()
            ms_g, _ = timed(g.launch, 10, 3)
            del g
            _lib.call("slas_cuda_set_tunable", b"pdl", 0)
            rec(f"{opname}_2^20_single", ms_g / ncalls, f"{opname} f32 LEN=2^20, one call per pair, {ncalls} distinct pairs (2-3 GiB >> L2), replayed from one "
                f"CUDA graph (slas_cuda_graph_*), back-to-back calls overlapping (dependent launch): per-call time", bytes_=bytes_per * l20,
                extra={"us_per_call_graph": ms_g * 1e3 / ncalls, "us_per_call_c_loop": ms_c * 1e3 / ncalls,
                       "us_per_call_python_loop": ms_py * 1e3 / ncalls,
                       "frac_c_loop": bytes_per * l20 / (ms_c * 1e-3 / ncalls) / 1e9 / peak_gbs,
                       "shared_stream_no_overlap": {"us_per_call_graph": ms_g0 * 1e3 / ncalls, "us_per_call_c_loop": ms_c0 * 1e3 / ncalls,
                                                    "us_per_call_python_loop": ms_py0 * 1e3 / ncalls},
                       "hbm_floor_us": bytes_per * l20 / (peak_gbs * 1e9) * 1e6})
        # normalize of LEN 2^20 vectors, one call each: one launch with the vector held in registers between the norm and
        # the division (8 B per element), against the two-pass path (reduce + scale: 12 B per element, two launches)
        def norm_loop():
            _lib.call("slas_cuda_bench_call_loop", 3, l20, a.data_ptr(), b.data_ptr(), c.data_ptr(), ncalls, l20)

        c.copy_(a)
        _lib.call("slas_cuda_set_tunable", b"pdl", 2)
        with sl.Graph() as g:
            norm_loop()
        ms_n, _ = timed(g.launch, 10, 3)
        del g
        _lib.call("slas_cuda_set_tunable", b"normalize_fused", 1)
        with sl.Graph() as g:
            norm_loop()
        ms_n2, _ = timed(g.launch, 10, 3)
        del g
        _lib.call("slas_cuda_set_tunable", b"normalize_fused", 0)
        _lib.call("slas_cuda_set_tunable", b"pdl", 0)
        rec("normalize_2^20_single", ms_n / ncalls, f"normalize f32 LEN=2^20, one call per vector, {ncalls} distinct vectors, replayed from one CUDA graph: "
            f"single-launch kernel (vector in registers between norm and division), 8 B per element", bytes_=8.0 * l20,
            extra={"us_per_call": ms_n * 1e3 / ncalls, "us_per_call_two_pass": ms_n2 * 1e3 / ncalls})
        # the reference's own calling pattern: `let x = a.dot(&b)` -- the scalar comes back to the host after EVERY call
        # (tests/src/main.rs:589-601 benches exactly this loop).  Device-resident operands, distinct pairs, C loop.
        def sync_loop():
            _lib.call("slas_cuda_bench_call_loop", 2, l20, a.data_ptr(), b.data_ptr(), None, ncalls, l20)

        ms_sync, _ = timed(sync_loop, 5, 2)
        _lib.call("slas_cuda_set_tunable", b"host_result", 1)   # the round-1 way: device scalar + D2H copy + stream sync
        ms_sync_copy, _ = timed(sync_loop, 5, 2)
        _lib.call("slas_cuda_set_tunable", b"host_result", 0)
        rec("dot_2^20_single_host_result", ms_sync / ncalls, f"dot f32 LEN=2^20 returning its scalar to the HOST after every call (`a.dot(&b)`), {ncalls} distinct "
            f"device-resident pairs, C loop: per-call latency; the kernel stores the scalar into mapped pinned memory and the host spins on it",
            bytes_=8.0 * l20, extra={"us_per_call": ms_sync * 1e3 / ncalls, "us_per_call_with_d2h_copy_and_stream_sync": ms_sync_copy * 1e3 / ncalls})
        # ... and with the operands where a StaticCowVec keeps them: in pageable HOST memory (8 MiB down per call)
        ha_, hb_ = a[:8 * l20].cpu().numpy(), b[:8 * l20].cpu().numpy()
        hres = np.zeros(1, np.float32)

        def host_operand_loop():
            for i in range(8):
                L.slas_cuda_sdot(l20, ha_.ctypes.data + 4 * i * l20, hb_.ctypes.data + 4 * i * l20, hres.ctypes.data)

        host_operand_loop()
        t0_ = time.perf_counter()
        for _ in range(5):
            host_operand_loop()
        us_host = (time.perf_counter() - t0_) / 40 * 1e6
        configs["dot_2^20_single_host_result"]["us_per_call_host_resident_pageable_operands"] = us_host
        configs["dot_2^20_single_host_result"]["host_operands_gbs"] = 8.0 * l20 / (us_host * 1e-6) / 1e9
        del ha_, hb_
        # phases of one cold 2^20 dot kernel (%globaltimer inside the kernel)
        _lib.call("slas_cuda_set_tunable", b"reduce_phases", 1)
        ph = (C.c_uint64 * 5)()
        stamps = []
        for i in range(8):
            L.slas_cuda_sdot(l20, a.data_ptr() + 4 * (i + 100) * l20, b.data_ptr() + 4 * (i + 100) * l20, out.data_ptr())
            _lib.call("slas_cuda_get_reduce_phases", ph)
            stamps.append([int(ph[j]) - int(ph[0]) for j in range(5)])
        _lib.call("slas_cuda_set_tunable", b"reduce_phases", 0)
        configs["dot_2^20_single"]["kernel_phases_ns"] = {
            "loads_and_fmas_done": float(np.median([s[1] for s in stamps])), "block_sum_done": float(np.median([s[2] for s in stamps])),
            "all_partials_in": float(np.median([s[3] for s in stamps])), "result_written": float(np.median([s[4] for s in stamps]))}
        del a, b, c, out, va, vb, vc, vout
        torch.cuda.empty_cache()

        # config 5 (matrix half): batched 16x16 f32 matrix_mul over the N GPUs, weak (2^20 per GPU) and strong (2^20 in
        # total) -- batch-index shards, no collective; whole-job numbers (max over ranks), sampled objects vs the oracle
        for how, total in (("weak", (1 << 20) * world), ("strong", 1 << 20)):
            blo, bhi = sharding.batch_range(total, world, rank)
            nloc = bhi - blo
            a = torch.empty(nloc * 256, device=dev, dtype=torch.float32)
            b = torch.empty(nloc * 256, device=dev, dtype=torch.float32)
            # object i of the whole job carries the same values whatever the rank count (counter-based fill at its offset)
            _lib.call("slas_cuda_sfill_uniform", nloc * 256, a.data_ptr(), 12, blo * 256, -1.0, 1.0)
            _lib.call("slas_cuda_sfill_uniform", nloc * 256, b.data_ptr(), 13, blo * 256, -1.0, 1.0)
            c = torch.empty_like(a)
            va, vb, vc = vec(a), vec(b), vec(c)
            ms_, _ = timed(lambda: cuda.matrix_mul_batched(va, vb, vc, 16, 16, 16), ksteps, 3)
            par = check_sampled_products(a, b, c, 16, nloc, f"16x16 {how} shard of rank {rank}", samples=256, seed=rank)
            # a checksum of checksums: the sum over all C of the whole job must not depend on how it was sharded
            cs = gather_f64([float(c.double().sum().item())])
            rec(f"mm16_f32_{how}", ms_, f"batched 16x16 f32 matrix_mul, {total} objects over {world} GPU(s) ({how} scaling: "
                f"{nloc} per GPU), whole job, max over ranks", flop=2.0 * 4096 * total, bytes_=3072.0 * total, ranks=world,
                extra={"parity_error_over_bar": par, "objects_total": total, "checksum_of_c": float(sum(float(x[0]) for x in cs))})
            del a, b, c, va, vb, vc
        torch.cuda.empty_cache()

        # config 5 (vector half): dot / normalize over 2^32 f32 elements IN TOTAL, sharded by contiguous slice, each
        # rank's slice resident in its own HBM, ONE exchange of the f64 partial per op -- with the parity asserts
        total = 1 << 32
        lo_, hi_ = sharding.shard_range(total, 4, world, rank)
        local = hi_ - lo_
        a = torch.empty(local, device=dev, dtype=torch.float32)
        b = torch.empty(local, device=dev, dtype=torch.float32)
        res = np.zeros(1, np.float32)
        dres = torch.zeros(4, device=dev, dtype=torch.float32)
        part = np.zeros(1, np.float64)

        def f64_truth(x, y):
            """independent f64 sum of x*y over this rank's slice (torch, harness side, chunked)"""
            s = 0.0
            step_ = 1 << 26
            for o in range(0, x.numel(), step_):
                s += float((x[o:o + step_].double() * y[o:o + step_].double()).sum().item())
            return s

        for fused in ((True, False) if world > 1 else (True,)):
            how = sharding.init_comm(dist, fused=fused) if world > 1 else "single"
            _lib.call("slas_cuda_sfill_uniform", local, a.data_ptr(), 4, lo_, 0.0, 1.0)
            _lib.call("slas_cuda_sfill_uniform", local, b.data_ptr(), 5, lo_, 0.0, 1.0)
            # the exchange's inputs: every rank's own f64 partial, checked against an independent f64 truth of the slice
            # (a dropped or doubled slice of 1e-6 of the data fails) and added in rank order
            _lib.call("slas_cuda_sdot_partial", local, a.data_ptr(), b.data_ptr(), part.ctypes.data)
            truth = f64_truth(a, b)
            assert abs(part[0] - truth) <= 1e-6 * truth, f"rank {rank}: partial {part[0]} vs f64 truth {truth}"
            parts = gather_f64([float(part[0])])
            sum_parts = 0.0
            for p in parts:
                sum_parts += float(p[0])
            ms_, _ = timed(lambda: L.slas_cuda_sdot_sharded(local, a.data_ptr(), b.data_ptr(), res.ctypes.data), 5, 2)
            got = gather_f64([float(res[0])])
            assert all(np.array_equal(got[0], x) for x in got), f"sharded dot ({how}) differs between ranks: {got}"
            want32 = np.float32(sum_parts)
            if how != "nccl":  # rank-ordered sum inside the kernel: exactly the cast of the f64 sum of the partials
                assert res[0].tobytes() == want32.tobytes(), f"sharded dot ({how}) {res[0]} != f32(sum of partials) {want32}"
            else:              # NCCL's order is its own: one f64 rounding per rank, i.e. at most an f32 ulp after the cast
                assert abs(float(res[0]) - float(want32)) <= float(np.spacing(want32)), f"sharded dot (nccl) {res[0]} vs {want32}"
            assert abs(float(res[0]) / total - 0.25) < 1e-3, f"sharded dot {res[0]} is not ~ 2^32/4"
            rec(f"sharded_dot_{how}", ms_, f"sharded dot f32, 2^32 elements over {world} GPU(s), exchange: {how}, scalar returned to host",
                flop=2.0 * total, bytes_=8.0 * total, ranks=world,
                extra={"result": float(res[0]), "parity": "bit-identical on all ranks; == f32(rank-ordered sum of the f64 partials)" if how != "nccl"
                       else "bit-identical on all ranks; within 1 ulp of f32(sum of the f64 partials)"})
            ms_, _ = timed(lambda: L.slas_cuda_sdot_sharded(local, a.data_ptr(), b.data_ptr(), dres.data_ptr()), 5, 2)
            assert dres[:1].cpu().numpy().tobytes() == res.tobytes(), "device-resident result differs from the host-returned one"
            rec(f"sharded_dot_{how}_hbm", ms_, f"sharded dot f32, 2^32 elements over {world} GPU(s), exchange: {how}, result left in HBM",
                flop=2.0 * total, bytes_=8.0 * total, ranks=world)
            # normalize: the global norm first (its bits are what every element is divided by), then sampled windows
            L.slas_cuda_snrm2_sharded(local, a.data_ptr(), res.ctypes.data)
            nrm = np.float32(res[0])
            gn = gather_f64([float(nrm)])
            assert all(np.array_equal(gn[0], x) for x in gn), f"sharded norm ({how}) differs between ranks"
            offs = [0, max(0, local // 3 - 7), max(0, local - 4096)]
            before = [a[o:o + 4096].cpu().numpy() for o in offs]
            ms_first, _ = timed(lambda: L.slas_cuda_snormalize_sharded(local, a.data_ptr()), 1, 0)
            for o, bf in zip(offs, before):
                assert np.array_equal(a[o:o + 4096].cpu().numpy(), bf / nrm), f"sharded normalize ({how}): window at {o} is not a[i] / norm bit for bit"
            ms_, _ = timed(lambda: L.slas_cuda_snormalize_sharded(local, a.data_ptr()), 5, 1)
            rec(f"sharded_normalize_{how}", ms_, f"sharded normalize f32, 2^32 elements over {world} GPU(s), exchange: {how}",
                bytes_=12.0 * total, ranks=world, extra={"parity": "sampled windows == a[i] / norm bit for bit; norm bit-identical on all ranks"})
            L.slas_cuda_snrm2_sharded(local, a.data_ptr(), res.ctypes.data)
            assert abs(float(res[0]) - 1.0) < 1e-5, f"norm after sharded normalize is {res[0]}"
            if world > 1:
                sharding.destroy_comm(dist)
        del dres, a, b
        torch.cuda.empty_cache()
        # config 4: single 4096^2 f32 product, FFMA path and tcgen05 3xTF32 path
        if rank == 0:
            n_ = 4096
            a, b = dev_uniform(n_ * n_, torch.float32, 6), dev_uniform(n_ * n_, torch.float32, 7)
            c = torch.empty_like(a)
            va, vb, vc = vec(a), vec(b), vec(c)
            flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)

            def gemm():
                cuda.matrix_mul(va, vb, vc, n_, n_, n_, n_, n_, n_, False, False)

            def time_gemm():
                for _ in range(2):
                    gemm()
                torch.cuda.synchronize()
                ts = []
                n_before = sl.launch_count()
                with ClockSampler(local_rank) as ck:
                    for _ in range(7):
                        flush.zero_()  # > L2: evict the operands between timed iterations
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(stream); gemm(); e1.record(stream)
                        torch.cuda.synchronize()
                        ts.append(e0.elapsed_time(e1))
                return float(np.median(ts)), (sl.launch_count() - n_before) / 7, ck.summary()["sm_mhz"]

            def freivalds():
                """C x = A (B x) for a random x, in f64 on the harness side: the whole product, O(n^2)"""
                x = torch.rand(n_, device=dev, dtype=torch.float64) - 0.5
                lhs = c.view(n_, n_).double() @ x
                rhs = a.view(n_, n_).double() @ (b.view(n_, n_).double() @ x)
                mag = a.view(n_, n_).double().abs() @ (b.view(n_, n_).double().abs() @ x.abs())
                return float(((lhs - rhs).abs() / (1e-5 * np.sqrt(n_) * mag)).max().item())

            sl.set_sgemm_path("ffma")
            ms_, lpc, mhz = time_gemm()
            rec("sgemm4096_ffma", ms_, "single 4096^2 f32 matrix_mul, FFMA path (every C element an ascending-k fp32 chain)",
                flop=2.0 * n_ ** 3, bound="fp32", extra={"launches_per_call": lpc, "sm_mhz": mhz, "parity_error_over_bar": freivalds()})
            try:
                sl.set_sgemm_path("tcgen05")
                ms_, lpc, mhz = time_gemm()
                rec("sgemm4096_tcgen05", ms_, "single 4096^2 f32 matrix_mul, tcgen05 3xTF32 path (the default at this size), split pass included",
                    flop=2.0 * n_ ** 3, bound="tensor", extra={"launches_per_call": lpc, "sm_mhz": mhz, "parity_error_over_bar": freivalds()})
            finally:
                sl.set_sgemm_path("auto")
            del a, b, c, va, vb, vc, flush
        # ONE process driving all N GPUs (what a slas program is: a backend is a ZST made by B::default(), no launcher):
        # rank 0 alone, the other ranks wait at the barrier below with their HBM released
        if world > 1:
            torch.cuda.empty_cache()
            barrier()
            dist.barrier(group=host_group)
            if rank == 0 and torch.cuda.device_count() < world:
                configs["inproc_skipped"] = {"what": f"one-process rows skipped: this rank sees {torch.cuda.device_count()} device(s), not {world}"}
            elif rank == 0:
                from slas_b200.backends import init_devices
                ndev = init_devices(world)
                total = 1 << 32
                sa, sb = [], []
                for r in range(ndev):
                    lo_, hi_ = sharding.shard_range(total, 4, ndev, r)
                    va_, vb_ = CudaVec(hi_ - lo_, np.float32, device=r), CudaVec(hi_ - lo_, np.float32, device=r)
                    _lib.call("slas_cuda_sfill_uniform", hi_ - lo_, va_.as_ptr(), 4, lo_, 0.0, 1.0)
                    _lib.call("slas_cuda_sfill_uniform", hi_ - lo_, vb_.as_ptr(), 5, lo_, 0.0, 1.0)
                    sa.append(va_)
                    sb.append(vb_)
                part = np.zeros(1, np.float64)
                sum_parts = 0.0
                for r in range(ndev):  # slice order, like the exchange inside the kernels
                    _lib.call("slas_cuda_sdot_partial", sa[r].LEN, sa[r].as_ptr(), sb[r].as_ptr(), part.ctypes.data)
                    sum_parts += float(part[0])

                def wall(fn, reps=5):
                    fn()
                    ts = []
                    for _ in range(reps):
                        t0 = time.perf_counter()
                        fn()
                        ts.append((time.perf_counter() - t0) * 1e3)
                    return float(np.median(ts))

                got = cuda.dot_multi(sa, sb)
                assert np.float32(got).tobytes() == np.float32(sum_parts).tobytes(), f"in-process sharded dot {got} != f32(sum of partials) {sum_parts}"
                assert float(got) == configs["sharded_dot_p2p"]["result"], "one process / N processes disagree on the 2^32-element dot"
                ms_ = wall(lambda: cuda.dot_multi(sa, sb))
                rec("inproc_sharded_dot", ms_, f"ONE process, {ndev} GPUs (slas_cuda_init_devices): dot f32 over 2^32 elements, one device-resident slice per "
                    f"GPU, exchange inside the kernels over peer memory, scalar returned to the host (wall clock of the call)",
                    flop=2.0 * total, bytes_=8.0 * total, ranks=ndev,
                    extra={"result": float(got), "parity": "== f32(slice-ordered sum of the f64 partials) == the N-process result"})
                def normalize_all():
                    cuda.normalize_multi(sa)
                    for r_ in range(ndev):
                        torch.cuda.synchronize(r_)

                ms_ = wall(normalize_all)
                rec("inproc_sharded_normalize", ms_, f"ONE process, {ndev} GPUs: normalize f32 over 2^32 elements (wall clock incl. a final synchronize per device)",
                    bytes_=12.0 * total, ranks=ndev)
                assert abs(float(cuda.norm_multi(sa)) - 1.0) < 1e-5
                del sa, sb
                # host operands (pinned) split over the devices by batch index: N PCIe links from one process
                hb = 1 << 24
                hp = [C.c_void_p() for _ in range(3)]
                for p_ in hp:
                    _lib.call("slas_cuda_malloc_host", C.byref(p_), hb * 64)
                hA, hB, hC = [np.ctypeslib.as_array((C.c_float * (hb * 16)).from_address(p_.value)) for p_ in hp]
                rng_ = np.random.default_rng(3)
                hA[:] = rng_.random(hb * 16, dtype=np.float32)
                hB[:] = rng_.random(hb * 16, dtype=np.float32)
                ms_ = wall(lambda: cuda.matrix_mul_batched(hA, hB, hC, 4, 4, 4), reps=3)
                want = oracle.gemm_batched(hA[-4096 * 16:], hB[-4096 * 16:], 4096, 4, 4, 4)
                assert np.max(np.abs(hC[-4096 * 16:] - want)) <= 1e-5 * 2 * 4, "in-process split of the host batch differs from the oracle"
                configs["inproc_mm4_host_e2e"] = {
                    "what": f"ONE process, {ndev} GPUs: batched 4x4 f32 matrix_mul, 2^24 pairs in pinned HOST memory split over the devices by "
                            f"batch index (per-device worker thread + 3-slot copy pipeline), wall clock of the call",
                    "ms": ms_, "bound": "pcie / host memory", "achieved": 3.0 * hb * 64 / (ms_ * 1e-3) / 1e9, "unit": "GB/s over all links",
                    "gflops": FLOP_PER_4X4 * hb / (ms_ * 1e-3) / 1e9,
                    "n_process_e2e_ms_for_2^24_pairs_per_gpu": e2e["ms_per_step"] if e2e else None}
                for p_ in hp:
                    _lib.call("slas_cuda_free_host", p_.value)
            dist.barrier(group=host_group)  # CPU-side wait: no NCCL kernel spins on the GPUs rank 0 is driving

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_reference_4x4(min(args.batch_log2, CPU_SAMPLE_LOG2), 5, 1)
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
            "config": workload_config(args.batch_log2, world),
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": peak_gbs, "unit": "GB/s", "frac": achieved_gbs / peak_gbs,
                         "traffic": (traffic or {}).get("bytes") if batch == 1 << 24 else None,
                         "traffic_source": (traffic or {}).get("source"),
                         "peak_source": peak_src, "kernel": "gemm_small_kernel<float, 4x4x4>",
                         "algorithmic_bytes_per_launch": BYTES_PER_4X4 * batch,
                         "parity_error_over_bar": headline_parity,
                         "fma_peak_tflops_measured": fma_peak, "tf32_peak_measured": tf32_peak,
                         "configs": configs},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def bind_to_gpu_local_cpus(device_index: int) -> str:
    """One process per GPU: run the rank -- and with it the first touch of its pinned host buffers and the library's copy
    threads -- on the CPUs NVML reports as local to the GPU, as `numactl --cpunodebind` would.  With several ranks the
    host-to-device leg is bound by the host's memory system (bench: e2e.copy_only), and buffers on the far socket halve it."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return f"NVML reports no narrower CPU set than the {len(allowed)} allowed CPUs"
        os.sched_setaffinity(0, cpus)
        return f"bound to the {len(cpus)} CPUs local to GPU {device_index} (NVML)"
    except Exception as e:  # noqa: BLE001 -- a box without NVML topology information: run unbound
        return f"not bound ({type(e).__name__}: {e})"


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
