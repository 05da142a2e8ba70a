"""Worker of tests/test_cuda_multi_gpu.py: one rank per GPU under torchrun.  Checks the sharded
dot / norm / normalize (fused NVLink peer exchange and the NCCL path) against an f64 host truth of the
whole vector, against the rank-ordered sum of the ranks' own f64 partials, that every rank ends with the same
bits, a back-to-back stress of the mailbox / epoch protocol with alternating lengths, and batch-index sharding
of the batched GEMM.  Prints MULTI_GPU_OK <world> on rank 0 when every rank passed."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402
from slas_b200 import _lib as L  # noqa: E402
from slas_b200 import sharding  # noqa: E402
from slas_b200.backends import Cuda  # noqa: E402
from slas_b200.static_vec import CudaBatch, CudaVec  # noqa: E402


def gather_f64(vals, world):
    t = torch.tensor(vals, dtype=torch.float64, device="cuda")
    g = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(g, t)
    return [x.cpu().numpy() for x in g]


def partial_f64(v: CudaVec, w: CudaVec) -> float:
    out = np.zeros(1, np.float64)
    name = "slas_cuda_sdot_partial" if v.dtype == np.float32 else "slas_cuda_ddot_partial"
    L.call(name, v.LEN, v.as_ptr(), w.as_ptr(), out.ctypes.data)
    return float(out[0])


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    cuda = Cuda()
    n = (1 << 22) + 12
    rng = np.random.default_rng(3)  # every rank regenerates the same global vectors and keeps its slice
    a = rng.random(n, dtype=np.float32)
    b = rng.random(n, dtype=np.float32) * 2 - 1
    lo, hi = sharding.shard_range(n, 4, world, rank)
    truth_dot = float(np.dot(a.astype(np.float64), b.astype(np.float64)))
    truth_nrm = float(np.sqrt(np.dot(a.astype(np.float64), a.astype(np.float64))))
    scale = float(np.sum(np.abs(a.astype(np.float64) * b)))
    worst = {}
    for fused in (True, False):
        mode = sharding.init_comm(dist, fused=fused)
        assert mode == ("p2p" if fused else "nccl"), mode
        da, db = CudaVec.from_host(a[lo:hi]), CudaVec.from_host(b[lo:hi])
        # the collective's own inputs: the ranks' f64 partials, added here in rank order
        parts = gather_f64([partial_f64(da, db), partial_f64(da, da)], world)
        sum_dot = 0.0
        sum_sq = 0.0
        for p in parts:  # rank order, like the fused exchange
            sum_dot += float(p[0])
            sum_sq += float(p[1])
        for rep in range(5):  # repeated collectives walk both mailbox parities
            d = cuda.dot_sharded(da, db)
            assert abs(float(d) - truth_dot) <= 1e-5 * np.sqrt(n) * scale, (mode, float(d), truth_dot)
            worst[mode + "_dot"] = max(worst.get(mode + "_dot", 0.0), abs(float(d) - truth_dot) / (1e-5 * np.sqrt(n) * scale))
            nr = cuda.norm_sharded(da)
            assert abs(float(nr) - truth_nrm) <= 1e-5 * np.sqrt(n) * truth_nrm
            # the result is the typed cast of the rank-ordered f64 sum: exact for the fused exchange, and within one
            # f64 rounding per rank (1e-12 relative) of it whatever order NCCL adds in
            if fused:
                assert np.float32(sum_dot).tobytes() == np.float32(d).tobytes(), (mode, float(d), sum_dot)
                assert np.float32(np.sqrt(sum_sq)).tobytes() == np.float32(nr).tobytes(), (mode, float(nr), np.sqrt(sum_sq))
            else:
                assert abs(float(d) - np.float32(sum_dot)) <= np.spacing(np.float32(abs(sum_dot))), (mode, float(d), sum_dot)
                assert abs(float(nr) - np.float32(np.sqrt(sum_sq))) <= np.spacing(np.float32(np.sqrt(sum_sq)))
            # bit-identical on every rank
            g = gather_f64([float(d), float(nr)], world)
            assert all(np.array_equal(g[0], x) for x in g), (mode, [x.tolist() for x in g])
        # f64 data: the f64 result against the sum of the partials to 1e-12 relative (exact when fused)
        da64, db64 = CudaVec.from_host(a[lo:hi].astype(np.float64)), CudaVec.from_host(b[lo:hi].astype(np.float64))
        p64 = gather_f64([partial_f64(da64, db64)], world)
        s64 = 0.0
        for p in p64:
            s64 += float(p[0])
        d64 = float(cuda.dot_sharded(da64, db64))
        if fused:
            assert d64 == s64, (mode, d64, s64)
        assert abs(d64 - s64) <= 1e-12 * scale and abs(d64 - truth_dot) <= 1e-13 * np.sqrt(n) * scale
        cuda.normalize_sharded(da)
        assert np.array_equal(da.to_host(), a[lo:hi] / nr)          # true division by the global norm
        assert abs(float(cuda.norm_sharded(da)) - 1.0) < 1e-5
        # an empty slice on some ranks still takes part in the collective
        tiny = CudaVec.from_host(np.ones(4, np.float32)) if rank == 0 else CudaVec(0, np.float32)
        assert float(cuda.dot_sharded(tiny, tiny)) == 4.0
        # stress (racecheck stand-in for the ticket / flagged-slot / mailbox code): many back-to-back collectives
        # with alternating lengths, device-resident results (no host sync between them), both epoch parities
        lens = [hi - lo, 1, 4096, (hi - lo) // 3, 0, 1 << 16, 17]
        views = [(da.view(0, ln), db.view(0, ln)) for ln in lens]
        res = CudaVec(len(lens) * 32, np.float32).zero_()
        reps = 32
        for r in range(reps):
            for i, (va, vb) in enumerate(views):
                L.call("slas_cuda_sdot_sharded", va.LEN, va.as_ptr(), vb.as_ptr(), res.as_ptr() + 4 * (i * reps + r))
        got = res.to_host().reshape(len(lens), reps)
        for i in range(len(lens)):
            assert np.all(got[i] == got[i][0]), (mode, "stress: a repeated collective changed its result", lens[i])
        g = gather_f64(got[:, 0].astype(np.float64).tolist(), world)
        assert all(np.array_equal(g[0], x) for x in g), (mode, "stress: ranks disagree")
        sharding.destroy_comm(dist)
    # batched ops shard by batch index with no collective
    batch, size = 1000 + 7, 16
    A = rng.random(batch * size * size, dtype=np.float32)
    B = rng.random(batch * size * size, dtype=np.float32)
    blo, bhi = sharding.batch_range(batch, world, rank)
    e = size * size
    c = CudaBatch(bhi - blo, e, np.float32)
    cuda.matrix_mul_batched(CudaVec.from_host(A[blo * e:bhi * e]), CudaVec.from_host(B[blo * e:bhi * e]), c, size, size, size)
    want = oracle.gemm_batched(A[blo * e:bhi * e], B[blo * e:bhi * e], bhi - blo, size, size, size)
    assert np.max(np.abs(c.to_host().ravel() - want)) <= 1e-4
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_OK", world, {k: round(v, 6) for k, v in worst.items()}, "(observed error / bar)")


if __name__ == "__main__":
    main()
