"""Worker of tests/test_cuda_multi_gpu.py: one rank per GPU under torchrun.  Checks the sharded
dot / norm / normalize (fused NVLink peer exchange and the NCCL path) against an f64 host truth of the
whole vector, that every rank ends with the same bits, and batch-index sharding of the batched GEMM."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402
from slas_b200 import sharding  # noqa: E402
from slas_b200.backends import Cuda  # noqa: E402
from slas_b200.static_vec import CudaBatch, CudaVec  # noqa: E402


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
    for fused in (True, False):
        mode = sharding.init_comm(dist, fused=fused)
        assert mode == ("p2p" if fused else "nccl"), mode
        da, db = CudaVec.from_host(a[lo:hi]), CudaVec.from_host(b[lo:hi])
        for rep in range(5):  # repeated collectives walk both mailbox parities
            d = cuda.dot_sharded(da, db)
            assert abs(float(d) - truth_dot) <= 1e-5 * np.sqrt(n) * scale, (mode, float(d), truth_dot)
            nr = cuda.norm_sharded(da)
            assert abs(float(nr) - truth_nrm) <= 1e-5 * np.sqrt(n) * truth_nrm
            # bit-identical on every rank
            t = torch.tensor([float(d), float(nr)], dtype=torch.float64, device="cuda")
            g = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(g, t)
            assert all(torch.equal(g[0], x) for x in g), (mode, [x.tolist() for x in g])
        cuda.normalize_sharded(da)
        assert np.array_equal(da.to_host(), a[lo:hi] / nr)          # true division by the global norm
        assert abs(float(cuda.norm_sharded(da)) - 1.0) < 1e-5
        # an empty slice on some ranks still takes part in the collective
        tiny = CudaVec.from_host(np.ones(4, np.float32)) if rank == 0 else CudaVec(0, np.float32)
        assert float(cuda.dot_sharded(tiny, tiny)) == 4.0
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
        print("MULTI_GPU_OK", world)


if __name__ == "__main__":
    main()
