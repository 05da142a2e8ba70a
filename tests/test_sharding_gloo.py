"""N > 1 host logic on CPU: two processes over `gloo` walk the same protocol the GPU ranks do --
slice the vector / the batch with the library's own shard arithmetic, form a rank-local f64 partial,
sum-allreduce it, finish -- and must reproduce the single-process result.  The rank-local partial is
computed with numpy here (no GPU in this container); on the B200 box the same slices go through
slas_cuda_{s,d}dot_sharded and the allreduce is NCCL (bench.py, config 5)."""
import os
import socket

import numpy as np
import pytest

from slas_b200 import sharding


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, length: int, out_dir: str):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)  # every rank regenerates the same global vectors, then keeps its slice
        a = rng.random(length, dtype=np.float32)
        b = (rng.random(length, dtype=np.float32) * 2 - 1)
        lo, hi = sharding.shard_range(length, 4, world, rank)
        assert (lo * 4) % 16 == 0
        partial = float(np.dot(a[lo:hi].astype(np.float64), b[lo:hi].astype(np.float64)))
        dot = sharding.combine_partials(partial, dist, "dot")
        sq = float(np.dot(a[lo:hi].astype(np.float64), a[lo:hi].astype(np.float64)))
        nrm = sharding.combine_partials(sq, dist, "norm")
        # normalize: the local scale pass needs no second collective
        local_normalized = a[lo:hi] / np.float32(nrm)
        # batched ops: batch-index slices, no collective at all
        blo, bhi = sharding.batch_range(1000 + 3, world, rank)
        np.save(os.path.join(out_dir, f"r{rank}.npy"),
                np.array([dot, nrm, lo, hi, blo, bhi, float(np.sum(local_normalized.astype(np.float64) ** 2))]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("length", [1 << 16, (1 << 16) + 5, 7])
def test_sharded_dot_norm_protocol_world2(tmp_path, length):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, length, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"r{i}.npy") for i in range(world)]
    rng = np.random.default_rng(7)
    a = rng.random(length, dtype=np.float32)
    b = (rng.random(length, dtype=np.float32) * 2 - 1)
    want_dot = float(np.dot(a.astype(np.float64), b.astype(np.float64)))
    want_nrm = float(np.sqrt(np.dot(a.astype(np.float64), a.astype(np.float64))))
    for i in range(world):
        assert r[i][0] == pytest.approx(want_dot, rel=1e-12, abs=1e-12)  # every rank holds the full result
        assert r[i][1] == pytest.approx(want_nrm, rel=1e-12)
    assert r[0][0] == r[1][0] and r[0][1] == r[1][1]  # bit-identical on all ranks after the allreduce
    # slices tile the vector and the batch exactly
    assert r[0][2] == 0 and r[0][3] == r[1][2] and r[1][3] == length
    assert r[0][4] == 0 and r[0][5] == r[1][4] and r[1][5] == 1003
    # the normalized slices together have unit norm
    assert r[0][6] + r[1][6] == pytest.approx(1.0, rel=1e-5)


def test_batch_shards_cover_every_object():
    from slas_b200.static_vec import CudaBatch
    for batch in (1, 7, 1 << 20):
        for n in (1, 2, 4, 8):
            whole = CudaBatch(batch, 16, np.float32, _ptr=1 << 30)  # a view: no allocation, no GPU
            seen = 0
            for rnk in range(n):
                s = whole.shard(n, rnk)
                assert s.as_ptr() == (1 << 30) + seen * 16 * 4
                seen += s.batch
            assert seen == batch
