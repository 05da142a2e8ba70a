"""Host-side logic of the multi-GPU path: one process per GPU, each rank owns one
contiguous slice.  Batched ops shard by batch index and need no collective; a single huge
dot / norm / normalize shards by vector slice and needs exactly one sum-allreduce of an f64 partial.

`torch.distributed` is plumbing only: rendezvous and the broadcast of the NCCL unique id.  The
collective on the data path is the library's own ncclAllReduce (slas_b200/csrc/comm.cu), enqueued on
the stream right behind the kernel that produced the partial."""
from __future__ import annotations

import ctypes as C

from . import _lib


def shard_range(length: int, elem_size: int, nranks: int, rank: int) -> tuple[int, int]:
    """[begin, end) of a `length`-element vector owned by `rank`; interior boundaries fall on 16-byte
    multiples so that every slice keeps the 128-bit kernels.  Pure host arithmetic (no GPU needed)."""
    lo, hi = C.c_size_t(), C.c_size_t()
    _lib.call("slas_cuda_shard_range", length, elem_size, nranks, rank, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def batch_range(batch: int, nranks: int, rank: int) -> tuple[int, int]:
    """[begin, end) of the batch indices owned by `rank` (what CudaBatch.shard uses)."""
    return batch * rank // nranks, batch * (rank + 1) // nranks


def init_comm(dist, fused: bool = True) -> str:
    """Set up the collective of the sharded ops over the ranks of an initialised torch.distributed
    process group (any backend: it only carries a few bytes of handles).

    fused=True (default, <= 8 ranks of one NVSwitch domain): every rank exports a mailbox through CUDA
    IPC and maps its peers' -- the reduction kernel then does the exchange itself over NVLink peer
    memory (slas_b200/csrc/p2p.cuh).  fused=False: the library's NCCL communicator, one
    ncclAllReduce of the f64 partial behind the kernel.  Returns "p2p", "nccl" or "single"."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return "single"
    if fused and world <= 8:
        buf = (C.c_char * 64)()
        _lib.call("slas_cuda_p2p_handle", buf)
        handles = [None] * world
        dist.all_gather_object(handles, bytes(buf.raw))
        _lib.call("slas_cuda_p2p_init", world, rank, b"".join(handles))
        dist.barrier()
        return "p2p"
    ident = [None]
    if rank == 0:
        buf = (C.c_char * 128)()
        _lib.call("slas_cuda_nccl_unique_id", buf)
        ident = [bytes(buf.raw)]
    dist.broadcast_object_list(ident, src=0)
    _lib.call("slas_cuda_comm_init", world, rank, ident[0])
    return "nccl"


def destroy_comm(dist=None) -> None:
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()  # nobody unmaps a mailbox a peer may still write to
    _lib.call("slas_cuda_p2p_destroy")
    _lib.call("slas_cuda_comm_destroy")


def combine_partials(partial: float, dist, mode: str = "dot"):
    """The reduction protocol of the sharded ops, spelt on the host: sum the ranks' f64 partials, then
    finish (dot: the sum; norm: its square root).  The GPU path does the same on device; this is the
    form the CPU (gloo) tests exercise."""
    import torch
    t = torch.tensor([partial], dtype=torch.float64)
    if dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    s = float(t.item())
    return s ** 0.5 if mode == "norm" else s
