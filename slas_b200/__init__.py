"""slas_b200 -- `slas_backend::Cuda`: a B200 (sm_100a) compute backend behind the backend-operation
traits of unic0rn9k/slas.  Kernels + C ABI live in slas_b200/csrc (-> slas_b200/lib/libslas_cuda.so,
declared in include/slas_cuda.h); this package is the thin host-side mirror of the reference
interface for that path (backends.Cuda, static_vec.CudaVec/CudaBatch, tensor.Matrix)."""
from . import _lib
from .backends import (Cuda, Graph, WithStaticBackend, device_count, device_info, launch_count, set_sgemm_path, set_stream,
                       synchronize)
from .static_vec import CudaBatch, CudaVec, StaticCowVec, moo
from .tensor import BatchedMatrix, Matrix, MatrixShape, Tensor, matrix, reshape, tensor_index

slas_backend = type("slas_backend", (), {"Cuda": Cuda})

__all__ = ["Cuda", "Graph", "WithStaticBackend", "CudaVec", "CudaBatch", "StaticCowVec", "moo", "Matrix", "BatchedMatrix", "MatrixShape",
           "Tensor", "matrix", "reshape", "tensor_index", "device_count", "device_info", "launch_count",
           "set_sgemm_path", "set_stream", "synchronize", "slas_backend"]
