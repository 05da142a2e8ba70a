// jit.cuh -- run-time specialisation of kernel templates with NVRTC (gemm_jit.cu holds the machinery; gemm_jit.cu and
// staged_jit.cu are its users).  slas shapes are compile-time constants that rustc monomorphises; the C ABI receives
// them as run-time integers, so the library compiles the shapes it has no built-in instantiation for the first time they
// are seen.  NVRTC and the driver API are bound at run time (dlopen / cudaGetDriverEntryPoint): without them every user
// falls back to its run-time-shape kernel -- never to the CPU.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace slas {

// true once libnvrtc and the driver entry points are bound (tried once per process)
bool jit_ready();
// Compiles `source` for sm_100a, loads the cubin into the CURRENT device's context and returns the __global__ function
// that the template-instantiation expression `expr` names, with `smem` bytes of dynamic shared memory allowed.  On failure
// returns false and leaves the NVRTC log in slas_cuda_last_error().
bool jit_compile(const char *source, const char *file_name, const char *expr, size_t smem, CUfunction *fn);
int jit_launch(CUfunction fn, unsigned grid, unsigned block, size_t smem, cudaStream_t st, void **params);

// staged_jit.cu -- staged_objects_kernel.cuh with the object shape as template constants; nullptr = not available (no
// NVRTC, compile failure, tunable staged_jit = 1): launch the run-time-shape instantiation.  Every function is allowed
// kStagedJitSmem bytes of dynamic shared memory.
constexpr size_t kStagedJitSmem = 200 * 1024;
CUfunction staged_reduce_jit(Context *ctx, size_t elem_size, int mode, unsigned len);
CUfunction staged_transpose_jit(Context *ctx, size_t elem_size, unsigned len, unsigned rows, unsigned columns, bool per_object);

}  // namespace slas
