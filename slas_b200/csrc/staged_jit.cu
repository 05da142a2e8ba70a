// staged_jit.cu -- run-time specialisation (NVRTC, jit.cuh) of the staged small-object kernels of
// staged_objects_kernel.cuh for object shapes without a built-in instantiation: batched dot / norm / normalize of
// LEN-vectors and batched transposes of rows x columns objects with the shape as template constants, which is what
// rustc's monomorphisation gives the reference (StaticVec<T, LEN>, static_vec.rs:76; MatrixShape<M, K>,
// tensor.rs:108-115).  A shape costs ~1 s the first time it is seen; the CUfunction is cached per device.  Returns
// nullptr when NVRTC is not available or the compile failed: the caller launches its run-time-shape kernel instead.
#include <map>
#include <mutex>
#include <tuple>

#include "jit.cuh"

namespace slas {

static const char *kStagedSource =
#include "staged_jit_src.inc"
    ;

// device, kernel family (0 reduce, 1 transpose), element size, mode / per-object flag, len, rows, columns
using StagedKey = std::tuple<int, int, size_t, int, unsigned, unsigned, unsigned>;
static std::map<StagedKey, CUfunction> g_staged_cache;  // nullptr = compile failed, do not retry
static std::mutex g_staged_mu;

static CUfunction staged_lookup(const StagedKey &key, const char *expr) {
    if (tunable("staged_jit") == 1 || !jit_ready()) return nullptr;
    std::lock_guard<std::mutex> lk(g_staged_mu);
    auto it = g_staged_cache.find(key);
    if (it == g_staged_cache.end()) {
        CUfunction fn = nullptr;
        if (!jit_compile(kStagedSource, "staged_objects_jit.cu", expr, kStagedJitSmem, &fn)) fn = nullptr;
        it = g_staged_cache.emplace(key, fn).first;
    }
    return it->second;
}

CUfunction staged_reduce_jit(Context *ctx, size_t elem_size, int mode, unsigned len) {
    char expr[160];
    snprintf(expr, sizeof(expr), "slas::batched_reduce_staged_kernel<%s, %d, %uu>", elem_size == 4 ? "float" : "double", mode, len);
    return staged_lookup(StagedKey{ctx->device, 0, elem_size, mode, len, 0u, 0u}, expr);
}

CUfunction staged_transpose_jit(Context *ctx, size_t elem_size, unsigned len, unsigned rows, unsigned columns, bool per_object) {
    const char *tname = elem_size == 1 ? "uint8_t" : elem_size == 2 ? "uint16_t" : elem_size == 4 ? "uint32_t" : elem_size == 8 ? "uint64_t" : "uint4";
    char expr[200];
    snprintf(expr, sizeof(expr), "slas::transpose_staged_kernel<%s, %uu, %uu, %uu, %s>", tname, len, rows, columns,
             per_object ? "true" : "false");
    return staged_lookup(StagedKey{ctx->device, 1, elem_size, per_object ? 1 : 0, len, rows, columns}, expr);
}

}  // namespace slas
