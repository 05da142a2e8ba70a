// gemm_jit.cu -- run-time specialisation of the batched static-shape GEMM kernel for shapes that have no
// ahead-of-time instantiation.
//
// In slas a matrix shape is a compile-time constant (MatrixShape<M, K>, src/tensor.rs:108-115): rustc monomorphises
// Matrix::matrix_mul for every shape a program uses.  The C ABI receives the shape as run-time integers, and the
// library can only carry so many instantiations ({4, 8, 16, 32, 64}^3 and everything up to 4 x 4 x 4).  For the rest
// (8x16 . 16x4, 12x12 . 12x12, 6x6 . 6x6, ...) this file does what rustc does: it compiles
// gemm_small_kernel<T, SmallCfg<T, M, N, K, ...>, TA, TB> -- the very kernel text of gemm_small_kernel.cuh, embedded
// at build time -- for sm_100a with NVRTC the first time a shape is seen (~1 s), keeps the CUfunction in a cache and
// launches it like the built-in ones from then on.  NVRTC and the driver API are bound at run time (dlopen /
// cudaGetDriverEntryPoint); if either is missing, or a shape does not fit the kernel's tiling rules, the caller falls
// back to the runtime-shape kernel (SLAS_ERR_UNSUPPORTED) -- never to the CPU.
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <stdio.h>

#include <map>
#include <mutex>
#include <string>
#include <tuple>

#include "common.cuh"
#include "jit.cuh"
#include "kernels.cuh"

namespace slas {

static const char *kKernelSource =
#include "gemm_small_jit_src.inc"
    ;

// ---- NVRTC + driver entry points, bound once -------------------------------------------------------------------
struct JitApi {
    bool ok = false;
    nvrtcResult (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *);
    nvrtcResult (*DestroyProgram)(nvrtcProgram *);
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char *const *);
    nvrtcResult (*AddNameExpression)(nvrtcProgram, const char *);
    nvrtcResult (*GetLoweredName)(nvrtcProgram, const char *, const char **);
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t *);
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char *);
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t *);
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char *);
    CUresult (*ModuleLoadData)(CUmodule *, const void *);
    CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *);
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int);
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **,
                             void **);
};
static JitApi g_api;
static std::once_flag g_api_once;

template <typename F> static bool driver_fn(const char *name, F *out) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        return false;
    }
    *out = reinterpret_cast<F>(fn);
    return true;
}

static void bind_api() {
    void *h = nullptr;
    for (const char *name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
        h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        if (h) break;
    }
    if (!h) return;
    bool ok = true;
#define SLAS_NVRTC_SYM(field, sym) ok = ok && ((g_api.field = reinterpret_cast<decltype(g_api.field)>(dlsym(h, sym))) != nullptr)
    SLAS_NVRTC_SYM(CreateProgram, "nvrtcCreateProgram");
    SLAS_NVRTC_SYM(DestroyProgram, "nvrtcDestroyProgram");
    SLAS_NVRTC_SYM(CompileProgram, "nvrtcCompileProgram");
    SLAS_NVRTC_SYM(AddNameExpression, "nvrtcAddNameExpression");
    SLAS_NVRTC_SYM(GetLoweredName, "nvrtcGetLoweredName");
    SLAS_NVRTC_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
    SLAS_NVRTC_SYM(GetCUBIN, "nvrtcGetCUBIN");
    SLAS_NVRTC_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
    SLAS_NVRTC_SYM(GetProgramLog, "nvrtcGetProgramLog");
#undef SLAS_NVRTC_SYM
    ok = ok && driver_fn("cuModuleLoadData", &g_api.ModuleLoadData) && driver_fn("cuModuleGetFunction", &g_api.ModuleGetFunction) &&
         driver_fn("cuFuncSetAttribute", &g_api.FuncSetAttribute) && driver_fn("cuLaunchKernel", &g_api.LaunchKernel);
    g_api.ok = ok;
}

static size_t gcd_sz(size_t a, size_t b) { return b ? gcd_sz(b, a % b) : a; }

// ---- tile configuration for an arbitrary (M, N, K) ---------------------------------------------------------------
struct JitCfg {
    int tiny = 0;  // 0: gemm_small_kernel (vector register tiles), 1: gemm_tiny_kernel (a thread per matrix),
                   // 2: gemm_rect_kernel (scalar register tiles)
    int tm = 0, tn = 0, g = 0, cons = 256, stages = 3;
    size_t smem = 0, stage_elems = 0, cstage_elems = 0;
};

// Thread per matrix (gemm_tiny_kernel<T, M, N, K, G>): any m, n, k -- no vector-width rules -- as long as the operands
// and the result of one product fit a thread's registers (<= 200 32-bit registers) and three input stages plus two C
// stages of G matrices fit shared memory.  G = consumer threads: as many as leave room for two CTAs per SM.
static bool choose_tiny_cfg(size_t elem, size_t m, size_t n, size_t k, JitCfg *out) {
    const size_t words = (m * k + k * n + m * n) * (elem / 4);
    if (m == 0 || n == 0 || k == 0 || words > 200) return false;
    for (size_t g = 256; g >= 64; g /= 2) {
        const size_t stage = (g * (m * k + k * n) * elem + 127) / 128 * 128, cstage = (g * m * n * elem + 127) / 128 * 128;
        const size_t smem = 3 * stage + 2 * cstage + 6 * sizeof(uint64_t);
        if (smem > (g > 64 ? 110 : 200) * 1024) continue;
        out->tiny = 1;
        out->g = (int)g;
        out->cons = (int)g;
        out->smem = smem;
        out->stage_elems = stage / elem;
        out->cstage_elems = cstage / elem;
        return true;
    }
    return false;
}

// C through a double-buffered shared tile and one bulk store per chunk, or straight from the registers with per-warp
// 16-byte stores?  Measured both ways (profiles/r3x_tune_jitbulkc.log): the bulk store wins when per-warp stores leave
// partial lines -- rows that are not whole 128-byte lines in big objects (20x20 f32 0.88 -> 0.95, 24x24 0.85 -> 0.92,
// 20x20 f64 0.87 -> 0.97) or not whole 32-byte sectors (6x6 f64 0.86 -> 1.01) -- and loses a few per cent where the
// stores are already whole lines (16x8 . 8x32: 0.94 -> 0.91) or whole tiny objects (8x16 . 16x4: 1.04 -> 0.97).
// tunable gemm_jit_bulk_c: 1 = always, 2 = never.
static bool jit_bulk_c_store(size_t elem, size_t m, size_t n) {
    if (tunable("gemm_jit_bulk_c") == 1) return true;
    if (tunable("gemm_jit_bulk_c") == 2) return false;
    const size_t row = n * elem, obj = m * n * elem;
    return row % 128 != 0 && (obj >= 1536 || (row % 32 != 0 && row > 32));
}

// The rules of SmallCfg / gemm_small_kernel: K and TN whole 16-byte vectors, TM | M, TN | N, threads per matrix
// (M/TM)*(N/TN) dividing the consumer count (whole warps, at most 512 threads), a chunk = whole passes, the stage ring within the
// shared-memory budget.  Among the configurations that fit: big, squarish register tiles (fewer shared-memory loads
// per FMA), at least two CTAs per SM, three stages and 256 consumers when possible.
static bool choose_cfg(size_t elem, size_t m, size_t n, size_t k, JitCfg *out) {
    const size_t kvn = 16 / elem;
    if (m == 0 || n == 0 || k == 0 || n % kvn || k % kvn) return false;
    const size_t per_pair = (m * k + k * n) * elem;
    if (per_pair > 40 * 1024) return false;
    const size_t max_tile = 64;  // accumulators per thread (24^3 f32 needs 3 x 12, 20^3 f64 5 x 10)
    double best = -1.0;
    for (size_t tm = 1; tm <= m; ++tm) {
        if (m % tm) continue;
        for (size_t tn = kvn; tn <= n; tn += kvn) {
            if (n % tn || tm * tn > max_tile) continue;
            const size_t tpm = (m / tm) * (n / tn);
            if (tpm > 512) continue;
            // consumer threads: whole warps and whole matrices -- multiples of lcm(tpm, 32) up to 512 (20x20: 5 x 4 tiles,
            // 20 threads per matrix, 160 consumers; powers of two get the usual 256 / 128)
            const size_t unit = tpm / gcd_sz(tpm, 32) * 32;
            if (unit > 512) continue;
            const bool pow2 = (tpm & (tpm - 1)) == 0;
            // powers of two: 256 or 128 consumers (as measured for the built-in shapes); otherwise the largest whole
            // number of matrices within 256 threads, or one `unit` when that is already more
            const size_t cands[2] = {unit <= 256 ? 256 / unit * unit : unit, pow2 && tpm <= 128 ? (size_t)128 : 0};
            for (size_t ci = 0; ci < 2; ++ci) {
                const size_t cons = cands[ci];
                if (cons == 0 || cons < tpm) continue;
                const size_t mpp = cons / tpm;
                size_t g = (32 * 1024 / per_pair) / mpp * mpp;
                if (g < mpp) g = mpp;
                if (g > 256) g = 256 / mpp * mpp;
                if (g == 0) continue;
                const bool bulk_c = jit_bulk_c_store(elem, m, n);
                for (int stages = 3; stages >= 2; --stages) {
                    const size_t smem = (size_t)stages * g * per_pair + (bulk_c ? 2 * g * m * n * elem : 0) + 2 * stages * sizeof(uint64_t);
                    if (smem > 200 * 1024) continue;
                    const size_t per_sm = (size_t)227 * 1024 / (smem + 1024);
                    double score = 1.0 / (1.0 / (double)tn + 1.0 / (double)tm) + 1e-3 * (double)tn;
                    // K of one or two vectors: nothing to reuse, and a thread that owns most of a matrix reads its rows at
                    // the matrix stride -- lanes collide in shared memory (8x4 . 4x8 f32 as one thread per matrix: 0.53
                    // of HBM peak).  Prefer eight or more threads per matrix there, which read consecutive vectors.
                    if (k <= 2 * kvn && tpm < 8 && (m / 1) * (n / kvn) >= 8) score *= 0.25;
                    if (per_sm < 2) score *= 0.6;
                    if (stages == 2) score *= 0.9;
                    if (cons < 192) score *= 0.95;
                    if (score > best) {
                        best = score;
                        out->tm = (int)tm; out->tn = (int)tn; out->g = (int)g; out->cons = (int)cons; out->stages = stages;
                        out->smem = smem;
                    }
                }
            }
        }
    }
    return best > 0;
}

// Scalar register tiles (gemm_rect_kernel<T, M, N, K, TM, TN, G>): any divisors TM | M, TN | N with at most 256 threads per
// matrix; G = a whole number of passes of MPP = 256 / TPM matrices, sized so that three input stages and two C stages
// leave room for two CTAs per SM when possible.
static bool choose_rect_cfg(size_t elem, size_t m, size_t n, size_t k, JitCfg *out) {
    if (m == 0 || n == 0 || k == 0) return false;
    const size_t per_in = (m * k + k * n) * elem, per_out = m * n * elem;
    if (per_in > 48 * 1024) return false;
    // G objects per chunk must keep every chunk of A, B and C on a 16-byte boundary
    const size_t galign = 16 / gcd_sz(gcd_sz(gcd_sz(m * k * elem, k * n * elem), m * n * elem), 16);
    double best = -1.0;
    for (size_t tm = 1; tm <= m; ++tm) {
        if (m % tm) continue;
        for (size_t tn = 1; tn <= n; ++tn) {
            if (n % tn || tm * tn > 32) continue;
            const size_t tpm = (m / tm) * (n / tn);
            if (tpm > 256) continue;
            const size_t mpp = 256 / tpm;
            size_t g = (24 * 1024 / (per_in + per_out)) / galign * galign;
            if (g < mpp) g = (mpp + galign - 1) / galign * galign;
            if (g < galign) g = galign;
            if (g > 1024) g = 1024 / galign * galign;
            size_t stage = 0, cstage = 0, smem = 0;
            for (;; g -= galign) {  // big objects: fewer per chunk than a pass could take (some thread groups idle)
                stage = (g * per_in + 127) / 128 * 128;
                cstage = (g * per_out + 127) / 128 * 128;
                smem = 3 * stage + 2 * cstage + 6 * sizeof(uint64_t);
                if (smem <= 200 * 1024 || g <= galign) break;
            }
            if (smem > 200 * 1024) continue;
            double score = 1.0 / (1.0 / (double)tn + 1.0 / (double)tm);         // FMAs per shared-memory load
            score *= (double)((g < mpp ? g : mpp) * tpm) / 256.0;                // busy threads
            if ((size_t)227 * 1024 / (smem + 1024) < 2) score *= 0.6;
            if (score > best) {
                best = score;
                out->tiny = 2;
                out->tm = (int)tm; out->tn = (int)tn; out->g = (int)g; out->cons = 256; out->smem = smem;
                out->stage_elems = stage / elem; out->cstage_elems = cstage / elem;
            }
        }
    }
    return best > 0;
}

// ---- compile + cache ---------------------------------------------------------------------------------------------
struct JitKernel {
    CUfunction fn = nullptr;
    JitCfg cfg;
    bool failed = false;
};
using JitKey = std::tuple<int, int, size_t, size_t, size_t, int, int>;  // device (a module belongs to its context), elem size, m, n, k, ta, tb
static std::map<JitKey, JitKernel> g_cache;
static std::mutex g_cache_mu;

bool jit_ready() {
    std::call_once(g_api_once, bind_api);
    return g_api.ok;
}

bool jit_compile(const char *source, const char *file_name, const char *expr, size_t smem, CUfunction *fn) {
    if (!jit_ready()) return false;
    nvrtcProgram prog = nullptr;
    if (g_api.CreateProgram(&prog, source, file_name, 0, nullptr, nullptr) != NVRTC_SUCCESS) return false;
    bool ok = g_api.AddNameExpression(prog, expr) == NVRTC_SUCCESS;
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-default-device"};
    if (ok && g_api.CompileProgram(prog, 3, opts) != NVRTC_SUCCESS) {
        ok = false;
        size_t len = 0;
        if (g_api.GetProgramLogSize(prog, &len) == NVRTC_SUCCESS && len > 1) {
            std::string log(len, '\0');
            g_api.GetProgramLog(prog, &log[0]);
            fail(SLAS_ERR_CUDA, "NVRTC could not specialise %s: %.400s", expr, log.c_str());
        }
    }
    const char *lowered = nullptr;
    ok = ok && g_api.GetLoweredName(prog, expr, &lowered) == NVRTC_SUCCESS && lowered;
    size_t size = 0;
    ok = ok && g_api.GetCUBINSize(prog, &size) == NVRTC_SUCCESS && size > 0;
    std::string cubin;
    if (ok) {
        cubin.resize(size);
        ok = g_api.GetCUBIN(prog, &cubin[0]) == NVRTC_SUCCESS;
    }
    CUmodule mod = nullptr;
    ok = ok && g_api.ModuleLoadData(&mod, cubin.data()) == CUDA_SUCCESS;  // the module lives as long as the process
    ok = ok && g_api.ModuleGetFunction(fn, mod, lowered) == CUDA_SUCCESS;
    ok = ok && g_api.FuncSetAttribute(*fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem) == CUDA_SUCCESS;
    g_api.DestroyProgram(&prog);
    return ok;
}

int jit_launch(CUfunction fn, unsigned grid, unsigned block, size_t smem, cudaStream_t st, void **params) {
    const CUresult r = g_api.LaunchKernel(fn, grid, 1, 1, block, 1, 1, (unsigned)smem, (CUstream)st, params, nullptr);
    if (r != CUDA_SUCCESS) return fail(SLAS_ERR_CUDA, "launch of a run-time specialised kernel failed (CUresult %d)", (int)r);
    count_launch();
    return SLAS_OK;
}

static bool compile_kernel(size_t elem, size_t m, size_t n, size_t k, int ta, int tb, JitKernel *kern) {
    const char *tname = elem == 4 ? "float" : "double";
    char expr[256];
    if (kern->cfg.tiny == 1)
        snprintf(expr, sizeof(expr), "slas::gemm_tiny_kernel<%s, %zu, %zu, %zu, %d>", tname, m, n, k, kern->cfg.g);
    else if (kern->cfg.tiny == 2)
        snprintf(expr, sizeof(expr), "slas::gemm_rect_kernel<%s, %zu, %zu, %zu, %d, %d, %d>", tname, m, n, k, kern->cfg.tm, kern->cfg.tn,
                 kern->cfg.g);
    else
        snprintf(expr, sizeof(expr), "slas::gemm_small_kernel<%s, slas::SmallCfg<%s, %zu, %zu, %zu, %d, %d, %d, %d, %d, %s>, %s, %s>",
                 tname, tname, m, n, k, kern->cfg.tm, kern->cfg.tn, kern->cfg.cons, kern->cfg.g, kern->cfg.stages,
                 jit_bulk_c_store(elem, m, n) ? "true" : "false", ta ? "true" : "false", tb ? "true" : "false");
    return jit_compile(kKernelSource, "gemm_small_jit.cu", expr, kern->cfg.smem, &kern->fn);
}

template <typename T>
int launch_gemm_jit_batched(Context *ctx, cudaStream_t st, size_t batch, const T *a, const T *b, T *c, size_t m, size_t n,
                            size_t k, int ta, int tb) {
    if (tunable("gemm_jit") == 1 || batch < 64) return SLAS_ERR_UNSUPPORTED;
    if (!jit_ready()) return SLAS_ERR_UNSUPPORTED;
    JitKernel kern;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        // register tiles where the kernel's vector-width rules allow; else a thread per matrix if one fits a thread
        // (its a_trans / b_trans are run-time arguments: one specialisation serves all four flag pairs)
        JitCfg cfg;
        // ... and scalar register tiles for whatever is left (9x9, 10x10 f32, 5x9 ...)
        const bool tiled = choose_cfg(sizeof(T), m, n, k, &cfg);
        const bool tiny = !tiled && (choose_tiny_cfg(sizeof(T), m, n, k, &cfg) || choose_rect_cfg(sizeof(T), m, n, k, &cfg));
        const JitKey key{ctx->device, (int)sizeof(T) + (jit_bulk_c_store(sizeof(T), m, n) ? 100 : 0), m, n, k, tiny ? 0 : (ta ? 1 : 0), tiny ? 0 : (tb ? 1 : 0)};
        auto it = g_cache.find(key);
        if (it == g_cache.end()) {
            JitKernel fresh;
            fresh.cfg = cfg;
            fresh.failed = !(tiled || tiny) || !compile_kernel(sizeof(T), m, n, k, ta, tb, &fresh);
            it = g_cache.emplace(key, fresh).first;
        }
        kern = it->second;
    }
    if (kern.failed) return SLAS_ERR_UNSUPPORTED;
    const size_t G = (size_t)kern.cfg.g, nchunks = (batch + G - 1) / G;
    int per_sm = (int)((size_t)227 * 1024 / (kern.cfg.smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    size_t grid = (size_t)ctx->sm_count * per_sm;
    if (grid > nchunks) grid = nchunks;
    size_t batch_arg = batch, nchunks_arg = nchunks;
    int ta_arg = ta ? 1 : 0, tb_arg = tb ? 1 : 0;
    unsigned stage_arg = (unsigned)kern.cfg.stage_elems, cstage_arg = (unsigned)kern.cfg.cstage_elems;
    // gemm_small_kernel(a, b, c, batch, nchunks) / gemm_tiny_kernel(a, b, c, batch, nchunks, ta, tb, stage_elems, cstage_elems)
    void *params[] = {(void *)&a,      (void *)&b,      (void *)&c,         (void *)&batch_arg, (void *)&nchunks_arg,
                      (void *)&ta_arg, (void *)&tb_arg, (void *)&stage_arg, (void *)&cstage_arg};
    if (kern.cfg.tiny && ((batch * m * k * sizeof(T)) % 16 || (batch * k * n * sizeof(T)) % 16))
        return SLAS_ERR_UNSUPPORTED;  // its ragged last chunk is copied in whole 16-byte words
    return jit_launch(kern.fn, (unsigned)grid, (unsigned)kern.cfg.cons + 32, kern.cfg.smem, st, params);
}
template int launch_gemm_jit_batched<float>(Context *, cudaStream_t, size_t, const float *, const float *, float *, size_t, size_t,
                                            size_t, int, int);
template int launch_gemm_jit_batched<double>(Context *, cudaStream_t, size_t, const double *, const double *, double *, size_t,
                                             size_t, size_t, int, int);

}  // namespace slas
