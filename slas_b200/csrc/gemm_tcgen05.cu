// gemm_tcgen05.cu -- placeholder until the 3xTF32 tcgen05/TMEM kernel lands.
#include "common.cuh"
#include "kernels.cuh"

namespace slas {
int launch_sgemm_3xtf32(Context *, cudaStream_t, const float *, const float *, float *, size_t, size_t, size_t, size_t,
                        size_t, size_t, int, int) {
    return SLAS_ERR_UNSUPPORTED;
}
}  // namespace slas
