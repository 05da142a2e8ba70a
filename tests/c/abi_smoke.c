/* abi_smoke.c -- the C ABI from plain C99 (gcc -std=c99 -pedantic -Wall -Werror): the header must be C-clean, the
 * library must link from C, the pure-host entry points must work without a GPU and every compute entry must fail
 * loudly (SLAS_ERR_CUDA / SLAS_ERR_INVALID, never a CPU fallback) when there is none.  Built and run by
 * tests/test_host_logic.py::test_c99_program_links_and_runs (CPU suite). */
#include <stdio.h>
#include <string.h>

#include "slas_cuda.h"

int main(void) {
    size_t lo = 0, hi = 0, covered = 0;
    int r, ndev = 0;
    /* slice arithmetic of the sharded ops: pure host code */
    for (r = 0; r < 8; ++r) {
        if (slas_cuda_shard_range(1000003, sizeof(float), 8, r, &lo, &hi) != SLAS_OK) return 1;
        if (lo != covered || (r < 7 && (hi * sizeof(float)) % 16 != 0)) return 2;
        covered = hi;
    }
    if (covered != 1000003) return 3;
    if (slas_cuda_shard_range(10, 3, 2, 0, &lo, &hi) != SLAS_ERR_INVALID) return 4; /* element size must divide 16 */
    if (strlen(slas_cuda_last_error()) == 0) return 5;
    slas_cuda_device_count(&ndev);
    if (ndev == 0) {
        /* no GPU: compute entries report it, they do not compute on the CPU */
        float a[4] = {0.f, 1.f, 2.f, 3.f}, out = -1.f;
        const int st = slas_cuda_sdot(4, a, a, &out);
        if (st == SLAS_OK || out != -1.f) return 6;
        printf("no CUDA device: slas_cuda_sdot -> status %d (%s)\n", st, slas_cuda_last_error());
    } else {
        float a[4] = {0.f, 1.f, 2.f, 3.f}, b[4] = {0.f, -1.f, -2.f, 3.f}, out = 0.f; /* tests/src/main.rs:41-46 */
        if (slas_cuda_sdot(4, a, b, &out) != SLAS_OK || out != 4.f) return 7;
        printf("slas_cuda_sdot on the GPU: %g\n", out);
    }
    printf("C_ABI_OK %llu launches\n", (unsigned long long)slas_cuda_launch_count());
    return 0;
}
