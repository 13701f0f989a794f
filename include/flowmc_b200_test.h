/* flowmc_b200_test.h -- entry points of libflowmc_b200_test.so: TEST SCAFFOLDING, not part of the drop-in boundary.
 * Probe kernels for the tcgen05 building blocks (flowmc_b200/csrc/tc_common.cuh) that the tensor-core flow kernels are
 * built from; tests/test_gpu_tc.py checks the operand conventions through them. */
#ifndef FLOWMC_B200_TEST_H
#define FLOWMC_B200_TEST_H

#include "flowmc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* out[128, N] = A[128, K] W[N, K]^T through the tcgen05 path (A in TMEM, W as packed swizzled stages, kind::tf32,
 * terms = 1: plain TF32, 3: 3xTF32).  N % 16 == 0 (16..256), K % 32 == 0 (32..128); scratch: device,
 * >= (K/32) * 2 * N * 128 bytes. */
FLOWMC_API int flowmc_test_tc_gemm(const float* A, const float* W, int N, int K, int terms, float* out,
                                   float* scratch, void* stream);
/* out[256, N] = A[256, K] W[N, K]^T on a CTA pair: one 2-CTA cluster, tcgen05.mma.cta_group::2 with M = 256, each
 * CTA holding 128 rows of A / D in its tensor memory and half of W's rows in its shared memory.  Same limits
 * (N >= 32), same scratch. */
FLOWMC_API int flowmc_test_tc_gemm_pair(const float* A, const float* W, int N, int K, int terms, float* out,
                                        float* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLOWMC_B200_TEST_H */
