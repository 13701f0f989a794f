// Probe / unit-test kernel for the tcgen05 building blocks in tc_common.cuh: out[128, N] = A[128, K] W[N, K]^T with
// A written to TMEM by the epilogue warps (tcgen05.st), W pre-packed into swizzled K-major stages and brought in
// with cp.async.bulk, tcgen05.mma kind::tf32 (1 or 3 terms), result read back with tcgen05.ld.
// Exposed as flowmc_test_tc_gemm so tests/test_gpu_tc.py can check the operand conventions the flow kernels rely on.
// TEST SCAFFOLDING: built into flowmc_b200/lib/libflowmc_b200_test.so (flowmc_b200/build.py), not into the product
// library; declared in include/flowmc_b200_test.h.
#include <string>

#include "../../../include/flowmc_b200_test.h"
#include "../registry.h"
#include "../tc_common.cuh"

namespace flowmc {

// W [N, K] row-major fp32 -> per K-chunk of 32: [hi image: Npad rows x 128 B][lo image], swizzled
__global__ void tc_pack_b_kernel(const float* __restrict__ W, int N, int K, int Npad, float* __restrict__ img) {
  const int n_kc = (K + 31) / 32;
  const int total = n_kc * Npad * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int kc = i / (Npad * 32), rem = i - kc * Npad * 32;
    const int n = rem / 32, kk = rem - n * 32;
    const int k = kc * 32 + kk;
    const float w = (n < N && k < K) ? W[(int64_t)n * K + k] : 0.0f;
    uint32_t hi, lo;
    tc::split_tf32_rn(w, hi, lo);
    float* stage = img + (int64_t)kc * 2 * Npad * 32;
    const int off = tc::packed_b_offset(n, kk) / 4;
    stage[off] = __uint_as_float(hi);
    stage[Npad * 32 + off] = __uint_as_float(lo);
  }
}

__global__ void __launch_bounds__(160) tc_gemm_test_kernel(const float* __restrict__ A, const float* __restrict__ img,
                                                           int N, int K, int terms, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);  // SWIZZLE_128B atoms: 1024-B aligned
  __shared__ uint64_t full[4], a_ready, acc_full;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_kc = K / 32;
  const uint32_t stage_bytes = 2u * N * 128u;

  if (warp == 4 && lane == 0) {
    for (int i = 0; i < 4; ++i) tc::mbar_init(&full[i], 1);
    tc::mbar_init(&a_ready, 128);
    tc::mbar_init(&acc_full, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t t_ahi = tbase, t_alo = tbase + 128, t_acc = tbase + 256;

  if (warp < 4) {
    const int row = tid;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int k = 0; k < K; k += 8) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) tc::split_tf32(A[(int64_t)row * K + k + j], hi[j], lo[j]);
      tc::tmem_st8(t_ahi + lane_base + k, hi);
      tc::tmem_st8(t_alo + lane_base + k, lo);
    }
    tc::tmem_wait_st();
    tc::tc_fence_before();
    tc::mbar_arrive(&a_ready);
    tc::mbar_wait(&acc_full, 0);
    tc::tc_fence_after();
    for (int c = 0; c < N; c += 32) {
      float v[32];
      tc::tmem_ld32(t_acc + lane_base + c, v);
      tc::tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c + j < N) out[(int64_t)row * N + c + j] = v[j];
    }
    tc::tc_fence_before();
  } else if (lane == 0) {
    for (int kc = 0; kc < n_kc; ++kc) {
      tc::mbar_arrive_expect_tx(&full[kc], stage_bytes);
      tc::bulk_g2s(smem + (size_t)kc * stage_bytes, img + (int64_t)kc * 2 * N * 32, stage_bytes, &full[kc]);
    }
    tc::mbar_wait(&a_ready, 0);
    tc::tc_fence_after();
    const uint32_t idesc = tc::make_idesc_tf32(128, N);
    uint32_t accum = 0;
    for (int kc = 0; kc < n_kc; ++kc) {
      tc::mbar_wait(&full[kc], 0);
      tc::tc_fence_after();
      const uint32_t b_hi = tc::smem_u32(smem + (size_t)kc * stage_bytes);
      const uint32_t b_lo = b_hi + N * 128;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t acol = kc * 32 + ks * 8;
        const uint64_t dhi = tc::make_b_desc(b_hi + ks * 32), dlo = tc::make_b_desc(b_lo + ks * 32);
        tc::mma_tf32_ts(t_acc, t_ahi + acol, dhi, idesc, accum);
        accum = 1;
        if (terms == 3) {
          tc::mma_tf32_ts(t_acc, t_alo + acol, dhi, idesc, 1);
          tc::mma_tf32_ts(t_acc, t_ahi + acol, dlo, idesc, 1);
        }
      }
    }
    tc::mma_commit(&acc_full);
  }
  __syncthreads();
  tc::tc_fence_after();
  if (warp == 0) tc::tmem_dealloc<512>(tbase);
}

// The same product on a CTA PAIR: out[256, N] = A[256, K] W[N, K]^T with tcgen05.mma.cta_group::2 (M = 256).
// CTA r of the 2-CTA cluster holds rows 128 r .. 128 r + 127 of A (and of the result) in its tensor memory and rows
// r N/2 .. of W (hi and lo images) in its shared memory -- each SM streams HALF of the weights.  Probe for the
// conventions a pair version of the flow kernels needs: allocation, operand placement, cross-CTA barriers.
__global__ void __launch_bounds__(160) tc_gemm_pair_test_kernel(const float* __restrict__ A,
                                                                const float* __restrict__ img, int N, int K, int terms,
                                                                float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full[4], peer_full[4], a_ready, acc_full;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = tc::cluster_rank();
  const int n_kc = K / 32;
  const int NH = N / 2;                                // B rows held by this CTA
  const uint32_t half_bytes = (uint32_t)NH * 128u;     // one image (hi or lo) of this CTA's rows
  const uint32_t stage_bytes = 2u * half_bytes;

  if (warp == 4 && lane == 0) {
    for (int i = 0; i < 4; ++i) {
      tc::mbar_init(&full[i], 1);
      tc::mbar_init(&peer_full[i], 1);
    }
    tc::mbar_init(&a_ready, 256);   // leader: 128 local + 128 remote epilogue threads
    tc::mbar_init(&acc_full, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc_pair<512>(&tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync();
  tc::tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t t_ahi = tbase, t_alo = tbase + 128, t_acc = tbase + 256;

  if (warp < 4) {
    const int row = tid;
    const int64_t grow = (int64_t)rank * 128 + row;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int k = 0; k < K; k += 8) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) tc::split_tf32(A[grow * K + k + j], hi[j], lo[j]);
      tc::tmem_st8(t_ahi + lane_base + k, hi);
      tc::tmem_st8(t_alo + lane_base + k, lo);
    }
    tc::tmem_wait_st();
    tc::tc_fence_before();
    if (rank == 0) tc::mbar_arrive(&a_ready);
    else tc::mbar_arrive_remote(&a_ready, 0);
    tc::mbar_wait(&acc_full, 0);
    tc::tc_fence_after();
    for (int c = 0; c < N; c += 16) {
      float v[16];
      tc::tmem_ld16(t_acc + lane_base + c, v);
      tc::tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) out[grow * N + c + j] = v[j];
    }
    tc::tc_fence_before();
  } else if (lane == 0) {
    // this CTA's half of every stage: rows rank * NH .. of the hi image, then of the lo image
    for (int kc = 0; kc < n_kc; ++kc) {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(img) + (size_t)kc * 2 * N * 128;
      uint8_t* dst = smem + (size_t)kc * stage_bytes;
      tc::mbar_arrive_expect_tx(&full[kc], stage_bytes);
      tc::bulk_g2s(dst, src + (size_t)rank * half_bytes, half_bytes, &full[kc]);
      tc::bulk_g2s(dst + half_bytes, src + (size_t)N * 128 + (size_t)rank * half_bytes, half_bytes, &full[kc]);
    }
    if (rank != 0) {
      // relay: tell the leader when this CTA's half of a stage has landed
      for (int kc = 0; kc < n_kc; ++kc) {
        tc::mbar_wait(&full[kc], 0);
        tc::mbar_arrive_remote(&peer_full[kc], 0);
      }
    } else {
      tc::mbar_wait(&a_ready, 0);
      tc::tc_fence_after();
      const uint32_t idesc = tc::make_idesc_tf32(256, N);
      uint32_t accum = 0;
      for (int kc = 0; kc < n_kc; ++kc) {
        tc::mbar_wait(&full[kc], 0);
        tc::mbar_wait(&peer_full[kc], 0);
        tc::tc_fence_after();
        const uint32_t b_hi = tc::smem_u32(smem + (size_t)kc * stage_bytes);
        const uint32_t b_lo = b_hi + half_bytes;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t acol = kc * 32 + ks * 8;
          const uint64_t dhi = tc::make_b_desc(b_hi + ks * 32), dlo = tc::make_b_desc(b_lo + ks * 32);
          tc::mma_tf32_ts_pair(t_acc, t_ahi + acol, dhi, idesc, accum);
          accum = 1;
          if (terms == 3) {
            tc::mma_tf32_ts_pair(t_acc, t_alo + acol, dhi, idesc, 1);
            tc::mma_tf32_ts_pair(t_acc, t_ahi + acol, dlo, idesc, 1);
          }
        }
      }
      tc::mma_commit_pair(&acc_full, 3);
    }
  }
  __syncthreads();
  tc::cluster_sync();
  tc::tc_fence_after();
  if (warp == 0) tc::tmem_dealloc_pair<512>(tbase);
}

}  // namespace flowmc

extern "C" {

// scratch: device, >= ceil(K/32) * 2 * N * 128 bytes
int flowmc_test_tc_gemm(const float* A, const float* W, int N, int K, int terms, float* out, float* scratch,
                         void* stream_) {
  using namespace flowmc;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!A || !W || !out || !scratch || N < 16 || N > 256 || (N % 16) || K < 32 || K > 128 || (K % 32) ||
      (terms != 1 && terms != 3)) {
    flowmc_set_error("test_tc_gemm: need N % 16 == 0 (16..256), K % 32 == 0 (32..128), terms in {1, 3}");
    return FLOWMC_ERR_INVALID;
  }
  const size_t smem = (size_t)(K / 32) * 2 * N * 128;
  if (smem > 200 * 1024) {
    flowmc_set_error("test_tc_gemm: stages do not fit shared memory");
    return FLOWMC_ERR_UNSUPPORTED;
  }
  tc_pack_b_kernel<<<64, 256, 0, stream>>>(W, N, K, N, scratch);
  flowmc_count_launch();
  cudaFuncSetAttribute(tc_gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem + 1024);
  tc_gemm_test_kernel<<<1, 160, smem + 1024, stream>>>(A, scratch, N, K, terms, out);
  flowmc_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    flowmc_set_error(cudaGetErrorString(e));
    return FLOWMC_ERR_CUDA;
  }
  return FLOWMC_OK;
}

// out[256, N] = A[256, K] W[N, K]^T on a CTA pair (cta_group::2).  scratch as above.
int flowmc_test_tc_gemm_pair(const float* A, const float* W, int N, int K, int terms, float* out, float* scratch,
                              void* stream_) {
  using namespace flowmc;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!A || !W || !out || !scratch || N < 32 || N > 256 || (N % 16) || K < 32 || K > 128 || (K % 32) ||
      (terms != 1 && terms != 3)) {
    flowmc_set_error("test_tc_gemm_pair: need N % 16 == 0 (32..256), K % 32 == 0 (32..128), terms in {1, 3}");
    return FLOWMC_ERR_INVALID;
  }
  const size_t smem = (size_t)(K / 32) * N * 128;   // per CTA: half of every stage
  tc_pack_b_kernel<<<64, 256, 0, stream>>>(W, N, K, N, scratch);
  flowmc_count_launch();
  cudaFuncSetAttribute(tc_gemm_pair_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem + 1024);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2);
  cfg.blockDim = dim3(160);
  cfg.dynamicSmemBytes = smem + 1024;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const float* img = scratch;
  cudaError_t e = cudaLaunchKernelEx(&cfg, tc_gemm_pair_test_kernel, A, img, N, K, terms, out);
  flowmc_count_launch();
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    flowmc_set_error(cudaGetErrorString(e));
    return FLOWMC_ERR_CUDA;
  }
  return FLOWMC_OK;
}

}  // extern "C"
