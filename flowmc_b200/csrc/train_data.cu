// Device-side data plumbing of TrainModel / NFModel.train: jax.random-compatible permutation and choice,
// selection of the training rows from the positions buffer, batch statistics.
//
// Reference: TrainModel.__call__ (src/flowMC/strategy/train_model.py:66-81: finite-row filter, last
// `history_window` steps per chain, jax.random.choice with replacement), NFModel.train /
// train_epoch (src/flowMC/resource/model/nf_model/base.py:141-144,187-188: jax.random.permutation
// batching, jnp.mean / jnp.cov of the training set).
#include <cub/cub.cuh>

#include <string>

#include "../../include/flowmc_b200.h"
#include "registry.h"
#include "rng.cuh"

namespace flowmc {

__global__ void iota_kernel(int32_t* out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (int32_t)i;
}
__global__ void bits_kernel(Key key, int64_t n, uint32_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = bits_at(key, (uint64_t)i);
}

// jax.random.randint(key, (m,), 0, span) (jax/_src/random.py: two 32-bit draws combined with
// multiplier (2^16 % span)^2 % span, all in uint32)
__global__ void randint_kernel(Key k1, Key k2, int64_t m, uint32_t span, int32_t* __restrict__ out) {
  uint32_t mult = 65536u % span;
  mult = (mult * mult) % span;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t hi = bits_at(k1, (uint64_t)i), lo = bits_at(k2, (uint64_t)i);
    const uint32_t off = ((hi % span) * mult + (lo % span)) % span;
    out[i] = (int32_t)off;
  }
}

// One warp per chain: rowmap[c][k] = step index of the k-th finite row (all d entries finite) of chain c,
// counts[c] = number of finite rows.
__global__ void __launch_bounds__(128) finite_rowmap_kernel(const float* __restrict__ buf, int64_t n_chains,
                                                            int64_t n_total, int d, int32_t* __restrict__ rowmap,
                                                            int32_t* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= n_chains) return;
  int cnt = 0;
  for (int64_t t0 = 0; t0 < n_total; t0 += 32) {
    const int64_t t = t0 + lane;
    bool fin = false;
    if (t < n_total) {
      fin = true;
      const float* row = buf + (c * n_total + t) * d;
      for (int j = 0; j < d; ++j) fin = fin && isfinite(row[j]);
    }
    const unsigned m = __ballot_sync(0xffffffffu, fin);
    if (fin) rowmap[c * n_total + cnt + __popc(m & ((1u << lane) - 1u))] = (int32_t)t;
    cnt += __popc(m);
  }
  if (lane == 0) counts[c] = cnt;
}

__global__ void minmax_kernel(const int32_t* __restrict__ v, int64_t n, int32_t* __restrict__ out) {
  int lo = INT32_MAX, hi = INT32_MIN;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    lo = min(lo, v[i]);
    hi = max(hi, v[i]);
  }
  __shared__ int slo[32], shi[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    slo[threadIdx.x >> 5] = lo;
    shi[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      lo = min(lo, slo[w]);
      hi = max(hi, shi[w]);
    }
    out[0] = lo;
    out[1] = hi;
  }
}

// out[i] = population row idx[i]; population row q = (chain q / w, the (m_finite - w + q % w)-th finite row)
__global__ void gather_training_rows_kernel(const float* __restrict__ buf, const int32_t* __restrict__ rowmap,
                                            int64_t n_total, int d, int w, int m_finite, int64_t chain_lo,
                                            int64_t chain_hi, const int32_t* __restrict__ idx, int64_t m,
                                            float* __restrict__ out) {
  const int64_t total = m * d;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / d;
    const int j = (int)(e - i * d);
    const int64_t q = idx[i];
    const int64_t c = q / w;
    if (c < chain_lo || c >= chain_hi) continue;  // row owned by another rank: left untouched
    const int64_t cl = c - chain_lo;
    const int t = rowmap[cl * n_total + (m_finite - w + (int)(q - c * w))];
    out[e] = buf[(cl * n_total + t) * d + j];
  }
}

// column sums of x [n, d] accumulated into sum[d] (zeroed by the caller)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int64_t n, int d,
                                                     float* __restrict__ sum) {
  // thread (ty, j): rows ty, ty + R, ... of this block's slab, column j
  const int R = 256 / d > 0 ? 256 / d : 1;
  const int ty = threadIdx.x / d;
  if (ty >= R) return;
  for (int j = threadIdx.x % d; j < d; j += 256) {  // more than one column per thread only if d > 256
    float s = 0.0f;
    for (int64_t r = (int64_t)blockIdx.x * R + ty; r < n; r += (int64_t)gridDim.x * R) s += x[r * d + j];
    atomicAdd(sum + j, s);
  }
}

// cov[a][b] += sum_rows (x[r][a] - mean[a]) (x[r][b] - mean[b]) / (n - 1); cov zeroed by the caller
__global__ void __launch_bounds__(256) cov_kernel(const float* __restrict__ x, int64_t n, int d,
                                                  const float* __restrict__ sum, float* __restrict__ cov) {
  extern __shared__ float sm[];  // [32][d] centred rows
  const float inv_n = 1.0f / (float)n;
  const float inv_nm1 = 1.0f / (float)(n - 1);
  const int dd = d * d;
  // output block ob: each thread owns outputs ob + tid, ob + tid + 256, ... (16 per thread; one block if d <= 64)
  for (int ob = 0; ob < dd; ob += 4096) {
    float acc[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = 0.0f;
    for (int64_t r0 = (int64_t)blockIdx.x * 32; r0 < n; r0 += (int64_t)gridDim.x * 32) {
      __syncthreads();
      for (int i = threadIdx.x; i < 32 * d; i += 256) {
        const int s = i / d, j = i - s * d;
        sm[i] = (r0 + s < n) ? x[(r0 + s) * d + j] - sum[j] * inv_n : 0.0f;
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int o = ob + threadIdx.x + 256 * q;
        if (o < dd) {
          const int a = o / d, b = o - a * d;
          float v = acc[q];
#pragma unroll 8
          for (int s = 0; s < 32; ++s) v = fmaf(sm[s * d + a], sm[s * d + b], v);
          acc[q] = v;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int o = ob + threadIdx.x + 256 * q;
      if (o < dd) atomicAdd(cov + o, acc[q] * inv_nm1);
    }
  }
}

__global__ void finish_mean_kernel(const float* __restrict__ sum, int d, float inv_n, float* __restrict__ mean) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < d) mean[j] = sum[j] * inv_n;
}

static int last_error(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    flowmc_set_error((std::string(what) + ": " + cudaGetErrorString(e)).c_str());
    return FLOWMC_ERR_CUDA;
  }
  return FLOWMC_OK;
}

static inline unsigned grid_for(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  const int64_t cap = 148 * 16;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

static inline int64_t align256(int64_t v) { return (v + 255) & ~(int64_t)255; }

static size_t sort_temp_bytes(int64_t n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  return bytes;
}

}  // namespace flowmc

extern "C" {

int64_t flowmc_random_permutation_workspace_bytes(int64_t n) {
  if (n <= 0) return 0;
  return 2 * flowmc::align256(n * 4) + flowmc::align256(n * 4) + flowmc::align256((int64_t)flowmc::sort_temp_bytes(n));
}

int flowmc_random_permutation(const uint32_t key[2], int64_t n, int32_t* out, void* workspace, int64_t workspace_bytes,
                              void* stream_) {
  using namespace flowmc;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!key || n < 0 || n > INT32_MAX || (n > 0 && (!out || !workspace)) ||
      workspace_bytes < flowmc_random_permutation_workspace_bytes(n)) {
    flowmc_set_error("random_permutation: bad arguments or workspace too small");
    return FLOWMC_ERR_INVALID;
  }
  if (n == 0) return FLOWMC_OK;
  char* ws = static_cast<char*>(workspace);
  uint32_t* keys_in = reinterpret_cast<uint32_t*>(ws);
  uint32_t* keys_out = reinterpret_cast<uint32_t*>(ws + align256(n * 4));
  int32_t* vals = reinterpret_cast<int32_t*>(ws + 2 * align256(n * 4));
  void* temp = ws + 3 * align256(n * 4);
  size_t temp_bytes = sort_temp_bytes(n);
  // jax._src.random._shuffle: num_rounds = ceil(3 ln(n) / ln(2^32 - 1)) rounds of a stable sort by fresh 32-bit keys
  const int rounds = (int)std::ceil(3.0 * std::log((double)(n > 1 ? n : 1)) / std::log(4294967295.0));
  iota_kernel<<<grid_for(n, 256), 256, 0, stream>>>(out, n);
  flowmc_count_launch();
  Key k{key[0], key[1]};
  for (int r = 0; r < rounds; ++r) {
    const Key sub = split_at(k, 1);
    k = split_at(k, 0);
    bits_kernel<<<grid_for(n, 256), 256, 0, stream>>>(sub, n, keys_in);
    cudaMemcpyAsync(vals, out, (size_t)n * 4, cudaMemcpyDeviceToDevice, stream);
    cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals, out, (int)n, 0, 32, stream);
    flowmc_count_launch();
    flowmc_count_launch();
  }
  return last_error("random_permutation");
}

int flowmc_random_choice(const uint32_t key[2], int64_t n_population, int64_t m, int32_t* out, void* stream_) {
  using namespace flowmc;
  if (!key || n_population < 1 || n_population > INT32_MAX || m < 0 || (m > 0 && !out)) {
    flowmc_set_error("random_choice: bad arguments");
    return FLOWMC_ERR_INVALID;
  }
  if (m == 0) return FLOWMC_OK;
  const Key k{key[0], key[1]};
  randint_kernel<<<grid_for(m, 256), 256, 0, (cudaStream_t)stream_>>>(split_at(k, 0), split_at(k, 1), m,
                                                                      (uint32_t)n_population, out);
  flowmc_count_launch();
  return last_error("random_choice");
}

int flowmc_buffer_finite_rows(const float* buf, int64_t n_chains, int64_t n_total, int d, int32_t* rowmap,
                              int32_t* counts, int32_t* minmax, void* stream_) {
  using namespace flowmc;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!buf || !rowmap || !counts || !minmax || n_chains < 1 || n_total < 1 || d < 1) {
    flowmc_set_error("buffer_finite_rows: bad arguments");
    return FLOWMC_ERR_INVALID;
  }
  finite_rowmap_kernel<<<(unsigned)((n_chains + 3) / 4), 128, 0, stream>>>(buf, n_chains, n_total, d, rowmap, counts);
  flowmc_count_launch();
  minmax_kernel<<<1, 1024, 0, stream>>>(counts, n_chains, minmax);
  flowmc_count_launch();
  return last_error("buffer_finite_rows");
}

int flowmc_gather_training_rows(const float* buf, const int32_t* rowmap, int64_t n_total, int d, int window,
                                int m_finite, int64_t chain_lo, int64_t chain_hi, const int32_t* idx, int64_t m,
                                float* out, void* stream_) {
  using namespace flowmc;
  if (!buf || !rowmap || !idx || !out || window < 1 || window > m_finite || m < 0 || chain_hi < chain_lo) {
    flowmc_set_error("gather_training_rows: bad arguments");
    return FLOWMC_ERR_INVALID;
  }
  if (m == 0) return FLOWMC_OK;
  gather_training_rows_kernel<<<grid_for(m * d, 256), 256, 0, (cudaStream_t)stream_>>>(
      buf, rowmap, n_total, d, window, m_finite, chain_lo, chain_hi, idx, m, out);
  flowmc_count_launch();
  return last_error("gather_training_rows");
}

int flowmc_data_mean_cov(const float* x, int64_t n, int d, float* mean, float* cov, float* scratch, void* stream_) {
  using namespace flowmc;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!x || !mean || !cov || !scratch || n < 2 || d < 1 || d > 512) {
    flowmc_set_error("data_mean_cov: bad arguments (2 <= n, 1 <= d <= 512)");
    return FLOWMC_ERR_INVALID;
  }
  cudaMemsetAsync(scratch, 0, (size_t)d * sizeof(float), stream);
  cudaMemsetAsync(cov, 0, (size_t)d * d * sizeof(float), stream);
  const int R = 256 / d > 0 ? 256 / d : 1;
  int64_t blocks = (n + R - 1) / R;
  if (blocks > 148 * 4) blocks = 148 * 4;
  colsum_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, n, d, scratch);
  flowmc_count_launch();
  blocks = (n + 31) / 32;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (32 * d * sizeof(float) > 48 * 1024) {  // d > 384: beyond the default dynamic shared-memory limit
    static bool opted_in = false;
    if (!opted_in) {
      cudaFuncSetAttribute(cov_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 512 * (int)sizeof(float));
      opted_in = true;
    }
  }
  cov_kernel<<<(unsigned)blocks, 256, 32 * d * sizeof(float), stream>>>(x, n, d, scratch, cov);
  flowmc_count_launch();
  finish_mean_kernel<<<(unsigned)((d + 63) / 64), 64, 0, stream>>>(scratch, d, 1.0f / (float)n, mean);
  flowmc_count_launch();
  return last_error("data_mean_cov");
}

}  // extern "C"
