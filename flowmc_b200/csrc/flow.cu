// MaskedCouplingRQSpline forward / inverse / log_prob / sample -- fp32 parity path for sm_100a.
//
// Reference: src/flowMC/resource/model/nf_model/rqSpline.py:392-504 (model), common.py:68-124
// (MLP conditioner), :150-168 (masked coupling), :211-240 (ScalarAffine), :285-293 (Gaussian base).
//
// One CTA owns a tile of 64 samples and walks ALL coupling layers with the tile resident in
// shared memory: x tile, the two hidden activations, per-sample log-det.  Per layer:
//   ScalarAffine -> h1 = tanh(W1 (x*mask) + b1) -> h2 = tanh(W2 h1 + b2) -> for every transformed
//   feature f: raw[3K+1] = W3[f] h2 + b3[f] -> spline parameters -> rational-quadratic transform
// so the conditioner's output (the dominant d*(3K+1) x h GEMM) never leaves registers: the spline
// is the GEMM's epilogue.  Only the transformed half of W3's rows is ever read (the reference
// computes and discards the masked half, common.py:155-157 -- identical results).
// Thread map: lane = 2 samples (lane, lane+32); warp = a block of 8 output columns (dense stages)
// or one transformed feature (spline stage).  Weight rows are read through L1 with warp-uniform
// 128-bit loads (one transaction per warp), activations with conflict-free 128-bit LDS.
#include <cstdio>
#include <cstring>
#include <string>

#include "flow_tile.cuh"
#include "registry.h"

namespace flowmc {

template <int K, bool INV>
__global__ void __launch_bounds__(NT) flow_transform_kernel(const FlowmcFlowDesc D, const float* __restrict__ P,
                                                            const float* __restrict__ xin, int64_t n,
                                                            float* __restrict__ yout, float* __restrict__ ldout,
                                                            float* __restrict__ layer_inputs, int pre, int post,
                                                            const uint32_t* __restrict__ keys, Key host_key,
                                                            int64_t rows_per_key, const int32_t* __restrict__ idx) {
  extern __shared__ __align__(16) float smem[];
  const FlowSmem S = flow_smem_layout(D);
  float* xs = smem + S.xs;
  float* ld = smem + S.ld;
  const int d = D.n_features;
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * TM;

  // ---- load / generate the tile ---------------------------------------------------------------
  for (int i = tid; i < TM * d; i += NT) {
    const int s = i / d, j = i - s * d;
    const int64_t r = min(row0 + s, n - 1);
    float v;
    if (pre == PRE_NORMAL) {
      const int64_t kidx = r / rows_per_key;
      const Key key = keys ? Key{keys[2 * kidx], keys[2 * kidx + 1]} : host_key;
      const float z = bits_to_normal(bits_at(key, (uint64_t)((r - kidx * rows_per_key) * d + j)));
      v = P[D.off_base_mean + j] + z * sqrtf(P[D.off_base_cov + (int64_t)j * d + j]);
    } else {
      v = xin[(idx ? (int64_t)idx[r] : r) * d + j];
      if (pre == PRE_WHITEN) v = (v - P[D.off_data_mean + j]) / sqrtf(P[D.off_data_cov + (int64_t)j * d + j]);
    }
    xs[s * S.xs_stride + j] = v;
  }
  if (tid < TM) ld[tid] = 0.0f;
  __syncthreads();

  flow_layers<K, INV>(D, P, S, smem, row0, n, layer_inputs);
  if (layer_inputs != nullptr) {  // slot n_layers: the final latent (the training backward starts from it)
    for (int i = tid; i < TM * d; i += NT) {
      const int s = i / d, j = i - s * d;
      if (row0 + s < n) layer_inputs[((int64_t)D.n_layers * n + row0 + s) * d + j] = xs[s * S.xs_stride + j];
    }
  }

  // ---- epilogue ---------------------------------------------------------------------------------
  if (post == POST_BASE_LOGP) {
    if (tid < TM && row0 + tid < n) ldout[row0 + tid] = ld[tid] + base_log_prob(D, P, xs + tid * S.xs_stride);
  } else {
    for (int i = tid; i < TM * d; i += NT) {
      const int s = i / d, j = i - s * d;
      if (row0 + s < n) {
        float v = xs[s * S.xs_stride + j];
        if (post == POST_UNWHITEN) v = v * sqrtf(P[D.off_data_cov + (int64_t)j * d + j]) + P[D.off_data_mean + j];
        yout[(row0 + s) * d + j] = v;
      }
    }
    if (ldout != nullptr && tid < TM && row0 + tid < n) ldout[row0 + tid] = ld[tid];
  }
}

template <int K, bool INV>
static int launch_transform(const FlowmcFlowDesc& D, const float* P, const float* x, int64_t n, float* y, float* ld,
                            float* layer_inputs, int pre, int post, const uint32_t* keys, Key hk, int64_t rpk,
                            const int32_t* idx, cudaStream_t stream) {
  const FlowSmem S = flow_smem_layout(D);
  const size_t bytes = (size_t)S.total * sizeof(float);
  auto kern = flow_transform_kernel<K, INV>;
  static size_t configured = 0;
  if (bytes > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      flowmc_set_error("flow: model too large for the shared-memory tile (hidden width / n_features)");
      return FLOWMC_ERR_UNSUPPORTED;
    }
    configured = bytes;
  }
  const unsigned grid = (unsigned)((n + TM - 1) / TM);
  kern<<<grid, NT, bytes, stream>>>(D, P, x, n, y, ld, layer_inputs, pre, post, keys, hk, rpk, idx);
  flowmc_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    flowmc_set_error(cudaGetErrorString(e));
    return FLOWMC_ERR_CUDA;
  }
  return FLOWMC_OK;
}

int flow_transform(const FlowmcFlowDesc& D, bool inverse, const float* P, const float* x, int64_t n, float* y,
                   float* ld, float* layer_inputs, int pre, int post, const uint32_t* keys, Key hk, int64_t rpk,
                   cudaStream_t stream, const int32_t* idx) {
  if (n <= 0) return FLOWMC_OK;
  if (layer_inputs == nullptr && flow_tc_enabled(D))
    return flow_transform_tc(D, inverse, P, x, n, y, ld, pre, post, keys, hk, rpk, stream, idx);
#define FLOWMC_DISPATCH_K(KK)                                                                                      \
  case KK:                                                                                                         \
    return inverse                                                                                                 \
               ? launch_transform<KK, true>(D, P, x, n, y, ld, layer_inputs, pre, post, keys, hk, rpk, idx, stream) \
               : launch_transform<KK, false>(D, P, x, n, y, ld, layer_inputs, pre, post, keys, hk, rpk, idx, stream);
  switch (D.num_bins) {
    FLOWMC_DISPATCH_K(4)
    FLOWMC_DISPATCH_K(8)
    FLOWMC_DISPATCH_K(16)
    default:
      flowmc_set_error("flow: num_bins must be 4, 8 or 16");
      return FLOWMC_ERR_UNSUPPORTED;
  }
#undef FLOWMC_DISPATCH_K
}

}  // namespace flowmc

using flowmc::Key;

static int check_desc(const FlowmcFlowDesc* D, const char* who) {
  if (!D || D->n_features < 1 || D->n_layers < 1 || D->n_linear < 2 || D->n_linear > FLOWMC_FLOW_MAX_LINEAR) {
    flowmc_set_error((std::string(who) + ": bad flow descriptor").c_str());
    return FLOWMC_ERR_INVALID;
  }
  return FLOWMC_OK;
}

extern "C" {

int flowmc_flow_desc_init(FlowmcFlowDesc* D, int n_features, int n_layers, int n_hidden, const int* hidden,
                          int num_bins, float range_min, float range_max) {
  if (!D || !hidden || n_features < 1 || n_layers < 1 || n_hidden < 1 || n_hidden > FLOWMC_FLOW_MAX_LINEAR - 1 ||
      num_bins < 1) {
    flowmc_set_error("flow_desc_init: bad arguments (1..3 hidden layers supported)");
    return FLOWMC_ERR_INVALID;
  }
  std::memset(D, 0, sizeof(*D));
  D->n_features = n_features;
  D->n_layers = n_layers;
  D->n_linear = n_hidden + 1;
  D->num_bins = num_bins;
  D->range_min = range_min;
  D->range_max = range_max;
  D->dims[0] = n_features;
  for (int i = 0; i < n_hidden; ++i) D->dims[i + 1] = hidden[i];
  D->dims[n_hidden + 1] = n_features * (3 * num_bins + 1);
  auto pad4 = [](int64_t v) { return (v + 3) & ~(int64_t)3; };
  int64_t o = 0;
  for (int i = 0; i < D->n_linear; ++i) {
    D->off_W[i] = o;
    o = pad4(o + (int64_t)D->dims[i + 1] * D->dims[i]);
    D->off_b[i] = o;
    o = pad4(o + D->dims[i + 1]);
  }
  D->off_scale = o;
  D->off_shift = o + 1;
  D->layer_stride = pad4(o + 2);
  o = D->layer_stride * n_layers;
  D->off_data_mean = o;
  o = pad4(o + n_features);
  D->off_data_cov = o;
  o = pad4(o + (int64_t)n_features * n_features);
  D->off_base_mean = o;
  o = pad4(o + n_features);
  D->off_base_cov = o;
  o = pad4(o + (int64_t)n_features * n_features);
  D->n_params = o;
  return FLOWMC_OK;
}

int flowmc_flow_forward(const FlowmcFlowDesc* D, const float* params, const float* x, int64_t n, float* y,
                        float* logdet, void* stream) {
  if (int rc = check_desc(D, "flow_forward")) return rc;
  return flowmc::flow_transform(*D, false, params, x, n, y, logdet, nullptr, flowmc::PRE_NONE, flowmc::POST_NONE,
                                nullptr, Key{0, 0}, 1, (cudaStream_t)stream);
}

int flowmc_flow_inverse(const FlowmcFlowDesc* D, const float* params, const float* x, int64_t n, float* y,
                        float* logdet, void* stream) {
  if (int rc = check_desc(D, "flow_inverse")) return rc;
  return flowmc::flow_transform(*D, true, params, x, n, y, logdet, nullptr, flowmc::PRE_NONE, flowmc::POST_NONE,
                                nullptr, Key{0, 0}, 1, (cudaStream_t)stream);
}

int flowmc_flow_log_prob(const FlowmcFlowDesc* D, const float* params, const float* x, int64_t n, float* log_prob,
                         float* layer_inputs, void* stream) {
  if (int rc = check_desc(D, "flow_log_prob")) return rc;
  return flowmc::flow_transform(*D, false, params, x, n, nullptr, log_prob, layer_inputs, flowmc::PRE_WHITEN,
                                flowmc::POST_BASE_LOGP, nullptr, Key{0, 0}, 1, (cudaStream_t)stream);
}

int flowmc_flow_sample(const FlowmcFlowDesc* D, const float* params, const uint32_t* keys, const uint32_t host_key[2],
                       int64_t rows_per_key, int64_t n, float* x_out, void* stream) {
  if (int rc = check_desc(D, "flow_sample")) return rc;
  if (rows_per_key < 1 || (!keys && !host_key)) {
    flowmc_set_error("flow_sample: need keys and rows_per_key >= 1");
    return FLOWMC_ERR_INVALID;
  }
  const Key hk = host_key ? Key{host_key[0], host_key[1]} : Key{0, 0};
  return flowmc::flow_transform(*D, true, params, nullptr, n, x_out, nullptr, nullptr, flowmc::PRE_NORMAL,
                                flowmc::POST_UNWHITEN, keys, hk, rows_per_key, (cudaStream_t)stream);
}

}  // extern "C"
