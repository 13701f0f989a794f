// RealNVP forward / inverse / log_prob / sample and loss gradient -- fp32 CUDA-core kernels for sm_100a.
//
// Reference: src/flowMC/resource/model/nf_model/realNVP.py:102-228 (RealNVP), :18-100 (AffineCoupling),
// src/flowMC/resource/model/common.py:68-124 (MLP, default activation relu), :150-168 (MaskedCouplingLayer),
// :171-209 (MLPAffine), :285-293 (Gaussian base), nf_model/base.py:98-100 (loss_fn).
//
// One CTA owns a tile of 32 samples (lane = sample) and walks ALL coupling layers with the tile resident in shared
// memory.  Per layer, with the FLOAT mask m of the reference (a model leaf that weight decay shrinks, see
// oracle/realnvp.py):
//     cond  = x * m
//     scale = tanh(W2s relu(W1s cond + b1s) + b2s) * dt         shift = (W2t relu(W1t cond + b1t) + b2t) * dt
//     fwd:  y = (x + shift) * exp(scale)      inv:  y = x * exp(-scale) - shift
//     x'    = (1 - m) * y + m * x             log_det += sum_j (1 - m_j) * (+-scale_j)
// The two conditioner MLPs are the GEMMs; the affine transform is their epilogue (both outputs stay in shared memory,
// nothing leaves the SM between layers).  The loss gradient is a hand-derived reverse pass with the same tile
// structure: the forward pass leaves every layer's input in an L2-resident scratch, the backward CTA walks the layers
// in reverse, recomputes the hidden activations and accumulates dW / db with red.global.add.f32.
//
// The spline flow's tcgen05 pipeline (flow_tc.cu) is not used here: RealNVP's conditioners are two GEMMs of
// d x h x d (8 d h flop per sample and layer, 32 kflop at d = 32, h = 128 -- 1/40 of an RQ-spline layer), so the
// kernel is bound by its tanh / exp epilogue and shared-memory traffic, not by the FMA pipe.
#include <cstring>
#include <string>

#include "../../include/flowmc_b200.h"
#include "registry.h"
#include "rng.cuh"

namespace flowmc {
namespace nvp {

constexpr int TM = 32;   // samples per CTA (lane = sample)
constexpr int NT = 256;  // threads per CTA
constexpr int NW = NT / 32;

enum : int { PRE_NONE = 0, PRE_WHITEN = 1, PRE_NORMAL = 2 };
enum : int { POST_NONE = 0, POST_BASE_LOGP = 1, POST_UNWHITEN = 2 };

__host__ __device__ inline int round4(int v) { return (v + 3) & ~3; }

struct Smem {  // offsets in floats
  int xs, s, t, g, ds, dt, du, h, dh, ld, xs_stride, h_stride, total;
};
__host__ __device__ inline Smem smem_layout(const FlowmcRealNVPDesc& D, bool backward) {
  Smem m;
  m.xs_stride = round4(D.n_features) + 4;
  m.h_stride = round4(D.n_hidden) + 4;
  int o = 0;
  m.xs = o; o += TM * m.xs_stride;
  m.s = o; o += TM * m.xs_stride;
  m.t = o; o += TM * m.xs_stride;
  m.h = o; o += TM * m.h_stride;
  m.g = m.ds = m.dt = m.du = m.dh = 0;
  if (backward) {
    m.g = o; o += TM * m.xs_stride;
    m.ds = o; o += TM * m.xs_stride;
    m.dt = o; o += TM * m.xs_stride;
    m.du = o; o += TM * m.xs_stride;
    m.dh = o; o += TM * m.h_stride;
  }
  m.ld = o; o += TM;
  m.total = o;
  return m;
}

// out[r][n] = act(b[n] + sum_k in[r][k] * W[n][k])  for the tile's 32 rows; in / out in shared memory.
// MASKED: in[r][k] is multiplied by mask[k] (the conditioner sees x * mask).  RELU: act = max(., 0).
template <bool RELU, bool MASKED>
__device__ __forceinline__ void dense_stage(const float* __restrict__ in_s, int in_stride, int Kd,
                                            const float* __restrict__ W, const float* __restrict__ b, int N,
                                            float* __restrict__ out_s, int out_stride,
                                            const float* __restrict__ mask) {
  constexpr int NB = 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* row = in_s + lane * in_stride;
  for (int n0 = warp * NB; n0 < N; n0 += NW * NB) {
    float acc[NB];
    const float* wrow[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int n = min(n0 + j, N - 1);
      acc[j] = __ldg(b + n);
      wrow[j] = W + (int64_t)n * Kd;
    }
    for (int k = 0; k < Kd; ++k) {
      float a = row[k];
      if (MASKED) a = a * __ldg(mask + k);
#pragma unroll
      for (int j = 0; j < NB; ++j) acc[j] = fmaf(a, __ldg(wrow[j] + k), acc[j]);
    }
#pragma unroll
    for (int j = 0; j < NB; ++j)
      if (n0 + j < N) out_s[lane * out_stride + n0 + j] = RELU ? fmaxf(acc[j], 0.0f) : acc[j];
  }
}

// dA[r][k] (+)= sum_n G[r][n] * W[n][k], optionally gated by relu'(act[r][k]) = act > 0
template <bool GATE, bool ACCUM>
__device__ __forceinline__ void back_stage(const float* __restrict__ g_s, int g_stride, int N,
                                           const float* __restrict__ W, int Kd, float* __restrict__ out_s,
                                           int out_stride, const float* __restrict__ act_s, int act_stride) {
  constexpr int KB = 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* grow = g_s + lane * g_stride;
  for (int k0 = warp * KB; k0 < Kd; k0 += NW * KB) {
    float acc[KB];
#pragma unroll
    for (int j = 0; j < KB; ++j) acc[j] = 0.0f;
    for (int n = 0; n < N; ++n) {
      const float gv = grow[n];
      const float* wr = W + (int64_t)n * Kd + k0;
#pragma unroll
      for (int j = 0; j < KB; ++j)
        if (k0 + j < Kd) acc[j] = fmaf(gv, __ldg(wr + j), acc[j]);
    }
#pragma unroll
    for (int j = 0; j < KB; ++j) {
      if (k0 + j < Kd) {
        float v = acc[j];
        if (GATE) v = act_s[lane * act_stride + k0 + j] > 0.0f ? v : 0.0f;
        float* o = out_s + lane * out_stride + k0 + j;
        *o = ACCUM ? *o + v : v;
      }
    }
  }
}

// dW[n][k] += sum_r G[r][n] * A[r][k] * (mask ? mask[k] : 1)   and   db[n] += sum_r G[r][n]
// (reduction over the tile's samples; results go to global memory with atomics)
__device__ __forceinline__ void weight_grad(const float* __restrict__ g_s, int g_stride, int N,
                                            const float* __restrict__ a_s, int a_stride, int Kd,
                                            float* __restrict__ dW, float* __restrict__ db,
                                            const float* __restrict__ mask, int rows) {
  for (int i = threadIdx.x; i < N * Kd; i += NT) {
    const int n = i / Kd, k = i - n * Kd;
    float acc = 0.0f;
    for (int r = 0; r < rows; ++r) acc = fmaf(g_s[r * g_stride + n], a_s[r * a_stride + k], acc);
    if (mask != nullptr) acc *= __ldg(mask + k);
    if (acc != 0.0f) atomicAdd(dW + i, acc);
  }
  for (int n = threadIdx.x; n < N; n += NT) {
    float acc = 0.0f;
    for (int r = 0; r < rows; ++r) acc += g_s[r * g_stride + n];
    if (acc != 0.0f) atomicAdd(db + n, acc);
  }
}

// scale (pre-tanh) and shift of one layer for the tile: s_s <- W2s relu(W1s (x m) + b1s) + b2s, t_s likewise.
// Ends with a __syncthreads(); h_s holds the SHIFT MLP's hidden activations on return if shift_last, else the
// SCALE MLP's.
__device__ __forceinline__ void conditioners(const FlowmcRealNVPDesc& D, const float* __restrict__ PL, const Smem& S,
                                             float* smem, bool shift_last) {
  const int d = D.n_features, h = D.n_hidden;
  float* xs = smem + S.xs;
  float* hb = smem + S.h;
  const float* mask = PL + D.off_mask;
  for (int pass = 0; pass < 2; ++pass) {
    const bool do_shift = (pass == 1) == shift_last;
    const float* W1 = PL + (do_shift ? D.off_W1t : D.off_W1s);
    const float* b1 = PL + (do_shift ? D.off_b1t : D.off_b1s);
    const float* W2 = PL + (do_shift ? D.off_W2t : D.off_W2s);
    const float* b2 = PL + (do_shift ? D.off_b2t : D.off_b2s);
    dense_stage<true, true>(xs, S.xs_stride, d, W1, b1, h, hb, S.h_stride, mask);
    __syncthreads();
    dense_stage<false, false>(hb, S.h_stride, h, W2, b2, d, smem + (do_shift ? S.t : S.s), S.xs_stride, nullptr);
    __syncthreads();
  }
}

template <bool INV>
__device__ __forceinline__ void coupling_layers(const FlowmcRealNVPDesc& D, const float* __restrict__ P, const Smem& S,
                                                float* smem, int64_t row0, int64_t n,
                                                float* __restrict__ layer_inputs) {
  float* xs = smem + S.xs;
  float* ss = smem + S.s;
  float* ts = smem + S.t;
  float* ld = smem + S.ld;
  const int d = D.n_features;
  const int tid = threadIdx.x;
  for (int li = 0; li < D.n_layers; ++li) {
    const int l = INV ? D.n_layers - 1 - li : li;
    const float* PL = P + (int64_t)l * D.layer_stride;
    const float* mask = PL + D.off_mask;
    if (layer_inputs != nullptr) {
      for (int i = tid; i < TM * d; i += NT) {
        const int s = i / d, j = i - s * d;
        if (row0 + s < n) layer_inputs[((int64_t)l * n + row0 + s) * d + j] = xs[s * S.xs_stride + j];
      }
    }
    conditioners(D, PL, S, smem, true);
    // affine epilogue (common.py:186-209) + masked blend (common.py:155-157, 165-167)
    for (int i = tid; i < TM * d; i += NT) {
      const int s = i / d, j = i - s * d;
      const int o = s * S.xs_stride + j;
      const float m = __ldg(mask + j);
      const float scale = tanhf(ss[o]) * D.dt;
      const float shift = ts[o] * D.dt;
      const float x = xs[o];
      const float y = INV ? x * expf(-scale) - shift : (x + shift) * expf(scale);
      xs[o] = (1.0f - m) * y + m * x;
      ss[o] = (1.0f - m) * (INV ? -scale : scale);
    }
    __syncthreads();
    if (tid < TM) {  // log_det = sum over features in index order (fixed order: run-to-run reproducible)
      float acc = 0.0f;
      for (int j = 0; j < d; ++j) acc += ss[tid * S.xs_stride + j];
      ld[tid] += acc;
    }
    __syncthreads();
  }
}

template <bool INV>
__global__ void __launch_bounds__(NT) realnvp_kernel(const FlowmcRealNVPDesc D, const float* __restrict__ P,
                                                     const float* __restrict__ xin, int64_t n,
                                                     float* __restrict__ yout, float* __restrict__ ldout,
                                                     float* __restrict__ layer_inputs, int pre, int post,
                                                     const uint32_t* __restrict__ keys, Key host_key,
                                                     int64_t rows_per_key, const int32_t* __restrict__ idx) {
  extern __shared__ __align__(16) float smem[];
  const Smem S = smem_layout(D, false);
  float* xs = smem + S.xs;
  float* ld = smem + S.ld;
  const int d = D.n_features;
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * TM;
  for (int i = tid; i < TM * d; i += NT) {
    const int s = i / d, j = i - s * d;
    const int64_t r = min(row0 + s, n - 1);
    float v;
    if (pre == PRE_NORMAL) {  // Gaussian.sample (common.py:288-293): mean + chol(cov) z, diagonal covariance
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

  coupling_layers<INV>(D, P, S, smem, row0, n, layer_inputs);
  if (layer_inputs != nullptr) {  // slot n_layers: the final latent (the backward pass starts from it)
    for (int i = tid; i < TM * d; i += NT) {
      const int s = i / d, j = i - s * d;
      if (row0 + s < n) layer_inputs[((int64_t)D.n_layers * n + row0 + s) * d + j] = xs[s * S.xs_stride + j];
    }
  }
  if (post == POST_BASE_LOGP) {
    // realNVP.py:218-220: multivariate_normal.logpdf(y, zeros, eye) -- literal zeros / eye, not base_dist
    if (tid < TM && row0 + tid < n) {
      float q = 0.0f;
      for (int j = 0; j < d; ++j) {
        const float v = xs[tid * S.xs_stride + j];
        q += v * v;
      }
      ldout[row0 + tid] = ld[tid] + (-0.5f * q - (float)d * 0.5f * 1.8378770664093453f);
    }
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

// Reverse pass of loss = -inv_n * sum_i log_prob(x_i) for one tile (layer inputs from the forward pass).
__global__ void __launch_bounds__(NT) realnvp_backward_kernel(const FlowmcRealNVPDesc D, const float* __restrict__ P,
                                                              const float* __restrict__ layer_inputs,
                                                              const float* __restrict__ logp, int64_t n,
                                                              float inv_n, float* __restrict__ grad,
                                                              float* __restrict__ loss) {
  extern __shared__ __align__(16) float smem[];
  const Smem S = smem_layout(D, true);
  float* xs = smem + S.xs;
  float* ss = smem + S.s;
  float* ts = smem + S.t;
  float* gs = smem + S.g;
  float* dss = smem + S.ds;
  float* dts = smem + S.dt;
  float* dus = smem + S.du;
  float* hb = smem + S.h;
  float* dhb = smem + S.dh;
  const int d = D.n_features, h = D.n_hidden;
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * TM;
  const int rows = (int)min((int64_t)TM, n - row0);

  // dL/dy_L = inv_n * y_L (base term -1/2 |y|^2); rows beyond n carry zero gradient
  for (int i = tid; i < TM * d; i += NT) {
    const int s = i / d, j = i - s * d;
    gs[s * S.xs_stride + j] = s < rows ? inv_n * layer_inputs[((int64_t)D.n_layers * n + row0 + s) * d + j] : 0.0f;
  }
  if (tid < 32) {  // loss contribution of this tile
    float v = tid < rows ? -logp[row0 + tid] * inv_n : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (tid == 0) atomicAdd(loss, v);
  }
  const float g_ld = -inv_n;  // dL/dlog_det

  for (int l = D.n_layers - 1; l >= 0; --l) {
    const float* PL = P + (int64_t)l * D.layer_stride;
    float* GL = grad + (int64_t)l * D.layer_stride;
    const float* mask = PL + D.off_mask;
    __syncthreads();
    for (int i = tid; i < TM * d; i += NT) {
      const int s = i / d, j = i - s * d;
      xs[s * S.xs_stride + j] = s < rows ? layer_inputs[((int64_t)l * n + row0 + s) * d + j] : 0.0f;
    }
    __syncthreads();
    conditioners(D, PL, S, smem, false);  // ss, ts = pre-activation scale, shift; hb = SCALE MLP's hidden layer
    // adjoints of the affine epilogue
    for (int i = tid; i < TM * d; i += NT) {
      const int s = i / d, j = i - s * d;
      const int o = s * S.xs_stride + j;
      const float m = __ldg(mask + j), a = 1.0f - m;
      const float th = tanhf(ss[o]);
      const float scale = th * D.dt, shift = ts[o] * D.dt;
      const float e = expf(scale);
      const float gy = s < rows ? gs[o] : 0.0f;
      const float x = xs[o];
      const float dscale = a * (gy * (x + shift) * e + (s < rows ? g_ld : 0.0f));
      dss[o] = dscale * D.dt * (1.0f - th * th);   // d / d(pre-tanh output of the scale MLP)
      dts[o] = a * gy * e * D.dt;                  // d / d(output of the shift MLP)
      gs[o] = gy * (a * e + m);                    // direct path to x; the conditioner path (du * m) is added below
    }
    __syncthreads();
    // ---- scale MLP backward (hb = its hidden activations) ----
    weight_grad(dss, S.xs_stride, d, hb, S.h_stride, h, GL + D.off_W2s, GL + D.off_b2s, nullptr, rows);
    back_stage<true, false>(dss, S.xs_stride, d, PL + D.off_W2s, h, dhb, S.h_stride, hb, S.h_stride);
    __syncthreads();
    weight_grad(dhb, S.h_stride, h, xs, S.xs_stride, d, GL + D.off_W1s, GL + D.off_b1s, mask, rows);
    back_stage<false, false>(dhb, S.h_stride, h, PL + D.off_W1s, d, dus, S.xs_stride, nullptr, 0);
    __syncthreads();
    // ---- shift MLP backward: recompute its hidden layer into hb ----
    dense_stage<true, true>(xs, S.xs_stride, d, PL + D.off_W1t, PL + D.off_b1t, h, hb, S.h_stride, mask);
    __syncthreads();
    weight_grad(dts, S.xs_stride, d, hb, S.h_stride, h, GL + D.off_W2t, GL + D.off_b2t, nullptr, rows);
    back_stage<true, false>(dts, S.xs_stride, d, PL + D.off_W2t, h, dhb, S.h_stride, hb, S.h_stride);
    __syncthreads();
    weight_grad(dhb, S.h_stride, h, xs, S.xs_stride, d, GL + D.off_W1t, GL + D.off_b1t, mask, rows);
    back_stage<false, true>(dhb, S.h_stride, h, PL + D.off_W1t, d, dus, S.xs_stride, nullptr, 0);
    __syncthreads();
    for (int i = tid; i < TM * d; i += NT) {
      const int s = i / d, j = i - s * d;
      const int o = s * S.xs_stride + j;
      gs[o] += __ldg(mask + j) * dus[o];
    }
  }
}

static int launch_forward(const FlowmcRealNVPDesc& D, bool inverse, const float* P, const float* x, int64_t n,
                          float* y, float* ld, float* layer_inputs, int pre, int post, const uint32_t* keys, Key hk,
                          int64_t rpk, const int32_t* idx, cudaStream_t stream) {
  if (n <= 0) return FLOWMC_OK;
  const Smem S = smem_layout(D, false);
  const size_t bytes = (size_t)S.total * sizeof(float);
  static size_t configured[2] = {0, 0};
  if (bytes > configured[inverse ? 1 : 0]) {
    cudaError_t e = inverse ? cudaFuncSetAttribute(realnvp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)bytes)
                            : cudaFuncSetAttribute(realnvp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      flowmc_set_error("realnvp: model too large for the shared-memory tile (n_features / n_hidden)");
      return FLOWMC_ERR_UNSUPPORTED;
    }
    configured[inverse ? 1 : 0] = bytes;
  }
  const unsigned grid = (unsigned)((n + TM - 1) / TM);
  if (inverse)
    realnvp_kernel<true><<<grid, NT, bytes, stream>>>(D, P, x, n, y, ld, layer_inputs, pre, post, keys, hk, rpk, idx);
  else
    realnvp_kernel<false><<<grid, NT, bytes, stream>>>(D, P, x, n, y, ld, layer_inputs, pre, post, keys, hk, rpk, idx);
  flowmc_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    flowmc_set_error(cudaGetErrorString(e));
    return FLOWMC_ERR_CUDA;
  }
  return FLOWMC_OK;
}

static inline int64_t pad4i(int64_t v) { return (v + 3) & ~(int64_t)3; }

}  // namespace nvp
}  // namespace flowmc

using flowmc::Key;

static int check_nvp(const FlowmcRealNVPDesc* D, const char* who) {
  if (!D || D->n_features < 1 || D->n_layers < 1 || D->n_hidden < 1) {
    flowmc_set_error((std::string(who) + ": bad RealNVP descriptor").c_str());
    return FLOWMC_ERR_INVALID;
  }
  return FLOWMC_OK;
}

extern "C" {

int flowmc_realnvp_desc_init(FlowmcRealNVPDesc* D, int n_features, int n_layers, int n_hidden, float dt) {
  if (!D || n_features < 1 || n_layers < 1 || n_hidden < 1) {
    flowmc_set_error("realnvp_desc_init: bad arguments");
    return FLOWMC_ERR_INVALID;
  }
  std::memset(D, 0, sizeof(*D));
  D->n_features = n_features;
  D->n_layers = n_layers;
  D->n_hidden = n_hidden;
  D->dt = dt;
  auto pad4 = [](int64_t v) { return (v + 3) & ~(int64_t)3; };
  const int64_t d = n_features, h = n_hidden;
  int64_t o = 0;
  D->off_W1s = o; o = pad4(o + h * d);
  D->off_b1s = o; o = pad4(o + h);
  D->off_W2s = o; o = pad4(o + d * h);
  D->off_b2s = o; o = pad4(o + d);
  D->off_W1t = o; o = pad4(o + h * d);
  D->off_b1t = o; o = pad4(o + h);
  D->off_W2t = o; o = pad4(o + d * h);
  D->off_b2t = o; o = pad4(o + d);
  D->off_mask = o; o = pad4(o + d);
  D->layer_stride = o;
  o = D->layer_stride * n_layers;
  D->off_data_mean = o; o = pad4(o + d);
  D->off_data_cov = o; o = pad4(o + d * d);
  D->off_base_mean = o; o = pad4(o + d);
  D->off_base_cov = o; o = pad4(o + d * d);
  D->n_params = o;
  return FLOWMC_OK;
}

int flowmc_realnvp_forward(const FlowmcRealNVPDesc* D, const float* params, const float* x, int64_t n, float* y,
                           float* logdet, void* stream) {
  if (int rc = check_nvp(D, "realnvp_forward")) return rc;
  return flowmc::nvp::launch_forward(*D, false, params, x, n, y, logdet, nullptr, flowmc::nvp::PRE_NONE,
                                     flowmc::nvp::POST_NONE, nullptr, Key{0, 0}, 1, nullptr, (cudaStream_t)stream);
}

int flowmc_realnvp_inverse(const FlowmcRealNVPDesc* D, const float* params, const float* x, int64_t n, float* y,
                           float* logdet, void* stream) {
  if (int rc = check_nvp(D, "realnvp_inverse")) return rc;
  return flowmc::nvp::launch_forward(*D, true, params, x, n, y, logdet, nullptr, flowmc::nvp::PRE_NONE,
                                     flowmc::nvp::POST_NONE, nullptr, Key{0, 0}, 1, nullptr, (cudaStream_t)stream);
}

int flowmc_realnvp_log_prob(const FlowmcRealNVPDesc* D, const float* params, const float* x, int64_t n,
                            float* log_prob, void* stream) {
  if (int rc = check_nvp(D, "realnvp_log_prob")) return rc;
  return flowmc::nvp::launch_forward(*D, false, params, x, n, nullptr, log_prob, nullptr, flowmc::nvp::PRE_WHITEN,
                                     flowmc::nvp::POST_BASE_LOGP, nullptr, Key{0, 0}, 1, nullptr,
                                     (cudaStream_t)stream);
}

int flowmc_realnvp_sample(const FlowmcRealNVPDesc* D, const float* params, const uint32_t* keys,
                          const uint32_t host_key[2], int64_t rows_per_key, int64_t n, float* x_out, void* stream) {
  if (int rc = check_nvp(D, "realnvp_sample")) return rc;
  if (rows_per_key < 1 || (!keys && !host_key)) {
    flowmc_set_error("realnvp_sample: need keys and rows_per_key >= 1");
    return FLOWMC_ERR_INVALID;
  }
  const Key hk = host_key ? Key{host_key[0], host_key[1]} : Key{0, 0};
  return flowmc::nvp::launch_forward(*D, true, params, nullptr, n, x_out, nullptr, nullptr, flowmc::nvp::PRE_NORMAL,
                                     flowmc::nvp::POST_UNWHITEN, keys, hk, rows_per_key, nullptr,
                                     (cudaStream_t)stream);
}

int64_t flowmc_realnvp_loss_grad_workspace_bytes(const FlowmcRealNVPDesc* D, int64_t n) {
  if (!D || n <= 0) return 0;
  return 4 * (flowmc::nvp::pad4i((int64_t)(D->n_layers + 1) * n * D->n_features) + flowmc::nvp::pad4i(n));
}

int flowmc_realnvp_loss_grad(const FlowmcRealNVPDesc* D, const float* params, const float* x, const int32_t* idx,
                             int64_t n, float inv_n_total, float* grad, float* loss, void* workspace,
                             int64_t workspace_bytes, void* stream_) {
  using namespace flowmc::nvp;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_nvp(D, "realnvp_loss_grad")) return rc;
  if (!params || !grad || !loss || n < 0 || (n > 0 && (!x || !workspace)) ||
      workspace_bytes < flowmc_realnvp_loss_grad_workspace_bytes(D, n)) {
    flowmc_set_error("realnvp_loss_grad: null buffer or workspace too small");
    return FLOWMC_ERR_INVALID;
  }
  cudaMemsetAsync(grad, 0, (size_t)D->n_params * sizeof(float), stream);
  cudaMemsetAsync(loss, 0, sizeof(float), stream);
  if (n == 0) return FLOWMC_OK;
  float* layer_inputs = static_cast<float*>(workspace);
  float* logp = layer_inputs + pad4i((int64_t)(D->n_layers + 1) * n * D->n_features);
  if (int rc = launch_forward(*D, false, params, x, n, nullptr, logp, layer_inputs, PRE_WHITEN, POST_BASE_LOGP,
                              nullptr, Key{0, 0}, 1, idx, stream))
    return rc;
  const Smem S = smem_layout(*D, true);
  const size_t bytes = (size_t)S.total * sizeof(float);
  static size_t configured = 0;
  if (bytes > configured) {
    if (cudaFuncSetAttribute(realnvp_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) !=
        cudaSuccess) {
      cudaGetLastError();
      flowmc_set_error("realnvp_loss_grad: model too large for the shared-memory tile (n_features / n_hidden)");
      return FLOWMC_ERR_UNSUPPORTED;
    }
    configured = bytes;
  }
  realnvp_backward_kernel<<<(unsigned)((n + TM - 1) / TM), NT, bytes, stream>>>(*D, params, layer_inputs, logp, n,
                                                                               inv_n_total, grad, loss);
  flowmc_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    flowmc_set_error(cudaGetErrorString(e));
    return FLOWMC_ERR_CUDA;
  }
  return FLOWMC_OK;
}

}  // extern "C"
