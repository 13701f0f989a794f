// Tile machinery of the MaskedCouplingRQSpline kernels: a CTA keeps TM samples resident in shared
// memory and walks all coupling layers on them (fp32 parity path, CUDA cores).
//
// Reference: src/flowMC/resource/model/nf_model/rqSpline.py:450-488 (layer scan), common.py:109-112
// (MLP conditioner), :150-168 (masked coupling), :211-240 (ScalarAffine).
#pragma once
#include "flow_common.cuh"
#include "rng.cuh"

namespace flowmc {

constexpr int TM = 64;   // samples per CTA
constexpr int NT = 256;  // threads per CTA
constexpr int NW = NT / 32;

enum : int { PRE_NONE = 0, PRE_WHITEN = 1, PRE_NORMAL = 2 };
enum : int { POST_NONE = 0, POST_BASE_LOGP = 1, POST_UNWHITEN = 2 };

__host__ __device__ inline int round4(int v) { return (v + 3) & ~3; }

struct FlowSmem {  // offsets (floats) into dynamic shared memory
  int xs, xs_stride, a0, a1, a_stride, ldw, ld, total;
};
__host__ __device__ inline FlowSmem flow_smem_layout(const FlowmcFlowDesc& D) {
  FlowSmem s;
  int hmax = 4;
  for (int i = 1; i < D.n_linear; ++i) hmax = D.dims[i] > hmax ? D.dims[i] : hmax;
  s.xs_stride = round4(D.n_features) + 4;
  s.a_stride = round4(hmax) + 4;
  s.xs = 0;
  s.a0 = s.xs + TM * s.xs_stride;
  s.a1 = s.a0 + TM * s.a_stride;
  s.ldw = s.a1 + TM * s.a_stride;
  s.ld = s.ldw + NW * TM;
  s.total = s.ld + TM;
  return s;
}

// Gaussian.log_prob of one latent row with a diagonal covariance (common.py:285-286: multivariate_normal
// .logpdf, Cholesky form; the base covariance is c*I -- I at initialisation, shrunk uniformly by AdamW's
// weight decay afterwards)
__device__ __forceinline__ float base_log_prob(const FlowmcFlowDesc& D, const float* __restrict__ P,
                                               const float* __restrict__ y) {
  const int d = D.n_features;
  float q = 0.0f, logdiag = 0.0f;
  for (int j = 0; j < d; ++j) {
    const float L = sqrtf(P[D.off_base_cov + (int64_t)j * d + j]);
    const float r = (y[j] - P[D.off_base_mean + j]) / L;
    q += r * r;
    logdiag += logf(L);
  }
  return -0.5f * q - (float)d * 0.5f * 1.8378770664093453f - logdiag;
}

// host entry shared by the C-ABI functions (flow.cu)
int flow_transform(const FlowmcFlowDesc& D, bool inverse, const float* P, const float* x, int64_t n, float* y,
                   float* ld, float* layer_inputs, int pre, int post, const uint32_t* keys, Key hk, int64_t rpk,
                   cudaStream_t stream, const int32_t* idx = nullptr);

// tensor-core path (flow_tc.cu)
bool flow_tc_enabled(const FlowmcFlowDesc& D);
int tc_split_factor(const FlowmcFlowDesc& D, int64_t tiles);  // cluster size of the feature-split training kernels
void tc_split_note_max_clusters(int R, int n);
int flow_transform_tc(const FlowmcFlowDesc& D, bool inverse, const float* P, const float* x, int64_t n, float* y,
                      float* ld, int pre, int post, const uint32_t* keys, Key hk, int64_t rpk, cudaStream_t stream,
                      const int32_t* idx, float* save_x = nullptr, float* save_h = nullptr,
                      float* save_theta = nullptr, uint8_t* act_img = nullptr);
int flow_nf_propose_tc(const FlowmcFlowDesc& D, const float* P, Key subkey, const uint32_t* chain_keys,
                       int64_t chain_offset, int64_t n_chains, int n_steps, int n_batch, int n_sample, float* props,
                       float* lp_nf, cudaStream_t stream);

// out[s][n] = tanh(sum_k in[s][k] W[n][k] + b[n]) for the 64-sample tile; in/out in shared memory
__device__ __forceinline__ void dense_tanh_stage(const float* __restrict__ in_s, int in_stride, int Kdim,
                                                 const float* __restrict__ W, const float* __restrict__ b, int N,
                                                 float* __restrict__ out_s, int out_stride, bool mask_input,
                                                 int mask_parity) {
  constexpr int NB = 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* r0 = in_s + lane * in_stride;
  const float* r1 = in_s + (lane + 32) * in_stride;
  for (int n0 = warp * NB; n0 < N; n0 += NW * NB) {
    float acc0[NB], acc1[NB];
    const float* wrow[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int n = min(n0 + j, N - 1);
      acc0[j] = acc1[j] = __ldg(b + n);
      wrow[j] = W + (int64_t)n * Kdim;
    }
    if ((Kdim & 3) == 0) {
      // mask_input: the conditioner sees x * mask, mask True where (k + layer) % 2 == 1 (rqSpline.py:434)
      const float m_even = (!mask_input || (mask_parity & 1) == 1) ? 1.0f : 0.0f;  // components x, z
      const float m_odd = (!mask_input || (mask_parity & 1) == 0) ? 1.0f : 0.0f;   // components y, w
      for (int k = 0; k < Kdim; k += 4) {
        float4 a0 = *reinterpret_cast<const float4*>(r0 + k);
        float4 a1 = *reinterpret_cast<const float4*>(r1 + k);
        if (mask_input) {
          a0.x *= m_even; a0.z *= m_even; a0.y *= m_odd; a0.w *= m_odd;
          a1.x *= m_even; a1.z *= m_even; a1.y *= m_odd; a1.w *= m_odd;
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(wrow[j] + k));
          acc0[j] = fmaf(a0.x, w.x, acc0[j]); acc0[j] = fmaf(a0.y, w.y, acc0[j]);
          acc0[j] = fmaf(a0.z, w.z, acc0[j]); acc0[j] = fmaf(a0.w, w.w, acc0[j]);
          acc1[j] = fmaf(a1.x, w.x, acc1[j]); acc1[j] = fmaf(a1.y, w.y, acc1[j]);
          acc1[j] = fmaf(a1.z, w.z, acc1[j]); acc1[j] = fmaf(a1.w, w.w, acc1[j]);
        }
      }
    } else {
      // dimensions that are not a multiple of 4: scalar loads
      for (int k = 0; k < Kdim; ++k) {
        const bool keep = !mask_input || (((k + mask_parity) & 1) == 1);
        const float a0 = keep ? r0[k] : 0.0f;
        const float a1 = keep ? r1[k] : 0.0f;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const float w = __ldg(wrow[j] + k);
          acc0[j] = fmaf(a0, w, acc0[j]);
          acc1[j] = fmaf(a1, w, acc1[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      if (n0 + j < N) {
        out_s[lane * out_stride + n0 + j] = tanhf(acc0[j]);
        out_s[(lane + 32) * out_stride + n0 + j] = tanhf(acc1[j]);
      }
    }
  }
}


// Walks all coupling layers (forward: 0..L-1, inverse: L-1..0) over the tile in shared memory:
// xs [TM][xs_stride] is transformed in place, ld [TM] accumulates the log-determinant.
// Must be called by all NT threads; ends with a __syncthreads().
template <int K, bool INV>
__device__ __forceinline__ void flow_layers(const FlowmcFlowDesc& D, const float* __restrict__ P, const FlowSmem& S,
                                            float* smem, int64_t row0, int64_t n, float* __restrict__ layer_inputs) {
  constexpr int NP = 3 * K + 1;
  float* xs = smem + S.xs;
  float* a0 = smem + S.a0;
  float* a1 = smem + S.a1;
  float* ldw = smem + S.ldw;
  float* ld = smem + S.ld;
  const int d = D.n_features;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int li = 0; li < D.n_layers; ++li) {
    const int l = INV ? D.n_layers - 1 - li : li;
    const float* PL = P + (int64_t)l * D.layer_stride;
    const float scale = PL[D.off_scale], shift = PL[D.off_shift];
    if (layer_inputs != nullptr) {
      for (int i = tid; i < TM * d; i += NT) {
        const int s = i / d, j = i - s * d;
        if (row0 + s < n) layer_inputs[((int64_t)l * n + row0 + s) * d + j] = xs[s * S.xs_stride + j];
      }
    }
    // ScalarAffine under an all-False mask (rqSpline.py:435-436, common.py:224-240)
    {
      const float e = INV ? expf(-scale) : expf(scale);
      for (int i = tid; i < TM * d; i += NT) {
        const int s = i / d, j = i - s * d;
        float v = xs[s * S.xs_stride + j];
        v = INV ? v * e - shift : (v + shift) * e;
        xs[s * S.xs_stride + j] = v;
      }
      if (tid < TM) ld[tid] += INV ? -(float)d * scale : (float)d * scale;
    }
    __syncthreads();
    // conditioner MLP on x * mask (common.py:109-112,155)
    const float* in_s = xs;
    int in_stride = S.xs_stride;
    float* bufs[2] = {a0, a1};
    for (int i = 0; i < D.n_linear - 1; ++i) {
      float* out_s = bufs[i & 1];
      dense_tanh_stage(in_s, in_stride, D.dims[i], PL + D.off_W[i], PL + D.off_b[i], D.dims[i + 1], out_s,
                       S.a_stride, i == 0, l);
      __syncthreads();
      in_s = out_s;
      in_stride = S.a_stride;
    }
    // last linear + spline, one transformed feature per warp iteration
    {
      const int H = D.dims[D.n_linear - 1];
      const float* Wl = PL + D.off_W[D.n_linear - 1];
      const float* bl = PL + D.off_b[D.n_linear - 1];
      const float* h0 = in_s + lane * in_stride;
      const float* h1 = in_s + (lane + 32) * in_stride;
      float ldacc0 = 0.0f, ldacc1 = 0.0f;
      const int f0 = (l & 1);  // transformed features: (f + l) % 2 == 0
      for (int f = f0 + 2 * warp; f < d; f += 2 * NW) {
        float r0[NP], r1[NP];
        const float* wbase = Wl + (int64_t)f * NP * H;
#pragma unroll
        for (int r = 0; r < NP; ++r) r0[r] = r1[r] = __ldg(bl + f * NP + r);
        if ((H & 3) == 0) {
          for (int k = 0; k < H; k += 4) {
            const float4 u0 = *reinterpret_cast<const float4*>(h0 + k);
            const float4 u1 = *reinterpret_cast<const float4*>(h1 + k);
#pragma unroll
            for (int r = 0; r < NP; ++r) {
              const float4 w = __ldg(reinterpret_cast<const float4*>(wbase + (int64_t)r * H + k));
              r0[r] = fmaf(u0.x, w.x, r0[r]); r0[r] = fmaf(u0.y, w.y, r0[r]);
              r0[r] = fmaf(u0.z, w.z, r0[r]); r0[r] = fmaf(u0.w, w.w, r0[r]);
              r1[r] = fmaf(u1.x, w.x, r1[r]); r1[r] = fmaf(u1.y, w.y, r1[r]);
              r1[r] = fmaf(u1.z, w.z, r1[r]); r1[r] = fmaf(u1.w, w.w, r1[r]);
            }
          }
        } else {
          for (int k = 0; k < H; ++k) {
            const float u0 = h0[k], u1 = h1[k];
#pragma unroll
            for (int r = 0; r < NP; ++r) {
              const float w = __ldg(wbase + (int64_t)r * H + k);
              r0[r] = fmaf(u0, w, r0[r]);
              r1[r] = fmaf(u1, w, r1[r]);
            }
          }
        }
        float t;
        float* px = xs + lane * S.xs_stride + f;
        *px = rq_apply<K, INV>(r0, D.range_min, D.range_max, *px, t);
        ldacc0 += t;
        px = xs + (lane + 32) * S.xs_stride + f;
        *px = rq_apply<K, INV>(r1, D.range_min, D.range_max, *px, t);
        ldacc1 += t;
      }
      ldw[warp * TM + lane] = ldacc0;
      ldw[warp * TM + lane + 32] = ldacc1;
    }
    __syncthreads();
    if (tid < TM) {
      float s = 0.0f;
#pragma unroll
      for (int w = 0; w < NW; ++w) s += ldw[w * TM + tid];
      ld[tid] += s;
    }
    __syncthreads();
  }
}

}  // namespace flowmc
