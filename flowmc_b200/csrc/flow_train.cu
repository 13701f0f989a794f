// Flow training: loss + analytic gradient of -mean(log_prob) and the fused optimiser step.
//
// Reference: NFModel.loss_fn / train_step (src/flowMC/resource/model/nf_model/base.py:98-125, reverse-mode
// autodiff of MaskedCouplingRQSpline.log_prob) and Optimizer = optax.chain(clip_by_global_norm(1.0),
// adamw(lr, b1=momentum)) (src/flowMC/resource/optimizer.py:19-23).
//
// flowmc_flow_loss_grad = two launches:
//   1. the forward kernel of flow.cu (PRE_WHITEN, POST_BASE_LOGP) which also stores every layer's input and
//      the final latent (L2-resident scratch);
//   2. flow_backward_kernel: one CTA per 64-sample tile walks the layers in reverse with the tile's
//      gradient resident in shared memory.  Per layer it recomputes the conditioner activations, then
//        (a) one warp per transformed feature: theta_f = W3[f] h + b3[f], spline forward AND hand-derived
//            reverse pass in registers -> d(loss)/d(theta_f) (softmax / cumsum / softplus / bin-select
//            adjoints), d(loss)/dx_f;
//        (b) dh += dtheta W3 (register accumulators per warp-owned column block);
//        (c) dW3[f] = dtheta_f^T h  (reduction over the tile's samples) -> fp32 atomics on the flat grad;
//      then the two tanh layers' dW / db / dh the same way, the masked-coupling and ScalarAffine adjoints.
//   Weight gradients are reduced over the batch with red.global.add.f32 (order is not deterministic;
//   results agree with float64 autograd to ~1e-6 relative).
//
// flowmc_clip_adamw = two launches: per-block partial sums of g^2 (fixed order), then every block of the
// update kernel re-reduces the partials in the same fixed order (deterministic global norm) and applies
// clip + Adam moments + bias correction + decoupled weight decay + learning rate in one pass over the
// flat parameter / moment vectors.
#include <cmath>
#include <cstdlib>
#include <string>

#include "flow_tile.cuh"
#include "flow_train.cuh"
#include "registry.h"

namespace flowmc {

struct TrainSmem {  // offsets in floats
  int xa, g, xs_stride, h[FLOWMC_FLOW_MAX_LINEAR], a_stride, dact0, dact1, dth, dth_stride, fc, red, total;
};

__host__ __device__ inline TrainSmem train_smem_layout(const FlowmcFlowDesc& D, int fc) {
  TrainSmem s;
  int hmax = 4;
  for (int i = 1; i < D.n_linear; ++i) hmax = D.dims[i] > hmax ? D.dims[i] : hmax;
  s.xs_stride = round4(D.n_features) + 4;
  s.a_stride = round4(hmax) + 4;
  int o = 0;
  s.xa = o; o += TM * s.xs_stride;
  s.g = o; o += TM * s.xs_stride;
  for (int i = 0; i < D.n_linear - 1; ++i) { s.h[i] = o; o += TM * s.a_stride; }
  s.dact0 = o; o += TM * s.a_stride;
  s.dact1 = o; o += TM * s.a_stride;
  s.fc = fc;
  s.dth_stride = fc * round4(3 * D.num_bins + 1) + 4;
  s.dth = o; o += TM * s.dth_stride;
  s.red = o; o += 2 * NW;
  s.total = o;
  return s;
}

// dW[n][k] += sum_s A[s][n] * B[s][k]  and  db[n] += sum_s A[s][n]   (reduction over the tile's samples).
// A, B in shared memory ([TM][stride] rows); results go to global memory with atomics.
// mask_parity >= 0: B is the layer input and only the conditioning columns ((k + parity) odd) are non-zero.
__device__ __forceinline__ void weight_grad_stage(const float* __restrict__ A, int a_stride, int N,
                                                  const float* __restrict__ B, int b_stride, int Kd,
                                                  float* __restrict__ dW, float* __restrict__ db, int mask_parity) {
  constexpr int RB = 8;  // rows of dW per warp pass
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int n0 = warp * RB; n0 < N; n0 += NW * RB) {
    for (int k0 = 0; k0 < Kd; k0 += 128) {
      float acc[RB][4];
#pragma unroll
      for (int r = 0; r < RB; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;
      float bsum = 0.0f;  // lane r < RB accumulates db[n0 + r] (first column pass only)
      for (int s = 0; s < TM; ++s) {
        float a[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) a[r] = (n0 + r < N) ? A[s * a_stride + n0 + r] : 0.0f;
        float b[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int k = k0 + lane + 32 * c;
          b[c] = (k < Kd) ? B[s * b_stride + k] : 0.0f;
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
          if (lane == r) bsum += a[r];
        }
      }
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        if (n0 + r < N) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int k = k0 + lane + 32 * c;
            if (k < Kd && (mask_parity < 0 || ((k + mask_parity) & 1) == 1))
              atomicAdd(dW + (int64_t)(n0 + r) * Kd + k, acc[r][c]);
          }
        }
      }
      if (k0 == 0 && lane < RB && n0 + lane < N) atomicAdd(db + n0 + lane, bsum);
    }
  }
}

// out[s][k] = sum_n in[s][n] * W[n][k]   (back-propagation through a Linear: dh = da W), tile in smem.
__device__ __forceinline__ void dense_back_stage(const float* __restrict__ in_s, int in_stride, int N,
                                                 const float* __restrict__ W, int Kd, float* __restrict__ out_s,
                                                 int out_stride) {
  constexpr int KB = 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* r0 = in_s + lane * in_stride;
  const float* r1 = in_s + (lane + 32) * in_stride;
  for (int k0 = warp * KB; k0 < Kd; k0 += NW * KB) {
    float acc0[KB], acc1[KB];
#pragma unroll
    for (int j = 0; j < KB; ++j) acc0[j] = acc1[j] = 0.0f;
    if ((Kd & 3) == 0 && k0 + KB <= Kd) {
      for (int n = 0; n < N; ++n) {
        const float a0 = r0[n], a1 = r1[n];
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(W + (int64_t)n * Kd + k0));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(W + (int64_t)n * Kd + k0 + 4));
        acc0[0] = fmaf(a0, w0.x, acc0[0]); acc0[1] = fmaf(a0, w0.y, acc0[1]);
        acc0[2] = fmaf(a0, w0.z, acc0[2]); acc0[3] = fmaf(a0, w0.w, acc0[3]);
        acc0[4] = fmaf(a0, w1.x, acc0[4]); acc0[5] = fmaf(a0, w1.y, acc0[5]);
        acc0[6] = fmaf(a0, w1.z, acc0[6]); acc0[7] = fmaf(a0, w1.w, acc0[7]);
        acc1[0] = fmaf(a1, w0.x, acc1[0]); acc1[1] = fmaf(a1, w0.y, acc1[1]);
        acc1[2] = fmaf(a1, w0.z, acc1[2]); acc1[3] = fmaf(a1, w0.w, acc1[3]);
        acc1[4] = fmaf(a1, w1.x, acc1[4]); acc1[5] = fmaf(a1, w1.y, acc1[5]);
        acc1[6] = fmaf(a1, w1.z, acc1[6]); acc1[7] = fmaf(a1, w1.w, acc1[7]);
      }
    } else {
      for (int n = 0; n < N; ++n) {
        const float a0 = r0[n], a1 = r1[n];
#pragma unroll
        for (int j = 0; j < KB; ++j) {
          const float w = (k0 + j < Kd) ? __ldg(W + (int64_t)n * Kd + k0 + j) : 0.0f;
          acc0[j] = fmaf(a0, w, acc0[j]);
          acc1[j] = fmaf(a1, w, acc1[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < KB; ++j) {
      if (k0 + j < Kd) {
        out_s[lane * out_stride + k0 + j] = acc0[j];
        out_s[(lane + 32) * out_stride + k0 + j] = acc1[j];
      }
    }
  }
}

template <int K>
__global__ void __launch_bounds__(NT) flow_backward_kernel(const FlowmcFlowDesc D, const float* __restrict__ P,
                                                           const float* __restrict__ layer_inputs,
                                                           const float* __restrict__ logp, int64_t n, float inv_n,
                                                           int fc, float* __restrict__ grad,
                                                           float* __restrict__ loss,
                                                           const float* __restrict__ saved_h,
                                                           const float* __restrict__ saved_theta) {
  extern __shared__ __align__(16) float smem[];
  constexpr int NP = 3 * K + 1;
  constexpr int NP4 = (NP + 3) & ~3;
  const TrainSmem S = train_smem_layout(D, fc);
  float* xa = smem + S.xa;
  float* g = smem + S.g;
  float* dact[2] = {smem + S.dact0, smem + S.dact1};
  float* dth = smem + S.dth;
  float* red = smem + S.red;
  const int d = D.n_features;
  const int n_lin = D.n_linear;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * TM;
  const int n_valid = (int)min((int64_t)TM, n - row0);

  // ---- loss contribution and the gradient w.r.t. the final latent -------------------------------
  // loss = -mean(logdet + base.log_prob(y)):  dL/dy = (y - mean) / cov_jj / n,  dL/dlogdet = -1/n
  if (warp == 0) {
    float v = 0.0f;
    if (lane < n_valid) v += logp[row0 + lane];
    if (lane + 32 < n_valid) v += logp[row0 + lane + 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) atomicAdd(loss, -v * inv_n);
  }
  for (int i = tid; i < TM * d; i += NT) {
    const int s = i / d, j = i - s * d;
    float v = 0.0f;
    if (s < n_valid) {
      const float y = layer_inputs[((int64_t)D.n_layers * n + row0 + s) * d + j];
      v = inv_n * (y - P[D.off_base_mean + j]) / P[D.off_base_cov + (int64_t)j * d + j];
    }
    g[s * S.xs_stride + j] = v;
  }
  const float gld0 = (lane < n_valid) ? -inv_n : 0.0f;
  const float gld1 = (lane + 32 < n_valid) ? -inv_n : 0.0f;
  __syncthreads();

  for (int l = D.n_layers - 1; l >= 0; --l) {
    const float* PL = P + (int64_t)l * D.layer_stride;
    float* GL = grad + (int64_t)l * D.layer_stride;
    const float scale = PL[D.off_scale], shift = PL[D.off_shift];
    const float e = expf(scale);
    // ---- recompute the layer's forward activations ---------------------------------------------
    for (int i = tid; i < TM * d; i += NT) {
      const int s = i / d, j = i - s * d;
      const int64_t r = min(row0 + s, n - 1);
      xa[s * S.xs_stride + j] = (layer_inputs[((int64_t)l * n + r) * d + j] + shift) * e;
    }
    __syncthreads();
    if (saved_h != nullptr) {
      // activations kept by the tensor-core forward pass ([L][n_hidden][128][n], sample-contiguous)
      for (int i = 0; i < n_lin - 1; ++i) {
        float* out_s = smem + S.h[i];
        const int N = D.dims[i + 1];
        const float* src = saved_h + ((int64_t)(l * (n_lin - 1) + i) * 128) * n + row0;
        for (int e = tid; e < N * TM; e += NT) {
          const int k = e / TM, s = e - k * TM;
          out_s[s * S.a_stride + k] = (s < n_valid) ? src[(int64_t)k * n + s] : 0.0f;
        }
      }
      __syncthreads();
    } else {
      const float* in_s = xa;
      int in_stride = S.xs_stride;
      for (int i = 0; i < n_lin - 1; ++i) {
        float* out_s = smem + S.h[i];
        dense_tanh_stage(in_s, in_stride, D.dims[i], PL + D.off_W[i], PL + D.off_b[i], D.dims[i + 1], out_s,
                         S.a_stride, i == 0, l);
        __syncthreads();
        in_s = out_s;
        in_stride = S.a_stride;
      }
    }
    // ---- last linear + spline: chunks of `fc` transformed features -------------------------------
    const int H = D.dims[n_lin - 1];
    const float* hl = smem + S.h[n_lin - 2];
    const float* Wl = PL + D.off_W[n_lin - 1];
    const float* bl = PL + D.off_b[n_lin - 1];
    float* dWl = GL + D.off_W[n_lin - 1];
    float* dbl = GL + D.off_b[n_lin - 1];
    const int f0 = (l & 1);
    const int n_tf = (d - f0 + 1) / 2;  // transformed features f0, f0+2, ...
    float* dh = dact[0];
    for (int c0 = 0; c0 < n_tf; c0 += fc) {
      const int nc = min(fc, n_tf - c0);
      // (a) one warp per feature of the chunk
      if (warp < nc) {
        const int f = f0 + 2 * (c0 + warp);
        const float* h0 = hl + lane * S.a_stride;
        const float* h1 = hl + (lane + 32) * S.a_stride;
        float r0[NP], r1[NP];
        const float* wbase = Wl + (int64_t)f * NP * H;
        if (saved_theta != nullptr) {
          // spline parameters kept by the forward pass ([L][ceil(d/2) * NP][n], sample-contiguous): no GEMM
          const float* src = saved_theta + ((int64_t)l * ((d + 1) / 2) + (c0 + warp)) * NP * n + row0;
          const bool v0 = lane < n_valid, v1 = lane + 32 < n_valid;
#pragma unroll
          for (int r = 0; r < NP; ++r) {
            r0[r] = v0 ? src[(int64_t)r * n + lane] : 0.0f;
            r1[r] = v1 ? src[(int64_t)r * n + lane + 32] : 0.0f;
          }
        } else if ((H & 3) == 0) {
#pragma unroll
          for (int r = 0; r < NP; ++r) r0[r] = r1[r] = __ldg(bl + f * NP + r);
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
#pragma unroll
          for (int r = 0; r < NP; ++r) r0[r] = r1[r] = __ldg(bl + f * NP + r);
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
        float dr[NP], gx;
        float* gp = g + lane * S.xs_stride + f;
        rq_backward<K>(r0, D.range_min, D.range_max, xa[lane * S.xs_stride + f], *gp, gld0, gx, dr);
        *gp = gx;
        float* o = dth + lane * S.dth_stride + warp * NP4;
#pragma unroll
        for (int r = 0; r < NP; ++r) o[r] = dr[r];
        gp = g + (lane + 32) * S.xs_stride + f;
        rq_backward<K>(r1, D.range_min, D.range_max, xa[(lane + 32) * S.xs_stride + f], *gp, gld1, gx, dr);
        *gp = gx;
        o = dth + (lane + 32) * S.dth_stride + warp * NP4;
#pragma unroll
        for (int r = 0; r < NP; ++r) o[r] = dr[r];
      }
      __syncthreads();
      // (b) dh[s][k] += sum_{f in chunk, r} dtheta[s][f][r] * W3[f*NP + r][k]; warp owns column blocks
      for (int k0 = warp * 16; k0 < H; k0 += NW * 16) {
        float a0[16], a1[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const bool ok = (c0 > 0) && (k0 + j < H);
          a0[j] = ok ? dh[lane * S.a_stride + k0 + j] : 0.0f;
          a1[j] = ok ? dh[(lane + 32) * S.a_stride + k0 + j] : 0.0f;
        }
        const bool fast = ((H & 3) == 0) && (k0 + 16 <= H);
        for (int fi = 0; fi < nc; ++fi) {
          const int f = f0 + 2 * (c0 + fi);
          const float* t0 = dth + lane * S.dth_stride + fi * NP4;
          const float* t1 = dth + (lane + 32) * S.dth_stride + fi * NP4;
          const float* wrow = Wl + (int64_t)f * NP * H + k0;
#pragma unroll 5
          for (int r = 0; r < NP; ++r) {
            const float v0 = t0[r], v1 = t1[r];
            if (fast) {
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(wrow + (int64_t)r * H + 4 * j4));
                a0[4 * j4] = fmaf(v0, w.x, a0[4 * j4]); a0[4 * j4 + 1] = fmaf(v0, w.y, a0[4 * j4 + 1]);
                a0[4 * j4 + 2] = fmaf(v0, w.z, a0[4 * j4 + 2]); a0[4 * j4 + 3] = fmaf(v0, w.w, a0[4 * j4 + 3]);
                a1[4 * j4] = fmaf(v1, w.x, a1[4 * j4]); a1[4 * j4 + 1] = fmaf(v1, w.y, a1[4 * j4 + 1]);
                a1[4 * j4 + 2] = fmaf(v1, w.z, a1[4 * j4 + 2]); a1[4 * j4 + 3] = fmaf(v1, w.w, a1[4 * j4 + 3]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float w = (k0 + j < H) ? __ldg(wrow + (int64_t)r * H + j) : 0.0f;
                a0[j] = fmaf(v0, w, a0[j]);
                a1[j] = fmaf(v1, w, a1[j]);
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (k0 + j < H) {
            dh[lane * S.a_stride + k0 + j] = a0[j];
            dh[(lane + 32) * S.a_stride + k0 + j] = a1[j];
          }
        }
      }
      // (c) dW3[f][r][k] += sum_s dtheta[s][f][r] * h[s][k];  db3[f][r] += sum_s dtheta[s][f][r]
      if (warp < nc) {
        const int f = f0 + 2 * (c0 + warp);
        for (int k0 = 0; k0 < H; k0 += 128) {
          float acc[NP][4];
#pragma unroll
          for (int r = 0; r < NP; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;
          for (int s = 0; s < TM; ++s) {
            const float* t = dth + s * S.dth_stride + warp * NP4;
            float b[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int k = k0 + lane + 32 * c;
              b[c] = (k < H) ? hl[s * S.a_stride + k] : 0.0f;
            }
#pragma unroll
            for (int r = 0; r < NP; ++r) {
              const float v = t[r];
#pragma unroll
              for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(v, b[c], acc[r][c]);
            }
          }
#pragma unroll
          for (int r = 0; r < NP; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int k = k0 + lane + 32 * c;
              if (k < H) atomicAdd(dWl + ((int64_t)f * NP + r) * H + k, acc[r][c]);
            }
        }
        if (lane < NP) {
          float bs = 0.0f;
          for (int s = 0; s < TM; ++s) bs += dth[s * S.dth_stride + warp * NP4 + lane];
          atomicAdd(dbl + f * NP + lane, bs);
        }
        if (NP > 32 && lane + 32 < NP) {
          float bs = 0.0f;
          for (int s = 0; s < TM; ++s) bs += dth[s * S.dth_stride + warp * NP4 + lane + 32];
          atomicAdd(dbl + f * NP + lane + 32, bs);
        }
      }
      __syncthreads();
    }
    // ---- tanh layers in reverse: da = dh * (1 - h^2); dW, db; dh_prev = da W -----------------------
    int cur = 0;
    for (int j = n_lin - 2; j >= 0; --j) {
      const int N = D.dims[j + 1];
      const float* hj = smem + S.h[j];
      float* da = dact[cur];
      for (int i = tid; i < TM * N; i += NT) {
        const int s = i / N, k = i - s * N;
        const float hv = hj[s * S.a_stride + k];
        da[s * S.a_stride + k] *= (1.0f - hv * hv);
      }
      __syncthreads();
      const float* in_s = (j == 0) ? xa : smem + S.h[j - 1];
      const int in_stride = (j == 0) ? S.xs_stride : S.a_stride;
      weight_grad_stage(da, S.a_stride, N, in_s, in_stride, D.dims[j], GL + D.off_W[j], GL + D.off_b[j],
                        j == 0 ? l : -1);
      dense_back_stage(da, S.a_stride, N, PL + D.off_W[j], D.dims[j], dact[cur ^ 1], S.a_stride);
      __syncthreads();
      cur ^= 1;
    }
    // ---- masked coupling: conditioning features receive the conditioner's input gradient ----------
    {
      const float* dc = dact[cur];
      float ssc = 0.0f, ssh = 0.0f;
      for (int i = tid; i < TM * d; i += NT) {
        const int s = i / d, j = i - s * d;
        float ga = g[s * S.xs_stride + j];
        if (((j + l) & 1) == 1) ga += dc[s * S.a_stride + j];
        // ScalarAffine: x_a = (x + shift) e^scale ; logdet += d * scale
        ssc += ga * xa[s * S.xs_stride + j];
        ssh += ga * e;
        g[s * S.xs_stride + j] = ga * e;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ssc += __shfl_xor_sync(0xffffffffu, ssc, o);
        ssh += __shfl_xor_sync(0xffffffffu, ssh, o);
      }
      if (lane == 0) {
        red[warp] = ssc;
        red[NW + warp] = ssh;
      }
      __syncthreads();
      if (tid == 0) {
        float a = 0.0f, b = 0.0f;
        for (int w = 0; w < NW; ++w) {
          a += red[w];
          b += red[NW + w];
        }
        atomicAdd(GL + D.off_scale, a - inv_n * (float)d * (float)n_valid);
        atomicAdd(GL + D.off_shift, b);
      }
      __syncthreads();
    }
  }
}

template <int K>
static int launch_backward(const FlowmcFlowDesc& D, const float* P, const float* layer_inputs, const float* logp,
                           int64_t n, float inv_n, float* grad, float* loss, const float* saved_h,
                           const float* saved_theta, cudaStream_t stream) {
  auto kern = flow_backward_kernel<K>;
  static int max_smem = 0;
  if (max_smem == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  }
  int fc = NW;
  while (fc > 1 && (size_t)train_smem_layout(D, fc).total * sizeof(float) > (size_t)max_smem) --fc;
  const size_t bytes = (size_t)train_smem_layout(D, fc).total * sizeof(float);
  if (bytes > (size_t)max_smem) {
    flowmc_set_error("flow_loss_grad: model too large for the shared-memory tile (hidden width / n_features)");
    return FLOWMC_ERR_UNSUPPORTED;
  }
  static size_t configured = 0;
  if (bytes > configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
      flowmc_set_error("flow_loss_grad: cannot configure shared memory");
      return FLOWMC_ERR_CUDA;
    }
    configured = bytes;
  }
  kern<<<(unsigned)((n + TM - 1) / TM), NT, bytes, stream>>>(D, P, layer_inputs, logp, n, inv_n, fc, grad, loss,
                                                             saved_h, saved_theta);
  flowmc_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    flowmc_set_error(cudaGetErrorString(e));
    return FLOWMC_ERR_CUDA;
  }
  return FLOWMC_OK;
}

// ---- optimiser --------------------------------------------------------------------------------
constexpr int kNormBlocks = 256;

__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, int64_t n,
                                                            float* __restrict__ partial) {
  __shared__ float sh[8];
  float v = 0.0f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) v = fmaf(g[i], g[i], v);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.0f;
    for (int w = 0; w < 8; ++w) s += sh[w];
    partial[blockIdx.x] = s;
  }
}

struct AdamArgs {
  float lr, b1, b2, omb1, omb2, eps, wd, max_norm, bc1, bc2;  // omb = 1 - b (rounded from double), bc = 1 - b^count
};

__global__ void __launch_bounds__(256) clip_adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                         float* __restrict__ mu, float* __restrict__ nu, int64_t n,
                                                         const float* __restrict__ partial, AdamArgs a,
                                                         float* __restrict__ gnorm_out) {
  __shared__ float s_norm;
  if (threadIdx.x < 32) {
    float v = 0.0f;
    for (int i = threadIdx.x; i < kNormBlocks; i += 32) v += partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) {
      s_norm = sqrtf(v);
      if (blockIdx.x == 0 && gnorm_out != nullptr) *gnorm_out = s_norm;
    }
  }
  __syncthreads();
  const float gn = s_norm;
  const bool keep = gn < a.max_norm;  // optax.clip_by_global_norm: where(g_norm < max_norm, g, (g / g_norm) * max_norm)
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    float gi = g[i];
    if (!keep) gi = (gi / gn) * a.max_norm;
    const float m = a.omb1 * gi + a.b1 * mu[i];                  // optax.scale_by_adam / update_moment
    const float v = a.omb2 * (gi * gi) + a.b2 * nu[i];
    mu[i] = m;
    nu[i] = v;
    float u = (m / a.bc1) / (sqrtf(v / a.bc2) + a.eps);
    const float pi = p[i];
    u = u + a.wd * pi;                                           // optax.add_decayed_weights
    p[i] = pi + (-a.lr) * u;                                     // scale_by_learning_rate, apply_updates
  }
}

}  // namespace flowmc

extern "C" {

static int64_t pad4i(int64_t v) { return (v + 3) & ~(int64_t)3; }

int64_t flowmc_flow_loss_grad_workspace_bytes(const FlowmcFlowDesc* D, int64_t n) {
  if (!D || n <= 0) return 0;
  const int64_t NP = 3 * D->num_bins + 1;
  // layer inputs + final latent, log-probs, and (tensor-core forward) the hidden activations and spline parameters
  int64_t b = 4 * (pad4i((int64_t)(D->n_layers + 1) * n * D->n_features) + pad4i(n) +
                   pad4i((int64_t)D->n_layers * (D->n_linear - 1) * 128 * n) +
                   pad4i((int64_t)D->n_layers * ((D->n_features + 1) / 2) * NP * n));
  if (flowmc::flow_backward_tc_supported(*D))  // transposed weight image + per-tile activation images
    b += 2048 + ((flowmc::flow_backward_tc_wimg_bytes(*D) + 1023) & ~(int64_t)1023) +
         flowmc::flow_backward_tc_act_bytes(*D, n) + flowmc::flow_backward_tc_partial_bytes(*D, n);
  return b;
}

int flowmc_flow_loss_grad(const FlowmcFlowDesc* D, const float* params, const float* x, const int32_t* idx,
                          int64_t n, float inv_n_total, float* grad, float* loss, void* workspace,
                          int64_t workspace_bytes, void* stream_) {
  using namespace flowmc;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!D || D->n_features < 1 || D->n_layers < 1 || D->n_linear < 2 || D->n_linear > FLOWMC_FLOW_MAX_LINEAR) {
    flowmc_set_error("flow_loss_grad: bad flow descriptor");
    return FLOWMC_ERR_INVALID;
  }
  if (!params || !grad || !loss || n < 0 || (n > 0 && (!x || !workspace)) ||
      workspace_bytes < flowmc_flow_loss_grad_workspace_bytes(D, n)) {
    flowmc_set_error("flow_loss_grad: null buffer or workspace too small");
    return FLOWMC_ERR_INVALID;
  }
  cudaMemsetAsync(grad, 0, (size_t)D->n_params * sizeof(float), stream);
  cudaMemsetAsync(loss, 0, sizeof(float), stream);
  if (n == 0) return FLOWMC_OK;
  const int64_t NP = 3 * D->num_bins + 1;
  float* layer_inputs = static_cast<float*>(workspace);
  float* logp = layer_inputs + pad4i((int64_t)(D->n_layers + 1) * n * D->n_features);
  float* save_h = logp + pad4i(n);
  float* save_theta = save_h + pad4i((int64_t)D->n_layers * (D->n_linear - 1) * 128 * n);
  // (the training kernels address a feature's [NP][n] block of spline parameters with 32-bit element offsets: batches
  // beyond 2^26 rows take the CUDA-core path)
  const bool tcf = flow_tc_enabled(*D) && n <= ((int64_t)1 << 26);
  // The tensor-core forward can also hand its hidden activations and spline parameters to the backward kernel
  // (no conditioner recompute: -27 % instructions).  Measured on B200 (profiles/r01_flow_backward_c4_ncu.txt) the
  // CUDA-core backward is bound by the L2 latency of its weight loads, not by instruction count, and the extra
  // 430 MB/step of activation traffic makes it 12 % SLOWER -- so recompute stays the default; the hand-off is kept
  // (FLOWMC_BWD_SAVED=1, covered by tests) as the interface a tcgen05 backward will use.
  static const bool use_saved = [] {
    const char* e = std::getenv("FLOWMC_BWD_SAVED");
    return e != nullptr && e[0] == '1';
  }();
  static const bool bwd_tc = [] {
    const char* e = std::getenv("FLOWMC_BWD_TC");
    return e == nullptr || e[0] != '0';
  }();
  if (tcf && bwd_tc && flow_backward_tc_supported(*D)) {
    // tensor-core forward AND backward: the forward leaves the spline parameters and the packed activation
    // images behind, the backward (flow_train_tc.cu) runs every data / weight gradient GEMM on tcgen05
    uint8_t* wimg = reinterpret_cast<uint8_t*>(save_theta + pad4i((int64_t)D->n_layers * ((D->n_features + 1) / 2) * NP * n));
    wimg += (1024 - (reinterpret_cast<uintptr_t>(wimg) & 1023)) & 1023;
    uint8_t* act_img = wimg + ((flow_backward_tc_wimg_bytes(*D) + 1023) & ~(int64_t)1023);
    if (int rc = flow_transform_tc(*D, false, params, x, n, nullptr, logp, PRE_WHITEN, POST_BASE_LOGP, nullptr,
                                   Key{0, 0}, 1, stream, idx, layer_inputs, nullptr, save_theta, act_img))
      return rc;
    float* partial = reinterpret_cast<float*>(act_img + ((flow_backward_tc_act_bytes(*D, n) + 1023) & ~(int64_t)1023));
    return flow_backward_tc(*D, params, wimg, act_img, layer_inputs, save_theta, logp, n, inv_n_total, grad, loss,
                            partial, stream);
  }
  if (tcf) {
    if (!use_saved) {
      save_h = nullptr;
      save_theta = nullptr;
    }
    if (int rc = flow_transform_tc(*D, false, params, x, n, nullptr, logp, PRE_WHITEN, POST_BASE_LOGP, nullptr,
                                   Key{0, 0}, 1, stream, idx, layer_inputs, save_h, save_theta))
      return rc;
  } else {
    if (int rc = flow_transform(*D, false, params, x, n, nullptr, logp, layer_inputs, PRE_WHITEN, POST_BASE_LOGP,
                                nullptr, Key{0, 0}, 1, stream, idx))
      return rc;
    save_h = nullptr;
    save_theta = nullptr;
  }
  switch (D->num_bins) {
    case 4:
      return launch_backward<4>(*D, params, layer_inputs, logp, n, inv_n_total, grad, loss, save_h, save_theta, stream);
    case 8:
      return launch_backward<8>(*D, params, layer_inputs, logp, n, inv_n_total, grad, loss, save_h, save_theta, stream);
    case 16:
      return launch_backward<16>(*D, params, layer_inputs, logp, n, inv_n_total, grad, loss, save_h, save_theta, stream);
    default:
      flowmc_set_error("flow: num_bins must be 4, 8 or 16");
      return FLOWMC_ERR_UNSUPPORTED;
  }
}

int flowmc_clip_adamw(int64_t n_params, float* params, const float* grads, float* mu, float* nu, int64_t count,
                      double lr, double b1, double b2, double eps, double weight_decay, double max_norm,
                      float* scratch, float* gnorm_out, void* stream_) {
  using namespace flowmc;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_params < 0 || count < 1 || (n_params > 0 && (!params || !grads || !mu || !nu || !scratch))) {
    flowmc_set_error("clip_adamw: bad arguments (count is the 1-based step number)");
    return FLOWMC_ERR_INVALID;
  }
  if (n_params == 0) return FLOWMC_OK;
  sumsq_partial_kernel<<<kNormBlocks, 256, 0, stream>>>(grads, n_params, scratch);
  flowmc_count_launch();
  AdamArgs a;
  // optax works with Python-float hyperparameters: (1 - decay) is formed in double and only then rounded to
  // float32 by the multiply; decay ** count is a float32 power of the float32-rounded decay
  a.lr = (float)lr; a.b1 = (float)b1; a.b2 = (float)b2; a.eps = (float)eps; a.wd = (float)weight_decay;
  a.max_norm = (float)max_norm;
  a.omb1 = (float)(1.0 - b1);
  a.omb2 = (float)(1.0 - b2);
  a.bc1 = 1.0f - powf(a.b1, (float)count);
  a.bc2 = 1.0f - powf(a.b2, (float)count);
  int64_t blocks = (n_params + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  clip_adamw_kernel<<<(unsigned)blocks, 256, 0, stream>>>(params, grads, mu, nu, n_params, scratch, a, gnorm_out);
  flowmc_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    flowmc_set_error(cudaGetErrorString(e));
    return FLOWMC_ERR_CUDA;
  }
  return FLOWMC_OK;
}

}  // extern "C"
