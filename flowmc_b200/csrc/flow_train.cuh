// Pieces shared by the two backward kernels (flow_train.cu: fp32 CUDA cores, flow_train_tc.cu: tcgen05): the
// hand-derived reverse pass of the rational-quadratic spline and the host entry points of the tensor-core backward.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "flow_common.cuh"

namespace flowmc {

__device__ __forceinline__ float sigmoid_f(float t) { return 1.0f / (1.0f + expf(-t)); }

// Spline forward + reverse pass for one (sample, feature).
//   raw[3K+1]  conditioner output;  x  input;  Gy = dL/dy,  Gld = dL/dlogdet
//   -> gx = dL/dx (direct path),  draw[3K+1] = dL/draw.
// Same arithmetic as rq_params / rq_forward for everything that decides the bin.
// FAST = hardware ex2 / lg2 / rcp approximations (1-2 ulp, no slow-path branches) for the tensor-core kernel, whose
// epilogue is latency-bound; the gradients stay within ~1e-6 relative of the accurate version.
template <int K, bool FAST = false>
__device__ __forceinline__ void rq_backward(const float* raw, float rmin, float rmax, float x, float Gy, float Gld,
                                            float& gx, float* draw) {
  auto EXP = [](float v) { return FAST ? fast_exp(v) : expf(v); };
  auto RCP = [](float v) { return FAST ? fast_rcp(v) : 1.0f / v; };
  auto SOFTPLUS = [](float v) { return FAST ? fast_softplus(v) : softplus_f(v); };
  const float size = rmax - rmin;
  const float scale = size - (float)K * 1e-4f;
  const float offset = 0.5411666035652161f;
  float mw = raw[0], mh = raw[K];
#pragma unroll
  for (int i = 1; i < K; ++i) {
    mw = fmaxf(mw, raw[i]);
    mh = fmaxf(mh, raw[K + i]);
  }
  float pw[K], ph[K], sw = 0.0f, sh = 0.0f;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    pw[i] = EXP(raw[i] - mw);
    ph[i] = EXP(raw[K + i] - mh);
    sw += pw[i];
    sh += ph[i];
  }
  float xp[K + 1], yp[K + 1];
  xp[0] = rmin;
  yp[0] = rmin;
  float cx = 0.0f, cy = 0.0f;
  const float rsw = RCP(sw), rsh = RCP(sh);
#pragma unroll
  for (int i = 0; i < K; ++i) {
    pw[i] = FAST ? pw[i] * rsw : pw[i] / sw;
    ph[i] = FAST ? ph[i] * rsh : ph[i] / sh;
    if (i < K - 1) {
      const float bw = pw[i] * scale + 1e-4f;
      const float bh = ph[i] * scale + 1e-4f;
      cx = (i == 0) ? bw : cx + bw;
      cy = (i == 0) ? bh : cy + bh;
      xp[i + 1] = rmin + cx;
      yp[i + 1] = rmin + cy;
    }
  }
  xp[K] = rmax;
  yp[K] = rmax;
  // bin select (rqSpline.py:63-72: first bin if none)
  int kb = 0;
  float xl = xp[0], xr = xp[1], yl = yp[0], yr = yp[1], ul = raw[2 * K], ur = raw[2 * K + 1];
#pragma unroll
  for (int i = 1; i < K; ++i) {
    const bool in = (x >= xp[i]) && (x < xp[i + 1]);
    kb = in ? i : kb;
    xl = in ? xp[i] : xl; xr = in ? xp[i + 1] : xr;
    yl = in ? yp[i] : yl; yr = in ? yp[i + 1] : yr;
    ul = in ? raw[2 * K + i] : ul; ur = in ? raw[2 * K + i + 1] : ur;
  }
  const bool below = x <= xp[0], above = x >= xp[K];
  if (below) { ul = raw[2 * K]; }
  if (above) { ur = raw[3 * K]; }
  const float dl = SOFTPLUS(ul + offset) + 1e-4f, dr = SOFTPLUS(ur + offset) + 1e-4f;

  const float bw = xr - xl, bh = yr - yl;
  const float rbw = RCP(bw);
  const float s = FAST ? bh * rbw : bh / bw;
  float z = FAST ? (x - xl) * rbw : (x - xl) / bw;
  z = fminf(fmaxf(z, 0.0f), 1.0f);
  const float sq_z = z * z, z1mz = z - sq_z, omz = 1.0f - z, sq_1mz = omz * omz;
  const float st = dr + dl - 2.0f * s;
  const float nu = s * sq_z + dl * z1mz;  // num = bh * nu
  const float den = s + st * z1mz;
  const float q = dr * sq_z + 2.0f * s * z1mz + dl * sq_1mz;
  // ---- reverse pass, in-range branch --------------------------------------------------------
  const float rden = RCP(den);
  const float a_num = FAST ? Gy * rden : Gy / den;
  float a_den = FAST ? -Gy * (bh * nu) * (rden * rden) - 2.0f * Gld * rden
                     : -Gy * (bh * nu) / (den * den) - 2.0f * Gld / den;
  const float a_q = FAST ? Gld * RCP(q) : Gld / q;
  float a_s = (FAST ? 2.0f * Gld * RCP(s) : 2.0f * Gld / s) + a_q * 2.0f * z1mz + a_den * (1.0f - 2.0f * z1mz) +
              a_num * bh * sq_z;
  float a_dr = a_q * sq_z + a_den * z1mz;
  float a_dl = a_q * sq_1mz + a_den * z1mz + a_num * bh * z1mz;
  const float a_z = a_q * (2.0f * dr * z + 2.0f * s * (1.0f - 2.0f * z) - 2.0f * dl * omz) +
                    a_den * st * (1.0f - 2.0f * z) + a_num * bh * (2.0f * s * z + dl * (1.0f - 2.0f * z));
  float a_bh = FAST ? a_num * nu + a_s * rbw : a_num * nu + a_s / bw;
  float a_bw = FAST ? -(a_z * z + a_s * s) * rbw : -a_z * z / bw - a_s * s / bw;
  gx = FAST ? a_z * rbw : a_z / bw;
  float a_xl = -gx - a_bw;
  float a_xr = a_bw;
  float a_yl = Gy - a_bh;
  float a_yr = a_bh;
  // ---- linear tails (rqSpline.py:118-127): y = (x - x_e) d_e + y_e, logdet = log d_e ------------
  if (below || above) {
    const float de = below ? dl : dr;
    const float xe = below ? xp[0] : xp[K];
    gx = Gy * de;
    const float a_de = Gy * (x - xe) + (FAST ? Gld * RCP(de) : Gld / de);
    a_dl = below ? a_de : 0.0f;
    a_dr = above ? a_de : 0.0f;
    a_xl = a_xr = a_yl = a_yr = 0.0f;
  }
  const int kl = below ? 0 : (above ? -1 : kb);      // slope index receiving a_dl
  const int kr = above ? K : (below ? -1 : kb + 1);  // slope index receiving a_dr
  // ---- knots -> bin sizes -> softmax logits ------------------------------------------------------
  // knot j (1..K-1) = rmin + sum_{i<j} size_i ; knots 0 and K are constants (padding, rqSpline.py:329-333)
  const bool lk = (!below && !above) && kb >= 1;      // left knot is a function of the parameters
  const bool rk = (!below && !above) && kb + 1 <= K - 1;
  float dotw = 0.0f, doth = 0.0f;
  float apw[K], aph[K];
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const float sel_l = (lk && i < kb) ? 1.0f : 0.0f;
    const float sel_r = (rk && i < kb + 1) ? 1.0f : 0.0f;
    apw[i] = scale * (sel_l * a_xl + sel_r * a_xr);
    aph[i] = scale * (sel_l * a_yl + sel_r * a_yr);
    dotw += pw[i] * apw[i];
    doth += ph[i] * aph[i];
  }
#pragma unroll
  for (int i = 0; i < K; ++i) {
    draw[i] = pw[i] * (apw[i] - dotw);
    draw[K + i] = ph[i] * (aph[i] - doth);
  }
  // softplus'(u) = sigmoid(u): only the (at most) two slopes of the selected bin receive gradient
  const float gl = a_dl * (FAST ? fast_rcp(1.0f + fast_exp(-(ul + offset))) : sigmoid_f(ul + offset));
  const float gr = a_dr * (FAST ? fast_rcp(1.0f + fast_exp(-(ur + offset))) : sigmoid_f(ur + offset));
#pragma unroll
  for (int i = 0; i <= K; ++i) draw[2 * K + i] = (i == kl ? gl : 0.0f) + (i == kr ? gr : 0.0f);
}


// tensor-core backward (flow_train_tc.cu)
bool flow_backward_tc_supported(const FlowmcFlowDesc& D);
int64_t flow_backward_tc_wimg_bytes(const FlowmcFlowDesc& D);
int64_t flow_backward_tc_act_bytes(const FlowmcFlowDesc& D, int64_t n);
void flow_backward_tc_set_timing(long long* buf);
int64_t flow_backward_tc_partial_bytes(const FlowmcFlowDesc& D, int64_t n);
int flow_backward_tc(const FlowmcFlowDesc& D, const float* params, uint8_t* wimg, const uint8_t* act_img,
                     const float* save_x, const float* save_theta, const float* logp, int64_t n, float inv_n,
                     float* grad, float* loss, float* partial, cudaStream_t stream);

}  // namespace flowmc
