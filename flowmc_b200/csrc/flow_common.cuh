// Shared pieces of the MaskedCouplingRQSpline kernels: parameter-blob descriptor, rational-
// quadratic spline math (forward / inverse / parameter normalisation) on registers.
//
// Reference (paths under src/flowMC/resource/model/): nf_model/rqSpline.py:20-39 (bin / slope
// normalisation), :42-128 (spline forward), :131-155 (_safe_quadratic_root), :158-239 (spline
// inverse), :310-338 (get_params); common.py:150-168 (MaskedCouplingLayer), :211-240 (ScalarAffine).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/flowmc_b200.h"

namespace flowmc {

constexpr int kMaxBins = 16;

struct RQ {  // normalised spline parameters of one (sample, feature): K+1 knots
  float xp[kMaxBins + 1], yp[kMaxBins + 1], sl[kMaxBins + 1];
};

__device__ __forceinline__ float softplus_f(float x) {  // jnp.logaddexp(x, 0)
  return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x)));
}

// RQSpline.get_params for one feature: raw[3K+1] -> knots and slopes (rqSpline.py:310-338)
template <int K>
__device__ __forceinline__ void rq_params(const float* raw, float rmin, float rmax, RQ& q) {
  const float size = rmax - rmin;
  const float scale = size - (float)K * 1e-4f;
  const float offset = 0.5411666035652161f;  // log(exp(1 - 1e-4) - 1) evaluated in fp32
  float mw = raw[0], mh = raw[K];
#pragma unroll
  for (int i = 1; i < K; ++i) {
    mw = fmaxf(mw, raw[i]);
    mh = fmaxf(mh, raw[K + i]);
  }
  float ew[K], eh[K], sw = 0.0f, sh = 0.0f;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    ew[i] = expf(raw[i] - mw);
    eh[i] = expf(raw[K + i] - mh);
    sw += ew[i];
    sh += eh[i];
  }
  q.xp[0] = rmin;
  q.yp[0] = rmin;
  float cx = 0.0f, cy = 0.0f;
#pragma unroll
  for (int i = 0; i < K - 1; ++i) {
    const float bw = (ew[i] / sw) * scale + 1e-4f;
    const float bh = (eh[i] / sh) * scale + 1e-4f;
    cx = (i == 0) ? bw : cx + bw;
    cy = (i == 0) ? bh : cy + bh;
    q.xp[i + 1] = rmin + cx;
    q.yp[i + 1] = rmin + cy;
  }
  q.xp[K] = rmax;
  q.yp[K] = rmax;
#pragma unroll
  for (int i = 0; i <= K; ++i) q.sl[i] = softplus_f(raw[2 * K + i] + offset) + 1e-4f;
}

// select the bin of v on axis `pos` (first bin if none), gather both axes and the slopes
template <int K>
__device__ __forceinline__ void rq_select(const RQ& q, const float* pos, float v, float& xl, float& xr, float& yl,
                                          float& yr, float& dl, float& dr) {
  xl = q.xp[0]; xr = q.xp[1]; yl = q.yp[0]; yr = q.yp[1]; dl = q.sl[0]; dr = q.sl[1];
#pragma unroll
  for (int i = 1; i < K; ++i) {
    const bool in = (v >= pos[i]) && (v < pos[i + 1]);
    xl = in ? q.xp[i] : xl; xr = in ? q.xp[i + 1] : xr;
    yl = in ? q.yp[i] : yl; yr = in ? q.yp[i + 1] : yr;
    dl = in ? q.sl[i] : dl; dr = in ? q.sl[i + 1] : dr;
  }
}

// _rational_quadratic_spline_fwd (rqSpline.py:42-128)
template <int K>
__device__ __forceinline__ float rq_forward(const RQ& q, float x, float& logdet) {
  float xl, xr, yl, yr, dl, dr;
  rq_select<K>(q, q.xp, x, xl, xr, yl, yr, dl, dr);
  const float bw = xr - xl, bh = yr - yl;
  const float s = bh / bw;
  float z = (x - xl) / bw;
  z = fminf(fmaxf(z, 0.0f), 1.0f);
  const float sq_z = z * z, z1mz = z - sq_z, omz = 1.0f - z, sq_1mz = omz * omz;
  const float st = dr + dl - 2.0f * s;
  const float num = bh * (s * sq_z + dl * z1mz);
  const float den = s + st * z1mz;
  float y = yl + num / den;
  logdet = 2.0f * logf(s) + logf(dr * sq_z + 2.0f * s * z1mz + dl * sq_1mz) - 2.0f * logf(den);
  const bool below = x <= q.xp[0], above = x >= q.xp[K];
  y = below ? (x - q.xp[0]) * q.sl[0] + q.yp[0] : y;
  y = above ? (x - q.xp[K]) * q.sl[K] + q.yp[K] : y;
  logdet = below ? logf(q.sl[0]) : logdet;
  logdet = above ? logf(q.sl[K]) : logdet;
  return y;
}

// _rational_quadratic_spline_inv + _safe_quadratic_root (rqSpline.py:131-239)
template <int K>
__device__ __forceinline__ float rq_inverse(const RQ& q, float y, float& logdet) {
  float xl, xr, yl, yr, dl, dr;
  rq_select<K>(q, q.yp, y, xl, xr, yl, yr, dl, dr);
  const float bw = xr - xl, bh = yr - yl;
  const float s = bh / bw;
  float w = (y - yl) / bh;
  w = fminf(fmaxf(w, 0.0f), 1.0f);
  const float st = dr + dl - 2.0f * s;
  const float c = -s * w;
  const float b = dl - st * w;
  const float a = s - b;
  const float disc = b * b - 4.0f * a * c;
  float sq = sqrtf(fmaxf(disc, 1.17549435e-38f));
  sq = (disc > 0.0f) ? sq : 0.0f;
  const float num = (b >= 0.0f) ? 2.0f * c : -b + sq;
  const float den = (b >= 0.0f) ? -b - sq : 2.0f * a;
  float z = num / den;
  z = fminf(fmaxf(z, 0.0f), 1.0f);
  float x = bw * z + xl;
  const float sq_z = z * z, z1mz = z - sq_z, omz = 1.0f - z, sq_1mz = omz * omz;
  const float dn = s + st * z1mz;
  logdet = -2.0f * logf(s) - logf(dr * sq_z + 2.0f * s * z1mz + dl * sq_1mz) + 2.0f * logf(dn);
  const bool below = y <= q.yp[0], above = y >= q.yp[K];
  x = below ? (y - q.yp[0]) / q.sl[0] + q.xp[0] : x;
  x = above ? (y - q.yp[K]) / q.sl[K] + q.xp[K] : x;
  logdet = below ? -logf(q.sl[0]) : logdet;
  logdet = above ? -logf(q.sl[K]) : logdet;
  return x;
}

}  // namespace flowmc
