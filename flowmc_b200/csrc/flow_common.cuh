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

// Lean fused evaluation used by the flow kernels' epilogues: parameters + transform of ONE (sample, feature) with
// only the work the selected bin needs -- 2K exponentials for the two softmaxes, one reciprocal per softmax
// (bin size = e_i * (scale / sum) + 1e-4; the reference forms (e_i / sum) * scale, equal up to one rounding), and
// softplus only for the two knot slopes of the selected bin (or the boundary slope in the linear tails) instead of
// all K+1.  Same bin-selection rule and transform formulas as rq_forward / rq_inverse above.
template <int K, bool INV>
__device__ __forceinline__ float rq_apply(const float* raw, float rmin, float rmax, float v, float& logdet) {
  const float scale = (rmax - rmin) - (float)K * 1e-4f;
  const float offset = 0.5411666035652161f;
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
  const float rw = scale / sw, rh = scale / sh;
  // walk the knots, keeping the selected bin's corners and slope logits
  float xk = rmin, yk = rmin;  // knot i
  float xl = rmin, yl = rmin, xr = 0.0f, yr = 0.0f, ul = raw[2 * K], ur = raw[2 * K + 1];
  float cx = 0.0f, cy = 0.0f;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    float xn, yn;  // knot i + 1
    if (i < K - 1) {
      const float bw = ew[i] * rw + 1e-4f, bh = eh[i] * rh + 1e-4f;
      cx = (i == 0) ? bw : cx + bw;
      cy = (i == 0) ? bh : cy + bh;
      xn = rmin + cx;
      yn = rmin + cy;
    } else {
      xn = rmax;
      yn = rmax;
    }
    const float pk = INV ? yk : xk, pn = INV ? yn : xn;
    const bool in = (i == 0) || ((v >= pk) && (v < pn));  // first bin if none matches (rqSpline.py:63-72)
    const bool take = (i == 0) ? true : in;
    xl = take ? xk : xl; xr = take ? xn : xr;
    yl = take ? yk : yl; yr = take ? yn : yr;
    if (i > 0) {
      ul = take ? raw[2 * K + i] : ul;
      ur = take ? raw[2 * K + i + 1] : ur;
    }
    xk = xn;
    yk = yn;
  }
  const bool below = v <= rmin, above = v >= rmax;  // knot 0 and knot K are the range ends (both axes)
  ul = below ? raw[2 * K] : ul;
  ur = above ? raw[3 * K] : ur;
  const float dl = softplus_f(ul + offset) + 1e-4f, dr = softplus_f(ur + offset) + 1e-4f;
  const float bw = xr - xl, bh = yr - yl;
  const float s = bh / bw;
  const float st = dr + dl - 2.0f * s;
  float out, z;
  if (!INV) {
    z = (v - xl) / bw;
    z = fminf(fmaxf(z, 0.0f), 1.0f);
  } else {
    float w = (v - yl) / bh;
    w = fminf(fmaxf(w, 0.0f), 1.0f);
    const float c = -s * w;
    const float b = dl - st * w;
    const float a = s - b;
    const float disc = b * b - 4.0f * a * c;
    float sq = sqrtf(fmaxf(disc, 1.17549435e-38f));
    sq = (disc > 0.0f) ? sq : 0.0f;
    const float num = (b >= 0.0f) ? 2.0f * c : -b + sq;
    const float den = (b >= 0.0f) ? -b - sq : 2.0f * a;
    z = num / den;
    z = fminf(fmaxf(z, 0.0f), 1.0f);
  }
  const float sq_z = z * z, z1mz = z - sq_z, omz = 1.0f - z, sq_1mz = omz * omz;
  const float den = s + st * z1mz;
  const float ldin = 2.0f * logf(s) + logf(dr * sq_z + 2.0f * s * z1mz + dl * sq_1mz) - 2.0f * logf(den);
  if (!INV) {
    out = yl + bh * (s * sq_z + dl * z1mz) / den;
    out = below ? (v - rmin) * dl + rmin : out;
    out = above ? (v - rmax) * dr + rmax : out;
    logdet = below ? logf(dl) : (above ? logf(dr) : ldin);
  } else {
    out = bw * z + xl;
    out = below ? (v - rmin) / dl + rmin : out;
    out = above ? (v - rmax) / dr + rmax : out;
    logdet = below ? -logf(dl) : (above ? -logf(dr) : -ldin);
  }
  return out;
}

// ---- fast variants for the tensor-core epilogue ---------------------------------------------------------------
// The epilogue of the tcgen05 path is latency-bound: one long dependent chain per (sample, feature).  These use the
// hardware approximations (ex2 / lg2 / rcp / sqrt .approx, 1-2 ulp, no slow-path branches) so the chain is ~2x
// shorter; every quantity stays within ~3e-7 relative of the accurate version, far inside the 1e-5 tolerance.
__device__ __forceinline__ float fast_exp(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
  return r;
}
__device__ __forceinline__ float fast_log(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * 0.6931471805599453f;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_softplus(float x) {  // max(x,0) + log1p(exp(-|x|)), log1p by series when tiny
  const float e = fast_exp(-fabsf(x));
  const float series = e * fmaf(e, fmaf(e, 0.33333334f, -0.5f), 1.0f);
  const float l = (e < 0.03125f) ? series : fast_log(1.0f + e);
  return fmaxf(x, 0.0f) + l;
}

template <int K, bool INV>
__device__ __forceinline__ float rq_apply_fast(const float* raw, float rmin, float rmax, float v, float& logdet) {
  const float scale = (rmax - rmin) - (float)K * 1e-4f;
  const float offset = 0.5411666035652161f;
  float mw = raw[0], mh = raw[K];
#pragma unroll
  for (int i = 1; i < K; ++i) {
    mw = fmaxf(mw, raw[i]);
    mh = fmaxf(mh, raw[K + i]);
  }
  float ew[K], eh[K], sw = 0.0f, sh = 0.0f;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    ew[i] = fast_exp(raw[i] - mw);
    eh[i] = fast_exp(raw[K + i] - mh);
    sw += ew[i];
    sh += eh[i];
  }
  const float rw = scale * fast_rcp(sw), rh = scale * fast_rcp(sh);
  float xk = rmin, yk = rmin;
  float xl = rmin, yl = rmin, xr = 0.0f, yr = 0.0f, ul = raw[2 * K], ur = raw[2 * K + 1];
  float cx = 0.0f, cy = 0.0f;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    float xn, yn;
    if (i < K - 1) {
      cx += fmaf(ew[i], rw, 1e-4f);
      cy += fmaf(eh[i], rh, 1e-4f);
      xn = rmin + cx;
      yn = rmin + cy;
    } else {
      xn = rmax;
      yn = rmax;
    }
    const float pk = INV ? yk : xk, pn = INV ? yn : xn;
    const bool take = (i == 0) ? true : ((v >= pk) && (v < pn));
    xl = take ? xk : xl; xr = take ? xn : xr;
    yl = take ? yk : yl; yr = take ? yn : yr;
    if (i > 0) {
      ul = take ? raw[2 * K + i] : ul;
      ur = take ? raw[2 * K + i + 1] : ur;
    }
    xk = xn;
    yk = yn;
  }
  const bool below = v <= rmin, above = v >= rmax;
  ul = below ? raw[2 * K] : ul;
  ur = above ? raw[3 * K] : ur;
  const float dl = fast_softplus(ul + offset) + 1e-4f, dr = fast_softplus(ur + offset) + 1e-4f;
  const float bw = xr - xl, bh = yr - yl;
  const float rbw = fast_rcp(bw);
  const float s = bh * rbw;
  const float st = dr + dl - 2.0f * s;
  float out, z;
  if (!INV) {
    z = (v - xl) * rbw;
    z = fminf(fmaxf(z, 0.0f), 1.0f);
  } else {
    float w = (v - yl) * fast_rcp(bh);
    w = fminf(fmaxf(w, 0.0f), 1.0f);
    const float c = -s * w;
    const float b = dl - st * w;
    const float a = s - b;
    const float disc = b * b - 4.0f * a * c;
    float sq = fast_sqrt(fmaxf(disc, 1.17549435e-38f));
    sq = (disc > 0.0f) ? sq : 0.0f;
    const float num = (b >= 0.0f) ? 2.0f * c : -b + sq;
    const float den = (b >= 0.0f) ? -b - sq : 2.0f * a;
    z = num * fast_rcp(den);
    z = fminf(fmaxf(z, 0.0f), 1.0f);
  }
  const float sq_z = z * z, z1mz = z - sq_z, omz = 1.0f - z, sq_1mz = omz * omz;
  const float den = s + st * z1mz;
  const float rden = fast_rcp(den);
  const float qd = dr * sq_z + 2.0f * s * z1mz + dl * sq_1mz;
  // 2 log s + log q - 2 log den = log(s^2 q / den^2): one logarithm
  const float ldin = fast_log((s * s) * qd * (rden * rden));
  if (!INV) {
    out = yl + bh * (s * sq_z + dl * z1mz) * rden;
    out = below ? (v - rmin) * dl + rmin : out;
    out = above ? (v - rmax) * dr + rmax : out;
    logdet = below ? fast_log(dl) : (above ? fast_log(dr) : ldin);
  } else {
    out = bw * z + xl;
    out = below ? (v - rmin) * fast_rcp(dl) + rmin : out;
    out = above ? (v - rmax) * fast_rcp(dr) + rmax : out;
    logdet = below ? -fast_log(dl) : (above ? -fast_log(dr) : -ldin);
  }
  return out;
}

// tanh through one exponential: 1 - 2 / (1 + 2^(2 log2(e) x)) with the hardware ex2 / rcp approximations
// (absolute error <= ~3e-7 over the whole line: the hidden activations feed dot products, where that is below the
// fp32 rounding of the sum).  5 instructions instead of tanhf's ~30.
__device__ __forceinline__ float tanh_ex2(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return fmaf(-2.0f, r, 1.0f);
}

}  // namespace flowmc
