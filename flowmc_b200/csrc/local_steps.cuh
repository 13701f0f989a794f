// Persistent local-step kernels for sm_100a: MALA, HMC and Gaussian random walk.
//
// One launch runs ALL n_steps of TakeSerialSteps for a batch of chains:
//   reference path  TakeSteps.__call__ -> jit(vmap(TakeSerialSteps.sample)) -> lax.scan(body)
//                   (src/flowMC/strategy/take_steps.py:60-144,156-180) with
//                   MALA.kernel (resource/kernel/MALA.py:26-89), HMC.kernel (HMC.py:98-151),
//                   GaussianRandomWalk.kernel (Gaussian_random_walk.py:25-61)
//   plus the three Buffer.update_buffer copies (resource/buffers.py:32-41), which become direct
//   streaming stores into the chain-major sample buffers at the strategy's cursor.
//
// Mapping.  A chain lives in a group of G lanes of one warp; each lane owns DPL dimensions in
// registers (position, cached gradient, proposal).  Dimension j of slot k on lane lg is
//   j = (k / VEC) * (G * VEC) + lg * VEC + (k % VEC)
// so that a warp-wide VEC-wide store of slot block k/VEC writes G*VEC contiguous floats per chain
// (full 32B sectors for VEC=4).  32/G chains share a warp, one warp per CTA (no CTA-level
// synchronisation anywhere).
//
// RNG.  Per chain and step the reference consumes d+5 threefry2x32 blocks:
//   (k_c, s) = split(k_c); (key1, key2) = split(s); z = normal(key1, (d,)); u = uniform(key2).
// The d normals are drawn by the owning lanes: first all DPL blocks of a lane (branch-free, so
// the compiler interleaves the independent 20-round chains), then the float transforms.  The
// five key-schedule blocks are warp-uniform work, so they are batched over chunks of 32 steps:
// the serial k_c chain runs once per chunk, then the 32 steps' (s, key1, key2, u) are computed
// one step per lane and parked in shared memory.  Bits are identical to jax.random's.
//
// Time slicing (optional).  With more warps than resident slots a plain grid runs in waves; alternatively the
// step range can be cut into S segments and the (segment, group) work items, numbered item = segment * n_groups + group, run as ROUNDS: launch r
// holds the items [r * R, (r + 1) * R) with R = min(slots, n_groups) CTAs, i.e. exactly one resident
// wave.  Item (s, g) runs steps [s*Ls, (s+1)*Ls) of group g and leaves the chain state (position,
// cached gradient, log-prob, chain key) in a small global workspace for item (s+1, g), which lies
// n_groups >= R items later and therefore always in a LATER launch on the same stream: the
// kernel boundary is the only synchronisation (no flags, no spinning, no fences), so the schedule
// and its duration are deterministic.  Slicing is OPT-IN (FlowmcLocalParams.force_n_seg > 1): on B200
// the plain one-launch grid measured faster at every BASELINE shape (see launch_local_one) -- the kernel is
// issue-bound, so a partially filled last wave still runs near full rate.
//
// The gradient of the current point is carried across steps (the reference recomputes it every
// step, MALA.py:59 -- same value, half the work).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "rng.cuh"
#include "registry.h"

namespace flowmc {

// KIND_*_PT: the same kernel on the TEMPERED density beta_c * logpdf(x) + log_prior(x) with a per-chain inverse
// temperature (ParallelTempering's individual steps with any ProposalBase, strategy/parallel_tempering.py:189-197;
// resource/logPDF.py:104-106).  KIND % 3 is the proposal, KIND >= 3 the tempered variant.
enum : int { KIND_MALA = 0, KIND_HMC = 1, KIND_GRW = 2, KIND_MALA_PT = 3, KIND_HMC_PT = 4, KIND_GRW_PT = 5 };

constexpr int kChunk = 32;  // steps per key-schedule chunk (= one step per lane)

// PAD = false promises d == G * DPL (no padding dimensions): every `j < d` test folds away.
template <int G, int DPL, int VEC, bool PAD = true>
struct Layout {
  static constexpr int kG = G, kDPL = DPL, kVEC = VEC;
  static constexpr bool kPAD = PAD;
  __device__ __forceinline__ static bool valid(int j, int d) { return !PAD || j < d; }
  static constexpr int CPW = 32 / G;     // chains per warp
  static constexpr int DS = G * DPL;     // padded dimension (smem row length)
  static constexpr int NSTATE = 2 * DPL + 3;  // floats per lane handed between time slices
  __device__ __forceinline__ static int dim(int k, int lg) {
    return (k / VEC) * (G * VEC) + lg * VEC + (k % VEC);
  }
};

constexpr int kHalo = 4;  // zero floats kept before x[0] and after x[DS-1] (x[-1] == x[d] == 0 for targets)

template <int CPW, int DS>
struct WarpSmem {
  float xrow[CPW][DS + 2 * kHalo];
  float scratch[CPW][DS];
  uint32_t k0[CPW][kChunk];
  uint32_t k1[CPW][kChunk];
  float logu[CPW][kChunk];
  float lpst[CPW][kChunk];
  float accst[CPW][kChunk];
};

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One evaluation of target T at the point held in xv (slot layout L); returns logp on every lane
// of the group and, if WANT_GRAD, the owned gradient entries in gv.
template <class T, class L, bool WANT_GRAD>
__device__ __forceinline__ float eval_target(const typename T::Consts& tc, const float (&xv)[L::kDPL],
                                             float (&gv)[L::kDPL], float* xrow, float* scratch,
                                             const float* data, int d, int lg) {
  constexpr int DPL = L::kDPL, VEC = L::kVEC, G = L::kG;
  // stage the point in shared memory so the target sees the whole vector
#pragma unroll
  for (int k = 0; k < DPL; k += VEC) {
    const int j = L::dim(k, lg);
    if (VEC == 4) {
      *reinterpret_cast<float4*>(xrow + j) = make_float4(xv[k], xv[k + 1], xv[k + 2], xv[k + 3]);
    } else if (VEC == 2) {
      *reinterpret_cast<float2*>(xrow + j) = make_float2(xv[k], xv[k + 1]);
    } else {
      xrow[j] = xv[k];
    }
  }
  __syncwarp();
  TargetCtx ctx{xrow, scratch, data, d};
  float red[T::NRED];
#pragma unroll
  for (int r = 0; r < T::NRED; ++r) red[r] = 0.0f;
  float aux[DPL];
#pragma unroll
  for (int k = 0; k < DPL; ++k) {
    const int j = L::dim(k, lg);
    aux[k] = 0.0f;
    if (L::valid(j, d)) aux[k] = T::partial(tc, ctx, j, xv[k], red);
  }
  if (T::USES_SCRATCH) __syncwarp();
#pragma unroll
  for (int r = 0; r < T::NRED; ++r) red[r] = group_sum<G>(red[r]);
  const float lp = T::finish(tc, ctx, red);
  if (WANT_GRAD) {
#pragma unroll
    for (int k = 0; k < DPL; ++k) {
      const int j = L::dim(k, lg);
      gv[k] = L::valid(j, d) ? T::grad(tc, ctx, j, xv[k], aux[k], red) : 0.0f;
    }
  }
  __syncwarp();  // xrow/scratch may be overwritten by the next evaluation
  return lp;
}

// Tempered evaluation: beta * logpdf(x) + log_prior(x) and its gradient.  The prior is the fixed-function family
// log_prior(x) = -sum_j c_j (x_j - m_j)^2 inside the box [lo, hi], -inf outside (Gaussian, uniform and flat priors);
// prior == nullptr is the flat prior 0 (test/unit/test_strategies.py:345-350).  prior layout: [4][d] = c, m, lo, hi.
template <class T, class L, bool WANT_GRAD>
__device__ __forceinline__ float eval_tempered(const typename T::Consts& tc, const float (&xv)[L::kDPL],
                                               float (&gv)[L::kDPL], float* xrow, float* scratch, const float* data,
                                               int d, int lg, float beta, const float* prior) {
  constexpr int DPL = L::kDPL, G = L::kG;
  float lp = beta * eval_target<T, L, WANT_GRAD>(tc, xv, gv, xrow, scratch, data, d, lg);
  if (WANT_GRAD) {
#pragma unroll
    for (int k = 0; k < DPL; ++k) gv[k] = beta * gv[k];
  }
  if (prior != nullptr) {
    float ps = 0.0f, outside = 0.0f;
#pragma unroll
    for (int k = 0; k < DPL; ++k) {
      const int j = L::dim(k, lg);
      if (L::valid(j, d)) {
        const float c = __ldg(prior + j), r = xv[k] - __ldg(prior + d + j);
        ps += c * r * r;
        if (xv[k] < __ldg(prior + 2 * d + j) || xv[k] > __ldg(prior + 3 * d + j)) outside = 1.0f;
        if (WANT_GRAD) gv[k] += -2.0f * c * r;
      }
    }
    ps = group_sum<G>(ps);
    outside = group_sum<G>(outside);
    lp += (outside > 0.0f) ? -INFINITY : -ps;
  }
  return lp;
}

template <class L>
__device__ __forceinline__ void store_row(float* dst, const float (&xv)[L::kDPL], int d, int lg) {
  constexpr int DPL = L::kDPL, VEC = L::kVEC;
#pragma unroll
  for (int k = 0; k < DPL; k += VEC) {
    const int j = L::dim(k, lg);
    if (L::valid(j, d)) {
      if (VEC == 4) {
        __stcs(reinterpret_cast<float4*>(dst + j), make_float4(xv[k], xv[k + 1], xv[k + 2], xv[k + 3]));
      } else if (VEC == 2) {
        __stcs(reinterpret_cast<float2*>(dst + j), make_float2(xv[k], xv[k + 1]));
      } else {
        __stcs(dst + j, xv[k]);
      }
    }
  }
}

// z[k] = normal(key, (d,))[dim(k)] for the DPL owned dimensions (0 for padding dimensions).
template <class L>
__device__ __forceinline__ void draw_normals(Key key, int d, int lg, float (&z)[L::kDPL]) {
  constexpr int DPL = L::kDPL;
  uint32_t bits[DPL];
#pragma unroll
  for (int k = 0; k < DPL; ++k) bits[k] = bits_at(key, (uint64_t)L::dim(k, lg));  // branch-free: interleaves
  bool tail = false;
#pragma unroll
  for (int k = 0; k < DPL; ++k) {
    float w;
    const float u = normal_arg(bits[k], w, L::valid(L::dim(k, lg), d));
    tail |= (w >= 5.0f);
    z[k] = 1.41421356237309515f * (erf_inv_central(w) * u);
  }
  if (tail) {  // |z| > ~2.9: 0.34 % of draws
#pragma unroll
    for (int k = 0; k < DPL; ++k) {
      float w;
      const float u = normal_arg(bits[k], w, L::valid(L::dim(k, lg), d));
      if (w >= 5.0f) z[k] = 1.41421356237309515f * (erf_inv_tail(w) * u);
    }
  }
}

// a / b with r = __frcp_rn(b) precomputed: quotient + one residual correction (the IEEE-rounded
// result for operands in the normal range; b is a per-launch constant here)
__device__ __forceinline__ float div_by(float a, float b, float r) {
  const float q = a * r;
  return fmaf(fmaf(-q, b, a), r, q);
}

struct Slice {       // time-slicing parameters (host-computed)
  int n_seg;         // number of segments S (1 = no slicing)
  int seg_len;       // steps per segment
  int n_groups;      // ceil(n_chains / CPW)
  int item_base;     // first (segment, group) work item of this launch
  float* state;      // workspace: n_groups * NSTATE * 32 floats (hand-off between segments)
};


template <class T, int KIND, class L, int MINB>
__global__ void __launch_bounds__(32, MINB) local_steps_kernel(const LocalArgs a, const Slice sl) {
  constexpr int G = L::kG, DPL = L::kDPL, CPW = L::CPW, DS = L::DS, NS = L::NSTATE;
  __shared__ __align__(16) WarpSmem<CPW, DS> sm;

  const int lane = threadIdx.x;
  const int lg = lane % G;
  const int cw = lane / G;

  // which (segment, group) work item this CTA runs; item (seg-1, grp) ran in an earlier launch
  const int item = sl.item_base + (int)blockIdx.x;
  const int seg = item / sl.n_groups;
  const int grp = item - seg * sl.n_groups;
  const int t_begin = seg * sl.seg_len;
  const int t_end = min(a.n_steps, t_begin + sl.seg_len);

  int64_t chain = (int64_t)grp * CPW + cw;
  const bool active = chain < a.n_chains;
  if (!active) chain = a.n_chains - 1;  // idle groups shadow the last chain; their stores are masked
  const int d = a.d;
  float* xrow = sm.xrow[cw] + kHalo;
  float* scratch = sm.scratch[cw];
  if (lg < kHalo) {  // zero halo: targets may read x[-1] and x[d] (== 0) without bounds tests
    xrow[-1 - lg] = 0.0f;
    xrow[DS + lg] = 0.0f;
  }
  if (G < kHalo && lg == 0) {
#pragma unroll
    for (int q = 0; q < kHalo; ++q) {
      xrow[-1 - q] = 0.0f;
      xrow[DS + q] = 0.0f;
    }
  }
  __syncwarp();
  const typename T::Consts tc = T::prepare(a.data, d);
  constexpr int BASE = KIND % 3;
  constexpr bool IS_MALA = BASE == KIND_MALA, IS_HMC = BASE == KIND_HMC, IS_GRW = BASE == KIND_GRW;
  constexpr bool TEMPERED = KIND >= KIND_MALA_PT;
  const float beta = (TEMPERED && a.beta != nullptr) ? a.beta[chain] : 1.0f;
  // one evaluation of the (tempered) target at a point, with or without its gradient
  auto eval_grad = [&](const float (&xv)[DPL], float (&gv)[DPL]) -> float {
    if (TEMPERED) return eval_tempered<T, L, true>(tc, xv, gv, xrow, scratch, a.data, d, lg, beta, a.prior);
    return eval_target<T, L, true>(tc, xv, gv, xrow, scratch, a.data, d, lg);
  };
  auto eval_value = [&](const float (&xv)[DPL], float (&gv)[DPL]) -> float {
    if (TEMPERED) return eval_tempered<T, L, false>(tc, xv, gv, xrow, scratch, a.data, d, lg, beta, a.prior);
    return eval_target<T, L, false>(tc, xv, gv, xrow, scratch, a.data, d, lg);
  };

  Key kc;
  float x[DPL], g[DPL], lp;
  if (seg == 0) {
    // per-chain key: split(subkey, n_chains_global)[global chain index]   (take_steps.py:72)
    kc = split_at(a.subkey, (uint64_t)(a.chain_offset + chain));
    // ... or an explicit per-chain key (ParallelTempering: split(split(subkey, n_chains)[c], n_temps)[t])
    if (a.chain_keys != nullptr) kc = Key{a.chain_keys[2 * chain], a.chain_keys[2 * chain + 1]};
#pragma unroll
    for (int k = 0; k < DPL; ++k) {
      const int j = L::dim(k, lg);
      x[k] = L::valid(j, d) ? a.x0[chain * d + j] : 0.0f;
      g[k] = 0.0f;
    }
    // logpdf(initial_position) seeds the scan carry (take_steps.py:177); MALA/HMC cache the gradient
    lp = IS_GRW ? eval_value(x, g) : eval_grad(x, g);
    // ProposalBase.kernel(): HMC and GRW use the caller-supplied log_prob (HMC.py:137,
    // Gaussian_random_walk.py:56); MALA ignores it and re-evaluates logpdf(position) (MALA.py:59,75,87)
    if (a.lp0 != nullptr && !IS_MALA) lp = a.lp0[chain];
  } else {
    // pick up the state the previous time slice of this group left (an earlier launch on this stream)
    const float* st = sl.state + ((int64_t)grp * NS) * 32 + lane;
#pragma unroll
    for (int k = 0; k < DPL; ++k) {
      x[k] = __ldcg(st + k * 32);
      g[k] = __ldcg(st + (DPL + k) * 32);
    }
    lp = __ldcg(st + (2 * DPL) * 32);
    kc.k0 = __float_as_uint(__ldcg(st + (2 * DPL + 1) * 32));
    kc.k1 = __float_as_uint(__ldcg(st + (2 * DPL + 2) * 32));
  }

  const float dt = a.step_size;
  const float dt2 = dt * dt;
  // scalar-covariance multivariate_normal.logpdf constant: n/2 * (log(2 pi) + log(cov))
  const float mvn_c = (float)d * 0.5f * (1.8378770664093453f + logf(dt2));
  const float rdt2 = __frcp_rn(dt2);

  float cs[DPL], ld[DPL];  // HMC: column sums of the metric, diagonal of L
#pragma unroll
  for (int k = 0; k < DPL; ++k) {
    cs[k] = 0.0f;
    ld[k] = 0.0f;
    if (IS_HMC) {
      const int j = L::dim(k, lg);
      if (L::valid(j, d)) {
        cs[k] = a.hmc_colsum[j];
        ld[k] = a.hmc_chol[(int64_t)j * d + j];
      }
    }
  }

  const int thin = a.thinning;
  const int t_last_stored = ((a.n_steps - 1) / thin) * thin;
  int o_next = (t_begin + thin - 1) / thin;  // next output index
  int t_next = o_next * thin;                // ... and the step that produces it
  float* out_row = a.pos_buf + (chain * a.n_total + a.cursor + o_next) * d;  // advanced by d per output

  for (int t0 = t_begin; t0 < t_end; t0 += kChunk) {
    const int nb = min(kChunk, t_end - t0);
    // ---- key schedule for this chunk -------------------------------------------------
    // serial part: k_c^{t+1} = split(k_c^t)[0]   (take_steps.py:158)
    for (int t = 0; t < nb; ++t) {
      if (lg == 0) {
        sm.k0[cw][t] = kc.k0;
        sm.k1[cw][t] = kc.k1;
      }
      kc = split_at(kc, 0);
    }
    __syncwarp();
    // parallel part, one step per lane: s = split(k_c^t)[1]; key1, key2 = split(s)  (MALA.py:66);
    // log(uniform(key2))  (MALA.py:83)
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
      Key kt{sm.k0[c][lane], sm.k1[c][lane]};
      Key s = split_at(kt, 1);
      if (a.step_keys != nullptr) {  // single kernel() call with explicit per-chain keys
        const int64_t ch = min((int64_t)grp * CPW + c, a.n_chains - 1);
        s = Key{a.step_keys[2 * ch], a.step_keys[2 * ch + 1]};
      }
      Key key1 = split_at(s, 0);
      Key key2 = split_at(s, 1);
      const float u = bits_to_uniform01(bits_at(key2, 0));
      sm.k0[c][lane] = key1.k0;
      sm.k1[c][lane] = key1.k1;
      sm.logu[c][lane] = logf(u);
    }
    __syncwarp();

    const int o_first = o_next;  // first output index of this chunk
    // ---- steps ---------------------------------------------------------------------------
    for (int tt = 0; tt < nb; ++tt) {
      const int t = t0 + tt;
      const Key key1{sm.k0[cw][tt], sm.k1[cw][tt]};
      const float logu = sm.logu[cw][tt];
      bool acc;

      if (IS_MALA) {
        float prop[DPL], g1[DPL];
        draw_normals<L>(key1, d, lg, prop);  // prop holds z for now
        float qa = 0.0f;
#pragma unroll
        for (int k = 0; k < DPL; ++k) {
          const float mean = x[k] + (dt2 * g[k]) / 2.0f;  // MALA.py:60
          prop[k] = mean + dt * prop[k];                  // MALA.py:61-63
          const float y = prop[k] - mean;
          qa += y * y;
        }
        const float lp1 = eval_grad(prop, g1);
        float qb = 0.0f;
#pragma unroll
        for (int k = 0; k < DPL; ++k) {
          const float y = x[k] - (prop[k] + (dt2 * g1[k]) / 2.0f);
          qb += y * y;
        }
        qa = group_sum<G>(qa);
        qb = group_sum<G>(qb);
        // MALA.py:75-81
        float ratio = lp1 - lp;
        ratio -= (div_by(-0.5f * qa, dt2, rdt2) - mvn_c);
        ratio += (div_by(-0.5f * qb, dt2, rdt2) - mvn_c);
        acc = logu < ratio;
#pragma unroll
        for (int k = 0; k < DPL; ++k) {
          x[k] = acc ? prop[k] : x[k];
          g[k] = acc ? g1[k] : g[k];
        }
        lp = acc ? lp1 : lp;
      } else if (IS_GRW) {
        float prop[DPL], g1[DPL];
        draw_normals<L>(key1, d, lg, prop);
#pragma unroll
        for (int k = 0; k < DPL; ++k) prop[k] = x[k] + prop[k] * dt;  // Gaussian_random_walk.py:49-52
        const float lp1 = eval_value(prop, g1);
        acc = logu < (lp1 - lp);  // Gaussian_random_walk.py:56
#pragma unroll
        for (int k = 0; k < DPL; ++k) x[k] = acc ? prop[k] : x[k];
        lp = acc ? lp1 : lp;
      } else {  // HMC
        float xs[DPL], p[DPL], g1[DPL];
        draw_normals<L>(key1, d, lg, p);
        // momentum = normal(key1) @ chol(inv(M)).T   (HMC.py:133-136)
        if (a.hmc_diag) {
#pragma unroll
          for (int k = 0; k < DPL; ++k) p[k] = p[k] * ld[k];
        } else {
#pragma unroll
          for (int k = 0; k < DPL; ++k) {
            const int j = L::dim(k, lg);
            if (L::valid(j, d)) scratch[j] = p[k];
          }
          __syncwarp();
#pragma unroll
          for (int k = 0; k < DPL; ++k) {
            const int j = L::dim(k, lg);
            float s = 0.0f;
            if (L::valid(j, d)) {
              const float* Lrow = a.hmc_chol + (int64_t)j * d;
              for (int i = 0; i <= j; ++i) s = fmaf(scratch[i], Lrow[i], s);
            }
            p[k] = s;
          }
          __syncwarp();
        }
        float kin = 0.0f;
#pragma unroll
        for (int k = 0; k < DPL; ++k) kin += p[k] * p[k] * cs[k];
        kin = 0.5f * group_sum<G>(kin);
        const float H = -lp + kin;  // HMC.py:137
        // leapfrog_step (HMC.py:82-96): rows [0,.5], n_leapfrog x [1,1], [1,.5].
        // Row 0 moves the position by eps*0*dK/dp = 0 and uses the gradient at the current point,
        // which is the cached g.
        const float eps = a.step_size;
#pragma unroll
        for (int k = 0; k < DPL; ++k) {
          xs[k] = x[k];
          p[k] = p[k] - (eps * 0.5f) * (-g[k]);
        }
        float lp1 = lp;
        for (int it = 1; it <= a.n_leapfrog + 1; ++it) {
          const float c1 = (it == a.n_leapfrog + 1) ? 0.5f : 1.0f;
#pragma unroll
          for (int k = 0; k < DPL; ++k) xs[k] = xs[k] + (eps * 1.0f) * (p[k] * cs[k]);
          lp1 = eval_grad(xs, g1);
#pragma unroll
          for (int k = 0; k < DPL; ++k) p[k] = p[k] - (eps * c1) * (-g1[k]);
        }
        float kin1 = 0.0f;
#pragma unroll
        for (int k = 0; k < DPL; ++k) kin1 += p[k] * p[k] * cs[k];
        kin1 = 0.5f * group_sum<G>(kin1);
        const float ham = -lp1 + kin1;  // HMC.py:141-142
        acc = logu < (H - ham);         // HMC.py:143-146
#pragma unroll
        for (int k = 0; k < DPL; ++k) {
          x[k] = acc ? xs[k] : x[k];
          g[k] = acc ? g1[k] : g[k];
        }
        lp = acc ? lp1 : lp;
      }

      // ---- outputs (take_steps.py:134-142): thinned, written in place at the cursor ----------
      if (t == t_next) {
        if (active) store_row<L>(out_row, x, d, lg);
        out_row += d;
        if (lg == 0) {
          sm.lpst[cw][o_next - o_first] = lp;
          sm.accst[cw][o_next - o_first] = acc ? 1.0f : 0.0f;
        }
        if (t == t_last_stored && active) store_row<L>(a.last_pos + chain * d, x, d, lg);
        ++o_next;
        t_next += thin;
      }
    }
    __syncwarp();
    // flush this chunk's log-probs and accept flags: one coalesced row per chain
    const int n_out = o_next - o_first;
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
      const int64_t ch = (int64_t)grp * CPW + c;
      if (ch < a.n_chains && lane < n_out) {
        const int64_t off = ch * a.n_total + a.cursor + o_first + lane;
        __stcs(a.lp_buf + off, sm.lpst[c][lane]);
        __stcs(a.acc_buf + off, sm.accst[c][lane]);
      }
    }
    __syncwarp();
  }

  if (t_end < a.n_steps) {  // hand the chain state to the next time slice
    float* st = sl.state + ((int64_t)grp * NS) * 32 + lane;
#pragma unroll
    for (int k = 0; k < DPL; ++k) {
      __stcg(st + k * 32, x[k]);
      __stcg(st + (DPL + k) * 32, g[k]);
    }
    __stcg(st + (2 * DPL) * 32, lp);
    __stcg(st + (2 * DPL + 1) * 32, __uint_as_float(kc.k0));
    __stcg(st + (2 * DPL + 2) * 32, __uint_as_float(kc.k1));
  }
}

// ---- target evaluation kernel (LogPDF.__call__ / value_and_grad) -----------------------------
template <class T, class L>
__global__ void __launch_bounds__(32) target_eval_kernel(const float* __restrict__ data, const float* __restrict__ xin,
                                                         int64_t n, int d, float* __restrict__ lp_out,
                                                         float* __restrict__ grad_out) {
  constexpr int G = L::kG, DPL = L::kDPL, CPW = L::CPW, DS = L::DS;
  __shared__ __align__(16) float xrow_s[CPW][DS + 2 * kHalo];
  __shared__ __align__(16) float scratch[CPW][DS];
  const int lane = threadIdx.x, lg = lane % G, cw = lane / G;
  float* xrow = xrow_s[cw] + kHalo;
  if (lg == 0) {
#pragma unroll
    for (int q = 0; q < kHalo; ++q) {
      xrow[-1 - q] = 0.0f;
      xrow[DS + q] = 0.0f;
    }
  }
  __syncwarp();
  int64_t i = (int64_t)blockIdx.x * CPW + cw;
  const bool active = i < n;
  if (!active) i = n - 1;
  const typename T::Consts tc = T::prepare(data, d);
  float x[DPL], g[DPL];
#pragma unroll
  for (int k = 0; k < DPL; ++k) {
    const int j = L::dim(k, lg);
    x[k] = (j < d) ? xin[i * d + j] : 0.0f;
    g[k] = 0.0f;
  }
  float lp;
  if (grad_out != nullptr) {
    lp = eval_target<T, L, true>(tc, x, g, xrow, scratch[cw], data, d, lg);
  } else {
    lp = eval_target<T, L, false>(tc, x, g, xrow, scratch[cw], data, d, lg);
  }
  if (active) {
    if (lg == 0) lp_out[i] = lp;
    if (grad_out != nullptr) {
#pragma unroll
      for (int k = 0; k < DPL; ++k) {
        const int j = L::dim(k, lg);
        if (j < d) grad_out[i * d + j] = g[k];
      }
    }
  }
}

// ---- layout table and launchers --------------------------------------------------------------
// Scalar-store layouts (any d):   0:(1,8,1) 1:(4,8,1) 2:(8,8,1) 3:(32,4,1) 4:(32,16,1)
// float4-store layouts (d % 4 == 0): 5:(8,4,4) 6:(16,4,4) 7:(16,8,4) 8:(32,4,4) 9:(32,8,4) 10:(32,16,4)
// Measured on B200 (scripts/sweep_local.py): few dimensions per lane win -- more resident warps and
// fewer registers beat the smaller key-schedule redundancy of wide lanes.
constexpr int kNumLayouts = 11;

struct LayoutInfo {
  int G, DPL, VEC;
};
__host__ inline LayoutInfo layout_info(int idx) {
  static const LayoutInfo tab[kNumLayouts] = {{1, 8, 1},  {4, 8, 1},  {8, 8, 1},  {32, 4, 1}, {32, 16, 1}, {8, 4, 4},
                                              {16, 4, 4}, {16, 8, 4}, {32, 4, 4}, {32, 8, 4}, {32, 16, 4}};
  return tab[idx];
}

__host__ inline int pick_layout(int d, int hint) {
  if (hint > 0 && hint <= kNumLayouts) {
    const LayoutInfo li = layout_info(hint - 1);
    if (li.G * li.DPL >= d && (li.VEC == 1 || d % li.VEC == 0)) return hint - 1;
  }
  if (d % 4 == 0) {
    if (d <= 32) return 5;
    if (d <= 64) return 6;
    if (d <= 128) return 7;
    if (d <= 256) return 9;
    if (d <= 512) return 10;
    return -1;
  }
  if (d <= 8) return 0;
  if (d <= 32) return 1;
  if (d <= 64) return 2;
  if (d <= 128) return 3;
  if (d <= 512) return 4;
  return -1;
}

// bytes of workspace flowmc_local_steps needs for time slicing with the layout it would pick
__host__ inline int64_t local_workspace_bytes(int64_t n_chains, int d, int hint) {
  const int li = pick_layout(d, hint);
  if (li < 0) return 0;
  const LayoutInfo L = layout_info(li);
  const int64_t cpw = 32 / L.G;
  const int64_t n_groups = (n_chains + cpw - 1) / cpw;
  return n_groups * (2 * L.DPL + 3) * 32 * 4;
}

// Resident CTA slots of one kernel instantiation on the current device.  The shared-memory carve-out is pinned to what
// MINB CTAs per SM need (the driver's default heuristic may pick a smaller one and halve the occupancy), then the
// occupancy is queried once per device.
constexpr int kMaxDev = 64;
struct SlotCache {  // one per kernel instantiation (a static local of launch_local_one), indexed by device
  int slots[kMaxDev], per_sm_c[kMaxDev], smem_c[kMaxDev];
};
template <class K>
inline void local_slots(K kern, int minb, SlotCache& sc, int* slots_out, int* per_sm_out, int* smem_out) {
  int (&slots)[kMaxDev] = sc.slots;
  int (&per_sm_c)[kMaxDev] = sc.per_sm_c;
  int (&smem_c)[kMaxDev] = sc.smem_c;
  int dev = 0;
  cudaGetDevice(&dev);
  const int di = (dev >= 0 && dev < kMaxDev) ? dev : 0;
  if (slots[di] == 0) {
    int sms = 0, per_sm = 0, smem_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    const int64_t need = (int64_t)minb * ((int64_t)fa.sharedSizeBytes + 1024);  // + the per-CTA reservation
    int pct = smem_sm > 0 ? (int)((need * 100 + smem_sm - 1) / smem_sm) + 1 : 50;
    if (pct > 100) pct = 100;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, 0);
    if (per_sm <= 0) per_sm = 1;
    per_sm_c[di] = per_sm;
    smem_c[di] = (int)fa.sharedSizeBytes;
    slots[di] = sms * per_sm;
    cudaGetLastError();
  }
  *slots_out = slots[di];
  *per_sm_out = per_sm_c[di];
  *smem_out = smem_c[di];
}

template <class T, int KIND, int G, int DPL, int VEC, int MINB, bool PAD = true>
inline int launch_local_one(const LocalArgs* a, cudaStream_t stream) {
  using L = Layout<G, DPL, VEC, PAD>;
  auto kern = local_steps_kernel<T, KIND, L, MINB>;
  const int64_t n_groups = (a->n_chains + L::CPW - 1) / L::CPW;
  Slice sl{1, a->n_steps, (int)n_groups, 0, nullptr};

  int slots = 0, per_sm = 0, smem = 0;
  static SlotCache slot_cache = {};  // per instantiation: occupancy differs between layouts / targets / kinds
  local_slots(kern, MINB, slot_cache, &slots, &per_sm, &smem);
  if (a->slots_override > 0) slots = a->slots_override;  // tests: force rounds with small grids
  const int64_t need = n_groups * L::NSTATE * 32 * 4;
  const bool can_slice = a->workspace != nullptr && a->workspace_bytes >= need && a->step_keys == nullptr;
  if (a->force_n_seg > 1) {
    if (!can_slice) {
      flowmc_set_error("local_steps: force_n_seg needs a workspace (flowmc_local_steps_workspace_bytes) and no step_keys");
      return -1;
    }
    const int s = a->force_n_seg < a->n_steps ? a->force_n_seg : a->n_steps;
    sl.seg_len = (a->n_steps + s - 1) / s;
    sl.n_seg = (a->n_steps + sl.seg_len - 1) / sl.seg_len;
  }
  // force_n_seg <= 0: one launch.  Measured on B200 (profiles/r02_slice_sweep.jsonl), the plain grid beats every sliced
  // plan at the BASELINE shapes -- C2 4.88 vs 5.09 ms, C3 7.41 vs 8.44 ms: the kernel is issue-bound, so the last,
  // partially filled wave of a plain grid still runs near full rate (fewer warps per SM, each faster), while every
  // round of a sliced plan pays a drain + launch (~45 us).  Slicing therefore stays an explicit option.
  if (sl.n_seg > 1) sl.state = reinterpret_cast<float*>(a->workspace);
  // a round = one resident wave; item (s, g) depends on item (s-1, g) = n_groups items earlier, which is in an earlier
  // round as long as a round holds at most n_groups items
  const int64_t n_items = n_groups * sl.n_seg;
  const int64_t round = sl.n_seg > 1 ? (slots < n_groups ? slots : n_groups) : n_items;
  if (a->plan_out != nullptr) {
    LocalPlan* p = a->plan_out;
    p->G = G; p->DPL = DPL; p->VEC = VEC;
    p->n_groups = (int)n_groups; p->slots = slots; p->ctas_per_sm = per_sm; p->smem_per_cta = smem;
    p->n_seg = sl.n_seg; p->seg_len = sl.seg_len;
    p->n_rounds = (int)((n_items + round - 1) / round); p->round_size = (int)round;
    return 0;  // plan only
  }
  for (int64_t base = 0; base < n_items; base += round) {
    sl.item_base = (int)base;
    const int64_t nblk = (n_items - base < round) ? n_items - base : round;
    kern<<<(unsigned)nblk, 32, 0, stream>>>(*a, sl);
    flowmc_count_launch();
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    flowmc_set_error(cudaGetErrorString(e));
    return -3;
  }
  return 0;
}

// float4 layouts come in two flavours: exact (d == G*DPL, no padding tests) and padded
template <class T, int KIND, int G, int DPL, int MINB>
inline int launch_local_v4(const LocalArgs* a, cudaStream_t stream) {
  if (a->d == G * DPL) return launch_local_one<T, KIND, G, DPL, 4, MINB, false>(a, stream);
  return launch_local_one<T, KIND, G, DPL, 4, MINB, true>(a, stream);
}

template <class T, int KIND>
inline int launch_local_kind(const LocalArgs* a, cudaStream_t stream) {
  const int li = pick_layout(a->d, a->layout_hint);
  if (a->plan_out != nullptr) a->plan_out->layout = li;
  switch (li) {
    case 0: return launch_local_one<T, KIND, 1, 8, 1, 16>(a, stream);
    case 1: return launch_local_one<T, KIND, 4, 8, 1, 16>(a, stream);
    case 2: return launch_local_one<T, KIND, 8, 8, 1, 16>(a, stream);
    case 3: return launch_local_one<T, KIND, 32, 4, 1, 24>(a, stream);
    case 4: return launch_local_one<T, KIND, 32, 16, 1, 12>(a, stream);
    case 5: return launch_local_v4<T, KIND, 8, 4, 24>(a, stream);
    case 6: return launch_local_v4<T, KIND, 16, 4, 24>(a, stream);
    case 7: return launch_local_v4<T, KIND, 16, 8, 16>(a, stream);
    case 8: return launch_local_v4<T, KIND, 32, 4, 24>(a, stream);
    case 9: return launch_local_v4<T, KIND, 32, 8, 16>(a, stream);
    case 10: return launch_local_v4<T, KIND, 32, 16, 12>(a, stream);
    default:
      flowmc_set_error("local_steps: unsupported dimension (d must be <= 512)");
      return -2;
  }
}

template <class T>
int launch_local_steps(int kind, const LocalArgs* a, cudaStream_t stream) {
  switch (kind) {
    case KIND_MALA: return launch_local_kind<T, KIND_MALA>(a, stream);
    case KIND_HMC: return launch_local_kind<T, KIND_HMC>(a, stream);
    case KIND_GRW: return launch_local_kind<T, KIND_GRW>(a, stream);
    case KIND_MALA_PT: return launch_local_kind<T, KIND_MALA_PT>(a, stream);
    case KIND_HMC_PT: return launch_local_kind<T, KIND_HMC_PT>(a, stream);
    case KIND_GRW_PT: return launch_local_kind<T, KIND_GRW_PT>(a, stream);
    default:
      flowmc_set_error("local_steps: unknown kernel kind");
      return -1;
  }
}

template <class T, int G, int DPL, int VEC>
inline int launch_eval_one(const float* data, const float* x, int64_t n, int d, float* lp, float* grad,
                           cudaStream_t stream) {
  using L = Layout<G, DPL, VEC>;
  const int64_t nblk = (n + L::CPW - 1) / L::CPW;
  target_eval_kernel<T, L><<<(unsigned)nblk, 32, 0, stream>>>(data, x, n, d, lp, grad);
  flowmc_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    flowmc_set_error(cudaGetErrorString(e));
    return -3;
  }
  return 0;
}

template <class T>
int launch_target_eval(const float* data, const float* x, int64_t n, int d, float* lp, float* grad,
                       cudaStream_t stream) {
  if (n <= 0) return 0;
  if (d <= 8) return launch_eval_one<T, 1, 8, 1>(data, x, n, d, lp, grad, stream);
  if (d <= 64) return launch_eval_one<T, 8, 8, 1>(data, x, n, d, lp, grad, stream);
  if (d <= 512) return launch_eval_one<T, 32, 16, 1>(data, x, n, d, lp, grad, stream);
  flowmc_set_error("target_eval: unsupported dimension (d must be <= 512)");
  return -2;
}

}  // namespace flowmc
