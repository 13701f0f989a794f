// NFProposal global steps: independence Metropolis-Hastings with flow proposals, all chains and
// all n_steps proposals of one TakeGroupSteps call.
//
// Reference: TakeSteps.__call__ + TakeGroupSteps.sample (src/flowMC/strategy/take_steps.py:60-144,
// 191-206) driving NFProposal.kernel / sample_flow (src/flowMC/resource/kernel/NF_proposal.py:27-172).
//
// Passes (all enqueued on the caller's stream, no host synchronisation):
//   1. target logp of the initial positions            (take_steps.py:201)      -> lp0[n]
//   2. flow log_prob of the initial positions          (NF_proposal.py:44)      -> lp_nf_cur[n]
//   3. nf_propose_kernel: for every (chain, step) row, in ONE kernel with the 64-row tile resident
//      in shared memory: per-row key schedule, z = normal(key), flow INVERSE, un-whiten -> proposal
//      (stored once), re-whiten, flow FORWARD, base log-prob -> lp_nf_prop   (NF_proposal.py:130-172;
//      the reference recomputes log_prob with a forward pass rather than reusing the inverse's
//      log-det, and that is not equivalent once ScalarAffine trains -- SURVEY.md B.6)
//   4. target logp of all proposals                    (NF_proposal.py:50-89)   -> lp_prop
//   5. nf_accept_kernel: one warp per chain walks the n_steps accept/reject scan
//      (NF_proposal.py:91-126) and writes the thinned positions / log-probs / accept flags straight
//      into the sampler buffers at the cursor (take_steps.py:134-142).
#include <string>

#include "flow_tile.cuh"
#include "registry.h"

namespace flowmc {

struct NfArgs {
  Key subkey;                  // take_steps.py:71 subkey; chain key = split(subkey, n_chains_global)[global index]
  const uint32_t* chain_keys;  // optional explicit per-chain keys (NFProposal.kernel called directly)
  int64_t chain_offset, n_chains;
  int n_steps, n_batch, n_sample;  // n_batch == 0: un-batched branch of sample_flow
};

__device__ __forceinline__ Key nf_chain_key(const NfArgs& a, int64_t c) {
  if (a.chain_keys != nullptr) return Key{a.chain_keys[2 * c], a.chain_keys[2 * c + 1]};
  return split_at(a.subkey, (uint64_t)(a.chain_offset + c));
}

template <int K>
__global__ void __launch_bounds__(NT) nf_propose_kernel(const FlowmcFlowDesc D, const float* __restrict__ P,
                                                        const NfArgs a, float* __restrict__ props,
                                                        float* __restrict__ lp_nf) {
  extern __shared__ __align__(16) float smem[];
  const FlowSmem S = flow_smem_layout(D);
  float* xs = smem + S.xs;
  float* ld = smem + S.ld;
  uint32_t* rk = reinterpret_cast<uint32_t*>(smem + S.ldw);  // [TM][3]: key words, first counter
  const int d = D.n_features;
  const int tid = threadIdx.x;
  const int64_t n = a.n_chains * a.n_steps;
  const int64_t row0 = (int64_t)blockIdx.x * TM;

  // ---- per-row key schedule (NF_proposal.py:41,135-163) ------------------------------------------
  if (tid < TM) {
    const int64_t r = min(row0 + tid, n - 1);
    const int64_t c = r / a.n_steps;
    const int t = (int)(r - c * a.n_steps);
    Key key = split_at(nf_chain_key(a, c), 1);  // rng_key, subkey = split(rng_key): proposals use subkey
    int idx = t;
    if (a.n_batch > 0) {
      const int b = t / a.n_sample;
      for (int i = 0; i < b; ++i) key = split_at(key, 0);  // scan carry: rng_key, subkey = split(rng_key)
      key = split_at(key, 1);
      idx = t - b * a.n_sample;
    }
    rk[3 * tid] = key.k0;
    rk[3 * tid + 1] = key.k1;
    rk[3 * tid + 2] = (uint32_t)idx * (uint32_t)d;
  }
  __syncthreads();
  // ---- base sample: mean + chol(cov) z with z = normal(key, (n_sample, d))[idx] -------------------
  for (int i = tid; i < TM * d; i += NT) {
    const int s = i / d, j = i - s * d;
    const Key key{rk[3 * s], rk[3 * s + 1]};
    const float z = bits_to_normal(bits_at(key, (uint64_t)(rk[3 * s + 2] + (uint32_t)j)));
    xs[s * S.xs_stride + j] = P[D.off_base_mean + j] + z * sqrtf(P[D.off_base_cov + (int64_t)j * d + j]);
  }
  if (tid < TM) ld[tid] = 0.0f;
  __syncthreads();

  flow_layers<K, true>(D, P, S, smem, row0, n, nullptr);

  // ---- proposal = inverse * sqrt(diag data_cov) + data_mean (rqSpline.py:495); keep it, re-whiten ---
  for (int i = tid; i < TM * d; i += NT) {
    const int s = i / d, j = i - s * d;
    const float sd = sqrtf(P[D.off_data_cov + (int64_t)j * d + j]);
    const float mu = P[D.off_data_mean + j];
    const float x = xs[s * S.xs_stride + j] * sd + mu;
    if (row0 + s < n) __stcs(props + (row0 + s) * d + j, x);
    xs[s * S.xs_stride + j] = (x - mu) / sd;  // log_prob's whitening (rqSpline.py:501)
  }
  if (tid < TM) ld[tid] = 0.0f;
  __syncthreads();

  flow_layers<K, false>(D, P, S, smem, row0, n, nullptr);

  if (tid < TM && row0 + tid < n) lp_nf[row0 + tid] = ld[tid] + base_log_prob(D, P, xs + tid * S.xs_stride);
}

// One warp per chain: sequential accept scan over the n_steps proposals.
__global__ void __launch_bounds__(128) nf_accept_kernel(const NfArgs a, int d, const float* __restrict__ x0,
                                                        const float* __restrict__ lp0,
                                                        const float* __restrict__ lp_nf_cur,
                                                        const float* __restrict__ props,
                                                        const float* __restrict__ lp_prop,
                                                        const float* __restrict__ lp_nf_prop, int thinning,
                                                        float* __restrict__ pos_buf, float* __restrict__ lp_buf,
                                                        float* __restrict__ acc_buf, int64_t n_total, int64_t cursor,
                                                        float* __restrict__ last_pos) {
  const int lane = threadIdx.x & 31;
  const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= a.n_chains) return;
  Key rk = split_at(nf_chain_key(a, c), 0);  // NF_proposal.py:41: the accept scan continues with rng_key
  float lp = lp0[c], lpnf = lp_nf_cur[c];
  const float* src = x0 + c * d;
  const int t_last = ((a.n_steps - 1) / thinning) * thinning;
  for (int t = 0; t < a.n_steps; ++t) {
    const Key s = split_at(rk, 1);
    rk = split_at(rk, 0);
    const float logu = logf(bits_to_uniform01(bits_at(s, 0)));
    const int64_t r = c * a.n_steps + t;
    const float lpp = lp_prop[r], lpn = lp_nf_prop[r];
    const float ratio = (lpp - lp) - (lpn - lpnf);  // NF_proposal.py:100-102
    const bool acc = logu < ratio;
    if (acc) {
      src = props + r * d;
      lp = lpp;
      lpnf = lpn;
    }
    if (t % thinning == 0) {
      const int64_t o = c * n_total + cursor + t / thinning;
      float* dst = pos_buf + o * d;
      if ((d & 3) == 0) {
        for (int j = lane * 4; j < d; j += 128)
          __stcs(reinterpret_cast<float4*>(dst + j), *reinterpret_cast<const float4*>(src + j));
      } else {
        for (int j = lane; j < d; j += 32) __stcs(dst + j, src[j]);
      }
      if (lane == 0) {
        __stcs(lp_buf + o, lp);
        __stcs(acc_buf + o, acc ? 1.0f : 0.0f);
      }
      if (t == t_last)
        for (int j = lane; j < d; j += 32) last_pos[c * d + j] = src[j];
    }
  }
}

template <int K>
static int launch_propose(const FlowmcFlowDesc& D, const float* P, const NfArgs& a, float* props, float* lp_nf,
                          cudaStream_t stream) {
  const FlowSmem S = flow_smem_layout(D);
  const size_t bytes = (size_t)S.total * sizeof(float);
  auto kern = nf_propose_kernel<K>;
  static size_t configured = 0;
  if (bytes > configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
      flowmc_set_error("nf_global_steps: model too large for the shared-memory tile");
      return FLOWMC_ERR_UNSUPPORTED;
    }
    configured = bytes;
  }
  const int64_t n = a.n_chains * a.n_steps;
  kern<<<(unsigned)((n + TM - 1) / TM), NT, bytes, stream>>>(D, P, a, props, lp_nf);
  flowmc_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    flowmc_set_error(cudaGetErrorString(e));
    return FLOWMC_ERR_CUDA;
  }
  return FLOWMC_OK;
}

static inline int64_t pad4(int64_t v) { return (v + 3) & ~(int64_t)3; }

}  // namespace flowmc

extern "C" {

int64_t flowmc_nf_global_steps_workspace_bytes(int64_t n_chains, int d, int n_steps) {
  if (n_chains <= 0 || d <= 0 || n_steps <= 0) return 0;
  const int64_t rows = n_chains * n_steps;
  return 4 * (flowmc::pad4(rows * d) + 2 * flowmc::pad4(rows) + 2 * flowmc::pad4(n_chains));
}

int flowmc_nf_global_steps(const FlowmcFlowDesc* D, const float* params, int target_id, const float* target_data,
                           const uint32_t key[2], const float* x0, float* pos_buf, float* lp_buf, float* acc_buf,
                           int64_t n_total, int64_t cursor, int64_t n_chains, int n_steps, int thinning,
                           int64_t chain_offset, int64_t n_chains_global, const FlowmcGlobalParams* gp,
                           uint32_t key_out[2], float* last_pos, void* stream_) {
  using namespace flowmc;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!D || D->n_features < 1 || D->n_layers < 1 || D->n_linear < 2 || D->n_linear > FLOWMC_FLOW_MAX_LINEAR) {
    flowmc_set_error("nf_global_steps: bad flow descriptor");
    return FLOWMC_ERR_INVALID;
  }
  FlowmcTargetVTable vt;
  if (int rc = flowmc_get_target(target_id, &vt)) return rc;
  if (!key || !key_out || !gp || n_chains < 0 || n_steps < 0 || thinning <= 0 || gp->n_batch_size <= 0) {
    flowmc_set_error("nf_global_steps: bad arguments");
    return FLOWMC_ERR_INVALID;
  }
  if (chain_offset < 0 || chain_offset + n_chains > n_chains_global) {
    flowmc_set_error("nf_global_steps: chain shard outside [0, n_chains_global)");
    return FLOWMC_ERR_INVALID;
  }
  const int64_t n_out = (n_steps + thinning - 1) / thinning;
  if (cursor < 0 || cursor + n_out > n_total) {
    flowmc_set_error("nf_global_steps: cursor + n_steps/thinning exceeds the buffer length");
    return FLOWMC_ERR_INVALID;
  }
  // take_steps.py:71: rng_key, subkey = split(rng_key)
  const Key k{key[0], key[1]};
  const Key knew = split_at(k, 0), sub = split_at(k, 1);
  key_out[0] = knew.k0;
  key_out[1] = knew.k1;
  if (n_chains == 0 || n_steps == 0) return FLOWMC_OK;
  const int d = D->n_features;
  const int64_t rows = n_chains * n_steps;
  if (!params || !x0 || !pos_buf || !lp_buf || !acc_buf || !last_pos || !gp->workspace ||
      gp->workspace_bytes < flowmc_nf_global_steps_workspace_bytes(n_chains, d, n_steps)) {
    flowmc_set_error("nf_global_steps: null buffer or workspace too small");
    return FLOWMC_ERR_INVALID;
  }
  float* ws = static_cast<float*>(gp->workspace);
  float* props = ws;
  float* lp_nf_prop = props + pad4(rows * d);
  float* lp_prop = lp_nf_prop + pad4(rows);
  float* lp_nf_cur = lp_prop + pad4(rows);
  float* lp0 = lp_nf_cur + pad4(n_chains);

  NfArgs a;
  a.subkey = sub;
  a.chain_keys = gp->chain_keys;
  a.chain_offset = chain_offset;
  a.n_chains = n_chains;
  a.n_steps = n_steps;
  a.n_batch = 0;
  a.n_sample = n_steps;
  if (n_steps > gp->n_batch_size) {  // NF_proposal.py:135-137
    a.n_batch = (n_steps + gp->n_batch_size - 1) / gp->n_batch_size;
    a.n_sample = (n_steps + a.n_batch - 1) / a.n_batch;
  }

  const float* lp0_in = gp->lp0;
  if (lp0_in == nullptr) {
    if (int rc = vt.eval(target_data, x0, n_chains, d, lp0, nullptr, stream)) return rc;
    lp0_in = lp0;
  }
  if (int rc = flow_transform(*D, false, params, x0, n_chains, nullptr, lp_nf_cur, nullptr, PRE_WHITEN,
                              POST_BASE_LOGP, nullptr, Key{0, 0}, 1, stream))
    return rc;
  int rc;
  if (flow_tc_enabled(*D)) {
    rc = flow_nf_propose_tc(*D, params, a.subkey, a.chain_keys, a.chain_offset, a.n_chains, a.n_steps, a.n_batch,
                            a.n_sample, props, lp_nf_prop, stream);
  } else {
    switch (D->num_bins) {
      case 4: rc = launch_propose<4>(*D, params, a, props, lp_nf_prop, stream); break;
      case 8: rc = launch_propose<8>(*D, params, a, props, lp_nf_prop, stream); break;
      case 16: rc = launch_propose<16>(*D, params, a, props, lp_nf_prop, stream); break;
      default:
        flowmc_set_error("flow: num_bins must be 4, 8 or 16");
        return FLOWMC_ERR_UNSUPPORTED;
    }
  }
  if (rc) return rc;
  if (int rc2 = vt.eval(target_data, props, rows, d, lp_prop, nullptr, stream)) return rc2;
  const int wpb = 4;
  nf_accept_kernel<<<(unsigned)((n_chains + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
      a, d, x0, lp0_in, lp_nf_cur, props, lp_prop, lp_nf_prop, thinning, pos_buf, lp_buf, acc_buf, n_total, cursor,
      last_pos);
  flowmc_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    flowmc_set_error(cudaGetErrorString(e));
    return FLOWMC_ERR_CUDA;
  }
  return FLOWMC_OK;
}

int flowmc_nf_accept_scan(const uint32_t* chain_keys, int64_t n_chains, int d, int n_steps, int thinning,
                          const float* x0, const float* lp0, const float* lp_nf_cur, const float* props,
                          const float* lp_prop, const float* lp_nf_prop, float* pos_buf, float* lp_buf, float* acc_buf,
                          int64_t n_total, int64_t cursor, float* last_pos, void* stream_) {
  using namespace flowmc;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_chains < 0 || d < 1 || n_steps < 0 || thinning <= 0) {
    flowmc_set_error("nf_accept_scan: bad sizes");
    return FLOWMC_ERR_INVALID;
  }
  const int64_t n_out = (n_steps + thinning - 1) / thinning;
  if (cursor < 0 || cursor + n_out > n_total) {
    flowmc_set_error("nf_accept_scan: cursor + n_steps/thinning exceeds the buffer length");
    return FLOWMC_ERR_INVALID;
  }
  if (n_chains == 0 || n_steps == 0) return FLOWMC_OK;
  if (!chain_keys || !x0 || !lp0 || !lp_nf_cur || !props || !lp_prop || !lp_nf_prop || !pos_buf || !lp_buf ||
      !acc_buf || !last_pos) {
    flowmc_set_error("nf_accept_scan: null buffer");
    return FLOWMC_ERR_INVALID;
  }
  NfArgs a;
  a.subkey = Key{0, 0};
  a.chain_keys = chain_keys;
  a.chain_offset = 0;
  a.n_chains = n_chains;
  a.n_steps = n_steps;
  a.n_batch = 0;
  a.n_sample = n_steps;
  const int wpb = 4;
  nf_accept_kernel<<<(unsigned)((n_chains + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
      a, d, x0, lp0, lp_nf_cur, props, lp_prop, lp_nf_prop, thinning, pos_buf, lp_buf, acc_buf, n_total, cursor,
      last_pos);
  flowmc_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    flowmc_set_error(cudaGetErrorString(e));
    return FLOWMC_ERR_CUDA;
  }
  return FLOWMC_OK;
}

}  // extern "C"
