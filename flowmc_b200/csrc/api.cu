// C-ABI entry points of libflowmc_b200.so (declared in include/flowmc_b200.h): error handling,
// target registry, jax.random-compatible key management and draws, local-steps dispatch.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/flowmc_b200.h"
#include "registry.h"
#include "rng.cuh"
#include "../../include/flowmc_target.cuh"

namespace {
thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};

struct Registry {
  std::mutex mu;
  std::vector<FlowmcTargetVTable> targets;
};
Registry& registry() {
  static Registry r;  // constructed on first use: plugin static initialisers may run before ours
  return r;
}

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(FLOWMC_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return FLOWMC_OK;
}

__global__ void random_bits_kernel(flowmc::Key key, int64_t n, uint32_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = flowmc::bits_at(key, (uint64_t)i);
}
__global__ void random_uniform_kernel(flowmc::Key key, int64_t n, float lo, float hi, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float span = hi - lo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float f = flowmc::bits_to_unit(flowmc::bits_at(key, (uint64_t)i));
    out[i] = fmaxf(lo, __fadd_rn(__fmul_rn(f, span), lo));
  }
}
__global__ void random_normal_kernel(flowmc::Key key, int64_t n, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = flowmc::bits_to_normal(flowmc::bits_at(key, (uint64_t)i));
}

// ParallelTempering._exchange (strategy/parallel_tempering.py:291-398): one thread per chain walks the
// temperature ladder, idx = 0 .. n_temps - 2: key, sub = split(key); accept the swap of rungs idx / idx + 1 when
// log(uniform(sub)) < (1 / T[idx+1] - 1 / T[idx]) * (lp[idx] - lp[idx+1]), lp = the UNtempered log-density, which is
// swapped along with the position.
__global__ void pt_exchange_kernel(flowmc::Key subkey, int64_t chain_offset, int64_t n_chains, int n_temps, int d,
                                   float* __restrict__ pos, float* __restrict__ lp, const float* __restrict__ temps,
                                   float* __restrict__ acc) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chains) return;
  flowmc::Key key = flowmc::split_at(subkey, (uint64_t)(chain_offset + c));
  float* P = pos + c * n_temps * d;
  float* Lp = lp + c * n_temps;
  for (int idx = 0; idx + 1 < n_temps; ++idx) {
    const flowmc::Key sub = flowmc::split_at(key, 1);
    key = flowmc::split_at(key, 0);
    const float ratio = __fmul_rn(__fsub_rn(__fdiv_rn(1.0f, temps[idx + 1]), __fdiv_rn(1.0f, temps[idx])),
                                  __fsub_rn(Lp[idx], Lp[idx + 1]));
    const float log_uniform = logf(flowmc::bits_to_uniform01(flowmc::bits_at(sub, 0)));
    const bool accept = log_uniform < ratio;
    if (accept) {
      for (int j = 0; j < d; ++j) {
        const float t = P[idx * d + j];
        P[idx * d + j] = P[(idx + 1) * d + j];
        P[(idx + 1) * d + j] = t;
      }
      const float t = Lp[idx];
      Lp[idx] = Lp[idx + 1];
      Lp[idx + 1] = t;
    }
    acc[c * (n_temps - 1) + idx] = accept ? 1.0f : 0.0f;
  }
}

inline unsigned grid_for(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  const int64_t cap = 148 * 16;  // a few waves over the 148 SMs; kernels are grid-stride
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}
}  // namespace

extern "C" {

int flowmc_abi_version(void) { return FLOWMC_ABI_VERSION; }
const char* flowmc_last_error(void) { return g_err.c_str(); }
void flowmc_set_error(const char* msg) { g_err = msg ? msg : ""; }
void flowmc_count_launch(void) { g_launches.fetch_add(1, std::memory_order_relaxed); }
int64_t flowmc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int flowmc_register_target(const FlowmcTargetVTable* vt) {
  if (!vt || vt->abi_version != FLOWMC_TARGET_ABI || !vt->name) return FLOWMC_ERR_INVALID;
  Registry& r = registry();
  std::lock_guard<std::mutex> lk(r.mu);
  for (size_t i = 0; i < r.targets.size(); ++i) {
    if (std::strcmp(r.targets[i].name, vt->name) == 0) {
      r.targets[i] = *vt;  // re-registration replaces (plugin reloaded)
      return (int)i;
    }
  }
  r.targets.push_back(*vt);
  return (int)r.targets.size() - 1;
}

int flowmc_target_count(void) {
  Registry& r = registry();
  std::lock_guard<std::mutex> lk(r.mu);
  return (int)r.targets.size();
}

int flowmc_target_lookup(const char* name) {
  if (!name) return fail(FLOWMC_ERR_INVALID, "target_lookup: null name");
  Registry& r = registry();
  std::lock_guard<std::mutex> lk(r.mu);
  for (size_t i = 0; i < r.targets.size(); ++i)
    if (std::strcmp(r.targets[i].name, name) == 0) return (int)i;
  return fail(FLOWMC_ERR_NOT_FOUND, std::string("target not registered: ") + name);
}

const char* flowmc_target_name(int id) {
  Registry& r = registry();
  std::lock_guard<std::mutex> lk(r.mu);
  if (id < 0 || id >= (int)r.targets.size()) return nullptr;
  return r.targets[id].name;
}

int flowmc_get_target(int id, FlowmcTargetVTable* out) {
  Registry& r = registry();
  std::lock_guard<std::mutex> lk(r.mu);
  if (id < 0 || id >= (int)r.targets.size()) return fail(FLOWMC_ERR_NOT_FOUND, "invalid target id");
  *out = r.targets[id];
  return FLOWMC_OK;
}

int flowmc_target_eval(int target_id, const float* data, const float* x, int64_t n, int d, float* logp_out,
                       float* grad_out, void* stream) {
  FlowmcTargetVTable vt;
  if (int rc = flowmc_get_target(target_id, &vt)) return rc;
  if (n < 0 || d <= 0 || !x || !logp_out) return fail(FLOWMC_ERR_INVALID, "target_eval: bad arguments");
  if (n == 0) return FLOWMC_OK;
  return vt.eval(data, x, n, d, logp_out, grad_out, (cudaStream_t)stream);
}

int flowmc_key_split(const uint32_t key[2], int64_t num, uint32_t* out) {
  if (!key || !out || num < 0) return fail(FLOWMC_ERR_INVALID, "key_split: bad arguments");
  const flowmc::Key k{key[0], key[1]};
  for (int64_t i = 0; i < num; ++i) {
    const flowmc::Key r = flowmc::split_at(k, (uint64_t)i);
    out[2 * i] = r.k0;
    out[2 * i + 1] = r.k1;
  }
  return FLOWMC_OK;
}

int flowmc_key_split_batch(const uint32_t* keys, int64_t n_keys, int64_t num, uint32_t* out) {
  if (!keys || !out || n_keys < 0 || num < 0) return fail(FLOWMC_ERR_INVALID, "key_split_batch: bad arguments");
  for (int64_t c = 0; c < n_keys; ++c) {
    const flowmc::Key k{keys[2 * c], keys[2 * c + 1]};
    for (int64_t i = 0; i < num; ++i) {
      const flowmc::Key r = flowmc::split_at(k, (uint64_t)i);
      out[2 * (c * num + i)] = r.k0;
      out[2 * (c * num + i) + 1] = r.k1;
    }
  }
  return FLOWMC_OK;
}

int flowmc_random_bits(const uint32_t key[2], int64_t n, uint32_t* out, void* stream) {
  if (!key || n < 0 || (n > 0 && !out)) return fail(FLOWMC_ERR_INVALID, "random_bits: bad arguments");
  if (n == 0) return FLOWMC_OK;
  random_bits_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(flowmc::Key{key[0], key[1]}, n, out);
  flowmc_count_launch();
  return check_launch("random_bits");
}

int flowmc_random_uniform(const uint32_t key[2], int64_t n, float minval, float maxval, float* out, void* stream) {
  if (!key || n < 0 || (n > 0 && !out)) return fail(FLOWMC_ERR_INVALID, "random_uniform: bad arguments");
  if (n == 0) return FLOWMC_OK;
  random_uniform_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(flowmc::Key{key[0], key[1]}, n, minval,
                                                                          maxval, out);
  flowmc_count_launch();
  return check_launch("random_uniform");
}

int flowmc_random_normal(const uint32_t key[2], int64_t n, float* out, void* stream) {
  if (!key || n < 0 || (n > 0 && !out)) return fail(FLOWMC_ERR_INVALID, "random_normal: bad arguments");
  if (n == 0) return FLOWMC_OK;
  random_normal_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(flowmc::Key{key[0], key[1]}, n, out);
  flowmc_count_launch();
  return check_launch("random_normal");
}

int flowmc_pt_exchange(const uint32_t subkey[2], int64_t chain_offset, int64_t n_chains_global, int64_t n_chains,
                       int n_temps, int d, float* positions, float* log_probs, const float* temperatures,
                       float* accepts, void* stream) {
  if (!subkey || n_chains < 0 || n_temps < 1 || d <= 0)
    return fail(FLOWMC_ERR_INVALID, "pt_exchange: bad arguments");
  if (chain_offset < 0 || chain_offset + n_chains > n_chains_global)
    return fail(FLOWMC_ERR_INVALID, "pt_exchange: chain shard outside [0, n_chains_global)");
  if (n_chains == 0 || n_temps == 1) return FLOWMC_OK;
  if (!positions || !log_probs || !temperatures || !accepts) return fail(FLOWMC_ERR_INVALID, "pt_exchange: null buffer");
  pt_exchange_kernel<<<(unsigned)((n_chains + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      flowmc::Key{subkey[0], subkey[1]}, chain_offset, n_chains, n_temps, d, positions, log_probs, temperatures, accepts);
  flowmc_count_launch();
  return check_launch("pt_exchange");
}

int64_t flowmc_local_steps_workspace_bytes(int64_t n_chains, int d, int layout_hint) {
  if (n_chains <= 0 || d <= 0) return 0;
  return flowmc::local_workspace_bytes(n_chains, d, layout_hint);
}

int flowmc_local_steps(int kind, int target_id, const float* target_data, const uint32_t key[2], const float* x0,
                       float* pos_buf, float* lp_buf, float* acc_buf, int64_t n_total, int64_t cursor,
                       int64_t n_chains, int d, int n_steps, int thinning, int64_t chain_offset,
                       int64_t n_chains_global, const FlowmcLocalParams* params, uint32_t key_out[2],
                       float* last_pos, void* stream) {
  FlowmcTargetVTable vt;
  if (int rc = flowmc_get_target(target_id, &vt)) return rc;
  if (!key || !params || !key_out) return fail(FLOWMC_ERR_INVALID, "local_steps: null key/params");
  if (n_chains < 0 || d <= 0 || n_steps < 0 || thinning <= 0)
    return fail(FLOWMC_ERR_INVALID, "local_steps: bad sizes");
  if (chain_offset < 0 || chain_offset + n_chains > n_chains_global)
    return fail(FLOWMC_ERR_INVALID, "local_steps: chain shard outside [0, n_chains_global)");
  const int64_t n_out = (n_steps + thinning - 1) / thinning;
  if (cursor < 0 || cursor + n_out > n_total)
    return fail(FLOWMC_ERR_INVALID, "local_steps: cursor + n_steps/thinning exceeds the buffer length");
  if (params->step_keys && n_steps != 1)
    return fail(FLOWMC_ERR_INVALID, "local_steps: explicit step_keys require n_steps == 1");
  if ((kind == FLOWMC_KERNEL_HMC || kind == FLOWMC_KERNEL_HMC_TEMPERED) && (!params->hmc_chol || !params->hmc_colsum || params->n_leapfrog < 0))
    return fail(FLOWMC_ERR_INVALID, "local_steps: HMC needs hmc_chol, hmc_colsum and n_leapfrog >= 0");

  // take_steps.py:71: rng_key, subkey = split(rng_key)
  const flowmc::Key k{key[0], key[1]};
  const flowmc::Key knew = flowmc::split_at(k, 0);
  const flowmc::Key sub = flowmc::split_at(k, 1);
  key_out[0] = knew.k0;
  key_out[1] = knew.k1;
  if (n_chains == 0 || n_steps == 0) return FLOWMC_OK;
  if (!x0 || !pos_buf || !lp_buf || !acc_buf || !last_pos)
    return fail(FLOWMC_ERR_INVALID, "local_steps: null buffer");

  flowmc::LocalArgs a;
  a.data = target_data;
  a.x0 = x0;
  a.pos_buf = pos_buf;
  a.lp_buf = lp_buf;
  a.acc_buf = acc_buf;
  a.last_pos = last_pos;
  a.n_total = n_total;
  a.cursor = cursor;
  a.n_chains = n_chains;
  a.chain_offset = chain_offset;
  a.d = d;
  a.n_steps = n_steps;
  a.thinning = thinning;
  a.subkey = sub;
  a.step_size = params->step_size;
  a.n_leapfrog = params->n_leapfrog;
  a.hmc_chol = params->hmc_chol;
  a.hmc_colsum = params->hmc_colsum;
  a.hmc_diag = params->hmc_chol_diagonal;
  a.layout_hint = params->layout_hint;
  a.step_keys = params->step_keys;
  a.lp0 = params->lp0;
  a.workspace = params->workspace;
  a.workspace_bytes = params->workspace_bytes;
  a.chain_keys = params->chain_keys;
  a.beta = params->beta;
  a.prior = params->prior;
  a.force_n_seg = params->force_n_seg;
  a.slots_override = params->slots_override;
  a.plan_out = nullptr;
  return vt.local_steps(kind, &a, (cudaStream_t)stream);
}

int flowmc_local_steps_plan(int kind, int target_id, int64_t n_chains, int d, int n_steps,
                            const FlowmcLocalParams* params, int plan[12]) {
  FlowmcTargetVTable vt;
  if (int rc = flowmc_get_target(target_id, &vt)) return rc;
  if (!params || !plan || n_chains <= 0 || d <= 0 || n_steps <= 0)
    return fail(FLOWMC_ERR_INVALID, "local_steps_plan: bad arguments");
  flowmc::LocalArgs a;
  memset(&a, 0, sizeof(a));
  a.n_chains = n_chains;
  a.d = d;
  a.n_steps = n_steps;
  a.thinning = 1;
  a.layout_hint = params->layout_hint;
  a.step_keys = params->step_keys;
  a.workspace = params->workspace;
  a.workspace_bytes = params->workspace_bytes;
  a.force_n_seg = params->force_n_seg;
  a.slots_override = params->slots_override;
  flowmc::LocalPlan p;
  memset(&p, 0, sizeof(p));
  a.plan_out = &p;
  if (int rc = vt.local_steps(kind, &a, nullptr)) return rc;
  const int v[12] = {p.layout, p.G, p.DPL, p.VEC, p.n_groups, p.slots, p.ctas_per_sm, p.smem_per_cta,
                     p.n_seg, p.seg_len, p.n_rounds, p.round_size};
  memcpy(plan, v, sizeof(v));
  return FLOWMC_OK;
}

int flowmc_adam_optimize(int target_id, const float* target_data, const uint32_t key[2], const float* x0,
                         int64_t n_chains, int d, int n_steps, float learning_rate, float noise_level,
                         const float* bounds_lo, const float* bounds_hi, const float* bias_corrections,
                         int64_t chain_offset, int64_t n_chains_global, uint32_t key_out[2], float* x_out,
                         float* logp_out, void* stream) {
  FlowmcTargetVTable vt;
  if (int rc = flowmc_get_target(target_id, &vt)) return rc;
  if (!key || !key_out) return fail(FLOWMC_ERR_INVALID, "adam_optimize: null key");
  if (n_chains < 0 || d <= 0 || n_steps < 0) return fail(FLOWMC_ERR_INVALID, "adam_optimize: bad sizes");
  if (chain_offset < 0 || chain_offset + n_chains > n_chains_global)
    return fail(FLOWMC_ERR_INVALID, "adam_optimize: chain shard outside [0, n_chains_global)");
  if (!vt.adam_opt) return fail(FLOWMC_ERR_UNSUPPORTED, "adam_optimize: target plugin built against an older header");
  // optimization.py:149: rng_key, subkey = split(rng_key)
  const flowmc::Key k{key[0], key[1]};
  const flowmc::Key knew = flowmc::split_at(k, 0);
  const flowmc::Key sub = flowmc::split_at(k, 1);
  key_out[0] = knew.k0;
  key_out[1] = knew.k1;
  if (n_chains == 0) return FLOWMC_OK;
  if (!x0 || !x_out || !bounds_lo || !bounds_hi || (n_steps > 0 && !bias_corrections))
    return fail(FLOWMC_ERR_INVALID, "adam_optimize: null buffer");
  flowmc::AdamOptArgs a;
  a.data = target_data;
  a.x0 = x0;
  a.x_out = x_out;
  a.lp_out = logp_out;
  a.n_chains = n_chains;
  a.chain_offset = chain_offset;
  a.d = d;
  a.n_steps = n_steps;
  a.subkey = sub;
  // optax.adam defaults (optimization.py:61-63): b1 = 0.9, b2 = 0.999, eps = 1e-8; Python-float hyper-parameters
  // meet float32 arrays as float32 scalars
  a.neg_lr = -learning_rate;
  a.noise_level = noise_level;
  a.eps = 1e-8f;
  a.b1 = 0.9f;
  a.b2 = 0.999f;
  a.one_minus_b1 = (float)(1.0 - 0.9);
  a.one_minus_b2 = (float)(1.0 - 0.999);
  a.bc = bias_corrections;
  a.lo = bounds_lo;
  a.hi = bounds_hi;
  return vt.adam_opt(&a, (cudaStream_t)stream);
}

}  // extern "C"
