// Internal ABI between target plugins (instantiated kernel launchers) and the core library.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "rng.cuh"

#define FLOWMC_TARGET_ABI 3

namespace flowmc {

// what the local-steps launcher decided for a call (flowmc_local_steps_plan; bench.py and tests print / assert it)
struct LocalPlan {
  int layout, G, DPL, VEC;
  int n_groups, slots, ctas_per_sm, smem_per_cta;
  int n_seg, seg_len, n_rounds, round_size;
};

// Arguments of one persistent local-steps launch (POD, passed by value to the kernel).
struct LocalArgs {
  const float* data;      // packed target parameters (device)
  const float* x0;        // [n_chains, d]
  float* pos_buf;         // [n_chains, n_total, d]
  float* lp_buf;          // [n_chains, n_total]
  float* acc_buf;         // [n_chains, n_total]
  float* last_pos;        // [n_chains, d]
  int64_t n_total;
  int64_t cursor;
  int64_t n_chains;
  int64_t chain_offset;   // global index of local chain 0
  int d;
  int n_steps;
  int thinning;
  Key subkey;             // chain c uses split(subkey, n_chains_global)[chain_offset + c]
  float step_size;
  int n_leapfrog;
  const float* hmc_chol;  // [d,d] lower-triangular L
  const float* hmc_colsum;// [d]
  int hmc_diag;
  int layout_hint;
  const uint32_t* step_keys;  // optional [n_chains,2]; n_steps must be 1
  const float* lp0;           // optional [n_chains]
  const uint32_t* chain_keys; // optional [n_chains,2]: initial per-chain keys instead of split(subkey)[c]
  const float* beta;          // KIND_MALA_PT: [n_chains] inverse temperatures (NULL = 1)
  const float* prior;         // KIND_MALA_PT: [4][d] = c, m, lo, hi of the quadratic / box log-prior (NULL = 0)
  void* workspace;            // optional: time-slicing scratch (see local_steps.cuh)
  int64_t workspace_bytes;
  int force_n_seg;            // 0 = auto; > 1: cut the step range into this many segments; < 0: never slice
  int slots_override;         // 0 = the device's resident CTA slots; > 0: pretend there are this many (tests)
  LocalPlan* plan_out;        // host, optional: fill in the launch plan and return WITHOUT launching
};

// arguments of the AdamOptimization kernel (adam_opt.cuh)
struct AdamOptArgs {
  const float* data;      // packed target parameters
  const float* x0;        // [n_chains, d]
  float* x_out;           // [n_chains, d]
  float* lp_out;          // [n_chains] logpdf(x_out), optional
  int64_t n_chains;
  int64_t chain_offset;   // global index of local chain 0
  int d;
  int n_steps;
  Key subkey;             // chain c uses split(subkey, n_chains_global)[chain_offset + c]
  float neg_lr, noise_level, eps, b1, b2, one_minus_b1, one_minus_b2;
  const float* bc;        // device [n_steps, 2]
  const float* lo;        // device [d] lower bounds (-inf allowed)
  const float* hi;        // device [d]
};


}  // namespace flowmc

extern "C" {
typedef struct FlowmcTargetVTable {
  int abi_version;
  const char* name;
  // returns 0, or a negative FLOWMC_ERR_* code (message via flowmc_set_error)
  int (*local_steps)(int kind, const flowmc::LocalArgs* args, cudaStream_t stream);
  int (*eval)(const float* data, const float* x, int64_t n, int d, float* logp_out, float* grad_out,
              cudaStream_t stream);
  int (*adam_opt)(const flowmc::AdamOptArgs* args, cudaStream_t stream);  // AdamOptimization (ABI >= 2)
} FlowmcTargetVTable;

__attribute__((visibility("default"))) int flowmc_register_target(const FlowmcTargetVTable* vt);
__attribute__((visibility("default"))) void flowmc_set_error(const char* msg);
__attribute__((visibility("default"))) void flowmc_count_launch(void);
// copies the vtable of a registered target (0, or FLOWMC_ERR_NOT_FOUND with the error message set)
__attribute__((visibility("default"))) int flowmc_get_target(int target_id, FlowmcTargetVTable* out);
}
