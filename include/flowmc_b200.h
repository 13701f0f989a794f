/* flowmc_b200 -- C ABI of the B200-native flowMC sampling hot path.
 *
 * This is the drop-in boundary: every entry point below is what a jax.ffi / ctypes / torch
 * binding of the reference's hot path binds to.  All functions are extern "C", take plain
 * pointers and sizes, never allocate device memory, never synchronise the device and only
 * enqueue work on the given stream (a cudaStream_t passed as void*).  Return 0 on success, <0 on
 * error; flowmc_last_error() returns a thread-local message.  "device" pointers are device
 * memory owned by the caller; "host" pointers are ordinary host memory read/written
 * synchronously inside the call.
 *
 * Reference interfaces replaced (paths relative to the flowMC tree, v0.4.5):
 *   flowmc_local_steps        TakeSteps.__call__ + TakeSerialSteps.sample/body
 *                             (src/flowMC/strategy/take_steps.py:60-144,156-180) driving
 *                             MALA.kernel (resource/kernel/MALA.py:26-89), HMC.kernel
 *                             (resource/kernel/HMC.py:98-151) or GaussianRandomWalk.kernel
 *                             (resource/kernel/Gaussian_random_walk.py:25-61), including the
 *                             three Buffer.update_buffer writes (resource/buffers.py:32-41).
 *   flowmc_target_eval        LogPDF.__call__ and jax.value_and_grad(logpdf)
 *                             (resource/logPDF.py:60-61, MALA.py:59).
 *   flowmc_key_* / random_*   jax.random.split / bits / uniform / normal as used by the
 *                             strategies (take_steps.py:71-72, train_model.py:72-81) and by user
 *                             scripts that draw initial positions.
 *   flowmc_flow_*             MaskedCouplingRQSpline.forward/inverse/log_prob/sample
 *                             (resource/model/nf_model/rqSpline.py:392-504).
 *   flowmc_nf_global_steps    TakeGroupSteps.sample + NFProposal.kernel
 *                             (strategy/take_steps.py:191-206, resource/kernel/NF_proposal.py:27-172).
 *   flowmc_flow_loss_grad,    NFModel.loss_fn/train_step (resource/model/nf_model/base.py:98-125)
 *   flowmc_clip_adamw         and Optimizer (resource/optimizer.py:19-23).
 */
#ifndef FLOWMC_B200_H
#define FLOWMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLOWMC_ABI_VERSION 1

#if defined(__GNUC__)
#define FLOWMC_API __attribute__((visibility("default")))
#else
#define FLOWMC_API
#endif

/* local kernel kinds (resource/kernel/{MALA,HMC,Gaussian_random_walk}.py) */
#define FLOWMC_KERNEL_MALA 0
#define FLOWMC_KERNEL_HMC 1
#define FLOWMC_KERNEL_GRW 2

/* error codes */
#define FLOWMC_OK 0
#define FLOWMC_ERR_INVALID (-1)
#define FLOWMC_ERR_UNSUPPORTED (-2)
#define FLOWMC_ERR_CUDA (-3)
#define FLOWMC_ERR_NOT_FOUND (-4)

FLOWMC_API int flowmc_abi_version(void);
FLOWMC_API const char* flowmc_last_error(void);

/* ---- target registry (targets are compiled device functions, see flowmc_target.cuh) ---- */
FLOWMC_API int flowmc_target_count(void);
/* returns the id (>=0) of a registered target, FLOWMC_ERR_NOT_FOUND otherwise */
FLOWMC_API int flowmc_target_lookup(const char* name);
FLOWMC_API const char* flowmc_target_name(int target_id);

/* logp (and optionally grad) of n points: x device [n,d], data device (packed, target-defined),
 * logp_out device [n], grad_out device [n,d] or NULL. */
FLOWMC_API int flowmc_target_eval(int target_id, const float* data, const float* x, int64_t n, int d,
                       float* logp_out, float* grad_out, void* stream);

/* ---- jax.random-compatible key management (host, synchronous, tiny) ---- */
/* split(key, num) -> out host uint32[num][2] */
FLOWMC_API int flowmc_key_split(const uint32_t key[2], int64_t num, uint32_t* out);
/* ---- jax.random-compatible draws written to device memory ---- */
FLOWMC_API int flowmc_random_bits(const uint32_t key[2], int64_t n, uint32_t* out, void* stream);
FLOWMC_API int flowmc_random_uniform(const uint32_t key[2], int64_t n, float minval, float maxval, float* out, void* stream);
FLOWMC_API int flowmc_random_normal(const uint32_t key[2], int64_t n, float* out, void* stream);

/* ---- local steps ---- */
typedef struct FlowmcLocalParams {
  float step_size;        /* MALA, GRW, HMC */
  int n_leapfrog;         /* HMC */
  const float* hmc_chol;  /* HMC: device [d,d] row-major L = chol(inv(condition_matrix)) */
  const float* hmc_colsum;/* HMC: device [d] column sums of condition_matrix */
  int hmc_chol_diagonal;  /* HMC: 1 if L is diagonal (fast path, identical results) */
  int layout_hint;        /* 0 = auto; otherwise index+1 into the launcher's layout table */
  const uint32_t* step_keys; /* optional, device [n_chains,2]: explicit per-chain keys for a single
                              * ProposalBase.kernel() call (resource/kernel/base.py:16-27); requires
                              * n_steps == 1; the key is used exactly as kernel()'s rng_key argument */
  const float* lp0;       /* optional, device [n_chains]: incoming log_prob (kernel()'s log_prob
                           * argument); NULL = logpdf(x0) as in take_steps.py:177 */
  void* workspace;        /* optional, device: scratch for time slicing (chain-state hand-off between
                           * CTAs when there are more chain groups than resident slots); size from
                           * flowmc_local_steps_workspace_bytes().  NULL = plain one-CTA-per-group grid */
  int64_t workspace_bytes;
} FlowmcLocalParams;

/* bytes of `workspace` that flowmc_local_steps can use for (n_chains, d, layout_hint) */
FLOWMC_API int64_t flowmc_local_steps_workspace_bytes(int64_t n_chains, int d, int layout_hint);

/* Runs n_steps of one local kernel for n_chains chains in ONE persistent kernel launch and
 * writes the thinned samples straight into the (chain-major) sampler buffers at `cursor`:
 *   pos_buf  device [n_chains, n_total, d]      lp_buf, acc_buf  device [n_chains, n_total]
 * key:      host, the strategy-level rng_key; key_out: host, the rng_key the strategy returns.
 * x0:       device [n_chains, d]                last_pos: device [n_chains, d] (positions[:, -1])
 * chain_offset / n_chains_global: this call owns global chains [chain_offset, chain_offset+n_chains)
 *           of n_chains_global; per-chain keys are split(subkey, n_chains_global)[global index], so
 *           any sharding over processes gives bit-identical chains. */
FLOWMC_API int flowmc_local_steps(int kind, int target_id, const float* target_data, const uint32_t key[2],
                       const float* x0, float* pos_buf, float* lp_buf, float* acc_buf,
                       int64_t n_total, int64_t cursor, int64_t n_chains, int d, int n_steps, int thinning,
                       int64_t chain_offset, int64_t n_chains_global, const FlowmcLocalParams* params,
                       uint32_t key_out[2], float* last_pos, void* stream);

/* number of kernel launches issued by this library since load (for bench.py's gpu_launches) */
FLOWMC_API int64_t flowmc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FLOWMC_B200_H */
