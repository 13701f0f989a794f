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

#define FLOWMC_ABI_VERSION 3

#if defined(__GNUC__)
#define FLOWMC_API __attribute__((visibility("default")))
#else
#define FLOWMC_API
#endif

/* local kernel kinds (resource/kernel/{MALA,HMC,Gaussian_random_walk}.py) */
#define FLOWMC_KERNEL_MALA 0
#define FLOWMC_KERNEL_HMC 1
#define FLOWMC_KERNEL_GRW 2
#define FLOWMC_KERNEL_MALA_TEMPERED 3 /* MALA on beta_c * logpdf + log_prior (ParallelTempering) */
#define FLOWMC_KERNEL_HMC_TEMPERED 4  /* HMC on the tempered density */
#define FLOWMC_KERNEL_GRW_TEMPERED 5  /* Gaussian random walk on the tempered density */

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
/* vmapped split: out[c, i] = jax.random.split(keys[c], num)[i]; keys host [n_keys, 2], out host [n_keys, num, 2] */
FLOWMC_API int flowmc_key_split_batch(const uint32_t* keys, int64_t n_keys, int64_t num, uint32_t* out);
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
  const uint32_t* chain_keys; /* optional, device [n_chains,2]: the chains' initial keys; NULL = split(subkey,
                               * n_chains_global)[global chain index] (take_steps.py:72).  ParallelTempering passes
                               * split(split(subkey, n_chains)[c], n_temps)[t] (parallel_tempering.py:289-293) */
  const float* beta;      /* FLOWMC_KERNEL_*_TEMPERED: device [n_chains] inverse temperatures 1 / T (NULL = 1) */
  const float* prior;     /* FLOWMC_KERNEL_*_TEMPERED: device [4, d] = c, m, lo, hi of
                           * log_prior(x) = -sum_j c_j (x_j - m_j)^2 inside [lo, hi], -inf outside; NULL = flat 0 */
  int force_n_seg;        /* time slicing: <= 0 = one launch for the whole call (default: measured fastest on B200,
                           * profiles/r02_slice_sweep.jsonl); > 1 = cut the n_steps into this many segments that run
                           * as successive resident waves.  Results do not depend on it (tests/test_gpu_local_sliced.py) */
  int slots_override;     /* 0 = the device's resident CTA slots; > 0 = plan the rounds as if there were this many
                           * (test hook: forces multi-round launches with a handful of chains) */
} FlowmcLocalParams;

/* bytes of `workspace` that flowmc_local_steps can use for (n_chains, d, layout_hint) */
FLOWMC_API int64_t flowmc_local_steps_workspace_bytes(int64_t n_chains, int d, int layout_hint);

/* The launch plan flowmc_local_steps would use for (kind, target, n_chains, d, n_steps, params), without launching:
 * plan[12] (host) = {layout index, G lanes per chain, DPL dims per lane, VEC store width, chain groups, resident CTA
 * slots, CTAs per SM, static shared memory per CTA, n_seg, seg_len, n_rounds (= kernel launches), CTAs per round}. */
FLOWMC_API int flowmc_local_steps_plan(int kind, int target_id, int64_t n_chains, int d, int n_steps,
                                       const FlowmcLocalParams* params, int plan[12]);

/* Runs n_steps of one local kernel for n_chains chains in ONE persistent kernel launch (or, with
 * params->force_n_seg > 1, one launch per resident wave of (time segment, chain group) work items) and
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

/* ---- MaskedCouplingRQSpline flow (resource/model/nf_model/rqSpline.py:392-504) ---------------
 * The model is ONE flat float32 parameter blob on the device (so that the optimiser and the
 * gradient all-reduce see a single vector) described by FlowmcFlowDesc.  Offsets are in floats;
 * every array starts on a 4-float boundary (padding is zero and stays zero under AdamW).
 *   per layer l at l*layer_stride:  for i in 0..n_linear-1: W_i [out_i, in_i] row-major (equinox
 *     nn.Linear layout, common.py:95-101), b_i [out_i];  then scale, shift (ScalarAffine, trainable)
 *   tail: data_mean [d], data_cov [d,d], base mean [d], base cov [d,d]
 * Linear i maps dims[i] -> dims[i+1] with dims = [d, hidden..., d*(3*num_bins+1)].
 * Layer l transforms the features f with (f + l) % 2 == 0 (mask False, rqSpline.py:434). */
#define FLOWMC_FLOW_MAX_LINEAR 4
typedef struct FlowmcFlowDesc {
  int n_features, n_layers, n_linear, num_bins;
  int dims[FLOWMC_FLOW_MAX_LINEAR + 1];
  float range_min, range_max;
  int64_t off_W[FLOWMC_FLOW_MAX_LINEAR], off_b[FLOWMC_FLOW_MAX_LINEAR], off_scale, off_shift;
  int64_t layer_stride;
  int64_t off_data_mean, off_data_cov, off_base_mean, off_base_cov;
  int64_t n_params; /* total floats in the blob */
  /* execution path: tc_image == NULL or tc_terms == 0 -> fp32 CUDA-core kernels; otherwise the tcgen05 kernels
   * read the conditioner weights from tc_image (device, flowmc_flow_tc_image_bytes() bytes, refreshed with
   * flowmc_flow_tc_pack() whenever params change).  tc_terms = 3: 3xTF32 (fp32-grade, inside the 1e-5 parity
   * tolerance), 1: plain TF32 (about 1e-3 relative in the spline parameters). */
  const void* tc_image;
  int tc_terms;
} FlowmcFlowDesc;

/* fills `desc` for (n_features, n_layers, hidden[n_hidden], num_bins, spline range) */
FLOWMC_API int flowmc_flow_desc_init(FlowmcFlowDesc* desc, int n_features, int n_layers, int n_hidden,
                                     const int* hidden, int num_bins, float range_min, float range_max);

/* tensor-core weight image: size for a model shape (0 if the shape is not supported by the tcgen05 path:
 * n_features <= 128, hidden widths multiples of 16 and <= 128, num_bins in {4, 8, 16}) and the packing kernel
 * (splits every weight into tf32 hi + lo and lays the rows out as SWIZZLE_128B K-major UMMA stages) */
FLOWMC_API int64_t flowmc_flow_tc_image_bytes(const FlowmcFlowDesc* desc);
FLOWMC_API int flowmc_flow_tc_pack(const FlowmcFlowDesc* desc, const float* params, void* image, void* stream);

/* forward / inverse of the bijection on n rows: x device [n,d] -> y device [n,d], logdet device [n]
 * (MaskedCouplingRQSpline.forward / .inverse, rqSpline.py:450-488; no whitening) */
FLOWMC_API int flowmc_flow_forward(const FlowmcFlowDesc* desc, const float* params, const float* x, int64_t n,
                                   float* y, float* logdet, void* stream);
FLOWMC_API int flowmc_flow_inverse(const FlowmcFlowDesc* desc, const float* params, const float* x, int64_t n,
                                   float* y, float* logdet, void* stream);
/* log_prob(x) = logdet(forward((x - data_mean)/sqrt(diag data_cov))) + base.log_prob  (rqSpline.py:498-504);
 * layer_inputs (optional, device [n_layers + 1, n, d]) receives each layer's input and, in the last slot,
 * the final latent (what the training backward pass starts from) */
FLOWMC_API int flowmc_flow_log_prob(const FlowmcFlowDesc* desc, const float* params, const float* x, int64_t n,
                                    float* log_prob, float* layer_inputs, void* stream);
/* sample: row r draws z = normal(key_k, (rows_per_key, d))[r % rows_per_key] with k = r / rows_per_key, then
 * x = inverse(base_mean + sqrt(diag base_cov) z) * sqrt(diag data_cov) + data_mean  (rqSpline.py:490-496).
 * keys: device uint32 [ceil(n / rows_per_key), 2], or NULL to use the single host key `host_key`. */
FLOWMC_API int flowmc_flow_sample(const FlowmcFlowDesc* desc, const float* params, const uint32_t* keys,
                                  const uint32_t host_key[2], int64_t rows_per_key, int64_t n, float* x_out,
                                  void* stream);

/* ---- NFProposal global steps (strategy/take_steps.py:191-206 + resource/kernel/NF_proposal.py:27-172) ---- */
typedef struct FlowmcGlobalParams {
  int n_batch_size;           /* NFProposal.n_batch_size: only changes the proposals' key schedule */
  const uint32_t* chain_keys; /* optional, device [n_chains,2]: explicit per-chain rng_key (NFProposal.kernel
                               * called directly); NULL = split(subkey, n_chains_global)[global chain index] */
  const float* lp0;           /* optional, device [n_chains]: incoming log_prob; NULL = logpdf(x0) (take_steps.py:201) */
  void* workspace;            /* device scratch: proposals and their log-probs; size from
                               * flowmc_nf_global_steps_workspace_bytes() */
  int64_t workspace_bytes;
} FlowmcGlobalParams;

FLOWMC_API int64_t flowmc_nf_global_steps_workspace_bytes(int64_t n_chains, int d, int n_steps);

/* n_steps independence-MH steps with flow proposals for n_chains chains; buffers, cursor, thinning, key,
 * key_out, last_pos and chain sharding exactly as flowmc_local_steps.  Proposals, their flow log-probs
 * (a forward pass, as the reference does), the target log-probs and the sequential accept scan all run on
 * the device; nothing is read back. */
FLOWMC_API int flowmc_nf_global_steps(const FlowmcFlowDesc* desc, const float* params, int target_id,
                                      const float* target_data, const uint32_t key[2], const float* x0,
                                      float* pos_buf, float* lp_buf, float* acc_buf, int64_t n_total, int64_t cursor,
                                      int64_t n_chains, int n_steps, int thinning, int64_t chain_offset,
                                      int64_t n_chains_global, const FlowmcGlobalParams* params_g,
                                      uint32_t key_out[2], float* last_pos, void* stream);

/* ---- flow training (resource/model/nf_model/base.py:98-210, resource/optimizer.py:19-23) ------------------ */
FLOWMC_API int64_t flowmc_flow_loss_grad_workspace_bytes(const FlowmcFlowDesc* desc, int64_t n);
/* NFModel.loss_fn + its gradient (base.py:98-100,122) for the rows x[idx[0..n)] (idx NULL = rows 0..n):
 *   loss[0]  = -sum_i log_prob(x_i) * inv_n_total,   grad[0..n_params) = d loss / d params
 * (both device, OVERWRITTEN).  inv_n_total = 1 / (global batch size): data-parallel ranks each pass their slice
 * of the batch and sum-all-reduce grad and loss.  The non-trainable tail (data_mean, data_cov, base mean/cov)
 * gets zero gradient, as under the reference's stop_gradient. */
FLOWMC_API int flowmc_flow_loss_grad(const FlowmcFlowDesc* desc, const float* params, const float* x,
                                     const int32_t* idx, int64_t n, float inv_n_total, float* grad, float* loss,
                                     void* workspace, int64_t workspace_bytes, void* stream);
/* optax.chain(clip_by_global_norm(max_norm), adamw(lr, b1, b2, eps, weight_decay)) applied in place to the flat
 * vectors (all device, n_params floats): params, Adam moments mu / nu.  count = 1-based step number (optax's
 * count after increment); hyperparameters are doubles because optax holds them as Python floats (1 - b1 is
 * formed in double before rounding to float32); scratch: device, >= 256 floats; gnorm_out: optional device float (pre-clip norm). */
FLOWMC_API int flowmc_clip_adamw(int64_t n_params, float* params, const float* grads, float* mu, float* nu,
                                 int64_t count, double lr, double b1, double b2, double eps, double weight_decay,
                                 double max_norm, float* scratch, float* gnorm_out, void* stream);

/* ---- RealNVP (resource/model/nf_model/realNVP.py:102-228) ---------------------------------
 * n_layers x MaskedCouplingLayer(MLPAffine(scale_MLP, shift_MLP), mask), both MLPs [d, n_hidden, d] with relu
 * (common.py:68-124 default activation).  ONE flat float32 blob like the spline flow:
 *   per layer l at l*layer_stride: W1s [h,d], b1s [h], W2s [d,h], b2s [d] (scale MLP), W1t, b1t, W2t, b2t (shift
 *     MLP), mask [d] -- the reference's coupling mask is a FLOAT leaf (1 = conditioning / unchanged, 0 = transformed)
 *     that gets zero gradient but AdamW weight decay; the kernels evaluate the layer with the general float mask
 *   tail: data_mean [d], data_cov [d,d], base mean [d], base cov [d,d]
 * dt: MLPAffine / AffineCoupling's scaling factor (scale = tanh(.) * dt, shift = (.) * dt; RealNVP uses 1). */
typedef struct FlowmcRealNVPDesc {
  int n_features, n_layers, n_hidden;
  float dt;
  int64_t off_W1s, off_b1s, off_W2s, off_b2s, off_W1t, off_b1t, off_W2t, off_b2t, off_mask;
  int64_t layer_stride;
  int64_t off_data_mean, off_data_cov, off_base_mean, off_base_cov;
  int64_t n_params;
} FlowmcRealNVPDesc;

FLOWMC_API int flowmc_realnvp_desc_init(FlowmcRealNVPDesc* desc, int n_features, int n_layers, int n_hidden, float dt);
/* RealNVP.forward / .inverse (realNVP.py:172-206): x device [n,d] -> y device [n,d], logdet device [n] */
FLOWMC_API int flowmc_realnvp_forward(const FlowmcRealNVPDesc* desc, const float* params, const float* x, int64_t n,
                                      float* y, float* logdet, void* stream);
FLOWMC_API int flowmc_realnvp_inverse(const FlowmcRealNVPDesc* desc, const float* params, const float* x, int64_t n,
                                      float* y, float* logdet, void* stream);
/* RealNVP.log_prob (realNVP.py:214-221): whitening, forward, + multivariate_normal.logpdf(y, zeros, eye) */
FLOWMC_API int flowmc_realnvp_log_prob(const FlowmcRealNVPDesc* desc, const float* params, const float* x, int64_t n,
                                       float* log_prob, void* stream);
/* RealNVP.sample (realNVP.py:208-212); keys / host_key / rows_per_key as in flowmc_flow_sample */
FLOWMC_API int flowmc_realnvp_sample(const FlowmcRealNVPDesc* desc, const float* params, const uint32_t* keys,
                                     const uint32_t host_key[2], int64_t rows_per_key, int64_t n, float* x_out,
                                     void* stream);
/* NFModel.loss_fn + gradient for a RealNVP (same contract as flowmc_flow_loss_grad; the masks and the tail get zero
 * gradient).  Weight gradients are accumulated with float atomics (not bit-reproducible run to run). */
FLOWMC_API int64_t flowmc_realnvp_loss_grad_workspace_bytes(const FlowmcRealNVPDesc* desc, int64_t n);
FLOWMC_API int flowmc_realnvp_loss_grad(const FlowmcRealNVPDesc* desc, const float* params, const float* x,
                                        const int32_t* idx, int64_t n, float inv_n_total, float* grad, float* loss,
                                        void* workspace, int64_t workspace_bytes, void* stream);

/* The sequential accept scan of NFProposal.kernel (NF_proposal.py:91-126) on its own, for proposal models other than
 * the fused spline path: chain c continues with rng_key = split(chain_keys[c])[0]; per step key, sub = split(key),
 * accept when log(uniform(sub)) < (lp_prop - lp) - (lp_nf_prop - lp_nf).  chain_keys device [n_chains,2]; x0 device
 * [n_chains,d]; lp0, lp_nf_cur device [n_chains]; props device [n_chains,n_steps,d]; lp_prop, lp_nf_prop device
 * [n_chains,n_steps]; thinned results go into the sampler buffers at `cursor` as in flowmc_nf_global_steps. */
FLOWMC_API int flowmc_nf_accept_scan(const uint32_t* chain_keys, int64_t n_chains, int d, int n_steps, int thinning,
                                     const float* x0, const float* lp0, const float* lp_nf_cur, const float* props,
                                     const float* lp_prop, const float* lp_nf_prop, float* pos_buf, float* lp_buf,
                                     float* acc_buf, int64_t n_total, int64_t cursor, float* last_pos, void* stream);

/* ---- data-parallel optimiser step over NVLink peer memory (csrc/peer_reduce.cu) ------------------------------
 * One kernel = all-reduce of the ranks' gradients (reduce-scatter through peer loads, all-gather through peer stores,
 * sums in rank order) + clip_by_global_norm + AdamW on the full vector, replacing NCCL all-reduce + flowmc_clip_adamw
 * in NFModel.train_step when the ranks of one box train data-parallel.  Every rank owns one exchange block
 * (flowmc_peer_block_bytes; layout through flowmc_peer_block_offset: 0 = grad [n_params + 4, loss at index n_params],
 * 1 = reduced gradient, 2 = slice norms, 3 = flags, 4 = local counters) allocated with flowmc_ipc_alloc and opened
 * by every peer with flowmc_ipc_open (CUDA IPC; the 64-byte handles travel over torch.distributed).  blocks[k] =
 * rank k's block as mapped into THIS process (own block: the local pointer).  epoch = 1, 2, 3, ...: the same on every
 * rank, one per call.  Every rank must make the call (it spins on its peers' flags). */
FLOWMC_API int64_t flowmc_peer_block_bytes(int64_t n_params);
FLOWMC_API int64_t flowmc_peer_block_offset(int64_t n_params, int what);
FLOWMC_API int flowmc_ipc_alloc(int64_t bytes, void** ptr, unsigned char handle[64]);
FLOWMC_API int flowmc_ipc_open(const unsigned char handle[64], void** ptr);
FLOWMC_API int flowmc_ipc_close(void* ptr);
FLOWMC_API int flowmc_ipc_free(void* ptr);
FLOWMC_API int flowmc_dp_reduce_adamw(int rank, int world, void* const* blocks, int64_t n_params, float* params,
                                      float* mu, float* nu, int64_t count, double lr, double b1, double b2, double eps,
                                      double weight_decay, double max_norm, uint32_t epoch, float* loss_out,
                                      void* stream);

/* ---- training-set plumbing (strategy/train_model.py:66-81, nf_model/base.py:141-144,187-188) --------------- */
/* jax.random.permutation(key, n) -> out device int32[n] */
FLOWMC_API int64_t flowmc_random_permutation_workspace_bytes(int64_t n);
FLOWMC_API int flowmc_random_permutation(const uint32_t key[2], int64_t n, int32_t* out, void* workspace,
                                         int64_t workspace_bytes, void* stream);
/* jax.random.choice(key, arange(n_population), (m,), replace=True) -> out device int32[m] */
FLOWMC_API int flowmc_random_choice(const uint32_t key[2], int64_t n_population, int64_t m, int32_t* out,
                                    void* stream);
/* per chain of buf [n_chains, n_total, d]: rowmap[c][k] = step of the k-th finite row, counts[c] = number of
 * finite rows, minmax = {min, max} over chains of counts (all device int32) */
FLOWMC_API int flowmc_buffer_finite_rows(const float* buf, int64_t n_chains, int64_t n_total, int d, int32_t* rowmap,
                                         int32_t* counts, int32_t* minmax, void* stream);
/* out[i] = row idx[i] of the population "last `window` of each chain's m_finite finite rows" (row q belongs to
 * global chain q / window); only rows of chains in [chain_lo, chain_hi) (this rank's slab buf) are written */
FLOWMC_API int flowmc_gather_training_rows(const float* buf, const int32_t* rowmap, int64_t n_total, int d,
                                           int window, int m_finite, int64_t chain_lo, int64_t chain_hi,
                                           const int32_t* idx, int64_t m, float* out, void* stream);
/* jnp.mean(x, 0) and jnp.cov(x.T) of x device [n, d]; scratch: device, >= d floats */
FLOWMC_API int flowmc_data_mean_cov(const float* x, int64_t n, int d, float* mean, float* cov, float* scratch,
                                    void* stream);

/* ---- ParallelTempering exchange step (src/flowMC/strategy/parallel_tempering.py:291-398) ----------------- */
/* For every chain, in ladder order idx = 0 .. n_temps - 2: key, sub = split(key) (key = split(subkey,
 * n_chains_global)[global chain]); swap rungs idx and idx + 1 (positions AND log_probs) when
 * log(uniform(sub)) < (1 / T[idx+1] - 1 / T[idx]) * (log_probs[idx] - log_probs[idx+1]).  In place.
 * positions: device [n_chains, n_temps, d]; log_probs: device [n_chains, n_temps] UNtempered logpdf(positions);
 * temperatures: device [n_temps]; accepts: device [n_chains, n_temps - 1] (0 / 1).  The tempered individual steps
 * that precede it are flowmc_local_steps(FLOWMC_KERNEL_MALA_TEMPERED) with chain_keys / beta / prior. */
FLOWMC_API int flowmc_pt_exchange(const uint32_t subkey[2], int64_t chain_offset, int64_t n_chains_global,
                                  int64_t n_chains, int n_temps, int d, float* positions, float* log_probs,
                                  const float* temperatures, float* accepts, void* stream);

/* ---- AdamOptimization (src/flowMC/strategy/optimization.py:85-164) ---------------------------------------- */
/* n_steps of optax.adam(learning_rate) on -logpdf for every chain, gradient multiplied by (1 + normal * noise_level),
 * box projection after every step; one launch, chain state in registers.  key: the strategy's rng_key (host);
 * key_out: the key the strategy returns.  x0 / x_out: device [n_chains, d]; bounds_lo / bounds_hi: device [d];
 * bias_corrections: device [n_steps, 2] float32 = {1 - 0.9^(t+1), 1 - 0.999^(t+1)}; logp_out: device [n_chains] or
 * NULL.  chain_offset / n_chains_global: this process's shard of the global chain index space (keys come from the
 * global split, so any sharding gives the same result). */
FLOWMC_API int flowmc_adam_optimize(int target_id, const float* target_data, const uint32_t key[2], const float* x0,
                                    int64_t n_chains, int d, int n_steps, float learning_rate, float noise_level,
                                    const float* bounds_lo, const float* bounds_hi, const float* bias_corrections,
                                    int64_t chain_offset, int64_t n_chains_global, uint32_t key_out[2], float* x_out,
                                    float* logp_out, void* stream);

/* ---- tracing ---------------------------------------------------------------------------------------------- */
/* pipeline timeline of the tensor-core flow kernels (the reference's only tracing is its trace-time "Compiling ..."
 * prints): CTA 0 of subsequent launches stamps clock64() at its pipeline events into buf (device,
 * 4 * 256 int64: weight producer / MMA issuer / epilogue thread 0 of the forward kernel, epilogue thread 0 of the
 * tensor-core backward kernel); NULL switches it off.  scripts/tc_timeline.py and scripts/bt_timeline.py print it. */
FLOWMC_API void flowmc_trace_tc_timeline(long long* buf);

/* number of kernel launches issued by this library since load (for bench.py's gpu_launches) */
FLOWMC_API int64_t flowmc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FLOWMC_B200_H */
