// flowmc_xla_ffi.cc -- XLA FFI (jax.ffi) handlers over the C ABI of libflowmc_b200.so.
//
// This is the thin adapter BASELINE.json's north_star names: flowMC's Python strategies keep calling JAX, and the
// `eqx.filter_jit(eqx.filter_vmap(self.sample))` expression of src/flowMC/strategy/take_steps.py:127-142 (and the
// corresponding expressions listed beside each handler below) becomes one `jax.ffi.ffi_call`.  Every handler only
// unpacks XLA buffers / attributes and forwards to ONE function of include/flowmc_b200.h on XLA's stream; nothing is
// computed here.
//
// Build (in an environment that has jaxlib >= 0.4.31; its headers are header-only):
//   g++ -O2 -std=c++17 -shared -fPIC -I$(python -c "import jax; print(jax.ffi.include_dir())") -Iinclude \
//       -I/usr/local/cuda/include flowmc_b200/csrc/flowmc_xla_ffi.cc -Lflowmc_b200/lib -lflowmc_b200 \
//       -Wl,-rpath,'$ORIGIN' -o flowmc_b200/lib/libflowmc_xla_ffi.so
// Python side:
//   jax.ffi.register_ffi_target("flowmc_local_steps", jax.ffi.pycapsule(lib.FlowmcLocalSteps), platform="CUDA")
//
// jaxlib is NOT installable in this image (no wheel, no network), so `xla/ffi/api/ffi.h` is absent: the whole file is
// guarded by __has_include and compiles to an empty translation unit here.  tests/test_abi.py compiles it against a
// minimal mock of the XLA FFI surface (tests/mock_xla/) so that the calls into the C ABI stay type-checked.
//
// Conventions shared by all handlers:
//   * PRNG keys are jax.random.key_data(key) words passed as uint32 ATTRIBUTES (the strategies' __call__ is not jitted
//     in flowMC -- only the inner sample() is -- so the key is concrete); the key the strategy returns is
//     jax.random.split(key)[0], computed in JAX (== key_out of the C call);
//   * sample buffers are donated (input_output_aliases) so the C side's in-place store at `cursor` replaces the three
//     Buffer.update_buffer copies (src/flowMC/resource/buffers.py:32-41);
//   * scratch memory comes from ffi::ScratchAllocator, sized by the library's *_workspace_bytes() functions;
//   * a negative return code becomes ffi::Error(kInternal, flowmc_last_error()).
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define FLOWMC_HAVE_XLA_FFI 1
#endif
#endif

#ifdef FLOWMC_HAVE_XLA_FFI
#include <cstdint>
#include <cstring>
#include <cuda_runtime_api.h>

#include "xla/ffi/api/ffi.h"

#include "flowmc_b200.h"

namespace ffi = xla::ffi;

namespace {

using F32 = ffi::Buffer<ffi::F32>;
using U32 = ffi::Buffer<ffi::U32>;
using S32 = ffi::Buffer<ffi::S32>;
using RF32 = ffi::ResultBuffer<ffi::F32>;
using RS32 = ffi::ResultBuffer<ffi::S32>;
using Stream = ffi::PlatformStream<cudaStream_t>;

inline ffi::Error done(int rc) {
  if (rc >= 0) return ffi::Error::Success();
  return ffi::Error(ffi::ErrorCode::kInternal, flowmc_last_error());
}

// an optional operand is passed as a zero-element buffer
template <class B>
inline auto opt(B& b) -> decltype(b.typed_data()) {
  return b.element_count() == 0 ? nullptr : b.typed_data();
}

inline ffi::Error scratch(ffi::ScratchAllocator& alloc, int64_t bytes, void** out) {
  *out = nullptr;
  if (bytes <= 0) return ffi::Error::Success();
  auto p = alloc.Allocate((size_t)bytes);
  if (!p.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "flowmc: scratch allocation failed");
  *out = *p;
  return ffi::Error::Success();
}

// FlowmcFlowDesc from attributes (MaskedCouplingRQSpline.__init__ arguments, rqSpline.py:392-401); tc_image is the
// packed tensor-core weight image operand (zero elements = fp32 CUDA-core path)
inline int make_desc(FlowmcFlowDesc* desc, int32_t n_features, int32_t n_layers, ffi::Span<const int32_t> hidden,
                     int32_t num_bins, float range_min, float range_max, const void* tc_image, int32_t tc_terms) {
  int h[FLOWMC_FLOW_MAX_LINEAR];
  const int nh = (int)hidden.size();
  for (int i = 0; i < nh && i < FLOWMC_FLOW_MAX_LINEAR; ++i) h[i] = hidden[i];
  int rc = flowmc_flow_desc_init(desc, n_features, n_layers, nh, h, num_bins, range_min, range_max);
  desc->tc_image = tc_image;
  desc->tc_terms = tc_image ? tc_terms : 0;
  return rc;
}

// ---- TakeSerialSteps.sample + MALA / HMC / GaussianRandomWalk kernel (take_steps.py:127-142,156-180) ------------
ffi::Error LocalStepsImpl(cudaStream_t stream, ffi::ScratchAllocator alloc, F32 x0, F32 target_data, F32 hmc_chol,
                          F32 hmc_colsum, F32 pos_in, F32 lp_in, F32 acc_in, RF32 pos, RF32 lp, RF32 acc, RF32 last,
                          int32_t kind, int32_t target_id, int64_t cursor, int32_t n_steps, int32_t thinning,
                          float step_size, int32_t n_leapfrog, int32_t hmc_chol_diagonal, int64_t chain_offset,
                          int64_t n_chains_global, uint32_t key0, uint32_t key1) {
  (void)pos_in; (void)lp_in; (void)acc_in;  // aliased to pos / lp / acc (input_output_aliases)
  const auto dims = x0.dimensions();
  const int64_t n_chains = dims[0];
  const int d = (int)dims[1];
  FlowmcLocalParams p;
  std::memset(&p, 0, sizeof(p));
  p.step_size = step_size;
  p.n_leapfrog = n_leapfrog;
  p.hmc_chol = opt(hmc_chol);
  p.hmc_colsum = opt(hmc_colsum);
  p.hmc_chol_diagonal = hmc_chol_diagonal;
  p.workspace_bytes = flowmc_local_steps_workspace_bytes(n_chains, d, 0);
  if (auto e = scratch(alloc, p.workspace_bytes, &p.workspace); e.failure()) return e;
  const uint32_t key[2] = {key0, key1};
  uint32_t key_out[2];
  return done(flowmc_local_steps(kind, target_id, target_data.typed_data(), key, x0.typed_data(), pos->typed_data(),
                                 lp->typed_data(), acc->typed_data(), pos->dimensions()[1], cursor, n_chains, d,
                                 n_steps, thinning, chain_offset, n_chains_global, &p, key_out, last->typed_data(),
                                 stream));
}

// ---- ProposalBase.kernel: one application with explicit per-chain keys (resource/kernel/base.py:16-27) -----------
ffi::Error KernelStepImpl(cudaStream_t stream, U32 keys, F32 x, F32 log_prob, F32 target_data, F32 hmc_chol,
                          F32 hmc_colsum, RF32 pos, RF32 lp, RF32 acc, int32_t kind, int32_t target_id,
                          float step_size, int32_t n_leapfrog, int32_t hmc_chol_diagonal) {
  const auto dims = x.dimensions();
  FlowmcLocalParams p;
  std::memset(&p, 0, sizeof(p));
  p.step_size = step_size;
  p.n_leapfrog = n_leapfrog;
  p.hmc_chol = opt(hmc_chol);
  p.hmc_colsum = opt(hmc_colsum);
  p.hmc_chol_diagonal = hmc_chol_diagonal;
  p.step_keys = keys.typed_data();
  p.lp0 = log_prob.typed_data();
  const uint32_t key[2] = {0, 0};
  uint32_t key_out[2];
  // buffers of length 1: positions[:, 0] IS the new position, so `last` aliases pos
  return done(flowmc_local_steps(kind, target_id, target_data.typed_data(), key, x.typed_data(), pos->typed_data(),
                                 lp->typed_data(), acc->typed_data(), 1, 0, dims[0], (int)dims[1], 1, 1, 0, dims[0], &p,
                                 key_out, pos->typed_data(), stream));
}

// ---- LogPDF.__call__ / jax.value_and_grad(logpdf) (resource/logPDF.py:60-61, MALA.py:59) -------------------------
ffi::Error TargetEvalImpl(cudaStream_t stream, F32 x, F32 target_data, RF32 logp, RF32 grad, int32_t target_id) {
  const auto dims = x.dimensions();
  return done(flowmc_target_eval(target_id, target_data.typed_data(), x.typed_data(), dims[0], (int)dims[1],
                                 logp->typed_data(), grad->element_count() ? grad->typed_data() : nullptr, stream));
}

// ---- TakeGroupSteps.sample + NFProposal.kernel (take_steps.py:191-206, NF_proposal.py:27-172) --------------------
ffi::Error GlobalStepsImpl(cudaStream_t stream, ffi::ScratchAllocator alloc, F32 params, ffi::AnyBuffer tc_image,
                           F32 x0, F32 target_data, F32 pos_in, F32 lp_in, F32 acc_in, RF32 pos, RF32 lp, RF32 acc,
                           RF32 last, int32_t n_features, int32_t n_layers, ffi::Span<const int32_t> hidden,
                           int32_t num_bins, float range_min, float range_max, int32_t tc_terms, int32_t target_id,
                           int64_t cursor, int32_t n_steps, int32_t thinning, int32_t n_batch_size,
                           int64_t chain_offset, int64_t n_chains_global, uint32_t key0, uint32_t key1) {
  (void)pos_in; (void)lp_in; (void)acc_in;
  FlowmcFlowDesc desc;
  if (int rc = make_desc(&desc, n_features, n_layers, hidden, num_bins, range_min, range_max,
                         tc_image.element_count() ? tc_image.untyped_data() : nullptr, tc_terms))
    return done(rc);
  const int64_t n_chains = x0.dimensions()[0];
  FlowmcGlobalParams g;
  std::memset(&g, 0, sizeof(g));
  g.n_batch_size = n_batch_size;
  g.workspace_bytes = flowmc_nf_global_steps_workspace_bytes(n_chains, n_features, n_steps);
  if (auto e = scratch(alloc, g.workspace_bytes, &g.workspace); e.failure()) return e;
  const uint32_t key[2] = {key0, key1};
  uint32_t key_out[2];
  return done(flowmc_nf_global_steps(&desc, params.typed_data(), target_id, target_data.typed_data(), key,
                                     x0.typed_data(), pos->typed_data(), lp->typed_data(), acc->typed_data(),
                                     pos->dimensions()[1], cursor, n_chains, n_steps, thinning, chain_offset,
                                     n_chains_global, &g, key_out, last->typed_data(), stream));
}

// ---- MaskedCouplingRQSpline.forward / inverse / log_prob / sample (rqSpline.py:450-504) --------------------------
// mode: 0 forward, 1 inverse (y, logdet); 2 log_prob (logdet only is written, y unused)
ffi::Error FlowApplyImpl(cudaStream_t stream, F32 params, ffi::AnyBuffer tc_image, F32 x, RF32 y, RF32 logdet,
                         int32_t mode, int32_t n_features, int32_t n_layers, ffi::Span<const int32_t> hidden,
                         int32_t num_bins, float range_min, float range_max, int32_t tc_terms) {
  FlowmcFlowDesc desc;
  if (int rc = make_desc(&desc, n_features, n_layers, hidden, num_bins, range_min, range_max,
                         tc_image.element_count() ? tc_image.untyped_data() : nullptr, tc_terms))
    return done(rc);
  const int64_t n = x.dimensions()[0];
  if (mode == 0)
    return done(flowmc_flow_forward(&desc, params.typed_data(), x.typed_data(), n, y->typed_data(),
                                    logdet->typed_data(), stream));
  if (mode == 1)
    return done(flowmc_flow_inverse(&desc, params.typed_data(), x.typed_data(), n, y->typed_data(),
                                    logdet->typed_data(), stream));
  return done(flowmc_flow_log_prob(&desc, params.typed_data(), x.typed_data(), n, logdet->typed_data(), nullptr,
                                   stream));
}

ffi::Error FlowSampleImpl(cudaStream_t stream, F32 params, ffi::AnyBuffer tc_image, RF32 x_out, int32_t n_features,
                          int32_t n_layers, ffi::Span<const int32_t> hidden, int32_t num_bins, float range_min,
                          float range_max, int32_t tc_terms, uint32_t key0, uint32_t key1) {
  FlowmcFlowDesc desc;
  if (int rc = make_desc(&desc, n_features, n_layers, hidden, num_bins, range_min, range_max,
                         tc_image.element_count() ? tc_image.untyped_data() : nullptr, tc_terms))
    return done(rc);
  const int64_t n = x_out->dimensions()[0];
  const uint32_t key[2] = {key0, key1};
  return done(flowmc_flow_sample(&desc, params.typed_data(), nullptr, key, n, n, x_out->typed_data(), stream));
}

// refresh of the tensor-core weight image after a parameter update
ffi::Error FlowTcPackImpl(cudaStream_t stream, F32 params, ffi::Result<ffi::AnyBuffer> image, int32_t n_features,
                          int32_t n_layers, ffi::Span<const int32_t> hidden, int32_t num_bins, float range_min,
                          float range_max, int32_t tc_terms) {
  FlowmcFlowDesc desc;
  if (int rc = make_desc(&desc, n_features, n_layers, hidden, num_bins, range_min, range_max, nullptr, tc_terms))
    return done(rc);
  desc.tc_terms = tc_terms;
  return done(flowmc_flow_tc_pack(&desc, params.typed_data(), image->untyped_data(), stream));
}

// ---- NFModel.loss_fn + gradient, Optimizer update (nf_model/base.py:98-125, optimizer.py:19-23) ------------------
ffi::Error FlowLossGradImpl(cudaStream_t stream, ffi::ScratchAllocator alloc, F32 params, ffi::AnyBuffer tc_image,
                            F32 x, S32 idx, RF32 grad, RF32 loss, int32_t n_features, int32_t n_layers,
                            ffi::Span<const int32_t> hidden, int32_t num_bins, float range_min, float range_max,
                            int32_t tc_terms, float inv_n_total) {
  FlowmcFlowDesc desc;
  if (int rc = make_desc(&desc, n_features, n_layers, hidden, num_bins, range_min, range_max,
                         tc_image.element_count() ? tc_image.untyped_data() : nullptr, tc_terms))
    return done(rc);
  const int64_t n = idx.element_count() ? (int64_t)idx.element_count() : x.dimensions()[0];
  const int64_t ws_bytes = flowmc_flow_loss_grad_workspace_bytes(&desc, n);
  void* ws = nullptr;
  if (auto e = scratch(alloc, ws_bytes, &ws); e.failure()) return e;
  return done(flowmc_flow_loss_grad(&desc, params.typed_data(), x.typed_data(), opt(idx), n, inv_n_total,
                                    grad->typed_data(), loss->typed_data(), ws, ws_bytes, stream));
}

// params / mu / nu are donated and updated in place (outputs alias inputs 0, 2, 3)
ffi::Error ClipAdamWImpl(cudaStream_t stream, ffi::ScratchAllocator alloc, F32 params_in, F32 grads, F32 mu_in,
                         F32 nu_in, RF32 params, RF32 mu, RF32 nu, int64_t count, double lr, double b1, double b2,
                         double eps, double weight_decay, double max_norm) {
  (void)params_in; (void)mu_in; (void)nu_in;
  void* ws = nullptr;
  if (auto e = scratch(alloc, 256 * sizeof(float), &ws); e.failure()) return e;
  return done(flowmc_clip_adamw((int64_t)params->element_count(), params->typed_data(), grads.typed_data(),
                                mu->typed_data(), nu->typed_data(), count, lr, b1, b2, eps, weight_decay, max_norm,
                                static_cast<float*>(ws), nullptr, stream));
}

// ---- TrainModel's data selection (strategy/train_model.py:66-81, nf_model/base.py:141-144,187-188) ---------------
ffi::Error PermutationImpl(cudaStream_t stream, ffi::ScratchAllocator alloc, RS32 out, uint32_t key0, uint32_t key1) {
  const int64_t n = (int64_t)out->element_count();
  const int64_t ws_bytes = flowmc_random_permutation_workspace_bytes(n);
  void* ws = nullptr;
  if (auto e = scratch(alloc, ws_bytes, &ws); e.failure()) return e;
  const uint32_t key[2] = {key0, key1};
  return done(flowmc_random_permutation(key, n, out->typed_data(), ws, ws_bytes, stream));
}

ffi::Error ChoiceImpl(cudaStream_t stream, RS32 out, int64_t n_population, uint32_t key0, uint32_t key1) {
  const uint32_t key[2] = {key0, key1};
  return done(flowmc_random_choice(key, n_population, (int64_t)out->element_count(), out->typed_data(), stream));
}

ffi::Error FiniteRowsImpl(cudaStream_t stream, F32 buf, RS32 rowmap, RS32 counts, RS32 minmax) {
  const auto dims = buf.dimensions();
  return done(flowmc_buffer_finite_rows(buf.typed_data(), dims[0], dims[1], (int)dims[2], rowmap->typed_data(),
                                        counts->typed_data(), minmax->typed_data(), stream));
}

ffi::Error GatherRowsImpl(cudaStream_t stream, F32 buf, S32 rowmap, S32 idx, RF32 out, int32_t window,
                          int32_t m_finite, int64_t chain_lo, int64_t chain_hi) {
  const auto dims = buf.dimensions();
  return done(flowmc_gather_training_rows(buf.typed_data(), rowmap.typed_data(), dims[1], (int)dims[2], window,
                                          m_finite, chain_lo, chain_hi, idx.typed_data(),
                                          (int64_t)idx.element_count(), out->typed_data(), stream));
}

ffi::Error MeanCovImpl(cudaStream_t stream, ffi::ScratchAllocator alloc, F32 x, RF32 mean, RF32 cov) {
  const auto dims = x.dimensions();
  void* ws = nullptr;
  if (auto e = scratch(alloc, (int64_t)(dims[1] < 256 ? 256 : dims[1]) * sizeof(float), &ws); e.failure()) return e;
  return done(flowmc_data_mean_cov(x.typed_data(), dims[0], (int)dims[1], mean->typed_data(), cov->typed_data(),
                                   static_cast<float*>(ws), stream));
}

// ---- ParallelTempering exchange sweep, AdamOptimization ------------------------------------
ffi::Error PtExchangeImpl(cudaStream_t stream, F32 pos_in, F32 lp_in, F32 temperatures, RF32 positions,
                          RF32 log_probs, RF32 accepts, int64_t chain_offset, int64_t n_chains_global, uint32_t key0,
                          uint32_t key1) {
  (void)pos_in; (void)lp_in;  // aliased to positions / log_probs
  const auto dims = positions->dimensions();
  const uint32_t key[2] = {key0, key1};
  return done(flowmc_pt_exchange(key, chain_offset, n_chains_global, dims[0], (int)dims[1], (int)dims[2],
                                 positions->typed_data(), log_probs->typed_data(), temperatures.typed_data(),
                                 accepts->typed_data(), stream));
}

ffi::Error AdamOptimizeImpl(cudaStream_t stream, F32 x0, F32 target_data, F32 bounds_lo, F32 bounds_hi,
                            F32 bias_corrections, RF32 x_out, RF32 logp_out, int32_t target_id, int32_t n_steps,
                            float learning_rate, float noise_level, int64_t chain_offset, int64_t n_chains_global,
                            uint32_t key0, uint32_t key1) {
  const auto dims = x0.dimensions();
  const uint32_t key[2] = {key0, key1};
  uint32_t key_out[2];
  return done(flowmc_adam_optimize(target_id, target_data.typed_data(), key, x0.typed_data(), dims[0], (int)dims[1],
                                   n_steps, learning_rate, noise_level, bounds_lo.typed_data(), bounds_hi.typed_data(),
                                   bias_corrections.typed_data(), chain_offset, n_chains_global, key_out,
                                   x_out->typed_data(), logp_out->typed_data(), stream));
}

}  // namespace

#define FLOWMC_FLOW_ATTRS()                                                                                     \
  Attr<int32_t>("n_features").Attr<int32_t>("n_layers").Attr<ffi::Span<const int32_t>>("hidden")                \
      .Attr<int32_t>("num_bins").Attr<float>("range_min").Attr<float>("range_max").Attr<int32_t>("tc_terms")
#define FLOWMC_KEY_ATTRS() Attr<uint32_t>("key0").Attr<uint32_t>("key1")

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    FlowmcLocalSteps, LocalStepsImpl,
    ffi::Ffi::Bind().Ctx<Stream>().Ctx<ffi::ScratchAllocator>()
        .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
        .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>()
        .Attr<int32_t>("kind").Attr<int32_t>("target_id").Attr<int64_t>("cursor").Attr<int32_t>("n_steps")
        .Attr<int32_t>("thinning").Attr<float>("step_size").Attr<int32_t>("n_leapfrog")
        .Attr<int32_t>("hmc_chol_diagonal").Attr<int64_t>("chain_offset").Attr<int64_t>("n_chains_global")
        .FLOWMC_KEY_ATTRS());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    FlowmcKernelStep, KernelStepImpl,
    ffi::Ffi::Bind().Ctx<Stream>()
        .Arg<U32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
        .Ret<F32>().Ret<F32>().Ret<F32>()
        .Attr<int32_t>("kind").Attr<int32_t>("target_id").Attr<float>("step_size").Attr<int32_t>("n_leapfrog")
        .Attr<int32_t>("hmc_chol_diagonal"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    FlowmcTargetEval, TargetEvalImpl,
    ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>().Attr<int32_t>("target_id"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    FlowmcGlobalSteps, GlobalStepsImpl,
    ffi::Ffi::Bind().Ctx<Stream>().Ctx<ffi::ScratchAllocator>()
        .Arg<F32>().Arg<ffi::AnyBuffer>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
        .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>()
        .FLOWMC_FLOW_ATTRS()
        .Attr<int32_t>("target_id").Attr<int64_t>("cursor").Attr<int32_t>("n_steps").Attr<int32_t>("thinning")
        .Attr<int32_t>("n_batch_size").Attr<int64_t>("chain_offset").Attr<int64_t>("n_chains_global")
        .FLOWMC_KEY_ATTRS());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    FlowmcFlowApply, FlowApplyImpl,
    ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<ffi::AnyBuffer>().Arg<F32>().Ret<F32>().Ret<F32>()
        .Attr<int32_t>("mode").FLOWMC_FLOW_ATTRS());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    FlowmcFlowSample, FlowSampleImpl,
    ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<ffi::AnyBuffer>().Ret<F32>().FLOWMC_FLOW_ATTRS().FLOWMC_KEY_ATTRS());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    FlowmcFlowTcPack, FlowTcPackImpl,
    ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Ret<ffi::AnyBuffer>().FLOWMC_FLOW_ATTRS());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    FlowmcFlowLossGrad, FlowLossGradImpl,
    ffi::Ffi::Bind().Ctx<Stream>().Ctx<ffi::ScratchAllocator>()
        .Arg<F32>().Arg<ffi::AnyBuffer>().Arg<F32>().Arg<S32>().Ret<F32>().Ret<F32>()
        .FLOWMC_FLOW_ATTRS().Attr<float>("inv_n_total"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    FlowmcClipAdamW, ClipAdamWImpl,
    ffi::Ffi::Bind().Ctx<Stream>().Ctx<ffi::ScratchAllocator>()
        .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>().Ret<F32>()
        .Attr<int64_t>("count").Attr<double>("lr").Attr<double>("b1").Attr<double>("b2").Attr<double>("eps")
        .Attr<double>("weight_decay").Attr<double>("max_norm"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    FlowmcRandomPermutation, PermutationImpl,
    ffi::Ffi::Bind().Ctx<Stream>().Ctx<ffi::ScratchAllocator>().Ret<S32>().FLOWMC_KEY_ATTRS());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    FlowmcRandomChoice, ChoiceImpl,
    ffi::Ffi::Bind().Ctx<Stream>().Ret<S32>().Attr<int64_t>("n_population").FLOWMC_KEY_ATTRS());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    FlowmcBufferFiniteRows, FiniteRowsImpl,
    ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Ret<S32>().Ret<S32>().Ret<S32>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    FlowmcGatherTrainingRows, GatherRowsImpl,
    ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<S32>().Arg<S32>().Ret<F32>()
        .Attr<int32_t>("window").Attr<int32_t>("m_finite").Attr<int64_t>("chain_lo").Attr<int64_t>("chain_hi"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    FlowmcDataMeanCov, MeanCovImpl,
    ffi::Ffi::Bind().Ctx<Stream>().Ctx<ffi::ScratchAllocator>().Arg<F32>().Ret<F32>().Ret<F32>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    FlowmcPtExchange, PtExchangeImpl,
    ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>().Ret<F32>()
        .Attr<int64_t>("chain_offset").Attr<int64_t>("n_chains_global").FLOWMC_KEY_ATTRS());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    FlowmcAdamOptimize, AdamOptimizeImpl,
    ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>()
        .Attr<int32_t>("target_id").Attr<int32_t>("n_steps").Attr<float>("learning_rate").Attr<float>("noise_level")
        .Attr<int64_t>("chain_offset").Attr<int64_t>("n_chains_global").FLOWMC_KEY_ATTRS());

#endif  // FLOWMC_HAVE_XLA_FFI
