// flowmc_target.cuh -- plugin header: how a target log-density reaches the B200 kernels.
//
// The reference hands an arbitrary Python callable to jax.value_and_grad
// (src/flowMC/resource/logPDF.py:60-61, resource/kernel/MALA.py:59, HMC.py:76-79).  Here a
// target is a struct of __device__ functions carrying its *analytic* gradient; the sampling
// kernels (flowmc_b200/csrc/local_steps.cuh, nf kernels) are templates instantiated per target,
// and FLOWMC_REGISTER_TARGET puts the instantiations into the library's registry under a name
// that Python's LogPDF refers to.  A plugin is a .cu file that includes this header, defines a
// struct and registers it; compile it into libflowmc_b200.so or into its own shared object
// loaded after the library (flowmc_b200.targets.compile_target()).
//
// Execution model.  A chain's position x[0..d) is spread over a group of G lanes of one warp
// (each lane owns a few dimensions, held in registers) and mirrored in a shared-memory row that
// the target may read freely.  One evaluation is three calls:
//
//   k     = T::prepare(data, d)            once per kernel: a small struct of constants (registers)
//   aux_j = T::partial(k, ctx, j, x_j, red)   for every owned dimension j: add this dimension's
//                                          contribution(s) to red[0..NRED); return any per-
//                                          dimension value grad() will want (e.g. (P x)_j)
//   -- the kernel sums red[] over the chain's lanes --
//   logp  = T::finish(k, ctx, red)            every lane of the group: turn the reduced scalars into
//                                          log p(x); may overwrite red[] with whatever grad() needs
//   g_j   = T::grad(k, ctx, j, x_j, aux_j, red)  for every owned dimension j: d log p / d x_j
//
// `ctx.x` is the full vector (read-only; x[-1] and x[d] are readable and hold 0, so nearest-
// neighbour couplings need no bounds tests), `ctx.data` the packed float32 parameter block built on
// the host (the reference's `data` pytree; layout is target-defined), `ctx.scratch` a private
// row of >= d floats in shared memory (valid only if USES_SCRATCH).  partial() of all owned
// dimensions completes (with a warp barrier) before finish()/grad() run.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace flowmc {

struct TargetCtx {
  const float* x;      // shared memory: x[0..d)
  float* scratch;      // shared memory: >= d floats, private to this chain (if USES_SCRATCH)
  const float* data;   // global memory: packed target parameters
  int d;
};


// a / b for a CONSTANT b with r = RN(1 / b) given as a literal: product + one residual correction = the IEEE
// round-to-nearest quotient for every normal a (checked exhaustively over all float32 significands for b = 20), in 3
// FMA-pipe instructions instead of the ~10 + slow-path branch of a true division.  Non-finite quotients pass through
// as a * r would give them (inf / nan like a / b).
__device__ __forceinline__ float div_const(float a, float b, float r) {
  const float q = a * r;
  const float c = fmaf(fmaf(-q, b, a), r, q);
  return (fabsf(q) <= 3.402823466e+38f) ? c : q;
}
}  // namespace flowmc

#include "../flowmc_b200/csrc/local_steps.cuh"
#include "../flowmc_b200/csrc/adam_opt.cuh"
#include "../flowmc_b200/csrc/registry.h"

// Registers target struct T under `NAME` at load time (static initialiser).
#define FLOWMC_REGISTER_TARGET(T, NAME)                                                         \
  namespace {                                                                                   \
  struct FlowmcRegistrar_##T {                                                                  \
    FlowmcRegistrar_##T() {                                                                     \
      static FlowmcTargetVTable vt;                                                             \
      vt.abi_version = FLOWMC_TARGET_ABI;                                                       \
      vt.name = NAME;                                                                           \
      vt.local_steps = &flowmc::launch_local_steps<T>;                                          \
      vt.eval = &flowmc::launch_target_eval<T>;                                                 \
      vt.adam_opt = &flowmc::launch_adam_opt<T>;                                                \
      flowmc_register_target(&vt);                                                              \
    }                                                                                           \
  };                                                                                            \
  static FlowmcRegistrar_##T flowmc_registrar_instance_##T;                                     \
  }
