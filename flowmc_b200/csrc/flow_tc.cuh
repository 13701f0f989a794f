// Shared definitions of the tensor-core flow kernels (flow_tc.cu: forward / inverse / proposals,
// flow_train_tc.cu: backward): CTA shape, warp roles, and the layout of the per-tile activation images the
// forward pass leaves behind for the weight-gradient GEMMs.
#pragma once
#include "flow_common.cuh"
#include "tc_common.cuh"

namespace flowmc {

constexpr int TC_M = 128;          // samples per CTA (= TMEM lanes)
constexpr int TC_PARTS = 2;        // epilogue threads per sample row (they split columns / features)
constexpr int TC_EPI_WARPS = 4 * TC_PARTS;
constexpr int TC_EPI = TC_EPI_WARPS * 32;
constexpr int TC_THREADS = TC_EPI + 64;        // + weight-producer warp + MMA-issuer warp
constexpr int TC_STAGE_BYTES = 2 * 128 * 128;  // hi + lo images of up to 128 rows x 128 B (32 K-elements)

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI) : "memory"); }

__host__ __device__ inline int tc_pad16(int v) { return (v + 15) & ~15; }

// Activation image of one (tile, layer): item 0 = the conditioner input x*mask (N = pad16(d) rows), item i+1 = hidden
// activation h_i (N = dims[i+1] rows).  Each item is 4 K-stages (the tile's rows 32q..32q+31 = stage q) of
// [hi: N x 128 B][lo: N x 128 B] in the packed SWIZZLE_128B K-major form (tc_common.cuh), i.e. exactly the B operand
// of dW = dY^T X with K = the tile's 128 samples.
__host__ __device__ inline size_t tc_act_item_rows(const FlowmcFlowDesc& D, int item) {
  return item == 0 ? (size_t)tc_pad16(D.n_features) : (size_t)D.dims[item];
}
__host__ __device__ inline size_t tc_act_item_off(const FlowmcFlowDesc& D, int item) {
  size_t o = 0;
  for (int i = 0; i < item; ++i) o += 4 * 2 * tc_act_item_rows(D, i) * 128;
  return o;
}
__host__ __device__ inline size_t tc_act_layer_bytes(const FlowmcFlowDesc& D) { return tc_act_item_off(D, D.n_linear); }

bool tc_supported(const FlowmcFlowDesc& D);

}  // namespace flowmc
