// Backward pass of -mean(log_prob) on the sm_100a tensor cores.
//
// Reference: NFModel.loss_fn / train_step (src/flowMC/resource/model/nf_model/base.py:98-125): reverse-mode
// autodiff of MaskedCouplingRQSpline.log_prob (rqSpline.py:392-504).  Same hand-derived adjoints as the CUDA-core
// kernel in flow_train.cu (rq_backward), but every matrix product is a tcgen05.mma kind::tf32 (3xTF32) with M = 128:
//
//   data gradients   dh = dY W            A = dY  (TMEM, lane = sample: written by the epilogue threads)
//                                          B = W^T (pre-packed transposed weight image, streamed by cp.async.bulk)
//   weight gradients dW = dY^T X          A = dY^T (TMEM, lane = output unit: the epilogue transposes dY through
//                                               shared memory), K = the tile's 128 samples
//                                          B = X as left behind by the tensor-core FORWARD pass: its epilogue writes
//                                              the conditioner input and hidden activations straight into the packed
//                                              B-stage layout (flow_tc.cuh: tc_act_*), so no recompute and no
//                                              re-layout pass
//   the dW tile (lane = output unit) is read back with tcgen05.ld and stored into this CTA's private, permuted
//   accumulator; the accumulators are summed in CTA order by reducer CTAs inside the same launch (no atomics).
//
// One CTA = 128 samples, layers in reverse.  Per layer: for every chunk of 4 transformed features
// {spline adjoints -> dtheta; dh_last += dtheta W3_c (acc slot 0); dW3_c = dtheta^T h_last (acc slot 1)}, then per
// tanh layer {da = dh (1 - h^2); dh_prev = da W; dW = da^T h_prev}, masked-coupling and ScalarAffine adjoints.
// The epilogue is software-pipelined against the MMA warp (see the kernel: adjoints of unit u+1 under the weight-
// gradient MMAs of unit u, dW store of unit u-1 under the data-gradient MMAs of unit u); the weight producer
// prefetches the next items' stages through a 3-deep ring meanwhile.
// TMEM map: [0,128) A hi | [128,256) A lo | [256,384) acc 0 (data gradient) | [384,512) acc 1 (weight gradient).
#include <cstdlib>
#include <cstring>
#include <string>

#include "flow_tc.cuh"
#include "flow_tile.cuh"
#include "flow_train.cuh"
#include "registry.h"

namespace flowmc {

constexpr int BT_STAGES = 3;
constexpr int BT_TS = 132;  // row stride (floats) of the transpose buffer T[k][row]: conflict-free 128-bit reads
constexpr int BT_MAX_ITEMS = 80;
constexpr int BT_MAX_CTAS = 148;  // persistent CTAs, each with a private gradient accumulator

enum : int { BK_DG3 = 0, BK_WG3 = 1, BK_DGH = 2, BK_WGH = 3 };

struct BtItem {
  int kind;
  int N;        // MMA N (multiple of 16)
  int n_kc;     // stages of 32 K-elements
  int K;        // reduction length (columns of A that are valid)
  int lin;      // hidden Linear index, or first transformed-feature ordinal of the chunk
  int n_feat;   // chunk: features
  int act;      // B from the activation image (1) or the transposed-weight image (0)
  uint32_t off; // byte offset inside the layer's weight image / the (tile, layer) activation image
  uint32_t aoff; // BK_DGH: byte offset of the unit's own activations h_lin inside the (tile, layer) activation image
};
struct BtProgram {
  int n_items[2];
  uint32_t layer_bytes[2];
  int fc;
  uint32_t act_layer_bytes;  // tc_act_layer_bytes(D)
  BtItem items[2][BT_MAX_ITEMS];
};

static int bt_build_program(const FlowmcFlowDesc& D, BtProgram* P) {
  const int d = D.n_features, NP = 3 * D.num_bins + 1, nh = D.n_linear - 1;
  int fc = (128 / NP) & ~1;
  if (fc > 4) fc = 4;  // one 32-column slot of the A region per feature
  P->fc = fc;
  P->act_layer_bytes = (uint32_t)tc_act_layer_bytes(D);
  const int H = D.dims[nh];
  for (int p = 0; p < 2; ++p) {
    int n = 0;
    uint32_t off = 0;
    const int ntf = (d - p + 1) / 2;
    for (int c0 = 0; c0 < ntf; c0 += fc) {
      if (n + 2 > BT_MAX_ITEMS) return FLOWMC_ERR_UNSUPPORTED;
      const int nf = (ntf - c0 < fc) ? ntf - c0 : fc;
      BtItem& dg = P->items[p][n++];
      dg.kind = BK_DG3; dg.N = H; dg.n_kc = nf; dg.K = nf * 32; dg.lin = c0; dg.n_feat = nf; dg.act = 0; dg.off = off;
      dg.aoff = 0;
      off += (uint32_t)dg.n_kc * 2u * dg.N * 128u;
      BtItem& wg = P->items[p][n++];
      wg.kind = BK_WG3; wg.N = H; wg.n_kc = 4; wg.K = 128; wg.lin = c0; wg.n_feat = nf; wg.act = 1;
      wg.off = (uint32_t)tc_act_item_off(D, nh);  // h_last
      wg.aoff = 0;
    }
    for (int i = nh - 1; i >= 0; --i) {
      if (n + 2 > BT_MAX_ITEMS) return FLOWMC_ERR_UNSUPPORTED;
      const int Nin = (i == 0) ? tc_pad16(d) : D.dims[i];
      BtItem& dg = P->items[p][n++];
      dg.kind = BK_DGH; dg.N = Nin; dg.K = D.dims[i + 1]; dg.n_kc = (dg.K + 31) / 32; dg.lin = i; dg.n_feat = 0;
      dg.act = 0; dg.off = off; dg.aoff = (uint32_t)tc_act_item_off(D, i + 1);
      off += (uint32_t)dg.n_kc * 2u * dg.N * 128u;
      BtItem& wg = P->items[p][n++];
      wg.kind = BK_WGH; wg.N = Nin; wg.n_kc = 4; wg.K = 128; wg.lin = i; wg.n_feat = 0; wg.act = 1;
      wg.off = (uint32_t)tc_act_item_off(D, i);  // i == 0: x * mask, else h_{i-1}
      wg.aoff = 0;
    }
    P->n_items[p] = n;
    P->layer_bytes[p] = off;
  }
  return FLOWMC_OK;
}

__host__ __device__ inline int64_t bt_layer_base(const BtProgram& P, int l) {
  return (int64_t)((l + 1) / 2) * P.layer_bytes[0] + (int64_t)(l / 2) * P.layer_bytes[1];
}

// transposed weight images for the data-gradient GEMMs.  grid = (items, layers, element slices)
__global__ void bt_pack_kernel(const FlowmcFlowDesc D, const BtProgram P, const float* __restrict__ params,
                               uint8_t* __restrict__ image, int* __restrict__ done) {
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
    for (int i = threadIdx.x; i <= D.n_layers; i += blockDim.x) done[i] = 0;  // flags + work counter
  const int l = blockIdx.y, p = l & 1;
  if ((int)blockIdx.x >= P.n_items[p]) return;
  const BtItem it = P.items[p][blockIdx.x];
  if (it.act) return;
  const int NP = 3 * D.num_bins + 1, nh = D.n_linear - 1;
  const float* PL = params + (int64_t)l * D.layer_stride;
  float* dst = reinterpret_cast<float*>(image + bt_layer_base(P, l) + it.off);
  const int per_stage = it.N * 32;
  const int total = it.n_kc * per_stage;
  for (int i = blockIdx.z * blockDim.x + threadIdx.x; i < total; i += blockDim.x * gridDim.z) {
    const int kc = i / per_stage, rem = i - kc * per_stage;
    const int n = rem >> 5, kk = rem & 31;
    float w = 0.0f;
    if (it.kind == BK_DG3) {
      // B[n = hidden unit][k = (feature slot kc, parameter kk)] = W3[(f NP + kk)][n]
      const int f = p + 2 * (it.lin + kc);
      if (kk < NP && n < D.dims[nh]) w = PL[D.off_W[nh] + ((int64_t)f * NP + kk) * D.dims[nh] + n];
    } else {
      // B[n = input unit][k = output unit] = W_i[k][n]
      const int k = kc * 32 + kk;
      if (k < D.dims[it.lin + 1] && n < D.dims[it.lin]) w = PL[D.off_W[it.lin] + (int64_t)k * D.dims[it.lin] + n];
    }
    uint32_t hi, lo;
    tc::split_tf32_rn(w, hi, lo);
    float* stage = dst + (int64_t)kc * 2 * per_stage;
    const int o = tc::packed_b_offset(n, kk) >> 2;
    stage[o] = __uint_as_float(hi);
    stage[per_stage + o] = __uint_as_float(lo);
  }
}

struct BtArgs {
  const float* params;
  const uint8_t* wimg;     // transposed weight image (bt_pack_kernel)
  const uint8_t* act_img;  // activation images written by the forward pass
  const float* save_x;     // [L + 1][n][d]
  const float* save_theta; // [L][ceil(d/2) NP][n]
  const float* logp;       // [n]
  int64_t n;
  float inv_n;
  float* grad;         // final gradient (written by bt_reduce_kernel)
  float* loss;
  float* partial;      // [gridDim.x][pstride]: this CTA's private gradient accumulator (no atomics, fixed order)
  int64_t pstride;     // floats per CTA: L * layer_stride + 4 (the last 4: loss partial)
  int64_t n_tiles;
  long long* timing;  // optional diagnostics: clock64 stamps of epilogue thread 0 of CTA 0
  int split_r;        // feature split: CTAs per tile (cluster size), 1 = one CTA per tile
  int dbg_skip;       // timing experiments only (FLOWMC_BT_DBG_SKIP): 1 = no dW stores, 2 = no dW tile reads either
  int n_cta;          // CTAs that process tiles; CTAs beyond them (if any) are REDUCERS, see bt_reduce_layer
  int* done;          // [L] tile CTAs that have finished layer l (reducer hand-off) + [1] reduction work counter;
                      // zeroed by bt_pack_kernel
};

#define BT_STAMP()                                                                                \
  do {                                                                                            \
    if (a.timing != nullptr && blockIdx.x == 0 && tid == 0 && n_stamp < 252) a.timing[n_stamp++] = clock64(); \
  } while (0)

__device__ __forceinline__ void bt_reduce_layer(const FlowmcFlowDesc& D, const float* partial, int64_t pstride,
                                                int n_cta, float* __restrict__ grad, int l, int g0, int gstep,
                                                int g_end, int R = 1, int fc = 1);
__device__ __forceinline__ float bt_reduce_loss(const float* partial, int64_t pstride, int n_cta);

// PARTS = epilogue threads per sample row: 2 (default) or 4 (16 epilogue warps, 576 threads, 96 registers per thread:
// an experiment that measured slower, see flow_backward_tc).
template <int PARTS>
struct BtSmem {
  uint64_t stage_full[BT_STAGES], stage_empty[BT_STAGES], acc_full;
  uint64_t a_ready[4];  // per K-chunk (32 columns) of the A operand, one arrival per epilogue warp
  uint64_t xbar[2];  // feature split: the two exchange rounds of finish_layer (one arrival per epilogue warp of
                     // every CTA of the cluster)
  uint32_t tmem_base;
  float red[2 * 4 * PARTS];
  float bsum[PARTS][TC_M];
};

template <int N>
__device__ __forceinline__ void bt_bar() { asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory"); }

// SPLIT: a cluster of R = a.split_r CTAs shares ONE tile, like the split training forward (flow_tc.cu).  Every CTA takes
// the spline chunks with index % R == its cluster rank; the data gradient it accumulates (dh_last) is therefore only a
// PARTIAL sum -- and stays one: the tanh units are linear in it, so every CTA pushes its partial through them and
// stores partial weight gradients (the accumulator reduction sums them like it sums tiles).  What the CTAs must
// exchange per layer is small: the gradient w.r.t. the layer input -- for the transformed features the owner's value,
// for the conditioning features the sum of the CTAs' partials -- pulled through distributed shared memory in
// finish_layer (two cluster rounds, partials added in rank order: deterministic).
template <int KB, int PARTS, bool SPLIT = false>
__global__ void __launch_bounds__(4 * PARTS * 32 + 64, 1) flow_backward_tc_kernel(const FlowmcFlowDesc D,
                                                                                 const BtProgram PR, const BtArgs a) {
  constexpr int NP = 3 * KB + 1;
  const uint32_t R = SPLIT ? (uint32_t)a.split_r : 1u;
  const uint32_t crank = SPLIT ? tc::cluster_rank() : 0u;
  const int64_t tile_first = SPLIT ? (int64_t)(blockIdx.x / R) : (int64_t)blockIdx.x;
  const int64_t tile_step = SPLIT ? (int64_t)(a.n_cta / (int)R) : (int64_t)a.n_cta;
  // items 2c, 2c + 1 (c < number of chunks) are the data / weight gradient GEMMs of spline chunk c
  auto skip_item = [&](const BtItem& it, int ii) -> bool {
    return SPLIT && (it.kind == BK_DG3 || it.kind == BK_WG3) && (uint32_t)(ii >> 1) % R != crank;
  };
  constexpr int EPI_WARPS = 4 * PARTS, EPI = EPI_WARPS * 32;
  constexpr int CPT = 128 / PARTS;  // accumulator / operand columns per epilogue thread of a row
  auto epi_bar = [] { bt_bar<EPI>(); };
  using BtSmem = flowmc::BtSmem<PARTS>;
  if ((int)blockIdx.x < a.n_cta) {  // ===== tile CTA (CTAs beyond n_cta only reduce, below) =====
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stages = smem;
  BtSmem* S = reinterpret_cast<BtSmem*>(smem + BT_STAGES * TC_STAGE_BYTES);
  float* T = reinterpret_cast<float*>(smem + BT_STAGES * TC_STAGE_BYTES + ((sizeof(BtSmem) + 15) & ~15));  // [128][BT_TS]
  const int d = D.n_features;
  const int gs = d + 1;
  float* g = T + 128 * BT_TS;  // [128][d + 1] gradient w.r.t. the current layer output / input
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* P = a.params;
  const int L = D.n_layers, nh = D.n_linear - 1;
  const int64_t n = a.n;

  if (warp == EPI_WARPS + 1 && lane == 0) {
    for (int i = 0; i < BT_STAGES; ++i) {
      tc::mbar_init(&S->stage_full[i], 1);
      tc::mbar_init(&S->stage_empty[i], 1);
    }
    tc::mbar_init(&S->acc_full, 1);
    for (int i = 0; i < 4; ++i) tc::mbar_init(&S->a_ready[i], EPI_WARPS);
    for (int i = 0; i < 2; ++i) tc::mbar_init(&S->xbar[i], R * EPI_WARPS);
    tc::fence_mbar_init();
  }
  if (a.timing != nullptr && blockIdx.x == 0 && tid == 0) a.timing[252] = clock64();  // CTA 0 entered the kernel
  if (warp == EPI_WARPS) tc::tmem_alloc<512>(&S->tmem_base);
  tc::tc_fence_before();
  __syncthreads();
  if (SPLIT) tc::cluster_sync();  // every CTA's barriers exist before a peer arrives on them
  tc::tc_fence_after();
  const uint32_t tbase = S->tmem_base;
  const uint32_t t_ahi = tbase, t_alo = tbase + 128;

  if (warp == EPI_WARPS) {
    // ===== B-stage producer ====================================================================
    uint32_t s = 0, ph = 0;
    for (int64_t tile = tile_first; tile < a.n_tiles; tile += tile_step)
    for (int l = L - 1; l >= 0; --l) {
      const int p = l & 1;
      const uint8_t* wbase = a.wimg + bt_layer_base(PR, l);
      const uint8_t* abase = a.act_img + (tile * L + l) * (size_t)PR.act_layer_bytes;
      for (int ii = 0; ii < PR.n_items[p]; ++ii) {
        const BtItem it = PR.items[p][ii];
        if (skip_item(it, ii)) continue;
        const uint32_t bytes = 2u * it.N * 128u;
        const uint8_t* src = (it.act ? abase : wbase) + it.off;
        for (int kc = 0; kc < it.n_kc; ++kc) {
          tc::mbar_wait(&S->stage_empty[s], ph ^ 1);
          if (tc::elect_one()) {
            tc::mbar_arrive_expect_tx(&S->stage_full[s], bytes);
            tc::bulk_g2s(stages + (size_t)s * TC_STAGE_BYTES, src + (size_t)kc * bytes, bytes, &S->stage_full[s]);
          }
          __syncwarp();
          if (++s == BT_STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // ===== MMA issuer ==========================================================================
    uint32_t s = 0, ph = 0, a_ph = 0;
    for (int64_t tile = tile_first; tile < a.n_tiles; tile += tile_step)
    for (int l = L - 1; l >= 0; --l) {
      const int p = l & 1;
      bool first_chunk = true;  // the first spline chunk THIS CTA runs in the layer starts dh_last afresh
      for (int ii = 0; ii < PR.n_items[p]; ++ii) {
        const BtItem it = PR.items[p][ii];
        if (skip_item(it, ii)) continue;
        // (the epilogue signals the item's A operand K-chunk by K-chunk, see below)
        const bool wg = (it.kind == BK_WG3) || (it.kind == BK_WGH);
        const uint32_t t_acc = tbase + (wg ? 384 : 256);
        const uint32_t idesc = tc::make_idesc_tf32(TC_M, it.N);
        const bool cont = (it.kind == BK_DG3) && !first_chunk;  // later chunks accumulate into dh_last
        if (it.kind == BK_DG3) first_chunk = false;
        for (int kc = 0; kc < it.n_kc; ++kc) {
          tc::mbar_wait(&S->a_ready[kc], (a_ph >> kc) & 1);  // K-chunk kc of the A operand is written
          a_ph ^= 1u << kc;
          tc::mbar_wait(&S->stage_full[s], ph);
          tc::tc_fence_after();
          const uint32_t b_hi = tc::smem_u32(stages + (size_t)s * TC_STAGE_BYTES);
          const uint64_t dhi = tc::make_b_desc(b_hi), dlo = tc::make_b_desc(b_hi + it.N * 128);
          const int ksteps = min(4, (it.K - kc * 32 + 7) >> 3);
          const uint32_t acol = kc * 32;
          if (tc::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (ks < ksteps) {
                tc::mma_tf32_ts(t_acc, t_ahi + acol + ks * 8, dhi + 2 * ks, idesc, (uint32_t)(cont || (kc | ks) != 0));
                tc::mma_tf32_ts(t_acc, t_alo + acol + ks * 8, dhi + 2 * ks, idesc, 1);
                tc::mma_tf32_ts(t_acc, t_ahi + acol + ks * 8, dlo + 2 * ks, idesc, 1);
              }
            }
            tc::mma_commit(&S->stage_empty[s]);
          }
          __syncwarp();
          if (++s == BT_STAGES) { s = 0; ph ^= 1; }
        }
        if (tc::elect_one()) tc::mma_commit(&S->acc_full);
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue warps ======================================================================
    const int q = warp & 3, hf = warp >> 2;
    const int t = q * 32 + lane;                 // sample row (row work) or output unit (transposed work)
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float* gr = g + t * gs;
    uint32_t f_ph = 0;
    int n_stamp = 0;
    auto part = [&](int cnt, int& lo, int& hi) {
      const int per = (cnt + PARTS - 1) / PARTS;
      lo = min(cnt, hf * per);
      hi = min(cnt, lo + per);
    };
    int j_lo, j_hi;
    part(d, j_lo, j_hi);
    float* PB = a.partial + (int64_t)blockIdx.x * a.pstride;  // private accumulator of this CTA

    // ---- software pipeline -------------------------------------------------------------------------
    // A "unit" is one (data-gradient, weight-gradient) pair of GEMMs: a chunk of spline features, or one tanh
    // layer.  For unit u the epilogue threads
    //   [F] compute dY of u (spline adjoints / tanh derivative; at a layer boundary first the masked-coupling and
    //       ScalarAffine adjoints)                                 -- under the wgrad MMAs of u-1
    //   [G] wait for those MMAs (the A region of tensor memory is theirs until then): normally free by now
    //   [A] write dY as the A operand straight from the registers, K-chunk by K-chunk (a_ready[kc]): the dgrad MMAs
    //       of u start on the first features while the later ones are still in [F]
    //   [C] store the dW tile of u-1 (acc 1) into the accumulator  -- under the dgrad MMAs of u
    //   [D] wait for the dgrad MMAs
    //   [E] read dY back TRANSPOSED from shared memory and write it as the A operand -> wgrad MMAs of u start.
    // whole warps call this after every lane has written (and fenced) its part of K-chunk kc: one arrival per warp
    auto arrive_chunk = [&](int kc) {
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&S->a_ready[kc]);
    };
    auto wait_item = [&]() {
      tc::mbar_wait(&S->acc_full, f_ph);
      f_ph ^= 1;
      tc::tc_fence_after();
    };
    // feature split: this thread's column of the exchange buffer T[.][t] in every CTA of the cluster; a control round =
    // release my shared-memory writes at cluster scope, arrive on every CTA's barrier, wait (acquire) for mine
    uint32_t t_remote[8];
    uint32_t x_ph[2] = {0, 0};
    if (SPLIT) {
#pragma unroll
      for (uint32_t k = 0; k < 8; ++k) t_remote[k] = tc::mapa_u32(T + t, k < R ? k : 0u);
    }
    auto cluster_round = [&](int k) {
      tc::fence_cluster();
      __syncwarp();
      if (lane == 0)
        for (uint32_t rk = 0; rk < R; ++rk) tc::mbar_arrive_remote(&S->xbar[k], rk);
      tc::mbar_wait_cluster(&S->xbar[k], x_ph[k]);
      x_ph[k] ^= 1;
    };
    const uint32_t img_sw = (uint32_t)(((lane >> 2) << 4) | ((lane & 3) << 2));  // lane part of an image byte offset
    // 8 consecutive A columns of this thread's row, from registers
    auto write_a8 = [&](int col, const float* v) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) tc::split_tf32_trunc(v[u], hi[u], lo[u]);
      tc::tmem_st8(t_ahi + lane_base + col, hi);
      tc::tmem_st8(t_alo + lane_base + col, lo);
    };
    for (int64_t tile = tile_first; tile < a.n_tiles; tile += tile_step) {
    const bool first = tile == tile_first;           // first tile of this CTA: store, later tiles: accumulate
    const bool last = tile + tile_step >= a.n_tiles; // last tile: publish the layers to the reducers
    auto acc_to = [&](float* ptr, float v) { *ptr = first ? v : *ptr + v; };
    auto store4 = [&](float* ptr, const float* v) {
      float4 w = make_float4(v[0], v[1], v[2], v[3]);
      float4* gp = reinterpret_cast<float4*>(ptr);
      if (!first) {
        const float4 o = *gp;
        w.x += o.x; w.y += o.y; w.z += o.z; w.w += o.w;
      }
      *gp = w;
    };
    const int64_t row0 = tile * TC_M;
    const int64_t grow = row0 + t;
    const bool valid = grow < n;
    const int64_t r = valid ? grow : n - 1;
    const float gld = valid ? -a.inv_n : 0.0f;  // dL/dlogdet of this row
    // loss contribution and dL/dy of the final latent: loss = -mean(logdet + base.log_prob(y))
    if (hf == 0) {
      float v = valid ? a.logp[grow] : 0.0f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) S->red[warp] = -v * a.inv_n;
    }
    for (int j = j_lo; j < j_hi; ++j) {
      const float y = a.save_x[((int64_t)L * n + r) * d + j];
      gr[j] = valid ? a.inv_n * (y - P[D.off_base_mean + j]) / P[D.off_base_cov + (int64_t)j * d + j] : 0.0f;
    }
    epi_bar();
    if (tid == 0)  // (split: every CTA of the cluster sees the same rows; rank 0 accounts for them)
      acc_to(PB + a.pstride - 4, crank == 0 ? (S->red[0] + S->red[1]) + (S->red[2] + S->red[3]) : 0.0f);

    // dY of the unit (layer l, items ii / ii + 1): into the transpose buffer T[column][row] (for the weight-gradient
    // operand) AND, straight from the registers, into the A region (lane = this row) for the data-gradient GEMM.
    // The A region is still being read by the previous unit's weight-gradient MMAs while the adjoints are computed:
    // the wait for them ([G], need_g) sits right before the first A write, where it is normally free.  Spline chunks
    // hand their operand over feature by feature (K-chunk = one feature's 32-column slot; features alternate between
    // the row's two threads), so the MMAs of the first features run under the adjoints of the later ones.  tanh units
    // read the previous data gradient from acc 0, which their own GEMM overwrites: all chunks at the end.
    auto prepare = [&](int l, int ii, bool need_g) {
      const int p = l & 1;
      const BtItem it = PR.items[p][ii];
      const float* PL = P + (int64_t)l * D.layer_stride;
      bool g_pending = need_g;
      if (it.kind == BK_DG3) {
        const float shift = PL[D.off_shift], e = expf(PL[D.off_scale]);
        const float* xin = a.save_x + ((int64_t)l * n + r) * d;
        // chunks owned by the row's other threads: nothing to add
        for (int kc = 0; kc < it.n_feat; ++kc)
          if (kc % PARTS != hf) arrive_chunk(kc);
#pragma unroll
        for (int sg = 0; sg < 4 / PARTS; ++sg) {
          const int fi = PARTS * sg + hf;
          if (fi < it.n_feat) {
            const int fo = it.lin + fi, f = p + 2 * fo;
            float raw[NP], dr[32], gx;
            // feature block [NP][n]: warp-uniform 64-bit base + 32-bit element offsets (n < 2^25 rows, checked by the host)
            const float* th = a.save_theta + ((int64_t)l * ((d + 1) / 2) + fo) * NP * n;
            const uint32_t o0 = (uint32_t)r, nn = (uint32_t)n;
#pragma unroll
            for (int u = 0; u < NP; ++u) raw[u] = th[o0 + (uint32_t)u * nn];
            const float xa = (xin[f] + shift) * e;
            rq_backward<KB, true>(raw, D.range_min, D.range_max, xa, gr[f], gld, gx, dr);
            gr[f] = gx;
#pragma unroll
            for (int u = NP; u < 32; ++u) dr[u] = 0.0f;
#pragma unroll
            for (int u = 0; u < 32; ++u) T[(fi * 32 + u) * BT_TS + t] = dr[u];
            if (g_pending) {
              wait_item();
              g_pending = false;
            }
#pragma unroll
            for (int c = 0; c < 32; c += 8) write_a8(fi * 32 + c, dr + c);
            tc::tmem_wait_st();
            tc::tc_fence_before();
            arrive_chunk(fi);
          }
        }
        if (g_pending) wait_item();
      } else {
        // da = dh (1 - h^2): dh from acc 0, h from the forward pass's activation image (hi + lo)
        const int i = it.lin, N = D.dims[i + 1];
        const uint8_t* abase = a.act_img + (tile * L + l) * (size_t)PR.act_layer_bytes;
        const uint32_t* himg =
            reinterpret_cast<const uint32_t*>(abase + it.aoff + (size_t)q * 2 * N * 128);
        int c_lo, c_hi;
        part(N / 16, c_lo, c_hi);
        const int c0 = c_lo * 16, cn = (c_hi - c_lo) * 16;  // this thread's columns [c0, c0 + cn), cn <= 64
        // 16 columns at a time; the activation words of the next group are requested before this group's are
        // consumed (the image comes from L2 / HBM: ~1 us away)
        uint32_t hb[2][32];
        // word offset of (image row c0 + g4 * 16 + u, this lane) = tc::packed_b_offset(row, lane) / 4: c0 is a multiple
        // of 16, so the row's 8-row group is (c0 >> 3) + 2 g4 + (u >> 3) and its swizzle phase u & 7 -- the lane part
        // (img_sw) is hoisted, one XOR with a constant is left per element
        const uint32_t* hrow = himg + (c0 >> 3) * 256;
        auto request = [&](int g4, uint32_t* dst) {
          const uint32_t* h0 = hrow + g4 * 512;
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const uint32_t o = (uint32_t)((u >> 3) * 256 + (u & 7) * 32) + ((img_sw ^ (uint32_t)((u & 7) << 4)) >> 2);
            dst[u] = __ldg(h0 + o);
            dst[16 + u] = __ldg(h0 + N * 32 + o);
          }
        };
        if (cn > 0) request(0, hb[0]);
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
          if (g4 * 16 < cn) {
            if ((g4 + 1) * 16 < cn) request(g4 + 1, hb[(g4 + 1) & 1]);
            float v[16];
            tc::tmem_ld16(tbase + 256 + lane_base + c0 + g4 * 16, v);
            tc::tmem_wait_ld();
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const float hv = __uint_as_float(hb[g4 & 1][u]) + __uint_as_float(hb[g4 & 1][16 + u]);
              v[u] = v[u] * (1.0f - hv * hv);
              T[(c0 + g4 * 16 + u) * BT_TS + t] = v[u];
            }
            if (g_pending) {
              wait_item();
              g_pending = false;
            }
            write_a8(c0 + g4 * 16, v);
            write_a8(c0 + g4 * 16 + 8, v + 8);
          }
        }
        if (g_pending) wait_item();
        tc::tmem_wait_st();
        tc::tc_fence_before();   // (also orders this thread's acc 0 reads before the GEMM that overwrites acc 0)
        for (int kc = 0; kc < it.n_kc; ++kc) arrive_chunk(kc);
      }
    };
    // the unit after (l, ii): pull the spline parameters its prepare() will read (HBM, written by the forward pass)
    // into L2 one unit ahead
    auto prefetch_next = [&](int l, int ii) {
      int nl = l, nii = ii + 2;
      while (nii < PR.n_items[l & 1] && skip_item(PR.items[l & 1][nii], nii)) nii += 2;
      if (nii >= PR.n_items[l & 1]) {
        // last unit of layer l: finish_layer(l) will read this row's layer input (its share of the d columns)
        const float* xin = a.save_x + ((int64_t)l * n + r) * d;
        for (int j = j_lo; j < j_hi; j += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(xin + j));
        --nl;
        nii = 0;
        while (nl >= 0 && nii < PR.n_items[nl & 1] && skip_item(PR.items[nl & 1][nii], nii)) nii += 2;
      }
      if (nl < 0) return;
      const BtItem it = PR.items[nl & 1][nii];
      if (it.kind != BK_DG3) return;
      int i_lo, i_hi;
      part(it.n_feat, i_lo, i_hi);
      for (int fi = i_lo; fi < i_hi; ++fi) {
        const float* th = a.save_theta + ((int64_t)nl * ((d + 1) / 2) + it.lin + fi) * NP * n + r;
        // one request per 128-byte line (parameter u of the warp's 32 rows): lane u asks for line u
        const float* th0 = th - lane;  // the warp's first row (clamped rows at the end of the batch: still in bounds)
        if (lane < NP && grow - lane < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(th0 + (int64_t)lane * n));
      }
    };
    // dW tile of a finished unit (acc 1; lane = output unit m = t, columns = input units) -> this CTA's private
    // accumulator, straight from tensor memory.  Inside the accumulator a weight block is stored PERMUTED,
    // [column / 4][row][column % 4] (bt_private_to_canonical), so that the 32 lanes of a warp (32 consecutive
    // rows) write 512 contiguous bytes per store; bt_reduce_kernel undoes the permutation while it sums.
    auto reduce_unit = [&](int l, int ii) {
      const int p = l & 1;
      const BtItem it = PR.items[p][ii];
      float* GL = PB + (int64_t)l * D.layer_stride;
      int rows, ncols, ldc;  // rows per block, valid columns, canonical row length
      float* blk;
      bool active;
      int64_t boff;
      if (it.kind == BK_DG3) {
        const int H = D.dims[nh];
        const int f = p + 2 * (it.lin + q);  // warp q of each half owns feature slot q: rows q*32 .. q*32 + NP - 1
        active = q < it.n_feat && lane < NP;
        rows = NP; ncols = H; ldc = H;
        blk = GL + D.off_W[nh] + (int64_t)f * NP * H + lane * 4;
        boff = D.off_b[nh] + (int64_t)f * NP + lane;
      } else {
        const int i = it.lin;
        rows = D.dims[i + 1]; ncols = D.dims[i]; ldc = ncols;
        active = t < rows;
        blk = GL + D.off_W[i] + t * 4;
        boff = D.off_b[i] + t;
      }
      const bool permuted = (ncols & 3) == 0;
      const int npad = tc_pad16(ncols);
#pragma unroll 1
      for (int g2 = 0; g2 < CPT / 32; ++g2) {
        const int c0 = hf * CPT + g2 * 32;
        if (c0 < npad) {
          float v[32];
          if (a.dbg_skip & 2) continue;
          tc::tmem_ld16(tbase + 384 + lane_base + c0, v);
          if (c0 + 16 < npad) tc::tmem_ld16(tbase + 384 + lane_base + c0 + 16, v + 16);
          tc::tmem_wait_ld();
          if (active && !(a.dbg_skip & 1)) {
            if (permuted) {
              // column group k (4 columns) of this row: blk + k * rows * 4; 32-bit element offsets from one base
              float* pb = blk + (int64_t)(c0 >> 2) * rows * 4;
              const uint32_t gs4 = (uint32_t)rows * 4u;
              const int ng = min(8, (ncols - c0 + 3) >> 2);
              if (first) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                  if (k < ng)
                    *reinterpret_cast<float4*>(pb + (uint32_t)k * gs4) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
              } else {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                  if (k < ng) store4(pb + (uint32_t)k * gs4, v + 4 * k);
              }
            } else {
              // first Linear with n_features not a multiple of 4: canonical layout, live (conditioning) columns only
              float* dst = GL + D.off_W[it.lin] + (int64_t)t * ldc;
#pragma unroll
              for (int u = 0; u < 32; ++u)
                if (c0 + u < ncols && ((c0 + u + l) & 1) == 1) acc_to(dst + c0 + u, v[u]);
            }
          }
        }
      }
      // bias gradient = row sum of dY^T (both halves of the tile's samples)
      if (hf == 0 && active) {
        float bs = S->bsum[0][t];
#pragma unroll
        for (int h2 = 1; h2 < PARTS; ++h2) bs += S->bsum[h2][t];
        acc_to(GL + boff, bs);
      }
    };
    // masked coupling + ScalarAffine adjoints of layer l (acc 0 holds the conditioner-input gradient)
    auto finish_layer = [&](int l) {
      const float* PL = P + (int64_t)l * D.layer_stride;
      float* GL = PB + (int64_t)l * D.layer_stride;
      const float shift = PL[D.off_shift], e = expf(PL[D.off_scale]);
      const float* xin = a.save_x + ((int64_t)l * n + r) * d;
      float ssc = 0.0f, ssh = 0.0f;
      if (SPLIT) {
        // Exchange buffer X[j][row] = T (free here: the last unit's transposed reads are behind an epi_bar).  Every CTA
        // publishes, for its share of the columns: conditioning feature j -> ITS partial of the conditioner-input
        // gradient (acc 0); transformed feature j it owns -> the spline adjoint's dL/dx (already in gr[j]).
        const int fc = PR.fc;  // features per (full) chunk: 2 or 4
        for (int c = (j_lo / 16) * 16; c < j_hi; c += 16) {
          float v[16];
          tc::tmem_ld16(tbase + 256 + lane_base + c, v);
          tc::tmem_wait_ld();
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int j = c + u;
            if (j >= j_lo && j < j_hi) T[j * BT_TS + t] = (((j + l) & 1) == 1) ? v[u] : gr[j];
          }
        }
        BT_STAMP();        // (split) contributions written
        cluster_round(0);  // every CTA's contributions are in its T
        BT_STAMP();        // (split) round 0 complete
        // The peers' values come through distributed shared memory (~200 cycles a load): all loads of a block of JB
        // columns are issued before the first one is consumed.  fc (2 or 4) and R (2, 4 or 8) are powers of two: the
        // owner of transformed feature j is a shift and a mask -- the warp is latency-bound here and the two integer
        // divisions per column of the first version were most of the loop (profiles/r02_bt_timeline_c5_split.txt:
        // 38K -> 23K cycles per layer).
        constexpr int JB = 4;
        const int fc_shift = fc >= 4 ? 2 : 1;
        const uint32_t rmask = R - 1u;
        auto owner = [&](int j) -> uint32_t { return ((uint32_t)((j - (l & 1)) >> 1) >> fc_shift) & rmask; };
        // One copy of the loop per cluster size, unrolled over exactly RR ranks (a generic copy unrolled over 8 ranks with
        // `k < R` predicates was 150 instructions per column; the warp runs at IPC ~0.2 here).
        auto gather = [&](auto rr_tag) {
          constexpr uint32_t RR = decltype(rr_tag)::value;
          uint32_t tr[RR];  // this thread's column of T in every CTA of the cluster
#pragma unroll
          for (uint32_t k = 0; k < RR; ++k) tr[k] = t_remote[k];
          for (int j0 = j_lo; j0 < j_hi; j0 += JB) {
            float pv[JB][RR];  // [column][rank]
            float xv[JB];
#pragma unroll
            for (int u = 0; u < JB; ++u) {
              const int j = j0 + u;
              const bool cond = ((j + l) & 1) == 1;
              const uint32_t own = owner(j);
              xv[u] = j < j_hi ? xin[j] : 0.0f;
#pragma unroll
              for (uint32_t k = 0; k < RR; ++k) {
                pv[u][k] = 0.0f;
                if (j < j_hi && k != crank && (cond || k == own))
                  pv[u][k] = tc::ld_cluster_f32(tr[k] + 4u * (uint32_t)(j * BT_TS));
              }
            }
#pragma unroll
            for (int u = 0; u < JB; ++u) {
              const int j = j0 + u;
              if (j < j_hi) {
                float ga = gr[j];
                if (((j + l) & 1) == 1) {  // conditioning: dL/dy_j + the sum of the CTAs' partials, in rank order
                  float tot = 0.0f;
#pragma unroll
                  for (uint32_t k = 0; k < RR; ++k) tot += (k == crank) ? T[j * BT_TS + t] : pv[u][k];
                  ga += tot;
                } else {                   // transformed: the owner's dL/dx (this CTA's own gr[j] if it is the owner)
                  const uint32_t own = owner(j);
#pragma unroll
                  for (uint32_t k = 0; k < RR; ++k)
                    if (k == own && k != crank) ga = pv[u][k];
                }
                const float xa = (xv[u] + shift) * e;
                if (valid) {
                  ssc += ga * xa;
                  ssh += ga * e;
                }
                gr[j] = ga * e;
              }
            }
          }
        };
        if (R == 2) gather(std::integral_constant<uint32_t, 2>{});
        else if (R == 4) gather(std::integral_constant<uint32_t, 4>{});
        else gather(std::integral_constant<uint32_t, 8>{});
        BT_STAMP();        // (split) peers' values gathered
        cluster_round(1);  // every CTA has read what it needs: T may be rewritten by the next unit
        BT_STAMP();        // (split) round 1 complete
      } else
      for (int c = (j_lo / 16) * 16; c < j_hi; c += 16) {
        float v[16];
        tc::tmem_ld16(tbase + 256 + lane_base + c, v);  // N = pad16(d) columns
        tc::tmem_wait_ld();
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int j = c + u;
          if (j >= j_lo && j < j_hi) {
            float ga = gr[j];
            if (((j + l) & 1) == 1) ga += v[u];
            const float xa = (xin[j] + shift) * e;
            if (valid) {
              ssc += ga * xa;
              ssh += ga * e;
            }
            gr[j] = ga * e;
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ssc += __shfl_xor_sync(0xffffffffu, ssc, o);
        ssh += __shfl_xor_sync(0xffffffffu, ssh, o);
      }
      if (lane == 0) {
        S->red[warp] = ssc;
        S->red[EPI_WARPS + warp] = ssh;
      }
      epi_bar();
      if (tid == 0) {
        float sa = 0.0f, sb = 0.0f;
        for (int w = 0; w < EPI_WARPS; ++w) {
          sa += S->red[w];
          sb += S->red[EPI_WARPS + w];
        }
        const int n_valid = (int)min((int64_t)TC_M, n - row0);
        // (split: every CTA of the cluster computed the same sums from the exchanged gradient; rank 0 accounts for them)
        acc_to(GL + D.off_scale, crank == 0 ? sa - a.inv_n * (float)d * (float)n_valid : 0.0f);
        acc_to(GL + D.off_shift, crank == 0 ? sb : 0.0f);
      }
      epi_bar();  // gr complete for the next layer
    };

    // One loop, one call site per stage; the iteration after the last unit only drains: it finishes the last
    // layer and stores the last dW tile.
    int l = L - 1, ii = 0, pl = -1, pii = 0;  // current unit; previous unit (its dW tile is still in acc 1)
    while (ii < PR.n_items[l & 1] && skip_item(PR.items[l & 1][ii], ii)) ii += 2;  // split: the first unit this CTA owns
    int pending_pub = -1;                     // layer whose accumulator is complete but not yet published
    while (true) {
      const bool have = l >= 0;
      if (pl >= 0 && (!have || l != pl)) finish_layer(pl);  // [F] layer boundary: needs the dgrad of the last unit
      if (have) prepare(l, ii, pl >= 0);                    // [F] + [G] + [A]: overlaps the previous unit's wgrad MMAs,
      else if (pl >= 0) wait_item();                        //   hands the dgrad operand over chunk by chunk
      BT_STAMP();
      if (have) prefetch_next(l, ii);
      if (pl >= 0) reduce_unit(pl, pii);                    // [C] overlaps the dgrad MMAs
      if (last && a.done != nullptr) {
        // Layer pl of this CTA's accumulator is final after the store above: publish it to the reducer CTAs -- one
        // unit LATER (its stores have drained by then, so the fence is cheap), at once when nothing follows.
        if (pending_pub >= 0) {
          __threadfence();
          epi_bar();
          if (tid == 0) atomicAdd(a.done + pending_pub, 1);
          pending_pub = -1;
        }
        if (pl >= 0 && (!have || l != pl)) {
          if (have) {
            pending_pub = pl;
          } else {
            __threadfence();
            epi_bar();
            if (tid == 0) atomicAdd(a.done + pl, 1);
          }
        }
      }
      BT_STAMP();
      if (!have) break;
      wait_item();                                          // [D] dgrad done: the A region is free
      epi_bar();                                            // T of this unit complete (long since), bsum consumed
      BT_STAMP();
      {                                                     // [E] dY^T: lane = output unit t, columns = the tile's rows
        const float* src = T + t * BT_TS + hf * CPT;
        float bsum = 0.0f;
#pragma unroll
        for (int c = 0; c < CPT; c += 8) {
          const float4 v0 = *reinterpret_cast<const float4*>(src + c);
          const float4 v1 = *reinterpret_cast<const float4*>(src + c + 4);
          const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            bsum += v[u];
            tc::split_tf32_trunc(v[u], hi[u], lo[u]);
          }
          tc::tmem_st8(t_ahi + lane_base + hf * CPT + c, hi);
          tc::tmem_st8(t_alo + lane_base + hf * CPT + c, lo);
        }
        S->bsum[hf][t] = bsum;
      }
      tc::tmem_wait_st();
      tc::tc_fence_before();
      for (int kc = 0; kc < 4; ++kc) arrive_chunk(kc);      // -> wgrad MMAs start (K = the tile's 128 rows)
      epi_bar();                                            // T free for the next unit's prepare; bsum visible
      BT_STAMP();
      pl = l;
      pii = ii;
      ii += 2;
      while (ii < PR.n_items[l & 1] && skip_item(PR.items[l & 1][ii], ii)) ii += 2;
      if (ii >= PR.n_items[l & 1]) {
        --l;
        ii = 0;
        while (l >= 0 && ii < PR.n_items[l & 1] && skip_item(PR.items[l & 1][ii], ii)) ii += 2;
      }
    }
    }  // tiles
    tc::tc_fence_before();
  }
  __syncthreads();
  if (SPLIT) tc::cluster_sync();  // no CTA leaves while a peer may still read its shared memory / arrive on its barriers
  tc::tc_fence_after();
  if (warp == EPI_WARPS) tc::tmem_dealloc<512>(tbase);
  if (a.timing != nullptr && blockIdx.x == 0 && threadIdx.x == 0) a.timing[253] = clock64();  // CTA 0: tiles done
  }  // tile CTA
  if (a.done == nullptr) return;  // reduction by bt_reduce_kernel

  // ===== in-kernel reduction of the private accumulators ==============================================
  // The CTAs beyond n_cta (SMs the tiles leave idle) start here at once; the tile CTAs join when their tiles are
  // done.  Work = (layer, slice of the layer) chunks handed out in order L-1 .. 0 through an atomic counter; a chunk
  // of layer l can start as soon as every tile CTA has published that layer (a.done[l]): the sum over the accumulators
  // -- still in L2 -- overlaps the backward pass of the layers below, and only the tail runs on all SMs.  The
  // order of the sum inside an element is fixed (CTA order), whoever computes it.
  {
    __shared__ int s_chunk;
    const int Lr = D.n_layers;
    const int n_groups = (int)(D.layer_stride >> 2);
    const int gpc = 2 * (int)blockDim.x;
    const int per_layer = (n_groups + gpc - 1) / gpc;
    for (;;) {
      __syncthreads();
      if (threadIdx.x == 0) {
        const int c = atomicAdd(a.done + Lr, 1);
        if (c < Lr * per_layer) {
          const int* flag = a.done + (Lr - 1 - c / per_layer);
          int v;
          do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            if (v < a.n_cta) __nanosleep(128);
          } while (v < a.n_cta);
        }
        s_chunk = c;
      }
      __syncthreads();
      const int c = s_chunk;
      if (c >= Lr * per_layer) break;
      const int l = Lr - 1 - c / per_layer, ch = c % per_layer;
      bt_reduce_layer(D, a.partial, a.pstride, a.n_cta, a.grad, l, ch * gpc + (int)threadIdx.x, (int)blockDim.x,
                      min(n_groups, (ch + 1) * gpc), SPLIT ? a.split_r : 1, SPLIT ? PR.fc : 1);
      if (c == Lr * per_layer - 1 && threadIdx.x == 0) *a.loss = bt_reduce_loss(a.partial, a.pstride, a.n_cta);
    }
    if (a.timing != nullptr && blockIdx.x == 0 && threadIdx.x == 0) a.timing[254] = clock64();  // CTA 0: reduction done
  }
}

// grad = sum over CTAs of their private accumulators, in CTA order (deterministic).  The accumulators are read in
// their own (permuted) order, four consecutive floats per thread -- coalesced 16-byte loads, they are the bulk of the
// traffic -- and the sum goes to the canonical position (a permuted group of four is four consecutive columns of one
// canonical row).  Entries no CTA writes (alignment padding, W3 / b3 rows of the features a layer does not
// transform, W1 columns of the masked inputs) are recognised from the index and left at zero.
// Reduces the groups g = g0, g0 + gstep, ... < g_end of layer l.
// Feature split (R > 1 CTAs per tile, accumulator index = cluster * R + rank): the rows of W3 / b3 that belong to
// spline chunk c were written only by rank c % R of every cluster (fc = features per chunk); everything else holds a
// partial (or zero) in every accumulator.
__device__ __forceinline__ void bt_reduce_layer(const FlowmcFlowDesc& D, const float* partial, int64_t pstride,
                                                int n_cta, float* __restrict__ grad, int l, int g0, int gstep,
                                                int g_end, int R, int fc) {
  const int nh = D.n_linear - 1, NP = 3 * D.num_bins + 1, d = D.n_features;
  const int H = D.dims[nh];
  const bool w0_permuted = (d & 3) == 0;
  for (int g = g0; g < g_end; g += gstep) {
    const int64_t jo = (int64_t)g * 4;  // private offset inside the layer
    int64_t o = jo;                     // canonical offset of the group's first element
    bool live[4] = {false, false, false, false};
    int owner[4] = {-1, -1, -1, -1};    // feature split: the only rank that wrote the element (-1: all ranks did)
    if (jo >= D.off_W[nh] && jo < D.off_W[nh] + (int64_t)D.dims[nh + 1] * H) {
      const int jj = (int)(jo - D.off_W[nh]);
      const int f = jj / (NP * H), r2 = jj - f * NP * H;
      const int c = (r2 / (4 * NP)) * 4, rr = (r2 >> 2) % NP;
      o = D.off_W[nh] + (int64_t)(f * NP + rr) * H + c;
      live[0] = live[1] = live[2] = live[3] = ((f ^ l) & 1) == 0;
      if (R > 1) owner[0] = owner[1] = owner[2] = owner[3] = (((f - (l & 1)) >> 1) / fc) % R;
    } else if (jo >= D.off_W[0] && jo < D.off_W[0] + (int64_t)D.dims[1] * d) {
      if (w0_permuted) {
        const int jj = (int)(jo - D.off_W[0]), N = D.dims[1];
        const int c = (jj / (4 * N)) * 4;
        o = D.off_W[0] + (int64_t)((jj >> 2) % N) * d + c;
        live[0] = live[2] = ((c + l) & 1) == 1;
        live[1] = live[3] = !live[0];
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          live[e] = jo + e < D.off_W[0] + (int64_t)D.dims[1] * d && ((((int)((jo + e - D.off_W[0]) % d)) + l) & 1) == 1;
      }
    } else {
      bool found = false;
      for (int k = 1; k < nh; ++k) {
        if (jo >= D.off_W[k] && jo < D.off_W[k] + (int64_t)D.dims[k + 1] * D.dims[k]) {
          const int jj = (int)(jo - D.off_W[k]), N = D.dims[k + 1];
          o = D.off_W[k] + (int64_t)((jj >> 2) % N) * D.dims[k] + (jj / (4 * N)) * 4;
          live[0] = live[1] = live[2] = live[3] = true;
          found = true;
          break;
        }
      }
      if (!found) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int64_t oe = jo + e;
          if (oe >= D.off_b[nh] && oe < D.off_b[nh] + D.dims[nh + 1]) {
            const int f = (int)(oe - D.off_b[nh]) / NP;
            live[e] = ((f ^ l) & 1) == 0;
            if (R > 1 && live[e]) owner[e] = (((f - (l & 1)) >> 1) / fc) % R;
          } else if (oe == D.off_scale || oe == D.off_shift) {
            live[e] = true;
          } else {
            for (int k = 0; k < nh; ++k)
              if (oe >= D.off_b[k] && oe < D.off_b[k] + D.dims[k + 1]) live[e] = true;
          }
        }
      }
    }
    if (R > 1 && (owner[0] >= 0 || owner[1] >= 0 || owner[2] >= 0 || owner[3] >= 0)) {
      // owner-written elements: sum over the clusters, reading the owning rank's accumulator of each
      const float* base = partial + (int64_t)l * D.layer_stride + jo;
      float sv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      const int n_cl = n_cta / R;
      if (owner[0] == owner[1] && owner[0] == owner[2] && owner[0] == owner[3]) {
        for (int cl = 0; cl < n_cl; ++cl) {
          const float4 v = __ldcg(reinterpret_cast<const float4*>(base + (int64_t)(cl * R + owner[0]) * pstride));
          sv[0] += v.x; sv[1] += v.y; sv[2] += v.z; sv[3] += v.w;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (!live[e]) continue;
          if (owner[e] >= 0) {
            for (int cl = 0; cl < n_cl; ++cl) sv[e] += __ldcg(base + (int64_t)(cl * R + owner[e]) * pstride + e);
          } else {
            for (int c2 = 0; c2 < n_cta; ++c2) sv[e] += __ldcg(base + (int64_t)c2 * pstride + e);
          }
        }
      }
      float* dst = grad + (int64_t)l * D.layer_stride + o;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (live[e]) dst[e] = sv[e];
    } else if (live[0] | live[1] | live[2] | live[3]) {
      const float4* src = reinterpret_cast<const float4*>(partial + (int64_t)l * D.layer_stride + jo);
      const int64_t step4 = pstride >> 2;
      float4 s = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      int c = 0;
      for (; c + 16 <= n_cta; c += 16) {  // 16 independent 16-byte loads in flight, summed in CTA order
        float4 v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = __ldcg(src + (int64_t)(c + u) * step4);
#pragma unroll
        for (int u = 0; u < 16; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
      }
      for (; c < n_cta; ++c) {
        const float4 v = __ldcg(src + (int64_t)c * step4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      float* dst = grad + (int64_t)l * D.layer_stride + o;
      if (live[0] & live[1] & live[2] & live[3]) {
        *reinterpret_cast<float4*>(dst) = s;
      } else {
        if (live[0]) dst[0] = s.x;
        if (live[1]) dst[1] = s.y;
        if (live[2]) dst[2] = s.z;
        if (live[3]) dst[3] = s.w;
      }
    }
  }
}

__device__ __forceinline__ float bt_reduce_loss(const float* partial, int64_t pstride, int n_cta) {
  float s = 0.0f;
  for (int c = 0; c < n_cta; ++c) s += __ldcg(partial + (int64_t)c * pstride + pstride - 4);
  return s;
}

// stand-alone reduction (used when the backward kernel has no spare SMs for in-kernel reducers)
__global__ void bt_reduce_kernel(const FlowmcFlowDesc D, const float* __restrict__ partial, int64_t pstride, int n_cta,
                                 float* __restrict__ grad, float* __restrict__ loss, int R, int fc) {
  for (int l = blockIdx.y; l < D.n_layers; l += gridDim.y)
    bt_reduce_layer(D, partial, pstride, n_cta, grad, l, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x,
                    (int)(D.layer_stride >> 2), R, fc);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *loss = bt_reduce_loss(partial, pstride, n_cta);
}

static int bt_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = BT_MAX_CTAS;
    if (sms > BT_MAX_CTAS) sms = BT_MAX_CTAS;
  }
  return sms;
}

template <int KB, int PARTS>
static int launch_bt(const FlowmcFlowDesc& D, const BtProgram& PR, const BtArgs& a, cudaStream_t stream) {
  auto kern = flow_backward_tc_kernel<KB, PARTS>;
  const size_t bytes = 1024 + (size_t)BT_STAGES * TC_STAGE_BYTES + ((sizeof(BtSmem<PARTS>) + 15) & ~15) +
                       (size_t)128 * BT_TS * sizeof(float) + (size_t)TC_M * (D.n_features + 1) * sizeof(float);
  static size_t configured = 0;
  if (bytes > configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
      flowmc_set_error("flow backward (tensor-core path): cannot configure shared memory");
      return FLOWMC_ERR_CUDA;
    }
    configured = bytes;
  }
  // Spare SMs (fewer tiles than SMs) become in-kernel reducers: the sum over the private accumulators then overlaps
  // the backward pass instead of following it.  All CTAs are co-resident (one per SM), so the reducers' spin-wait on
  // the tile CTAs cannot starve them.
  BtArgs b = a;
  b.n_cta = (int)(a.n_tiles < BT_MAX_CTAS ? a.n_tiles : BT_MAX_CTAS);
  const int n_red = bt_sm_count() - b.n_cta;
  const bool fused = n_red >= 8;
  if (!fused) b.done = nullptr;
  kern<<<b.n_cta + (fused ? n_red : 0), 4 * PARTS * 32 + 64, bytes, stream>>>(D, PR, b);
  flowmc_count_launch();
  if (!fused) {
    bt_reduce_kernel<<<dim3(32, D.n_layers), 256, 0, stream>>>(D, b.partial, b.pstride, b.n_cta, b.grad, b.loss, 1, 1);
    flowmc_count_launch();
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    flowmc_set_error(cudaGetErrorString(e));
    return FLOWMC_ERR_CUDA;
  }
  return FLOWMC_OK;
}

// Feature split: clusters of R CTAs per tile (see the kernel); the accumulators are reduced by the stand-alone kernel.
template <int KB>
static int launch_bt_split(const FlowmcFlowDesc& D, const BtProgram& PR, const BtArgs& a, int R, cudaStream_t stream) {
  auto kern = flow_backward_tc_kernel<KB, 2, true>;
  const size_t bytes = 1024 + (size_t)BT_STAGES * TC_STAGE_BYTES + ((sizeof(BtSmem<2>) + 15) & ~15) +
                       (size_t)128 * BT_TS * sizeof(float) + (size_t)TC_M * (D.n_features + 1) * sizeof(float);
  static size_t configured = 0;
  if (bytes > configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess ||
        cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      flowmc_set_error("flow backward (tensor-core path, feature split): cannot configure the kernel");
      return FLOWMC_ERR_CUDA;
    }
    configured = bytes;
  }
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(4 * 2 * 32 + 64);
  cfg.dynamicSmemBytes = bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)R;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static int max_clusters[9] = {0};
  if (max_clusters[R] == 0) {
    cfg.gridDim = dim3((unsigned)R);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = 148 / (R == 8 ? 12 : R);
    }
    if (n * R > BT_MAX_CTAS) n = BT_MAX_CTAS / R;
    max_clusters[R] = n;
    tc_split_note_max_clusters(R, n);
  }
  BtArgs b = a;
  const int n_clusters = (int)(a.n_tiles < max_clusters[R] ? a.n_tiles : max_clusters[R]);
  b.split_r = R;
  b.n_cta = n_clusters * R;
  // clusters the tiles leave free become in-kernel reducers, like the spare CTAs of the one-CTA-per-tile launch (all
  // clusters of the grid are co-resident: the reducers' spin-wait cannot starve a tile cluster)
  const int n_red = (max_clusters[R] - n_clusters) * R;
  const bool fused = n_red >= 8 && a.done != nullptr;
  if (!fused) b.done = nullptr;
  cfg.gridDim = dim3((unsigned)(b.n_cta + (fused ? n_red : 0)));
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, D, PR, b);
  flowmc_count_launch();
  if (e == cudaSuccess && !fused) {
    bt_reduce_kernel<<<dim3(32, D.n_layers), 256, 0, stream>>>(D, b.partial, b.pstride, b.n_cta, b.grad, b.loss, R,
                                                               PR.fc);
    flowmc_count_launch();
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    flowmc_set_error(cudaGetErrorString(e));
    return FLOWMC_ERR_CUDA;
  }
  return FLOWMC_OK;
}

static long long* g_bt_timing = nullptr;
void flow_backward_tc_set_timing(long long* buf) { g_bt_timing = buf; }

bool flow_backward_tc_supported(const FlowmcFlowDesc& D) {
  if (!tc_supported(D)) return false;
  if (D.num_bins == 16) return false;  // 49 parameters per feature do not fit the 32-column A slot
  const size_t bytes = 4096 + (size_t)BT_STAGES * TC_STAGE_BYTES + (size_t)128 * BT_TS * 4 +
                       (size_t)TC_M * (D.n_features + 1) * 4;
  return bytes <= 227 * 1024;
}

int64_t flow_backward_tc_wimg_bytes(const FlowmcFlowDesc& D) {
  BtProgram PR;
  if (bt_build_program(D, &PR)) return 0;
  int64_t b = 0;
  for (int l = 0; l < D.n_layers; ++l) b += PR.layer_bytes[l & 1];
  return b;
}

int64_t flow_backward_tc_act_bytes(const FlowmcFlowDesc& D, int64_t n) {
  return ((n + TC_M - 1) / TC_M) * (int64_t)D.n_layers * (int64_t)tc_act_layer_bytes(D);
}

int64_t flow_backward_tc_partial_bytes(const FlowmcFlowDesc& D, int64_t n) {
  // one accumulator per CTA of the backward kernel: up to one per SM (with the feature split a handful of tiles
  // still occupies most SMs)
  (void)n;
  return (int64_t)BT_MAX_CTAS * ((int64_t)D.n_layers * D.layer_stride + 4) * 4 +
         (int64_t)(D.n_layers + 4) * 4;  // + the per-layer hand-off counters
}

int flow_backward_tc(const FlowmcFlowDesc& D, const float* params, uint8_t* wimg, const uint8_t* act_img,
                     const float* save_x, const float* save_theta, const float* logp, int64_t n, float inv_n,
                     float* grad, float* loss, float* partial, cudaStream_t stream) {
  BtProgram PR;
  if (int rc = bt_build_program(D, &PR)) return rc;
  const int items = PR.n_items[0] > PR.n_items[1] ? PR.n_items[0] : PR.n_items[1];
  const int64_t tiles = (n + TC_M - 1) / TC_M;
  const int64_t pstride = (int64_t)D.n_layers * D.layer_stride + 4;
  int* done = reinterpret_cast<int*>(partial + (int64_t)BT_MAX_CTAS * pstride);
  bt_pack_kernel<<<dim3(items, D.n_layers, 8), 256, 0, stream>>>(D, PR, params, wimg, done);
  flowmc_count_launch();
  BtArgs a;
  a.params = params; a.wimg = wimg; a.act_img = act_img; a.save_x = save_x; a.save_theta = save_theta; a.logp = logp;
  a.n = n; a.inv_n = inv_n; a.grad = grad; a.loss = loss; a.timing = g_bt_timing;
  a.partial = partial; a.pstride = pstride; a.n_tiles = tiles; a.done = done; a.n_cta = 0; a.split_r = 1;
  static const int dbg_skip = [] {
    const char* e = std::getenv("FLOWMC_BT_DBG_SKIP");
    return e != nullptr ? std::atoi(e) : 0;
  }();
  a.dbg_skip = dbg_skip;
  {
    // few tiles (a data-parallel rank's slice of the batch): clusters of CTAs share a tile, as in the forward pass
    const int R = tc_split_factor(D, tiles);
    if (R > 1 && D.num_bins <= 8) {
      switch (D.num_bins) {
        case 4: return launch_bt_split<4>(D, PR, a, R, stream);
        case 8: return launch_bt_split<8>(D, PR, a, R, stream);
      }
    }
  }
  // epilogue threads per sample row: 2.  FLOWMC_BT_PARTS=4 selects 16 epilogue warps -- measured SLOWER on B200
  // (profiles/r02_prof_train_c4_parts4.txt: C4 backward 562 vs 512 us, C5 783 vs 749 us): the epilogue is a
  // latency-bound serial program (IPC 0.22 per warp) that more warps would help, but a 640-thread CTA
  // leaves 96 registers per thread and the adjoint stage then spills 240 bytes per thread.  Kept for A/B runs.
  static const int parts = [] {
    const char* e = std::getenv("FLOWMC_BT_PARTS");
    return (e != nullptr && e[0] == '4') ? 4 : 2;
  }();
  switch (D.num_bins * 10 + parts) {
    case 42: return launch_bt<4, 2>(D, PR, a, stream);
    case 44: return launch_bt<4, 4>(D, PR, a, stream);
    case 82: return launch_bt<8, 2>(D, PR, a, stream);
    case 84: return launch_bt<8, 4>(D, PR, a, stream);
  }
  return FLOWMC_ERR_UNSUPPORTED;
}

}  // namespace flowmc
