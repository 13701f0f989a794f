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
//   the dW tile (lane = output unit) is read back with tcgen05.ld and reduced over tiles with red.global.add.f32.
//
// One CTA = 128 samples, layers in reverse.  Per layer: for every chunk of 4 transformed features
// {spline adjoints -> dtheta; dh_last += dtheta W3_c (acc slot 0); dW3_c = dtheta^T h_last (acc slot 1)}, then per
// tanh layer {da = dh (1 - h^2); dh_prev = da W; dW = da^T h_prev}, masked-coupling and ScalarAffine adjoints.
// Epilogue and MMA strictly alternate (they share the A region of tensor memory); the weight producer prefetches
// the next items' stages through a 3-deep ring meanwhile.
// TMEM map: [0,128) A hi | [128,256) A lo | [256,384) acc 0 (data gradient) | [384,512) acc 1 (weight gradient).
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
};
struct BtProgram {
  int n_items[2];
  uint32_t layer_bytes[2];
  int fc;
  int pad_;
  BtItem items[2][BT_MAX_ITEMS];
};

static int bt_build_program(const FlowmcFlowDesc& D, BtProgram* P) {
  const int d = D.n_features, NP = 3 * D.num_bins + 1, nh = D.n_linear - 1;
  int fc = (128 / NP) & ~1;
  if (fc > 4) fc = 4;  // one 32-column slot of the A region per feature
  P->fc = fc;
  P->pad_ = 0;
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
      off += (uint32_t)dg.n_kc * 2u * dg.N * 128u;
      BtItem& wg = P->items[p][n++];
      wg.kind = BK_WG3; wg.N = H; wg.n_kc = 4; wg.K = 128; wg.lin = c0; wg.n_feat = nf; wg.act = 1;
      wg.off = (uint32_t)tc_act_item_off(D, nh);  // h_last
    }
    for (int i = nh - 1; i >= 0; --i) {
      if (n + 2 > BT_MAX_ITEMS) return FLOWMC_ERR_UNSUPPORTED;
      const int Nin = (i == 0) ? tc_pad16(d) : D.dims[i];
      BtItem& dg = P->items[p][n++];
      dg.kind = BK_DGH; dg.N = Nin; dg.K = D.dims[i + 1]; dg.n_kc = (dg.K + 31) / 32; dg.lin = i; dg.n_feat = 0;
      dg.act = 0; dg.off = off;
      off += (uint32_t)dg.n_kc * 2u * dg.N * 128u;
      BtItem& wg = P->items[p][n++];
      wg.kind = BK_WGH; wg.N = Nin; wg.n_kc = 4; wg.K = 128; wg.lin = i; wg.n_feat = 0; wg.act = 1;
      wg.off = (uint32_t)tc_act_item_off(D, i);  // i == 0: x * mask, else h_{i-1}
    }
    P->n_items[p] = n;
    P->layer_bytes[p] = off;
  }
  return FLOWMC_OK;
}

__host__ __device__ inline int64_t bt_layer_base(const BtProgram& P, int l) {
  return (int64_t)((l + 1) / 2) * P.layer_bytes[0] + (int64_t)(l / 2) * P.layer_bytes[1];
}

// transposed weight images for the data-gradient GEMMs.  grid = (items, layers)
__global__ void bt_pack_kernel(const FlowmcFlowDesc D, const BtProgram P, const float* __restrict__ params,
                               uint8_t* __restrict__ image) {
  const int l = blockIdx.y, p = l & 1;
  if ((int)blockIdx.x >= P.n_items[p]) return;
  const BtItem it = P.items[p][blockIdx.x];
  if (it.act) return;
  const int NP = 3 * D.num_bins + 1, nh = D.n_linear - 1;
  const float* PL = params + (int64_t)l * D.layer_stride;
  float* dst = reinterpret_cast<float*>(image + bt_layer_base(P, l) + it.off);
  const int per_stage = it.N * 32;
  const int total = it.n_kc * per_stage;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
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
    tc::split_tf32(w, hi, lo);
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
};

#define BT_STAMP()                                                                                \
  do {                                                                                            \
    if (a.timing != nullptr && blockIdx.x == 0 && tid == 0 && n_stamp < 256) a.timing[n_stamp++] = clock64(); \
  } while (0)

struct BtSmem {
  uint64_t stage_full[BT_STAGES], stage_empty[BT_STAGES], acc_full, a_ready;
  uint32_t tmem_base;
  float red[2 * TC_EPI_WARPS];
  float bsum[TC_PARTS][TC_M];
};

template <int KB>
__global__ void __launch_bounds__(TC_THREADS, 1) flow_backward_tc_kernel(const FlowmcFlowDesc D, const BtProgram PR,
                                                                         const BtArgs a) {
  constexpr int NP = 3 * KB + 1;
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

  if (warp == TC_EPI_WARPS + 1 && lane == 0) {
    for (int i = 0; i < BT_STAGES; ++i) {
      tc::mbar_init(&S->stage_full[i], 1);
      tc::mbar_init(&S->stage_empty[i], 1);
    }
    tc::mbar_init(&S->acc_full, 1);
    tc::mbar_init(&S->a_ready, TC_EPI);
    tc::fence_mbar_init();
  }
  if (warp == TC_EPI_WARPS) tc::tmem_alloc<512>(&S->tmem_base);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tbase = S->tmem_base;
  const uint32_t t_ahi = tbase, t_alo = tbase + 128;

  if (warp == TC_EPI_WARPS) {
    // ===== B-stage producer ====================================================================
    uint32_t s = 0, ph = 0;
    for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x)
    for (int l = L - 1; l >= 0; --l) {
      const int p = l & 1;
      const uint8_t* wbase = a.wimg + bt_layer_base(PR, l);
      const uint8_t* abase = a.act_img + (tile * L + l) * tc_act_layer_bytes(D);
      for (int ii = 0; ii < PR.n_items[p]; ++ii) {
        const BtItem it = PR.items[p][ii];
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
  } else if (warp == TC_EPI_WARPS + 1) {
    // ===== MMA issuer ==========================================================================
    uint32_t s = 0, ph = 0, a_ph = 0;
    for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x)
    for (int l = L - 1; l >= 0; --l) {
      const int p = l & 1;
      for (int ii = 0; ii < PR.n_items[p]; ++ii) {
        const BtItem it = PR.items[p][ii];
        tc::mbar_wait(&S->a_ready, a_ph);  // the epilogue has written this item's A operand
        a_ph ^= 1;
        tc::tc_fence_after();
        const bool wg = (it.kind == BK_WG3) || (it.kind == BK_WGH);
        const uint32_t t_acc = tbase + (wg ? 384 : 256);
        const uint32_t idesc = tc::make_idesc_tf32(TC_M, it.N);
        const bool cont = (it.kind == BK_DG3) && (it.lin > 0);  // later chunks accumulate into dh_last
        for (int kc = 0; kc < it.n_kc; ++kc) {
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
      const int per = (cnt + TC_PARTS - 1) / TC_PARTS;
      lo = min(cnt, hf * per);
      hi = min(cnt, lo + per);
    };
    int j_lo, j_hi;
    part(d, j_lo, j_hi);
    float* PB = a.partial + (int64_t)blockIdx.x * a.pstride;  // private accumulator of this CTA

    // hands the A operand to the MMA warp and waits for the item's accumulator
    auto run_item = [&]() {
      tc::tmem_wait_st();
      tc::tc_fence_before();
      tc::mbar_arrive(&S->a_ready);
      tc::mbar_wait(&S->acc_full, f_ph);
      f_ph ^= 1;
      tc::tc_fence_after();
    };
    // A^T: output unit m = t reads its row of the transpose buffer (this thread: half of the tile's samples),
    // writes it as TMEM lane m, and returns the row sum (= the bias gradient of unit m)
    auto write_transposed = [&]() -> float {
      float bsum = 0.0f;
      const float* src = T + t * BT_TS;
      for (int c = hf * 64; c < hf * 64 + 64; c += 8) {
        const float4 v0 = *reinterpret_cast<const float4*>(src + c);
        const float4 v1 = *reinterpret_cast<const float4*>(src + c + 4);
        const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          bsum += v[u];
          tc::split_tf32(v[u], hi[u], lo[u]);
        }
        tc::tmem_st8(t_ahi + lane_base + c, hi);
        tc::tmem_st8(t_alo + lane_base + c, lo);
      }
      return bsum;
    };

    // acc 1 (lane = output unit m, columns [0, ncols)) -> T[m][c] row-major, so that the rows can then be written to
    // global memory with full 512-byte coalescing (a thread owning a whole row would touch 32 sectors per store)
    auto stage_acc1 = [&](int ncols) {
      for (int c = hf * 64; c < min(ncols, hf * 64 + 64); c += 16) {
        float v[16];
        tc::tmem_ld16(tbase + 384 + lane_base + c, v);
        tc::tmem_wait_ld();
#pragma unroll
        for (int u = 0; u < 16; u += 4)
          *reinterpret_cast<float4*>(T + t * BT_TS + c + u) = make_float4(v[u], v[u + 1], v[u + 2], v[u + 3]);
      }
    };
    for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const bool first = tile == (int64_t)blockIdx.x;  // first tile of this CTA: store, later tiles: accumulate
    auto acc_to = [&](float* ptr, float v) { *ptr = first ? v : *ptr + v; };
    // one warp per row of the staged dW tile: lanes cover 4 consecutive columns each
    auto write_row = [&](int m, float* dst, int ncols) {
      const int c = lane * 4;
      if (c < ncols) {
        float4 v = *reinterpret_cast<const float4*>(T + m * BT_TS + c);
        float4* gp = reinterpret_cast<float4*>(dst + c);
        if (!first) {
          const float4 o = *gp;
          v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
        *gp = v;
      }
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
    if (tid == 0) acc_to(PB + a.pstride - 4, (S->red[0] + S->red[1]) + (S->red[2] + S->red[3]));
    for (int l = L - 1; l >= 0; --l) {
      const int p = l & 1;
      const float* PL = P + (int64_t)l * D.layer_stride;
      float* GL = PB + (int64_t)l * D.layer_stride;
      const float scale = PL[D.off_scale], shift = PL[D.off_shift];
      const float e = expf(scale);
      const float* xin = a.save_x + ((int64_t)l * n + r) * d;  // this row's layer input (before the ScalarAffine)
      const uint8_t* abase = a.act_img + (tile * L + l) * tc_act_layer_bytes(D);
      const int H = D.dims[nh];
      int ii = 0;
      // ---- spline chunks -----------------------------------------------------------------------
      for (; ii < PR.n_items[p] && PR.items[p][ii].kind == BK_DG3; ii += 2) {
        const BtItem it = PR.items[p][ii];
        int i_lo, i_hi;
        part(it.n_feat, i_lo, i_hi);
        for (int fi = i_lo; fi < i_hi; ++fi) {
          const int fo = it.lin + fi, f = p + 2 * fo;
          float raw[NP], dr[NP + 7], gx;
          const float* th = a.save_theta + ((int64_t)l * ((d + 1) / 2) + fo) * NP * n + r;
#pragma unroll
          for (int u = 0; u < NP; ++u) raw[u] = th[(int64_t)u * n];
          const float xa = (xin[f] + shift) * e;
          rq_backward<KB, true>(raw, D.range_min, D.range_max, xa, gr[f], gld, gx, dr);
          gr[f] = gx;
          // dtheta -> A (lane = this row, columns fi*32 .. fi*32+31) and the transpose buffer T[column][row]
          if (NP <= 32) {
#pragma unroll
            for (int u = NP; u < NP + 7; ++u) dr[u] = 0.0f;
#pragma unroll
            for (int c = 0; c < 32; c += 8) {
              uint32_t hi[8], lo[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const float v = (c + u < NP) ? dr[(c + u < NP) ? c + u : 0] : 0.0f;
                tc::split_tf32(v, hi[u], lo[u]);
                T[(fi * 32 + c + u) * BT_TS + t] = v;
              }
              tc::tmem_st8(t_ahi + lane_base + fi * 32 + c, hi);
              tc::tmem_st8(t_alo + lane_base + fi * 32 + c, lo);
            }
          }
        }
        BT_STAMP();  // spline adjoints + operand written
        run_item();  // dh_last (+)= dtheta W3_c   (acc 0)
        BT_STAMP();  // data-gradient MMAs done
        epi_bar();   // T complete
        S->bsum[hf][t] = write_transposed();
        BT_STAMP();  // transposed operand written
        run_item();  // dW3_c = dtheta^T h_last   (acc 1)
        BT_STAMP();  // weight-gradient MMAs done
        stage_acc1(H);
        epi_bar();
        for (int m = warp; m < it.n_feat * 32; m += TC_EPI_WARPS) {
          const int fi = m >> 5, rr = m & 31;
          if (rr < NP) {
            const int64_t prow = (int64_t)(p + 2 * (it.lin + fi)) * NP + rr;
            write_row(m, GL + D.off_W[nh] + prow * H, H);
            if (lane == 0)  // bias gradient = row sum of dtheta^T (both halves of the tile's samples)
              acc_to(GL + D.off_b[nh] + prow, S->bsum[0][m] + S->bsum[1][m]);
          }
        }
        epi_bar();  // T, bsum and acc 1 free for the next chunk
        BT_STAMP();  // dW tile reduced into the gradient
      }
      // ---- tanh layers in reverse ----------------------------------------------------------------
      for (; ii < PR.n_items[p]; ii += 2) {
        const BtItem it = PR.items[p][ii];
        const int i = it.lin;
        const int N = D.dims[i + 1];  // width of this hidden layer
        // da = dh (1 - h^2): dh from acc 0, h from the forward pass's activation image (hi + lo)
        {
          const uint32_t* himg = reinterpret_cast<const uint32_t*>(abase + tc_act_item_off(D, i + 1) + (size_t)q * 2 * N * 128);
          int c_lo, c_hi;
          part(N / 16, c_lo, c_hi);
          for (int c = c_lo * 16; c < c_hi * 16; c += 16) {
            float v[16];
            tc::tmem_ld16(tbase + 256 + lane_base + c, v);
            tc::tmem_wait_ld();
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const int o = tc::packed_b_offset(c + u, lane) >> 2;
              const float hv = __uint_as_float(himg[o]) + __uint_as_float(himg[N * 32 + o]);
              const float da = v[u] * (1.0f - hv * hv);
              tc::split_tf32(da, hi[u], lo[u]);
              T[(c + u) * BT_TS + t] = da;
            }
            tc::tmem_st8(t_ahi + lane_base + c, hi);
            tc::tmem_st8(t_ahi + lane_base + c + 8, hi + 8);
            tc::tmem_st8(t_alo + lane_base + c, lo);
            tc::tmem_st8(t_alo + lane_base + c + 8, lo + 8);
          }
        }
        run_item();  // dh_prev = da W_i  (acc 0; for i == 0: the conditioner-input gradient)
        epi_bar();
        S->bsum[hf][t] = write_transposed();
        run_item();  // dW_i = da^T in_i   (acc 1)
        {
          const int Kin = D.dims[i];
          stage_acc1(tc_pad16(Kin));
          epi_bar();
          for (int m = warp; m < N; m += TC_EPI_WARPS) {
            float* dst = GL + D.off_W[i] + (int64_t)m * Kin;
            if (i > 0) {
              write_row(m, dst, Kin);  // hidden widths are multiples of 16: whole float4 columns
            } else {
              // first Linear: only the conditioning inputs ((j + l) odd) carry gradient; d need not be a multiple of 4
              for (int c = lane; c < Kin; c += 32)
                if (((c + l) & 1) == 1) acc_to(dst + c, T[m * BT_TS + c]);
            }
            if (lane == 0) acc_to(GL + D.off_b[i] + m, S->bsum[0][m] + S->bsum[1][m]);
          }
        }
        epi_bar();
      }
      // ---- masked coupling + ScalarAffine adjoints -------------------------------------------------
      {
        float ssc = 0.0f, ssh = 0.0f;
        for (int c = (j_lo / 16) * 16; c < j_hi; c += 16) {
          float v[16];
          tc::tmem_ld16(tbase + 256 + lane_base + c, v);  // conditioner-input gradient (acc 0, N = pad16(d))
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
          S->red[TC_EPI_WARPS + warp] = ssh;
        }
        epi_bar();
        if (tid == 0) {
          float sa = 0.0f, sb = 0.0f;
          for (int w = 0; w < TC_EPI_WARPS; ++w) {
            sa += S->red[w];
            sb += S->red[TC_EPI_WARPS + w];
          }
          const int n_valid = (int)min((int64_t)TC_M, n - row0);
          acc_to(GL + D.off_scale, sa - a.inv_n * (float)d * (float)n_valid);
          acc_to(GL + D.off_shift, sb);
        }
        epi_bar();
      }
    }
    }  // tiles
    tc::tc_fence_before();
  }
  __syncthreads();
  tc::tc_fence_after();
  if (warp == TC_EPI_WARPS) tc::tmem_dealloc<512>(tbase);
}

// grad[i] = sum over CTAs of their private accumulators, in CTA order (deterministic).  Entries no CTA writes
// (alignment padding, W3 / b3 rows of the features a layer does not transform, W1 columns of the masked inputs) are
// recognised from the index and left at zero.
__global__ void bt_reduce_kernel(const FlowmcFlowDesc D, const float* __restrict__ partial, int64_t pstride, int n_cta,
                                 float* __restrict__ grad, float* __restrict__ loss) {
  const int nh = D.n_linear - 1, NP = 3 * D.num_bins + 1, d = D.n_features;
  const int64_t total = (int64_t)D.n_layers * D.layer_stride;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int l = (int)(i / D.layer_stride);
    const int64_t o = i - (int64_t)l * D.layer_stride;
    bool live = false;
    if (o >= D.off_W[nh] && o < D.off_W[nh] + (int64_t)D.dims[nh + 1] * D.dims[nh]) {
      const int f = (int)((o - D.off_W[nh]) / D.dims[nh]) / NP;
      live = ((f ^ l) & 1) == 0;
    } else if (o >= D.off_b[nh] && o < D.off_b[nh] + D.dims[nh + 1]) {
      live = ((((int)(o - D.off_b[nh]) / NP) ^ l) & 1) == 0;
    } else if (o >= D.off_W[0] && o < D.off_W[0] + (int64_t)D.dims[1] * d) {
      live = ((((int)((o - D.off_W[0]) % d)) + l) & 1) == 1;
    } else if (o == D.off_scale || o == D.off_shift) {
      live = true;
    } else {
      for (int k = 0; k < nh; ++k) {
        if (k > 0 && o >= D.off_W[k] && o < D.off_W[k] + (int64_t)D.dims[k + 1] * D.dims[k]) live = true;
        if (o >= D.off_b[k] && o < D.off_b[k] + D.dims[k + 1]) live = true;
      }
    }
    if (live) {
      float s = 0.0f;
      for (int c = 0; c < n_cta; ++c) s += partial[(int64_t)c * pstride + i];
      grad[i] = s;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float s = 0.0f;
    for (int c = 0; c < n_cta; ++c) s += partial[(int64_t)c * pstride + pstride - 4];
    *loss = s;
  }
}

template <int KB>
static int launch_bt(const FlowmcFlowDesc& D, const BtProgram& PR, const BtArgs& a, cudaStream_t stream) {
  auto kern = flow_backward_tc_kernel<KB>;
  const size_t bytes = 1024 + (size_t)BT_STAGES * TC_STAGE_BYTES + ((sizeof(BtSmem) + 15) & ~15) +
                       (size_t)128 * BT_TS * sizeof(float) + (size_t)TC_M * (D.n_features + 1) * sizeof(float);
  static size_t configured = 0;
  if (bytes > configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
      flowmc_set_error("flow backward (tensor-core path): cannot configure shared memory");
      return FLOWMC_ERR_CUDA;
    }
    configured = bytes;
  }
  const int n_cta = (int)(a.n_tiles < BT_MAX_CTAS ? a.n_tiles : BT_MAX_CTAS);
  kern<<<n_cta, TC_THREADS, bytes, stream>>>(D, PR, a);
  flowmc_count_launch();
  bt_reduce_kernel<<<148 * 4, 256, 0, stream>>>(D, a.partial, a.pstride, n_cta, a.grad, a.loss);
  flowmc_count_launch();
  cudaError_t e = cudaGetLastError();
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
  const size_t bytes = 2048 + (size_t)BT_STAGES * TC_STAGE_BYTES + (size_t)128 * BT_TS * 4 +
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
  const int64_t tiles = (n + TC_M - 1) / TC_M;
  return (tiles < BT_MAX_CTAS ? tiles : BT_MAX_CTAS) * ((int64_t)D.n_layers * D.layer_stride + 4) * 4;
}

int flow_backward_tc(const FlowmcFlowDesc& D, const float* params, uint8_t* wimg, const uint8_t* act_img,
                     const float* save_x, const float* save_theta, const float* logp, int64_t n, float inv_n,
                     float* grad, float* loss, float* partial, cudaStream_t stream) {
  BtProgram PR;
  if (int rc = bt_build_program(D, &PR)) return rc;
  const int items = PR.n_items[0] > PR.n_items[1] ? PR.n_items[0] : PR.n_items[1];
  bt_pack_kernel<<<dim3(items, D.n_layers), 256, 0, stream>>>(D, PR, params, wimg);
  flowmc_count_launch();
  BtArgs a;
  a.params = params; a.wimg = wimg; a.act_img = act_img; a.save_x = save_x; a.save_theta = save_theta; a.logp = logp;
  a.n = n; a.inv_n = inv_n; a.grad = grad; a.loss = loss; a.timing = g_bt_timing;
  a.partial = partial; a.pstride = (int64_t)D.n_layers * D.layer_stride + 4; a.n_tiles = (n + TC_M - 1) / TC_M;
  switch (D.num_bins) {
    case 4: return launch_bt<4>(D, PR, a, stream);
    case 8: return launch_bt<8>(D, PR, a, stream);
  }
  return FLOWMC_ERR_UNSUPPORTED;
}

}  // namespace flowmc
