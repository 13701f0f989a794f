// MaskedCouplingRQSpline forward / inverse / log_prob / sample / NF proposals on the sm_100a tensor cores.
//
// Reference: src/flowMC/resource/model/nf_model/rqSpline.py:392-504, resource/model/common.py:68-168, and for the
// proposal mode resource/kernel/NF_proposal.py:130-172.
//
// One CTA owns 128 samples (= the 128 TMEM lanes) and walks ALL coupling layers with the tile on chip:
//   * the conditioner's three Linear layers are tcgen05.mma kind::tf32 GEMMs with M = 128:
//       A (activations)  lives in TENSOR MEMORY (written by the epilogue warps with tcgen05.st), so shared-memory
//                        bandwidth only carries B;
//       B (weights)      is streamed from a pre-packed, pre-swizzled global image (tc_pack_flow_kernel) with
//                        cp.async.bulk (TMA engine) through a ring of 32 KB shared-memory stages;
//       D (accumulators) two 128-column TMEM slots, so the MMAs of spline chunk c+1 overlap the epilogue of chunk c.
//     3xTF32: every fp32 operand is split hi + lo and a_hi b_hi + a_lo b_hi + a_hi b_lo is accumulated in fp32, which
//     reproduces the fp32 product to ~2^-21 -- the results stay inside the 1e-5 parity tolerance (tc_terms = 1 runs
//     plain TF32 at 3x the tensor rate with ~1e-3 relative error in the spline parameters).
//   * the rational-quadratic spline is the last GEMM's EPILOGUE: the output columns are ordered feature-major, a
//     chunk of 4 features (100 of 112 columns) is read back with tcgen05.ld (TMEM lane = sample) and softmax /
//     cumsum / softplus / bin search / transform / log-det run in registers (flow_common.cuh), exactly the code the
//     CUDA-core path uses.  Only the transformed half of W3 is ever loaded.
//   * warp roles: warps 0-7 epilogue (TC_PARTS = 2 threads per sample row, splitting columns / features; a warp can
//     only touch its own quarter of the TMEM lanes), warp 8 weight producer, warp 9 MMA issuer (whole warps walk the
//     schedule, one elected lane issues);
//   * hand-offs: per K-chunk of 32 operand columns (a_ready[4], one arrival per epilogue warp), so a GEMM starts on
//     its first chunk while the tanh epilogue is still producing the later ones; persistent CTAs loop over tiles;
//     TRAIN mode also leaves the layer inputs, the spline parameters and packed activation images (through
//     shared-memory staging + bulk stores) for the backward pass; PAIR mode = CTA pairs (cta_group::2), optional.
//   TMEM map (512 columns): [0,128) A hi | [128,256) A lo | [256,384) acc slot 0 | [384,512) acc slot 1.
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>

#include "flow_tc.cuh"
#include "flow_tile.cuh"
#include "flow_train.cuh"
#include "registry.h"

namespace flowmc {

constexpr int TC_STAGES = 4;
constexpr int TC_MAX_ITEMS = 40;

enum : int { TC_FWD = 0, TC_INV = 1, TC_NF = 2, TC_TRAIN = 3 };  // TRAIN = FWD + activation dump for the backward

struct TcItem {
  int kind;      // 0: hidden Linear (+tanh), 1: chunk of spline features
  int K;         // reduction length
  int n_kc;      // ceil(K / 32) stages
  int npad;      // MMA N (multiple of 16, <= 128) = rows per stage image
  int lin;       // hidden: index of the Linear; chunk: first transformed-feature ordinal
  int n_feat;    // chunk: features in it
  uint32_t off;  // byte offset of the item's first stage from the layer's image base
  uint32_t act_off;  // hidden: byte offset of the item's OUTPUT inside a (tile, layer) block of the activation image
};
struct TcProgram {
  int n_items[2];  // by layer parity (which features are transformed alternates, rqSpline.py:434)
  uint32_t layer_bytes[2];
  uint32_t act_layer_bytes;  // tc_act_layer_bytes(D)
  uint32_t pad_;
  TcItem items[2][TC_MAX_ITEMS];
};

bool tc_supported(const FlowmcFlowDesc& D) {
  if (D.n_features < 2 || D.n_features > 128) return false;
  {  // shared memory: weight stages + x tile + every layer's biases must fit one CTA
    const size_t bytes = 4096 + (size_t)4 * 32768 + (size_t)128 * (D.n_features + 1) * 4 +
                         (size_t)D.n_layers * ((D.n_linear - 1) * 128 + ((D.n_features + 1) / 2) * (3 * D.num_bins + 1)) * 4;
    if (bytes > 227 * 1024) return false;
  }
  if (D.num_bins != 4 && D.num_bins != 8 && D.num_bins != 16) return false;
  for (int i = 1; i < D.n_linear; ++i)
    if (D.dims[i] > 128 || (D.dims[i] % 16) != 0) return false;
  return true;
}

static int tc_build_program(const FlowmcFlowDesc& D, TcProgram* P) {
  const int d = D.n_features, NP = 3 * D.num_bins + 1, nh = D.n_linear - 1;
  int fc = 128 / NP;
  fc &= ~1;  // even, so that the two epilogue threads of a row split a chunk evenly
  for (int p = 0; p < 2; ++p) {
    int n = 0;
    uint32_t off = 0;
    for (int i = 0; i < nh; ++i) {
      TcItem& it = P->items[p][n++];
      it.kind = 0; it.K = D.dims[i]; it.n_kc = (it.K + 31) / 32; it.npad = D.dims[i + 1]; it.lin = i; it.n_feat = 0;
      it.off = off; it.act_off = (uint32_t)tc_act_item_off(D, i + 1);
      off += (uint32_t)it.n_kc * 2u * it.npad * 128u;
    }
    const int ntf = (d - p + 1) / 2;
    for (int c0 = 0; c0 < ntf; c0 += fc) {
      if (n >= TC_MAX_ITEMS) return FLOWMC_ERR_UNSUPPORTED;
      TcItem& it = P->items[p][n++];
      it.kind = 1; it.K = D.dims[nh]; it.n_kc = (it.K + 31) / 32; it.lin = c0;
      it.n_feat = (ntf - c0 < fc) ? ntf - c0 : fc;
      it.npad = (it.n_feat * NP + 15) & ~15;
      it.off = off; it.act_off = 0;
      off += (uint32_t)it.n_kc * 2u * it.npad * 128u;
    }
    P->n_items[p] = n;
    P->layer_bytes[p] = off;
  }
  P->act_layer_bytes = (uint32_t)tc_act_layer_bytes(D);
  P->pad_ = 0;
  return FLOWMC_OK;
}

static int64_t tc_image_bytes(const FlowmcFlowDesc& D, const TcProgram& P) {
  int64_t b = 0;
  for (int l = 0; l < D.n_layers; ++l) b += P.layer_bytes[l & 1];
  return b;
}

__device__ __forceinline__ int64_t tc_layer_base(const FlowmcFlowDesc& D, const TcProgram& P, int l) {
  // layers alternate parity 0,1,0,1,...
  return (int64_t)((l + 1) / 2) * P.layer_bytes[0] + (int64_t)(l / 2) * P.layer_bytes[1];
}

// blob -> packed image.  grid = (items, layers, element slices)
__global__ void tc_pack_flow_kernel(const FlowmcFlowDesc D, const TcProgram P, const float* __restrict__ params,
                                    uint8_t* __restrict__ image) {
  const int l = blockIdx.y, p = l & 1;
  if ((int)blockIdx.x >= P.n_items[p]) return;
  const TcItem it = P.items[p][blockIdx.x];
  const int NP = 3 * D.num_bins + 1, nh = D.n_linear - 1;
  const float* PL = params + (int64_t)l * D.layer_stride;
  float* dst = reinterpret_cast<float*>(image + tc_layer_base(D, P, l) + it.off);
  const int per_stage = it.npad * 32;
  const int total = it.n_kc * per_stage;
  for (int i = blockIdx.z * blockDim.x + threadIdx.x; i < total; i += blockDim.x * gridDim.z) {
    const int kc = i / per_stage, rem = i - kc * per_stage;
    const int n = rem >> 5, kk = rem & 31;
    const int k = kc * 32 + kk;
    float w = 0.0f;
    if (k < it.K) {
      if (it.kind == 0) {
        if (n < D.dims[it.lin + 1]) w = PL[D.off_W[it.lin] + (int64_t)n * it.K + k];
      } else if (n < it.n_feat * NP) {
        const int fi = n / NP, r = n - fi * NP;
        const int f = p + 2 * (it.lin + fi);
        w = PL[D.off_W[nh] + ((int64_t)f * NP + r) * it.K + k];
      }
    }
    uint32_t hi, lo;
    tc::split_tf32_rn(w, hi, lo);
    float* stage = dst + (int64_t)kc * 2 * per_stage;
    const int o = tc::packed_b_offset(n, kk) >> 2;
    stage[o] = __uint_as_float(hi);
    stage[per_stage + o] = __uint_as_float(lo);
  }
}

struct TcArgs {
  const float* params;
  const uint8_t* image;
  const float* xin;       // [n, d] (FWD / INV with PRE_NONE / PRE_WHITEN)
  const int32_t* idx;     // optional row gather for xin
  float* yout;            // [n, d] transformed rows (POST_NONE / POST_UNWHITEN; NF: the proposals)
  float* ldout;           // [n] log-det, or log_prob with POST_BASE_LOGP (NF: flow log-prob of the proposals)
  int64_t n;
  int pre, post, terms;
  const uint32_t* keys;   // PRE_NORMAL: device keys [ceil(n / rows_per_key), 2] or NULL -> host_key
  Key host_key;
  int64_t rows_per_key;
  // NF mode: per-row key schedule of NFProposal.sample_flow
  Key subkey;
  const uint32_t* chain_keys;
  int64_t chain_offset;
  int n_steps, n_batch, n_sample;
  // training forward (MODE FWD): optional activation dump for the backward pass (flow_train.cu)
  float* save_x;      // [L + 1][n][d]            every layer's input, then the final latent
  float* save_h;      // [L][n_hidden][128][n]    hidden activations, column-major over samples
  float* save_theta;  // [L][ceil(d/2) * NP][n]   spline parameters (bias included) of the transformed features
  uint8_t* act_img;   // per (tile, layer): conditioner input and hidden activations as packed B stages (K = the
                      // tile's 128 rows) for the tensor-core weight-gradient GEMMs (flow_tc.cuh: tc_act_*)
  long long* timing;  // optional diagnostics: [3][256] clock64 stamps of CTA 0 (producer, MMA issuer, epilogue thread 0)
  int dump_vec;       // training forward: activation images leave with st.global.v4 (1) or bulk stores (0)
  int dbg_skip;       // TIMING EXPERIMENTS ONLY (FLOWMC_TC_DBG_SKIP): 1 = no activation-image dump, 2 = no theta dump,
                      // 4 = no layer-input save -- the results are then useless to the backward pass
};

#define TC_STAMP(role)                                                     \
  do {                                                                     \
    if (a.timing != nullptr && blockIdx.x == 0 && n_stamp < 256) a.timing[(role) * 256 + n_stamp++] = clock64(); \
  } while (0)

struct TcSmem {
  uint64_t stage_full[2 * TC_STAGES], stage_empty[2 * TC_STAGES], acc_full[2], acc_empty[2];
  uint64_t a_ready[4];  // per K-chunk (32 columns) of the A operand: a GEMM starts on the first chunk while the
                        // epilogue threads are still writing the later ones
  uint64_t peer_full[2 * TC_STAGES];  // CTA pair, leader only: the peer CTA's half of a weight stage has landed
  uint64_t xbar[2];  // feature split: [0] "my tile may be written" (every epilogue warp of every CTA of the cluster
                     // arrives once per layer), [1] "all transformed features / log-det partials have been exchanged"
  uint32_t tmem_base;
  float ldpart[TC_PARTS][TC_M];
  alignas(16) float ldx[8][TC_M];  // feature split: per-CTA log-det partial sums of the tile (summed on rank 0)
};


// SPLIT (training forward only): a cluster of R = 2, 4 or 8 CTAs shares ONE tile.  Every CTA keeps the whole tile in
// its shared memory and runs the two tanh layers redundantly, but takes only every R-th chunk of spline features
// (chunk index % R == cluster rank): the W3 GEMMs and the spline epilogues -- 3/4 of a layer's MMAs, nearly all of its
// MUFU work -- divide by R.  After a layer each CTA has written its transformed features into every peer's tile
// through distributed shared memory (st.shared::cluster) and one mbarrier round synchronises the cluster; log-det
// partial sums are reduced on rank 0 in rank order (deterministic).  This is what lets a data-parallel rank that
// holds only a few tiles (batch / n_gpus rows) use all of its SMs: per-tile latency, not throughput, bounds a
// training step.
template <int KB, int MODE, bool PAIR, bool SPLIT = false>
__global__ void __launch_bounds__(TC_THREADS, 1) flow_tc_kernel(const FlowmcFlowDesc D, const TcProgram PR,
                                                                const TcArgs a) {
  constexpr int NP = 3 * KB + 1;
  static_assert(!(PAIR && SPLIT), "CTA pairs and the feature split are alternative cluster modes");
  static_assert(!SPLIT || MODE == TC_TRAIN, "the feature split is built for the training forward pass");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stages = smem;                                             // the weight ring, 1024-aligned slots
  // A pair streams half-size stages, so the same 128 KB hold twice as many: 8 slots = two spline chunks of
  // prefetch.  (The ring is latency-bound, not bandwidth-bound: a slot comes back only after MMA completion ->
  // commit -> producer -> L2 round trip, ~3.5K cycles; 4 slots give one chunk per round trip.)
  // Training gives the last 32 KB of the ring region to the activation-image staging buffers (4 KB per epilogue
  // warp): the weight stream is rate-bound, not depth-bound (a 5th slot changed nothing), 3 slots keep it fed.
  constexpr int NST = (PAIR ? 2 : 1) * (MODE == TC_TRAIN ? TC_STAGES - 1 : TC_STAGES);
  constexpr int SLOT = PAIR ? TC_STAGE_BYTES / 2 : TC_STAGE_BYTES;
  TcSmem* S = reinterpret_cast<TcSmem*>(smem + TC_STAGES * TC_STAGE_BYTES);
  const int d = D.n_features;
  const int xs_stride = d + 1;
  float* xs = reinterpret_cast<float*>(smem + TC_STAGES * TC_STAGE_BYTES + ((sizeof(TcSmem) + 15) & ~15));
  // every layer's biases, staged once: per layer [hidden Linears: 128 each][spline-parameter biases of the layer's
  // transformed features, feature-ordinal major]
  float* sbias_all = xs + TC_M * xs_stride;
  const int bias_stride = (D.n_linear - 1) * 128 + ((d + 1) / 2) * NP;
  // feature split: exchange buffer [transformed-feature ordinal][row] (16-byte aligned), after the biases
  float* xstage = sbias_all + ((D.n_layers * bias_stride + 3) & ~3);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* astage_all = stages + (size_t)(TC_STAGES - 1) * TC_STAGE_BYTES;  // training only, see NST
  const float* P = a.params;
  const int L = D.n_layers, nh = D.n_linear - 1;
  const int n_pass = (MODE == TC_NF) ? 2 : 1;
  // persistent CTAs: tile = blockIdx.x, blockIdx.x + gridDim.x, ... (barrier phases, the weight ring and the staged
  // biases carry over from tile to tile; the weight stream of the next tile runs under the tail of this one)
  const int64_t n_tiles_real = (a.n + TC_M - 1) / TC_M;
  const int64_t n_tiles = PAIR ? ((n_tiles_real + 1) & ~(int64_t)1) : n_tiles_real;
  // feature split: cluster c takes tiles c, c + n_clusters, ...; rank r of the cluster takes the chunks with index % R == r
  const uint32_t R = SPLIT ? tc::cluster_size() : 1u;
  const int64_t tile_first = SPLIT ? (int64_t)(blockIdx.x / R) : (int64_t)blockIdx.x;
  const int64_t tile_step = SPLIT ? (int64_t)(gridDim.x / R) : (int64_t)gridDim.x;
  // PAIR: two CTAs on neighbouring SMs (a 2-CTA cluster) run their two tiles through the same schedule as ONE
  // M = 256 problem (tcgen05.mma.cta_group::2): each CTA keeps its own rows of A and D in its tensor memory and
  // streams only HALF of every weight stage (N / 2 rows of the hi and lo images) into its shared memory -- the
  // weight stream, which paces the spline-parameter GEMMs at the chip's L2 read rate, halves per SM.  The leader
  // (cluster rank 0) issues every MMA; the peer's MMA warp only relays "my half of stage s has landed"; epilogue
  // threads of both CTAs signal the leader's operand / accumulator barriers; MMA completions are committed to the
  // barriers of both CTAs.
  const uint32_t crank = (PAIR || SPLIT) ? tc::cluster_rank() : 0u;
  auto skip_item = [&](const TcItem& it, int ii) -> bool {  // a spline chunk that another CTA of the cluster owns
    return SPLIT && it.kind == 1 && (uint32_t)(ii - (D.n_linear - 1)) % R != crank;
  };
  constexpr uint32_t kEpiArrivals = (PAIR ? 2 : 1) * TC_EPI_WARPS;  // one arrival per epilogue warp
  if (warp == TC_EPI_WARPS + 1 && lane == 0) {
    for (int i = 0; i < NST; ++i) {
      tc::mbar_init(&S->stage_full[i], 1);
      tc::mbar_init(&S->stage_empty[i], 1);
      tc::mbar_init(&S->peer_full[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&S->acc_full[i], 1);
      tc::mbar_init(&S->acc_empty[i], kEpiArrivals);
    }
    for (int i = 0; i < 4; ++i) tc::mbar_init(&S->a_ready[i], kEpiArrivals);
    tc::mbar_init(&S->xbar[0], (SPLIT ? R : 1u) * TC_EPI_WARPS);  // control: one arrival per epilogue warp of the cluster
    tc::mbar_init(&S->xbar[1], 1);  // data: one local arrive.expect_tx per round + the peers' st.async bytes
    tc::fence_mbar_init();
  }
  if (warp == TC_EPI_WARPS) {
    if (PAIR) tc::tmem_alloc_pair<512>(&S->tmem_base);
    else tc::tmem_alloc<512>(&S->tmem_base);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (PAIR || SPLIT) tc::cluster_sync();  // every CTA's barriers and tensor memory exist before anything crosses over
  tc::tc_fence_after();
  // epilogue -> MMA warp signal, called by whole warps (every lane has fenced its tensor-memory accesses): one
  // arrival per warp -- in a pair the barrier lives in the leader CTA, and 256 remote arrivals per hand-off cost
  // more than the GEMM they release
  auto arrive_mma = [&](uint64_t* bar) {
    __syncwarp();
    if (lane == 0) {
      if (!PAIR || crank == 0) tc::mbar_arrive(bar);
      else tc::mbar_arrive_remote(bar, 0);
    }
  };
  const uint32_t tbase = S->tmem_base;
  const uint32_t t_ahi = tbase, t_alo = tbase + 128;

  if (warp == TC_EPI_WARPS) {
    // ===== weight producer (whole warp walks the schedule; one elected lane issues) ==============
    {
      uint32_t s = 0, ph = 0;
      int n_stamp = 0;
      for (int64_t tile = tile_first; tile < n_tiles; tile += tile_step)
      for (int pass = 0; pass < n_pass; ++pass) {
        const bool inv = (MODE == TC_INV) || (MODE == TC_NF && pass == 0);
        for (int li = 0; li < L; ++li) {
          const int l = inv ? L - 1 - li : li, p = l & 1;
          const uint8_t* lbase = a.image + tc_layer_base(D, PR, l);
          for (int ii = 0; ii < PR.n_items[p]; ++ii) {
            const TcItem& it = PR.items[p][ii];
            if (skip_item(it, ii)) continue;
            const uint32_t bytes = 2u * it.npad * 128u;
            for (int kc = 0; kc < it.n_kc; ++kc) {
              tc::mbar_wait(&S->stage_empty[s], ph ^ 1);
              if (tc::elect_one()) {
                TC_STAMP(0);
                uint8_t* dst = stages + (size_t)s * SLOT;
                const uint8_t* src = lbase + it.off + (size_t)kc * bytes;
                if (!PAIR) {
                  tc::mbar_arrive_expect_tx(&S->stage_full[s], bytes);
                  tc::bulk_g2s(dst, src, bytes, &S->stage_full[s]);
                } else {
                  // this CTA's rows crank * npad / 2 .. of the hi image, then of the lo image
                  const uint32_t half = (uint32_t)it.npad * 64u;
                  tc::mbar_arrive_expect_tx(&S->stage_full[s], 2u * half);
                  tc::bulk_g2s(dst, src + (size_t)crank * half, half, &S->stage_full[s]);
                  tc::bulk_g2s(dst + half, src + (size_t)it.npad * 128u + (size_t)crank * half, half, &S->stage_full[s]);
                }
              }
              __syncwarp();
              if (++s == NST) { s = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == TC_EPI_WARPS + 1) {
    // ===== MMA issuer (whole warp walks the schedule; one elected lane issues) ====================
    if (PAIR && crank != 0) {
      // peer CTA of a pair: no MMAs to issue; relay the arrival of this CTA's half of every weight stage
      uint32_t s = 0, ph = 0;
      for (int64_t tile = tile_first; tile < n_tiles; tile += tile_step)
      for (int pass = 0; pass < n_pass; ++pass)
        for (int li = 0; li < L; ++li) {
          const int p = (((MODE == TC_INV) || (MODE == TC_NF && pass == 0)) ? L - 1 - li : li) & 1;
          for (int ii = 0; ii < PR.n_items[p]; ++ii)
            for (int kc = 0; kc < PR.items[p][ii].n_kc; ++kc) {
              tc::mbar_wait(&S->stage_full[s], ph);
              if (tc::elect_one()) tc::mbar_arrive_remote(&S->peer_full[s], 0);
              __syncwarp();
              if (++s == NST) { s = 0; ph ^= 1; }
            }
        }
    } else {
      uint32_t s = 0, ph = 0, seq = 0, a_ph = 0;
      int n_stamp = 0;
      for (int64_t tile = tile_first; tile < n_tiles; tile += tile_step)
      for (int pass = 0; pass < n_pass; ++pass) {
        const bool inv = (MODE == TC_INV) || (MODE == TC_NF && pass == 0);
        for (int li = 0; li < L; ++li) {
          const int l = inv ? L - 1 - li : li, p = l & 1;
          bool first_chunk = true;  // the first spline chunk THIS CTA runs in the layer picks up the h_last operand
          for (int ii = 0; ii < PR.n_items[p]; ++ii) {
            const TcItem it = PR.items[p][ii];
            if (skip_item(it, ii)) continue;
            const bool new_a = it.kind == 0 || first_chunk;  // a new A operand: masked x, h1, ..., h_last
            if (it.kind == 1) first_chunk = false;
            const uint32_t slot = seq & 1;
            tc::mbar_wait(&S->acc_empty[slot], ((seq >> 1) & 1) ^ 1);
            tc::tc_fence_after();
            if (lane == 0) TC_STAMP(1);  // operands + accumulator slot available
            const uint32_t t_acc = tbase + 256 + slot * 128;
            const uint32_t idesc = tc::make_idesc_tf32(PAIR ? 2 * TC_M : TC_M, it.npad);
            const bool three = a.terms == 3;
            for (int kc = 0; kc < it.n_kc; ++kc) {
              if (new_a) {  // K-chunk kc of the operand has been written
                tc::mbar_wait(&S->a_ready[kc], (a_ph >> kc) & 1);
                a_ph ^= 1u << kc;
              }
              tc::mbar_wait(&S->stage_full[s], ph);
              if (PAIR) tc::mbar_wait(&S->peer_full[s], ph);
              tc::tc_fence_after();
              if (kc == 0 && lane == 0) TC_STAMP(1);  // first weight stage landed
              const uint32_t b_hi = tc::smem_u32(stages + (size_t)s * SLOT);
              const uint64_t dhi = tc::make_b_desc(b_hi), dlo = tc::make_b_desc(b_hi + it.npad * (PAIR ? 64 : 128));
              const int ksteps = min(4, (it.K - kc * 32 + 7) >> 3);
              const uint32_t acol = kc * 32;
              if (tc::elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  if (ks < ksteps) {
                    // +32 bytes per k-step inside the 128-byte swizzle atom = +2 in the descriptor's address field
                    if (!PAIR) {
                      tc::mma_tf32_ts(t_acc, t_ahi + acol + ks * 8, dhi + 2 * ks, idesc, (kc | ks) != 0);
                      if (three) {
                        tc::mma_tf32_ts(t_acc, t_alo + acol + ks * 8, dhi + 2 * ks, idesc, 1);
                        tc::mma_tf32_ts(t_acc, t_ahi + acol + ks * 8, dlo + 2 * ks, idesc, 1);
                      }
                    } else {
                      tc::mma_tf32_ts_pair(t_acc, t_ahi + acol + ks * 8, dhi + 2 * ks, idesc, (kc | ks) != 0);
                      if (three) {
                        tc::mma_tf32_ts_pair(t_acc, t_alo + acol + ks * 8, dhi + 2 * ks, idesc, 1);
                        tc::mma_tf32_ts_pair(t_acc, t_ahi + acol + ks * 8, dlo + 2 * ks, idesc, 1);
                      }
                    }
                  }
                }
                if (!PAIR) tc::mma_commit(&S->stage_empty[s]);
                else tc::mma_commit_pair(&S->stage_empty[s], 3);
              }
              __syncwarp();
              if (++s == NST) { s = 0; ph ^= 1; }
            }
            if (tc::elect_one()) {
              if (!PAIR) tc::mma_commit(&S->acc_full[slot]);
              else tc::mma_commit_pair(&S->acc_full[slot], 3);
            }
            __syncwarp();
            if (lane == 0) TC_STAMP(1);  // all MMAs of the item issued
            ++seq;
          }
        }
      }
    }
  } else {
    // ===== epilogue warps: two threads per sample row ==========================================
    const int q = warp & 3, hf = warp >> 2;  // hf: which of the row's TC_PARTS threads this is
    auto part = [&](int n, int& lo, int& hi) {  // this thread's share [lo, hi) of n work items
      const int per = (n + TC_PARTS - 1) / TC_PARTS;
      lo = min(n, hf * per);
      hi = min(n, lo + per);
    };
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    int64_t row0 = 0, grow = 0, r = 0, tile = 0;  // set per tile below
    float* xr = xs + row * xs_stride;
    // Activation image rows of NB (8 or 16) operand columns n0 .. n0 + NB - 1 (NB image rows of 128 B: this warp's
    // 32 samples) -> global.  The 32 lanes write their words into the warp's shared-memory buffer in image order and
    // one lane sends the hi and the lo block with two bulk stores (asynchronous, full lines, off the LSU store
    // path, which the scattered 4-byte stores of a direct dump saturate).
    uint8_t* astage = astage_all + warp * 4096;
    const uint32_t dump_sw = (uint32_t)(((lane >> 2) << 4) | ((lane & 3) << 2));  // see dump_rows
    // a.dump_vec (FLOWMC_TC_DUMP=vec, the default): the staged block leaves with coalesced 16-byte st.global from
    // all 32 lanes instead -- the bulk stores queue behind the weight stream in the SM's TMA unit, and waiting for
    // the previous one to release the buffer (bulk_wait_read) was what tripled the tanh epilogues of the training
    // forward (10.5K against 3.4K cycles, profiles/r02 timelines).
    auto dump_rows = [&](auto nb_tag, const uint32_t* hi, const uint32_t* lo, int n0, uint8_t* gimg, int n_rows) {
      constexpr int NB = decltype(nb_tag)::value;
      if (a.dbg_skip & 1) return;
      if (!a.dump_vec && lane == 0) tc::bulk_wait_read<0>();  // the previous rows have left the buffer
      __syncwarp();
      // tc::packed_b_offset(u, lane) = (u >> 3) * 1024 + (u & 7) * 128 + (dump_sw ^ ((u & 7) << 4)): the lane part is
      // hoisted (dump_sw), what is left per element is one XOR with a constant and an immediate offset
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        uint8_t* o = astage + ((u >> 3) * 1024 + (u & 7) * 128) + (dump_sw ^ (uint32_t)((u & 7) << 4));
        *reinterpret_cast<uint32_t*>(o) = hi[u];
        *reinterpret_cast<uint32_t*>(o + NB * 128) = lo[u];
      }
      if (a.dump_vec) {
        __syncwarp();
        const float4* s4 = reinterpret_cast<const float4*>(astage) + lane;
        float4* ghi = reinterpret_cast<float4*>(gimg + (size_t)n0 * 128) + lane;
        float4* glo = ghi + (size_t)n_rows * 8;
#pragma unroll
        for (int v = 0; v < NB / 4; ++v) {  // NB * 128 bytes per block = NB * 8 float4 = NB / 4 per lane
          __stcs(ghi + v * 32, s4[v * 32]);
          __stcs(glo + v * 32, s4[NB * 8 + v * 32]);
        }
        return;
      }
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tc::bulk_s2g(gimg + (size_t)n0 * 128, astage, NB * 128);
        tc::bulk_s2g(gimg + (size_t)(n_rows + n0) * 128, astage + NB * 128, NB * 128);
        tc::bulk_commit();
      }
    };
    // this thread's share of the d features (affine / operand writes / pre / post): 8-column groups
    const int g_all = (d + 7) >> 3;
    int g_lo, g_hi;
    part(g_all, g_lo, g_hi);
    const int j_lo = g_lo * 8, j_hi = min(d, g_hi * 8);
    // feature split: remote addresses of this thread's tile row in the peer CTAs, phases of the two cluster barriers
    uint32_t x_ph0 = 0, x_ph1 = 0;
    // which transformed-feature ordinals (feature f = parity + 2 * ordinal) this CTA owns: bit o of the mask
    // (one mask serves both layer parities: chunk c = ordinals [c * fc, (c + 1) * fc) belongs to rank c % R)
    uint64_t own_mask = ~0ull;
    if (SPLIT) {
      const int fc_split = PR.items[0][D.n_linear - 1].n_feat;  // features per (full) spline chunk
      own_mask = 0;
      for (int o = 0; o < (d + 1) / 2; ++o)
        if ((uint32_t)(o / fc_split) % R == crank) own_mask |= 1ull << o;
    }
    // does this CTA transform feature j in layer l?  (j is one of the layer's transformed features)
    auto owns_feature = [&](int j, int l) -> bool { return !SPLIT || ((own_mask >> ((j - (l & 1)) >> 1)) & 1ull) != 0; };
    // Data exchange round: the peers' values arrive as bulk copies (shared memory -> peer shared memory) that complete
    // bytes on THIS CTA's xbar[1]; one thread announces how many bytes the round brings (arrive.expect_tx), everyone
    // waits for the phase.  No release fences: the sender's global stores (activation dumps) are not drained, and
    // no per-element remote stores (4-byte st.async packets cost ~16K cycles per layer, profiles/r02 notes).
    // send n_bytes at float offset `off` of xstage / of S->ldx to the same place in every peer (tid 0 only)
    auto send_block = [&](const float* src, uint32_t n_bytes, bool to_rank0_only) {
      for (uint32_t rk = 0; rk < (to_rank0_only ? 1u : R); ++rk)
        if (rk != crank) tc::bulk_s2peer(tc::mapa_u32(src, rk), src, n_bytes, tc::mapa_u32(&S->xbar[1], rk));
    };
    auto expect_round = [&](uint32_t bytes) {
      if (tid == 0) tc::mbar_arrive_expect_tx(&S->xbar[1], bytes);
    };
    auto wait_round = [&]() {
      tc::mbar_wait(&S->xbar[1], x_ph1);
      x_ph1 ^= 1;
    };

    // stage every layer's biases (once per CTA; visible after the epi_bar that follows).  Flattened over (layer, entry) with 8
    // independent loads in flight per thread: the naive per-layer loop cost ~16K cycles of serialised L2 latency.
    {
      const int total = L * bias_stride;
      for (int base = 0; base < total; base += 8 * TC_EPI) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = base + u * TC_EPI + tid;
          v[u] = 0.0f;
          if (e < total) {
            const int l = e / bias_stride, c = e - l * bias_stride;
            const float* PL = P + (int64_t)l * D.layer_stride;
            if (c < nh * 128) {
              const int i = c >> 7, cc = c & 127;
              if (cc < D.dims[i + 1]) v[u] = PL[D.off_b[i] + cc];
            } else {
              const int c2 = c - nh * 128, o = c2 / NP, rr = c2 - o * NP, p = l & 1;
              if (o < (d - p + 1) / 2) v[u] = PL[D.off_b[nh] + (p + 2 * o) * NP + rr];
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = base + u * TC_EPI + tid;
          if (e < total) sbias_all[e] = v[u];
        }
      }
    }
    epi_bar();
    uint32_t seq = 0;
    int n_stamp = (tid == 0) ? 0 : 256;
    for (tile = tile_first; tile < n_tiles; tile += tile_step) {
    row0 = tile * TC_M;
    grow = row0 + row;
    r = min(grow, a.n - 1);
    // training: this tile's block of the activation image (NULL: no image wanted / idle tile)
    uint8_t* const act_tile = (MODE == TC_TRAIN && a.act_img != nullptr && row0 < a.n)
                                  ? a.act_img + (size_t)tile * L * PR.act_layer_bytes : nullptr;
    // ---- load / generate the tile ------------------------------------------------------------
    if (MODE == TC_NF) {
      const int64_t c = r / a.n_steps;
      const int t = (int)(r - c * a.n_steps);
      Key ck = a.chain_keys ? Key{a.chain_keys[2 * c], a.chain_keys[2 * c + 1]}
                            : split_at(a.subkey, (uint64_t)(a.chain_offset + c));
      Key key = split_at(ck, 1);
      int idx = t;
      if (a.n_batch > 0) {
        const int b = t / a.n_sample;
        for (int i = 0; i < b; ++i) key = split_at(key, 0);
        key = split_at(key, 1);
        idx = t - b * a.n_sample;
      }
      for (int j = j_lo; j < j_hi; ++j) {
        const float z = bits_to_normal(bits_at(key, (uint64_t)((uint32_t)idx * (uint32_t)d + (uint32_t)j)));
        xr[j] = P[D.off_base_mean + j] + z * sqrtf(P[D.off_base_cov + (int64_t)j * d + j]);
      }
    } else if (a.pre == PRE_NORMAL) {
      const int64_t kidx = r / a.rows_per_key;
      const Key key = a.keys ? Key{a.keys[2 * kidx], a.keys[2 * kidx + 1]} : a.host_key;
      for (int j = j_lo; j < j_hi; ++j) {
        const float z = bits_to_normal(bits_at(key, (uint64_t)((r - kidx * a.rows_per_key) * d + j)));
        xr[j] = P[D.off_base_mean + j] + z * sqrtf(P[D.off_base_cov + (int64_t)j * d + j]);
      }
    } else {
      const float* src = a.xin + (a.idx ? (int64_t)a.idx[r] : r) * d;
      for (int j = j_lo; j < j_hi; ++j) {
        float v = src[j];
        if (a.pre == PRE_WHITEN) v = (v - P[D.off_data_mean + j]) / sqrtf(P[D.off_data_cov + (int64_t)j * d + j]);
        xr[j] = v;
      }
    }
    // the tile (all rows, complete after an epi_bar) -> save_x[slot]: rows are contiguous in global memory, so the
    // CTA writes them as one coalesced stream
    const int st_q = TC_EPI / d, st_r = TC_EPI - st_q * d;  // TC_EPI = st_q * d + st_r
    auto save_tile = [&](int slot) {
      if (a.dbg_skip & 4) return;
      const int64_t rows = min((int64_t)TC_M, a.n - row0);
      float* dst = a.save_x + ((int64_t)slot * a.n + row0) * d;
      // feature split: every CTA holds the whole tile; rank r writes its share of the rows
      const int e_lo = SPLIT ? (int)(crank * (TC_M / R)) * d : 0;
      const int e_hi = SPLIT ? min((int)rows, (int)((crank + 1) * (TC_M / R))) * d : (int)rows * d;
      // flat (coalesced) index e = rr * d + cc, advanced by TC_EPI per step without a division
      int e = e_lo + tid;
      int rr = e / d, cc = e - rr * d;
      for (; e < e_hi; e += TC_EPI) {
        dst[e] = xs[rr * xs_stride + cc];
        rr += st_q;
        cc += st_r;
        if (cc >= d) {
          cc -= d;
          ++rr;
        }
      }
    };
    float ldacc = 0.0f;
    TC_STAMP(2);  // tile loaded
    if (MODE == TC_TRAIN) epi_bar();  // save_tile(0) reads every thread's part of the tile

    for (int pass = 0; pass < n_pass; ++pass) {
      const bool inv = (MODE == TC_INV) || (MODE == TC_NF && pass == 0);
      for (int li = 0; li < L; ++li) {
        const int l = inv ? L - 1 - li : li, p = l & 1;
        const float* PL = P + (int64_t)l * D.layer_stride;
        const float scale = PL[D.off_scale], shift = PL[D.off_shift];
        const float* sbias = sbias_all + l * bias_stride;
        if (MODE == TC_TRAIN && a.save_x != nullptr) {
          save_tile(l);
          epi_bar();  // the affine below rewrites the tile
          if (SPLIT) TC_STAMP(2);  // (split) layer input saved
        }
        // ---- ScalarAffine (rqSpline.py:435-436) + A operand = x * mask, hi / lo ------------------
        {
          const float e = inv ? expf(-scale) : expf(scale);
          for (int g = g_lo; g < g_hi; ++g) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int j = g * 8 + u;
              float v = 0.0f;
              if (j < d) {
                const bool transformed = ((j + l) & 1) == 0;
                // feature split: a transformed feature another CTA owns stays untouched here -- the owner's result
                // will overwrite it (its old value is needed by nobody in this CTA)
                if (!(SPLIT && transformed && !owns_feature(j, l))) {
                  v = xr[j];
                  v = inv ? v * e - shift : (v + shift) * e;
                  xr[j] = v;
                }
                if (transformed) v = 0.0f;  // transformed features do not feed the conditioner
              }
              tc::split_tf32(v, hi[u], lo[u]);
            }
            tc::tmem_st8(t_ahi + lane_base + g * 8, hi);
            tc::tmem_st8(t_alo + lane_base + g * 8, lo);
            if (MODE == TC_TRAIN && act_tile != nullptr && (!SPLIT || (uint32_t)g % R == crank)) {
              const int npx = tc_pad16(d);
              dump_rows(std::integral_constant<int, 8>{}, hi, lo, g * 8,
                        act_tile + (size_t)l * PR.act_layer_bytes + (size_t)q * 2 * npx * 128, npx);
            }
          }
          if (hf == 0 && crank == 0) ldacc += inv ? -(float)d * scale : (float)d * scale;
          if (SPLIT) TC_STAMP(2);  // (split) affine loop done
          tc::tmem_wait_st();
          tc::tc_fence_before();
          for (int kc = 0; kc < PR.items[p][0].n_kc; ++kc) arrive_mma(&S->a_ready[kc]);
          if (SPLIT) {
            // announce this layer's exchange round BEFORE any peer can be released into sending: the layer brings
            // (transformed features of the layer - those this CTA owns) x 128 rows x 4 bytes
            int owned = 0;
            for (int ii = nh; ii < PR.n_items[p]; ++ii)
              if (!skip_item(PR.items[p][ii], ii)) owned += PR.items[p][ii].n_feat;
            expect_round((uint32_t)(((d - p + 1) / 2 - owned) * TC_M * 4));
          }
          epi_bar();  // the row's other thread reads these x values in the spline stage
          TC_STAMP(2);  // affine + operand written
          if (SPLIT && lane == 0) {
            // this warp no longer reads the layer's input (save_tile and the affine are done; the bar.sync above
            // ordered every thread's shared-memory accesses): the peers may write their transformed features into
            // this CTA's tile
            for (uint32_t rk = 0; rk < R; ++rk) tc::mbar_arrive_remote_relaxed(&S->xbar[0], rk);
          }
        }
        bool peers_ready = !SPLIT;
        for (int ii = 0; ii < PR.n_items[p]; ++ii) {
          const TcItem& it = PR.items[p][ii];
          if (skip_item(it, ii)) continue;
          const uint32_t slot = seq & 1;
          const uint32_t t_acc = tbase + 256 + slot * 128 + lane_base;
          tc::mbar_wait(&S->acc_full[slot], (seq >> 1) & 1);
          tc::tc_fence_after();
          TC_STAMP(2);  // accumulator observed full
          if (it.kind == 0) {
            // ---- tanh(acc + b) -> next A operand (this thread: half of the columns) ----------------
            // Columns in 16-wide groups, interleaved between the row's two threads (group 2 j + hf), so that K-chunk j
            // (32 columns) of the next GEMM's operand is complete after step j and its MMAs start while the later
            // columns are still being computed.
            static_assert(TC_PARTS == 2, "column interleave assumes two threads per row");
            const int N = it.npad;               // hidden width, multiple of 16
            const float* bias = sbias + it.lin * 128;
            for (int j = 0; j < (N + 31) / 32; ++j) {
              const int c = (2 * j + hf) * 16;
              if (c < N) {
                float v[16];
                tc::tmem_ld16(t_acc + c, v);
                tc::tmem_wait_ld();
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                  v[u] = tanh_ex2(v[u] + bias[c + u]);
                  tc::split_tf32(v[u], hi[u], lo[u]);
                }
                tc::tmem_st8(t_ahi + lane_base + c, hi);
                tc::tmem_st8(t_ahi + lane_base + c + 8, hi + 8);
                tc::tmem_st8(t_alo + lane_base + c, lo);
                tc::tmem_st8(t_alo + lane_base + c + 8, lo + 8);
                tc::tmem_wait_st();
                tc::tc_fence_before();
                arrive_mma(&S->a_ready[j]);  // the next GEMM starts on this K-chunk
                if (MODE == TC_TRAIN && a.save_h != nullptr && grow < a.n) {  // (CUDA-core backward only)
                  float* sh = a.save_h + ((int64_t)(l * nh + it.lin) * 128 + c) * a.n + grow;
#pragma unroll
                  for (int u = 0; u < 16; ++u) sh[(int64_t)u * a.n] = v[u];
                }
              } else {
                tc::tc_fence_before();
                arrive_mma(&S->a_ready[j]);
              }
            }
            arrive_mma(&S->acc_empty[slot]);
            if (MODE == TC_TRAIN && act_tile != nullptr) {
              // Activation image of this layer's output, AFTER every K-chunk has been handed over: the words are read
              // back from the operand region (the next GEMM only reads it; this thread rewrites it no earlier than its
              // next tanh epilogue / the next layer's affine, in program order), so the 128 KB of stores per tile run
              // under the next GEMM's MMAs instead of in front of them.
              uint8_t* gimg = act_tile + (size_t)l * PR.act_layer_bytes + it.act_off + (size_t)q * 2 * N * 128;
              for (int j = 0; j < (N + 31) / 32; ++j) {
                const int c = (2 * j + hf) * 16;
                if (c < N && (!SPLIT || (uint32_t)(c >> 4) % R == crank)) {  // split: the redundant copies share the dump
                  float fh[16], fl[16];
                  tc::tmem_ld16(t_ahi + lane_base + c, fh);
                  tc::tmem_ld16(t_alo + lane_base + c, fl);
                  tc::tmem_wait_ld();
                  uint32_t hi[16], lo[16];
#pragma unroll
                  for (int u = 0; u < 16; ++u) {
                    hi[u] = __float_as_uint(fh[u]);
                    lo[u] = __float_as_uint(fl[u]);
                  }
                  dump_rows(std::integral_constant<int, 16>{}, hi, lo, c, gimg, N);
                }
              }
            }
          } else {
            // ---- spline epilogue: this thread's half of the chunk's features ------------------------
            const int nf = it.n_feat;
            int i_lo, i_hi;
            part(nf, i_lo, i_hi);
            const float* bl = sbias + nh * 128 + it.lin * NP;  // ordinal-major: feature i of the chunk at i * NP
            auto load_raw = [&](int i, float* raw) {
              float v[32];
              tc::tmem_ld32(t_acc + i * NP, v);
              if (NP > 32) {
                float v2[32];
                tc::tmem_ld32(t_acc + i * NP + 32, v2);
                tc::tmem_wait_ld();
#pragma unroll
                for (int u = 32; u < NP; ++u) raw[u] = v2[u - 32] + bl[i * NP + u];
              } else {
                tc::tmem_wait_ld();
              }
#pragma unroll
              for (int u = 0; u < NP && u < 32; ++u) raw[u] = v[u] + bl[i * NP + u];
              if (MODE == TC_TRAIN && a.save_theta != nullptr && grow < a.n && !(a.dbg_skip & 2)) {
                // feature block [NP][n]: warp-uniform 64-bit base + 32-bit element offsets (tc path: n < 2^25 rows)
                float* dst = a.save_theta + ((int64_t)l * ((d + 1) / 2) + it.lin + i) * NP * a.n;
                const uint32_t o0 = (uint32_t)grow, nn = (uint32_t)a.n;
#pragma unroll
                for (int u = 0; u < NP; ++u) dst[o0 + (uint32_t)u * nn] = raw[u];
              }
            };
            int i = i_lo;
            for (; i + 1 < i_hi; i += 2) {  // two independent features per iteration: ILP for the MUFU chains
              float raw0[NP], raw1[NP], t0, t1;
              load_raw(i, raw0);
              load_raw(i + 1, raw1);
              const int f0 = p + 2 * (it.lin + i), f1 = f0 + 2;
              const float y0 = inv ? rq_apply_fast<KB, true>(raw0, D.range_min, D.range_max, xr[f0], t0)
                                   : rq_apply_fast<KB, false>(raw0, D.range_min, D.range_max, xr[f0], t0);
              const float y1 = inv ? rq_apply_fast<KB, true>(raw1, D.range_min, D.range_max, xr[f1], t1)
                                   : rq_apply_fast<KB, false>(raw1, D.range_min, D.range_max, xr[f1], t1);
              xr[f0] = y0;
              xr[f1] = y1;
              ldacc += t0;
              ldacc += t1;
              if (SPLIT) {
                xstage[(it.lin + i) * TC_M + row] = y0;
                xstage[(it.lin + i + 1) * TC_M + row] = y1;
              }
            }
            if (i < i_hi) {
              float raw0[NP], t0;
              load_raw(i, raw0);
              const int f0 = p + 2 * (it.lin + i);
              const float y0 = inv ? rq_apply_fast<KB, true>(raw0, D.range_min, D.range_max, xr[f0], t0)
                                   : rq_apply_fast<KB, false>(raw0, D.range_min, D.range_max, xr[f0], t0);
              xr[f0] = y0;
              ldacc += t0;
              if (SPLIT) xstage[(it.lin + i) * TC_M + row] = y0;
            }
            tc::tc_fence_before();
            arrive_mma(&S->acc_empty[slot]);
            if (SPLIT) {
              // the chunk's features [it.lin, it.lin + n_feat) x 128 rows are one contiguous block of the exchange
              // buffer: one bulk copy per peer, once every CTA of the cluster has finished reading the layer's input
              tc::fence_proxy_async_smem();  // every thread's exchange-buffer writes -> visible to the bulk copy engine
              epi_bar();
              if (tid == 0) {
                if (!peers_ready) tc::mbar_wait(&S->xbar[0], x_ph0);
                send_block(xstage + it.lin * TC_M, (uint32_t)(nf * TC_M * 4), false);
              }
              if (!peers_ready) {
                if (tid != 0) tc::mbar_wait(&S->xbar[0], x_ph0);
                peers_ready = true;
              }
            }
          }
          TC_STAMP(2);  // item epilogue done
          ++seq;
        }
        if (SPLIT) {
          // threads that wrote nothing still consume this layer's phase of xbar[0]; then the exchange round: every
          // CTA's transformed features have landed in every tile copy (also orders the row's two threads, like epi_bar)
          if (!peers_ready) tc::mbar_wait(&S->xbar[0], x_ph0);
          x_ph0 ^= 1;
          TC_STAMP(2);   // (split) guard consumed
          wait_round();  // every peer's chunks have landed in the exchange buffer
          TC_STAMP(2);   // (split) exchange round complete
          {              // ... copy the features other CTAs transformed into this CTA's tile (the row's two threads
                         // take alternate ordinals)
            const int ntf = (d - p + 1) / 2;
            for (int o = hf; o < ntf; o += TC_PARTS)
              if (((own_mask >> o) & 1ull) == 0) xr[p + 2 * o] = xstage[o * TC_M + row];
          }
          epi_bar();  // both threads of a row see each other's (and the peers') feature updates
        } else {
          epi_bar();  // both threads of a row see each other's feature updates
        }
      }
      if (MODE == TC_NF && pass == 0) {
        // proposal = inverse * sqrt(diag data_cov) + data_mean (rqSpline.py:495); keep it; re-whiten (rqSpline.py:501)
        for (int j = j_lo; j < j_hi; ++j) {
          const float sd = sqrtf(P[D.off_data_cov + (int64_t)j * d + j]);
          const float mu = P[D.off_data_mean + j];
          const float x = xr[j] * sd + mu;
          if (grow < a.n) __stcs(a.yout + grow * d + j, x);
          xr[j] = (x - mu) / sd;
        }
        ldacc = 0.0f;
        epi_bar();
      }
    }

    // ---- epilogue of the tile ------------------------------------------------------------------
    if (MODE == TC_TRAIN && a.save_x != nullptr) save_tile(L);  // (the layer loop ended with an epi_bar)
    S->ldpart[hf][row] = ldacc;
    epi_bar();
    const int post = (MODE == TC_NF) ? POST_BASE_LOGP : a.post;
    if (SPLIT) {
      // log-det of the row = sum over the cluster's CTAs of their chunks' contributions: every CTA sends its sum to
      // rank 0, which adds them in rank order
      expect_round(crank == 0 ? (R - 1u) * TC_M * 4u : 0u);
      if (hf == 0) {
        float ldsum = S->ldpart[0][row];
#pragma unroll
        for (int w = 1; w < TC_PARTS; ++w) ldsum += S->ldpart[w][row];
        S->ldx[crank][row] = ldsum;
      }
      tc::fence_proxy_async_smem();
      epi_bar();
      if (tid == 0 && crank != 0) send_block(&S->ldx[crank][0], TC_M * 4u, true);
      wait_round();
      if (crank == 0 && hf == 0 && grow < a.n) {
        float ldsum = S->ldx[0][row];
        for (uint32_t k = 1; k < R; ++k) ldsum += S->ldx[k][row];
        if (post == POST_BASE_LOGP) a.ldout[grow] = ldsum + base_log_prob(D, P, xr);
        else if (a.ldout != nullptr) a.ldout[grow] = ldsum;
      }
    } else if (post == POST_BASE_LOGP) {
      if (hf == 0 && grow < a.n) {
        float ldsum = S->ldpart[0][row];
#pragma unroll
        for (int w = 1; w < TC_PARTS; ++w) ldsum += S->ldpart[w][row];
        a.ldout[grow] = ldsum + base_log_prob(D, P, xr);
      }
    } else {
      if (grow < a.n) {
        for (int j = j_lo; j < j_hi; ++j) {
          float v = xr[j];
          if (post == POST_UNWHITEN) v = v * sqrtf(P[D.off_data_cov + (int64_t)j * d + j]) + P[D.off_data_mean + j];
          a.yout[grow * d + j] = v;
        }
        if (hf == 0 && a.ldout != nullptr) {
          float ldsum = S->ldpart[0][row];
#pragma unroll
          for (int w = 1; w < TC_PARTS; ++w) ldsum += S->ldpart[w][row];
          a.ldout[grow] = ldsum;
        }
      }
    }
    epi_bar();  // the tile in shared memory is free for the next one
    }  // tiles
    if (MODE == TC_TRAIN && a.act_img != nullptr && lane == 0) tc::bulk_wait<0>();  // this lane's bulk stores have landed
    tc::tc_fence_before();
  }
  __syncthreads();
  if (PAIR || SPLIT) tc::cluster_sync();  // no CTA leaves while another may still touch its barriers / shared / tensor memory
  tc::tc_fence_after();
  if (warp == TC_EPI_WARPS) {
    if (PAIR) tc::tmem_dealloc_pair<512>(tbase);
    else tc::tmem_dealloc<512>(tbase);
  }
}

template <int KB, int MODE, bool PAIR>
static int launch_tc_impl(const FlowmcFlowDesc& D, const TcProgram& PR, const TcArgs& a, cudaStream_t stream) {
  auto kern = flow_tc_kernel<KB, MODE, PAIR>;
  size_t bytes = 1024 + (size_t)TC_STAGES * TC_STAGE_BYTES + ((sizeof(TcSmem) + 15) & ~15) +
                 (size_t)TC_M * (D.n_features + 1) * sizeof(float) +
                 (size_t)D.n_layers * ((D.n_linear - 1) * 128 + ((D.n_features + 1) / 2) * (3 * D.num_bins + 1)) *
                     sizeof(float);
  TcArgs b = a;
  static size_t configured = 0;
  if (bytes > configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
      flowmc_set_error("flow (tensor-core path): cannot configure shared memory");
      return FLOWMC_ERR_CUDA;
    }
    configured = bytes;
  }
  const unsigned tiles = (unsigned)((a.n + TC_M - 1) / TC_M);
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
  }
  const unsigned persistent = (unsigned)(PAIR ? (n_sm & ~1) : n_sm);  // one CTA per SM (shared memory allows no more)
  cudaError_t e;
  if (PAIR) {
    // 2-CTA clusters: consecutive tiles pair up (an odd tile count gets one idle partner: its rows clamp to the last
    // row and nothing of it is stored)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(((tiles + 1u) & ~1u) < persistent ? ((tiles + 1u) & ~1u) : persistent);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, D, PR, b);
  } else {
    kern<<<tiles < persistent ? tiles : persistent, TC_THREADS, bytes, stream>>>(D, PR, b);
    e = cudaSuccess;
  }
  flowmc_count_launch();
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    flowmc_set_error(cudaGetErrorString(e));
    return FLOWMC_ERR_CUDA;
  }
  return FLOWMC_OK;
}

// dynamic shared memory of the split training-forward kernel: the one-CTA layout + the exchange buffer
static size_t tc_split_smem_bytes(const FlowmcFlowDesc& D) {
  return 1024 + (size_t)TC_STAGES * TC_STAGE_BYTES + ((sizeof(TcSmem) + 15) & ~15) +
         (size_t)TC_M * (D.n_features + 1) * sizeof(float) +
         (size_t)D.n_layers * ((D.n_linear - 1) * 128 + ((D.n_features + 1) / 2) * (3 * D.num_bins + 1)) * sizeof(float) +
         16 + (size_t)((D.n_features + 1) / 2) * TC_M * sizeof(float);
}

// co-resident clusters of R CTAs of the split training-forward kernel on this device (0 = unknown yet)
static int g_split_max_clusters[9] = {0};

// the backward kernel reports how many of ITS clusters are co-resident: the split factor honours the smaller number
void tc_split_note_max_clusters(int R, int n) {
  if (R >= 2 && R <= 8 && n > 0 && (g_split_max_clusters[R] == 0 || n < g_split_max_clusters[R])) g_split_max_clusters[R] = n;
}

// Feature split of the training forward pass: clusters of R CTAs per tile (see the kernel).  Returns the cluster size
// to use (1 = not applicable: the caller runs the one-CTA-per-tile kernel).
int tc_split_factor(const FlowmcFlowDesc& D, int64_t tiles) {
  TcProgram PR;
  if (tc_build_program(D, &PR)) return 1;
  static const int forced = [] {
    const char* e = std::getenv("FLOWMC_TC_SPLIT");  // 0 = off, 2 / 4 / 8 = force this cluster size, unset = automatic
    return e != nullptr ? std::atoi(e) : -1;
  }();
  if (forced == 0 || tc_split_smem_bytes(D) > 227 * 1024) return 1;
  const int nh = D.n_linear - 1;
  const int chunks = (PR.n_items[0] < PR.n_items[1] ? PR.n_items[0] : PR.n_items[1]) - nh;  // per layer, both parities
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
  }
  int r = 1;
  if (forced > 1) {
    r = forced;
  } else {
    // automatic: the largest cluster that still gives every tile its own cluster in ONE wave.  Clusters are placed
    // inside a GPC, so fewer 8-CTA clusters fit than SMs / 8 (measured through cudaOccupancyMaxActiveClusters by the
    // first launch of each size; until then a conservative guess).
    auto fits = [&](int rr) {
      const int cap = g_split_max_clusters[rr] > 0 ? g_split_max_clusters[rr] : (rr == 8 ? 12 : n_sm / rr - (rr == 4 ? 4 : 0));
      return tiles <= cap;
    };
    while (r < 8 && fits(2 * r)) r *= 2;
  }
  while (r > 1 && r > chunks) r /= 2;  // every CTA needs at least one chunk per layer (barrier phases stay aligned)
  return (r == 2 || r == 4 || r == 8) ? r : 1;
}

template <int KB>
static int launch_tc_split(const FlowmcFlowDesc& D, const TcProgram& PR, const TcArgs& a, int R, cudaStream_t stream) {
  auto kern = flow_tc_kernel<KB, TC_TRAIN, false, true>;
  const size_t bytes = tc_split_smem_bytes(D);
  static size_t configured = 0;
  if (bytes > configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess ||
        cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      flowmc_set_error("flow (tensor-core path, feature split): cannot configure the kernel");
      return FLOWMC_ERR_CUDA;
    }
    configured = bytes;
  }
  const unsigned tiles = (unsigned)((a.n + TC_M - 1) / TC_M);
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)R;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // co-resident clusters of this size on this device (cached per cluster size)
  int* max_clusters = g_split_max_clusters;
  if (max_clusters[R] == 0) {
    cfg.gridDim = dim3((unsigned)R);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = 148 / (R == 8 ? 9 : R);  // conservative
    }
    max_clusters[R] = n;
  }
  const unsigned n_clusters = tiles < (unsigned)max_clusters[R] ? tiles : (unsigned)max_clusters[R];
  cfg.gridDim = dim3(n_clusters * (unsigned)R);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, D, PR, a);
  flowmc_count_launch();
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    flowmc_set_error(cudaGetErrorString(e));
    return FLOWMC_ERR_CUDA;
  }
  return FLOWMC_OK;
}

// FLOWMC_TC_PAIR=1 runs the tiles as CTA pairs (cta_group::2, M = 256: each SM streams half of the weights, 8-slot
// ring).  Verified against the oracle like the default path, but measured SLOWER on B200 (C4 log_prob 86 M vs
// 100 M samples/s): the flow's GEMMs are short and separated by dependent epilogues, so every one of the ~15
// hand-offs per layer now waits for the slower of two SMs plus a cross-SM signal (layer period 30.9K vs 25.3K
// cycles, scripts/tc_timeline.py), which costs more than the halved weight stream gains.  Default: one CTA per tile.
template <int KB, int MODE>
static int launch_tc(const FlowmcFlowDesc& D, const TcProgram& PR, const TcArgs& a, cudaStream_t stream) {
  static const bool pair_ok = []() {
    const char* e = std::getenv("FLOWMC_TC_PAIR");
    return e != nullptr && e[0] == '1';
  }();
  if (pair_ok && a.n > TC_M) return launch_tc_impl<KB, MODE, true>(D, PR, a, stream);
  return launch_tc_impl<KB, MODE, false>(D, PR, a, stream);
}

template <int MODE>
static int dispatch_tc(const FlowmcFlowDesc& D, const TcProgram& PR, const TcArgs& a, cudaStream_t stream) {
  switch (D.num_bins) {
    case 4: return launch_tc<4, MODE>(D, PR, a, stream);
    case 8: return launch_tc<8, MODE>(D, PR, a, stream);
    case 16: return launch_tc<16, MODE>(D, PR, a, stream);
  }
  return FLOWMC_ERR_UNSUPPORTED;
}

static long long* g_tc_timing = nullptr;  // diagnostics hook (flowmc_trace_tc_timeline)

// true if the descriptor asks for (and the model shape allows) the tensor-core path
bool flow_tc_enabled(const FlowmcFlowDesc& D) { return D.tc_image != nullptr && D.tc_terms != 0 && tc_supported(D); }

int flow_transform_tc(const FlowmcFlowDesc& D, bool inverse, const float* P, const float* x, int64_t n, float* y,
                      float* ld, int pre, int post, const uint32_t* keys, Key hk, int64_t rpk, cudaStream_t stream,
                      const int32_t* idx, float* save_x, float* save_h, float* save_theta, uint8_t* act_img) {
  if (n <= 0) return FLOWMC_OK;
  TcProgram PR;
  if (int rc = tc_build_program(D, &PR)) return rc;
  TcArgs a;
  std::memset(&a, 0, sizeof(a));
  a.params = P; a.image = static_cast<const uint8_t*>(D.tc_image); a.xin = x; a.idx = idx; a.yout = y; a.ldout = ld;
  a.n = n; a.pre = pre; a.post = post; a.terms = D.tc_terms == 1 ? 1 : 3; a.keys = keys; a.host_key = hk;
  a.rows_per_key = rpk;
  a.save_x = save_x; a.save_h = save_h; a.save_theta = save_theta; a.act_img = act_img;
  a.timing = g_tc_timing;
  static const int dump_vec = [] {
    const char* e = std::getenv("FLOWMC_TC_DUMP");
    return (e != nullptr && e[0] == 'b') ? 0 : 1;  // "bulk" selects the cp.async.bulk stores
  }();
  a.dump_vec = dump_vec;
  static const int dbg_skip = [] {
    const char* e = std::getenv("FLOWMC_TC_DBG_SKIP");
    return e != nullptr ? std::atoi(e) : 0;
  }();
  a.dbg_skip = dbg_skip;
  if (inverse) return dispatch_tc<TC_INV>(D, PR, a, stream);
  if (save_x != nullptr || save_h != nullptr || save_theta != nullptr || act_img != nullptr) {
    // few tiles (a data-parallel rank's slice of the batch): clusters of CTAs share a tile, see flow_tc_kernel
    const int R = (save_h == nullptr && D.num_bins <= 8) ? tc_split_factor(D, (n + TC_M - 1) / TC_M) : 1;
    if (R > 1) {
      switch (D.num_bins) {
        case 4: return launch_tc_split<4>(D, PR, a, R, stream);
        case 8: return launch_tc_split<8>(D, PR, a, R, stream);
      }
    }
    return dispatch_tc<TC_TRAIN>(D, PR, a, stream);
  }
  return dispatch_tc<TC_FWD>(D, PR, a, stream);
}

int flow_nf_propose_tc(const FlowmcFlowDesc& D, const float* P, Key subkey, const uint32_t* chain_keys,
                       int64_t chain_offset, int64_t n_chains, int n_steps, int n_batch, int n_sample, float* props,
                       float* lp_nf, cudaStream_t stream) {
  TcProgram PR;
  if (int rc = tc_build_program(D, &PR)) return rc;
  TcArgs a;
  std::memset(&a, 0, sizeof(a));
  a.params = P; a.image = static_cast<const uint8_t*>(D.tc_image); a.yout = props; a.ldout = lp_nf;
  a.n = n_chains * n_steps; a.terms = D.tc_terms == 1 ? 1 : 3;
  a.subkey = subkey; a.chain_keys = chain_keys; a.chain_offset = chain_offset;
  a.n_steps = n_steps; a.n_batch = n_batch; a.n_sample = n_sample;
  return dispatch_tc<TC_NF>(D, PR, a, stream);
}

}  // namespace flowmc

extern "C" {

// diagnostics: device buffer of 3 * 256 int64 that CTA 0 of the next tensor-core flow launches stamps with clock64()
// (NULL switches it off).  Not thread-safe; used by scripts/tc_timeline.py only.
void flowmc_trace_tc_timeline(long long* buf) {
  flowmc::g_tc_timing = buf;
  flowmc::flow_backward_tc_set_timing(buf ? buf + 3 * 256 : nullptr);  // 4th row: backward epilogue thread 0
}

int64_t flowmc_flow_tc_image_bytes(const FlowmcFlowDesc* D) {
  using namespace flowmc;
  if (!D || !tc_supported(*D)) return 0;
  TcProgram PR;
  if (tc_build_program(*D, &PR)) return 0;
  return tc_image_bytes(*D, PR);
}

int flowmc_flow_tc_pack(const FlowmcFlowDesc* D, const float* params, void* image, void* stream) {
  using namespace flowmc;
  if (!D || !params || !image) {
    flowmc_set_error("flow_tc_pack: null argument");
    return FLOWMC_ERR_INVALID;
  }
  if (!tc_supported(*D)) {
    flowmc_set_error("flow_tc_pack: model shape not supported by the tensor-core path (need n_features <= 128, hidden "
                     "widths multiples of 16 and <= 128)");
    return FLOWMC_ERR_UNSUPPORTED;
  }
  TcProgram PR;
  if (int rc = tc_build_program(*D, &PR)) return rc;
  const int items = PR.n_items[0] > PR.n_items[1] ? PR.n_items[0] : PR.n_items[1];
  tc_pack_flow_kernel<<<dim3(items, D->n_layers, 8), 256, 0, (cudaStream_t)stream>>>(*D, PR, params,
                                                                                  static_cast<uint8_t*>(image));
  flowmc_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    flowmc_set_error(cudaGetErrorString(e));
    return FLOWMC_ERR_CUDA;
  }
  return FLOWMC_OK;
}

}  // extern "C"
