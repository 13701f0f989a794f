// Data-parallel optimiser step over NVLink peer memory: gradient all-reduce + global-norm clip + AdamW in ONE kernel.
//
// Reference semantics: NFModel.train_step (src/flowMC/resource/model/nf_model/base.py:102-125) with the Optimizer's
// optax.chain(clip_by_global_norm, adamw) (resource/optimizer.py:19-23) applied to the gradient of the GLOBAL batch;
// the reference is single-device, the split of a batch over ranks is the B200 addition.  The baseline
// for that step is NCCL all-reduce -> sumsq kernel -> clip_adamw kernel (flowmc_clip_adamw); this kernel does the same
// arithmetic with the collective inside it:
//
//   every rank owns one exchange block in its own HBM, mapped into every peer (CUDA IPC over NVLink / NVSwitch):
//     grad  [n + 4]   this rank's gradient of its slice of the batch (+ its loss), written by the backward pass
//     gsum  [n]       the all-reduced gradient, assembled from the ranks' slices
//     sumsq [world]   squared norm of each rank's slice of the reduced gradient (+ slot `world`: unused)
//     flags           cross-GPU barrier words, see barrier()
//   phase 1 (reduce-scatter)  rank r sums slice r of every rank's grad through peer LOADS, in rank order
//   phase 2 (all-gather)      ... and STORES the reduced slice and its squared norm into every rank's block
//   phase 3                   every rank: global norm from the world's slice norms (rank order), clip, Adam moments,
//                             decoupled weight decay, parameter update on the FULL vector (optimizer state stays
//                             replicated and bit-identical on every rank: every element of gsum was computed once, by
//                             its owner, and copied).
// Cross-GPU synchronisation: monotonically increasing epochs in flag words -- rank k release-stores the epoch into
// word [b][k] of every rank and acquire-polls its own words (no atomics across GPUs, nothing to reset).  All ranks must
// launch the kernel (it is part of the collective step like an NCCL call); a bounded spin turns a missing peer into an
// error flag instead of a hang.
#include <cstring>
#include <string>

#include "../../include/flowmc_b200.h"
#include "registry.h"

namespace flowmc {
namespace peer {

constexpr int kMaxWorld = 8;
constexpr int kMaxCtas = 1024;  // all CTAs must be co-resident (one wave): the kernel spins on flags; the launcher
                                // takes min(4 per SM, what the occupancy query allows)
constexpr int kThreads = 256;

struct Block {  // byte offsets inside a rank's exchange block
  int64_t grad, gsum, sumsq, flags, local, total;
};
__host__ __device__ inline int64_t pad256(int64_t v) { return (v + 255) & ~(int64_t)255; }
__host__ __device__ inline Block block_layout(int64_t n) {
  Block b;
  b.grad = 0;
  b.gsum = pad256((n + 4) * 4);
  b.sumsq = b.gsum + pad256(n * 4);
  b.flags = b.sumsq + 256;          // [3 barriers][kMaxWorld] uint32
  b.local = b.flags + 256;          // [0..2] grid-sync counters, [3] error flag, [4 .. 4 + kMaxCtas) float partials
  b.total = b.local + 256 + kMaxCtas * 4;
  return b;
}

struct Args {
  int rank, world;
  char* base[kMaxWorld];  // exchange block of every rank as seen from THIS rank (own = local pointer)
  int64_t n;
  float* params;
  float* mu;
  float* nu;
  float* loss_out;        // device, optional: the all-reduced loss
  uint32_t epoch;         // 1, 2, 3, ... one per call (host counter, identical on every rank)
  float lr, b1, b2, omb1, omb2, eps, wd, max_norm, bc1, bc2;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Grid-wide, then world-wide barrier number b of this call.  Every thread's prior global / peer stores are visible
// to every rank's threads after it.
__device__ __forceinline__ void barrier(const Args& a, const Block& L, int b) {
  uint32_t* local = reinterpret_cast<uint32_t*>(a.base[a.rank] + L.local);
  uint32_t* flags = reinterpret_cast<uint32_t*>(a.base[a.rank] + L.flags);
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    // last CTA of this GPU to arrive announces the rank to the world
    const uint32_t prev = atomicAdd(local + b, 1u);
    if (prev == a.epoch * gridDim.x - 1u) {
      __threadfence_system();
      for (int k = 0; k < a.world; ++k)
        st_release_sys(reinterpret_cast<uint32_t*>(a.base[k] + L.flags) + b * kMaxWorld + a.rank, a.epoch);
    }
    // everyone waits until every rank has announced this epoch
    for (int k = 0; k < a.world; ++k) {
      uint32_t spins = 0;
      while ((int32_t)(ld_acquire_sys(flags + b * kMaxWorld + k) - a.epoch) < 0) {
        if (++spins > (1u << 27)) {  // ~ seconds: a peer never arrived
          local[3] = 1u;
          break;
        }
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads) dp_reduce_adamw_kernel(const Args a) {
  const Block L = block_layout(a.n);
  const int64_t n = a.n;
  const int tid = threadIdx.x;
  const int64_t gtid = (int64_t)blockIdx.x * kThreads + tid, gstride = (int64_t)gridDim.x * kThreads;
  char* mine = a.base[a.rank];
  float* gsum_mine = reinterpret_cast<float*>(mine + L.gsum);
  float* sumsq_mine = reinterpret_cast<float*>(mine + L.sumsq);
  float* partials = reinterpret_cast<float*>(mine + L.local) + 4;
  __shared__ float sh[kThreads / 32];

  barrier(a, L, 0);  // every rank's gradient is in its block (each rank's kernel follows its backward pass in stream order)

  // the all-reduced loss (read here: after barrier 1 nobody touches a peer's `grad` any more, so the next backward
  // pass may overwrite it at once)
  float loss_sum = 0.0f;
  if (a.loss_out != nullptr && gtid == 0)
    for (int k = 0; k < a.world; ++k) loss_sum += __ldcg(reinterpret_cast<const float*>(a.base[k] + L.grad) + n);

  // ---- phase 1 + 2: reduce slice `rank` through peer loads, store it into every rank's gsum -------------------
  const int64_t n4 = n >> 2;  // n is a multiple of 4 (flat blob alignment)
  const int64_t per = (n4 + a.world - 1) / a.world;
  const int64_t lo = min(n4, (int64_t)a.rank * per), hi = min(n4, lo + per);
  float ss = 0.0f;
  for (int64_t i = lo + gtid; i < hi; i += gstride) {
    // all peers' loads in flight together (an NVLink round trip is ~2 us), summed in rank order: the sum is the same
    // whoever computes it
    float4 v[kMaxWorld];
#pragma unroll
    for (int k = 0; k < kMaxWorld; ++k)
      v[k] = (k < a.world) ? __ldcg(reinterpret_cast<const float4*>(a.base[k] + L.grad) + i)
                           : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float4 s = v[0];
#pragma unroll
    for (int k = 1; k < kMaxWorld; ++k)
      if (k < a.world) { s.x += v[k].x; s.y += v[k].y; s.z += v[k].z; s.w += v[k].w; }
    ss = fmaf(s.x, s.x, ss); ss = fmaf(s.y, s.y, ss); ss = fmaf(s.z, s.z, ss); ss = fmaf(s.w, s.w, ss);
#pragma unroll
    for (int k = 0; k < kMaxWorld; ++k)
      if (k < a.world) reinterpret_cast<float4*>(a.base[k] + L.gsum)[i] = s;
  }
  // squared norm of the slice: fixed-order two-stage sum (deterministic)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((tid & 31) == 0) sh[tid >> 5] = ss;
  __syncthreads();
  if (tid == 0) {
    float s = 0.0f;
    for (int w = 0; w < kThreads / 32; ++w) s += sh[w];
    partials[blockIdx.x] = s;
  }
  // the last CTA to finish sums the CTAs' partials in order and publishes the slice norm to every rank
  {
    uint32_t* local = reinterpret_cast<uint32_t*>(mine + L.local);
    __shared__ uint32_t s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(local + 2, 1u) == a.epoch * gridDim.x - 1u) ? 1u : 0u;
    __syncthreads();
    if (s_last && tid < 32) {  // fixed order: lane l takes partials l, l + 32, ...; then a fixed shuffle tree
      __threadfence();
      float s = 0.0f;
      for (unsigned c = tid; c < gridDim.x; c += 32) s += __ldcg(partials + c);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (tid == 0)
        for (int k = 0; k < a.world; ++k) reinterpret_cast<float*>(a.base[k] + L.sumsq)[a.rank] = s;
    }
  }
  barrier(a, L, 1);  // every slice of gsum and every slice norm has landed everywhere

  // ---- phase 3: clip by global norm + AdamW on the full vector (flow_train.cu clip_adamw_kernel's arithmetic) ----
  float tot = 0.0f;
  for (int k = 0; k < a.world; ++k) tot += __ldcg(sumsq_mine + k);
  const float gn = sqrtf(tot);
  const bool keep = gn < a.max_norm;
  for (int64_t i = gtid; i < n4; i += gstride) {  // four elements per thread and iteration, 16-byte accesses
    const float4 g4 = __ldcg(reinterpret_cast<const float4*>(gsum_mine) + i);
    float4 m4 = reinterpret_cast<float4*>(a.mu)[i], v4 = reinterpret_cast<float4*>(a.nu)[i];
    float4 p4 = reinterpret_cast<float4*>(a.params)[i];
    const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
    float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w}, pp[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float gi = gg[e];
      if (!keep) gi = (gi / gn) * a.max_norm;
      const float m = a.omb1 * gi + a.b1 * mm[e];
      const float v = a.omb2 * (gi * gi) + a.b2 * vv[e];
      mm[e] = m;
      vv[e] = v;
      float u = (m / a.bc1) / (sqrtf(v / a.bc2) + a.eps);
      u = u + a.wd * pp[e];
      pp[e] = pp[e] + (-a.lr) * u;
    }
    reinterpret_cast<float4*>(a.mu)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(a.nu)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    reinterpret_cast<float4*>(a.params)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
  }
  if (a.loss_out != nullptr && gtid == 0) *a.loss_out = loss_sum;
  // No third barrier: gsum / sumsq of this rank are rewritten only in phase 2 of the NEXT call, i.e. after that call's
  // barrier 0, which every rank reaches only after this kernel of its own has finished (stream order).
}

}  // namespace peer
}  // namespace flowmc

extern "C" {

int64_t flowmc_peer_block_bytes(int64_t n_params) {
  if (n_params <= 0 || (n_params & 3)) return 0;
  return flowmc::peer::block_layout(n_params).total;
}

int64_t flowmc_peer_block_offset(int64_t n_params, int what) {
  const flowmc::peer::Block b = flowmc::peer::block_layout(n_params);
  switch (what) {
    case 0: return b.grad;
    case 1: return b.gsum;
    case 2: return b.sumsq;
    case 3: return b.flags;
    case 4: return b.local;
  }
  return -1;
}

int flowmc_ipc_alloc(int64_t bytes, void** ptr, unsigned char handle[64]) {
  if (!ptr || !handle || bytes <= 0) {
    flowmc_set_error("ipc_alloc: bad arguments");
    return FLOWMC_ERR_INVALID;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    if (p) cudaFree(p);
    cudaGetLastError();
    flowmc_set_error((std::string("ipc_alloc: ") + cudaGetErrorString(e)).c_str());
    return FLOWMC_ERR_CUDA;
  }
  std::memcpy(handle, &h, 64);
  *ptr = p;
  return FLOWMC_OK;
}

int flowmc_ipc_open(const unsigned char handle[64], void** ptr) {
  if (!ptr || !handle) {
    flowmc_set_error("ipc_open: bad arguments");
    return FLOWMC_ERR_INVALID;
  }
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    flowmc_set_error((std::string("ipc_open: ") + cudaGetErrorString(e)).c_str());
    return FLOWMC_ERR_CUDA;
  }
  return FLOWMC_OK;
}

int flowmc_ipc_close(void* ptr) { return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? FLOWMC_OK : FLOWMC_ERR_CUDA; }
int flowmc_ipc_free(void* ptr) { return cudaFree(ptr) == cudaSuccess ? FLOWMC_OK : FLOWMC_ERR_CUDA; }

int flowmc_dp_reduce_adamw(int rank, int world, void* const* blocks, int64_t n_params, float* params, float* mu,
                           float* nu, int64_t count, double lr, double b1, double b2, double eps, double weight_decay,
                           double max_norm, uint32_t epoch, float* loss_out, void* stream_) {
  using namespace flowmc::peer;
  if (rank < 0 || world < 1 || world > kMaxWorld || rank >= world || !blocks || n_params <= 0 || (n_params & 3) ||
      !params || !mu || !nu || count < 1 || epoch < 1) {
    flowmc_set_error("dp_reduce_adamw: bad arguments (world <= 8, n_params % 4 == 0, epoch >= 1)");
    return FLOWMC_ERR_INVALID;
  }
  Args a;
  std::memset(&a, 0, sizeof(a));
  a.rank = rank;
  a.world = world;
  for (int k = 0; k < world; ++k) {
    if (!blocks[k]) {
      flowmc_set_error("dp_reduce_adamw: null peer block");
      return FLOWMC_ERR_INVALID;
    }
    a.base[k] = static_cast<char*>(blocks[k]);
  }
  a.n = n_params;
  a.params = params; a.mu = mu; a.nu = nu; a.loss_out = loss_out;
  a.epoch = epoch;
  a.lr = (float)lr; a.b1 = (float)b1; a.b2 = (float)b2; a.eps = (float)eps; a.wd = (float)weight_decay;
  a.max_norm = (float)max_norm;
  a.omb1 = (float)(1.0 - b1);
  a.omb2 = (float)(1.0 - b2);
  a.bc1 = 1.0f - powf(a.b1, (float)count);
  a.bc2 = 1.0f - powf(a.b2, (float)count);
  // one co-resident wave: 4 CTAs per SM if the occupancy allows it (the kernel has no shared memory to speak of)
  static int grid = 0;
  if (grid == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dp_reduce_adamw_kernel, kThreads, 0);
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    grid = sms * per_sm;
    if (grid > kMaxCtas) grid = kMaxCtas;
    if (grid < 1) grid = 64;
  }
  dp_reduce_adamw_kernel<<<grid, kThreads, 0, (cudaStream_t)stream_>>>(a);
  flowmc_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    flowmc_set_error(cudaGetErrorString(e));
    return FLOWMC_ERR_CUDA;
  }
  return FLOWMC_OK;
}

}  // extern "C"
