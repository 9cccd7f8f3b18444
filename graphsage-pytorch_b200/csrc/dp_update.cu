// Data-parallel exchange step fused with the update: gradient all-reduce over NVLink peer
// memory + per-model clip_grad_norm_ + SGD + gradient zeroing, in ONE kernel.
//
// Reference semantics: src/utils.py:184-191 (loss.backward(); clip_grad_norm_(model, 5) for
// each of the two models; optimizer.step(); zero_grad()), executed per data-parallel rank on
// the mean gradient.  The reference has no distributed code; the stock way to
// write this step is ncclAllReduce followed by 2 x (norm kernel + update kernel) = 5 launches
// and two latency-bound round trips for a 258 KB buffer.  Here:
//
//   push    every CTA copies its slice of the local flat gradient into a receive slot inside
//           every peer's memory (plain 16-byte stores over NVLink; fire and forget), then
//           publishes flag[rank][cta] = epoch with a system-scope release store;
//   wait    the CTA spins on the W-1 flags of ITS OWN slice only (no grid-wide dependency on
//           the network), acquire at system scope;
//   reduce  sums the W copies in rank order 0..W-1 -- every rank adds the same numbers in the
//           same order, so replicas stay bit-identical -- scales by 1/W, accumulates the
//           per-model sum of squares;
//   clip    one grid barrier (all CTAs are co-resident: <= 64 CTAs of 256 threads); the
//           per-CTA partial sums are combined in CTA order by every CTA (deterministic);
//   update  p -= lr * min(1, max_norm / (norm + 1e-6)) * g, gradient slice zeroed for the
//           next step.
//
// Receive slots are double-buffered by epoch parity: a rank can only start epoch e+2 (and
// overwrite parity e&1 in a peer) after that peer published epoch e+1, i.e. after the peer's
// epoch-e kernel has completed.  A wait that exceeds `timeout_ns` records status = 1 and
// proceeds (the host checks gs_dp_status) instead of hanging the GPU.
#include "common.cuh"

namespace gs {

constexpr int kDpThreads = 256;
constexpr int kDpMaxCtas = 64;
constexpr int kDpMaxWorld = 8;     // one NVSwitch domain of 8 GPUs
constexpr int kDpMaxSegs = 16;
constexpr int kDpMaxGroups = 4;

struct DpPeers {
  float* recv[kDpMaxWorld];        // rank r's receive region: [2 parities][world][n_total] floats
  uint32_t* flags[kDpMaxWorld];    // rank r's flag region:    [world][kDpMaxCtas] epochs
  int rank, world;
};
struct DpSegs {
  float* param[kDpMaxSegs];
  float* param_lo[kDpMaxSegs];     // nullable: receives p - trunc_tf32(p) of the updated parameter (K4's split operand)
  float* extra[kDpMaxSegs];        // nullable: extra_n[k] more partial gradients of the tensor, extra_stride[k] floats apart
  int extra_n[kDpMaxSegs];         //   (replicas a producer spread its atomics over); folded into the sum and cleared
  long long extra_stride[kDpMaxSegs];
  long long off[kDpMaxSegs];       // offset of the tensor's gradient inside the flat buffer (multiple of 4)
  long long numel[kDpMaxSegs];
  int group[kDpMaxSegs];           // clip group (model) of the tensor
  int n, groups, any_extra;
};
struct DpState {                   // device memory, zeroed once by the caller
  unsigned int epoch;
  unsigned int status;
  unsigned long long arrive;
  float partial[kDpMaxCtas][kDpMaxGroups];
  float norm[kDpMaxGroups];        // last step's total gradient norms (diagnostics / tests)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

__global__ void split_lo_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n) {
  pdl_sync();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] = tf32_lo(src[i]);
}

__device__ __forceinline__ int seg_of(const DpSegs& s, long long elem) {
  int k = -1;
#pragma unroll 4
  for (int i = 0; i < s.n; ++i)
    if (elem >= s.off[i] && elem < s.off[i] + ((s.numel[i] + 3) & ~3LL)) k = i;
  return k;
}

#ifdef GS_TOP_TRACE
__device__ long long g_dp_trace[16];
#define GS_DP_MARK(slot) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_dp_trace[(slot)] = clock64(); } while (0)
#else
#define GS_DP_MARK(slot) do { } while (0)
#endif

// Elements are float4s; thread t of CTA c owns float4s  c*256 + t + j*G*256  (j < kDpPer): consecutive threads touch
// consecutive float4s, and everything a thread needs -- its gradient pieces, the segment they fall in, the parameter
// pieces they update -- is fetched in ONE round of independent loads before the grid barrier, so the kernel is
// (one memory round trip) + (the barrier) + (stores).  The tables stay in the constant bank (__grid_constant__).
constexpr int kDpPer = 4;          // float4s per thread held in registers; larger buffers take the looping path

__global__ void __launch_bounds__(kDpThreads)
dp_update_kernel(float* __restrict__ flat, long long n_total, const __grid_constant__ DpPeers peers,
                 const __grid_constant__ DpSegs segs, DpState* __restrict__ st, float max_norm, float lr,
                 unsigned long long timeout_ns, long long* __restrict__ step_counter) {
  GS_DP_MARK(0);
  pdl_sync();
  GS_DP_MARK(1);
  const int G = gridDim.x, c = blockIdx.x, tid = threadIdx.x;
  const unsigned int e = *reinterpret_cast<volatile unsigned int*>(&st->epoch) + 1u;
  const int par = static_cast<int>(e & 1u);
  const long long n4 = n_total >> 2;
  float4* flat4 = reinterpret_cast<float4*>(flat);
  const int W = peers.world, me = peers.rank;
  const long long first = static_cast<long long>(c) * kDpThreads + tid;
  const long long stride = static_cast<long long>(G) * kDpThreads;
  GS_DP_MARK(2);

  // partial-gradient replicas of a segment are folded into the flat buffer (and cleared) before anything reads it
  auto fold = [&](long long i, int k, float4 g) -> float4 {
    if (k >= 0 && segs.extra[k] != nullptr) {
      const long long o = 4 * i - segs.off[k];
      const int n = segs.extra_n[k];
      for (int x0 = 0; x0 < n; x0 += 8) {                  // 8 independent loads in flight, then the adds in replica order
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          v[u] = x0 + u < n ? __ldcg(reinterpret_cast<const float4*>(segs.extra[k] + (x0 + u) * segs.extra_stride[k] + o))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          g.x += v[u].x; g.y += v[u].y; g.z += v[u].z; g.w += v[u].w;
          if (x0 + u < n) *reinterpret_cast<float4*>(segs.extra[k] + (x0 + u) * segs.extra_stride[k] + o) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    return g;
  };
  if (segs.any_extra)                                      // (uniform branch; off in the default configuration)
    for (long long i = first; i < n4; i += stride) {
      const int k = seg_of(segs, 4 * i);
      if (k >= 0 && segs.extra[k] != nullptr) flat4[i] = fold(i, k, flat4[i]);
    }
  if (W > 1) {
    // ---- push my pieces into every peer's slot [par][me] ----
    for (int p = 0; p < W; ++p) {
      if (p == me) continue;
      float4* dst = reinterpret_cast<float4*>(peers.recv[p]) + static_cast<long long>(par * W + me) * n4;
      for (long long i = first; i < n4; i += stride) dst[i] = flat4[i];
    }
    // No per-thread system fence here: the CTA barrier orders every thread's remote stores before the flag
    // writers, and their release at system scope is cumulative -- one fence round trip over NVLink, not two.
    __syncthreads();
    if (tid < W && tid != me) {
      st_release_sys(peers.flags[tid] + me * kDpMaxCtas + c, e);
      // ---- wait for peer `tid`'s copy of the pieces CTA c owns ----
      const uint32_t* f = peers.flags[me] + tid * kDpMaxCtas + c;
      unsigned long long t0 = 0;
      for (unsigned spins = 0; static_cast<int>(ld_acquire_sys(f) - e) < 0; ++spins) {
        if ((spins & 63u) == 63u) {                        // the timer is only consulted once the wait is long
          const unsigned long long now = global_ns();
          if (t0 == 0) t0 = now;
          if (now - t0 > timeout_ns) { atomicExch(&st->status, 1u); break; }
          __nanosleep(64);
        }
      }
    }
    __syncthreads();
  }
  GS_DP_MARK(3);

  // ---- one round of loads: gradient pieces (all ranks' copies), their segments, the parameters they update ----
  const float inv_w = 1.0f / static_cast<float>(W);
  const float4* mine = reinterpret_cast<const float4*>(peers.recv[me]) + static_cast<long long>(par) * W * n4;
  const bool in_regs = n4 <= stride * kDpPer;
  float4 gsum[kDpPer], wv[kDpPer];
  int segk[kDpPer];
  float ss[kDpMaxGroups];
#pragma unroll
  for (int g = 0; g < kDpMaxGroups; ++g) ss[g] = 0.f;
  auto reduce_one = [&](long long i, int k) -> float4 {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < W; ++r) {                          // rank order: every rank adds the same numbers in the same order
      const float4 v = (r == me) ? flat4[i] : __ldcg(mine + static_cast<long long>(r) * n4 + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    s.x *= inv_w; s.y *= inv_w; s.z *= inv_w; s.w *= inv_w;
    return s;
  };
  auto add_ss = [&](const float4& s, int k) {
    if (k < 0) return;
    const float q = s.x * s.x + s.y * s.y + s.z * s.z + s.w * s.w;     // padding elements are zero
    const int g = segs.group[k];
#pragma unroll
    for (int gg = 0; gg < kDpMaxGroups; ++gg) if (gg == g) ss[gg] += q;
  };
  if (in_regs) {
#pragma unroll
    for (int j = 0; j < kDpPer; ++j) {
      const long long i = first + j * stride;
      segk[j] = -1;
      if (i < n4) {
        const int k = seg_of(segs, 4 * i);
        gsum[j] = reduce_one(i, k);
        segk[j] = k;
        if (k >= 0) {
          const long long o = 4 * i - segs.off[k];
          const float* p = segs.param[k] + o;
          if (segs.numel[k] - o >= 4 && (reinterpret_cast<uintptr_t>(p) & 15u) == 0) wv[j] = *reinterpret_cast<const float4*>(p);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kDpPer; ++j) add_ss(gsum[j], segk[j]);
  } else {
    for (long long i = first; i < n4; i += stride) {
      const int k = seg_of(segs, 4 * i);
      const float4 s = reduce_one(i, k);
      flat4[i] = s;
      add_ss(s, k);
    }
  }
  __shared__ float s_red[kDpThreads / 32][kDpMaxGroups];
  __shared__ float s_coef[kDpMaxGroups];
#pragma unroll
  for (int g = 0; g < kDpMaxGroups; ++g) {
    const float v = warp_sum(ss[g]);
    if ((tid & 31) == 0) s_red[tid >> 5][g] = v;
  }
  __syncthreads();
  if (tid < kDpMaxGroups) {
    float v = 0.f;
    for (int w = 0; w < kDpThreads / 32; ++w) v += s_red[w][tid];
    st->partial[c][tid] = v;
  }
  GS_DP_MARK(4);
  // ---- grid barrier (arrive counter grows by G every epoch) ----
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    atomicAdd(&st->arrive, 1ULL);
    const unsigned long long target = static_cast<unsigned long long>(e) * static_cast<unsigned long long>(G);
    unsigned long long t0 = 0;
    for (unsigned spins = 0; ld_acquire_gpu_u64(&st->arrive) < target; ++spins) {
      if ((spins & 255u) == 255u) {
        const unsigned long long now = global_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > timeout_ns) { atomicExch(&st->status, 2u); break; }
      }
    }
  }
  __syncthreads();
  GS_DP_MARK(5);
  if (tid < 32) {
    // all G x groups partials in flight at once (lane = CTA, G <= 64), then a shuffle tree: the order of the
    // additions is fixed by the lane numbers, hence identical on every CTA and every rank
    float v[kDpMaxGroups];
#pragma unroll
    for (int g = 0; g < kDpMaxGroups; ++g) {
      const float a = tid < G ? __ldcg(&st->partial[tid][g]) : 0.f;
      const float b = tid + 32 < G ? __ldcg(&st->partial[tid + 32][g]) : 0.f;
      v[g] = a + b;
    }
#pragma unroll
    for (int g = 0; g < kDpMaxGroups; ++g) {
      const float tot = warp_sum(v[g]);
      if (tid == g) {
        const float norm = sqrtf(tot);
        // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to <= 1
        s_coef[g] = max_norm > 0.f ? fminf(max_norm / (norm + 1e-6f), 1.0f) : 1.0f;
        if (c == 0) st->norm[g] = norm;
      }
    }
  }
  __syncthreads();
  GS_DP_MARK(6);

  // ---- SGD on my pieces, gradient zeroed for the next step ----
  auto sgd_one = [&](long long i, const float4& g, int k, const float4* w_have) {
    if (k >= 0) {
      const float step = lr * s_coef[segs.group[k]];
      const long long o = 4 * i - segs.off[k];
      float* p = segs.param[k] + o;
      const long long left = segs.numel[k] - o;
      if (left >= 4 && (reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
        float4 w = w_have ? *w_have : *reinterpret_cast<float4*>(p);
        w.x = fmaf(-step, g.x, w.x); w.y = fmaf(-step, g.y, w.y); w.z = fmaf(-step, g.z, w.z); w.w = fmaf(-step, g.w, w.w);
        *reinterpret_cast<float4*>(p) = w;
        if (segs.param_lo[k] != nullptr)
          *reinterpret_cast<float4*>(segs.param_lo[k] + o) = make_float4(tf32_lo(w.x), tf32_lo(w.y), tf32_lo(w.z), tf32_lo(w.w));
      } else {
        const float gv[4] = {g.x, g.y, g.z, g.w};
        for (int j = 0; j < 4 && j < left; ++j) {
          p[j] = fmaf(-step, gv[j], p[j]);
          if (segs.param_lo[k] != nullptr) segs.param_lo[k][o + j] = tf32_lo(p[j]);
        }
      }
    }
    flat4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  };
  if (in_regs) {
#pragma unroll
    for (int j = 0; j < kDpPer; ++j) {
      const long long i = first + j * stride;
      if (i < n4) sgd_one(i, gsum[j], segk[j], &wv[j]);
    }
  } else {
    for (long long i = first; i < n4; i += stride) sgd_one(i, flat4[i], seg_of(segs, 4 * i), nullptr);
  }
  GS_DP_MARK(7);
  if (c == 0 && tid == 0) {
    st->epoch = e;
    if (step_counter) *step_counter += 1;          // Philox offset of the next step's sampler (trainer.py)
  }
}

}  // namespace gs

using namespace gs;

extern "C" size_t gs_dp_state_bytes(void) { return sizeof(DpState); }

extern "C" size_t gs_dp_region_bytes(int64_t n_total, int32_t world) {
  if (n_total < 0 || world < 1 || world > kDpMaxWorld) return 0;
  const size_t flags = static_cast<size_t>(kDpMaxWorld) * kDpMaxCtas * sizeof(uint32_t);       // 2 KB
  return flags + 2 * static_cast<size_t>(world) * static_cast<size_t>(n_total) * sizeof(float);
}

extern "C" size_t gs_dp_region_recv_offset(void) {
  return static_cast<size_t>(kDpMaxWorld) * kDpMaxCtas * sizeof(uint32_t);
}

extern "C" int gs_dp_allreduce_clip_sgd(float* flat_grad, int64_t n_total, void* const* peer_regions_host, int32_t rank,
                                        int32_t world, float* const* seg_params_host, const int64_t* seg_offsets_host,
                                        const int64_t* seg_numels_host, const int32_t* seg_groups_host, int32_t num_segs,
                                        float max_norm, float lr, void* state, uint64_t timeout_ns,
                                        int64_t* step_counter, float* const* seg_params_lo_host,
                                        float* const* seg_extra_host, const int32_t* seg_extra_n_host,
                                        const int64_t* seg_extra_stride_host, gs_stream_t stream) {
  if (!flat_grad || !state || n_total < 4 || (n_total & 3) || !aligned16(flat_grad)) return GS_ERR_BAD_ARG;
  if (world < 1 || world > kDpMaxWorld || rank < 0 || rank >= world) return GS_ERR_BAD_ARG;
  if (num_segs < 1 || num_segs > kDpMaxSegs || !seg_params_host || !seg_offsets_host || !seg_numels_host) return GS_ERR_BAD_ARG;
  if (world > 1 && !peer_regions_host) return GS_ERR_BAD_ARG;
  DpPeers peers{};
  peers.rank = rank;
  peers.world = world;
  for (int r = 0; r < world; ++r) {
    unsigned char* base = world > 1 ? static_cast<unsigned char*>(peer_regions_host[r]) : nullptr;
    if (world > 1 && (!base || !aligned16(base))) return GS_ERR_BAD_ARG;
    peers.flags[r] = reinterpret_cast<uint32_t*>(base);
    peers.recv[r] = reinterpret_cast<float*>(base + gs_dp_region_recv_offset());
  }
  DpSegs segs{};
  segs.n = num_segs;
  int groups = 1;
  for (int i = 0; i < num_segs; ++i) {
    segs.param[i] = seg_params_host[i];
    segs.param_lo[i] = seg_params_lo_host ? seg_params_lo_host[i] : nullptr;
    if (segs.param_lo[i] && !aligned16(segs.param_lo[i])) return GS_ERR_ALIGNMENT;
    segs.extra[i] = seg_extra_host ? seg_extra_host[i] : nullptr;
    segs.extra_n[i] = (segs.extra[i] && seg_extra_n_host) ? seg_extra_n_host[i] : 0;
    segs.extra_stride[i] = (segs.extra[i] && seg_extra_stride_host) ? seg_extra_stride_host[i] : 0;
    if (segs.extra[i]) {
      if (!aligned16(segs.extra[i]) || (segs.extra_stride[i] & 3) || segs.extra_stride[i] < ((segs.numel[i] + 3) & ~3LL) ||
          segs.extra_n[i] < 1 || segs.extra_n[i] > 16)
        return GS_ERR_BAD_ARG;
      segs.any_extra = 1;
    } else {
      segs.extra_n[i] = 0;
    }
    segs.off[i] = seg_offsets_host[i];
    segs.numel[i] = seg_numels_host[i];
    segs.group[i] = seg_groups_host ? seg_groups_host[i] : 0;
    if (!segs.param[i] || (segs.off[i] & 3) || segs.off[i] < 0 || segs.numel[i] < 0 ||
        segs.off[i] + segs.numel[i] > n_total || segs.group[i] < 0 || segs.group[i] >= kDpMaxGroups)
      return GS_ERR_BAD_ARG;
    if (segs.group[i] + 1 > groups) groups = segs.group[i] + 1;
  }
  segs.groups = groups;
  // the grid size is a pure function of n_total: the barrier counter of `state` relies on it
  const int64_t n4 = n_total >> 2;
  int grid = static_cast<int>((n4 + kDpThreads - 1) / kDpThreads);      // one float4 per thread while <= 64 CTAs suffice
  if (grid > kDpMaxCtas) grid = kDpMaxCtas;
  if (grid < 1) grid = 1;
  launch(dp_update_kernel, grid, kDpThreads, 0, as_stream(stream), flat_grad, n_total, peers, segs,
                                                              static_cast<DpState*>(state), max_norm, lr,
                                                              timeout_ns ? timeout_ns : 2000000000ULL,
                                                              reinterpret_cast<long long*>(step_counter));
  return finish_launch();
}

// dst[i] = src[i] - trunc_tf32(src[i]): the low half of the 3-term tf32 split of a weight (kept current by the update
// kernel afterwards, see seg_params_lo_host)
extern "C" int gs_split_lo(const float* src, float* dst, int64_t n, gs_stream_t stream) {
  if (!src || !dst || n < 0) return GS_ERR_BAD_ARG;
  if (n == 0) return GS_OK;
  const int blocks = static_cast<int>(n / 256 + 1 < 1184 ? n / 256 + 1 : 1184);
  launch(split_lo_kernel, blocks, 256, 0, as_stream(stream), src, dst, static_cast<long long>(n));
  return finish_launch();
}

// status (0 ok, 1 peer wait timed out, 2 grid barrier timed out), epoch and the last norms; synchronises the stream
extern "C" int gs_dp_status(const void* state, uint32_t* epoch_host, uint32_t* status_host, float* norms_host,
                            gs_stream_t stream) {
  if (!state) return GS_ERR_BAD_ARG;
  DpState h;
  cudaError_t e = cudaMemcpyAsync(&h, state, sizeof(DpState), cudaMemcpyDeviceToHost, as_stream(stream));
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaStreamSynchronize(as_stream(stream));
  if (e != cudaSuccess) return static_cast<int>(e);
  if (epoch_host) *epoch_host = h.epoch;
  if (status_host) *status_host = h.status;
  if (norms_host) for (int g = 0; g < kDpMaxGroups; ++g) norms_host[g] = h.norm[g];
  return GS_OK;
}

// ---------------------------------------------------------------------------------------------
// peer memory: cudaMalloc'ed regions shared between the ranks of one box with CUDA IPC
// ---------------------------------------------------------------------------------------------
extern "C" int gs_peer_alloc(size_t bytes, void** out_ptr_host) {
  if (!out_ptr_host || bytes == 0) return GS_ERR_BAD_ARG;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaMemset(p, 0, bytes);
  if (e != cudaSuccess) { cudaFree(p); return static_cast<int>(e); }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { cudaFree(p); return static_cast<int>(e); }
  *out_ptr_host = p;
  return GS_OK;
}
extern "C" int gs_peer_free(void* ptr) {
  if (!ptr) return GS_OK;
  return static_cast<int>(cudaFree(ptr));
}
extern "C" int gs_peer_export(void* ptr, unsigned char* handle64_host) {
  if (!ptr || !handle64_host) return GS_ERR_BAD_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) return static_cast<int>(e);
  memcpy(handle64_host, &h, 64);
  return GS_OK;
}
extern "C" int gs_peer_open(const unsigned char* handle64_host, void** out_ptr_host) {
  if (!handle64_host || !out_ptr_host) return GS_ERR_BAD_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64_host, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return static_cast<int>(e);
  *out_ptr_host = p;
  return GS_OK;
}
extern "C" int gs_peer_close(void* ptr) {
  if (!ptr) return GS_OK;
  return static_cast<int>(cudaIpcCloseMemHandle(ptr));
}

#ifdef GS_TOP_TRACE
extern "C" int gs_debug_dp_trace_read(long long* host_out, int n) {
  if (n > 16) n = 16;
  cudaDeviceSynchronize();
  return static_cast<int>(cudaMemcpyFromSymbol(host_out, gs::g_dp_trace, sizeof(long long) * n));
}
#endif
