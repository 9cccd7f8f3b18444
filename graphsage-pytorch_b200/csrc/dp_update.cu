// Data-parallel exchange step fused with the update: gradient all-reduce over NVLink peer
// memory + per-model clip_grad_norm_ + SGD + gradient zeroing, in ONE kernel.
//
// Reference semantics: src/utils.py:184-191 (loss.backward(); clip_grad_norm_(model, 5) for
// each of the two models; optimizer.step(); zero_grad()), executed per data-parallel rank on
// the mean gradient.  The reference has no distributed code; the stock way to
// write this step is ncclAllReduce followed by 2 x (norm kernel + update kernel) = 5 launches
// and two latency-bound round trips for a 258 KB buffer.  Here:
//
//   push    every thread copies its float4s of the local flat gradient into a receive slot inside every peer's
//           memory (plain 16-byte stores over NVLink; fire and forget).  There is NO flag and NO fence: an empty
//           receive word holds the bit pattern 0xffffffff (a NaN payload arithmetic never produces; the sender
//           rewrites a gradient word that happens to carry it to the canonical NaN), so the data are their own
//           arrival signal, word by word -- no assumption about the atomicity of a 16-byte store either;
//   wait    the thread polls ITS OWN float4s of the W-1 peer slots until none of the words is empty (no CTA- or
//           grid-wide dependency on the network) and puts the empty pattern back for the epoch after next;
//   reduce  sums the W copies in rank order 0..W-1 -- every rank adds the same numbers in the same order, so
//           replicas stay bit-identical -- scales by 1/W, accumulates the per-model sum of squares;
//   clip    the per-CTA partial sums of squares meet through tagged 8-byte records every CTA polls (all CTAs are
//           co-resident: <= 64 CTAs of 256 threads) and are combined in CTA order by every CTA (deterministic);
//   update  p -= lr * min(1, max_norm / (norm + 1e-6)) * g, gradient slice zeroed for the
//           next step.
//
// Compared with the first version (data, CTA barrier, release flag at system scope, acquire poll of the flag) the
// critical path loses the fence's round trip over NVLink and the flag's flight: measured in lockstep chains on
// 2 x B200 (scratch/dp_probe_mp.py), push + wait was 9.4-11.8K cycles per launch.
//
// Receive slots are double-buffered by epoch parity: a rank can only start epoch e+2 (and overwrite parity e&1 in a
// peer) after it has seen that peer's epoch-e+1 data, i.e. after the peer's epoch-e kernel -- which emptied the slot --
// has completed.  The regions start out filled with 0xff bytes (peer.py).  A wait that exceeds `timeout_ns` records
// status = 1 and proceeds (the host checks gs_dp_status) instead of hanging the GPU.
#include "common.cuh"

namespace gs {

constexpr int kDpThreads = 256;
constexpr int kDpMaxCtas = 64;
constexpr int kDpMaxWorld = 8;     // one NVSwitch domain of 8 GPUs
constexpr int kDpMaxSegs = 16;
constexpr int kDpMaxGroups = 4;

struct DpPeers {
  float* recv[kDpMaxWorld];        // rank r's receive region: [2 parities][world][n_total] floats, empty words = 0xffffffff
  uint32_t* flags[kDpMaxWorld];    // rank r's 2 KB header (the first protocol's flags; reserved)
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
  unsigned long long rec[kDpMaxCtas][kDpMaxGroups];   // {epoch << 32 | partial sum of squares} of CTA c, clip group g
  float norm[kDpMaxGroups];        // last step's total gradient norms (diagnostics / tests)
};

constexpr uint32_t kDpEmpty = 0xffffffffu;   // a receive word nobody has written yet this epoch

__device__ __forceinline__ uint4 ld_relaxed_sys_v4(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_v4(void* p, uint4 v) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ bool dp_landed(const uint4& v) {
  return v.x != kDpEmpty && v.y != kDpEmpty && v.z != kDpEmpty && v.w != kDpEmpty;
}
__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
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

// slow path of the wait: poll one float4 of a receive slot until every word has landed (or the timeout strikes)
__device__ __noinline__ uint4 dp_poll(const float4* slot, unsigned long long timeout_ns, unsigned int* status) {
  unsigned long long t0 = 0;
  uint4 v = ld_relaxed_sys_v4(slot);
  for (unsigned spins = 0; !dp_landed(v); ++spins) {
    if ((spins & 63u) == 63u) {                            // the timer is only consulted once the wait is long
      const unsigned long long now = global_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > timeout_ns) { atomicExch(status, 1u); break; }
      __nanosleep(64);
    }
    v = ld_relaxed_sys_v4(slot);
  }
  return v;
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

// Elements are float4s; thread t of CTA c owns float4s  c*256 + t + j*G*256: consecutive threads touch consecutive
// float4s.  While the buffer fits one float4 per thread (G <= kDpMaxCtas CTAs; 64432 floats at the headline
// configuration = 63 CTAs) everything a thread needs -- its gradient piece (all ranks' copies), the segment it falls in,
// the parameter piece it updates -- is fetched in ONE round of independent loads before the grid barrier and stays in
// registers, so the kernel is (one memory round trip) + (the barrier) + (stores); larger buffers loop and park the reduced
// gradient in the flat buffer across the barrier.  One copy of the code serves both (the kernel is launch-latency
// bound: its instruction footprint is part of its cost).  The tables stay in the constant bank (__grid_constant__).
__global__ void __launch_bounds__(kDpThreads, 4)     // <= 64 registers: the CTAs slot in beside the previous kernel's (PDL)
dp_update_kernel(float* __restrict__ flat, long long n_total, const __grid_constant__ DpPeers peers,
                 const __grid_constant__ DpSegs segs, DpState* __restrict__ st, float max_norm, float lr,
                 unsigned long long timeout_ns, long long* __restrict__ step_counter) {
  GS_DP_MARK(0);
  pdl_sync();
  GS_DP_MARK(1);
  const int G = gridDim.x, c = blockIdx.x, tid = threadIdx.x;
  const unsigned int e = *reinterpret_cast<volatile unsigned int*>(&st->epoch) + 1u;
  const int par = static_cast<int>(e & 1u);
  const int n4 = static_cast<int>(n_total >> 2);          // (the host refuses buffers of 2^31 floats or more)
  float4* flat4 = reinterpret_cast<float4*>(flat);
  const int W = peers.world, me = peers.rank;
  const int first = c * kDpThreads + tid;
  const int stride = G * kDpThreads;
  GS_DP_MARK(2);

  // partial-gradient replicas of a segment are folded into the flat buffer (and cleared) before anything reads it
  auto fold = [&](int i, int k, float4 g) -> float4 {
    if (k >= 0 && segs.extra[k] != nullptr) {
      const long long o = 4LL * i - segs.off[k];
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
    for (int i = first; i < n4; i += stride) {
      const int k = seg_of(segs, 4LL * i);
      if (k >= 0 && segs.extra[k] != nullptr) flat4[i] = fold(i, k, flat4[i]);
    }
  if (W > 1) {
    // ---- push my pieces into every peer's slot [par][me]; a word that carries the empty pattern goes out as the
    //      canonical NaN (it is a NaN either way) ----
    for (int i = first; i < n4; i += stride) {
      uint4 v = *reinterpret_cast<const uint4*>(flat4 + i);
      v.x = v.x == kDpEmpty ? 0x7fffffffu : v.x; v.y = v.y == kDpEmpty ? 0x7fffffffu : v.y;
      v.z = v.z == kDpEmpty ? 0x7fffffffu : v.z; v.w = v.w == kDpEmpty ? 0x7fffffffu : v.w;
      for (int p = 0; p < W; ++p)
        if (p != me) st_relaxed_sys_v4(reinterpret_cast<float4*>(peers.recv[p]) + static_cast<long long>(par * W + me) * n4 + i, v);
    }
  }
  GS_DP_MARK(3);

  // ---- one round of loads: gradient pieces (all ranks' copies), their segments, the parameters they update ----
  const float inv_w = 1.0f / static_cast<float>(W);
  const bool in_regs = n4 <= stride;                       // one float4 per thread: nothing goes through memory
  float4 g_keep = make_float4(0.f, 0.f, 0.f, 0.f), w_keep = make_float4(0.f, 0.f, 0.f, 0.f);
  int k_keep = -1;
  float ss[kDpMaxGroups];
#pragma unroll
  for (int g = 0; g < kDpMaxGroups; ++g) ss[g] = 0.f;
  float4* mine_w = reinterpret_cast<float4*>(peers.recv[me]) + static_cast<long long>(par) * W * n4;
  auto reduce_one = [&](int i, int k) -> float4 {
    // the W-1 peer copies of this float4, four ranks at a time: their loads in flight together, stragglers polled one
    // by one; summed in rank order 0..W-1 -- every rank adds the same numbers in the same order
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 own = flat4[i];
#pragma unroll 1
    for (int r0 = 0; r0 < W; r0 += 4) {
      uint4 got[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (r0 + u < W && r0 + u != me) got[u] = ld_relaxed_sys_v4(mine_w + static_cast<long long>(r0 + u) * n4 + i);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u;
        if (r >= W) break;
        float4 v = own;
        if (r != me) {
          float4* slot = mine_w + static_cast<long long>(r) * n4 + i;
          if (!dp_landed(got[u])) got[u] = dp_poll(slot, timeout_ns, &st->status);
          // the slot is empty again for the epoch after next (nobody writes it before this kernel has completed)
          *reinterpret_cast<uint4*>(slot) = make_uint4(kDpEmpty, kDpEmpty, kDpEmpty, kDpEmpty);
          v = make_float4(__uint_as_float(got[u].x), __uint_as_float(got[u].y), __uint_as_float(got[u].z), __uint_as_float(got[u].w));
        }
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
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
#pragma unroll 1
  for (int i = first; i < n4; i += stride) {
    const int k = seg_of(segs, 4LL * i);
    if (in_regs && k >= 0) {                               // the parameter piece: in flight while the peers' copies arrive
      const long long o = 4LL * i - segs.off[k];
      const float* p = segs.param[k] + o;
      if (segs.numel[k] - o >= 4 && (reinterpret_cast<uintptr_t>(p) & 15u) == 0) w_keep = *reinterpret_cast<const float4*>(p);
    }
    const float4 s = reduce_one(i, k);
    add_ss(s, k);
    if (in_regs) {
      g_keep = s;
      k_keep = k;
    } else {
      flat4[i] = s;
    }
  }
  __shared__ float s_red[kDpThreads / 32][kDpMaxGroups];
  __shared__ float s_part[kDpMaxCtas][kDpMaxGroups];
  __shared__ float s_coef[kDpMaxGroups];
#pragma unroll
  for (int g = 0; g < kDpMaxGroups; ++g) {
    const float v = warp_sum(ss[g]);
    if ((tid & 31) == 0) s_red[tid >> 5][g] = v;
  }
  __syncthreads();
  // ---- the per-CTA partial sums meet without a barrier of their own: CTA c publishes {partial, epoch} as ONE 8-byte
  //      word per group (naturally atomic: no fence, no counter), and every CTA polls the G x groups words -- thread t
  //      the word of CTA t / 4, group t % 4 -- until they carry this epoch.  One store and one load round trip instead of
  //      store + fence + atomic, poll, load. ----
  static_assert(kDpThreads == kDpMaxCtas * kDpMaxGroups, "one polling thread per (CTA, group) record");
  if (tid < kDpMaxGroups) {
    float v = 0.f;
    for (int w = 0; w < kDpThreads / 32; ++w) v += s_red[w][tid];
    st_relaxed_gpu_u64(&st->rec[c][tid], (static_cast<unsigned long long>(e) << 32) | __float_as_uint(v));
  }
  GS_DP_MARK(4);
  {
    const int cc = tid / kDpMaxGroups, g = tid % kDpMaxGroups;
    float v = 0.f;
    if (cc < G) {
      unsigned long long rec = ld_relaxed_gpu_u64(&st->rec[cc][g]);
      unsigned long long t0 = 0;
      for (unsigned spins = 0; static_cast<unsigned int>(rec >> 32) != e; ++spins) {
        if ((spins & 255u) == 255u) {
          const unsigned long long now = global_ns();
          if (t0 == 0) t0 = now;
          if (now - t0 > timeout_ns) { atomicExch(&st->status, 2u); break; }
        }
        rec = ld_relaxed_gpu_u64(&st->rec[cc][g]);
      }
      v = __uint_as_float(static_cast<unsigned int>(rec));
    }
    s_part[cc][g] = v;
  }
  __syncthreads();
  GS_DP_MARK(5);
  if (tid < 32) {
    // lane = CTA (G <= 64), then a shuffle tree: the order of the additions is fixed by the lane numbers, hence
    // identical on every CTA and every rank
    float v[kDpMaxGroups];
#pragma unroll
    for (int g = 0; g < kDpMaxGroups; ++g) {
      v[g] = s_part[tid][g] + s_part[tid + 32][g];          // (CTAs beyond G hold zeros)
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
  auto sgd_one = [&](int i, const float4& g, int k, const float4* w_have) {
    if (k >= 0) {
      const float step = lr * s_coef[segs.group[k]];
      const long long o = 4LL * i - segs.off[k];
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
#pragma unroll 1
  for (int i = first; i < n4; i += stride) {
    const float4 g = in_regs ? g_keep : flat4[i];
    const int k = in_regs ? k_keep : seg_of(segs, 4LL * i);
    sgd_one(i, g, k, in_regs ? &w_keep : nullptr);
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
  if (!flat_grad || !state || n_total < 4 || (n_total & 3) || n_total >= (1LL << 31) || !aligned16(flat_grad)) return GS_ERR_BAD_ARG;
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
  // the grid size is a pure function of n_total (every CTA polls the records of all of them)
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
