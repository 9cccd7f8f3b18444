// K3 gather-segment-reduce (MEAN / MAX) and its scatter backward.
//
// Replaces GraphSage.aggregate, src/models.py:300-326: the reference gathers
// embed_matrix = h[U], builds a dense [rows x |U|] 0/1 mask on the CPU, row-normalises it
// and multiplies (MEAN), or loops over rows in Python (MAX).  Here one warp owns one
// destination row: lanes span the feature dimension in 128-bit pieces, the row's (<= 11)
// neighbour ids are held one per lane and broadcast by shuffle, and up to kBatch
// independent 16-byte loads per lane are put in flight before any is consumed.  The kernel
// is HBM-bound: algorithmic bytes per row = cnt*dim*4 (gathered rows) + dim*4 (output)
// + cnt*4 (ids) + 4 (count)   [SURVEY.md §8(d)].
#include <stdlib.h>

#include "common.cuh"

namespace gs {

constexpr int kAggWarps = 8;          // warps (= rows) per CTA
constexpr int kBatch = 12;            // loads in flight per lane; covers fanout 10 + self in one go

// Register-staged forward.  One warp per destination row, lane = float4 column.  The row's count,
// its id list and the live-row counter are three INDEPENDENT loads (one latency), then up to
// kBatch 16-byte row pieces per lane are in flight before the first is consumed.  <= 80 registers
// keep 24 warps (= 24 rows, ~100 KB of gathers) resident per SM; a non-persistent grid lets the
// hardware scheduler overlap the index latency of one wave with the gathers of the previous one.
template <int MODE>
__global__ void __launch_bounds__(kAggWarps * 32, MODE == GS_AGG_MEAN ? 3 : 2)
agg_fwd_kernel(const float* __restrict__ table, uint32_t ld_bytes, int dim4,
               const int32_t* __restrict__ nbr, int stride, const int32_t* __restrict__ cnt,
               const int32_t* __restrict__ num_rows_dev, int max_rows,
               float* __restrict__ out, int64_t ld_out, int32_t* __restrict__ argmax, int64_t ld_arg) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * kAggWarps + (threadIdx.x >> 5);
  if (r >= max_rows) return;
  const int32_t* row_ids = nbr + static_cast<int64_t>(r) * stride;
  const int n_raw = __ldg(cnt + r);
  const int mine_raw = lane < stride ? __ldg(row_ids + lane) : -1;
  if (r >= live_rows(num_rows_dev, max_rows)) return;
  const int n = min(n_raw, stride);
  const float inv = 1.0f / static_cast<float>(n);         // n == 0 -> inf; 0 * inf = NaN as in the reference (0/0)
  const float qnan = __int_as_float(0x7fc00000);
  const char* tbase = reinterpret_cast<const char*>(table);

  for (int cbase = 0; cbase < dim4; cbase += 32) {
    const int c4 = cbase + lane;
    const bool active = c4 < dim4;
    const char* col = tbase + 16 * c4;
    float4 acc = (MODE == GS_AGG_MEAN) ? make_float4(0.f, 0.f, 0.f, 0.f)
                                       : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    int4 arg = make_int4(-1, -1, -1, -1);
    for (int jc = 0; jc < n; jc += 32) {                  // lists longer than a warp: compatibility callers only
      const int mine = jc == 0 ? (lane < n ? mine_raw : -1) : (jc + lane < n ? __ldg(row_ids + jc + lane) : -1);
      const int here = min(32, n - jc);
      for (int j0 = 0; j0 < here; j0 += kBatch) {         // a single pass for fan-out 10 (+ self)
        float4 v[kBatch];
        int id[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int got = __shfl_sync(0xffffffffu, mine, (j0 + u) & 31);
          id[u] = (active && j0 + u < here) ? got : -1;
          const char* src = col + static_cast<size_t>(static_cast<uint32_t>(id[u])) * ld_bytes;
          v[u] = ldg_stream_f4_if<(MODE == GS_AGG_MEAN) ? 0u : 0xff800000u>(src, id[u] >= 0);   // 0 / -inf when off
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          if (MODE == GS_AGG_MEAN) {
            acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
          } else {
            if (v[u].x > acc.x) { acc.x = v[u].x; arg.x = id[u]; }
            if (v[u].y > acc.y) { acc.y = v[u].y; arg.y = id[u]; }
            if (v[u].z > acc.z) { acc.z = v[u].z; arg.z = id[u]; }
            if (v[u].w > acc.w) { acc.w = v[u].w; arg.w = id[u]; }
          }
        }
      }
    }
    if (active) {
      if (MODE == GS_AGG_MEAN) {
        acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
      } else if (n == 0) {
        acc = make_float4(qnan, qnan, qnan, qnan);
      }
      *reinterpret_cast<float4*>(out + static_cast<int64_t>(r) * ld_out + 4 * c4) = acc;
      if (MODE == GS_AGG_MAX && argmax != nullptr)
        *reinterpret_cast<int4*>(argmax + static_cast<int64_t>(r) * ld_arg + 4 * c4) = arg;
    }
  }
}

// Backward.  MEAN: every gathered row receives grad/cnt (vector red.add, 16 B per lane);
// MAX: only the winning row per element.  The self-row gather of src/models.py:265 is the
// same scatter with weight 1, folded in here so layer l's input gradient is one launch.
template <int MODE>
__global__ void __launch_bounds__(kAggWarps * 32)
agg_bwd_kernel(const float* __restrict__ grad_agg, int64_t ld_ga, const float* __restrict__ grad_self, int64_t ld_gs,
               int dim4, const int32_t* __restrict__ nbr, int stride, const int32_t* __restrict__ cnt,
               const int32_t* __restrict__ self_idx, const int32_t* __restrict__ argmax, int64_t ld_arg,
               const int32_t* __restrict__ num_rows_dev, int max_rows, float* __restrict__ grad_table, int64_t ld_gt) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * kAggWarps + (threadIdx.x >> 5);
  if (r >= live_rows(num_rows_dev, max_rows)) return;
  const int n = grad_agg != nullptr ? min(cnt[r], stride) : 0;
  const int32_t* row_ids = nbr + static_cast<int64_t>(r) * stride;
  const float inv = n > 0 ? 1.0f / static_cast<float>(n) : 0.f;
  const int me = (grad_self != nullptr) ? (self_idx != nullptr ? self_idx[r] : r) : -1;

  for (int cbase = 0; cbase < dim4; cbase += 32) {
    const int c4 = cbase + lane;
    const bool active = c4 < dim4;
    if (me >= 0 && active) {
      const float4 g = *reinterpret_cast<const float4*>(grad_self + static_cast<int64_t>(r) * ld_gs + 4 * c4);
      atomicAdd(reinterpret_cast<float4*>(grad_table + static_cast<int64_t>(me) * ld_gt + 4 * c4), g);
    }
    if (grad_agg == nullptr) continue;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) g = *reinterpret_cast<const float4*>(grad_agg + static_cast<int64_t>(r) * ld_ga + 4 * c4);
    if (MODE == GS_AGG_MEAN) {
      g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
      for (int jc = 0; jc < n; jc += 32) {
        const int mine = (jc + lane < n) ? __ldg(row_ids + jc + lane) : -1;
        const int here = min(32, n - jc);
        for (int j = 0; j < here; ++j) {
          const int id = __shfl_sync(0xffffffffu, mine, j);
          if (active && id >= 0)
            atomicAdd(reinterpret_cast<float4*>(grad_table + static_cast<int64_t>(id) * ld_gt + 4 * c4), g);
        }
      }
    } else if (active) {
      const int4 a = *reinterpret_cast<const int4*>(argmax + static_cast<int64_t>(r) * ld_arg + 4 * c4);
      if (a.x >= 0) atomicAdd(grad_table + static_cast<int64_t>(a.x) * ld_gt + 4 * c4 + 0, g.x);
      if (a.y >= 0) atomicAdd(grad_table + static_cast<int64_t>(a.y) * ld_gt + 4 * c4 + 1, g.y);
      if (a.z >= 0) atomicAdd(grad_table + static_cast<int64_t>(a.z) * ld_gt + 4 * c4 + 2, g.z);
      if (a.w >= 0) atomicAdd(grad_table + static_cast<int64_t>(a.w) * ld_gt + 4 * c4 + 3, g.w);
    }
  }
}


// ---------------------------------------------------------------------------------------
// K3 forward, asynchronous-copy pipeline (the default for sampled lists, stride <= 16).
//
// The register-staged kernel above keeps <= 16 warps per SM resident (92 registers) and each
// warp spends most of its life in the dependent chain cnt -> ids -> rows, so only ~1/3 of the
// HBM latency x bandwidth product is ever in flight (measured 1.6 TB/s, ncu r1).  Here every
// warp owns a ring of STAGES shared-memory slots; a slot receives the <= stride gathered rows of
// one destination row by cp.async (LDGSTS, 16 bytes per lane, lane = float4 column, so a warp
// instruction moves one whole 400-512 B row piece, coalesced).  Copies are issued STAGES rows
// ahead of the row being reduced and hold no registers while in flight: 8 warps x 4 slots x
// ~4 KB keep > 100 KB per SM outstanding.  Every lane reduces exactly the bytes it copied, so
// cp.async.wait_group is the only synchronisation.  A persistent grid (one CTA per SM) walks
// the rows round-robin; wide tables are processed in column chunks of 128 floats.
//
// (A first version used one cp.async.bulk (TMA, UBLKCP) per gathered row: 2.2 TB/s.  With 400 B
// copies the per-SM TMA unit, not HBM, was the limit -- ~45 cycles per copy -- so the LSU path
// is used; ncu numbers in profiles/.)
// ---------------------------------------------------------------------------------------
constexpr int kPipeWarps = 16;          // 4 per scheduler: enough to cover each other's instruction latencies
constexpr int kPipeMaxStride = 16;
constexpr int kPipeChunkF4 = 32;            // float4 per column chunk = 512 B per copied row piece

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// cursor over a warp's work items (row, column chunk) without any division
struct ItemCursor {
  int row, chunk;
  __device__ __forceinline__ void advance(int n_chunks, int row_step) {
    if (++chunk == n_chunks) { chunk = 0; row += row_step; }
  }
};

template <int MODE, int STAGES>
__global__ void __launch_bounds__(kPipeWarps * 32, 1)
agg_fwd_pipe_kernel(const float* __restrict__ table, int64_t ld, int dim4, const int32_t* __restrict__ nbr, int stride,
                    const int32_t* __restrict__ cnt, const int32_t* __restrict__ num_rows_dev, int max_rows,
                    float* __restrict__ out, int64_t ld_out, int32_t* __restrict__ argmax, int64_t ld_arg,
                    int slot_bytes) {
  pdl_sync();
  extern __shared__ __align__(128) unsigned char pipe_smem[];
  __shared__ int32_t s_ids[kPipeWarps][STAGES][kPipeMaxStride];
  __shared__ int32_t s_n[kPipeWarps][STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = live_rows(num_rows_dev, max_rows);
  const int total_warps = gridDim.x * kPipeWarps;
  const int gw = blockIdx.x * kPipeWarps + warp;
  const int n_chunks = (dim4 + kPipeChunkF4 - 1) / kPipeChunkF4;
  unsigned char* ring = pipe_smem + static_cast<size_t>(warp) * STAGES * slot_bytes;
  const int piece_stride = slot_bytes / stride;                 // bytes between neighbour pieces inside a slot
  const float qnan = __int_as_float(0x7fc00000);
  const uint32_t lane_dst = smem_addr(ring) + lane * 16;
  const float* lane_src = table + lane * 4;

  // index fetch for the item under cursor c (this lane's neighbour id, -1 beyond the row's count);
  // issued two items ahead of its use so the dependent chain cnt/ids -> row addresses never sits
  // on the critical path
  auto fetch = [&](const ItemCursor& c) -> int {
    int mine = -1;
    if (c.row < rows) {
      const int n = __ldg(cnt + c.row);              // the two loads are independent: one latency, not two
      const int v = lane < stride ? __ldg(nbr + static_cast<int64_t>(c.row) * stride + lane) : -1;
      mine = lane < n ? v : -1;
    }
    return mine;
  };
  auto issue = [&](const ItemCursor& c, int s, int mine) {       // always commits one group (possibly empty)
    if (c.row < rows) {                              // warp-uniform
      const int f4 = min(kPipeChunkF4, dim4 - c.chunk * kPipeChunkF4);
      const int n = __popc(__ballot_sync(0xffffffffu, mine >= 0));   // ids occupy slots [0, n)
      if (lane < kPipeMaxStride) s_ids[warp][s][lane] = mine;
      if (lane == 0) s_n[warp][s] = n;
      __syncwarp();
      if (lane < f4) {
        const uint32_t dst = lane_dst + s * slot_bytes;
        const float* src0 = lane_src + c.chunk * (kPipeChunkF4 * 4);
        const int32_t* ids = s_ids[warp][s];
#pragma unroll 4
        for (int j = 0; j < n; ++j)                  // n is warp-uniform: no divergence, ids by LDS broadcast
          cp_async_16(dst + j * piece_stride, src0 + static_cast<int64_t>(ids[j]) * ld);
      }
    }
    cp_async_commit();
  };

  ItemCursor ci{gw, 0}, cc{gw, 0}, cf{gw, 0};        // issue, consume and index-fetch cursors
  int pre[STAGES + 2];
#pragma unroll
  for (int i = 0; i < STAGES + 2; ++i) {             // all index loads of the pipeline fill fly together
    pre[i] = fetch(cf);
    cf.advance(n_chunks, total_warps);
  }
#pragma unroll
  for (int t = 0; t < STAGES; ++t) {
    issue(ci, t, pre[t]);
    ci.advance(n_chunks, total_warps);
  }
  int ma = pre[STAGES], mb = pre[STAGES + 1];
  int s = 0;
  while (cc.row < rows) {
    const int f4 = min(kPipeChunkF4, dim4 - cc.chunk * kPipeChunkF4);
    cp_async_wait<STAGES - 1>();                    // the oldest outstanding group (this item) has landed
    __syncwarp();                                   // s_ids written by other lanes
    const unsigned char* slot = ring + s * slot_bytes + lane * 16;
    if (lane < f4) {
      float4 acc = (MODE == GS_AGG_MEAN) ? make_float4(0.f, 0.f, 0.f, 0.f)
                                         : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      int4 arg = make_int4(-1, -1, -1, -1);
      const int n = s_n[warp][s];
      const int32_t* ids = s_ids[warp][s];
#pragma unroll 4
      for (int j = 0; j < n; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(slot + j * piece_stride);
        if (MODE == GS_AGG_MEAN) {
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        } else {
          const int id = ids[j];
          if (v.x > acc.x) { acc.x = v.x; arg.x = id; }
          if (v.y > acc.y) { acc.y = v.y; arg.y = id; }
          if (v.z > acc.z) { acc.z = v.z; arg.z = id; }
          if (v.w > acc.w) { acc.w = v.w; arg.w = id; }
        }
      }
      if (MODE == GS_AGG_MEAN) {
        const float inv = 1.0f / static_cast<float>(n);        // n == 0: 0 * inf = NaN, the reference's 0/0
        acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
      } else if (n == 0) {
        acc = make_float4(qnan, qnan, qnan, qnan);
      }
      const int c4 = cc.chunk * kPipeChunkF4 + lane;
      *reinterpret_cast<float4*>(out + static_cast<int64_t>(cc.row) * ld_out + 4 * c4) = acc;
      if (MODE == GS_AGG_MAX && argmax != nullptr)
        *reinterpret_cast<int4*>(argmax + static_cast<int64_t>(cc.row) * ld_arg + 4 * c4) = arg;
    }
    __syncwarp();                                   // every lane is done with the slot before it is refilled
    issue(ci, s, ma);
    ci.advance(n_chunks, total_warps);
    ma = mb;
    mb = fetch(cf);
    cf.advance(n_chunks, total_warps);
    cc.advance(n_chunks, total_warps);
    s = (s + 1 == STAGES) ? 0 : s + 1;
  }
  cp_async_wait<0>();
}

// GS_AGG_IMPL=pipe selects the cp.async ring kernel (kept for A/B measurements); default: register kernel
static int agg_impl() {
  static int impl = -1;
  if (impl < 0) {
    const char* e = getenv("GS_AGG_IMPL");
    impl = (e && e[0] == 'p') ? 1 : 0;
  }
  return impl;
}

template <int MODE, int STAGES>
static int launch_pipe(const float* table, int64_t ld, int dim4, const int32_t* nbr, int stride, const int32_t* cnt,
                       const int32_t* num_rows_dev, int max_rows, float* out, int64_t ld_out, int32_t* argmax,
                       int64_t ld_arg, int slot_bytes, cudaStream_t st) {
  const int smem = kPipeWarps * STAGES * slot_bytes;
  cudaError_t e = cudaFuncSetAttribute(agg_fwd_pipe_kernel<MODE, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return static_cast<int>(e);
  int grid = (max_rows + kPipeWarps - 1) / kPipeWarps;
  if (grid > kNumSMs) grid = kNumSMs;
  launch(agg_fwd_pipe_kernel<MODE, STAGES>, grid, kPipeWarps * 32, smem, st, table, ld, dim4, nbr, stride, cnt, num_rows_dev,
                                                                        max_rows, out, ld_out, argmax, ld_arg, slot_bytes);
  return finish_launch();
}

// ---------------------------------------------------------------------------------------
// K3 forward over a ROW-PARTITIONED bf16 feature table (BASELINE.json configs[4]: 100M nodes,
// 128 bf16 features, 8 shards).  Shard s holds the rows of nodes [s*rows_per_shard,
// (s+1)*rows_per_shard); its base pointer is either local HBM or a CUDA-IPC mapping of a
// peer GPU's HBM, so a gathered row is read directly over NVLink by the 16-byte loads below
// -- no collective, no staging copy.  A 128-feature row is 256 B = 16 lanes
// x 16 B, so a warp fetches TWO neighbour rows per load instruction (one per half-warp) and
// keeps up to kShardBatch of them in flight per lane before reducing in fp32; the peer
// latency (~2-3 us) is covered by 16+ warps/SM x 6 x 512 B outstanding.  The same launch
// converts the destination node's own row to fp32 (`out_self`), because the SageLayer GEMM
// consumes fp32 operands and its gather index cannot cross shards.
// ---------------------------------------------------------------------------------------
constexpr int kMaxShards = 8;
constexpr int kShardBatch = 8;
struct ShardTable {
  const uint16_t* base[kMaxShards];
  int num_shards;
  long long rows_per_shard;
};

__device__ __forceinline__ uint4 ldg_stream_u4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void bf16x8_add(float (&acc)[8], const uint4& v) {
  acc[0] += __uint_as_float(v.x << 16); acc[1] += __uint_as_float(v.x & 0xffff0000u);
  acc[2] += __uint_as_float(v.y << 16); acc[3] += __uint_as_float(v.y & 0xffff0000u);
  acc[4] += __uint_as_float(v.z << 16); acc[5] += __uint_as_float(v.z & 0xffff0000u);
  acc[6] += __uint_as_float(v.w << 16); acc[7] += __uint_as_float(v.w & 0xffff0000u);
}

// LPR = lanes per gathered row (16: two rows per warp instruction, dim <= 128; 32: one row, column loop)
template <int LPR>
__global__ void __launch_bounds__(kAggWarps * 32)
agg_fwd_bf16_sharded_kernel(const ShardTable tab_arg, int64_t ld, int dim8, const int32_t* __restrict__ nbr, int stride,
                            const int32_t* __restrict__ cnt, const int32_t* __restrict__ self_nodes,
                            const int32_t* __restrict__ num_rows_dev, int max_rows, float* __restrict__ out_agg,
                            int64_t ld_agg, float* __restrict__ out_self, int64_t ld_self) {
  pdl_sync();
  __shared__ const uint16_t* s_base[kMaxShards];
  if (threadIdx.x < kMaxShards) s_base[threadIdx.x] = threadIdx.x < tab_arg.num_shards ? tab_arg.base[threadIdx.x] : nullptr;
  __syncthreads();
  constexpr int SUB = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane / LPR, sl = lane % LPR;
  const int r = blockIdx.x * kAggWarps + (threadIdx.x >> 5);
  if (r >= live_rows(num_rows_dev, max_rows)) return;
  const int n = min(__ldg(cnt + r), stride);
  const int mine = lane < n ? __ldg(nbr + static_cast<int64_t>(r) * stride + lane) : -1;      // stride <= 32
  const int me = (out_self != nullptr) ? __ldg(self_nodes + r) : -1;
  const int rps = static_cast<int>(tab_arg.rows_per_shard);   // node ids are int32, so is the shard height
  const float inv = 1.0f / static_cast<float>(n);             // n == 0: 0 * inf = NaN, the reference's 0/0
  auto row_ptr = [&](int id) -> const uint16_t* {
    const int sh = id / rps;
    return s_base[sh] + static_cast<long long>(id - sh * rps) * ld;
  };
  for (int cbase = 0; cbase < dim8; cbase += LPR) {
    const int c8 = cbase + sl;
    const bool active = c8 < dim8;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    uint4 sv = make_uint4(0u, 0u, 0u, 0u);
    const bool self_here = active && me >= 0 && sub == 0;
    if (self_here) sv = ldg_stream_u4(row_ptr(me) + 8 * c8);
    for (int j0 = 0; j0 < n; j0 += SUB * kShardBatch) {
      uint4 v[kShardBatch];
      bool ok[kShardBatch];
#pragma unroll
      for (int u = 0; u < kShardBatch; ++u) {
        const int j = j0 + u * SUB + sub;
        const int id = __shfl_sync(0xffffffffu, mine, j & 31);
        ok[u] = active && j < n && id >= 0;
        if (ok[u]) v[u] = ldg_stream_u4(row_ptr(id) + 8 * c8);
      }
#pragma unroll
      for (int u = 0; u < kShardBatch; ++u)
        if (ok[u]) bf16x8_add(acc, v[u]);
    }
    if (SUB == 2) {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 16);
    }
    if (active) {
      // 8 fp32 outputs per column piece: with two half-warps each writes one float4 of them
      float* dst = out_agg + static_cast<int64_t>(r) * ld_agg + 8 * c8;
      const float4 lo = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
      const float4 hi = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
      if (SUB == 1 || sub == 0) *reinterpret_cast<float4*>(dst) = lo;
      if (SUB == 1 || sub == 1) *reinterpret_cast<float4*>(dst + 4) = hi;
      if (self_here) {
        float s8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) s8[k] = 0.f;
        bf16x8_add(s8, sv);
        float* sd = out_self + static_cast<int64_t>(r) * ld_self + 8 * c8;
        *reinterpret_cast<float4*>(sd) = make_float4(s8[0], s8[1], s8[2], s8[3]);
        *reinterpret_cast<float4*>(sd + 4) = make_float4(s8[4], s8[5], s8[6], s8[7]);
      }
    }
  }
}

}  // namespace gs

using namespace gs;

extern "C" int gs_agg_fwd(const float* table, int64_t ld, int32_t dim, const int32_t* nbr, int32_t stride,
                          const int32_t* cnt, const int32_t* num_rows_dev, int32_t max_rows, int32_t mode,
                          float* out, int64_t ld_out, int32_t* argmax, int64_t ld_arg, gs_stream_t stream) {
  if (!table || !nbr || !cnt || !out || dim < 1 || stride < 1 || max_rows < 0) return GS_ERR_BAD_ARG;
  if (mode != GS_AGG_MEAN && mode != GS_AGG_MAX) return GS_ERR_BAD_ARG;
  const int dim4 = (dim + 3) / 4;
  if ((ld & 3) || (ld_out & 3) || ld < 4 * dim4 || ld_out < 4 * dim4) return GS_ERR_ALIGNMENT;
  if (!aligned16(table) || !aligned16(out)) return GS_ERR_ALIGNMENT;
  if (argmax && ((ld_arg & 3) || ld_arg < 4 * dim4 || !aligned16(argmax))) return GS_ERR_ALIGNMENT;
  if (max_rows == 0) return GS_OK;
  cudaStream_t st = as_stream(stream);
  if (agg_impl() == 1 && stride <= kPipeMaxStride) {
    // asynchronous-copy pipeline: slot = `stride` pieces of one column chunk
    const int f4 = dim4 < kPipeChunkF4 ? dim4 : kPipeChunkF4;
    const int slot_bytes = stride * f4 * 16;
    const int budget = 216 * 1024 / kPipeWarps;
    int32_t* am = (mode == GS_AGG_MAX) ? argmax : nullptr;
    const int64_t la = (mode == GS_AGG_MAX) ? ld_arg : 0;
#define GS_PIPE(MODE_, STAGES_) \
    return launch_pipe<MODE_, STAGES_>(table, ld, dim4, nbr, stride, cnt, num_rows_dev, max_rows, out, ld_out, am, la, slot_bytes, st)
    if (3 * slot_bytes <= budget) { if (mode == GS_AGG_MEAN) GS_PIPE(GS_AGG_MEAN, 3); else GS_PIPE(GS_AGG_MAX, 3); }
    if (2 * slot_bytes <= budget) { if (mode == GS_AGG_MEAN) GS_PIPE(GS_AGG_MEAN, 2); else GS_PIPE(GS_AGG_MAX, 2); }
    if (slot_bytes <= budget) { if (mode == GS_AGG_MEAN) GS_PIPE(GS_AGG_MEAN, 1); else GS_PIPE(GS_AGG_MAX, 1); }
#undef GS_PIPE
  }
  if (ld * 4 > 0xffffffffLL) return GS_ERR_UNSUPPORTED;
  const int blocks = (max_rows + kAggWarps - 1) / kAggWarps;
  const uint32_t ld_bytes = static_cast<uint32_t>(ld * 4);
  if (mode == GS_AGG_MEAN)
    launch(agg_fwd_kernel<GS_AGG_MEAN>, blocks, kAggWarps * 32, 0, st, 
        table, ld_bytes, dim4, nbr, stride, cnt, num_rows_dev, max_rows, out, ld_out, nullptr, 0);
  else
    launch(agg_fwd_kernel<GS_AGG_MAX>, blocks, kAggWarps * 32, 0, st, 
        table, ld_bytes, dim4, nbr, stride, cnt, num_rows_dev, max_rows, out, ld_out, argmax, ld_arg);
  return finish_launch();
}

extern "C" int gs_agg_bwd(const float* grad_agg, int64_t ld_ga, const float* grad_self, int64_t ld_gs, int32_t dim,
                          const int32_t* nbr, int32_t stride, const int32_t* cnt, const int32_t* self_idx,
                          const int32_t* argmax, int64_t ld_arg, const int32_t* num_rows_dev, int32_t max_rows,
                          int32_t mode, float* grad_table, int64_t ld_gt, gs_stream_t stream) {
  if (!grad_table || dim < 1 || max_rows < 0) return GS_ERR_BAD_ARG;
  if (!grad_agg && !grad_self) return GS_ERR_BAD_ARG;
  if (grad_agg && (!nbr || !cnt || stride < 1)) return GS_ERR_BAD_ARG;
  if (mode != GS_AGG_MEAN && mode != GS_AGG_MAX) return GS_ERR_BAD_ARG;
  if (grad_agg && mode == GS_AGG_MAX && !argmax) return GS_ERR_BAD_ARG;
  const int dim4 = (dim + 3) / 4;
  if ((ld_gt & 3) || ld_gt < 4 * dim4 || !aligned16(grad_table)) return GS_ERR_ALIGNMENT;
  if (grad_agg && ((ld_ga & 3) || ld_ga < 4 * dim4 || !aligned16(grad_agg))) return GS_ERR_ALIGNMENT;
  if (grad_self && ((ld_gs & 3) || ld_gs < 4 * dim4 || !aligned16(grad_self))) return GS_ERR_ALIGNMENT;
  if (grad_agg && mode == GS_AGG_MAX && ((ld_arg & 3) || ld_arg < 4 * dim4 || !aligned16(argmax))) return GS_ERR_ALIGNMENT;
  if (max_rows == 0) return GS_OK;
  // a dummy index list keeps the kernel's pointer arithmetic valid when only grad_self is scattered
  const int blocks = (max_rows + kAggWarps - 1) / kAggWarps;
  if (mode == GS_AGG_MEAN)
    launch(agg_bwd_kernel<GS_AGG_MEAN>, blocks, kAggWarps * 32, 0, as_stream(stream), 
        grad_agg, ld_ga, grad_self, ld_gs, dim4, nbr, stride, cnt, self_idx, argmax, ld_arg, num_rows_dev, max_rows,
        grad_table, ld_gt);
  else
    launch(agg_bwd_kernel<GS_AGG_MAX>, blocks, kAggWarps * 32, 0, as_stream(stream), 
        grad_agg, ld_ga, grad_self, ld_gs, dim4, nbr, stride, cnt, self_idx, argmax, ld_arg, num_rows_dev, max_rows,
        grad_table, ld_gt);
  return finish_launch();
}

extern "C" int gs_agg_fwd_bf16_sharded(const void* const* shard_bases_host, int32_t num_shards, int64_t rows_per_shard,
                                       int64_t ld, int32_t dim, const int32_t* nbr, int32_t stride, const int32_t* cnt,
                                       const int32_t* self_nodes, const int32_t* num_rows_dev, int32_t max_rows,
                                       float* out_agg, int64_t ld_agg, float* out_self, int64_t ld_self,
                                       gs_stream_t stream) {
  if (!shard_bases_host || num_shards < 1 || num_shards > kMaxShards || rows_per_shard < 1 || rows_per_shard > 0x7fffffffLL)
    return GS_ERR_BAD_ARG;
  if (!nbr || !cnt || !out_agg || dim < 1 || stride < 1 || stride > 32 || max_rows < 0) return GS_ERR_BAD_ARG;
  if (out_self && !self_nodes) return GS_ERR_BAD_ARG;
  const int dim8 = (dim + 7) / 8;
  if ((ld & 7) || ld < 8 * dim8) return GS_ERR_ALIGNMENT;                 // bf16 rows in 16-byte pieces
  if ((ld_agg & 3) || ld_agg < 8 * dim8 || !aligned16(out_agg)) return GS_ERR_ALIGNMENT;
  if (out_self && ((ld_self & 3) || ld_self < 8 * dim8 || !aligned16(out_self))) return GS_ERR_ALIGNMENT;
  ShardTable tab{};
  tab.num_shards = num_shards;
  tab.rows_per_shard = rows_per_shard;
  for (int s = 0; s < num_shards; ++s) {
    if (!shard_bases_host[s] || !aligned16(shard_bases_host[s])) return GS_ERR_ALIGNMENT;
    tab.base[s] = static_cast<const uint16_t*>(shard_bases_host[s]);
  }
  if (max_rows == 0) return GS_OK;
  const int blocks = (max_rows + kAggWarps - 1) / kAggWarps;
  if (dim8 <= 16)
    launch(agg_fwd_bf16_sharded_kernel<16>, blocks, kAggWarps * 32, 0, as_stream(stream), 
        tab, ld, dim8, nbr, stride, cnt, self_nodes, num_rows_dev, max_rows, out_agg, ld_agg, out_self, ld_self);
  else
    launch(agg_fwd_bf16_sharded_kernel<32>, blocks, kAggWarps * 32, 0, as_stream(stream), 
        tab, ld, dim8, nbr, stride, cnt, self_nodes, num_rows_dev, max_rows, out_agg, ld_agg, out_self, ld_self);
  return finish_launch();
}
