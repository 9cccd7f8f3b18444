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

constexpr int kAggWarps = 8;          // warps per CTA
#ifndef GS_AGG_BATCH
#define GS_AGG_BATCH 6
#endif
#ifndef GS_AGG_CTAS
#define GS_AGG_CTAS 5
#endif
constexpr int kBatch = GS_AGG_BATCH;  // 16-byte loads in flight per lane: fan-out 10 (+ self) goes in two batches
constexpr int kAggCtasPerSM = GS_AGG_CTAS;   // MEAN: 48 registers -> 5 CTAs = 40 warps per SM

// Register-staged forward, persistent warps.  A warp owns destination rows gw, gw+W, ... (W = warps
// in the grid); lane = float4 column.  The count and id list of a row are loaded one row AHEAD, and
// for the first row before the live-row counter is known, so the dependent chain
// ids -> row addresses is paid once per warp, not once per row.  Per row, kBatch independent 16-byte
// loads per lane are in flight before the first is consumed.  What matters on B200 is resident warps,
// not loads per warp: at 48 registers 40 warps/SM keep ~96 KB of gathers outstanding per SM (measured,
// scratch/agg_probe.cu: 12 loads/lane at 24 warps/SM = 0.47 of the HBM copy peak at b_sz 1024 and 0.73
// at 88K rows; 6 loads/lane at 40 warps/SM = 0.565 and 0.82).  L2 prefetch of looked-ahead rows
// (prefetch.global.L2) and cp.async rings were measured slower than this.  With a grid that covers
// every row the same code is the one-row-per-warp kernel (GS_AGG_GRID=rows).
__device__ __forceinline__ float4 tf32_lo4(const float4& v) {      // x - trunc_tf32(x): the low half of the 3-term split
  return make_float4(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u), v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u),
                     v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u), v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u));
}

// XOUT: the kernel writes the SageLayer's whole input row  X[r] = [ table[self_nodes[r]] | agg[r] ]  (self part at
// column 0, aggregate at column agg_off; src/models.py:265 + :260 -- the operands torch.cat joins at :217) and, when
// out_lo is given, the low halves x - trunc_tf32(x) of both parts in a second buffer of the same layout.  The layer-1
// GEMMs of a train step then read dense operands by TMA instead of gathering 400-byte rows of the feature table and
// splitting them on the critical chain; the self row is loaded with the first gather batch.
template <int MODE, bool XOUT>
__global__ void __launch_bounds__(kAggWarps * 32, MODE == GS_AGG_MEAN ? (XOUT ? kAggCtasPerSM - 1 : kAggCtasPerSM) : 3)
agg_fwd_kernel(const float* __restrict__ table, uint32_t ld_bytes, int dim4,
               const int32_t* __restrict__ nbr, int stride, const int32_t* __restrict__ cnt,
               const int32_t* __restrict__ num_rows_dev, int max_rows,
               float* __restrict__ out, int64_t ld_out, int32_t* __restrict__ argmax, int64_t ld_arg,
               const int32_t* __restrict__ self_nodes, int agg_off, float* __restrict__ out_lo) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int W = gridDim.x * kAggWarps;
  const int gw = blockIdx.x * kAggWarps + (threadIdx.x >> 5);
  const float qnan = __int_as_float(0x7fc00000);
  const char* tbase = reinterpret_cast<const char*>(table);

  int n_next = 0, v_next = -1, s_next = -1;       // count, this lane's id and the node id of the row the warp handles next
  if (gw < max_rows) {
    n_next = __ldg(cnt + gw);
    if (lane < stride) v_next = __ldg(nbr + static_cast<int64_t>(gw) * stride + lane);
    if (XOUT && self_nodes != nullptr) s_next = __ldg(self_nodes + gw);
  }
  const int rows = live_rows(num_rows_dev, max_rows);
  for (int r = gw; r < rows; r += W) {
    const int32_t* row_ids = nbr + static_cast<int64_t>(r) * stride;
    const int n = min(n_next, stride);
    const int mine_raw = v_next;
    const int self_id = s_next;
    if (r + W < max_rows) {                       // next row's indices fly beside this row's gathers
      n_next = __ldg(cnt + r + W);
      if (lane < stride) v_next = __ldg(row_ids + static_cast<int64_t>(W) * stride + lane);
      if (XOUT && self_nodes != nullptr) s_next = __ldg(self_nodes + r + W);
    }
    const float inv = 1.0f / static_cast<float>(n);         // n == 0 -> inf; 0 * inf = NaN as in the reference (0/0)
    for (int cbase = 0; cbase < dim4; cbase += 32) {
      const int c4 = cbase + lane;
      const bool active = c4 < dim4;
      const char* col = tbase + 16 * c4;
      float4 acc = (MODE == GS_AGG_MEAN) ? make_float4(0.f, 0.f, 0.f, 0.f)
                                         : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      int4 arg = make_int4(-1, -1, -1, -1);
      bool self_todo = XOUT && self_id >= 0;
      const int64_t xoff = static_cast<int64_t>(r) * ld_out + 4 * c4;      // addresses are formed at the stores (registers)
      for (int jc = 0; jc < n; jc += 32) {                  // lists longer than a warp: compatibility callers only
        const int mine = jc == 0 ? (lane < n ? mine_raw : -1) : (jc + lane < n ? __ldg(row_ids + jc + lane) : -1);
        const int here = min(32, n - jc);
        for (int j0 = 0; j0 < here; j0 += kBatch) {         // two passes for fan-out 10
          float4 v[kBatch];
          int id[kBatch];
#pragma unroll
          for (int u = 0; u < kBatch; ++u) {
            const int got = __shfl_sync(0xffffffffu, mine, (j0 + u) & 31);
            id[u] = (active && j0 + u < here) ? got : -1;
            const char* src = col + static_cast<size_t>(static_cast<uint32_t>(id[u])) * ld_bytes;
            v[u] = ldg_stream_f4_if<(MODE == GS_AGG_MEAN) ? 0u : 0xff800000u>(src, id[u] >= 0);   // 0 / -inf when off
          }
          float4 sv;
          const bool self_now = XOUT && self_todo;          // the self row flies with the first batch
          if (self_now)
            sv = ldg_stream_f4_if<0u>(col + static_cast<size_t>(static_cast<uint32_t>(self_id)) * ld_bytes, active);
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
          if (self_now) {
            if (active) {
              *reinterpret_cast<float4*>(out + xoff) = sv;
              if (out_lo != nullptr) *reinterpret_cast<float4*>(out_lo + xoff) = tf32_lo4(sv);
            }
            self_todo = false;
          }
        }
      }
      if (XOUT && self_todo && active) {           // a row without neighbours still has its self part
        const float4 sv = ldg_stream_f4(reinterpret_cast<const float*>(col + static_cast<size_t>(static_cast<uint32_t>(self_id)) * ld_bytes));
        *reinterpret_cast<float4*>(out + xoff) = sv;
        if (out_lo != nullptr) *reinterpret_cast<float4*>(out_lo + xoff) = tf32_lo4(sv);
      }
      if (active) {
        if (MODE == GS_AGG_MEAN) {
          acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
        } else if (n == 0) {
          acc = make_float4(qnan, qnan, qnan, qnan);
        }
        *reinterpret_cast<float4*>(out + xoff + (XOUT ? agg_off : 0)) = acc;
        if (XOUT && out_lo != nullptr) *reinterpret_cast<float4*>(out_lo + xoff + agg_off) = tf32_lo4(acc);
        if (MODE == GS_AGG_MAX && argmax != nullptr)
          *reinterpret_cast<int4*>(argmax + static_cast<int64_t>(r) * ld_arg + 4 * c4) = arg;
      }
    }
  }
}

// Backward.  MEAN: every gathered row receives grad/cnt (vector red.add, 16 B per lane);
// MAX: only the winning row per element.  The self-row gather of src/models.py:265 is the
// same scatter with weight 1, folded in here so layer l's input gradient is one launch.
// mask_table (nullable) = the ReLU output h the scattered-into rows came from (src/models.py:219):
// a contribution to element (u, c) is dropped when h[u, c] <= 0.  The mask depends only on the
// destination element, so masking every contribution equals masking the sum -- the separate
// d(relu) pass over grad_table (one launch on the critical path of the step) disappears.
__device__ __forceinline__ void scatter_add4(float* __restrict__ grad_table, int64_t ld_gt, const float* __restrict__ mask_table,
                                             int64_t ld_mask, int row, int c4, float4 g) {
  if (mask_table != nullptr) {
    const float4 h = __ldg(reinterpret_cast<const float4*>(mask_table + static_cast<int64_t>(row) * ld_mask + 4 * c4));
    g.x = h.x > 0.f ? g.x : 0.f; g.y = h.y > 0.f ? g.y : 0.f; g.z = h.z > 0.f ? g.z : 0.f; g.w = h.w > 0.f ? g.w : 0.f;
  }
  atomicAdd(reinterpret_cast<float4*>(grad_table + static_cast<int64_t>(row) * ld_gt + 4 * c4), g);
}

template <int MODE>
__global__ void __launch_bounds__(kAggWarps * 32)
agg_bwd_kernel(const float* __restrict__ grad_agg, int64_t ld_ga, const float* __restrict__ grad_self, int64_t ld_gs,
               int dim4, const int32_t* __restrict__ nbr, int stride, const int32_t* __restrict__ cnt,
               const int32_t* __restrict__ self_idx, const int32_t* __restrict__ argmax, int64_t ld_arg,
               const int32_t* __restrict__ num_rows_dev, int max_rows, float* __restrict__ grad_table, int64_t ld_gt,
               const float* __restrict__ mask_table, int64_t ld_mask) {
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
      scatter_add4(grad_table, ld_gt, mask_table, ld_mask, me, c4, g);
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
          if (active && id >= 0) scatter_add4(grad_table, ld_gt, mask_table, ld_mask, id, c4, g);
        }
      }
    } else if (active) {
      const int4 a = *reinterpret_cast<const int4*>(argmax + static_cast<int64_t>(r) * ld_arg + 4 * c4);
      auto on = [&](int row, int j) -> bool {      // the winning row's ReLU output decides (MAX: one row per element)
        return row >= 0 && (mask_table == nullptr || __ldg(mask_table + static_cast<int64_t>(row) * ld_mask + 4 * c4 + j) > 0.f);
      };
      if (on(a.x, 0)) atomicAdd(grad_table + static_cast<int64_t>(a.x) * ld_gt + 4 * c4 + 0, g.x);
      if (on(a.y, 1)) atomicAdd(grad_table + static_cast<int64_t>(a.y) * ld_gt + 4 * c4 + 1, g.y);
      if (on(a.z, 2)) atomicAdd(grad_table + static_cast<int64_t>(a.z) * ld_gt + 4 * c4 + 2, g.z);
      if (on(a.w, 3)) atomicAdd(grad_table + static_cast<int64_t>(a.w) * ld_gt + 4 * c4 + 3, g.w);
    }
  }
}


// ---------------------------------------------------------------------------------------
// K3 forward over a ROW-PARTITIONED bf16 feature table (BASELINE.json configs[4]: 100M nodes,
// 128 bf16 features, 8 shards).  Shard s holds the rows of nodes [s*rows_per_shard,
// (s+1)*rows_per_shard); its base pointer is either local HBM or a CUDA-IPC mapping of a
// peer GPU's HBM, so a gathered row is read directly over NVLink by the 16-byte loads below
// -- no collective, no staging copy.  A 128-feature row is 256 B = 16 lanes
// x 16 B, so a warp fetches TWO neighbour rows per load instruction (one per half-warp) and
// keeps up to kShardBatch of them in flight per lane before reducing in fp32; the peer
// latency (~2-3 us) is covered by 32 warps/SM (64 registers) x 5 x 512 B outstanding.  The same launch
// converts the destination node's own row to fp32 (`out_self`), because the SageLayer GEMM
// consumes fp32 operands and its gather index cannot cross shards.
// ---------------------------------------------------------------------------------------
constexpr int kMaxShards = 8;
constexpr int kShardBatch = 5;      // x2 rows per instruction at 16 lanes per row: fan-out 10 in one batch
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

__device__ __forceinline__ void bf16x8_max(float (&acc)[8], const uint4& v) {
  acc[0] = fmaxf(acc[0], __uint_as_float(v.x << 16)); acc[1] = fmaxf(acc[1], __uint_as_float(v.x & 0xffff0000u));
  acc[2] = fmaxf(acc[2], __uint_as_float(v.y << 16)); acc[3] = fmaxf(acc[3], __uint_as_float(v.y & 0xffff0000u));
  acc[4] = fmaxf(acc[4], __uint_as_float(v.z << 16)); acc[5] = fmaxf(acc[5], __uint_as_float(v.z & 0xffff0000u));
  acc[6] = fmaxf(acc[6], __uint_as_float(v.w << 16)); acc[7] = fmaxf(acc[7], __uint_as_float(v.w & 0xffff0000u));
}

// LPR = lanes per gathered row (16: two rows per warp instruction, dim <= 128; 32: one row, column loop).
// MODE: GS_AGG_MEAN (src/models.py:311-314) or GS_AGG_MAX (:316-326; no argmax: the raw features take no gradient).
template <int LPR, int MODE>
__global__ void __launch_bounds__(kAggWarps * 32, 4)
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
    for (int k = 0; k < 8; ++k) acc[k] = MODE == GS_AGG_MEAN ? 0.f : -INFINITY;
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
        if (ok[u]) { if (MODE == GS_AGG_MEAN) bf16x8_add(acc, v[u]); else bf16x8_max(acc, v[u]); }
    }
    if (SUB == 2) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float o = __shfl_xor_sync(0xffffffffu, acc[k], 16);
        acc[k] = MODE == GS_AGG_MEAN ? acc[k] + o : fmaxf(acc[k], o);
      }
    }
    if (active) {
      // 8 fp32 outputs per column piece: with two half-warps each writes one float4 of them
      float* dst = out_agg + static_cast<int64_t>(r) * ld_agg + 8 * c8;
      const float sc = MODE == GS_AGG_MEAN ? inv : (n == 0 ? __int_as_float(0x7fc00000) : 1.0f);   // MAX of nothing: NaN
      const float4 lo = make_float4(acc[0] * sc, acc[1] * sc, acc[2] * sc, acc[3] * sc);
      const float4 hi = make_float4(acc[4] * sc, acc[5] * sc, acc[6] * sc, acc[7] * sc);
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

// CTAs per SM of the persistent forward grid (gs_set_agg_ctas; 0 = as many as fit).  A trainer that runs
// the aggregation in a branch BESIDE its critical chain lowers it: at full occupancy the persistent CTAs
// hold every register of every SM until the kernel ends, and the (larger) GEMM / classifier CTAs of the
// other branch wait for the whole kernel (measured: 30 us of stall per step, profiles/r1_timeline_*).
static int g_agg_ctas = 0;

// GS_AGG_GRID=rows launches one warp per row (kept for A/B measurements); default: persistent grid
static bool agg_grid_persistent() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GS_AGG_GRID");
    v = (e && e[0] == 'r') ? 0 : 1;
  }
  return v == 1;
}

}  // namespace gs

using namespace gs;

extern "C" void gs_set_agg_ctas(int32_t ctas_per_sm) { g_agg_ctas = ctas_per_sm > 0 ? ctas_per_sm : 0; }

extern "C" int gs_agg_fwd_x(const float* table, int64_t ld, int32_t dim, const int32_t* nbr, int32_t stride,
                            const int32_t* cnt, const int32_t* self_nodes, const int32_t* num_rows_dev, int32_t max_rows,
                            int32_t mode, float* out_x, int64_t ld_x, int32_t agg_off, float* out_x_lo,
                            gs_stream_t stream) {
  if (!table || !nbr || !cnt || !out_x || dim < 1 || stride < 1 || max_rows < 0 || agg_off < 0) return GS_ERR_BAD_ARG;
  if (mode != GS_AGG_MEAN && mode != GS_AGG_MAX) return GS_ERR_BAD_ARG;
  const int dim4 = (dim + 3) / 4;
  if (self_nodes && agg_off < 4 * dim4) return GS_ERR_BAD_ARG;            // the self part occupies columns [0, 4*dim4)
  if ((ld & 3) || (ld_x & 3) || (agg_off & 3) || ld < 4 * dim4 || ld_x < agg_off + 4 * dim4) return GS_ERR_ALIGNMENT;
  if (!aligned16(table) || !aligned16(out_x) || (out_x_lo && !aligned16(out_x_lo))) return GS_ERR_ALIGNMENT;
  if (max_rows == 0) return GS_OK;
  if (ld * 4 > 0xffffffffLL) return GS_ERR_UNSUPPORTED;
  int blocks = (max_rows + kAggWarps - 1) / kAggWarps;
  int per_sm = mode == GS_AGG_MEAN ? kAggCtasPerSM - 1 : 3;       // the X variant: 64 registers, 4 CTAs = 32 warps per SM
  if (g_agg_ctas > 0 && g_agg_ctas < per_sm) per_sm = g_agg_ctas;
  const int persistent = kNumSMs * per_sm;
  if (agg_grid_persistent() && blocks > persistent) blocks = persistent;
  const uint32_t ld_bytes = static_cast<uint32_t>(ld * 4);
  cudaStream_t st = as_stream(stream);
  if (mode == GS_AGG_MEAN) {
    set_kernel_carveout(reinterpret_cast<const void*>(agg_fwd_kernel<GS_AGG_MEAN, true>), background_launches());
    launch(agg_fwd_kernel<GS_AGG_MEAN, true>, blocks, kAggWarps * 32, 0, st, table, ld_bytes, dim4, nbr, stride, cnt,
           num_rows_dev, max_rows, out_x, ld_x, nullptr, 0, self_nodes, agg_off, out_x_lo);
  } else {
    set_kernel_carveout(reinterpret_cast<const void*>(agg_fwd_kernel<GS_AGG_MAX, true>), background_launches());
    launch(agg_fwd_kernel<GS_AGG_MAX, true>, blocks, kAggWarps * 32, 0, st, table, ld_bytes, dim4, nbr, stride, cnt,
           num_rows_dev, max_rows, out_x, ld_x, nullptr, 0, self_nodes, agg_off, out_x_lo);
  }
  return finish_launch();
}

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
  if (ld * 4 > 0xffffffffLL) return GS_ERR_UNSUPPORTED;
  int blocks = (max_rows + kAggWarps - 1) / kAggWarps;
  int per_sm = mode == GS_AGG_MEAN ? kAggCtasPerSM : 3;
  if (g_agg_ctas > 0 && g_agg_ctas < per_sm) per_sm = g_agg_ctas;
  const int persistent = kNumSMs * per_sm;
  if (agg_grid_persistent() && blocks > persistent) blocks = persistent;
  const uint32_t ld_bytes = static_cast<uint32_t>(ld * 4);
  set_kernel_carveout(mode == GS_AGG_MEAN ? reinterpret_cast<const void*>(agg_fwd_kernel<GS_AGG_MEAN, false>)
                                          : reinterpret_cast<const void*>(agg_fwd_kernel<GS_AGG_MAX, false>),
                      background_launches());
  if (mode == GS_AGG_MEAN)
    launch(agg_fwd_kernel<GS_AGG_MEAN, false>, blocks, kAggWarps * 32, 0, st,
        table, ld_bytes, dim4, nbr, stride, cnt, num_rows_dev, max_rows, out, ld_out, nullptr, 0, nullptr, 0, nullptr);
  else
    launch(agg_fwd_kernel<GS_AGG_MAX, false>, blocks, kAggWarps * 32, 0, st,
        table, ld_bytes, dim4, nbr, stride, cnt, num_rows_dev, max_rows, out, ld_out, argmax, ld_arg, nullptr, 0, nullptr);
  return finish_launch();
}

extern "C" int gs_agg_bwd(const float* grad_agg, int64_t ld_ga, const float* grad_self, int64_t ld_gs, int32_t dim,
                          const int32_t* nbr, int32_t stride, const int32_t* cnt, const int32_t* self_idx,
                          const int32_t* argmax, int64_t ld_arg, const int32_t* num_rows_dev, int32_t max_rows,
                          int32_t mode, float* grad_table, int64_t ld_gt, const float* mask_table, int64_t ld_mask,
                          gs_stream_t stream) {
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
  if (mask_table && ((ld_mask & 3) || ld_mask < 4 * dim4 || !aligned16(mask_table))) return GS_ERR_ALIGNMENT;
  if (max_rows == 0) return GS_OK;
  // a dummy index list keeps the kernel's pointer arithmetic valid when only grad_self is scattered
  const int blocks = (max_rows + kAggWarps - 1) / kAggWarps;
  if (mode == GS_AGG_MEAN)
    launch(agg_bwd_kernel<GS_AGG_MEAN>, blocks, kAggWarps * 32, 0, as_stream(stream), 
        grad_agg, ld_ga, grad_self, ld_gs, dim4, nbr, stride, cnt, self_idx, argmax, ld_arg, num_rows_dev, max_rows,
        grad_table, ld_gt, mask_table, ld_mask);
  else
    launch(agg_bwd_kernel<GS_AGG_MAX>, blocks, kAggWarps * 32, 0, as_stream(stream), 
        grad_agg, ld_ga, grad_self, ld_gs, dim4, nbr, stride, cnt, self_idx, argmax, ld_arg, num_rows_dev, max_rows,
        grad_table, ld_gt, mask_table, ld_mask);
  return finish_launch();
}

extern "C" int gs_agg_fwd_bf16_sharded(const void* const* shard_bases_host, int32_t num_shards, int64_t rows_per_shard,
                                       int64_t ld, int32_t dim, const int32_t* nbr, int32_t stride, const int32_t* cnt,
                                       const int32_t* self_nodes, const int32_t* num_rows_dev, int32_t max_rows,
                                       float* out_agg, int64_t ld_agg, float* out_self, int64_t ld_self,
                                       int32_t mode, gs_stream_t stream) {
  if (mode != GS_AGG_MEAN && mode != GS_AGG_MAX) return GS_ERR_BAD_ARG;
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
#define GS_SHARD_LAUNCH(LPR, MODE)                                                                                   \
  do {                                                                                                               \
    set_kernel_carveout(reinterpret_cast<const void*>(agg_fwd_bf16_sharded_kernel<LPR, MODE>), background_launches()); \
    launch(agg_fwd_bf16_sharded_kernel<LPR, MODE>, blocks, kAggWarps * 32, 0, as_stream(stream), tab, ld, dim8, nbr,   \
           stride, cnt, self_nodes, num_rows_dev, max_rows, out_agg, ld_agg, out_self, ld_self);                     \
  } while (0)
  if (dim8 <= 16) { if (mode == GS_AGG_MEAN) GS_SHARD_LAUNCH(16, GS_AGG_MEAN); else GS_SHARD_LAUNCH(16, GS_AGG_MAX); }
  else { if (mode == GS_AGG_MEAN) GS_SHARD_LAUNCH(32, GS_AGG_MEAN); else GS_SHARD_LAUNCH(32, GS_AGG_MAX); }
#undef GS_SHARD_LAUNCH
  return finish_launch();
}
