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
#include "common.cuh"

namespace gs {

constexpr int kAggWarps = 8;          // warps (= rows) per CTA
constexpr int kBatch = 12;            // loads in flight per lane; covers fanout 10 + self in one go

template <int MODE>
__global__ void __launch_bounds__(kAggWarps * 32)
agg_fwd_kernel(const float* __restrict__ table, int64_t ld, int dim4,
               const int32_t* __restrict__ nbr, int stride, const int32_t* __restrict__ cnt,
               const int32_t* __restrict__ num_rows_dev, int max_rows,
               float* __restrict__ out, int64_t ld_out, int32_t* __restrict__ argmax, int64_t ld_arg) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * kAggWarps + (threadIdx.x >> 5);
  if (r >= live_rows(num_rows_dev, max_rows)) return;
  const int n = min(cnt[r], stride);
  const int32_t* row_ids = nbr + static_cast<int64_t>(r) * stride;
  const float inv = 1.0f / static_cast<float>(n);         // n == 0 -> inf; 0 * inf = NaN as in the reference (0/0)
  const float qnan = __int_as_float(0x7fc00000);

  for (int cbase = 0; cbase < dim4; cbase += 32) {
    const int c4 = cbase + lane;
    const bool active = c4 < dim4;
    float4 acc = (MODE == GS_AGG_MEAN) ? make_float4(0.f, 0.f, 0.f, 0.f)
                                       : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    int4 arg = make_int4(-1, -1, -1, -1);
    for (int jc = 0; jc < n; jc += 32) {
      const int mine = (jc + lane < n) ? __ldg(row_ids + jc + lane) : -1;
      const int here = min(32, n - jc);
      for (int j0 = 0; j0 < here; j0 += kBatch) {
        float4 v[kBatch];
        int id[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          id[u] = __shfl_sync(0xffffffffu, mine, (j0 + u) & 31);
          const bool ok = active && (j0 + u < here) && id[u] >= 0;
          if (!ok) id[u] = -1;
          if (ok) v[u] = ldg_stream_f4(table + static_cast<int64_t>(id[u]) * ld + 4 * c4);
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          if (id[u] >= 0) {
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
  const int blocks = (max_rows + kAggWarps - 1) / kAggWarps;
  if (mode == GS_AGG_MEAN)
    agg_fwd_kernel<GS_AGG_MEAN><<<blocks, kAggWarps * 32, 0, as_stream(stream)>>>(
        table, ld, dim4, nbr, stride, cnt, num_rows_dev, max_rows, out, ld_out, nullptr, 0);
  else
    agg_fwd_kernel<GS_AGG_MAX><<<blocks, kAggWarps * 32, 0, as_stream(stream)>>>(
        table, ld, dim4, nbr, stride, cnt, num_rows_dev, max_rows, out, ld_out, argmax, ld_arg);
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
    agg_bwd_kernel<GS_AGG_MEAN><<<blocks, kAggWarps * 32, 0, as_stream(stream)>>>(
        grad_agg, ld_ga, grad_self, ld_gs, dim4, nbr, stride, cnt, self_idx, argmax, ld_arg, num_rows_dev, max_rows,
        grad_table, ld_gt);
  else
    agg_bwd_kernel<GS_AGG_MAX><<<blocks, kAggWarps * 32, 0, as_stream(stream)>>>(
        grad_agg, ld_ga, grad_self, ld_gs, dim4, nbr, stride, cnt, self_idx, argmax, ld_arg, num_rows_dev, max_rows,
        grad_table, ld_gt);
  return finish_launch();
}
