// K6 pair losses of UnsupervisedLoss, src/models.py:65-132, forward and backward.
//
// The reference loops over seeds in Python and issues ~10 tiny torch ops per seed; here one
// warp owns one seed.  Lanes span the embedding dimension in 128-bit pieces; every pair costs
// one row load and one 2-value warp reduction (dot, |v|^2).
//   mode 0 ("normal", get_loss_sage :65-98):
//       score_s = mean_p(-log sig(cos_p)) - Q * mean_n(log sig(-cos_n))
//   mode 1 (margin, get_loss_margin :100-132):
//       score_s = max(0, max_n log sig(cos_n) - min_p log sig(cos_p) + MARGIN)
//   loss = mean over seeds that have at least one positive and one negative pair (:75-76).
// Forward stores d score / d cos for every pair (coef_*); backward turns those into
// embedding gradients with d cos/d u = (v^ - cos u^)/|u|, d cos/d v = (u^ - cos v^)/|v|
// (F.cosine_similarity semantics: each norm clamped at eps = 1e-8).
#include "common.cuh"

namespace gs {

constexpr int kPairWarps = 4;
constexpr float kCosEps = 1e-8f;

__device__ __forceinline__ float log_sigmoid_ref(float x) { return logf(1.0f / (1.0f + expf(-x))); }   // log(sigmoid(x)) as torch evaluates it
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// cos(u, v) with u held in registers (one float4 per lane per 128-column slab)
template <int SLABS>
__device__ __forceinline__ float pair_cos(const float4 (&u)[SLABS], float inv_nu, const float* __restrict__ vrow,
                                          int dim4, int lane, float* inv_nv_out) {
  float dot = 0.f, vv = 0.f;
#pragma unroll
  for (int s = 0; s < SLABS; ++s) {
    const int c4 = s * 32 + lane;
    if (c4 < dim4) {
      const float4 v = *reinterpret_cast<const float4*>(vrow + 4 * c4);
      dot += u[s].x * v.x + u[s].y * v.y + u[s].z * v.z + u[s].w * v.w;
      vv += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
    vv += __shfl_xor_sync(0xffffffffu, vv, o);
  }
  const float inv_nv = 1.0f / fmaxf(sqrtf(vv), kCosEps);
  if (inv_nv_out) *inv_nv_out = inv_nv;
  return dot * inv_nu * inv_nv;
}

template <int SLABS>
__global__ void __launch_bounds__(kPairWarps * 32)
pair_loss_fwd_kernel(const float* __restrict__ emb, int64_t ld, int dim4, const int32_t* __restrict__ seed_idx,
                     int num_seeds, const int32_t* __restrict__ pos_ptr, const int32_t* __restrict__ pos_idx,
                     const int32_t* __restrict__ neg_ptr, const int32_t* __restrict__ neg_idx, int mode, float q,
                     float margin, float* __restrict__ loss_sum, float* __restrict__ coef_pos,
                     float* __restrict__ coef_neg, int32_t* __restrict__ num_active) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * kPairWarps + (threadIdx.x >> 5);
  if (s >= num_seeds) return;
  const int pb = pos_ptr[s], pe = pos_ptr[s + 1], nb = neg_ptr[s], ne = neg_ptr[s + 1];
  // lists are fixed-stride with -1 holes (a walk that produced no pair, fewer far nodes than num_neg)
  int np = 0, nn = 0;
  for (int p = pb + lane; p < pe; p += 32) np += pos_idx[p] >= 0;
  for (int p = nb + lane; p < ne; p += 32) nn += neg_idx[p] >= 0;
  np = __reduce_add_sync(0xffffffffu, np);
  nn = __reduce_add_sync(0xffffffffu, nn);
  if (np == 0 || nn == 0) {                       // skipped seed (:75-76 / :110-111): no gradient
    for (int p = pb + lane; p < pe; p += 32) coef_pos[p] = 0.f;
    for (int p = nb + lane; p < ne; p += 32) coef_neg[p] = 0.f;
    return;
  }
  const float* urow = emb + static_cast<int64_t>(seed_idx[s]) * ld;
  float4 u[SLABS];
  float uu = 0.f;
#pragma unroll
  for (int k = 0; k < SLABS; ++k) {
    const int c4 = k * 32 + lane;
    u[k] = c4 < dim4 ? *reinterpret_cast<const float4*>(urow + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    uu += u[k].x * u[k].x + u[k].y * u[k].y + u[k].z * u[k].z + u[k].w * u[k].w;
  }
  uu = warp_sum(uu);
  const float inv_nu = 1.0f / fmaxf(sqrtf(uu), kCosEps);

  float score;
  if (mode == 0) {
    float pos_sum = 0.f, neg_sum = 0.f;
    for (int p = pb; p < pe; ++p) {
      if (pos_idx[p] < 0) { if (lane == 0) coef_pos[p] = 0.f; continue; }
      const float c = pair_cos<SLABS>(u, inv_nu, emb + static_cast<int64_t>(pos_idx[p]) * ld, dim4, lane, nullptr);
      pos_sum += log_sigmoid_ref(c);
      if (lane == 0) coef_pos[p] = -sigmoidf_(-c) / static_cast<float>(np);        // d(-mean log sig(c))/dc
    }
    for (int p = nb; p < ne; ++p) {
      if (neg_idx[p] < 0) { if (lane == 0) coef_neg[p] = 0.f; continue; }
      const float c = pair_cos<SLABS>(u, inv_nu, emb + static_cast<int64_t>(neg_idx[p]) * ld, dim4, lane, nullptr);
      neg_sum += log_sigmoid_ref(-c);
      if (lane == 0) coef_neg[p] = q * sigmoidf_(c) / static_cast<float>(nn);         // d(-Q mean log sig(-c))/dc
    }
    score = -pos_sum / static_cast<float>(np) - q * neg_sum / static_cast<float>(nn);
  } else {
    float pos_min = INFINITY, neg_max = -INFINITY, c_pos = 0.f, c_neg = 0.f;
    int arg_pos = pb, arg_neg = nb;
    for (int p = pb; p < pe; ++p) {
      if (lane == 0) coef_pos[p] = 0.f;
      if (pos_idx[p] < 0) continue;
      const float c = pair_cos<SLABS>(u, inv_nu, emb + static_cast<int64_t>(pos_idx[p]) * ld, dim4, lane, nullptr);
      const float l = log_sigmoid_ref(c);
      if (l < pos_min) { pos_min = l; arg_pos = p; c_pos = c; }                      // first index on ties
    }
    for (int p = nb; p < ne; ++p) {
      if (lane == 0) coef_neg[p] = 0.f;
      if (neg_idx[p] < 0) continue;
      const float c = pair_cos<SLABS>(u, inv_nu, emb + static_cast<int64_t>(neg_idx[p]) * ld, dim4, lane, nullptr);
      const float l = log_sigmoid_ref(c);
      if (l > neg_max) { neg_max = l; arg_neg = p; c_neg = c; }
    }
    const float raw = neg_max - pos_min + margin;
    score = fmaxf(raw, 0.f);
    __syncwarp();
    if (lane == 0 && raw > 0.f) {                  // torch.max(0, x): gradient to x only when x > 0
      coef_pos[arg_pos] = -sigmoidf_(-c_pos);      // d(-log sig(c))/dc
      coef_neg[arg_neg] = sigmoidf_(-c_neg);       // d(+log sig(c))/dc
    }
  }
  if (lane == 0) {
    atomicAdd(loss_sum, score);
    atomicAdd(num_active, 1);
  }
}

__global__ void pair_loss_finalize_kernel(float* __restrict__ loss, const float* __restrict__ loss_sum,
                                          const int32_t* __restrict__ num_active) {
  pdl_sync();
  loss[0] = loss_sum[0] / static_cast<float>(num_active[0]);   // 0/0 = NaN when every seed was skipped (reference raises)
}

template <int SLABS>
__global__ void __launch_bounds__(kPairWarps * 32)
pair_loss_bwd_kernel(const float* __restrict__ emb, int64_t ld, int dim4, const int32_t* __restrict__ seed_idx,
                     int num_seeds, const int32_t* __restrict__ pos_ptr, const int32_t* __restrict__ pos_idx,
                     const int32_t* __restrict__ neg_ptr, const int32_t* __restrict__ neg_idx,
                     const float* __restrict__ coef_pos, const float* __restrict__ coef_neg,
                     const int32_t* __restrict__ num_active, const float* __restrict__ grad_loss,
                     float* __restrict__ grad_emb, int64_t ld_ge) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * kPairWarps + (threadIdx.x >> 5);
  if (s >= num_seeds) return;
  const int pb = pos_ptr[s], pe = pos_ptr[s + 1], nb = neg_ptr[s], ne = neg_ptr[s + 1];
  const float scale = (grad_loss ? grad_loss[0] : 1.0f) / static_cast<float>(num_active[0]);
  const int urow_i = seed_idx[s];
  const float* urow = emb + static_cast<int64_t>(urow_i) * ld;
  float4 u[SLABS], gu[SLABS];
  float uu = 0.f;
#pragma unroll
  for (int k = 0; k < SLABS; ++k) {
    const int c4 = k * 32 + lane;
    u[k] = c4 < dim4 ? *reinterpret_cast<const float4*>(urow + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    gu[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    uu += u[k].x * u[k].x + u[k].y * u[k].y + u[k].z * u[k].z + u[k].w * u[k].w;
  }
  uu = warp_sum(uu);
  const float inv_nu = 1.0f / fmaxf(sqrtf(uu), kCosEps);
  for (int side = 0; side < 2; ++side) {
    const int b = side ? nb : pb, e = side ? ne : pe;
    const int32_t* idx = side ? neg_idx : pos_idx;
    const float* coef = side ? coef_neg : coef_pos;
    for (int p = b; p < e; ++p) {
      const float w = coef[p] * scale;
      if (w == 0.f) continue;                       // warp-uniform: coef is read by every lane
      const int vrow_i = idx[p];
      const float* vrow = emb + static_cast<int64_t>(vrow_i) * ld;
      float inv_nv;
      const float c = pair_cos<SLABS>(u, inv_nu, vrow, dim4, lane, &inv_nv);
#pragma unroll
      for (int k = 0; k < SLABS; ++k) {
        const int c4 = k * 32 + lane;
        if (c4 >= dim4) continue;
        const float4 v = *reinterpret_cast<const float4*>(vrow + 4 * c4);
        // u^ = u*inv_nu, v^ = v*inv_nv
        float4 dv, du;
        du.x = w * (v.x * inv_nv - c * u[k].x * inv_nu) * inv_nu;
        du.y = w * (v.y * inv_nv - c * u[k].y * inv_nu) * inv_nu;
        du.z = w * (v.z * inv_nv - c * u[k].z * inv_nu) * inv_nu;
        du.w = w * (v.w * inv_nv - c * u[k].w * inv_nu) * inv_nu;
        dv.x = w * (u[k].x * inv_nu - c * v.x * inv_nv) * inv_nv;
        dv.y = w * (u[k].y * inv_nu - c * v.y * inv_nv) * inv_nv;
        dv.z = w * (u[k].z * inv_nu - c * v.z * inv_nv) * inv_nv;
        dv.w = w * (u[k].w * inv_nu - c * v.w * inv_nv) * inv_nv;
        gu[k].x += du.x; gu[k].y += du.y; gu[k].z += du.z; gu[k].w += du.w;
        atomicAdd(reinterpret_cast<float4*>(grad_emb + static_cast<int64_t>(vrow_i) * ld_ge + 4 * c4), dv);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < SLABS; ++k) {
    const int c4 = k * 32 + lane;
    if (c4 < dim4) atomicAdd(reinterpret_cast<float4*>(grad_emb + static_cast<int64_t>(urow_i) * ld_ge + 4 * c4), gu[k]);
  }
}

}  // namespace gs

using namespace gs;

#define GS_DISPATCH_SLABS(dim4, CALL)                  \
  do {                                                 \
    if ((dim4) <= 32) { CALL(1); }                     \
    else if ((dim4) <= 64) { CALL(2); }                \
    else if ((dim4) <= 128) { CALL(4); }               \
    else if ((dim4) <= 256) { CALL(8); }               \
    else return GS_ERR_UNSUPPORTED;                    \
  } while (0)

extern "C" int gs_pair_loss_fwd(const float* emb, int64_t ld, int32_t dim, const int32_t* seed_idx, int32_t num_seeds,
                                const int32_t* pos_ptr, const int32_t* pos_idx, const int32_t* neg_ptr,
                                const int32_t* neg_idx, int32_t mode, float q, float margin, float* loss,
                                float* coef_pos, float* coef_neg, float* loss_sum_scratch, int32_t* num_active,
                                gs_stream_t stream) {
  if (!emb || !seed_idx || !pos_ptr || !pos_idx || !neg_ptr || !neg_idx || !loss || !coef_pos || !coef_neg ||
      !loss_sum_scratch || !num_active)
    return GS_ERR_BAD_ARG;
  if (num_seeds < 1 || dim < 1 || (mode != 0 && mode != 1)) return GS_ERR_BAD_ARG;
  const int dim4 = (dim + 3) / 4;
  if ((ld & 3) || ld < 4 * dim4 || !aligned16(emb)) return GS_ERR_ALIGNMENT;
  cudaStream_t st = as_stream(stream);
  cudaError_t ce = cudaMemsetAsync(loss_sum_scratch, 0, sizeof(float), st);
  if (ce == cudaSuccess) ce = cudaMemsetAsync(num_active, 0, sizeof(int32_t), st);
  if (ce != cudaSuccess) return static_cast<int>(ce);
  const int blocks = (num_seeds + kPairWarps - 1) / kPairWarps;
#define CALL(S)                                                                                                    \
  launch(pair_loss_fwd_kernel<S>, blocks, kPairWarps * 32, 0, st, emb, ld, dim4, seed_idx, num_seeds, pos_ptr, pos_idx, \
                                                              neg_ptr, neg_idx, mode, q, margin, loss_sum_scratch, \
                                                              coef_pos, coef_neg, num_active)
  GS_DISPATCH_SLABS(dim4, CALL);
#undef CALL
  launch(pair_loss_finalize_kernel, 1, 1, 0, st, loss, loss_sum_scratch, num_active);
  return finish_launch(2);
}

extern "C" int gs_pair_loss_bwd(const float* emb, int64_t ld, int32_t dim, const int32_t* seed_idx, int32_t num_seeds,
                                const int32_t* pos_ptr, const int32_t* pos_idx, const int32_t* neg_ptr,
                                const int32_t* neg_idx, const float* coef_pos, const float* coef_neg,
                                const int32_t* num_active, const float* grad_loss, float* grad_emb, int64_t ld_ge,
                                gs_stream_t stream) {
  if (!emb || !seed_idx || !pos_ptr || !pos_idx || !neg_ptr || !neg_idx || !coef_pos || !coef_neg || !num_active ||
      !grad_emb)
    return GS_ERR_BAD_ARG;
  if (num_seeds < 1 || dim < 1) return GS_ERR_BAD_ARG;
  const int dim4 = (dim + 3) / 4;
  if ((ld & 3) || ld < 4 * dim4 || !aligned16(emb)) return GS_ERR_ALIGNMENT;
  if ((ld_ge & 3) || ld_ge < 4 * dim4 || !aligned16(grad_emb)) return GS_ERR_ALIGNMENT;
  cudaStream_t st = as_stream(stream);
  const int blocks = (num_seeds + kPairWarps - 1) / kPairWarps;
#define CALL(S)                                                                                                    \
  launch(pair_loss_bwd_kernel<S>, blocks, kPairWarps * 32, 0, st, emb, ld, dim4, seed_idx, num_seeds, pos_ptr, pos_idx, \
                                                              neg_ptr, neg_idx, coef_pos, coef_neg, num_active,    \
                                                              grad_loss, grad_emb, ld_ge)
  GS_DISPATCH_SLABS(dim4, CALL);
#undef CALL
  return finish_launch();
}
