// Classification head (src/models.py:25-27) and the supervised loss of src/utils.py:162-163.
//   fwd : logits = emb . W^T (K4 forward kernel, gcn-style single operand, no ReLU), then
//         one warp per row adds the bias and applies log_softmax in place.
//   bwd : one warp per row turns grad_logp into grad_logits (log_softmax backward) and
//         column-sums it into grad_b; grad_W and grad_emb are the K4 backward kernels.
#include <algorithm>
#include "common.cuh"

namespace gs {

constexpr int kRowWarps = 8;

__global__ void __launch_bounds__(kRowWarps * 32)
bias_logsoftmax_kernel(float* __restrict__ logp, const float* __restrict__ bias, int rows, int classes, int64_t ld) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (r >= rows) return;
  float* row = logp + static_cast<int64_t>(r) * ld;
  float mx = -INFINITY;
  for (int c = lane; c < classes; c += 32) {
    const float v = row[c] + (bias ? bias[c] : 0.f);
    row[c] = v;
    mx = fmaxf(mx, v);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int c = lane; c < classes; c += 32) sum += expf(row[c] - mx);
  sum = warp_sum(sum);
  const float lse = mx + logf(sum);
  for (int c = lane; c < classes; c += 32) row[c] -= lse;
}

// grad_logits[r,c] = g[r,c] - exp(logp[r,c]) * sum_c' g[r,c'];  grad_b[c] += sum_r grad_logits[r,c]
__global__ void __launch_bounds__(kRowWarps * 32)
logsoftmax_bwd_kernel(const float* __restrict__ grad_logp, const float* __restrict__ logp, int rows, int classes,
                      int64_t ld, float* __restrict__ grad_logits, int64_t ld_gl, float* __restrict__ grad_b) {
  pdl_sync();
  extern __shared__ float s_db[];
  for (int c = threadIdx.x; c < classes; c += blockDim.x) s_db[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (r < rows) {
    const float* g = grad_logp + static_cast<int64_t>(r) * ld;
    const float* lp = logp + static_cast<int64_t>(r) * ld;
    float s = 0.f;
    for (int c = lane; c < classes; c += 32) s += g[c];
    s = warp_sum(s);
    for (int c = lane; c < classes; c += 32) {
      const float d = g[c] - expf(lp[c]) * s;
      grad_logits[static_cast<int64_t>(r) * ld_gl + c] = d;
      if (grad_b) atomicAdd(&s_db[c], d);
    }
  }
  __syncthreads();
  if (grad_b)
    for (int c = threadIdx.x; c < classes; c += blockDim.x) atomicAdd(&grad_b[c], s_db[c]);
}

__global__ void __launch_bounds__(256)
nll_kernel(const float* __restrict__ logp, const int64_t* __restrict__ labels, const int32_t* __restrict__ label_index,
           int rows, int classes, float* __restrict__ loss, float* __restrict__ grad_logp) {
  pdl_sync();
  const float inv = 1.0f / static_cast<float>(rows);
  float part = 0.f;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
    const int y = static_cast<int>(labels[label_index ? label_index[r] : r]);
    part -= logp[static_cast<int64_t>(r) * classes + y];
    if (grad_logp) {
      float* g = grad_logp + static_cast<int64_t>(r) * classes;
      for (int c = 0; c < classes; ++c) g[c] = (c == y) ? -inv : 0.f;
    }
  }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) atomicAdd(loss, part * inv);
}

// Fused tail of the supervised step: bias + log_softmax + NLL(mean) + d(loss)/d(logits) + grad_b
// in one pass over the logits (one warp per row).  d logits = (softmax - onehot(y)) / rows.
__global__ void __launch_bounds__(kRowWarps * 32)
softmax_nll_kernel(float* __restrict__ logp, const float* __restrict__ bias, const int64_t* __restrict__ labels,
                   const int32_t* __restrict__ label_index, int rows, int classes, float* __restrict__ loss,
                   float* __restrict__ dlogits, float* __restrict__ grad_b) {
  pdl_sync();
  extern __shared__ float s_db[];            // [classes] + 1 (loss partial)
  for (int c = threadIdx.x; c <= classes; c += blockDim.x) s_db[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (r < rows) {
    float* row = logp + static_cast<int64_t>(r) * classes;
    float* drow = dlogits + static_cast<int64_t>(r) * classes;
    const int y = static_cast<int>(labels[label_index ? label_index[r] : r]);
    const float inv = 1.0f / static_cast<float>(rows);
    float mx = -INFINITY;
    for (int c = lane; c < classes; c += 32) {
      const float v = row[c] + (bias ? bias[c] : 0.f);
      row[c] = v;
      mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int c = lane; c < classes; c += 32) sum += expf(row[c] - mx);
    sum = warp_sum(sum);
    const float lse = mx + logf(sum);
    for (int c = lane; c < classes; c += 32) {
      const float lp = row[c] - lse;
      row[c] = lp;
      const float d = (expf(lp) - (c == y ? 1.f : 0.f)) * inv;
      drow[c] = d;
      if (grad_b) atomicAdd(&s_db[c], d);
      if (c == y) atomicAdd(&s_db[classes], -lp * inv);
    }
  }
  __syncthreads();
  if (grad_b)
    for (int c = threadIdx.x; c < classes; c += blockDim.x) atomicAdd(&grad_b[c], s_db[c]);
  if (threadIdx.x == 0) atomicAdd(loss, s_db[classes]);
}


// ---------------------------------------------------------------------------------------
// The whole supervised tail in ONE launch (classes <= 64, dim <= 256, dim % 4 == 0):
//   logits = emb . W^T + b -> log_softmax -> NLL(mean) -> d logits -> grad_b, grad_emb, grad_W.
// The problem is tiny (1024 x 128 x 47 at the bench config: 19 MFMA in total) and was four
// launches of ~10 us each; here a CTA takes 16 rows (one warp per row), keeps W and W^T in
// shared memory, and only the [classes x dim] grad_W partial leaves the CTA through vector REDs.
// Arithmetic is plain fp32 FMA in ascending-k order (the 1e-5 parity mode).
// ---------------------------------------------------------------------------------------
constexpr int kClsRows = 16;                 // rows (= warps) per CTA
constexpr int kClsMaxC = 64;
constexpr int kClsMaxD = 256;

__global__ void __launch_bounds__(kClsRows * 32)
cls_fused_kernel(const float* __restrict__ emb, int64_t ld_emb, int rows, int dim, const float* __restrict__ weight,
                 const float* __restrict__ bias, int classes, const int64_t* __restrict__ labels,
                 const int32_t* __restrict__ label_index, float* __restrict__ logp, float* __restrict__ loss,
                 float* __restrict__ grad_emb, int64_t ld_ge, float* __restrict__ grad_w, float* __restrict__ grad_b,
                 int mask_relu, const int32_t* __restrict__ num_rows_dev) {
  pdl_sync();
  rows = live_rows(num_rows_dev, rows);          // a batch extended on the device: its size is only known here
  extern __shared__ __align__(16) float cls_smem[];
  const int cp = kClsMaxC + 1;                               // W^T row stride: lanes read consecutive classes
  float* w_s = cls_smem;                                     // [classes][dim]
  float* wt_s = w_s + classes * dim;                         // [dim][cp], zero beyond `classes`
  float* e_s = wt_s + dim * cp;                              // [kClsRows][dim]
  float* d_s = e_s + kClsRows * dim;                         // [kClsRows][kClsMaxC]
  float* red_s = d_s + kClsRows * kClsMaxC;                  // [kClsRows] loss partials
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = blockIdx.x * kClsRows;

  // the label sits behind two dependent global loads (index, then label): start them before the staging
  const int r = r0 + warp;
  const bool live = r < rows;
  int y = -1;
  if (live) y = static_cast<int>(labels[label_index ? label_index[r] : r]);
  // W^T columns >= classes are never initialised: the lanes that read them are masked by selects below
  const int d4 = dim >> 2;
  for (int i = tid; i < classes * d4; i += blockDim.x) {
    const int c = i / d4, k = (i - c * d4) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(weight + static_cast<int64_t>(c) * dim + k));
    *reinterpret_cast<float4*>(w_s + c * dim + k) = v;
    wt_s[(k + 0) * cp + c] = v.x; wt_s[(k + 1) * cp + c] = v.y; wt_s[(k + 2) * cp + c] = v.z; wt_s[(k + 3) * cp + c] = v.w;
  }
  for (int k = lane * 4; k < dim; k += 128) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) v = *reinterpret_cast<const float4*>(emb + static_cast<int64_t>(r) * ld_emb + k);
    *reinterpret_cast<float4*>(e_s + warp * dim + k) = v;
  }
  __syncthreads();

  // ---- logits of my row: lane owns classes `lane` and `lane + 32` ----
  const float* e = e_s + warp * dim;
  float a0 = 0.f, a1 = 0.f;
  {
    // dim % 4 == 0: four independent partial sums per class shorten the FMA dependency chain; they are
    // combined pairwise, so the result differs from a serial dot product only by fp32 rounding order
    float p0[4] = {0.f, 0.f, 0.f, 0.f}, p1[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < dim; k += 4) {
      const float4 x = *reinterpret_cast<const float4*>(e + k);
      const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        p0[j] = fmaf(xs[j], wt_s[(k + j) * cp + lane], p0[j]);
        p1[j] = fmaf(xs[j], wt_s[(k + j) * cp + lane + 32], p1[j]);
      }
    }
    a0 = (p0[0] + p0[1]) + (p0[2] + p0[3]);
    a1 = (p1[0] + p1[1]) + (p1[2] + p1[3]);
  }
  const bool v0 = lane < classes, v1 = lane + 32 < classes;
  if (bias) { if (v0) a0 += bias[lane]; if (v1) a1 += bias[lane + 32]; }
  const float mx = warp_max(fmaxf(v0 ? a0 : -INFINITY, v1 ? a1 : -INFINITY));
  const float sum = warp_sum((v0 ? expf(a0 - mx) : 0.f) + (v1 ? expf(a1 - mx) : 0.f));
  const float lse = mx + logf(sum);
  const float lp0 = a0 - lse, lp1 = a1 - lse;
  float dl0 = 0.f, dl1 = 0.f, lpart = 0.f;
  if (live) {
    const float inv = 1.0f / static_cast<float>(rows);
    if (v0) { dl0 = (expf(lp0) - (lane == y ? 1.f : 0.f)) * inv; if (lane == y) lpart = -lp0 * inv; }
    if (v1) { dl1 = (expf(lp1) - (lane + 32 == y ? 1.f : 0.f)) * inv; if (lane + 32 == y) lpart = -lp1 * inv; }
    if (logp) {
      if (v0) logp[static_cast<int64_t>(r) * classes + lane] = lp0;
      if (v1) logp[static_cast<int64_t>(r) * classes + lane + 32] = lp1;
    }
  }
  d_s[warp * kClsMaxC + lane] = dl0;
  d_s[warp * kClsMaxC + lane + 32] = dl1;
  lpart = warp_sum(lpart);
  if (lane == 0) red_s[warp] = lpart;
  __syncwarp();

  // ---- grad_emb[r, :] = d logits . W : lane owns 4 consecutive k ----
  if (grad_emb && live) {
    const float* dr = d_s + warp * kClsMaxC;
    for (int k = lane * 4; k < dim; k += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = 0; c < classes; ++c) {
        const float d = dr[c];
        const float4 w = *reinterpret_cast<const float4*>(w_s + c * dim + k);
        acc.x = fmaf(d, w.x, acc.x); acc.y = fmaf(d, w.y, acc.y); acc.z = fmaf(d, w.z, acc.z); acc.w = fmaf(d, w.w, acc.w);
      }
      if (mask_relu) {        // emb is a ReLU output (src/models.py:219): hand back d(pre-activation) directly
        const float4 x = *reinterpret_cast<const float4*>(e + k);
        acc.x = x.x > 0.f ? acc.x : 0.f; acc.y = x.y > 0.f ? acc.y : 0.f;
        acc.z = x.z > 0.f ? acc.z : 0.f; acc.w = x.w > 0.f ? acc.w : 0.f;
      }
      *reinterpret_cast<float4*>(grad_emb + static_cast<int64_t>(r) * ld_ge + k) = acc;
    }
  }
  __syncthreads();

  // ---- CTA partials: loss, grad_b[c] = sum_r d[r,c], grad_W[c,k] = sum_r d[r,c] emb[r,k] ----
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < kClsRows; ++w) t += red_s[w];
    atomicAdd(loss, t);
  }
  if (grad_b && tid < classes) {
    float t = 0.f;
    for (int w = 0; w < kClsRows; ++w) t += d_s[w * kClsMaxC + tid];
    atomicAdd(grad_b + tid, t);
  }
  if (grad_w) {
    for (int i = tid; i < classes * d4; i += blockDim.x) {
      const int c = i / d4, k = (i - c * d4) * 4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int w = 0; w < kClsRows; ++w) {
        const float d = d_s[w * kClsMaxC + c];
        const float4 x = *reinterpret_cast<const float4*>(e_s + w * dim + k);
        acc.x = fmaf(d, x.x, acc.x); acc.y = fmaf(d, x.y, acc.y); acc.z = fmaf(d, x.z, acc.z); acc.w = fmaf(d, x.w, acc.w);
      }
      atomicAdd(reinterpret_cast<float4*>(grad_w + static_cast<int64_t>(c) * dim + k), acc);
    }
  }
}

static size_t cls_fused_smem(int dim, int classes) {
  return sizeof(float) * (static_cast<size_t>(classes) * dim + static_cast<size_t>(dim) * (kClsMaxC + 1) +
                          static_cast<size_t>(kClsRows) * dim + kClsRows * kClsMaxC + kClsRows);
}

}  // namespace gs

using namespace gs;

extern "C" int gs_cls_nll_fwd_bwd(const float* emb, int64_t ld_emb, int32_t rows, int32_t dim,
                                  const float* weight, const float* bias, int32_t num_classes,
                                  const int64_t* labels, const int32_t* label_index,
                                  float* logp, float* loss, float* grad_emb, int64_t ld_ge,
                                  float* grad_w, float* grad_b, float* scratch, int32_t mask_relu_input,
                                  int32_t zero_loss, const int32_t* num_rows_dev, int32_t precision, gs_stream_t stream) {
  if (!emb || !weight || !labels || !logp || !loss || !scratch || rows < 1 || dim < 1 || num_classes < 1)
    return GS_ERR_BAD_ARG;
  cudaStream_t st = as_stream(stream);
  cudaError_t ce = cudaSuccess;
  if (zero_loss) ce = cudaMemsetAsync(loss, 0, sizeof(float), st);     // 0: the caller zeroed it off the critical path
  if (ce != cudaSuccess) return static_cast<int>(ce);
  // small heads (every configuration of the reference: 3..47 classes, 128 features): one fused launch
  const bool fused_ok = num_classes <= kClsMaxC && dim <= kClsMaxD && (dim & 3) == 0 && (ld_emb & 3) == 0 &&
                        aligned16(emb) && aligned16(weight) && (!grad_emb || ((ld_ge & 3) == 0 && aligned16(grad_emb))) &&
                        (!grad_w || aligned16(grad_w));
  if (fused_ok) {
    const size_t smem = cls_fused_smem(dim, num_classes);
    ce = cudaFuncSetAttribute(cls_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (ce != cudaSuccess) return static_cast<int>(ce);
    launch(cls_fused_kernel, (rows + kClsRows - 1) / kClsRows, kClsRows * 32, smem, st, 
        emb, ld_emb, rows, dim, weight, bias, num_classes, labels, label_index, logp, loss, grad_emb, ld_ge, grad_w,
        grad_b, mask_relu_input, num_rows_dev);
    return finish_launch();
  }
  if (num_rows_dev) return GS_ERR_UNSUPPORTED;       // device-resident row counts: one-launch heads only
  int e = gs_sage_gemm_fwd(nullptr, 0, nullptr, emb, ld_emb, dim, weight, dim, num_classes, /*gcn=*/1, nullptr, rows,
                           logp, num_classes, /*relu=*/0, precision, stream);
  if (e) return e;
  launch(softmax_nll_kernel, (rows + kRowWarps - 1) / kRowWarps, kRowWarps * 32, (num_classes + 1) * sizeof(float), st, 
      logp, bias, labels, label_index, rows, num_classes, loss, scratch, grad_b);
  e = finish_launch();
  if (e) return e;
  if (grad_w) {
    e = gs_sage_gemm_bwd_w(nullptr, 0, nullptr, emb, ld_emb, dim, scratch, num_classes, nullptr, 0, num_classes,
                           /*gcn=*/1, /*relu=*/0, nullptr, rows, grad_w, dim, precision, stream);
    if (e) return e;
  }
  if (grad_emb) {
    e = gs_sage_gemm_bwd_x(scratch, num_classes, nullptr, 0, weight, dim, dim, num_classes, /*gcn=*/1, /*relu=*/0,
                           nullptr, rows, nullptr, 0, grad_emb, ld_ge, precision, stream);
    if (e) return e;
    if (mask_relu_input) {
      e = gs_relu_bwd_inplace(grad_emb, ld_ge, emb, ld_emb, dim, nullptr, rows, stream);
      if (e) return e;
    }
  }
  return GS_OK;
}

extern "C" int gs_cls_fwd(const float* emb, int64_t ld_emb, int32_t rows, int32_t dim,
                          const float* weight, const float* bias, int32_t num_classes,
                          float* logp, int32_t precision, gs_stream_t stream) {
  if (!emb || !weight || !logp || rows < 0 || dim < 1 || num_classes < 1) return GS_ERR_BAD_ARG;
  if (rows == 0) return GS_OK;
  int e = gs_sage_gemm_fwd(nullptr, 0, nullptr, emb, ld_emb, dim, weight, dim, num_classes, /*gcn=*/1, nullptr, rows,
                           logp, num_classes, /*relu=*/0, precision, stream);
  if (e) return e;
  launch(bias_logsoftmax_kernel, (rows + kRowWarps - 1) / kRowWarps, kRowWarps * 32, 0, as_stream(stream), 
      logp, bias, rows, num_classes, num_classes);
  return finish_launch();
}

extern "C" int gs_cls_bwd(const float* grad_logp, const float* logp, const float* emb, int64_t ld_emb,
                          int32_t rows, int32_t dim, const float* weight, int32_t num_classes,
                          float* grad_emb, int64_t ld_ge, float* grad_w, float* grad_b, float* scratch,
                          int32_t precision, gs_stream_t stream) {
  if (!grad_logp || !logp || !emb || !weight || !scratch || rows < 0 || dim < 1 || num_classes < 1) return GS_ERR_BAD_ARG;
  if (rows == 0) return GS_OK;
  launch(logsoftmax_bwd_kernel, (rows + kRowWarps - 1) / kRowWarps, kRowWarps * 32, num_classes * sizeof(float),
                          as_stream(stream), grad_logp, logp, rows, num_classes, num_classes, scratch, num_classes,
                                               grad_b);
  int e = finish_launch();
  if (e) return e;
  if (grad_w) {
    e = gs_sage_gemm_bwd_w(nullptr, 0, nullptr, emb, ld_emb, dim, scratch, num_classes, nullptr, 0, num_classes,
                           /*gcn=*/1, /*relu=*/0, nullptr, rows, grad_w, dim, precision, stream);
    if (e) return e;
  }
  if (grad_emb) {
    e = gs_sage_gemm_bwd_x(scratch, num_classes, nullptr, 0, weight, dim, dim, num_classes, /*gcn=*/1, /*relu=*/0,
                           nullptr, rows, nullptr, 0, grad_emb, ld_ge, precision, stream);
    if (e) return e;
  }
  return GS_OK;
}

extern "C" int gs_nll_fwd_bwd(const float* logp, const int64_t* labels, const int32_t* label_index, int32_t rows,
                              int32_t num_classes, float* loss, float* grad_logp, gs_stream_t stream) {
  if (!logp || !labels || !loss || rows < 1 || num_classes < 1) return GS_ERR_BAD_ARG;
  cudaError_t ce = cudaMemsetAsync(loss, 0, sizeof(float), as_stream(stream));
  if (ce != cudaSuccess) return static_cast<int>(ce);
  const int blocks = std::min((rows + 255) / 256, kNumSMs);
  launch(nll_kernel, blocks, 256, 0, as_stream(stream), logp, labels, label_index, rows, num_classes, loss, grad_logp);
  return finish_launch();
}
