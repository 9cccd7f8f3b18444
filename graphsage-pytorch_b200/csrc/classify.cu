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

}  // namespace gs

using namespace gs;

extern "C" int gs_cls_nll_fwd_bwd(const float* emb, int64_t ld_emb, int32_t rows, int32_t dim,
                                  const float* weight, const float* bias, int32_t num_classes,
                                  const int64_t* labels, const int32_t* label_index,
                                  float* logp, float* loss, float* grad_emb, int64_t ld_ge,
                                  float* grad_w, float* grad_b, float* scratch, int32_t precision, gs_stream_t stream) {
  if (!emb || !weight || !labels || !logp || !loss || !scratch || rows < 1 || dim < 1 || num_classes < 1)
    return GS_ERR_BAD_ARG;
  cudaStream_t st = as_stream(stream);
  cudaError_t ce = cudaMemsetAsync(loss, 0, sizeof(float), st);
  if (ce != cudaSuccess) return static_cast<int>(ce);
  int e = gs_sage_gemm_fwd(nullptr, 0, nullptr, emb, ld_emb, dim, weight, dim, num_classes, /*gcn=*/1, nullptr, rows,
                           logp, num_classes, /*relu=*/0, precision, stream);
  if (e) return e;
  softmax_nll_kernel<<<(rows + kRowWarps - 1) / kRowWarps, kRowWarps * 32, (num_classes + 1) * sizeof(float), st>>>(
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
  bias_logsoftmax_kernel<<<(rows + kRowWarps - 1) / kRowWarps, kRowWarps * 32, 0, as_stream(stream)>>>(
      logp, bias, rows, num_classes, num_classes);
  return finish_launch();
}

extern "C" int gs_cls_bwd(const float* grad_logp, const float* logp, const float* emb, int64_t ld_emb,
                          int32_t rows, int32_t dim, const float* weight, int32_t num_classes,
                          float* grad_emb, int64_t ld_ge, float* grad_w, float* grad_b, float* scratch,
                          int32_t precision, gs_stream_t stream) {
  if (!grad_logp || !logp || !emb || !weight || !scratch || rows < 0 || dim < 1 || num_classes < 1) return GS_ERR_BAD_ARG;
  if (rows == 0) return GS_OK;
  logsoftmax_bwd_kernel<<<(rows + kRowWarps - 1) / kRowWarps, kRowWarps * 32, num_classes * sizeof(float),
                          as_stream(stream)>>>(grad_logp, logp, rows, num_classes, num_classes, scratch, num_classes,
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
  nll_kernel<<<blocks, 256, 0, as_stream(stream)>>>(logp, labels, label_index, rows, num_classes, loss, grad_logp);
  return finish_launch();
}
