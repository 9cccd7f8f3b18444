// Update step of src/utils.py:185-187: clip_grad_norm_(model.parameters(), max_norm) per
// model, then SGD (no momentum, no weight decay).  Two launches per model: a sum-of-squares
// reduction over all its gradient tensors, then the fused scale + update (+ optional
// zeroing of the gradient for the next step, src/utils.py:189-191).
#include "common.cuh"

namespace gs {

__global__ void __launch_bounds__(256)
grad_sqnorm_kernel(float* const* __restrict__ grads, const int64_t* __restrict__ numels, float grad_div,
                   float* __restrict__ sqnorm) {
  pdl_sync();
  const float* g = grads[blockIdx.y];
  const int64_t n = numels[blockIdx.y];
  const float inv = 1.0f / grad_div;
  float part = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = g[i] * inv;
    part = fmaf(v, v, part);
  }
  part = warp_sum(part);
  __shared__ float s_part[8];
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? s_part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0 && v != 0.f) atomicAdd(sqnorm, v);
  }
}

__global__ void __launch_bounds__(256)
clip_sgd_kernel(float* const* __restrict__ params, float* const* __restrict__ grads,
                const int64_t* __restrict__ numels, float max_norm, float lr, float grad_div,
                const float* __restrict__ sqnorm, int zero_grads) {
  pdl_sync();
  float* p = params[blockIdx.y];
  float* g = grads[blockIdx.y];
  const int64_t n = numels[blockIdx.y];
  // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to <= 1
  float coef = 1.0f;
  if (max_norm > 0.f) {
    const float total = sqrtf(*sqnorm);
    coef = fminf(max_norm / (total + 1e-6f), 1.0f);
  }
  const float step = lr * coef / grad_div;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    p[i] = fmaf(-step, g[i], p[i]);
    if (zero_grads) g[i] = 0.f;
  }
}

}  // namespace gs

using namespace gs;

extern "C" int gs_clip_sgd(float* const* params, float* const* grads, const int64_t* numels, int32_t num_tensors,
                           int64_t max_numel, float max_norm, float lr, float grad_div, int32_t zero_grads,
                           float* norm_scratch, gs_stream_t stream) {
  if (!params || !grads || !numels || !norm_scratch || num_tensors < 1 || max_numel < 1 || grad_div == 0.f)
    return GS_ERR_BAD_ARG;
  cudaStream_t st = as_stream(stream);
  int launches = 0;
  int bx = static_cast<int>((max_numel + 255) / 256);
  if (bx > 2 * kNumSMs) bx = 2 * kNumSMs;
  dim3 grid(bx, num_tensors);
  if (max_norm > 0.f) {
    cudaError_t ce = cudaMemsetAsync(norm_scratch, 0, sizeof(float), st);
    if (ce != cudaSuccess) return static_cast<int>(ce);
    launch(grad_sqnorm_kernel, grid, 256, 0, st, grads, numels, grad_div, norm_scratch);
    ++launches;
  }
  launch(clip_sgd_kernel, grid, 256, 0, st, params, grads, numels, max_norm, lr, grad_div, norm_scratch, zero_grads);
  ++launches;
  return finish_launch(launches);
}
