// K4 SageLayer GEMM, fp32 SIMT path (GS_PREC_FP32): the 1e-5 parity mode.
//
// Replaces SageLayer.forward, src/models.py:215-219:  out = relu(W . cat[self, agg]^T)^T.
// The concat is never materialised: the K loop walks a *virtual* operand
//   X[r, :] = [ self_table[self_idx[r], 0:Dp) | agg[r, 0:Dp) ]        Dp = round_up(D, 4)
// (gcn: X = agg only), reading the self rows straight from the previous layer's table
// through the index list (the `pre_hidden_embs[nb]` gather of :265) and the aggregate from
// K3's output.  Feature tables are stored with rows padded to Dp floats (zero filled), so
// every operand load is a 128-bit load; the weight keeps its native [H x 2D] layout and a
// virtual column kv maps to weight column kv (self part) or kv - Dp + D (agg part).
//
// Three kernels: forward (X.W^T, ReLU epilogue), bwd_x (dZ.W), bwd_w (dZ^T.X, split over
// row chunks, accumulated with fp32 atomics).  dZ = grad_out * (out > 0) is applied while
// the tile is loaded, so the ReLU backward never touches HBM on its own.
// The tcgen05 tensor-core path lives in sage_gemm_tc.cu.
#include <stdlib.h>

#include "common.cuh"

namespace gs {

constexpr int BM = 64, BN = 64, BK = 16, kGemmThreads = 256, kPad = 4;

struct XOperand {
  const float* self_table; int64_t ld_self; const int32_t* self_idx;
  const float* agg; int64_t ld_agg;
  int dim, dim_pad, gcn;
  __device__ __forceinline__ int kv_total() const { return gcn ? dim_pad : 2 * dim_pad; }
  // weight column of virtual column kv, or -1 for a padding column
  __device__ __forceinline__ int wcol(int kv) const {
    if (kv < dim) return kv;
    if (gcn) return -1;
    const int k2 = kv - dim_pad;
    return (k2 >= 0 && k2 < dim) ? dim + k2 : -1;
  }
  __device__ __forceinline__ float4 load4(int r, int self_row, int kv) const {
    if (gcn) return *reinterpret_cast<const float4*>(agg + static_cast<int64_t>(r) * ld_agg + kv);
    if (kv < dim_pad) return *reinterpret_cast<const float4*>(self_table + static_cast<int64_t>(self_row) * ld_self + kv);
    return *reinterpret_cast<const float4*>(agg + static_cast<int64_t>(r) * ld_agg + (kv - dim_pad));
  }
};

__device__ __forceinline__ void fma_tile(float (&acc)[4][4], const float* __restrict__ a_col, const float* __restrict__ b_col) {
  const float4 a = *reinterpret_cast<const float4*>(a_col);
  const float4 b = *reinterpret_cast<const float4*>(b_col);
  const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
}

// ---------------------------------------------------------------------------------------
// forward: out[r,h] = act( sum_kv X[r,kv] * W[h, wcol(kv)] )
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGemmThreads)
sage_fwd_kernel(XOperand x, const float* __restrict__ weight, int64_t ldw, int out_dim,
                const int32_t* __restrict__ num_rows_dev, int max_rows, float* __restrict__ out, int64_t ld_out, int relu) {
  pdl_sync();
  __shared__ __align__(16) float Xs[BK][BM + kPad];
  __shared__ __align__(16) float Ws[BK][BN + kPad];
  const int rows = live_rows(num_rows_dev, max_rows);
  const int row0 = blockIdx.x * BM, col0 = blockIdx.y * BN;
  if (row0 >= rows) return;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  // loader roles: X tile = 64 rows x 4 float4; W tile = 64 cols(h) x 16 k scalars
  const int lx_row = tid >> 2, lx_kq = (tid & 3) * 4;
  const int xr = row0 + lx_row;
  const bool xr_ok = xr < rows;
  const int self_row = (xr_ok && !x.gcn) ? (x.self_idx ? x.self_idx[xr] : xr) : 0;
  const int lw_h = tid >> 2, lw_kq = (tid & 3) * 4;
  const int wh = col0 + lw_h;
  float acc[4][4] = {};
  const int kt = x.kv_total();
  float4 xv;
  float wv[4];
  auto fetch = [&](int k0) {                       // global -> registers for the tile starting at k0
    xv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (xr_ok && k0 + lx_kq < kt) xv = x.load4(xr, self_row, k0 + lx_kq);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kv = k0 + lw_kq + i;
      const int wc = kv < kt ? x.wcol(kv) : -1;
      wv[i] = (wc >= 0 && wh < out_dim) ? __ldg(weight + static_cast<int64_t>(wh) * ldw + wc) : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < kt; k0 += BK) {
    __syncthreads();
    Xs[lx_kq + 0][lx_row] = xv.x; Xs[lx_kq + 1][lx_row] = xv.y; Xs[lx_kq + 2][lx_row] = xv.z; Xs[lx_kq + 3][lx_row] = xv.w;
#pragma unroll
    for (int i = 0; i < 4; ++i) Ws[lw_kq + i][lw_h] = wv[i];
    __syncthreads();
    if (k0 + BK < kt) fetch(k0 + BK);              // next tile's loads fly while this one is multiplied
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) fma_tile(acc, &Xs[kk][ty * 4], &Ws[kk][tx * 4]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + ty * 4 + i;
    if (r >= rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int h = col0 + tx * 4 + j;
      if (h < out_dim) out[static_cast<int64_t>(r) * ld_out + h] = relu ? fmaxf(acc[i][j], 0.f) : acc[i][j];
    }
  }
}

// ---------------------------------------------------------------------------------------
// bwd_x: dX[r, c] = sum_h dZ[r,h] W[h,c]   (c over the native 2D / D weight columns)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGemmThreads)
sage_bwd_x_kernel(const float* __restrict__ grad_out, int64_t ld_go, const float* __restrict__ out, int64_t ld_out,
                  const float* __restrict__ weight, int64_t ldw, int dim, int out_dim, int gcn, int relu,
                  const int32_t* __restrict__ num_rows_dev, int max_rows,
                  float* __restrict__ grad_self, int64_t ld_gs, float* __restrict__ grad_agg, int64_t ld_ga) {
  pdl_sync();
  __shared__ __align__(16) float As[BK][BM + kPad];   // dZ^T tile: [h][row]
  __shared__ __align__(16) float Bs[BK][BN + kPad];   // W tile:    [h][col]
  const int rows = live_rows(num_rows_dev, max_rows);
  const int row0 = blockIdx.x * BM, col0 = blockIdx.y * BN;
  if (row0 >= rows) return;
  const int ncols = gcn ? dim : 2 * dim;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int la_row = tid >> 2, la_hq = (tid & 3) * 4;          // dZ: 64 rows x 16 h
  const int lb_h = tid >> 4, lb_c = (tid & 15) * 4;            // W : 16 h x 64 cols
  const int ar = row0 + la_row;
  float acc[4][4] = {};
  float av[4], bv[4];
  auto fetch = [&](int h0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int h = h0 + la_hq + i;
      float g = 0.f;
      if (ar < rows && h < out_dim) {
        g = grad_out[static_cast<int64_t>(ar) * ld_go + h];
        if (relu && !(out[static_cast<int64_t>(ar) * ld_out + h] > 0.f)) g = 0.f;
      }
      av[i] = g;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int h = h0 + lb_h, c = col0 + lb_c + j;
      bv[j] = (h < out_dim && c < ncols) ? __ldg(weight + static_cast<int64_t>(h) * ldw + c) : 0.f;
    }
  };
  fetch(0);
  for (int h0 = 0; h0 < out_dim; h0 += BK) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) As[la_hq + i][la_row] = av[i];
#pragma unroll
    for (int j = 0; j < 4; ++j) Bs[lb_h][lb_c + j] = bv[j];
    __syncthreads();
    if (h0 + BK < out_dim) fetch(h0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) fma_tile(acc, &As[kk][ty * 4], &Bs[kk][tx * 4]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + ty * 4 + i;
    if (r >= rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col0 + tx * 4 + j;
      if (c >= ncols) continue;
      if (gcn) grad_agg[static_cast<int64_t>(r) * ld_ga + c] = acc[i][j];
      else if (c < dim) grad_self[static_cast<int64_t>(r) * ld_gs + c] = acc[i][j];
      else grad_agg[static_cast<int64_t>(r) * ld_ga + (c - dim)] = acc[i][j];
    }
  }
}

// ---------------------------------------------------------------------------------------
// bwd_w: dW[h, wcol(kv)] += sum_{r in chunk} dZ[r,h] X[r,kv]
// grid = (kv tiles, h tiles, row chunks); each CTA reduces `rows_per_chunk` rows.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGemmThreads)
sage_bwd_w_kernel(XOperand x, const float* __restrict__ grad_out, int64_t ld_go, const float* __restrict__ out,
                  int64_t ld_out, int out_dim, int relu, const int32_t* __restrict__ num_rows_dev, int max_rows,
                  int rows_per_chunk, float* __restrict__ grad_w, int64_t ldw) {
  pdl_sync();
  __shared__ __align__(16) float As[BK][BM + kPad];   // dZ tile: [row][h]
  __shared__ __align__(16) float Bs[BK][BN + kPad];   // X tile : [row][kv]
  const int rows = live_rows(num_rows_dev, max_rows);
  const int kv0 = blockIdx.x * BN, h0 = blockIdx.y * BM;
  const int r_begin = blockIdx.z * rows_per_chunk;
  const int r_end = min(rows, r_begin + rows_per_chunk);
  if (r_begin >= r_end) return;
  const int kt = x.kv_total();
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int l_row = tid >> 4, l_q = (tid & 15) * 4;            // 16 rows x 16 float4
  float acc[4][4] = {};
  float4 av, bv;
  auto fetch = [&](int r0) {
    const int r = r0 + l_row;
    av = make_float4(0.f, 0.f, 0.f, 0.f);
    bv = av;
    if (r < r_end) {
      float g[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int h = h0 + l_q + i;
        g[i] = 0.f;
        if (h < out_dim) {
          g[i] = grad_out[static_cast<int64_t>(r) * ld_go + h];
          if (relu && !(out[static_cast<int64_t>(r) * ld_out + h] > 0.f)) g[i] = 0.f;
        }
      }
      av = make_float4(g[0], g[1], g[2], g[3]);
      const int kv = kv0 + l_q;
      if (kv < kt) {
        const int self_row = x.gcn ? 0 : (x.self_idx ? x.self_idx[r] : r);
        bv = x.load4(r, self_row, kv);
      }
    }
  };
  fetch(r_begin);
  for (int r0 = r_begin; r0 < r_end; r0 += BK) {
    __syncthreads();
    *reinterpret_cast<float4*>(&As[l_row][l_q]) = av;
    *reinterpret_cast<float4*>(&Bs[l_row][l_q]) = bv;
    __syncthreads();
    if (r0 + BK < r_end) fetch(r0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) fma_tile(acc, &As[kk][ty * 4], &Bs[kk][tx * 4]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int h = h0 + ty * 4 + i;
    if (h >= out_dim) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kv = kv0 + tx * 4 + j;
      const int wc = kv < kt ? x.wcol(kv) : -1;
      if (wc >= 0) atomicAdd(grad_w + static_cast<int64_t>(h) * ldw + wc, acc[i][j]);
    }
  }
}

static int check_x(const float* self_table, int64_t ld_self, const float* agg, int64_t ld_agg, int dim, int gcn) {
  if (!agg || dim < 1) return GS_ERR_BAD_ARG;
  const int dp = (dim + 3) & ~3;
  if ((ld_agg & 3) || ld_agg < dp || !aligned16(agg)) return GS_ERR_ALIGNMENT;
  if (!gcn) {
    if (!self_table) return GS_ERR_BAD_ARG;
    if ((ld_self & 3) || ld_self < dp || !aligned16(self_table)) return GS_ERR_ALIGNMENT;
  }
  return GS_OK;
}

}  // namespace gs

using namespace gs;

int gs_sage_gemm_fwd_tc(const float*, int64_t, const int32_t*, const float*, int64_t, int32_t, const float*, int64_t,
                        int32_t, int32_t, const int32_t*, int32_t, float*, int64_t, int32_t, int32_t, float*, int64_t,
                        int32_t, gs_stream_t);
int gs_sage_gemm_fwd_tma(const float*, const float*, int64_t, int32_t, const float*, const float*, int64_t, int32_t,
                         const int32_t*, int32_t, float*, int64_t, int32_t, int32_t, float*, int64_t, gs_stream_t);
int gs_sage_gemm_fwd_tma_gather(const float*, int64_t, const int32_t*, const float*, int64_t, int32_t, const float*,
                                const float*, int64_t, int32_t, int32_t, const int32_t*, int32_t, float*, int64_t, int32_t,
                                int32_t, float*, int64_t, gs_stream_t);
int gs_sage_gemm_bwd_x_tc(const float*, int64_t, const float*, int64_t, const float*, int64_t, int32_t, int32_t, int32_t,
                          int32_t, const int32_t*, int32_t, float*, int64_t, float*, int64_t, int32_t, gs_stream_t);
int gs_sage_gemm_bwd_w_tc(const float*, int64_t, const int32_t*, const float*, int64_t, int32_t, const float*, int64_t,
                          const float*, int64_t, int32_t, int32_t, int32_t, const int32_t*, int32_t, float*, int64_t,
                          int32_t, gs_stream_t);
int gs_sage_gemm_bwd_w_group_tc(int32_t, const float* const*, const int64_t*, const int32_t* const*, const float* const*,
                                const int64_t*, const int32_t*, const float* const*, const int64_t*, const float* const*,
                                const int64_t*, const int32_t*, const int32_t*, const int32_t*, const int32_t*,
                                const int32_t* const*, const int32_t*, float* const*, const int64_t*, int32_t, gs_stream_t);

extern "C" int gs_sage_gemm_fwd(const float* self_table, int64_t ld_self, const int32_t* self_idx,
                                const float* agg, int64_t ld_agg, int32_t dim,
                                const float* weight, int64_t ldw, int32_t out_dim, int32_t gcn,
                                const int32_t* num_rows_dev, int32_t max_rows,
                                float* out, int64_t ld_out, int32_t relu, int32_t precision, gs_stream_t stream) {
  return gs_sage_gemm_fwd_ex(self_table, ld_self, self_idx, agg, ld_agg, dim, weight, ldw, out_dim, gcn, num_rows_dev,
                             max_rows, out, ld_out, relu, precision, nullptr, 0, nullptr, nullptr, 0, stream);
}

extern "C" int gs_sage_gemm_fwd_ex(const float* self_table, int64_t ld_self, const int32_t* self_idx,
                                   const float* agg, int64_t ld_agg, int32_t dim,
                                   const float* weight, int64_t ldw, int32_t out_dim, int32_t gcn,
                                   const int32_t* num_rows_dev, int32_t max_rows,
                                   float* out, int64_t ld_out, int32_t relu, int32_t precision,
                                   float* zero_out, int64_t ld_zero, const float* x_lo, const float* weight_lo,
                                   int32_t l2_normalize, gs_stream_t stream) {
  if (!weight || !out || out_dim < 1 || max_rows < 0) return GS_ERR_BAD_ARG;
  if (l2_normalize && precision == GS_PREC_FP32) return GS_ERR_UNSUPPORTED;    // an epilogue of the tensor-core kernel
  if (zero_out && ld_zero < out_dim) return GS_ERR_BAD_ARG;
  if (int e = check_x(self_table, ld_self, agg, ld_agg, dim, gcn)) return e;
  if (ldw < (gcn ? dim : 2 * dim) || ld_out < out_dim) return GS_ERR_BAD_ARG;
  if (max_rows == 0) return GS_OK;
  // The tensor core accumulates its fp32 partial sums with truncation, an error that grows
  // linearly with the number of K steps (measured 2.8e-6 at K=256, 2.3e-5 at K=2866); the
  // fp32-faithful mode therefore keeps contractions longer than 1024 on the FFMA path.
  const int k_total = gcn ? ((dim + 3) & ~3) : 2 * ((dim + 3) & ~3);
  if (precision == GS_PREC_TF32X3 && k_total > 1024) precision = GS_PREC_FP32;
  if (precision != GS_PREC_FP32) {
    // DENSE input rows (X = [self | agg] written side by side by gs_agg_fwd_x, no gather index) with the low halves of
    // both operands supplied: the all-TMA kernel (sage_gemm_tma.cu).  Anything else: the gathered-operand kernel.
    const bool dense = self_idx == nullptr && (dim & 3) == 0 &&
                       (gcn ? true : (self_table != nullptr && agg == self_table + dim && ld_self == ld_agg));
    if (dense && !l2_normalize && max_rows > 0 && (precision == GS_PREC_TF32 || (x_lo != nullptr && weight_lo != nullptr))) {
      const int e = gs_sage_gemm_fwd_tma(gcn ? agg : self_table, x_lo, ld_agg, gcn ? dim : 2 * dim, weight, weight_lo, ldw,
                                         out_dim, num_rows_dev, max_rows, out, ld_out, relu, precision, zero_out, ld_zero,
                                         stream);
      if (e != GS_ERR_UNSUPPORTED) return e;
    }
    // GATHERED input rows of a width the 32-column boxes handle (dim % 4 == 0): the self half by TMA gather4, the
    // aggregate half and W by tiled copies (sage_gemm_tma.cu); GS_TMA_GATHER=0 keeps the thread-staged kernel (A/B runs)
    // (GS_TMA_GATHER=2: no fallback -- a refused shape is reported, for tests that must know which kernel ran)
    static const int tma_gather = [] { const char* e = getenv("GS_TMA_GATHER"); return e ? atoi(e) : 1; }();
    if (tma_gather && !l2_normalize && (dim & 3) == 0 && agg != nullptr) {
      const int e = gs_sage_gemm_fwd_tma_gather(self_table, ld_self, self_idx, agg, ld_agg, dim, weight, weight_lo, ldw,
                                                out_dim, gcn, num_rows_dev, max_rows, out, ld_out, relu, precision,
                                                zero_out, ld_zero, stream);
      if (e != GS_ERR_UNSUPPORTED || tma_gather == 2) return e;
    }
    return gs_sage_gemm_fwd_tc(self_table, ld_self, self_idx, agg, ld_agg, dim, weight, ldw, out_dim, gcn,
                               num_rows_dev, max_rows, out, ld_out, relu, precision, zero_out, ld_zero, l2_normalize, stream);
  }
  if (zero_out) {       // FFMA path: a memset node in front (the tensor-core path folds the fill into its epilogue)
    cudaError_t e = cudaMemset2DAsync(zero_out, static_cast<size_t>(ld_zero) * 4, 0, static_cast<size_t>(out_dim) * 4,
                                      static_cast<size_t>(max_rows), as_stream(stream));
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  XOperand x{self_table, ld_self, self_idx, agg, ld_agg, dim, (dim + 3) & ~3, gcn};
  dim3 grid((max_rows + BM - 1) / BM, (out_dim + BN - 1) / BN);
  launch(sage_fwd_kernel, grid, kGemmThreads, 0, as_stream(stream), x, weight, ldw, out_dim, num_rows_dev, max_rows, out,
                                                                ld_out, relu);
  return finish_launch();
}

extern "C" int gs_sage_gemm_bwd_x(const float* grad_out, int64_t ld_go, const float* out, int64_t ld_out,
                                  const float* weight, int64_t ldw, int32_t dim, int32_t out_dim, int32_t gcn,
                                  int32_t relu, const int32_t* num_rows_dev, int32_t max_rows,
                                  float* grad_self, int64_t ld_gs, float* grad_agg, int64_t ld_ga, int32_t precision,
                                  gs_stream_t stream) {
  if (!grad_out || !weight || !grad_agg || dim < 1 || out_dim < 1 || max_rows < 0) return GS_ERR_BAD_ARG;
  if (relu && !out) return GS_ERR_BAD_ARG;
  if (!gcn && !grad_self) return GS_ERR_BAD_ARG;
  if (ld_ga < dim || (!gcn && ld_gs < dim) || ldw < (gcn ? dim : 2 * dim)) return GS_ERR_BAD_ARG;
  if (max_rows == 0) return GS_OK;
  if (precision != GS_PREC_FP32)
    return gs_sage_gemm_bwd_x_tc(grad_out, ld_go, out, ld_out, weight, ldw, dim, out_dim, gcn, relu, num_rows_dev,
                                 max_rows, grad_self, ld_gs, grad_agg, ld_ga, precision, stream);
  const int ncols = gcn ? dim : 2 * dim;
  dim3 grid((max_rows + BM - 1) / BM, (ncols + BN - 1) / BN);
  launch(sage_bwd_x_kernel, grid, kGemmThreads, 0, as_stream(stream), grad_out, ld_go, out, ld_out, weight, ldw, dim,
                                                                  out_dim, gcn, relu, num_rows_dev, max_rows,
                                                                  grad_self, ld_gs, grad_agg, ld_ga);
  return finish_launch();
}

extern "C" int gs_sage_gemm_bwd_w(const float* self_table, int64_t ld_self, const int32_t* self_idx,
                                  const float* agg, int64_t ld_agg, int32_t dim,
                                  const float* grad_out, int64_t ld_go, const float* out, int64_t ld_out,
                                  int32_t out_dim, int32_t gcn, int32_t relu,
                                  const int32_t* num_rows_dev, int32_t max_rows,
                                  float* grad_w, int64_t ldw, int32_t precision, gs_stream_t stream) {
  if (!grad_out || !grad_w || out_dim < 1 || max_rows < 0) return GS_ERR_BAD_ARG;
  if (relu && !out) return GS_ERR_BAD_ARG;
  if (int e = check_x(self_table, ld_self, agg, ld_agg, dim, gcn)) return e;
  if (ldw < (gcn ? dim : 2 * dim)) return GS_ERR_BAD_ARG;
  if (max_rows == 0) return GS_OK;
  if (precision != GS_PREC_FP32)
    return gs_sage_gemm_bwd_w_tc(self_table, ld_self, self_idx, agg, ld_agg, dim, grad_out, ld_go, out, ld_out, out_dim,
                                 gcn, relu, num_rows_dev, max_rows, grad_w, ldw, precision, stream);
  XOperand x{self_table, ld_self, self_idx, agg, ld_agg, dim, (dim + 3) & ~3, gcn};
  const int kt = gcn ? x.dim_pad : 2 * x.dim_pad;
  const int tiles = ((kt + BN - 1) / BN) * ((out_dim + BM - 1) / BM);
  // enough row chunks to fill the machine (~2 CTAs per SM), at least 64 rows per chunk
  int chunks = (2 * kNumSMs + tiles - 1) / tiles;
  const int max_chunks = (max_rows + 63) / 64;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  int rows_per_chunk = (max_rows + chunks - 1) / chunks;
  rows_per_chunk = ((rows_per_chunk + BK - 1) / BK) * BK;
  chunks = (max_rows + rows_per_chunk - 1) / rows_per_chunk;
  dim3 grid((kt + BN - 1) / BN, (out_dim + BM - 1) / BM, chunks);
  launch(sage_bwd_w_kernel, grid, kGemmThreads, 0, as_stream(stream), x, grad_out, ld_go, out, ld_out, out_dim, relu,
                                                                  num_rows_dev, max_rows, rows_per_chunk, grad_w, ldw);
  return finish_launch();
}

// Up to three weight-gradient problems (the layers of one step and its classifier) in one call.  Arrays of n (HOST
// arrays of device pointers / sizes).  Tensor-core precisions run them as ONE grid; otherwise, or when they cannot share
// a kernel, one after the other.
extern "C" int gs_sage_gemm_bwd_w_group(int32_t n, const float* const* self_table_host, const int64_t* ld_self_host,
                                        const int32_t* const* self_idx_host, const float* const* agg_host,
                                        const int64_t* ld_agg_host, const int32_t* dim_host,
                                        const float* const* grad_out_host, const int64_t* ld_go_host,
                                        const float* const* out_host, const int64_t* ld_out_host,
                                        const int32_t* out_dim_host, const int32_t* grad_out_cols_host,
                                        const int32_t* gcn_host, const int32_t* relu_host,
                                        const int32_t* const* num_rows_dev_host, const int32_t* max_rows_host,
                                        float* const* grad_w_host, const int64_t* ldw_host, int32_t precision,
                                        gs_stream_t stream) {
  if (n < 1 || n > 3 || !self_table_host || !ld_self_host || !self_idx_host || !agg_host || !ld_agg_host || !dim_host ||
      !grad_out_host || !ld_go_host || !out_host || !ld_out_host || !out_dim_host || !gcn_host || !relu_host ||
      !num_rows_dev_host || !max_rows_host || !grad_w_host || !ldw_host)
    return GS_ERR_BAD_ARG;
  bool fused = precision != GS_PREC_FP32;
  for (int i = 0; i < n; ++i) {
    if (max_rows_host[i] <= 0) { fused = false; continue; }
    if (!grad_out_host[i] || !grad_w_host[i] || out_dim_host[i] < 1 || (relu_host[i] && !out_host[i])) return GS_ERR_BAD_ARG;
    if (int e = check_x(self_table_host[i], ld_self_host[i], agg_host[i], ld_agg_host[i], dim_host[i], gcn_host[i])) return e;
    if (ldw_host[i] < (gcn_host[i] ? dim_host[i] : 2 * dim_host[i])) return GS_ERR_BAD_ARG;
    if (grad_out_cols_host && grad_out_cols_host[i] > 0 &&
        (grad_out_cols_host[i] < out_dim_host[i] || grad_out_cols_host[i] > ld_go_host[i]))
      return GS_ERR_BAD_ARG;
  }
  if (fused) {
    const int e = gs_sage_gemm_bwd_w_group_tc(n, self_table_host, ld_self_host, self_idx_host, agg_host, ld_agg_host,
                                              dim_host, grad_out_host, ld_go_host, out_host, ld_out_host, out_dim_host,
                                              grad_out_cols_host, gcn_host, relu_host, num_rows_dev_host, max_rows_host,
                                              grad_w_host, ldw_host, precision, stream);
    if (e != GS_ERR_UNSUPPORTED) return e;
  }
  for (int i = 0; i < n; ++i) {
    const int e = gs_sage_gemm_bwd_w(self_table_host[i], ld_self_host[i], self_idx_host[i], agg_host[i], ld_agg_host[i],
                                     dim_host[i], grad_out_host[i], ld_go_host[i], out_host[i], ld_out_host[i],
                                     out_dim_host[i], gcn_host[i], relu_host[i], num_rows_dev_host[i], max_rows_host[i],
                                     grad_w_host[i], ldw_host[i], precision, stream);
    if (e) return e;
  }
  return GS_OK;
}

// The two-problem form with common gcn / relu (the layers of a two-layer step).
extern "C" int gs_sage_gemm_bwd_w_pair(const float* const* self_table_host, const int64_t* ld_self_host,
                                       const int32_t* const* self_idx_host, const float* const* agg_host,
                                       const int64_t* ld_agg_host, const int32_t* dim_host,
                                       const float* const* grad_out_host, const int64_t* ld_go_host,
                                       const float* const* out_host, const int64_t* ld_out_host,
                                       const int32_t* out_dim_host, int32_t gcn, int32_t relu,
                                       const int32_t* const* num_rows_dev_host, const int32_t* max_rows_host,
                                       float* const* grad_w_host, const int64_t* ldw_host, int32_t precision,
                                       gs_stream_t stream) {
  const int32_t g[2] = {gcn, gcn}, r[2] = {relu, relu};
  return gs_sage_gemm_bwd_w_group(2, self_table_host, ld_self_host, self_idx_host, agg_host, ld_agg_host, dim_host,
                                  grad_out_host, ld_go_host, out_host, ld_out_host, out_dim_host, nullptr, g, r,
                                  num_rows_dev_host, max_rows_host, grad_w_host, ldw_host, precision, stream);
}

// ---------------------------------------------------------------------------------------
// ReLU backward in place: grad[r, c] = 0 where out[r, c] <= 0.  The tensor-core backward
// kernels stream dZ with cp.async (no registers in between), so the mask is applied once
// here instead of inside both of them.
// ---------------------------------------------------------------------------------------
namespace gs {
__global__ void __launch_bounds__(256)
relu_bwd_kernel(float* __restrict__ grad, int64_t ld_g, const float* __restrict__ out, int64_t ld_out, int dim4,
                const int32_t* __restrict__ num_rows_dev, int max_rows) {
  pdl_sync();
  const int rows = live_rows(num_rows_dev, max_rows);
  const int64_t total = static_cast<int64_t>(rows) * dim4;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / dim4;
    const int c = static_cast<int>(i - r * dim4) * 4;
    float4 g = *reinterpret_cast<const float4*>(grad + r * ld_g + c);
    const float4 o = *reinterpret_cast<const float4*>(out + r * ld_out + c);
    g.x = o.x > 0.f ? g.x : 0.f;
    g.y = o.y > 0.f ? g.y : 0.f;
    g.z = o.z > 0.f ? g.z : 0.f;
    g.w = o.w > 0.f ? g.w : 0.f;
    *reinterpret_cast<float4*>(grad + r * ld_g + c) = g;
  }
}
}  // namespace gs

extern "C" int gs_relu_bwd_inplace(float* grad, int64_t ld_g, const float* out, int64_t ld_out, int32_t dim,
                                   const int32_t* num_rows_dev, int32_t max_rows, gs_stream_t stream) {
  if (!grad || !out || dim < 1 || max_rows < 0) return GS_ERR_BAD_ARG;
  const int dim4 = (dim + 3) / 4;
  if ((ld_g & 3) || (ld_out & 3) || ld_g < 4 * dim4 || ld_out < 4 * dim4 || !aligned16(grad) || !aligned16(out))
    return GS_ERR_ALIGNMENT;
  if (max_rows == 0) return GS_OK;
  const int64_t total = static_cast<int64_t>(max_rows) * dim4;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  launch(relu_bwd_kernel, blocks, 256, 0, as_stream(stream), grad, ld_g, out, ld_out, dim4, num_rows_dev, max_rows);
  return finish_launch();
}
