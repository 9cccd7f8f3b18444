// The TOP SageLayer of a supervised step, forward and backward, in ONE launch.
//
// A step's last layer works on the b_sz batch rows only (1024 at the headline configuration): as separate
// kernels -- aggregation, SageLayer GEMM, classifier/loss tail, dX GEMM, scatter -- it was five launches of 5-18 us
// with 8 CTAs each on a 148-SM machine, 58 us of a 95 us training chain.  Here one CTA owns a tile of 16 batch rows
// from the gather to the scatter; a batch of 1024 rows is 64 CTAs, each doing per tile
//
//   gather   X[r] = [ table[self_idx[r]] | mean_j table[nbr_idx[r][j]] ]   src/models.py:260-266, 300-314
//   layer    h    = relu(X . W^T)                                          src/models.py:215-219
//   head     logp = log_softmax(h . Wc^T + bc),  loss -= logp[y] / rows    src/models.py:25-27, src/utils.py:162-163
//   backward dlogits -> grad Wc, grad bc, dh -> dZ = dh * (h > 0) -> dX = dZ . W            (autograd of the above)
//   scatter  grad_table[t] += (dX_self | dX_agg / cnt) * (table[t] > 0)    (autograd of the gather + the ReLU of the
//                                                                           layer below, src/models.py:219)
//
// and saving agg (B operand) and dZ (A operand) for the dW GEMM of this layer (gs_sage_gemm_bwd_w), which runs beside
// the layer below's dW GEMM.  The two [16 x K] x [K x 128] contractions run on the tensor cores with warp-level
// mma.sync (m16n8k8, tf32 inputs, fp32 accumulate) in the same 3-term split the tcgen05 kernels use
// (x = hi + lo, D = A_lo.B_hi + A_hi.B_lo + A_hi.B_hi: fp32-faithful), because a 16-row tile is an eighth of the
// smallest tcgen05 tile; W (128 KB) sits in shared memory for both contractions.  Everything else is fp32 FFMA.
// The ReLU gates of the gathered rows are kept as bits in registers from the gather to the scatter, so the table is
// read once.  The loss is reduced without atomics on floats (per-CTA partials, last CTA sums them in a fixed order).
// Two options the trainers use: out_dlog -- d(logits) is saved and grad Wc = dlog^T . h is left to the grouped
// weight-gradient launch (64 CTAs adding into one [C x 128] block cost each of them 3.5-4.4K cycles here); and
// gs_set_early_reads -- the tile's index lists and labels are loaded before the wait for the previous kernel.
#include <cuda.h>
#include <string.h>

#include "common.cuh"

namespace gs {
namespace tc {      // csrc/sage_gemm_tc.cu: 2-D fp32 tensor map, box = 32 columns x box_rows rows, SWIZZLE_128B, OOB reads zero
bool make_tmap_2d(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);
}
namespace top {

constexpr int kH = 128;             // layer width (SageLayer out_size); the reference's hidden size
constexpr int kTM = 16;             // batch rows per tile = the M of one mma.sync
constexpr int kThreads = 512;       // 16 warps: warp w owns tile row w in the row-wise phases, n-tile w in the MMAs
constexpr int kWarps = kThreads / 32;
constexpr int kMaxStride = 16;      // sampled-list stride (fan-out 10, +1 with gcn)
constexpr int kMaxClasses = 64;
constexpr int kHs = kH + 8;         // row stride of the [.. x 128] tiles: = 8 mod 32 -> conflict-free fragment loads
constexpr int kCs = kMaxClasses + 8;   // row stride of the [16 x classes] tiles, = 8 mod 32

// optional phase trace (build with -DGS_TOP_TRACE): CTA 0 records the SM clock at the phase boundaries of its first tile
#ifdef GS_TOP_TRACE
__device__ long long g_top_trace[16];
#define GS_TOP_MARK(slot) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_top_trace[(slot)] = clock64(); } while (0)
#else
#define GS_TOP_MARK(slot) do { } while (0)
#endif

struct Params {
  const float* table; int64_t ld_table;
  const int32_t* nbr_idx; int stride; const int32_t* cnt; const int32_t* self_idx;
  const int32_t* num_rows_dev; int max_rows;
  const float* weight; int64_t ldw;
  const float* cls_w; const float* cls_b; int num_classes;
  const int64_t* labels; const int32_t* label_index;
  float* out_h; int64_t ld_h;
  float* out_agg; int64_t ld_agg;
  float* out_dz; int64_t ld_dz;
  float* out_dlog; int64_t ld_dlog;                       // d(logits), kMaxClasses columns per row (zeros beyond num_classes)
  float* logp;
  float* loss; float* grad_cls_w; float* grad_cls_b;
  float* cls_w_rep; float* cls_b_rep; int cls_reps;       // replicas 1..cls_reps-1 of the classifier gradients (see P4c)
  float* grad_table; int64_t ld_gt;
  float* partials; unsigned int* ticket;
  int early_lists;                                        // gs_set_early_reads at launch time
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
// TMA: the weights arrive as a handful of tiled bulk copies issued by one thread (the LSU path, cp.async per 16 bytes,
// cost ~5K cycles of issue per CTA for the 128 KB of W)
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
// A [rows x 32k] fp32 matrix as the TMA writes it with SWIZZLE_128B boxes of 32 columns: box b (columns 32b..32b+31)
// is `box_bytes` long, its rows are 128 bytes, the 16-byte pieces of a row XOR-ed with (row & 7)
__device__ __forceinline__ int swz(int row, int col, int box_bytes) {
  return (col >> 5) * box_bytes + row * 128 + ((((col & 31) >> 2) ^ (row & 7)) << 4) + ((col & 3) << 2);
}
__device__ __forceinline__ float lds_swz(const unsigned char* base, int row, int col, int box_bytes) {
  return *reinterpret_cast<const float*>(base + swz(row, col, box_bytes));
}
// The 8 piece offsets of one swizzled row: piece j (columns 4j..4j+3 of a 32-column box) of row `row` sits at
// xor8[j]; with them a fragment address is  box * box_bytes + row * 128 + xor8[piece] + (col & 3) * 4  -- all but the
// (compile-time) box term are per-lane constants, so the unrolled MMA loops issue bare loads.
__device__ __forceinline__ void swz_pieces(int row, int (&xor8)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) xor8[j] = (j ^ (row & 7)) << 4;
}

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;                  // the 19 bits a tf32 operand carries (truncation)
  lo = __float_as_uint(x - __uint_as_float(hi));          // exact in fp32; the tensor core truncates it in turn
}
// D += A(16x8, row) . B(8x8, col); tf32 in, fp32 accumulate
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// d: the hi.hi product, e: the two correction products -- separate accumulators, i.e. two independent dependency
// chains through the tensor pipe (and the small terms are summed among themselves first); the caller adds e to d
template <bool SPLIT3>
__device__ __forceinline__ void mma_split(float (&d)[4], float (&e)[4], const float (&a)[4], const float (&b)[2]) {
  uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_tf32(a[i], ah[i], al[i]);
#pragma unroll
  for (int i = 0; i < 2; ++i) split_tf32(b[i], bh[i], bl[i]);
  mma_tf32(d, ah, bh);
  if (SPLIT3) {
    mma_tf32(e, al, bh);
    mma_tf32(e, ah, bl);
  }
}

__device__ __forceinline__ void red_add_f4(float* dst, float4 v) { atomicAdd(reinterpret_cast<float4*>(dst), v); }

constexpr int kWBox = kH * 128;        // bytes of one TMA box of W: 128 rows x 32 columns
constexpr int kWcBox = kMaxClasses * 128;

template <bool GCN>
struct Smem {
  static constexpr int kK = GCN ? kH : 2 * kH;            // contraction length of the layer (X columns)
  static constexpr int kXs = kK + 8;                      // row stride of the X tile, = 8 mod 32
  unsigned char w[(kK / 32) * kWBox];                     // W[h][k] in TMA boxes (1024-byte aligned: first member)
  unsigned char wc[(kH / 32) * kWcBox];                   // classifier weight in TMA boxes, rows >= num_classes zero
  float x[kTM * kXs];                                     // X tile; later dX
  float h[kTM * kHs];                                     // relu output
  float dz[kTM * kHs];                                    // d(pre-activation)
  float logit[kTM * kCs];
  float dlog[kTM * kCs];                                  // d(logits), columns >= num_classes zero
  float red[kWarps];
  float bias[kMaxClasses];
  int32_t nbr[kTM * kMaxStride];
  int32_t cnt[kTM];
  int32_t self_row[kTM];
  int32_t label[kTM];
  uint64_t bar;                                           // the weights have landed
};

template <bool GCN, bool SPLIT3>
__global__ void __maxnreg__(72)
sage_top_sup_kernel(const Params p, const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_wc) {
  GS_TOP_MARK(0);
  // With gs_set_early_reads the wait for the previous kernels of the stream comes LATE: the index lists and labels of
  // the first tile are read before it (in a pipelined train step they come from the preparation branch, which the
  // step's first kernel waits for with a full dependency -- common.cuh: pdl_wait).  Weights, bias and the rows of the
  // layer below are always read after it.
  if (!p.early_lists) pdl_sync();
  extern __shared__ unsigned char smem_raw[];
  using S = Smem<GCN>;
  // the TMA boxes (first members) need 1024-byte alignment; the launch asks for 1 KB of slack
  S& s = *reinterpret_cast<S*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  constexpr int K = S::kK, XS = S::kXs;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;                  // mma fragment coordinates
  const int rows = live_rows(p.num_rows_dev, p.max_rows);
  const int tiles = (rows + kTM - 1) / kTM;
  const int C = p.num_classes;
  const int c_pad = (C + 15) & ~15;                       // classes rounded up to the MMA shapes (<= 64)
  const float inv_rows = 1.0f / static_cast<float>(rows > 0 ? rows : 1);
  float loss_part = 0.f;                                  // this thread's share of -sum logp[y] / rows
  bool weights_pending = static_cast<int>(blockIdx.x) < tiles;

  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int row0 = tile * kTM;
    // ---- P1: the tile's lists ----
    if (tid < kTM * kMaxStride) {
      const int r = tid / kMaxStride, j = tid - r * kMaxStride;
      const int row = row0 + r;
      s.nbr[tid] = (row < rows && j < p.stride) ? __ldg(p.nbr_idx + static_cast<int64_t>(row) * p.stride + j) : -1;
    } else if (tid < kTM * kMaxStride + kTM) {
      const int r = tid - kTM * kMaxStride, row = row0 + r;
      const bool live = row < rows;
      s.cnt[r] = live ? __ldg(p.cnt + row) : 0;
      s.self_row[r] = (live && !GCN) ? (p.self_idx ? __ldg(p.self_idx + row) : row) : -1;
    } else if (tid < kTM * kMaxStride + 2 * kTM) {
      const int r = tid - kTM * kMaxStride - kTM, row = row0 + r;
      s.label[r] = row < rows ? static_cast<int>(__ldg(p.labels + (p.label_index ? __ldg(p.label_index + row) : row))) : -1;
    }
    if (tile == static_cast<int>(blockIdx.x)) {           // first tile: from here on, data the previous kernels wrote
      if (p.early_lists) { pdl_wait(); pdl_trigger(); }
      GS_TOP_MARK(1);
      if (tid < kMaxClasses) s.bias[tid] = (tid < C && p.cls_b) ? __ldg(p.cls_b + tid) : 0.f;   // visible after the first barrier
      if (weights_pending && tid == 0) {
        // W and Wc -> shared memory: K/32 + 4 tiled bulk copies, landing while the first tile is gathered
        const uint32_t bar = smem_u32(&s.bar);
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar, static_cast<uint32_t>((K / 32) * kWBox + (kH / 32) * kWcBox));
    #pragma unroll
        for (int b = 0; b < K / 32; ++b) tma_load_2d(smem_u32(s.w + b * kWBox), &tmap_w, 32 * b, 0, bar);
    #pragma unroll
        for (int b = 0; b < kH / 32; ++b) tma_load_2d(smem_u32(s.wc + b * kWcBox), &tmap_wc, 32 * b, 0, bar);
      }
    }
    __syncthreads();
    GS_TOP_MARK(2);

    // ---- P2: gather + mean.  Warp w owns tile row w; a lane owns 4 consecutive columns of a table row (128-bit
    //      loads, a 512-byte row per warp instruction); the whole fan-out is in flight at once.  The ReLU gates of
    //      what was read stay in registers for the scatter. ----
    uint64_t gate = 0;                                    // bit 4j+c: table[nbr_j][4 lane + c] > 0
    uint32_t gate_self = 0;
    const int my_n = s.cnt[warp];
    {
      const int r = warp;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 self_v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int sr = s.self_row[r];
      if (!GCN && sr >= 0) self_v = ldg_stream_f4(p.table + static_cast<int64_t>(sr) * p.ld_table + 4 * lane);
      for (int j0 = 0; j0 < my_n; j0 += 12) {
        float4 v[12];
#pragma unroll
        for (int u = 0; u < 12; ++u) {
          const int j = j0 + u;
          const int tr = (j < my_n && j < kMaxStride) ? s.nbr[r * kMaxStride + j] : -1;
          v[u] = ldg_stream_f4_if<0u>(p.table + static_cast<int64_t>(tr < 0 ? 0 : tr) * p.ld_table + 4 * lane, tr >= 0);
        }
#pragma unroll
        for (int u = 0; u < 12; ++u) {
          acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
          const uint64_t bits = (v[u].x > 0.f ? 1u : 0u) | (v[u].y > 0.f ? 2u : 0u) | (v[u].z > 0.f ? 4u : 0u) | (v[u].w > 0.f ? 8u : 0u);
          if (j0 + u < kMaxStride) gate |= bits << (4 * (j0 + u));
        }
      }
      const int row = row0 + r;
      if (row < rows) {
        const float inv = 1.0f / static_cast<float>(my_n);   // n == 0: 0 * inf = NaN, the reference's 0/0 row (src/models.py:312)
        acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
        if (p.out_agg) *reinterpret_cast<float4*>(p.out_agg + static_cast<int64_t>(row) * p.ld_agg + 4 * lane) = acc;
      }
      if (!GCN) {
        gate_self = (self_v.x > 0.f ? 1u : 0u) | (self_v.y > 0.f ? 2u : 0u) | (self_v.z > 0.f ? 4u : 0u) | (self_v.w > 0.f ? 8u : 0u);
        *reinterpret_cast<float4*>(&s.x[r * XS + 4 * lane]) = self_v;
        *reinterpret_cast<float4*>(&s.x[r * XS + kH + 4 * lane]) = acc;
      } else {
        *reinterpret_cast<float4*>(&s.x[r * XS + 4 * lane]) = acc;
      }
    }
    GS_TOP_MARK(3);
    __syncthreads();                                      // the X tile is complete; (first tile) the barrier is initialised
    if (weights_pending) {
      mbar_wait(smem_u32(&s.bar), 0);                     // the weights have landed
      weights_pending = false;
    }
    GS_TOP_MARK(4);

    // ---- P3: h = relu(X . W^T) on the tensor cores.  Warp w owns output columns 8w .. 8w+7 (one n8 tile). ----
    {
      float acc[4] = {0.f, 0.f, 0.f, 0.f}, cor[4] = {0.f, 0.f, 0.f, 0.f};
      const float* xa = &s.x[g * XS + t];
      const float* xb = &s.x[(g + 8) * XS + t];
      const int n = 8 * warp + g;
      int pc[8];
      swz_pieces(n, pc);
      const unsigned char* wrow = s.w + n * 128 + t * 4;
#pragma unroll
      for (int b = 0; b < K / 32; ++b) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k0 = 32 * b + 8 * j;
          const float a[4] = {xa[k0], xb[k0], xa[k0 + 4], xb[k0 + 4]};
          const float bv[2] = {*reinterpret_cast<const float*>(wrow + b * kWBox + pc[2 * j]),
                               *reinterpret_cast<const float*>(wrow + b * kWBox + pc[2 * j + 1])};
          mma_split<SPLIT3>(acc, cor, a, bv);
        }
      }
      const int col = 8 * warp + 2 * t;
      const float2 top = make_float2(fmaxf(acc[0] + cor[0], 0.f), fmaxf(acc[1] + cor[1], 0.f));      // src/models.py:219
      const float2 bot = make_float2(fmaxf(acc[2] + cor[2], 0.f), fmaxf(acc[3] + cor[3], 0.f));
      *reinterpret_cast<float2*>(&s.h[g * kHs + col]) = top;
      *reinterpret_cast<float2*>(&s.h[(g + 8) * kHs + col]) = bot;
      if (p.out_h) {
        if (row0 + g < rows) *reinterpret_cast<float2*>(p.out_h + static_cast<int64_t>(row0 + g) * p.ld_h + col) = top;
        if (row0 + g + 8 < rows) *reinterpret_cast<float2*>(p.out_h + static_cast<int64_t>(row0 + g + 8) * p.ld_h + col) = bot;
      }
    }
    __syncthreads();
    GS_TOP_MARK(5);

    // ---- P4a: logits = h . Wc^T + bc on the tensor cores.  Warps 0..7 own the classes 8w .. 8w+7 over the first half
    //      of the contraction, warps 8..15 the same classes over the second half (two short dependency chains instead
    //      of one long one; the halves meet in shared memory). ----
    {
      const int cw = warp & 7, half = warp >> 3;
      if (8 * cw < c_pad) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f}, cor[4] = {0.f, 0.f, 0.f, 0.f};
        const float* ha = &s.h[g * kHs + t];
        const float* hb = &s.h[(g + 8) * kHs + t];
        const int n = 8 * cw + g;
        int pc[8];
        swz_pieces(n, pc);
        const unsigned char* wrow = s.wc + n * 128 + t * 4;
#pragma unroll
        for (int bb = 0; bb < kH / 64; ++bb) {
          const int b = half * (kH / 64) + bb;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k0 = 32 * b + 8 * j;
            const float a[4] = {ha[k0], hb[k0], ha[k0 + 4], hb[k0 + 4]};
            const float bv[2] = {*reinterpret_cast<const float*>(wrow + b * kWcBox + pc[2 * j]),
                                 *reinterpret_cast<const float*>(wrow + b * kWcBox + pc[2 * j + 1])};
            mma_split<SPLIT3>(acc, cor, a, bv);
          }
        }
        const int c = 8 * cw + 2 * t;
        float* dst = half == 0 ? s.logit : s.dlog;          // dlog is free until the softmax writes it
        *reinterpret_cast<float2*>(&dst[g * kCs + c]) = make_float2(acc[0] + cor[0], acc[1] + cor[1]);
        *reinterpret_cast<float2*>(&dst[(g + 8) * kCs + c]) = make_float2(acc[2] + cor[2], acc[3] + cor[3]);
      }
    }
    __syncthreads();
    GS_TOP_MARK(10);

    // ---- P4b: log-softmax, NLL and d(logits): 16 threads per row, thread j owns classes j, j+16, j+32, j+48 ----
    if (tid < kTM * 16) {
      const int r = tid >> 4, j = tid & 15;
      const int row = row0 + r;
      float z[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = j + 16 * q;
        z[q] = (c < C) ? s.logit[r * kCs + c] + s.dlog[r * kCs + c] + s.bias[c] : -INFINITY;
      }
      float m = fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3]));
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float se = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) if (j + 16 * q < C) se += expf(z[q] - m);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
      const float lse = m + logf(se);
      const int y = s.label[r];
      const bool live = row < rows;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = j + 16 * q;
        float d = 0.f;
        if (c < C) {
          const float lp = z[q] - lse;                                          // src/models.py:26
          if (live && p.logp) p.logp[static_cast<int64_t>(row) * C + c] = lp;
          if (live && c == y) loss_part -= lp * inv_rows;                       // src/utils.py:162-163
          if (live) d = (expf(lp) - (c == y ? 1.f : 0.f)) * inv_rows;
        }
        s.dlog[r * kCs + c] = d;                                                // columns C .. 63 are zero
      }
    }
    __syncthreads();
    if (p.out_dlog && tid < kTM * (kMaxClasses / 4)) {    // d(logits) for a classifier weight gradient computed elsewhere
      const int r = tid / (kMaxClasses / 4), c4 = (tid % (kMaxClasses / 4)) * 4;
      if (row0 + r < rows)
        *reinterpret_cast<float4*>(p.out_dlog + static_cast<int64_t>(row0 + r) * p.ld_dlog + c4) =
            *reinterpret_cast<const float4*>(&s.dlog[r * kCs + c4]);
    }
    GS_TOP_MARK(6);

    // ---- P4c: dh = dlog . Wc on the tensor cores, dZ = dh * (h > 0); then grad Wc = dlog^T . h and grad bc ----
    {
      const int n0 = 8 * warp;                            // this warp's 8 columns of h (both products)
      {
        float acc[4] = {0.f, 0.f, 0.f, 0.f}, cor[4] = {0.f, 0.f, 0.f, 0.f};   // dh[r][h] = sum_c dlog[r][c] Wc[c][h]
        // B[k = class][n = column]: rows c0 + t and c0 + t + 4 of the swizzled Wc tile, column n0 + g
        const int colw = n0 + g;
        const unsigned char* wb0 = s.wc + (colw >> 5) * kWcBox + t * 128 + (((((colw & 31) >> 2)) ^ t) << 4) + (colw & 3) * 4;
        const unsigned char* wb1 = s.wc + (colw >> 5) * kWcBox + (t + 4) * 128 + (((((colw & 31) >> 2)) ^ (t + 4)) << 4) + (colw & 3) * 4;
#pragma unroll 2
        for (int c0 = 0; c0 < c_pad; c0 += 8) {
          const float a[4] = {s.dlog[g * kCs + c0 + t], s.dlog[(g + 8) * kCs + c0 + t], s.dlog[g * kCs + c0 + t + 4],
                              s.dlog[(g + 8) * kCs + c0 + t + 4]};
          const float b[2] = {*reinterpret_cast<const float*>(wb0 + c0 * 128), *reinterpret_cast<const float*>(wb1 + c0 * 128)};
          mma_split<SPLIT3>(acc, cor, a, b);
        }
        const int col = n0 + 2 * t;
        const float2 h_top = *reinterpret_cast<const float2*>(&s.h[g * kHs + col]);
        const float2 h_bot = *reinterpret_cast<const float2*>(&s.h[(g + 8) * kHs + col]);
        const float2 top = make_float2(h_top.x > 0.f ? acc[0] + cor[0] : 0.f, h_top.y > 0.f ? acc[1] + cor[1] : 0.f);
        const float2 bot = make_float2(h_bot.x > 0.f ? acc[2] + cor[2] : 0.f, h_bot.y > 0.f ? acc[3] + cor[3] : 0.f);
        *reinterpret_cast<float2*>(&s.dz[g * kHs + col]) = top;
        *reinterpret_cast<float2*>(&s.dz[(g + 8) * kHs + col]) = bot;
        if (p.out_dz) {
          if (row0 + g < rows) *reinterpret_cast<float2*>(p.out_dz + static_cast<int64_t>(row0 + g) * p.ld_dz + col) = top;
          if (row0 + g + 8 < rows) *reinterpret_cast<float2*>(p.out_dz + static_cast<int64_t>(row0 + g + 8) * p.ld_dz + col) = bot;
        }
      }
      GS_TOP_MARK(11);
      // 64 CTAs adding into the same [C x 128] block serialise in the L2 atomic units (~60 cycles per add and address:
      // 4K cycles of tail).  CTA b therefore adds into replica b % cls_reps (replica 0 = the gradient buffer itself,
      // the others a side buffer the update kernel folds in and clears): 8 adds per address instead of 64.
      const int rep = p.cls_reps > 1 ? static_cast<int>(blockIdx.x) % p.cls_reps : 0;
      float* gcw = rep == 0 ? p.grad_cls_w : p.cls_w_rep + static_cast<int64_t>(rep - 1) * C * kH;
      float* gcb = rep == 0 ? p.grad_cls_b : p.cls_b_rep + static_cast<int64_t>(rep - 1) * kMaxClasses;
      if (p.grad_cls_w) {
        for (int m0 = 0; m0 < c_pad; m0 += 16) {          // D[c][h] = sum_r dlog[r][c] h[r][h]
          float acc[4] = {0.f, 0.f, 0.f, 0.f}, cor[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int r0 = 0; r0 < kTM; r0 += 8) {
            const float a[4] = {s.dlog[(r0 + t) * kCs + m0 + g], s.dlog[(r0 + t) * kCs + m0 + g + 8],
                                s.dlog[(r0 + t + 4) * kCs + m0 + g], s.dlog[(r0 + t + 4) * kCs + m0 + g + 8]};
            const float b[2] = {s.h[(r0 + t) * kHs + n0 + g], s.h[(r0 + t + 4) * kHs + n0 + g]};
            mma_split<SPLIT3>(acc, cor, a, b);
          }
          const int c = m0 + g, hcol = n0 + 2 * t;
          if (c < C) atomicAdd(reinterpret_cast<float2*>(gcw + static_cast<int64_t>(c) * kH + hcol),
                               make_float2(acc[0] + cor[0], acc[1] + cor[1]));
          if (c + 8 < C) atomicAdd(reinterpret_cast<float2*>(gcw + static_cast<int64_t>(c + 8) * kH + hcol),
                                   make_float2(acc[2] + cor[2], acc[3] + cor[3]));
        }
      }
      if (tid < C && p.grad_cls_b) {
        float sb = 0.f;
#pragma unroll
        for (int r = 0; r < kTM; ++r) sb += s.dlog[r * kCs + tid];
        atomicAdd(gcb + tid, sb);
      }
    }
    __syncthreads();
    GS_TOP_MARK(7);

    // ---- P5: dX = dZ . W on the tensor cores (contraction over the 128 outputs).  Warp w owns dX columns
    //      [w*K/16, (w+1)*K/16); the X tile in shared memory is dead and receives dX. ----
    {
      constexpr int NT = K / (8 * kWarps);                // n8 tiles per warp: 2 (K = 256) or 1 (gcn)
      float acc[NT][4] = {}, cor[NT][4] = {};
      const int c_base = warp * (K / kWarps);
      // B[k = output h][n = input column]: rows h0 + t and h0 + t + 4 of the swizzled W tile; (h & 7) = t or t + 4 for
      // every h0, so a lane's two addresses per n-tile only advance by h0 * 128
      const unsigned char* wb0[NT];
      const unsigned char* wb1[NT];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int colw = c_base + 8 * nt + g;
        const unsigned char* base = s.w + (colw >> 5) * kWBox + (colw & 3) * 4;
        wb0[nt] = base + t * 128 + ((((colw & 31) >> 2) ^ t) << 4);
        wb1[nt] = base + (t + 4) * 128 + ((((colw & 31) >> 2) ^ (t + 4)) << 4);
      }
      const float* da = &s.dz[g * kHs + t];
      const float* db = &s.dz[(g + 8) * kHs + t];
#pragma unroll
      for (int h0 = 0; h0 < kH; h0 += 8) {
        const float a[4] = {da[h0], db[h0], da[h0 + 4], db[h0 + 4]};
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const float b[2] = {*reinterpret_cast<const float*>(wb0[nt] + h0 * 128), *reinterpret_cast<const float*>(wb1[nt] + h0 * 128)};
          mma_split<SPLIT3>(acc[nt], cor[nt], a, b);
        }
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int col = c_base + 8 * nt + 2 * t;
        *reinterpret_cast<float2*>(&s.x[g * XS + col]) = make_float2(acc[nt][0] + cor[nt][0], acc[nt][1] + cor[nt][1]);
        *reinterpret_cast<float2*>(&s.x[(g + 8) * XS + col]) = make_float2(acc[nt][2] + cor[nt][2], acc[nt][3] + cor[nt][3]);
      }
    }
    __syncthreads();
    GS_TOP_MARK(8);

    // ---- P6: scatter into the gradient of the layer below, ReLU gate of the TARGET row applied (the gates were
    //      recorded while gathering): grad_table receives d(pre-activation) of layer L-1 directly. ----
    if (p.grad_table && row0 + warp < rows) {
      const int r = warp;
      if (!GCN && gate_self) {
        const float4 d = *reinterpret_cast<const float4*>(&s.x[r * XS + 4 * lane]);
        const float4 v = make_float4(gate_self & 1u ? d.x : 0.f, gate_self & 2u ? d.y : 0.f, gate_self & 4u ? d.z : 0.f,
                                     gate_self & 8u ? d.w : 0.f);
        red_add_f4(p.grad_table + static_cast<int64_t>(s.self_row[r]) * p.ld_gt + 4 * lane, v);
      }
      float4 d = *reinterpret_cast<const float4*>(&s.x[r * XS + (GCN ? 0 : kH) + 4 * lane]);
      const float inv = 1.0f / static_cast<float>(my_n > 0 ? my_n : 1);
      d.x *= inv; d.y *= inv; d.z *= inv; d.w *= inv;
      for (int j = 0; j < my_n; ++j) {
        const uint32_t gb = static_cast<uint32_t>(gate >> (4 * j)) & 15u;
        if (gb) {
          const float4 v = make_float4(gb & 1u ? d.x : 0.f, gb & 2u ? d.y : 0.f, gb & 4u ? d.z : 0.f, gb & 8u ? d.w : 0.f);
          red_add_f4(p.grad_table + static_cast<int64_t>(s.nbr[r * kMaxStride + j]) * p.ld_gt + 4 * lane, v);
        }
      }
    }
    __syncthreads();                                      // the tile buffers are rewritten by the next tile
    GS_TOP_MARK(9);
  }

  if (p.early_lists) pdl_wait();                          // (a CTA without a tile has not waited yet; the workspace is shared with the previous launch)
  // ---- loss: per-CTA partial, the last CTA to finish adds them up in CTA order (deterministic, nothing to zero) ----
  loss_part = warp_sum(loss_part);
  if (lane == 0) s.red[warp] = loss_part;
  __syncthreads();
  if (tid == 0) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) tot += s.red[w];
    p.partials[blockIdx.x] = tot;
    __threadfence();
    const unsigned int done = atomicAdd(p.ticket, 1u);
    if (done == gridDim.x - 1) {
      __threadfence();
      float sum = 0.f;
      for (unsigned int b = 0; b < gridDim.x; ++b) sum += *reinterpret_cast<volatile float*>(p.partials + b);
      p.loss[0] = sum;
      *p.ticket = 0u;                                     // ready for the next launch
    }
  }
}

constexpr int kMaxGrid = 1024;

}  // namespace top
}  // namespace gs

using namespace gs;

extern "C" size_t gs_sage_top_workspace_bytes(void) { return (top::kMaxGrid + 4) * sizeof(float); }

extern "C" int gs_sage_top_sup(const float* table, int64_t ld_table, const int32_t* nbr_idx, int32_t stride,
                               const int32_t* cnt, const int32_t* self_idx, const int32_t* num_rows_dev, int32_t max_rows,
                               const float* weight, int64_t ldw, int32_t dim, int32_t out_dim, int32_t gcn,
                               const float* cls_w, const float* cls_b, int32_t num_classes, const int64_t* labels,
                               const int32_t* label_index, float* out_h, int64_t ld_h, float* out_agg, int64_t ld_agg,
                               float* out_dz, int64_t ld_dz, float* out_dlog, int64_t ld_dlog, float* logp,
                               float* loss, float* grad_cls_w,
                               float* grad_cls_b, float* grad_table, int64_t ld_gt, void* workspace,
                               size_t workspace_bytes, int32_t precision, float* cls_w_replicas, float* cls_b_replicas,
                               int32_t cls_reps, gs_stream_t stream) {
  if (!table || !nbr_idx || !cnt || !weight || !cls_w || !labels || !loss || max_rows < 0) return GS_ERR_BAD_ARG;
  if (dim != top::kH || out_dim != top::kH || num_classes < 1 || num_classes > top::kMaxClasses || stride < 1 ||
      stride > top::kMaxStride)
    return GS_ERR_UNSUPPORTED;                            // the caller runs the layer as separate kernels instead
  if (precision != GS_PREC_TF32X3 && precision != GS_PREC_TF32) return GS_ERR_UNSUPPORTED;
  const int k = gcn ? top::kH : 2 * top::kH;
  if (ldw < k || ld_table < top::kH) return GS_ERR_BAD_ARG;
  if ((ld_table & 3) || (ldw & 3) || !aligned16(table) || !aligned16(weight) || !aligned16(cls_w)) return GS_ERR_ALIGNMENT;
  if ((out_h && ((ld_h & 1) || (reinterpret_cast<uintptr_t>(out_h) & 7u))) || (out_agg && ((ld_agg & 3) || !aligned16(out_agg))) ||
      (out_dz && ((ld_dz & 3) || !aligned16(out_dz))) || (grad_table && ((ld_gt & 3) || !aligned16(grad_table))) ||
      (grad_cls_w && !aligned16(grad_cls_w)) || (out_dlog && ((ld_dlog & 3) || !aligned16(out_dlog))))
    return GS_ERR_ALIGNMENT;
  if (out_dlog && ld_dlog < top::kMaxClasses) return GS_ERR_BAD_ARG;
  if (!workspace || workspace_bytes < gs_sage_top_workspace_bytes()) return GS_ERR_WORKSPACE;
  if (cls_reps < 1 || cls_reps > 16 || (cls_reps > 1 && (!cls_w_replicas || !cls_b_replicas || !aligned16(cls_w_replicas))))
    return GS_ERR_BAD_ARG;
  if (cls_reps > 1 && ((num_classes * top::kH) & 3)) return GS_ERR_BAD_ARG;
  if (max_rows == 0) return GS_OK;
  top::Params p{table, ld_table, nbr_idx, stride, cnt, self_idx, num_rows_dev, max_rows, weight, ldw, cls_w, cls_b,
                num_classes, labels, label_index, out_h, ld_h, out_agg, ld_agg, out_dz, ld_dz, out_dlog, ld_dlog, logp, loss, grad_cls_w,
                grad_cls_b, cls_w_replicas, cls_b_replicas, cls_reps, grad_table, ld_gt, reinterpret_cast<float*>(workspace) + 4,
                reinterpret_cast<unsigned int*>(workspace), early_reads() ? 1 : 0};
  int tiles = (max_rows + top::kTM - 1) / top::kTM;
  // one tile per CTA while the tiles fit one wave; beyond that a persistent grid (W stays in shared memory)
  int grid = tiles <= kNumSMs ? tiles : kNumSMs;
  if (grid > top::kMaxGrid) grid = top::kMaxGrid;
  const bool split3 = precision == GS_PREC_TF32X3;
  // the weights travel by TMA: W as K/32 boxes of [128 rows x 32 columns], Wc as 4 boxes of [64 x 32] (rows beyond
  // num_classes are out of range and arrive as zeros)
  CUtensorMap tmap_w, tmap_wc;
  memset(&tmap_w, 0, sizeof(tmap_w));
  memset(&tmap_wc, 0, sizeof(tmap_wc));
  if (!tc::make_tmap_2d(&tmap_w, weight, top::kH, k, ldw, top::kH) ||
      !tc::make_tmap_2d(&tmap_wc, cls_w, num_classes, top::kH, top::kH, top::kMaxClasses))
    return GS_ERR_UNSUPPORTED;
#define GS_TOP_LAUNCH(GCN, SP)                                                                                      \
  do {                                                                                                              \
    const int smem = static_cast<int>(sizeof(top::Smem<GCN>)) + 1024;                                               \
    cudaError_t e = cudaFuncSetAttribute(top::sage_top_sup_kernel<GCN, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
    if (e != cudaSuccess) return static_cast<int>(e);                                                               \
    launch(top::sage_top_sup_kernel<GCN, SP>, dim3(grid), dim3(top::kThreads), smem, as_stream(stream), p, tmap_w, tmap_wc); \
  } while (0)
  if (gcn) { if (split3) GS_TOP_LAUNCH(true, true); else GS_TOP_LAUNCH(true, false); }
  else { if (split3) GS_TOP_LAUNCH(false, true); else GS_TOP_LAUNCH(false, false); }
#undef GS_TOP_LAUNCH
  return finish_launch();
}

#ifdef GS_TOP_TRACE
extern "C" int gs_debug_top_trace_read(long long* host_out, int n) {
  if (n > 16) n = 16;
  cudaDeviceSynchronize();
  return static_cast<int>(cudaMemcpyFromSymbol(host_out, gs::top::g_top_trace, sizeof(long long) * n));
}
#endif
