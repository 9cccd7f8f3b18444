// K4 forward over DENSE operands, fed entirely by TMA:  out = relu(X . W^T)   (src/models.py:215-219)
//
// X [rows x K] is the SageLayer's input row [self | agg] as the aggregation kernel wrote it (gs_agg_fwd_x), W [H x K]
// the layer's weight; both come with their low halves x - trunc_tf32(x) in second buffers (X_lo from the same
// aggregation launch, W_lo from the update kernel), so the fp32-faithful 3-term product
//     D = X_lo.W_hi + X_hi.W_lo + X_hi.W_hi       (kind::tf32 reads the top 19 bits: "hi" is the raw fp32 word)
// needs no conversion pass in this kernel at all.  What is left is the textbook Blackwell pipeline:
//     warp 0, one thread   TMA producer: per 32-wide k-stage four tiled bulk copies (X_hi, X_lo: 128 x 32; W_hi, W_lo:
//                          N x 32) land in SWIZZLE_128B tiles and complete the stage's mbarrier by byte count
//     warp 1, one thread   tcgen05.mma issuer: 4 k8-steps x 3 products per stage into one fp32 accumulator in TMEM;
//                          tcgen05.commit hands the stage back to the producer
//     warps 2..9           epilogue: tcgen05.ld -> shared memory (transpose) -> ReLU -> coalesced 512-byte row stores,
//                          and the zero fill of the buffer the backward pass scatters d(out) into
// The gathered-operand kernel (sage_gemm_tc.cu) spends its time issuing 16-byte cp.async copies (2048 per stage) and
// splitting tiles in place; this one issues 4 copies per stage.  Used for layer 1 of a train step, whose X is dense.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace gs {
namespace tc {
bool make_tmap_2d(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);
}
namespace tma {

constexpr int kTileM = 128;
constexpr int kBK = 32;                       // fp32 elements per k-stage = one 128-byte swizzle row
constexpr int kEpiWarps = 8;
constexpr int kThreads = (2 + kEpiWarps) * 32;
constexpr int kMaxStages = 4;
constexpr int kSmemBudget = 200 * 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major SWIZZLE_128B operand descriptor: start >> 4 | LBO 16 B | SBO 1024 B (8-row groups) | version 1 | layout 2
__device__ __forceinline__ uint64_t make_desc_k(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(16u >> 4) << 16;
  d |= static_cast<uint64_t>(1024u >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// c = F32, a = b = TF32, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

#ifdef GS_TOP_TRACE
__device__ long long g_tma_trace[32];
#define GS_TMA_MARK(slot) do { if (blockIdx.x == 0) g_tma_trace[(slot)] = clock64(); } while (0)
#else
#define GS_TMA_MARK(slot) do { } while (0)
#endif

struct Args {
  const int32_t* num_rows_dev; int max_rows;
  float* out; int64_t ld_out; int out_dim, relu;
  float* zero_out; int64_t ld_zero;
  int n_tile, k_stages, num_stages;
};

template <bool SPLIT3>
__global__ void __launch_bounds__(kThreads, 1)
sage_fwd_tma_kernel(const __grid_constant__ CUtensorMap tm_x_hi, const __grid_constant__ CUtensorMap tm_x_lo,
                    const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo, const Args a) {
  if (threadIdx.x == 0) GS_TMA_MARK(0);
  pdl_sync();
  if (threadIdx.x == 0) GS_TMA_MARK(1);
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t s_full[kMaxStages];
  __shared__ __align__(8) uint64_t s_empty[kMaxStages];
  __shared__ __align__(8) uint64_t s_acc;
  __shared__ uint32_t s_tmem;
  const int rows = live_rows(a.num_rows_dev, a.max_rows);
  const int row0 = blockIdx.x * kTileM;
  if (row0 >= rows) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const int a_bytes = kTileM * 128, b_bytes = a.n_tile * 128;
  const int stage_bytes = (SPLIT3 ? 2 : 1) * (a_bytes + b_bytes);
  const uint32_t tmem_cols = a.n_tile <= 32 ? 32u : a.n_tile <= 64 ? 64u : 128u;

  if (tid == 0) {
    for (int s = 0; s < a.num_stages; ++s) {
      mbar_init(smem_u32(&s_full[s]), 1);
      mbar_init(smem_u32(&s_empty[s]), 1);
    }
    mbar_init(smem_u32(&s_acc), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&s_tmem), tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = s_tmem;
  if (threadIdx.x == 0) GS_TMA_MARK(2);

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      for (int ks = 0; ks < a.k_stages; ++ks) {
        const int s = ks % a.num_stages;
        if (ks >= a.num_stages) mbar_wait(smem_u32(&s_empty[s]), static_cast<uint32_t>(ks / a.num_stages - 1) & 1u);
        const uint32_t bar = smem_u32(&s_full[s]);
        const uint32_t a_hi = smem_base + s * stage_bytes;
        const uint32_t b_hi = a_hi + (SPLIT3 ? 2 : 1) * a_bytes;
        mbar_expect_tx(bar, static_cast<uint32_t>(stage_bytes));
        tma_load_2d(a_hi, &tm_x_hi, ks * kBK, row0, bar);
        tma_load_2d(b_hi, &tm_w_hi, ks * kBK, 0, bar);
        if (SPLIT3) {
          tma_load_2d(a_hi + a_bytes, &tm_x_lo, ks * kBK, row0, bar);
          tma_load_2d(b_hi + b_bytes, &tm_w_lo, ks * kBK, 0, bar);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = make_idesc(kTileM, a.n_tile);
    for (int ks = 0; ks < a.k_stages; ++ks) {
      const int s = ks % a.num_stages;
      mbar_wait(smem_u32(&s_full[s]), static_cast<uint32_t>(ks / a.num_stages) & 1u);
      tc_fence_after();
      if (lane == 0) {
        GS_TMA_MARK(8 + ks);
        const uint32_t a_hi = smem_base + s * stage_bytes;
        const uint32_t a_lo = a_hi + a_bytes;
        const uint32_t b_hi = a_hi + (SPLIT3 ? 2 : 1) * a_bytes;
        const uint32_t b_lo = b_hi + b_bytes;
#pragma unroll
        for (int kk = 0; kk < kBK / 8; ++kk) {     // UMMA_K = 8 tf32 = 32 bytes
          const uint32_t off = kk * 32;
          const uint32_t first = (ks == 0 && kk == 0) ? 0u : 1u;
          if (SPLIT3) {
            umma_tf32(tmem_acc, make_desc_k(a_lo + off), make_desc_k(b_hi + off), idesc, first);
            umma_tf32(tmem_acc, make_desc_k(a_hi + off), make_desc_k(b_lo + off), idesc, 1u);
            umma_tf32(tmem_acc, make_desc_k(a_hi + off), make_desc_k(b_hi + off), idesc, 1u);
          } else {
            umma_tf32(tmem_acc, make_desc_k(a_hi + off), make_desc_k(b_hi + off), idesc, first);
          }
        }
        umma_commit(smem_u32(&s_empty[s]));        // frees the stage when these MMAs have read it
        if (ks == a.k_stages - 1) umma_commit(smem_u32(&s_acc));
      }
      __syncwarp();
    }
    tc_fence_before();
  } else {
    // ================= epilogue: TMEM -> registers -> smem (transpose) -> global =================
    const int e = warp - 2;                        // 0..7
    const int quad = warp & 3;                     // a warp may only touch TMEM lanes 32*(warp%4) .. +31
    const int grp = e >> 2;                        // which half of the 32-column chunks
    mbar_wait(smem_u32(&s_acc), 0);                // every MMA has completed: the operand ring is idle
    tc_fence_after();
    if (warp == 2 && lane == 0) GS_TMA_MARK(3);
    const int m = quad * 32 + lane;
    const int n_chunks = (a.n_tile + 31) / 32;
    const int ldst = a.n_tile + 4;                 // floats per staged row (+4: rows land on different banks)
    float* stg = reinterpret_cast<float*>(smem_dyn + (smem_base - smem_u32(smem_dyn)));
    for (int c = grp; c < n_chunks; c += 2) {
      uint32_t v[32];
      tmem_ld32(tmem_acc + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(c * 32), v);
      float* dst = stg + m * ldst + c * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        if (c * 32 + j < a.n_tile)
          *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                            __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    }
    tc_fence_before();
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
    if (warp == 2 && lane == 0) GS_TMA_MARK(4);
    const int n4 = a.n_tile >> 2;
    for (int q = lane; q < n4; q += 32) {
      const int h = 4 * q;
#pragma unroll 4
      for (int mm = e; mm < kTileM; mm += kEpiWarps) {
        const int r = row0 + mm;
        if (r >= rows || h >= a.out_dim) continue;
        float4 v = *reinterpret_cast<const float4*>(stg + mm * ldst + h);
        if (a.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        *reinterpret_cast<float4*>(a.out + static_cast<int64_t>(r) * a.ld_out + h) = v;      // out_dim % 4 == 0, ld_out % 4 == 0
        if (a.zero_out) *reinterpret_cast<float4*>(a.zero_out + static_cast<int64_t>(r) * a.ld_zero + h) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) GS_TMA_MARK(5);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, tmem_cols);
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// The same pipeline for the GATHERED operand of a SageLayer: X[r] = [ self_table[self_idx[r]] | agg[r] ] is never
// materialised -- the self half arrives by TMA tile::gather4 (four table rows named by index per instruction, 32-column
// box; one instruction per lane of the producer warp fills the 128-row tile), the aggregate half by one tiled copy.
// K runs over the self columns first (ceil(dim/32) stages, the tail of the last one zero-filled by the tensor map's
// bounds), then over the aggregate columns; W's box starts at column 0 / dim of the matching half, so a product of
// a zero-filled X column with whatever W column lies there contributes nothing.  X has no low half in memory: the
// eight epilogue warps, idle during the main loop, write lo = x - trunc_tf32(x) of every landed stage next to it (a
// position-independent pass over the 16 KB tile) and hand the stage to the MMA issuer; W's low half comes from the
// buffer the update kernel maintains (or is split the same way when there is none).
// Measured (scratch/tma_gather_probe.py, trace build): a gather4 costs the TMA unit ~70 cycles, 2.3K cycles for the 32 of
// a stage, against 1.2K cycles of MMA issue per stage -- the self stages would set the pace.  By default
// (self_by_threads) the eight epilogue warps therefore fetch the self half themselves, four 16-byte cp.async per thread
// and stage straight into the swizzled tile, two stages ahead, and split their own pieces when they land; TMA keeps
// the aggregate half and W.  (The thread-staged kernel of sage_gemm_tc.cu moves BOTH halves and, without a W_lo
// buffer, W's split through the LSU pipes: 2048 pieces per stage.)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_gather4(uint32_t dst_smem, const CUtensorMap* map, int c0, int r0, int r1, int r2, int r3,
                                            uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float4 tf32_lo4(float4 v) {
  return make_float4(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u), v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u),
                     v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u), v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u));
}

struct GatherArgs {
  const int32_t* self_idx; int dim, ks_self, ks_agg, has_wlo;
  const float* self_table; int64_t ld_self; int self_by_threads;     // see the kernel: who fetches the self half
  int early_idx;                                                     // gs_set_early_reads at launch time
};
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes));
}

template <bool SPLIT3>
__global__ void __launch_bounds__(kThreads, 1)
sage_fwd_tma_gather_kernel(const __grid_constant__ CUtensorMap tm_self, const __grid_constant__ CUtensorMap tm_agg,
                           const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                           const Args a, const GatherArgs ga) {
  if (threadIdx.x == 0) GS_TMA_MARK(0);
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t s_full[kMaxStages];     // the stage's TMA copies have landed
  __shared__ __align__(8) uint64_t s_ready[kMaxStages];    // ... and its low halves are written / its self rows fetched
  __shared__ __align__(8) uint64_t s_empty[kMaxStages];
  __shared__ __align__(8) uint64_t s_acc;
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) int32_t s_idx[kTileM];
  const int row0 = blockIdx.x * kTileM;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* smem_al = smem_dyn + (smem_base - smem_u32(smem_dyn));
  const int a_bytes = kTileM * 128, b_bytes = a.n_tile * 128;
  const int stage_bytes = (SPLIT3 ? 2 : 1) * (a_bytes + b_bytes);
  const uint32_t tmem_cols = a.n_tile <= 32 ? 32u : a.n_tile <= 64 ? 64u : 128u;

  // everything that reads no global memory happens before the wait for the previous kernel of the stream
  if (tid == 0) {
    for (int s = 0; s < a.num_stages; ++s) {
      mbar_init(smem_u32(&s_full[s]), 1);
      mbar_init(smem_u32(&s_ready[s]), kEpiWarps);
      mbar_init(smem_u32(&s_empty[s]), 1);
    }
    mbar_init(smem_u32(&s_acc), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&s_tmem), tmem_cols);
  if (!ga.early_idx) pdl_sync();    // gs_set_early_reads: the row count and the index list may be read before the wait
  const int rows = live_rows(a.num_rows_dev, a.max_rows);
  if (warp == 0 && row0 < rows) {   // the tile's table rows (rows past the live ones read row 0: their results are dropped)
#pragma unroll
    for (int i = 0; i < kTileM / 32; ++i) {
      const int r = row0 + lane + 32 * i;
      s_idx[lane + 32 * i] = r < rows ? (ga.self_idx ? __ldg(ga.self_idx + r) : r) : 0;
    }
  }
  if (ga.early_idx) pdl_sync();
  if (threadIdx.x == 0) GS_TMA_MARK(1);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (row0 >= rows) {               // a tile beyond the live rows: nothing to do but hand the TMEM columns back
    if (warp == 1) tmem_dealloc(s_tmem, tmem_cols);
    return;
  }
  const uint32_t tmem_acc = s_tmem;
  if (threadIdx.x == 0) GS_TMA_MARK(2);
  // the epilogue warps take part in the main loop (and the MMA issuer waits for them) when there are low halves to
  // write or self rows to fetch
  const bool staged = SPLIT3 || (ga.self_by_threads && ga.ks_self > 0);

  if (warp == 0) {
    // ================= TMA producer: all 32 lanes issue (one gather4 each fills the 128 rows of a self stage) =================
    const int4 my = *reinterpret_cast<const int4*>(&s_idx[4 * lane]);
    const uint32_t tx_w = static_cast<uint32_t>(b_bytes * ((SPLIT3 && ga.has_wlo) ? 2 : 1));
    for (int ks = 0; ks < a.k_stages; ++ks) {
      const int s = ks % a.num_stages;
      if (ks >= a.num_stages) mbar_wait(smem_u32(&s_empty[s]), static_cast<uint32_t>(ks / a.num_stages - 1) & 1u);
      const uint32_t bar = smem_u32(&s_full[s]);
      const uint32_t a_hi = smem_base + s * stage_bytes;
      const uint32_t b_hi = a_hi + (SPLIT3 ? 2 : 1) * a_bytes;
      const bool self_stage = ks < ga.ks_self;
      const bool a_by_tma = !(self_stage && ga.self_by_threads);
      if (lane == 0) { GS_TMA_MARK(24 + ks); mbar_expect_tx(bar, tx_w + (a_by_tma ? static_cast<uint32_t>(a_bytes) : 0u)); }
      __syncwarp();
      const int wcol = self_stage ? ks * kBK : (ga.ks_self > 0 ? ga.dim : 0) + (ks - ga.ks_self) * kBK;
      if (self_stage) { if (a_by_tma) tma_gather4(a_hi + lane * 512, &tm_self, ks * kBK, my.x, my.y, my.z, my.w, bar); }
      else if (lane == 0) tma_load_2d(a_hi, &tm_agg, (ks - ga.ks_self) * kBK, row0, bar);
      if (lane == 1) {
        tma_load_2d(b_hi, &tm_w_hi, wcol, 0, bar);
        if (SPLIT3 && ga.has_wlo) tma_load_2d(b_hi + b_bytes, &tm_w_lo, wcol, 0, bar);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = make_idesc(kTileM, a.n_tile);
    for (int ks = 0; ks < a.k_stages; ++ks) {
      const int s = ks % a.num_stages;
      mbar_wait(smem_u32(staged ? &s_ready[s] : &s_full[s]), static_cast<uint32_t>(ks / a.num_stages) & 1u);
      tc_fence_after();
      if (lane == 0) {
        GS_TMA_MARK(8 + ks);
        const uint32_t a_hi = smem_base + s * stage_bytes;
        const uint32_t a_lo = a_hi + a_bytes;
        const uint32_t b_hi = a_hi + (SPLIT3 ? 2 : 1) * a_bytes;
        const uint32_t b_lo = b_hi + b_bytes;
        // the last stage of a half holds dim % 32 live columns (the rest is zero fill): only the k8-steps that see them
        const int k_local = (ks < ga.ks_self ? ks : ks - ga.ks_self) * kBK;
        const int kk_n = min(kBK, ga.dim - k_local + 7) / 8;
#pragma unroll
        for (int kk = 0; kk < kBK / 8; ++kk) {     // UMMA_K = 8 tf32 = 32 bytes
          if (kk >= kk_n) break;
          const uint32_t off = kk * 32;
          const uint32_t first = (ks == 0 && kk == 0) ? 0u : 1u;
          if (SPLIT3) {
            umma_tf32(tmem_acc, make_desc_k(a_lo + off), make_desc_k(b_hi + off), idesc, first);
            umma_tf32(tmem_acc, make_desc_k(a_hi + off), make_desc_k(b_lo + off), idesc, 1u);
            umma_tf32(tmem_acc, make_desc_k(a_hi + off), make_desc_k(b_hi + off), idesc, 1u);
          } else {
            umma_tf32(tmem_acc, make_desc_k(a_hi + off), make_desc_k(b_hi + off), idesc, first);
          }
        }
        umma_commit(smem_u32(&s_empty[s]));        // frees the stage when these MMAs have read it
        if (ks == a.k_stages - 1) umma_commit(smem_u32(&s_acc));
      }
      __syncwarp();
    }
    tc_fence_before();
  } else {
    const int e = warp - 2;                        // 0..7
    const int t = tid - 64;                        // 0..255
    if (staged) {
      // ================= self rows by cp.async (two stages ahead), low halves of every landed stage =================
      const bool fetch = ga.self_by_threads && ga.ks_self > 0;
      const int ahead = a.num_stages - 1;
      const int row = t >> 1, c0 = (t & 1) * 4;                 // my four 16-byte pieces of a self stage: row, chunks c0..c0+3
      const float* src_row = fetch ? ga.self_table + static_cast<int64_t>(s_idx[row]) * ga.ld_self : nullptr;
      uint32_t piece[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) piece[j] = static_cast<uint32_t>(row * 128 + (((c0 + j) ^ (row & 7)) << 4));
      int issued = 0;
      auto issue_self = [&]() {                                 // always commits exactly one group (empty past the self stages)
        if (fetch && issued < ga.ks_self) {
          const int s = issued % a.num_stages;
          if (issued >= a.num_stages) mbar_wait(smem_u32(&s_empty[s]), static_cast<uint32_t>(issued / a.num_stages - 1) & 1u);
          const uint32_t a_hi = smem_base + s * stage_bytes;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = issued * kBK + 4 * (c0 + j);
            const bool in = col < ga.dim;                       // dim % 4 == 0: a piece is whole or zero
            cp_async16(a_hi + piece[j], in ? static_cast<const void*>(src_row + col) : static_cast<const void*>(ga.self_table), in ? 16u : 0u);
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        ++issued;
      };
      if (fetch) for (int i = 0; i < ahead; ++i) issue_self();
      for (int ks = 0; ks < a.k_stages; ++ks) {
        const int s = ks % a.num_stages;
        const bool mine = fetch && ks < ga.ks_self;             // this stage's A tile came through my own copies
        if (mine) {
          if (ahead >= 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
          else asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        mbar_wait(smem_u32(&s_full[s]), static_cast<uint32_t>(ks / a.num_stages) & 1u);    // the TMA copies of the stage
        if (t == 0) GS_TMA_MARK(16 + ks);
        unsigned char* a_hi = smem_al + s * stage_bytes;
        if (SPLIT3) {
          if (mine) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<float4*>(a_hi + a_bytes + piece[j]) = tf32_lo4(*reinterpret_cast<const float4*>(a_hi + piece[j]));
          } else {
            float4 v[kTileM * 8 / (kEpiWarps * 32)];
#pragma unroll
            for (int i = 0; i < kTileM * 8 / (kEpiWarps * 32); ++i) v[i] = *reinterpret_cast<const float4*>(a_hi + (t + i * kEpiWarps * 32) * 16);
#pragma unroll
            for (int i = 0; i < kTileM * 8 / (kEpiWarps * 32); ++i)
              *reinterpret_cast<float4*>(a_hi + a_bytes + (t + i * kEpiWarps * 32) * 16) = tf32_lo4(v[i]);
          }
          if (!ga.has_wlo) {
            unsigned char* b_hi = a_hi + 2 * a_bytes;
            for (int q = t; q < a.n_tile * 8; q += kEpiWarps * 32)
              *reinterpret_cast<float4*>(b_hi + b_bytes + q * 16) = tf32_lo4(*reinterpret_cast<const float4*>(b_hi + q * 16));
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s_ready[s]));
        if (fetch) issue_self();                                // refill the stage the MMA of k-stage ks-1 is about to release
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    // ================= epilogue: TMEM -> registers -> smem (transpose) -> global =================
    const int quad = warp & 3;                     // a warp may only touch TMEM lanes 32*(warp%4) .. +31
    const int grp = e >> 2;                        // which half of the 32-column chunks
    mbar_wait(smem_u32(&s_acc), 0);                // every MMA has completed: the operand ring is idle
    tc_fence_after();
    if (t == 0) GS_TMA_MARK(3);
    const int m = quad * 32 + lane;
    const int n_chunks = (a.n_tile + 31) / 32;
    const int ldst = a.n_tile + 4;                 // floats per staged row (+4: rows land on different banks)
    float* stg = reinterpret_cast<float*>(smem_al);
    for (int c = grp; c < n_chunks; c += 2) {
      uint32_t v[32];
      tmem_ld32(tmem_acc + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(c * 32), v);
      float* dst = stg + m * ldst + c * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        if (c * 32 + j < a.n_tile)
          *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                            __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    }
    tc_fence_before();
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
    const int n4 = a.n_tile >> 2;
    for (int q = lane; q < n4; q += 32) {
      const int h = 4 * q;
#pragma unroll 4
      for (int mm = e; mm < kTileM; mm += kEpiWarps) {
        const int r = row0 + mm;
        if (r >= rows || h >= a.out_dim) continue;
        float4 v = *reinterpret_cast<const float4*>(stg + mm * ldst + h);
        if (a.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        *reinterpret_cast<float4*>(a.out + static_cast<int64_t>(r) * a.ld_out + h) = v;      // out_dim % 4 == 0, ld_out % 4 == 0
        if (a.zero_out) *reinterpret_cast<float4*>(a.zero_out + static_cast<int64_t>(r) * a.ld_zero + h) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) GS_TMA_MARK(5);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, tmem_cols);
  }
}

}  // namespace tma
}  // namespace gs

using namespace gs;

// x_hi / x_lo: [max_rows x kx] (leading dimension ld_x), w_hi / w_lo: [out_dim x kx] (ldw).  The lo operands may be
// NULL for GS_PREC_TF32 only.  Returns GS_ERR_UNSUPPORTED when the shape does not fit this kernel (the caller falls
// back to the gathered-operand kernel).
int gs_sage_gemm_fwd_tma(const float* x_hi, const float* x_lo, int64_t ld_x, int32_t kx, const float* w_hi,
                         const float* w_lo, int64_t ldw, int32_t out_dim, const int32_t* num_rows_dev, int32_t max_rows,
                         float* out, int64_t ld_out, int32_t relu, int32_t precision, float* zero_out, int64_t ld_zero,
                         gs_stream_t stream) {
  const bool split3 = precision == GS_PREC_TF32X3;
  if (precision != GS_PREC_TF32 && !split3) return GS_ERR_UNSUPPORTED;
  if (split3 && (!x_lo || !w_lo)) return GS_ERR_UNSUPPORTED;
  if (out_dim < 16 || out_dim > 128 || (out_dim & 15) || (kx & 3) || kx < 4) return GS_ERR_UNSUPPORTED;
  if ((ld_x & 3) || (ldw & 3) || (ld_out & 3) || !aligned16(x_hi) || !aligned16(w_hi) || !aligned16(out) ||
      (x_lo && !aligned16(x_lo)) || (w_lo && !aligned16(w_lo)) || (zero_out && ((ld_zero & 3) || !aligned16(zero_out))))
    return GS_ERR_UNSUPPORTED;
  const int n_tile = out_dim;
  CUtensorMap tm[4];
  memset(tm, 0, sizeof(tm));
  if (!tc::make_tmap_2d(&tm[0], x_hi, max_rows, kx, ld_x, tma::kTileM) || !tc::make_tmap_2d(&tm[2], w_hi, out_dim, kx, ldw, n_tile))
    return GS_ERR_UNSUPPORTED;
  if (split3) {
    if (!tc::make_tmap_2d(&tm[1], x_lo, max_rows, kx, ld_x, tma::kTileM) || !tc::make_tmap_2d(&tm[3], w_lo, out_dim, kx, ldw, n_tile))
      return GS_ERR_UNSUPPORTED;
  } else {
    tm[1] = tm[0];
    tm[3] = tm[2];
  }
  const int stage = (split3 ? 2 : 1) * (tma::kTileM * 128 + n_tile * 128);
  int stages = tma::kSmemBudget / stage;
  if (stages > tma::kMaxStages) stages = tma::kMaxStages;
  if (stages < 2) return GS_ERR_UNSUPPORTED;
  const int staging = tma::kTileM * (n_tile + 4) * 4;
  const int smem = (stages * stage > staging ? stages * stage : staging) + 1024;
  tma::Args a{num_rows_dev, max_rows, out, ld_out, out_dim, relu, zero_out, ld_zero, n_tile, (kx + tma::kBK - 1) / tma::kBK, stages};
  const dim3 grid((max_rows + tma::kTileM - 1) / tma::kTileM);
  cudaError_t e;
  if (split3) {
    e = cudaFuncSetAttribute(tma::sage_fwd_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    launch(tma::sage_fwd_tma_kernel<true>, grid, dim3(tma::kThreads), smem, as_stream(stream), tm[0], tm[1], tm[2], tm[3], a);
  } else {
    e = cudaFuncSetAttribute(tma::sage_fwd_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    launch(tma::sage_fwd_tma_kernel<false>, grid, dim3(tma::kThreads), smem, as_stream(stream), tm[0], tm[1], tm[2], tm[3], a);
  }
  return finish_launch();
}

// The gathered form: X[r] = [ self_table[self_idx[r]] | agg[r] ] (gcn: agg only), self_idx NULL = identity.  w_lo may be
// NULL (the kernel then splits W itself).  Returns GS_ERR_UNSUPPORTED when the shape does not fit (the caller falls
// back to the thread-staged kernel).
int gs_sage_gemm_fwd_tma_gather(const float* self_table, int64_t ld_self, const int32_t* self_idx, const float* agg,
                                int64_t ld_agg, int32_t dim, const float* w_hi, const float* w_lo, int64_t ldw,
                                int32_t out_dim, int32_t gcn, const int32_t* num_rows_dev, int32_t max_rows, float* out,
                                int64_t ld_out, int32_t relu, int32_t precision, float* zero_out, int64_t ld_zero,
                                gs_stream_t stream) {
  const bool split3 = precision == GS_PREC_TF32X3;
  if (precision != GS_PREC_TF32 && !split3) return GS_ERR_UNSUPPORTED;
  if (out_dim < 16 || out_dim > 128 || (out_dim & 15) || (dim & 3) || dim < 4) return GS_ERR_UNSUPPORTED;
  if ((ld_agg & 3) || (ldw & 3) || (ld_out & 3) || !aligned16(agg) || !aligned16(w_hi) || !aligned16(out) ||
      (w_lo && !aligned16(w_lo)) || (zero_out && ((ld_zero & 3) || !aligned16(zero_out))))
    return GS_ERR_UNSUPPORTED;
  if (!gcn && (!self_table || (ld_self & 3) || !aligned16(self_table))) return GS_ERR_UNSUPPORTED;
  const int n_tile = out_dim, kt = gcn ? dim : 2 * dim;
  CUtensorMap tm[4];
  memset(tm, 0, sizeof(tm));
  // the table's row count is not part of the interface: the map's bound only has to cover every index the lists hold
  if (!tc::make_tmap_2d(&tm[1], agg, max_rows, dim, ld_agg, tma::kTileM) || !tc::make_tmap_2d(&tm[2], w_hi, out_dim, kt, ldw, n_tile))
    return GS_ERR_UNSUPPORTED;
  if (gcn) tm[0] = tm[1];
  else if (!tc::make_tmap_2d(&tm[0], self_table, int64_t{1} << 30, dim, ld_self, 1)) return GS_ERR_UNSUPPORTED;
  if (split3 && w_lo) {
    if (!tc::make_tmap_2d(&tm[3], w_lo, out_dim, kt, ldw, n_tile)) return GS_ERR_UNSUPPORTED;
  } else {
    tm[3] = tm[2];
  }
  const int stage = (split3 ? 2 : 1) * (tma::kTileM * 128 + n_tile * 128);
  int stages = tma::kSmemBudget / stage;
  if (stages > tma::kMaxStages) stages = tma::kMaxStages;
  if (stages < 2) return GS_ERR_UNSUPPORTED;
  const int staging = tma::kTileM * (n_tile + 4) * 4;
  const int smem = (stages * stage > staging ? stages * stage : staging) + 1024;
  const int ks_half = (dim + tma::kBK - 1) / tma::kBK;
  static const int by_threads = [] { const char* e = getenv("GS_TMA_SELF"); return (e && e[0] == 'g') ? 0 : 1; }();   // GS_TMA_SELF=gather4: A/B runs
  tma::GatherArgs ga{self_idx, dim, gcn ? 0 : ks_half, ks_half, (split3 && w_lo) ? 1 : 0, self_table, ld_self, by_threads,
                     early_reads() ? 1 : 0};
  tma::Args a{num_rows_dev, max_rows, out, ld_out, out_dim, relu, zero_out, ld_zero, n_tile, ga.ks_self + ga.ks_agg, stages};
  const dim3 grid((max_rows + tma::kTileM - 1) / tma::kTileM);
  cudaError_t e;
  if (split3) {
    e = cudaFuncSetAttribute(tma::sage_fwd_tma_gather_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    launch(tma::sage_fwd_tma_gather_kernel<true>, grid, dim3(tma::kThreads), smem, as_stream(stream), tm[0], tm[1], tm[2], tm[3], a, ga);
  } else {
    e = cudaFuncSetAttribute(tma::sage_fwd_tma_gather_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    launch(tma::sage_fwd_tma_gather_kernel<false>, grid, dim3(tma::kThreads), smem, as_stream(stream), tm[0], tm[1], tm[2], tm[3], a, ga);
  }
  return finish_launch();
}

#ifdef GS_TOP_TRACE
extern "C" int gs_debug_tma_trace_read(long long* host_out, int n) {
  if (n > 32) n = 32;
  cudaDeviceSynchronize();
  return static_cast<int>(cudaMemcpyFromSymbol(host_out, gs::tma::g_tma_trace, sizeof(long long) * n));
}
#endif
