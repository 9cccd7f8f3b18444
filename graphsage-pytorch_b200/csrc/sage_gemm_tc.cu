// K4 on the 5th-generation tensor cores: tcgen05.mma kind::tf32, accumulators in TMEM.
//
// One templated core serves the three contractions of a SageLayer (src/models.py:215-219 and
// its autograd):
//     forward   out = relu(X . W^T)            A = X  (K-major, gathered),  B = W   (K-major)
//     bwd_x     dX  = dZ . W                   A = dZ (K-major, ReLU mask), B = W   (MN-major)
//     bwd_w     dW += dZ^T . X  (row chunks)   A = dZ (MN-major),           B = X   (MN-major)
// with X the virtual concat [ self_table[self_idx[r]] | agg[r] ] that is never materialised.
//
// Why the operands are staged by threads and not by TMA: the A rows of the self half are a
// gather through an index list and every operand needs an element-wise transform on the way
// (hi/lo split for the fp32-faithful mode), so the 8 producer warps cp.async 16-byte pieces
// (coalesced per row) straight into the 128-byte swizzled layout the UMMA shared-memory
// descriptors expect, num_stages-1 k-stages ahead.  Every thread owns the same pieces in every
// stage, so cp.async.wait_group is the only wait before it splits ITS pieces in place
// (hi = trunc_tf32(x), lo = x - hi); a proxy fence + mbarrier hands the stage to the single
// MMA-issuing thread and tcgen05.commit hands it back.  The accumulator (128 lanes x N fp32
// columns) lives in TMEM, is read once (tcgen05.ld), transposed through the then idle operand
// ring in shared memory and leaves the SM as full coalesced rows (float4 stores / vector REDs).
//
// Precision modes:
//   GS_PREC_TF32   one tf32 product  (10-bit mantissa operands, fp32 accumulate)
//   GS_PREC_TF32X3 x = hi + lo with hi = x truncated to tf32, lo = x - hi (exact in fp32);
//                  D = Alo.Bhi + Ahi.Blo + Ahi.Bhi.  Dropped term and the truncation of lo are
//                  O(2^-22) relative: fp32-faithful, meets the 1e-5 parity bound.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>          // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint)

#include "common.cuh"

namespace gs {
namespace tc {

constexpr int kTileM = 128;
constexpr int kBK = 32;                                  // tf32 elements per k-stage: 128 bytes, one swizzle row
#ifndef GS_TC_PRODUCER_WARPS
#define GS_TC_PRODUCER_WARPS 16
#endif
constexpr int kProducerWarps = GS_TC_PRODUCER_WARPS;     // multiple of 4 (TMEM lane quadrants), divides 32
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kThreads = kProducerThreads + 32;          // producers/epilogue + MMA/TMEM warp
constexpr int kMaxStages = 4;
// 544 threads x 72 registers = 39K of the SM's 64K: a GEMM CTA fits beside two 8-warp aggregation CTAs of
// the preparation branch (2 x 12K), so neither branch of the pipelined step waits for the other's kernel to end
constexpr int kMaxRegs = 72;
constexpr int kSmemBudget = 196 * 1024;
constexpr int kMaxChunkRows = 512;
// kind::tf32 reads the top 19 bits of each 32-bit operand element (the low 13 mantissa bits are
// ignored), so the "hi" operand of the 3-term split is the raw fp32 value and only lo = x - trunc(x)
// has to be written.  Checked by the 1e-5 parity tests: a rounding tensor core would miss them by 1e-4.
#ifndef GS_TC_IMPLICIT_TRUNC
#define GS_TC_IMPLICIT_TRUNC 1
#endif
constexpr bool kImplicitTrunc = GS_TC_IMPLICIT_TRUNC != 0;                       // bwd_w: rows reduced per CTA (index cache size)

// ---------------------------------------------------------------------------------------------
// optional pipeline trace (build with -DGS_TC_TRACE): CTA 0 records SM clock at pipeline events
// ---------------------------------------------------------------------------------------------
#ifdef GS_TC_TRACE
__device__ long long g_trace[512];
#define GS_TRACE(slot) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (threadIdx.x & 31) == 0 && (slot) < 512) g_trace[(slot)] = clock64(); } while (0)
#else
#define GS_TRACE(slot) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
// 16-byte asynchronous global->shared copy (LDGSTS); src_bytes == 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  // no "memory" clobber on purpose: the compiler may hoist the (independent) index loads of the
  // next pieces above this copy; ordering against the mbarrier operations is kept by `volatile`.
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes));
}
// TMA: one thread arms the stage's mbarrier with the byte count of the box and issues the tiled copy; the
// TMA unit writes the box in the 128-byte swizzled layout the UMMA descriptors read and completes the barrier.
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
// D[tmem] (+)= A[smem] . B[smem]^T, tf32 inputs, fp32 accumulate
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

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout type:
// 2 = SWIZZLE_128B (K-major operands), 1 = SWIZZLE_128B_BASE32B (the only layout tf32 accepts
// for MN-major operands).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 (1<<4), a=b=TF32 (2<<7, 2<<10),
// a_major bit 15, b_major bit 16 (1 = MN-major), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// shared-memory operand tiles.  Extent E = rows of the operand (M = 128 or N), kBK = 32 along K.
//   K-major : row e is 128 contiguous bytes (32 k), 8-row groups 1024 B apart, 16-byte pieces
//             XOR-swizzled with (e & 7).                       desc: LBO = 16, SBO = 1024
//   MN-major: SWIZZLE_128B_BASE32B atoms of [32 e x 4 k]: 4 k-rows of 128 B (32 e each), the
//             32-byte granules of a row XOR-swizzled with (k & 3) (Swizzle<2,5,2> on the byte
//             address); atoms ordered [k-atom][e-group].       desc: LBO = 512, SBO = G*512
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int tile_bytes(int extent, bool mn_major) {
  return mn_major ? 4 * ((extent + 31) / 32) * 1024 : ((extent + 7) / 8) * 1024;
}
__device__ __forceinline__ int chunks_in_tile(int extent, bool mn_major) {
  return mn_major ? kBK * ((extent + 31) / 32) * 8 : extent * 8;
}
// chunk q -> (tile-local coordinate along E of its first element, along K) and byte offset
__device__ __forceinline__ void chunk_coords(int q, int extent, bool mn_major, int& e, int& k, int& off) {
  if (!mn_major) {
    e = q >> 3;
    const int c = q & 7;
    k = 4 * c;
    off = e * 128 + ((c ^ (e & 7)) << 4);
  } else {
    const int groups = (extent + 31) / 32;
    const int per_k = groups * 8;
    k = q / per_k;
    const int cc = q - k * per_k;
    const int g = cc >> 3, c = cc & 7;
    e = g * 32 + 4 * c;
    off = ((k >> 2) * groups + g) * 512 + (k & 3) * 128 + ((((c >> 1) ^ (k & 3)) << 5) | ((c & 1) << 4));
  }
}

__device__ __forceinline__ float4 tf32_hi(const float4& v) {
  float4 h;
  h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
  h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
  h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
  h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
  return h;
}

// ---------------------------------------------------------------------------------------------
// operand loaders: load4(e, k, kstage) returns 4 consecutive elements along the operand's
// contiguous dimension (K for K-major, E for MN-major) at tile-local (e, k) of k-stage `ks`.
// ---------------------------------------------------------------------------------------------
struct XView {            // virtual concat operand, see sage_gemm.cu
  const float* self_table; int64_t ld_self; const int32_t* self_idx;
  const float* agg; int64_t ld_agg;
  int dim, dim_pad, gcn;
  // optional shared-memory copy of self_idx[cache_base .. cache_base + cache_n): the loaders look a
  // row index up per 16-byte piece, and a dependent global load there serialises the cp.async stream
  const int32_t* idx_cache; int cache_base, cache_n;
  __device__ __forceinline__ int self_row(int r) const {
    if (!self_idx) return r;
    const int o = r - cache_base;
    if (idx_cache && o >= 0 && o < cache_n) return idx_cache[o];
    return __ldg(self_idx + r);
  }
  __device__ __forceinline__ void fill_cache(int32_t* smem_buf, int base, int n, int rows) {   // all threads; sync after
    idx_cache = nullptr; cache_base = base; cache_n = 0;
    if (gcn || !self_idx) return;
    for (int i = threadIdx.x; i < n; i += blockDim.x) smem_buf[i] = (base + i < rows) ? __ldg(self_idx + base + i) : 0;
    idx_cache = smem_buf; cache_n = n;
  }
  __device__ __forceinline__ int kv_total() const { return gcn ? dim_pad : 2 * dim_pad; }
  __device__ __forceinline__ int wcol(int kv) const {
    if (kv < dim) return kv;
    if (gcn) return -1;
    const int k2 = kv - dim_pad;
    return (k2 >= 0 && k2 < dim) ? dim + k2 : -1;
  }
  __device__ __forceinline__ float4 load4(int r, int kv) const {          // kv % 4 == 0, kv < kv_total
    if (gcn) return *reinterpret_cast<const float4*>(agg + static_cast<int64_t>(r) * ld_agg + kv);
    if (kv < dim_pad) {
      const int sr = self_row(r);
      return *reinterpret_cast<const float4*>(self_table + static_cast<int64_t>(sr) * ld_self + kv);
    }
    return *reinterpret_cast<const float4*>(agg + static_cast<int64_t>(r) * ld_agg + (kv - dim_pad));
  }
  __device__ __forceinline__ const float* ptr4(int r, int kv) const {      // address of the same 16 bytes
    if (gcn) return agg + static_cast<int64_t>(r) * ld_agg + kv;
    if (kv < dim_pad) {
      const int sr = self_row(r);
      return self_table + static_cast<int64_t>(sr) * ld_self + kv;
    }
    return agg + static_cast<int64_t>(r) * ld_agg + (kv - dim_pad);
  }
};

__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

struct LoadX_K {           // forward A: X rows, K-major
  XView x; int row0, rows;
  __device__ __forceinline__ float4 operator()(int e, int k, int ks) const {
    const int r = row0 + e, kv = ks * kBK + k;
    if (r >= rows || kv >= x.kv_total()) return zero4();
    return x.load4(r, kv);
  }
  __device__ __forceinline__ const float* ptr(int e, int k, int ks) const {     // nullptr => zero fill
    const int r = row0 + e, kv = ks * kBK + k;
    if (r >= rows || kv >= x.kv_total()) return nullptr;
    return x.ptr4(r, kv);
  }
  // per-piece state computed once: the piece's source is affine in the k-stage
  struct Prep { const float* s; const float* a; int k; };
  __device__ __forceinline__ Prep prep(int e, int k) const {
    const int r = row0 + e;
    if (r >= rows) return Prep{nullptr, nullptr, 1 << 30};
    const float* a = x.agg + static_cast<int64_t>(r) * x.ld_agg + k - (x.gcn ? 0 : x.dim_pad);
    const float* sp = x.gcn ? a : x.self_table + static_cast<int64_t>(x.self_row(r)) * x.ld_self + k;
    return Prep{sp, a, k};
  }
  __device__ __forceinline__ const float* ptr(const Prep& p, int ks) const {
    const int kv = ks * kBK + p.k;
    if (kv >= x.kv_total()) return nullptr;                  // also rows beyond the tile (k = 2^30)
    return (kv < x.dim_pad ? p.s : p.a) + ks * kBK;
  }
};
struct LoadW_K {           // forward B: W[h, wcol(kv)], K-major
  XView x; const float* w; int64_t ldw; int h0, out_dim; bool vec_ok;
  __device__ __forceinline__ float4 operator()(int e, int k, int ks) const {
    const int h = h0 + e, kv = ks * kBK + k;
    if (h >= out_dim || kv >= x.kv_total()) return zero4();
    const float* row = w + static_cast<int64_t>(h) * ldw;
    if (vec_ok) return __ldg(reinterpret_cast<const float4*>(row + kv));          // dim % 4 == 0: wcol(kv) == kv
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { const int c = x.wcol(kv + i); v[i] = c >= 0 ? __ldg(row + c) : 0.f; }
    return make_float4(v[0], v[1], v[2], v[3]);
  }
  __device__ __forceinline__ const float* ptr(int e, int k, int ks) const {     // requires vec_ok
    const int h = h0 + e, kv = ks * kBK + k;
    if (h >= out_dim || kv >= x.kv_total()) return nullptr;
    return w + static_cast<int64_t>(h) * ldw + kv;
  }
  struct Prep { const float* p; int k; };
  __device__ __forceinline__ Prep prep(int e, int k) const {
    const int h = h0 + e;
    if (h >= out_dim) return Prep{nullptr, 1 << 30};
    return Prep{w + static_cast<int64_t>(h) * ldw + k, k};
  }
  __device__ __forceinline__ const float* ptr(const Prep& p, int ks) const {
    return ks * kBK + p.k < x.kv_total() ? p.p + ks * kBK : nullptr;
  }
};
struct LoadDZ_K {          // bwd_x A: dZ rows (ReLU mask), K-major over h
  const float* go; int64_t ld_go; const float* out; int64_t ld_out; int row0, rows, out_dim, relu;
  __device__ __forceinline__ float4 operator()(int e, int k, int ks) const {
    const int r = row0 + e, h = ks * kBK + k;
    if (r >= rows || h >= out_dim) return zero4();
    float g[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      g[i] = 0.f;
      if (h + i < out_dim) {
        g[i] = go[static_cast<int64_t>(r) * ld_go + h + i];
        if (relu && !(out[static_cast<int64_t>(r) * ld_out + h + i] > 0.f)) g[i] = 0.f;
      }
    }
    return make_float4(g[0], g[1], g[2], g[3]);
  }
  __device__ __forceinline__ const float* ptr(int e, int k, int ks) const {     // requires relu == 0, out_dim % 4 == 0
    const int r = row0 + e, h = ks * kBK + k;
    if (r >= rows || h >= out_dim) return nullptr;
    return go + static_cast<int64_t>(r) * ld_go + h;
  }
  struct Prep { const float* p; int k; };
  __device__ __forceinline__ Prep prep(int e, int k) const {
    const int r = row0 + e;
    if (r >= rows) return Prep{nullptr, 1 << 30};
    return Prep{go + static_cast<int64_t>(r) * ld_go + k, k};
  }
  __device__ __forceinline__ const float* ptr(const Prep& p, int ks) const {
    return ks * kBK + p.k < out_dim ? p.p + ks * kBK : nullptr;
  }
};
struct LoadW_MN {          // bwd_x B: W[h = k, c = n..n+3], MN-major (c contiguous)
  const float* w; int64_t ldw; int c0, ncols, out_dim; bool vec_ok;
  __device__ __forceinline__ float4 operator()(int e, int k, int ks) const {
    const int h = ks * kBK + k, c = c0 + e;
    if (h >= out_dim || c >= ncols) return zero4();
    const float* row = w + static_cast<int64_t>(h) * ldw;
    if (vec_ok && c + 3 < ncols) return __ldg(reinterpret_cast<const float4*>(row + c));
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (c + i < ncols) ? __ldg(row + c + i) : 0.f;
    return make_float4(v[0], v[1], v[2], v[3]);
  }
  __device__ __forceinline__ const float* ptr(int e, int k, int ks) const {     // requires vec_ok, ncols % 4 == 0
    const int h = ks * kBK + k, c = c0 + e;
    if (h >= out_dim || c >= ncols) return nullptr;
    return w + static_cast<int64_t>(h) * ldw + c;
  }
  struct Prep { const float* p; int k; };
  __device__ __forceinline__ Prep prep(int e, int k) const {
    const int c = c0 + e;
    if (c >= ncols) return Prep{nullptr, 1 << 30};
    return Prep{w + static_cast<int64_t>(k) * ldw + c, k};
  }
  __device__ __forceinline__ const float* ptr(const Prep& p, int ks) const {
    return ks * kBK + p.k < out_dim ? p.p + static_cast<int64_t>(ks * kBK) * ldw : nullptr;
  }
};
struct LoadDZ_MN {         // bwd_w A: dZ[r = k, h = m..m+3], MN-major (h contiguous)
  const float* go; int64_t ld_go; const float* out; int64_t ld_out; int r_begin, r_end, h0, out_dim, relu;
  __device__ __forceinline__ float4 operator()(int e, int k, int ks) const {
    const int r = r_begin + ks * kBK + k, h = h0 + e;
    if (r >= r_end || h >= out_dim) return zero4();
    float g[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      g[i] = 0.f;
      if (h + i < out_dim) {
        g[i] = go[static_cast<int64_t>(r) * ld_go + h + i];
        if (relu && !(out[static_cast<int64_t>(r) * ld_out + h + i] > 0.f)) g[i] = 0.f;
      }
    }
    return make_float4(g[0], g[1], g[2], g[3]);
  }
  __device__ __forceinline__ const float* ptr(int e, int k, int ks) const {     // requires relu == 0, out_dim % 4 == 0
    const int r = r_begin + ks * kBK + k, h = h0 + e;
    if (r >= r_end || h >= out_dim) return nullptr;
    return go + static_cast<int64_t>(r) * ld_go + h;
  }
  struct Prep { const float* p; int r; };
  __device__ __forceinline__ Prep prep(int e, int k) const {
    const int h = h0 + e;
    if (h >= out_dim) return Prep{nullptr, 1 << 30};
    return Prep{go + static_cast<int64_t>(r_begin + k) * ld_go + h, r_begin + k};
  }
  __device__ __forceinline__ const float* ptr(const Prep& p, int ks) const {
    return p.r + ks * kBK < r_end ? p.p + static_cast<int64_t>(ks * kBK) * ld_go : nullptr;
  }
};
struct LoadX_MN {          // bwd_w B: X[r = k, kv = n..n+3], MN-major (kv contiguous)
  XView x; int r_begin, r_end, kv0;
  __device__ __forceinline__ float4 operator()(int e, int k, int ks) const {
    const int r = r_begin + ks * kBK + k, kv = kv0 + e;
    if (r >= r_end || kv >= x.kv_total()) return zero4();
    return x.load4(r, kv);
  }
  __device__ __forceinline__ const float* ptr(int e, int k, int ks) const {
    const int r = r_begin + ks * kBK + k, kv = kv0 + e;
    if (r >= r_end || kv >= x.kv_total()) return nullptr;
    return x.ptr4(r, kv);
  }
  // agg half: affine in the k-stage; self half: one index lookup (shared-memory cache) per stage
  struct Prep { const float* a; int r; int kv; };
  __device__ __forceinline__ Prep prep(int e, int k) const {
    const int kv = kv0 + e;
    if (kv >= x.kv_total()) return Prep{nullptr, 1 << 30, 0};
    const bool self_half = !x.gcn && kv < x.dim_pad;
    const float* a = self_half ? nullptr
                               : x.agg + static_cast<int64_t>(r_begin + k) * x.ld_agg + kv - (x.gcn ? 0 : x.dim_pad);
    return Prep{a, r_begin + k, kv};
  }
  __device__ __forceinline__ const float* ptr(const Prep& p, int ks) const {
    const int r = p.r + ks * kBK;
    if (r >= r_end) return nullptr;
    if (p.a) return p.a + static_cast<int64_t>(ks * kBK) * x.ld_agg;
    return x.self_table + static_cast<int64_t>(x.self_row(r)) * x.ld_self + p.kv;
  }
};

// ---------------------------------------------------------------------------------------------
// epilogues: called with tile-local row m (= TMEM lane), tile-local column n (n % 4 == 0) and 4
// consecutive accumulator columns; the lanes of a warp hold consecutive n of ONE row, so global
// accesses are coalesced
// ---------------------------------------------------------------------------------------------
struct StoreOut {          // forward: out[r, h] = relu(acc); optionally zero-fills a second buffer of the same shape
  float* out; int64_t ld_out; int row0, rows, h0, out_dim, relu; float* zero; int64_t ld_zero;
  __device__ __forceinline__ void operator()(int m, int n, float4 v) const {
    const int r = row0 + m, h = h0 + n;
    if (r >= rows || h >= out_dim) return;
    if (zero) {            // the gradient w.r.t. this output: the backward scatter accumulates into it later in the step
      float* z = zero + static_cast<int64_t>(r) * ld_zero + h;
      if (h + 3 < out_dim && ((reinterpret_cast<uintptr_t>(z) & 15u) == 0)) *reinterpret_cast<float4*>(z) = zero4();
      else for (int j = 0; j < 4; ++j) if (h + j < out_dim) z[j] = 0.f;
    }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    float* dst = out + static_cast<int64_t>(r) * ld_out + h;
    if (h + 3 < out_dim && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
      *reinterpret_cast<float4*>(dst) = v;
    } else {
      const float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) if (h + j < out_dim) dst[j] = o[j];
    }
  }
};
struct StoreDX {           // bwd_x: dX[r, c] -> grad_self (c < dim) / grad_agg (c >= dim)
  float* gs; int64_t ld_gs; float* ga; int64_t ld_ga; int row0, rows, c0, dim, ncols, gcn;
  __device__ __forceinline__ void operator()(int m, int n, float4 v) const {
    const int r = row0 + m, c = c0 + n;
    if (r >= rows || c >= ncols) return;
    const bool to_self = !gcn && c < dim;
    float* dst = to_self ? gs + static_cast<int64_t>(r) * ld_gs + c
                         : ga + static_cast<int64_t>(r) * ld_ga + (gcn ? c : c - dim);
    const int lim = to_self ? dim : ncols;                    // a 4-group never straddles the seam when dim % 4 == 0
    if (c + 3 < lim && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
      *reinterpret_cast<float4*>(dst) = v;
    } else {
      const float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cj = c + j;
        if (cj >= ncols) continue;
        if (gcn) ga[static_cast<int64_t>(r) * ld_ga + cj] = o[j];
        else if (cj < dim) gs[static_cast<int64_t>(r) * ld_gs + cj] = o[j];
        else ga[static_cast<int64_t>(r) * ld_ga + (cj - dim)] = o[j];
      }
    }
  }
};
struct AddDW {             // bwd_w: grad_w[h, wcol(kv)] += acc   (row chunks reduced with fp32 REDs)
  XView x; float* gw; int64_t ldw; int h0, out_dim, kv0;
  __device__ __forceinline__ void operator()(int m, int n, float4 v) const {
    const int h = h0 + m, kv = kv0 + n;
    const int kt = x.kv_total();
    if (h >= out_dim || kv >= kt) return;
    float* row = gw + static_cast<int64_t>(h) * ldw;
    const int c = x.wcol(kv);
    if ((x.dim & 3) == 0 && c >= 0 && ((reinterpret_cast<uintptr_t>(row + c) & 15u) == 0)) {
      atomicAdd(reinterpret_cast<float4*>(row + c), v);       // dim % 4 == 0: the 4 columns are consecutive in W
    } else {
      const float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cj = kv + j < kt ? x.wcol(kv + j) : -1;
        if (cj >= 0) atomicAdd(row + cj, o[j]);
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// the core: one CTA computes a [128 x n_tile] accumulator over `k_stages` stages of 32.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_pending(int n) {      // at most n groups still in flight
  switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
  }
}
__device__ __forceinline__ void producers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kProducerThreads) : "memory"); }

template <bool A_MN, bool B_MN, bool SPLIT3, bool ASYNC, class LoadA, class LoadB, class Epi>
__device__ __forceinline__ void gemm_core(const LoadA& load_a, const LoadB& load_b, const Epi& epi, int n_tile,
                                          int k_stages, int num_stages, unsigned char* smem, const void* gdummy,
                                          const CUtensorMap* tmap_b = nullptr, int tma_b_row = 0, int tma_b_box_rows = 0,
                                          int l2norm = 0, bool late_pdl = false) {
  // l2norm (forward only, n_tile == 128 == the whole output row): rows leave as relu(acc) / max(||relu(acc)||_2, 1e-12)
  // tmap_b != nullptr (K-major B only): the B tile of k-stage ks is the TMA box {32 k from 32*ks, tma_b_box_rows
  // rows from tma_b_row}; the producers only split it (hi/lo) once it has landed
  // ---- carve shared memory: [stages][A_hi, (A_lo), B_hi, (B_lo)], 1024-byte aligned ----
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;
  unsigned char* smem_al = smem + (smem_base - smem_u32(smem));
  const int a_bytes = tile_bytes(kTileM, A_MN);
  const int b_bytes = tile_bytes(n_tile, B_MN);
  const int stage_bytes = (SPLIT3 ? 2 : 1) * (a_bytes + b_bytes);
  __shared__ __align__(8) uint64_t s_full[kMaxStages];
  __shared__ __align__(8) uint64_t s_empty[kMaxStages];
  __shared__ __align__(8) uint64_t s_acc;
  __shared__ __align__(8) uint64_t s_tma[kMaxStages];      // B tile landed (TMA transaction barrier)
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t tmem_cols = n_tile <= 32 ? 32u : n_tile <= 64 ? 64u : n_tile <= 128 ? 128u : 256u;

  if (tid == 0) {
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(smem_u32(&s_full[s]), kProducerWarps);      // one arrival per producer warp (512 serialised arrivals cost ~1 us per stage)
      mbar_init(smem_u32(&s_empty[s]), 1);
      mbar_init(smem_u32(&s_tma[s]), 1);
    }
    mbar_init(smem_u32(&s_acc), 1);
    fence_barrier_init();
  }
  if (warp == kProducerWarps) tmem_alloc(smem_u32(&s_tmem), tmem_cols);
  tc_fence_before();
  GS_TRACE(0);
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = s_tmem;
  if (late_pdl) pdl_sync();                        // the caller has read nothing the previous kernels wrote so far
  GS_TRACE(1);

  if (warp < kProducerWarps) {
    const int a_chunks = chunks_in_tile(kTileM, A_MN);
    const int b_chunks = chunks_in_tile(n_tile, B_MN);
    constexpr int kMaxA = 1024 / kProducerThreads, kMaxB = 2048 / kProducerThreads;   // 16-byte pieces per thread
    if (ASYNC) {
      // ================= producers, asynchronous path: cp.async ring, num_stages - 1 stages ahead =================
      const int ahead = num_stages - 1;            // 0 only if a single stage fits (then load and compute alternate)
      typename LoadA::Prep pa[kMaxA];
      typename LoadB::Prep pb[kMaxB];
      int oa[kMaxA], ob[kMaxB];                    // byte offset of my pieces inside an operand tile (-1: none)
#pragma unroll
      for (int i = 0; i < kMaxA; ++i) {
        const int q = tid + i * kProducerThreads;
        int e = 0, k = 0;
        oa[i] = -1;
        if (q < a_chunks) chunk_coords(q, kTileM, A_MN, e, k, oa[i]);
        pa[i] = load_a.prep(e, k);
      }
#pragma unroll
      for (int i = 0; i < kMaxB; ++i) {
        const int q = tid + i * kProducerThreads;
        int e = 0, k = 0;
        ob[i] = -1;
        if (q < b_chunks) chunk_coords(q, n_tile, B_MN, e, k, ob[i]);
        pb[i] = load_b.prep(e, k);
      }
      int issued = 0;
      auto issue_next = [&]() {                    // always commits exactly one group (empty past the last stage)
        if (issued < k_stages) {
          const int stage = issued % num_stages;
          if (issued >= num_stages) mbar_wait(smem_u32(&s_empty[stage]), static_cast<uint32_t>(issued / num_stages - 1) & 1u);
          const uint32_t a_hi = smem_base + stage * stage_bytes;
          const uint32_t b_hi = a_hi + (SPLIT3 ? 2 : 1) * a_bytes;
#pragma unroll
          for (int i = 0; i < kMaxA; ++i) {
            if (oa[i] >= 0) {
              const float* src = load_a.ptr(pa[i], issued);
              cp_async16(a_hi + oa[i], src ? static_cast<const void*>(src) : gdummy, src ? 16u : 0u);
            }
          }
          if (tmap_b == nullptr) {
#pragma unroll
            for (int i = 0; i < kMaxB; ++i) {
              if (ob[i] >= 0) {
                const float* src = load_b.ptr(pb[i], issued);
                cp_async16(b_hi + ob[i], src ? static_cast<const void*>(src) : gdummy, src ? 16u : 0u);
              }
            }
          } else if (tid == 0) {
            const uint32_t bar = smem_u32(&s_tma[stage]);
            mbar_expect_tx(bar, static_cast<uint32_t>(tma_b_box_rows) * 128u);
            tma_load_2d(b_hi, tmap_b, issued * kBK, tma_b_row, bar);
          }
        }
        cp_async_commit_group();
        if (warp == 0 && issued < k_stages) GS_TRACE(19 + 4 * issued);
        ++issued;
      };
      for (int i = 0; i < ahead; ++i) issue_next();
      for (int ks = 0; ks < k_stages; ++ks) {
        const int stage = ks % num_stages;
        if (ahead == 0) issue_next();
        cp_async_wait_pending(ahead == 0 ? 0 : ahead - 1);       // my pieces of stage ks have landed
        if (tmap_b != nullptr) mbar_wait(smem_u32(&s_tma[stage]), static_cast<uint32_t>(ks / num_stages) & 1u);
        if (warp == 0) GS_TRACE(16 + 4 * ks);
        if (SPLIT3) {                                            // split MY pieces in place: hi = trunc_tf32(x), lo = x - hi
          unsigned char* a_hi = smem_al + stage * stage_bytes;
          unsigned char* a_lo = a_hi + a_bytes;
          unsigned char* b_hi = a_hi + 2 * a_bytes;
          unsigned char* b_lo = b_hi + b_bytes;
#pragma unroll
          for (int i = 0; i < kMaxA; ++i) {
            if (oa[i] >= 0) {
              const float4 v = *reinterpret_cast<const float4*>(a_hi + oa[i]);
              const float4 h = tf32_hi(v);
              if (!kImplicitTrunc) *reinterpret_cast<float4*>(a_hi + oa[i]) = h;
              *reinterpret_cast<float4*>(a_lo + oa[i]) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
            }
          }
#pragma unroll
          for (int i = 0; i < kMaxB; ++i) {
            if (ob[i] >= 0) {
              const float4 v = *reinterpret_cast<const float4*>(b_hi + ob[i]);
              const float4 h = tf32_hi(v);
              if (!kImplicitTrunc) *reinterpret_cast<float4*>(b_hi + ob[i]) = h;
              *reinterpret_cast<float4*>(b_lo + ob[i]) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
            }
          }
        }
        fence_proxy_async();                       // generic-proxy writes -> visible to the tensor-core (async) proxy
        __syncwarp();                              // every lane's writes and fence are ordered before lane 0's arrival
        if (lane == 0) mbar_arrive(smem_u32(&s_full[stage]));
        if (warp == 0) GS_TRACE(17 + 4 * ks);
        if (ahead > 0) issue_next();               // refill the stage the MMA of k-stage ks-1 is about to release
      }
      cp_async_wait_pending(0);
    } else {
    // ================= producers, register-staged path: global -> registers -> swizzled smem =================
    // (operands whose rows are not 16-byte aligned, or that need the ReLU mask applied on the fly)
    for (int ks = 0; ks < k_stages; ++ks) {
      const int stage = ks % num_stages;
      const uint32_t use = static_cast<uint32_t>(ks / num_stages);
      if (ks >= num_stages) mbar_wait(smem_u32(&s_empty[stage]), (use - 1) & 1u);
      unsigned char* a_hi = smem_al + stage * stage_bytes;
      unsigned char* a_lo = a_hi + a_bytes;
      unsigned char* b_hi = a_hi + (SPLIT3 ? 2 : 1) * a_bytes;
      unsigned char* b_lo = b_hi + b_bytes;
      float4 va[kMaxA], vb[kMaxB];
      int oa[kMaxA], ob[kMaxB];
#pragma unroll
      for (int i = 0; i < kMaxA; ++i) {
        const int q = tid + i * kProducerThreads;
        oa[i] = -1;
        if (q < a_chunks) {
          int e, k;
          chunk_coords(q, kTileM, A_MN, e, k, oa[i]);
          va[i] = load_a(e, k, ks);
        }
      }
#pragma unroll
      for (int i = 0; i < kMaxB; ++i) {
        const int q = tid + i * kProducerThreads;
        ob[i] = -1;
        if (q < b_chunks) {
          int e, k;
          chunk_coords(q, n_tile, B_MN, e, k, ob[i]);
          vb[i] = load_b(e, k, ks);
        }
      }
#pragma unroll
      for (int i = 0; i < kMaxA; ++i) {
        if (oa[i] < 0) continue;
        if (SPLIT3) {
          const float4 h = tf32_hi(va[i]);
          *reinterpret_cast<float4*>(a_hi + oa[i]) = h;
          *reinterpret_cast<float4*>(a_lo + oa[i]) = make_float4(va[i].x - h.x, va[i].y - h.y, va[i].z - h.z, va[i].w - h.w);
        } else {
          *reinterpret_cast<float4*>(a_hi + oa[i]) = va[i];
        }
      }
#pragma unroll
      for (int i = 0; i < kMaxB; ++i) {
        if (ob[i] < 0) continue;
        if (SPLIT3) {
          const float4 h = tf32_hi(vb[i]);
          *reinterpret_cast<float4*>(b_hi + ob[i]) = h;
          *reinterpret_cast<float4*>(b_lo + ob[i]) = make_float4(vb[i].x - h.x, vb[i].y - h.y, vb[i].z - h.z, vb[i].w - h.w);
        } else {
          *reinterpret_cast<float4*>(b_hi + ob[i]) = vb[i];
        }
      }
      fence_proxy_async();                         // generic-proxy stores -> visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_full[stage]));
    }
    }
    // =========================== epilogue: TMEM -> registers -> smem (transpose) -> global ===========================
    mbar_wait(smem_u32(&s_acc), 0);                // every MMA has completed: the operand ring is idle
    tc_fence_after();
    if (warp == 0) GS_TRACE(2);
    const int quad = warp & 3;                     // a warp may only touch TMEM lanes 32*(warp%4) .. +31
    const int m = quad * 32 + lane;
    const int n_chunks = (n_tile + 31) / 32;
    const int ldst = n_tile + 4;                   // floats per staged row (+4: rows land on different banks)
    float* stg = reinterpret_cast<float*>(smem_al);
    for (int c = warp >> 2; c < n_chunks; c += kProducerWarps / 4) {
      uint32_t v[32];
      tmem_ld32(tmem_acc + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(c * 32), v);
      float* dst = stg + m * ldst + c * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        if (c * 32 + j < n_tile)
          *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                            __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    }
    tc_fence_before();
    if (warp == 0) GS_TRACE(4);
    producers_sync();
    if (warp == 0) GS_TRACE(5);
    const int n4 = n_tile >> 2;
    const uint32_t stg_s = smem_base;               // explicit shared-space loads: LDS, and no aliasing with the global stores
    for (int q = lane; q < n4; q += 32) {
#pragma unroll 1
      for (int m0 = warp; m0 < kTileM; m0 += 4 * kProducerWarps) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t a = stg_s + static_cast<uint32_t>(((m0 + u * kProducerWarps) * ldst + 4 * q) * 4);
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "r"(a));
        }
        if (l2norm) {             // n4 == 32: the 32 lanes of the warp hold one whole row per u
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            v[u].x = fmaxf(v[u].x, 0.f); v[u].y = fmaxf(v[u].y, 0.f); v[u].z = fmaxf(v[u].z, 0.f); v[u].w = fmaxf(v[u].w, 0.f);
            const float ss = warp_sum(v[u].x * v[u].x + v[u].y * v[u].y + v[u].z * v[u].z + v[u].w * v[u].w);
            const float sc = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
            v[u].x *= sc; v[u].y *= sc; v[u].z *= sc; v[u].w *= sc;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) epi(m0 + u * kProducerWarps, 4 * q, v[u]);
      }
    }
    if (warp == 0) GS_TRACE(3);
  } else if (warp == kProducerWarps) {
    // =========================== MMA issuer (one elected thread) ===========================
    const uint32_t idesc = make_idesc(kTileM, n_tile, A_MN, B_MN);
    const int a_groups = (kTileM + 31) / 32, b_groups = (n_tile + 31) / 32;
    for (int ks = 0; ks < k_stages; ++ks) {
      const int stage = ks % num_stages;
      mbar_wait(smem_u32(&s_full[stage]), static_cast<uint32_t>(ks / num_stages) & 1u);
      tc_fence_after();
      GS_TRACE(18 + 4 * ks);
      if (lane == 0) {
        const uint32_t a_hi = smem_base + stage * stage_bytes;
        const uint32_t a_lo = a_hi + a_bytes;
        const uint32_t b_hi = a_hi + (SPLIT3 ? 2 : 1) * a_bytes;
        const uint32_t b_lo = b_hi + b_bytes;
#pragma unroll
        for (int kk = 0; kk < kBK / 8; ++kk) {     // UMMA_K = 8 tf32 = 32 bytes
          const uint32_t a_off = A_MN ? kk * a_groups * 1024 : kk * 32;
          const uint32_t b_off = B_MN ? kk * b_groups * 1024 : kk * 32;
          const uint32_t a_lbo = A_MN ? 512 : 16, a_sbo = A_MN ? a_groups * 512 : 1024, a_ly = A_MN ? 1 : 2;
          const uint32_t b_lbo = B_MN ? 512 : 16, b_sbo = B_MN ? b_groups * 512 : 1024, b_ly = B_MN ? 1 : 2;
          const uint32_t first = (ks == 0 && kk == 0) ? 0u : 1u;
          if (SPLIT3) {
            umma_tf32(tmem_acc, make_desc(a_lo + a_off, a_lbo, a_sbo, a_ly), make_desc(b_hi + b_off, b_lbo, b_sbo, b_ly), idesc, first);
            umma_tf32(tmem_acc, make_desc(a_hi + a_off, a_lbo, a_sbo, a_ly), make_desc(b_lo + b_off, b_lbo, b_sbo, b_ly), idesc, 1u);
            umma_tf32(tmem_acc, make_desc(a_hi + a_off, a_lbo, a_sbo, a_ly), make_desc(b_hi + b_off, b_lbo, b_sbo, b_ly), idesc, 1u);
          } else {
            umma_tf32(tmem_acc, make_desc(a_hi + a_off, a_lbo, a_sbo, a_ly), make_desc(b_hi + b_off, b_lbo, b_sbo, b_ly), idesc, first);
          }
        }
        umma_commit(smem_u32(&s_empty[stage]));    // frees the stage when these MMAs have read it
        if (ks == k_stages - 1) umma_commit(smem_u32(&s_acc));
      }
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == kProducerWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, tmem_cols);
  }
}

template <bool SPLIT3, bool ASYNC>
__global__ void __maxnreg__(kMaxRegs)
sage_fwd_tc_kernel(XView x, const float* __restrict__ weight, int64_t ldw, int out_dim, bool vec_ok,
                   const int32_t* __restrict__ num_rows_dev, int max_rows, float* __restrict__ out, int64_t ld_out,
                   int relu, int n_tile, int k_stages, int num_stages, const __grid_constant__ CUtensorMap tmap_w,
                   int use_tma_w, float* __restrict__ zero_out, int64_t ld_zero, int l2norm) {
  pdl_sync();
  extern __shared__ unsigned char smem_dyn[];
  const int rows = live_rows(num_rows_dev, max_rows);
  const int row0 = blockIdx.x * kTileM, h0 = blockIdx.y * n_tile;
  if (row0 >= rows) return;
  const int nt = min(n_tile, ((out_dim - h0) + 15) & ~15);
  __shared__ int32_t s_rowidx[kTileM];
  x.fill_cache(s_rowidx, row0, kTileM, rows);            // made visible by the __syncthreads in gemm_core
  LoadX_K la{x, row0, rows};
  LoadW_K lb{x, weight, ldw, h0, out_dim, vec_ok};
  StoreOut epi{out, ld_out, row0, rows, h0, out_dim, relu, zero_out, ld_zero};
  gemm_core<false, false, SPLIT3, ASYNC>(la, lb, epi, nt, k_stages, num_stages, smem_dyn, weight,
                                         (ASYNC && use_tma_w) ? &tmap_w : nullptr, h0, n_tile, l2norm);
}

template <bool SPLIT3, bool ASYNC>
__global__ void __maxnreg__(kMaxRegs)
sage_bwd_x_tc_kernel(const float* __restrict__ grad_out, int64_t ld_go, const float* __restrict__ out, int64_t ld_out,
                     const float* __restrict__ weight, int64_t ldw, int dim, int out_dim, int gcn, int relu, bool vec_ok,
                     const int32_t* __restrict__ num_rows_dev, int max_rows, float* __restrict__ grad_self, int64_t ld_gs,
                     float* __restrict__ grad_agg, int64_t ld_ga, int n_tile, int k_stages, int num_stages) {
  pdl_sync();
  extern __shared__ unsigned char smem_dyn[];
  const int rows = live_rows(num_rows_dev, max_rows);
  const int row0 = blockIdx.x * kTileM, c0 = blockIdx.y * n_tile;
  if (row0 >= rows) return;
  const int ncols = gcn ? dim : 2 * dim;
  const int nt = min(n_tile, ((ncols - c0) + 15) & ~15);
  LoadDZ_K la{grad_out, ld_go, out, ld_out, row0, rows, out_dim, relu};
  LoadW_MN lb{weight, ldw, c0, ncols, out_dim, vec_ok};
  StoreDX epi{grad_self, ld_gs, grad_agg, ld_ga, row0, rows, c0, dim, ncols, gcn};
  gemm_core<false, true, SPLIT3, ASYNC>(la, lb, epi, nt, k_stages, num_stages, smem_dyn, weight);
}

// One weight-gradient problem dW[h,kv] += sum_r dZ[r,h] X[r,kv], cut into `chunks` row chunks.
// out_dim bounds the columns of dZ that are READ (a multiple of 4 on the cp.async path; a caller whose dZ rows are
// zero-padded passes the padded width), out_rows <= out_dim the rows of dW that exist.
struct BwdWProblem {
  XView x; const float* grad_out; int64_t ld_go; const float* out; int64_t ld_out; int out_dim, out_rows, relu;
  const int32_t* num_rows_dev; int max_rows, rows_per_chunk; float* grad_w; int64_t ldw; int n_tile, num_stages;
  int tiles_x, tiles_y, chunks;
  int early;                 // gs_set_early_reads at launch time: row count and index list are read before the PDL wait
};

// The grid's z axis runs over the row chunks of problem A, then those of problem B, then those of problem C (chunks == 0:
// problem absent).  The weight gradients of a step -- the layers' and the classifier's -- are independent leaves of its
// dependency graph; launched as one grid they share the machine (CTAs split in proportion to their work) instead of
// queueing behind each other.
template <bool SPLIT3, bool ASYNC>
__global__ void __maxnreg__(kMaxRegs)
sage_bwd_w_tc_kernel(const __grid_constant__ BwdWProblem pa, const __grid_constant__ BwdWProblem pb,
                     const __grid_constant__ BwdWProblem pc) {
  if (!pa.early) pdl_sync();   // else: after the setup inside gemm_core (row count and index cache are read early)
  extern __shared__ unsigned char smem_dyn[];
  const int z = blockIdx.z;
  const int which = z < pa.chunks ? 0 : (z < pa.chunks + pb.chunks ? 1 : 2);
  const BwdWProblem& q = *(which == 0 ? &pa : (which == 1 ? &pb : &pc));     // stays in the constant bank
  const int chunk = which == 0 ? z : (which == 1 ? z - pa.chunks : z - pa.chunks - pb.chunks);
  if (static_cast<int>(blockIdx.x) >= q.tiles_x || static_cast<int>(blockIdx.y) >= q.tiles_y) return;
  XView x = q.x;
  const int rows = live_rows(q.num_rows_dev, q.max_rows);
  const int kv0 = blockIdx.x * q.n_tile, h0 = blockIdx.y * kTileM;
  const int r_begin = chunk * q.rows_per_chunk;
  const int r_end = min(rows, r_begin + q.rows_per_chunk);
  if (r_begin >= r_end) return;
  const int nt = min(q.n_tile, ((x.kv_total() - kv0) + 15) & ~15);
  const int k_stages = (r_end - r_begin + kBK - 1) / kBK;
  __shared__ int32_t s_rowidx[kMaxChunkRows];
  x.fill_cache(s_rowidx, r_begin, min(q.rows_per_chunk, kMaxChunkRows), rows);
  LoadDZ_MN la{q.grad_out, q.ld_go, q.out, q.ld_out, r_begin, r_end, h0, q.out_dim, q.relu};
  LoadX_MN lb{x, r_begin, r_end, kv0};
  AddDW epi{x, q.grad_w, q.ldw, h0, q.out_rows, kv0};
  gemm_core<true, true, SPLIT3, ASYNC>(la, lb, epi, nt, k_stages, q.num_stages, smem_dyn, q.grad_out, nullptr, 0, 0, 0,
                                       pa.early != 0);
}

struct Plan { int n_tile, num_stages, smem; };

// n_total output columns, `m_tiles` CTAs along M.  n_tile is capped so that at least 3 stages of the
// (split) operand ring fit, and halved while the grid would leave most SMs idle (the extra CTAs
// re-read the A tile from L2, which is cheap next to an idle SM).
static Plan make_plan(int n_total, bool a_mn, bool b_mn, bool split3, int m_tiles) {
  Plan p{};
  int n_tile = (n_total + 15) & ~15;
  const int cap = split3 ? 128 : 256;
  if (n_tile > cap) n_tile = cap;
  while (n_tile > 32 && m_tiles * ((n_total + n_tile - 1) / n_tile) < kNumSMs / 2) n_tile = ((n_tile / 2) + 15) & ~15;
  p.n_tile = n_tile;
  const int stage = (split3 ? 2 : 1) * (tile_bytes(kTileM, a_mn) + tile_bytes(n_tile, b_mn));
  int stages = kSmemBudget / stage;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 1) stages = 1;
  p.num_stages = stages;
  const int staging = kTileM * (n_tile + 4) * 4;              // epilogue transpose buffer reuses the ring
  p.smem = (stages * stage > staging ? stages * stage : staging) + 1024;
  return p;
}

// 2-D fp32 tensor map over a row-major [rows x cols] matrix (leading dimension ld floats): box = 32 columns
// (one 128-byte swizzle row) x box_rows rows, SWIZZLE_128B, out-of-range elements read as zero.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    const char* e = getenv("GS_TC_TMA");
    if (e && e[0] == '0') fn = nullptr;                      // GS_TC_TMA=0: cp.async for every operand (A/B measurements)
  }
  return fn;
}
bool make_tmap_2d(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn || box_rows < 1 || box_rows > 256 || (ld & 3) || !aligned16(base)) return false;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 4u};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1u, 1u};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <class K>
static int set_smem(K kernel, int bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  return e == cudaSuccess ? GS_OK : static_cast<int>(e);
}

}  // namespace tc
}  // namespace gs

using namespace gs;
using namespace gs::tc;

// launch helper: picks the <SPLIT3, ASYNC> instantiation
#define GS_TC_LAUNCH(KERNEL, GRID, SMEM, STREAM, ...)                                                       \
  do {                                                                                                      \
    int e_ = 0;                                                                                             \
    if (split3 && async) { if ((e_ = set_smem(KERNEL<true, true>, SMEM))) return e_;                        \
      launch(KERNEL<true, true>, GRID, kThreads, SMEM, STREAM, __VA_ARGS__); }                                  \
    else if (split3) { if ((e_ = set_smem(KERNEL<true, false>, SMEM))) return e_;                           \
      launch(KERNEL<true, false>, GRID, kThreads, SMEM, STREAM, __VA_ARGS__); }                                 \
    else if (async) { if ((e_ = set_smem(KERNEL<false, true>, SMEM))) return e_;                            \
      launch(KERNEL<false, true>, GRID, kThreads, SMEM, STREAM, __VA_ARGS__); }                                 \
    else { if ((e_ = set_smem(KERNEL<false, false>, SMEM))) return e_;                                      \
      launch(KERNEL<false, false>, GRID, kThreads, SMEM, STREAM, __VA_ARGS__); }                                \
  } while (0)

int gs_sage_gemm_fwd_tc(const float* self_table, int64_t ld_self, const int32_t* self_idx, const float* agg,
                        int64_t ld_agg, int32_t dim, const float* weight, int64_t ldw, int32_t out_dim, int32_t gcn,
                        const int32_t* num_rows_dev, int32_t max_rows, float* out, int64_t ld_out, int32_t relu,
                        int32_t precision, float* zero_out, int64_t ld_zero, int32_t l2norm, gs_stream_t stream) {
  const bool split3 = precision == GS_PREC_TF32X3;
  if (precision != GS_PREC_TF32 && !split3) return GS_ERR_BAD_ARG;
  XView x{self_table, ld_self, self_idx, agg, ld_agg, dim, (dim + 3) & ~3, gcn, nullptr, 0, 0};
  const int kt = gcn ? x.dim_pad : 2 * x.dim_pad;
  Plan p = make_plan(out_dim, false, false, split3, (max_rows + kTileM - 1) / kTileM);
  if (l2norm) {                                    // the normalising epilogue needs the whole row in one CTA
    if (out_dim != 128 || !relu) return GS_ERR_UNSUPPORTED;
    p = make_plan(out_dim, false, false, split3, kNumSMs);
    if (p.n_tile != 128) return GS_ERR_UNSUPPORTED;
  }
  const int k_stages = (kt + kBK - 1) / kBK;
  const bool vec_ok = (dim % 4 == 0) && (ldw % 4 == 0) && aligned16(weight);
  const bool async = vec_ok;                       // X rows are always 16-byte aligned (padded tables)
  dim3 grid((max_rows + kTileM - 1) / kTileM, (out_dim + p.n_tile - 1) / p.n_tile);
  // W through TMA when every column tile is a full box (the box may not spill over the next operand tile)
  CUtensorMap tmap_w;
  memset(&tmap_w, 0, sizeof(tmap_w));
  const int use_tma_w = (async && out_dim % p.n_tile == 0 && make_tmap_2d(&tmap_w, weight, out_dim, kt, ldw, p.n_tile)) ? 1 : 0;
  GS_TC_LAUNCH(sage_fwd_tc_kernel, grid, p.smem, as_stream(stream), x, weight, ldw, out_dim, vec_ok, num_rows_dev,
               max_rows, out, ld_out, relu, p.n_tile, k_stages, p.num_stages, tmap_w, use_tma_w, zero_out, ld_zero, l2norm);
  return finish_launch();
}

int gs_sage_gemm_bwd_x_tc(const float* grad_out, int64_t ld_go, const float* out, int64_t ld_out, const float* weight,
                          int64_t ldw, int32_t dim, int32_t out_dim, int32_t gcn, int32_t relu,
                          const int32_t* num_rows_dev, int32_t max_rows, float* grad_self, int64_t ld_gs,
                          float* grad_agg, int64_t ld_ga, int32_t precision, gs_stream_t stream) {
  const bool split3 = precision == GS_PREC_TF32X3;
  if (precision != GS_PREC_TF32 && !split3) return GS_ERR_BAD_ARG;
  const int ncols = gcn ? dim : 2 * dim;
  const Plan p = make_plan(ncols, false, true, split3, (max_rows + kTileM - 1) / kTileM);
  const int k_stages = (out_dim + kBK - 1) / kBK;
  const bool vec_ok = (ldw % 4 == 0) && aligned16(weight);
  const bool async = vec_ok && !relu && (ncols % 4 == 0) && (out_dim % 4 == 0) && (ld_go % 4 == 0) && aligned16(grad_out);
  dim3 grid((max_rows + kTileM - 1) / kTileM, (ncols + p.n_tile - 1) / p.n_tile);
  GS_TC_LAUNCH(sage_bwd_x_tc_kernel, grid, p.smem, as_stream(stream), grad_out, ld_go, out, ld_out, weight, ldw, dim,
               out_dim, gcn, relu, vec_ok, num_rows_dev, max_rows, grad_self, ld_gs, grad_agg, ld_ga, p.n_tile, k_stages,
               p.num_stages);
  return finish_launch();
}

// Row chunks of one problem for a budget of `cta_budget` CTAs (about one per SM): at least 4 k-stages per CTA, at
// most kMaxChunkRows rows (the index cache); fixed_rows > 0 prescribes the chunk length instead.  Returns the dynamic
// shared memory its CTAs need.
static int plan_bwd_w(BwdWProblem& q, int kt, int out_dim, int max_rows, bool split3, int cta_budget, int fixed_rows = 0) {
  const Plan p = make_plan(kt, true, true, split3, kNumSMs);   // row chunks fill the machine
  q.n_tile = p.n_tile;
  q.num_stages = p.num_stages;
  q.tiles_x = (kt + p.n_tile - 1) / p.n_tile;
  q.tiles_y = (out_dim + kTileM - 1) / kTileM;
  const int tiles = q.tiles_x * q.tiles_y;
  int rows_per_chunk = fixed_rows;
  if (rows_per_chunk <= 0) {
    int chunks = (cta_budget + tiles - 1) / tiles;
    const int max_chunks = (max_rows + 4 * kBK - 1) / (4 * kBK);
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    rows_per_chunk = (max_rows + chunks - 1) / chunks;
    rows_per_chunk = ((rows_per_chunk + kBK - 1) / kBK) * kBK;
  }
  if (rows_per_chunk > kMaxChunkRows) rows_per_chunk = kMaxChunkRows;
  q.rows_per_chunk = rows_per_chunk;
  q.chunks = (max_rows + rows_per_chunk - 1) / rows_per_chunk;
  return p.smem;
}

static bool bwd_w_async_ok(const float* grad_out, int64_t ld_go, int out_dim, int relu) {
  return !relu && (out_dim % 4 == 0) && (ld_go % 4 == 0) && aligned16(grad_out);
}

static int launch_bwd_w(const BwdWProblem* pr, int n, int smem, bool split3, bool async, gs_stream_t stream) {
  BwdWProblem q[3] = {};
  int tx = 1, ty = 1, chunks = 0;
  for (int i = 0; i < n; ++i) {
    q[i] = pr[i];
    tx = q[i].tiles_x > tx ? q[i].tiles_x : tx;
    ty = q[i].tiles_y > ty ? q[i].tiles_y : ty;
    chunks += q[i].chunks;
  }
  dim3 grid(tx, ty, chunks);
  GS_TC_LAUNCH(sage_bwd_w_tc_kernel, grid, smem, as_stream(stream), q[0], q[1], q[2]);
  return finish_launch();
}

int gs_sage_gemm_bwd_w_tc(const float* self_table, int64_t ld_self, const int32_t* self_idx, const float* agg,
                          int64_t ld_agg, int32_t dim, const float* grad_out, int64_t ld_go, const float* out,
                          int64_t ld_out, int32_t out_dim, int32_t gcn, int32_t relu, const int32_t* num_rows_dev,
                          int32_t max_rows, float* grad_w, int64_t ldw, int32_t precision, gs_stream_t stream) {
  const bool split3 = precision == GS_PREC_TF32X3;
  if (precision != GS_PREC_TF32 && !split3) return GS_ERR_BAD_ARG;
  BwdWProblem pa{};
  pa.x = XView{self_table, ld_self, self_idx, agg, ld_agg, dim, (dim + 3) & ~3, gcn, nullptr, 0, 0};
  pa.grad_out = grad_out; pa.ld_go = ld_go; pa.out = out; pa.ld_out = ld_out; pa.out_dim = out_dim; pa.out_rows = out_dim;
  pa.relu = relu; pa.early = early_reads() ? 1 : 0;
  pa.num_rows_dev = num_rows_dev; pa.max_rows = max_rows; pa.grad_w = grad_w; pa.ldw = ldw;
  const int kt = gcn ? pa.x.dim_pad : 2 * pa.x.dim_pad;
  const int smem = plan_bwd_w(pa, kt, out_dim, max_rows, split3, kNumSMs);
  return launch_bwd_w(&pa, 1, smem, split3, bwd_w_async_ok(grad_out, ld_go, out_dim, relu), stream);
}

// Up to three problems in one launch (see sage_bwd_w_tc_kernel).  gcn / relu are per problem; go_cols[i] (nullable
// array; 0 = out_dim[i]) is the zero-padded width of problem i's dZ rows.  Returns GS_ERR_UNSUPPORTED when the problems
// cannot share a kernel instantiation (the caller then launches them one after the other).
int gs_sage_gemm_bwd_w_group_tc(int32_t n, const float* const* self_table, const int64_t* ld_self,
                                const int32_t* const* self_idx, const float* const* agg, const int64_t* ld_agg,
                                const int32_t* dim, const float* const* grad_out, const int64_t* ld_go,
                                const float* const* out, const int64_t* ld_out, const int32_t* out_dim,
                                const int32_t* go_cols, const int32_t* gcn, const int32_t* relu,
                                const int32_t* const* num_rows_dev, const int32_t* max_rows, float* const* grad_w,
                                const int64_t* ldw, int32_t precision, gs_stream_t stream) {
  const bool split3 = precision == GS_PREC_TF32X3;
  if (precision != GS_PREC_TF32 && !split3) return GS_ERR_BAD_ARG;
  if (n < 1 || n > 3) return GS_ERR_BAD_ARG;
  BwdWProblem pr[3] = {};
  int kt[3];
  bool async0 = false;
  for (int i = 0; i < n; ++i) {
    pr[i].x = XView{self_table[i], ld_self[i], self_idx[i], agg[i], ld_agg[i], dim[i], (dim[i] + 3) & ~3, gcn[i], nullptr, 0, 0};
    const int cols = (go_cols && go_cols[i] > 0) ? go_cols[i] : out_dim[i];
    if (cols < out_dim[i] || cols > ld_go[i]) return GS_ERR_BAD_ARG;
    pr[i].grad_out = grad_out[i]; pr[i].ld_go = ld_go[i]; pr[i].out = out[i]; pr[i].ld_out = ld_out[i];
    pr[i].out_dim = cols; pr[i].out_rows = out_dim[i]; pr[i].relu = relu[i];
    pr[i].num_rows_dev = num_rows_dev[i]; pr[i].max_rows = max_rows[i];
    pr[i].grad_w = grad_w[i]; pr[i].ldw = ldw[i]; pr[i].early = early_reads() ? 1 : 0;
    kt[i] = gcn[i] ? pr[i].x.dim_pad : 2 * pr[i].x.dim_pad;
    const bool a = bwd_w_async_ok(grad_out[i], ld_go[i], cols, relu[i]);
    if (i == 0) async0 = a;
    else if (a != async0) return GS_ERR_UNSUPPORTED;
  }
  // CTAs of EQUAL length: every problem is cut into chunks of the same R rows, R the smallest multiple of the k-stage
  // (>= 4 stages) with which all the CTAs fit one wave -- the grid ends when its longest CTA does, and a small problem
  // given CTAs in proportion to its flops would have had the longest ones
  int smem = 0, rows_fixed = kMaxChunkRows;
  for (int r = 4 * kBK; r <= kMaxChunkRows; r += kBK) {
    int ctas = 0;
    for (int i = 0; i < n; ++i) {
      const Plan p = make_plan(kt[i], true, true, split3, kNumSMs);
      const int tiles = ((kt[i] + p.n_tile - 1) / p.n_tile) * ((pr[i].out_dim + kTileM - 1) / kTileM);
      ctas += tiles * ((max_rows[i] + r - 1) / r);
    }
    if (ctas <= kNumSMs) { rows_fixed = r; break; }
  }
  for (int i = 0; i < n; ++i) {
    const int sm = plan_bwd_w(pr[i], kt[i], pr[i].out_dim, max_rows[i], split3, 0, rows_fixed);
    smem = sm > smem ? sm : smem;
  }
  return launch_bwd_w(pr, n, smem, split3, async0, stream);
}

#ifdef GS_TC_TRACE
extern "C" int gs_debug_trace_read(long long* host_out, int n) {
  if (n > 512) n = 512;
  cudaDeviceSynchronize();
  cudaError_t e = cudaMemcpyFromSymbol(host_out, gs::tc::g_trace, sizeof(long long) * n);
  return static_cast<int>(e);
}
#endif
