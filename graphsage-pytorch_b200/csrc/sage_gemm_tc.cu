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
// (ReLU mask for dZ, hi/lo split for the fp32-faithful mode), so 8 producer warps load 16-byte
// pieces (coalesced per row), transform them in registers and store them into the 128-byte
// swizzled layout the UMMA shared-memory descriptors expect; a proxy fence + mbarrier hands the
// stage to the single MMA-issuing thread; tcgen05.commit hands it back.  The accumulator
// (128 lanes x N fp32 columns) lives in TMEM and is read once by the epilogue (tcgen05.ld).
//
// Precision modes:
//   GS_PREC_TF32   one tf32 product  (10-bit mantissa operands, fp32 accumulate)
//   GS_PREC_TF32X3 x = hi + lo with hi = x truncated to tf32, lo = x - hi (exact in fp32);
//                  D = Alo.Bhi + Ahi.Blo + Ahi.Bhi.  Dropped term and the truncation of lo are
//                  O(2^-22) relative: fp32-faithful, meets the 1e-5 parity bound.
#include "common.cuh"

namespace gs {
namespace tc {

constexpr int kTileM = 128;
constexpr int kBK = 32;                                  // tf32 elements per k-stage: 128 bytes, one swizzle row
constexpr int kProducerWarps = 8;
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kLoaderWarps = 8;                          // cp.async issuers of the asynchronous path (4 for A, 4 for B)
constexpr int kThreads = kProducerThreads + 32 + kLoaderWarps * 32;   // converters/epilogue + MMA/TMEM warp + loaders
constexpr int kMaxStages = 4;
constexpr int kSmemBudget = 196 * 1024;
constexpr int kMaxChunkRows = 512;                       // bwd_w: rows reduced per CTA (index cache size)

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
// the mbarrier receives one arrival from this thread once all its earlier cp.async have landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
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

// One operand row as the loader warps see it: elements [0, split) come from p0, [split, limit)
// from p1 (stored pre-offset by -split so that p1 + x is the address), anything else is zero.
struct RowDesc {
  const float* p0; const float* p1; int split, limit;
  __device__ __forceinline__ const float* at(int x) const {
    return x < split ? p0 + x : (x < limit ? p1 + x : nullptr);
  }
};
__device__ __forceinline__ RowDesc empty_row() { return RowDesc{nullptr, nullptr, 0, 0}; }
__device__ __forceinline__ RowDesc xview_row(const XView& x, int r) {
  if (x.gcn) return RowDesc{x.agg + static_cast<int64_t>(r) * x.ld_agg, nullptr, x.dim_pad, x.dim_pad};
  return RowDesc{x.self_table + static_cast<int64_t>(x.self_row(r)) * x.ld_self,
                 x.agg + static_cast<int64_t>(r) * x.ld_agg - x.dim_pad, x.dim_pad, 2 * x.dim_pad};
}

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
  __device__ __forceinline__ RowDesc row(int e, int) const { return row0 + e < rows ? xview_row(x, row0 + e) : empty_row(); }
  __device__ __forceinline__ int origin() const { return 0; }
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
  __device__ __forceinline__ RowDesc row(int e, int) const {
    const int kt = x.kv_total();
    return h0 + e < out_dim ? RowDesc{w + static_cast<int64_t>(h0 + e) * ldw, nullptr, kt, kt} : empty_row();
  }
  __device__ __forceinline__ int origin() const { return 0; }
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
  __device__ __forceinline__ RowDesc row(int e, int) const {
    return row0 + e < rows ? RowDesc{go + static_cast<int64_t>(row0 + e) * ld_go, nullptr, out_dim, out_dim} : empty_row();
  }
  __device__ __forceinline__ int origin() const { return 0; }
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
  __device__ __forceinline__ RowDesc row(int k, int ks) const {
    const int h = ks * kBK + k;
    return h < out_dim ? RowDesc{w + static_cast<int64_t>(h) * ldw, nullptr, ncols, ncols} : empty_row();
  }
  __device__ __forceinline__ int origin() const { return c0; }
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
  __device__ __forceinline__ RowDesc row(int k, int ks) const {
    const int r = r_begin + ks * kBK + k;
    return r < r_end ? RowDesc{go + static_cast<int64_t>(r) * ld_go, nullptr, out_dim, out_dim} : empty_row();
  }
  __device__ __forceinline__ int origin() const { return h0; }
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
  __device__ __forceinline__ RowDesc row(int k, int ks) const {
    const int r = r_begin + ks * kBK + k;
    return r < r_end ? xview_row(x, r) : empty_row();
  }
  __device__ __forceinline__ int origin() const { return kv0; }
};

// ---------------------------------------------------------------------------------------------
// epilogues: called per thread with row m (tile-local, = TMEM lane) and 32 accumulator columns
// ---------------------------------------------------------------------------------------------
struct StoreOut {          // forward: out[r, h] = relu(acc)
  float* out; int64_t ld_out; int row0, rows, h0, out_dim, relu;
  __device__ __forceinline__ void operator()(int m, int n_base, const uint32_t (&v)[32]) const {
    const int r = row0 + m;
    if (r >= rows) return;
    float* dst = out + static_cast<int64_t>(r) * ld_out + h0 + n_base;
    const int lim = out_dim - h0 - n_base;
    if (lim >= 32 && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        *reinterpret_cast<float4*>(dst + j) = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < lim) { const float o = __uint_as_float(v[j]); dst[j] = relu ? fmaxf(o, 0.f) : o; }
    }
  }
};
struct StoreDX {           // bwd_x: dX[r, c] -> grad_self (c < dim) / grad_agg (c >= dim)
  float* gs; int64_t ld_gs; float* ga; int64_t ld_ga; int row0, rows, c0, dim, ncols, gcn;
  __device__ __forceinline__ void operator()(int m, int n_base, const uint32_t (&v)[32]) const {
    const int r = row0 + m;
    if (r >= rows) return;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int c = c0 + n_base + j;
      if (c >= ncols) continue;
      const float o = __uint_as_float(v[j]);
      if (gcn) ga[static_cast<int64_t>(r) * ld_ga + c] = o;
      else if (c < dim) gs[static_cast<int64_t>(r) * ld_gs + c] = o;
      else ga[static_cast<int64_t>(r) * ld_ga + (c - dim)] = o;
    }
  }
};
struct AddDW {             // bwd_w: grad_w[h, wcol(kv)] += acc   (row chunks reduced with fp32 atomics)
  XView x; float* gw; int64_t ldw; int h0, out_dim, kv0;
  __device__ __forceinline__ void operator()(int m, int n_base, const uint32_t (&v)[32]) const {
    const int h = h0 + m;
    if (h >= out_dim) return;
    const int kt = x.kv_total();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int kv = kv0 + n_base + j;
      const int c = kv < kt ? x.wcol(kv) : -1;
      if (c >= 0) atomicAdd(gw + static_cast<int64_t>(h) * ldw + c, __uint_as_float(v[j]));
    }
  }
};

// ---------------------------------------------------------------------------------------------
// the core: one CTA computes a [128 x n_tile] accumulator over `k_stages` stages of 32.
// ---------------------------------------------------------------------------------------------
template <bool A_MN, bool B_MN, bool SPLIT3, bool ASYNC, class LoadA, class LoadB, class Epi>
__device__ __forceinline__ void gemm_core(const LoadA& load_a, const LoadB& load_b, const Epi& epi, int n_tile,
                                          int k_stages, int num_stages, unsigned char* smem, const void* gdummy) {
  // ---- carve shared memory: [stages][A_hi, (A_lo), B_hi, (B_lo)], 1024-byte aligned ----
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;
  unsigned char* smem_al = smem + (smem_base - smem_u32(smem));
  const int a_bytes = tile_bytes(kTileM, A_MN);
  const int b_bytes = tile_bytes(n_tile, B_MN);
  const int stage_bytes = (SPLIT3 ? 2 : 1) * (a_bytes + b_bytes);
  __shared__ __align__(8) uint64_t s_full[kMaxStages];
  __shared__ __align__(8) uint64_t s_empty[kMaxStages];
  __shared__ __align__(8) uint64_t s_landed[kMaxStages];
  __shared__ __align__(8) uint64_t s_acc;
  __shared__ uint32_t s_tmem;
  __shared__ RowDesc s_rowdesc[kLoaderWarps][32];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t tmem_cols = n_tile <= 32 ? 32u : n_tile <= 64 ? 64u : n_tile <= 128 ? 128u : 256u;

  if (tid == 0) {
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(smem_u32(&s_full[s]), kProducerThreads);
      mbar_init(smem_u32(&s_empty[s]), 1);
      mbar_init(smem_u32(&s_landed[s]), kLoaderWarps * 32);
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
  GS_TRACE(1);

  if (warp < kProducerWarps) {
    const int a_chunks = chunks_in_tile(kTileM, A_MN);
    const int b_chunks = chunks_in_tile(n_tile, B_MN);
    constexpr int kMaxA = 4, kMaxB = 8;           // 1024 / 256 and 2048 / 256 pieces per thread
    if (ASYNC) {
      // ================= converters, asynchronous path =================
      // The two loader warps (below) cp.async every 16-byte piece straight into its swizzled
      // slot of the *hi* buffer, as many stages ahead as the ring allows and without holding
      // registers, so the gather latency of several stages overlaps.  When a stage has landed
      // these 8 warps rewrite hi = trunc_tf32(x) and derive lo = x - hi in shared memory
      // (3xTF32 only), then hand the stage to the tensor core.
      for (int ks = 0; ks < k_stages; ++ks) {
        const int stage = ks % num_stages;
        mbar_wait(smem_u32(&s_landed[stage]), static_cast<uint32_t>(ks / num_stages) & 1u);
        if (warp == 0) GS_TRACE(16 + 4 * ks);
        if (SPLIT3) {
          unsigned char* a_hi = smem_al + stage * stage_bytes;
          unsigned char* a_lo = a_hi + a_bytes;
          unsigned char* b_hi = a_hi + 2 * a_bytes;
          unsigned char* b_lo = b_hi + b_bytes;
#pragma unroll
          for (int i = 0; i < kMaxA; ++i) {
            const int q = tid + i * kProducerThreads;
            if (q < a_chunks) {
              int e, k, off;
              chunk_coords(q, kTileM, A_MN, e, k, off);
              const float4 v = *reinterpret_cast<const float4*>(a_hi + off);
              const float4 h = tf32_hi(v);
              *reinterpret_cast<float4*>(a_hi + off) = h;
              *reinterpret_cast<float4*>(a_lo + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
            }
          }
#pragma unroll
          for (int i = 0; i < kMaxB; ++i) {
            const int q = tid + i * kProducerThreads;
            if (q < b_chunks) {
              int e, k, off;
              chunk_coords(q, n_tile, B_MN, e, k, off);
              const float4 v = *reinterpret_cast<const float4*>(b_hi + off);
              const float4 h = tf32_hi(v);
              *reinterpret_cast<float4*>(b_hi + off) = h;
              *reinterpret_cast<float4*>(b_lo + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(smem_u32(&s_full[stage]));
        if (warp == 0) GS_TRACE(17 + 4 * ks);
      }
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
      mbar_arrive(smem_u32(&s_full[stage]));
    }
    }
    // =========================== epilogue: TMEM -> registers -> global ===========================
    mbar_wait(smem_u32(&s_acc), 0);
    tc_fence_after();
    if (warp == 0) GS_TRACE(2);
    const int quad = warp & 3;                     // a warp may only touch TMEM lanes 32*(warp%4) .. +31
    const int m = quad * 32 + lane;
    const int n_chunks = (n_tile + 31) / 32;
    for (int c = warp >> 2; c < n_chunks; c += 2) {
      uint32_t v[32];
      tmem_ld32(tmem_acc + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(c * 32), v);
      epi(m, c * 32, v);
    }
    if (warp == 0) GS_TRACE(3);
    tc_fence_before();
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
  else if (ASYNC) {
    // =========================== loader warps: 4 stream A, 4 stream B ===========================
    // A lone warp retires roughly one dependent instruction per 4-6 cycles, so the per-piece
    // address arithmetic is what bounds a cp.async stream (measured: ~400 cycles per piece with
    // generic per-piece addressing).  Each warp therefore first builds, lane-per-row, a small
    // table of row descriptors in shared memory (base pointers, the self|agg seam, the valid
    // length) and then walks the pieces lane-per-piece (coalesced) with ~8 instructions each.
    const int lw = warp - (kProducerWarps + 1);                  // 0..7
    const bool is_a = lw < 4;
    const int part = lw & 3;                                     // which quarter of the operand
    RowDesc* desc = s_rowdesc[lw];
    for (int ks = 0; ks < k_stages; ++ks) {
      const int stage = ks % num_stages;
      if (ks >= num_stages) mbar_wait(smem_u32(&s_empty[stage]), static_cast<uint32_t>(ks / num_stages - 1) & 1u);
      const uint32_t a_hi = smem_base + stage * stage_bytes;
      const uint32_t dst = is_a ? a_hi : a_hi + (SPLIT3 ? 2 : 1) * a_bytes;
      const bool mn = is_a ? A_MN : B_MN;
      const int extent = is_a ? kTileM : n_tile;
      if (!mn) {
        // K-major: tile rows e, 8 pieces (128 B) per row.  This warp owns rows part*32 + 128*j.
        const int k0 = ks * kBK;
        for (int e0 = part * 32; e0 < extent; e0 += 128) {
          __syncwarp();
          desc[lane] = is_a ? load_a.row(e0 + lane, ks) : load_b.row(e0 + lane, ks);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int el = 4 * i + (lane >> 3), c = lane & 7, e = e0 + el;
            if (e < extent) {
              const float* src = desc[el].at(k0 + 4 * c);
              cp_async16(dst + e * 128 + ((c ^ (e & 7)) << 4), src ? static_cast<const void*>(src) : gdummy, src ? 16u : 0u);
            }
          }
        }
      } else {
        // MN-major: tile rows k (32 per stage), G*8 pieces per row.  This warp owns k = part*8 .. +7.
        const int groups = (extent + 31) / 32, per_k = groups * 8;
        const int origin = is_a ? load_a.origin() : load_b.origin();
        __syncwarp();
        if (lane < 8) desc[lane] = is_a ? load_a.row(part * 8 + lane, ks) : load_b.row(part * 8 + lane, ks);
        __syncwarp();
        for (int kl = 0; kl < 8; ++kl) {
          const int k = part * 8 + kl;
          const uint32_t row_dst = dst + (k >> 2) * groups * 512 + (k & 3) * 128;
          for (int cc = lane; cc < per_k; cc += 32) {
            const int g = cc >> 3, c = cc & 7;
            const float* src = desc[kl].at(origin + g * 32 + 4 * c);
            cp_async16(row_dst + g * 512 + ((((c >> 1) ^ (k & 3)) << 5) | ((c & 1) << 4)),
                       src ? static_cast<const void*>(src) : gdummy, src ? 16u : 0u);
          }
        }
      }
      cp_async_mbar_arrive_noinc(smem_u32(&s_landed[stage]));
      if (lw == 0) GS_TRACE(19 + 4 * ks);
    }
  }
  __syncthreads();
  if (warp == kProducerWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, tmem_cols);
  }
}

template <bool SPLIT3, bool ASYNC>
__global__ void __launch_bounds__(kThreads, 1)
sage_fwd_tc_kernel(XView x, const float* __restrict__ weight, int64_t ldw, int out_dim, bool vec_ok,
                   const int32_t* __restrict__ num_rows_dev, int max_rows, float* __restrict__ out, int64_t ld_out,
                   int relu, int n_tile, int k_stages, int num_stages) {
  extern __shared__ unsigned char smem_dyn[];
  const int rows = live_rows(num_rows_dev, max_rows);
  const int row0 = blockIdx.x * kTileM, h0 = blockIdx.y * n_tile;
  if (row0 >= rows) return;
  const int nt = min(n_tile, ((out_dim - h0) + 15) & ~15);
  __shared__ int32_t s_rowidx[kTileM];
  x.fill_cache(s_rowidx, row0, kTileM, rows);            // made visible by the __syncthreads in gemm_core
  LoadX_K la{x, row0, rows};
  LoadW_K lb{x, weight, ldw, h0, out_dim, vec_ok};
  StoreOut epi{out, ld_out, row0, rows, h0, out_dim, relu};
  gemm_core<false, false, SPLIT3, ASYNC>(la, lb, epi, nt, k_stages, num_stages, smem_dyn, weight);
}

template <bool SPLIT3, bool ASYNC>
__global__ void __launch_bounds__(kThreads, 1)
sage_bwd_x_tc_kernel(const float* __restrict__ grad_out, int64_t ld_go, const float* __restrict__ out, int64_t ld_out,
                     const float* __restrict__ weight, int64_t ldw, int dim, int out_dim, int gcn, int relu, bool vec_ok,
                     const int32_t* __restrict__ num_rows_dev, int max_rows, float* __restrict__ grad_self, int64_t ld_gs,
                     float* __restrict__ grad_agg, int64_t ld_ga, int n_tile, int k_stages, int num_stages) {
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

template <bool SPLIT3, bool ASYNC>
__global__ void __launch_bounds__(kThreads, 1)
sage_bwd_w_tc_kernel(XView x, const float* __restrict__ grad_out, int64_t ld_go, const float* __restrict__ out,
                     int64_t ld_out, int out_dim, int relu, const int32_t* __restrict__ num_rows_dev, int max_rows,
                     int rows_per_chunk, float* __restrict__ grad_w, int64_t ldw, int n_tile, int num_stages) {
  extern __shared__ unsigned char smem_dyn[];
  const int rows = live_rows(num_rows_dev, max_rows);
  const int kv0 = blockIdx.x * n_tile, h0 = blockIdx.y * kTileM;
  const int r_begin = blockIdx.z * rows_per_chunk;
  const int r_end = min(rows, r_begin + rows_per_chunk);
  if (r_begin >= r_end) return;
  const int nt = min(n_tile, ((x.kv_total() - kv0) + 15) & ~15);
  const int k_stages = (r_end - r_begin + kBK - 1) / kBK;
  __shared__ int32_t s_rowidx[kMaxChunkRows];
  x.fill_cache(s_rowidx, r_begin, min(rows_per_chunk, kMaxChunkRows), rows);
  LoadDZ_MN la{grad_out, ld_go, out, ld_out, r_begin, r_end, h0, out_dim, relu};
  LoadX_MN lb{x, r_begin, r_end, kv0};
  AddDW epi{x, grad_w, ldw, h0, out_dim, kv0};
  gemm_core<true, true, SPLIT3, ASYNC>(la, lb, epi, nt, k_stages, num_stages, smem_dyn, grad_out);
}

struct Plan { int n_tile, num_stages, smem; };

static Plan make_plan(int n_total, bool a_mn, bool b_mn, bool split3) {
  Plan p{};
  int n_tile = (n_total + 15) & ~15;
  if (n_tile > 256) n_tile = 256;
  p.n_tile = n_tile;
  const int stage = (split3 ? 2 : 1) * (tile_bytes(kTileM, a_mn) + tile_bytes(n_tile, b_mn));
  int stages = kSmemBudget / stage;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 1) stages = 1;
  p.num_stages = stages;
  p.smem = stages * stage + 1024;
  return p;
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
      KERNEL<true, true><<<GRID, kThreads, SMEM, STREAM>>>(__VA_ARGS__); }                                  \
    else if (split3) { if ((e_ = set_smem(KERNEL<true, false>, SMEM))) return e_;                           \
      KERNEL<true, false><<<GRID, kThreads, SMEM, STREAM>>>(__VA_ARGS__); }                                 \
    else if (async) { if ((e_ = set_smem(KERNEL<false, true>, SMEM))) return e_;                            \
      KERNEL<false, true><<<GRID, kThreads, SMEM, STREAM>>>(__VA_ARGS__); }                                 \
    else { if ((e_ = set_smem(KERNEL<false, false>, SMEM))) return e_;                                      \
      KERNEL<false, false><<<GRID, kThreads, SMEM, STREAM>>>(__VA_ARGS__); }                                \
  } while (0)

int gs_sage_gemm_fwd_tc(const float* self_table, int64_t ld_self, const int32_t* self_idx, const float* agg,
                        int64_t ld_agg, int32_t dim, const float* weight, int64_t ldw, int32_t out_dim, int32_t gcn,
                        const int32_t* num_rows_dev, int32_t max_rows, float* out, int64_t ld_out, int32_t relu,
                        int32_t precision, gs_stream_t stream) {
  const bool split3 = precision == GS_PREC_TF32X3;
  if (precision != GS_PREC_TF32 && !split3) return GS_ERR_BAD_ARG;
  XView x{self_table, ld_self, self_idx, agg, ld_agg, dim, (dim + 3) & ~3, gcn, nullptr, 0, 0};
  const int kt = gcn ? x.dim_pad : 2 * x.dim_pad;
  const Plan p = make_plan(out_dim, false, false, split3);
  const int k_stages = (kt + kBK - 1) / kBK;
  const bool vec_ok = (dim % 4 == 0) && (ldw % 4 == 0) && aligned16(weight);
  const bool async = vec_ok;                       // X rows are always 16-byte aligned (padded tables)
  dim3 grid((max_rows + kTileM - 1) / kTileM, (out_dim + p.n_tile - 1) / p.n_tile);
  GS_TC_LAUNCH(sage_fwd_tc_kernel, grid, p.smem, as_stream(stream), x, weight, ldw, out_dim, vec_ok, num_rows_dev,
               max_rows, out, ld_out, relu, p.n_tile, k_stages, p.num_stages);
  return finish_launch();
}

int gs_sage_gemm_bwd_x_tc(const float* grad_out, int64_t ld_go, const float* out, int64_t ld_out, const float* weight,
                          int64_t ldw, int32_t dim, int32_t out_dim, int32_t gcn, int32_t relu,
                          const int32_t* num_rows_dev, int32_t max_rows, float* grad_self, int64_t ld_gs,
                          float* grad_agg, int64_t ld_ga, int32_t precision, gs_stream_t stream) {
  const bool split3 = precision == GS_PREC_TF32X3;
  if (precision != GS_PREC_TF32 && !split3) return GS_ERR_BAD_ARG;
  const int ncols = gcn ? dim : 2 * dim;
  const Plan p = make_plan(ncols, false, true, split3);
  const int k_stages = (out_dim + kBK - 1) / kBK;
  const bool vec_ok = (ldw % 4 == 0) && aligned16(weight);
  const bool async = vec_ok && !relu && (ncols % 4 == 0) && (out_dim % 4 == 0) && (ld_go % 4 == 0) && aligned16(grad_out);
  dim3 grid((max_rows + kTileM - 1) / kTileM, (ncols + p.n_tile - 1) / p.n_tile);
  GS_TC_LAUNCH(sage_bwd_x_tc_kernel, grid, p.smem, as_stream(stream), grad_out, ld_go, out, ld_out, weight, ldw, dim,
               out_dim, gcn, relu, vec_ok, num_rows_dev, max_rows, grad_self, ld_gs, grad_agg, ld_ga, p.n_tile, k_stages,
               p.num_stages);
  return finish_launch();
}

int gs_sage_gemm_bwd_w_tc(const float* self_table, int64_t ld_self, const int32_t* self_idx, const float* agg,
                          int64_t ld_agg, int32_t dim, const float* grad_out, int64_t ld_go, const float* out,
                          int64_t ld_out, int32_t out_dim, int32_t gcn, int32_t relu, const int32_t* num_rows_dev,
                          int32_t max_rows, float* grad_w, int64_t ldw, int32_t precision, gs_stream_t stream) {
  const bool split3 = precision == GS_PREC_TF32X3;
  if (precision != GS_PREC_TF32 && !split3) return GS_ERR_BAD_ARG;
  XView x{self_table, ld_self, self_idx, agg, ld_agg, dim, (dim + 3) & ~3, gcn, nullptr, 0, 0};
  const int kt = gcn ? x.dim_pad : 2 * x.dim_pad;
  const Plan p = make_plan(kt, true, true, split3);
  const int tiles = ((kt + p.n_tile - 1) / p.n_tile) * ((out_dim + kTileM - 1) / kTileM);
  int chunks = (kNumSMs + tiles - 1) / tiles;                 // about one CTA per SM
  const int max_chunks = (max_rows + 4 * kBK - 1) / (4 * kBK);   // at least 4 k-stages per CTA
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  int rows_per_chunk = (max_rows + chunks - 1) / chunks;
  rows_per_chunk = ((rows_per_chunk + kBK - 1) / kBK) * kBK;
  if (rows_per_chunk > kMaxChunkRows) rows_per_chunk = kMaxChunkRows;
  chunks = (max_rows + rows_per_chunk - 1) / rows_per_chunk;
  const bool async = !relu && (out_dim % 4 == 0) && (ld_go % 4 == 0) && aligned16(grad_out);
  dim3 grid((kt + p.n_tile - 1) / p.n_tile, (out_dim + kTileM - 1) / kTileM, chunks);
  GS_TC_LAUNCH(sage_bwd_w_tc_kernel, grid, p.smem, as_stream(stream), x, grad_out, ld_go, out, ld_out, out_dim, relu,
               num_rows_dev, max_rows, rows_per_chunk, grad_w, ldw, p.n_tile, p.num_stages);
  return finish_launch();
}

#ifdef GS_TC_TRACE
extern "C" int gs_debug_trace_read(long long* host_out, int n) {
  if (n > 512) n = 512;
  cudaDeviceSynchronize();
  cudaError_t e = cudaMemcpyFromSymbol(host_out, gs::tc::g_trace, sizeof(long long) * n);
  return static_cast<int>(e);
}
#endif
