// Stand-alone probe for K3 (layer-1 gather-segment-reduce) variants at the cfg-3 shape.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o scratch/agg_probe scratch/agg_probe.cu
// Times n back-to-back launches on DISTINCT frontiers between one CUDA-event pair (what bench.py does)
// and the same kernel at a saturating size.  Scratch code: the winner moves into csrc/aggregate.cu.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ int live_rows(const int32_t* p, int max_rows) {
  if (p == nullptr) return max_rows;
  int n = __ldg(p);
  return n < max_rows ? n : max_rows;
}
template <uint32_t FILL>
__device__ __forceinline__ float4 ldg_stream_f4_if(const void* p, bool on) {
  float4 v;
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "mov.b32 %0, %6;\n\t"
      "mov.b32 %1, %6;\n\t"
      "mov.b32 %2, %6;\n\t"
      "mov.b32 %3, %6;\n\t"
      "@q ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];\n\t"
      "}"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "l"(p), "r"(static_cast<int>(on)), "n"(FILL));
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

constexpr int kAggWarps = 8;
constexpr int kBatch = 12;

// ---- A: the shipped register-staged kernel (MEAN only) ----
template <int MINB>
__global__ void __launch_bounds__(kAggWarps * 32, MINB)
agg_base(const float* __restrict__ table, uint32_t ld_bytes, int dim4, const int32_t* __restrict__ nbr, int stride,
         const int32_t* __restrict__ cnt, const int32_t* __restrict__ num_rows_dev, int max_rows,
         float* __restrict__ out, int64_t ld_out) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * kAggWarps + (threadIdx.x >> 5);
  if (r >= max_rows) return;
  const int32_t* row_ids = nbr + static_cast<int64_t>(r) * stride;
  const int n_raw = __ldg(cnt + r);
  const int mine_raw = lane < stride ? __ldg(row_ids + lane) : -1;
  if (r >= live_rows(num_rows_dev, max_rows)) return;
  const int n = min(n_raw, stride);
  const float inv = 1.0f / static_cast<float>(n);
  const char* tbase = reinterpret_cast<const char*>(table);
  for (int cbase = 0; cbase < dim4; cbase += 32) {
    const int c4 = cbase + lane;
    const bool active = c4 < dim4;
    const char* col = tbase + 16 * c4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int mine = lane < n ? mine_raw : -1;
    for (int j0 = 0; j0 < n; j0 += kBatch) {
      float4 v[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int got = __shfl_sync(0xffffffffu, mine, (j0 + u) & 31);
        const int id = (active && j0 + u < n) ? got : -1;
        v[u] = ldg_stream_f4_if<0u>(col + static_cast<size_t>(static_cast<uint32_t>(id)) * ld_bytes, id >= 0);
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    if (active) {
      acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
      *reinterpret_cast<float4*>(out + static_cast<int64_t>(r) * ld_out + 4 * c4) = acc;
    }
  }
}

// ---- B: persistent warps, id look-ahead of LA rows, optional L2 prefetch of the looked-ahead rows ----
// PFG = prefetch granule in bytes (0: no prefetch)
template <int LA, int PFG, int MINB, int KB>
__global__ void __launch_bounds__(kAggWarps * 32, MINB)
agg_persist(const float* __restrict__ table, uint32_t ld_bytes, int dim4, const int32_t* __restrict__ nbr, int stride,
            const int32_t* __restrict__ cnt, const int32_t* __restrict__ num_rows_dev, int max_rows,
            float* __restrict__ out, int64_t ld_out) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int W = gridDim.x * kAggWarps;
  const int gw = blockIdx.x * kAggWarps + (threadIdx.x >> 5);
  const int rows = live_rows(num_rows_dev, max_rows);
  const char* tbase = reinterpret_cast<const char*>(table);
  const uint32_t row_bytes = 16u * dim4;

  auto fetch = [&](int row) -> int {
    int mine = -1;
    if (row < rows) {
      const int n = __ldg(cnt + row);
      const int v = lane < stride ? __ldg(nbr + static_cast<int64_t>(row) * stride + lane) : -1;
      mine = lane < n ? v : -1;
    }
    return mine;
  };
  auto prefetch = [&](int mine) {
    if (PFG == 0) return;
    const int per_row = (row_bytes + PFG - 1) / (PFG > 0 ? PFG : 1);
    const int n = __popc(__ballot_sync(0xffffffffu, mine >= 0));
    const int total = n * per_row;
    for (int p0 = 0; p0 < total; p0 += 32) {
      const int p = p0 + lane;
      const int j = p / per_row, q = p - j * per_row;
      const int id = __shfl_sync(0xffffffffu, mine, j & 31);
      if (p < total) prefetch_l2(tbase + static_cast<size_t>(static_cast<uint32_t>(id)) * ld_bytes + q * PFG);
    }
  };

  int q[LA];
#pragma unroll
  for (int k = 0; k < LA; ++k) q[k] = fetch(gw + k * W);
#pragma unroll
  for (int k = 0; k < LA; ++k) prefetch(q[k]);

  for (int r = gw; r < rows; r += W) {
    const int mine = q[0];
#pragma unroll
    for (int k = 0; k + 1 < LA; ++k) q[k] = q[k + 1];
    q[LA - 1] = fetch(r + LA * W);
    const int n = __popc(__ballot_sync(0xffffffffu, mine >= 0));
    const float inv = 1.0f / static_cast<float>(n);
    for (int cbase = 0; cbase < dim4; cbase += 32) {
      const int c4 = cbase + lane;
      const bool active = c4 < dim4;
      const char* col = tbase + 16 * c4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j0 = 0; j0 < n; j0 += KB) {
        float4 v[KB];
#pragma unroll
        for (int u = 0; u < KB; ++u) {
          const int got = __shfl_sync(0xffffffffu, mine, (j0 + u) & 31);
          const int id = (active && j0 + u < n) ? got : -1;
          v[u] = ldg_stream_f4_if<0u>(col + static_cast<size_t>(static_cast<uint32_t>(id)) * ld_bytes, id >= 0);
        }
        if (cbase == 0 && j0 == 0) prefetch(q[LA - 1]);     // behind this row's loads, ahead of their use
#pragma unroll
        for (int u = 0; u < KB; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
      }
      if (active) {
        acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
        *reinterpret_cast<float4*>(out + static_cast<int64_t>(r) * ld_out + 4 * c4) = acc;
      }
    }
  }
}

// ---- C: non-persistent, but every warp ALSO prefetches into L2 the rows of the warp `ahead` rows later ----
template <int PFG>
__global__ void __launch_bounds__(kAggWarps * 32, 3)
agg_base_pf(const float* __restrict__ table, uint32_t ld_bytes, int dim4, const int32_t* __restrict__ nbr, int stride,
            const int32_t* __restrict__ cnt, const int32_t* __restrict__ num_rows_dev, int max_rows,
            float* __restrict__ out, int64_t ld_out, int ahead) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * kAggWarps + (threadIdx.x >> 5);
  if (r >= max_rows) return;
  const int rows = live_rows(num_rows_dev, max_rows);
  const int32_t* row_ids = nbr + static_cast<int64_t>(r) * stride;
  const int n_raw = __ldg(cnt + r);
  const int mine_raw = lane < stride ? __ldg(row_ids + lane) : -1;
  const int r2 = r + ahead;
  int n2 = 0, mine2 = -1;
  if (r2 < rows) {
    n2 = min(__ldg(cnt + r2), stride);
    mine2 = lane < stride ? __ldg(nbr + static_cast<int64_t>(r2) * stride + lane) : -1;
  }
  if (r >= rows) return;
  const int n = min(n_raw, stride);
  const float inv = 1.0f / static_cast<float>(n);
  const char* tbase = reinterpret_cast<const char*>(table);
  const uint32_t row_bytes = 16u * dim4;
  for (int cbase = 0; cbase < dim4; cbase += 32) {
    const int c4 = cbase + lane;
    const bool active = c4 < dim4;
    const char* col = tbase + 16 * c4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int mine = lane < n ? mine_raw : -1;
    for (int j0 = 0; j0 < n; j0 += kBatch) {
      float4 v[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int got = __shfl_sync(0xffffffffu, mine, (j0 + u) & 31);
        const int id = (active && j0 + u < n) ? got : -1;
        v[u] = ldg_stream_f4_if<0u>(col + static_cast<size_t>(static_cast<uint32_t>(id)) * ld_bytes, id >= 0);
      }
      if (cbase == 0 && j0 == 0 && n2 > 0) {
        const int per_row = (row_bytes + PFG - 1) / PFG;
        const int total = n2 * per_row;
        for (int p0 = 0; p0 < total; p0 += 32) {
          const int p = p0 + lane;
          const int j = p / per_row, qq = p - j * per_row;
          const int id = __shfl_sync(0xffffffffu, mine2, j & 31);
          if (p < total) prefetch_l2(tbase + static_cast<size_t>(static_cast<uint32_t>(id)) * ld_bytes + qq * PFG);
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    if (active) {
      acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
      *reinterpret_cast<float4*>(out + static_cast<int64_t>(r) * ld_out + 4 * c4) = acc;
    }
  }
}


// ---- D: persistent, KB loads per batch, index loads not serialised behind the live-row count,
//         optional L2 prefetch of (a) the rest of the CURRENT row beyond the first batch (PFSELF) and
//         (b) the NEXT row of this warp (PFNEXT); EF = L2 evict_first on the gather loads ----
template <uint32_t FILL, int EF>
__device__ __forceinline__ float4 ldg_gather(const void* p, bool on) {
  float4 v;
  if (EF) {
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\tmov.b32 %0, %6;\n\tmov.b32 %1, %6;\n\tmov.b32 %2, %6;\n\tmov.b32 %3, %6;\n\t"
        "@q ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];\n\t}"
        : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "r"(static_cast<int>(on)), "n"(FILL));
  } else {
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\tmov.b32 %0, %6;\n\tmov.b32 %1, %6;\n\tmov.b32 %2, %6;\n\tmov.b32 %3, %6;\n\t"
        "@q ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];\n\t}"
        : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "r"(static_cast<int>(on)), "n"(FILL));
  }
  return v;
}

template <int KB, int MINB, int PFSELF, int PFNEXT, int EF>
__global__ void __launch_bounds__(kAggWarps * 32, MINB)
agg_d(const float* __restrict__ table, uint32_t ld_bytes, int dim4, const int32_t* __restrict__ nbr, int stride,
      const int32_t* __restrict__ cnt, const int32_t* __restrict__ num_rows_dev, int max_rows,
      float* __restrict__ out, int64_t ld_out) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int W = gridDim.x * kAggWarps;
  const int gw = blockIdx.x * kAggWarps + (threadIdx.x >> 5);
  const char* tbase = reinterpret_cast<const char*>(table);
  const uint32_t row_bytes = 16u * dim4;
  const int per_row = (row_bytes + 127) / 128;

  // raw index loads for a row < max_rows: independent of the live-row count (one latency for all three)
  auto fetch_raw = [&](int row, int& n) -> int {
    int v = -1; n = 0;
    if (row < max_rows) {
      n = __ldg(cnt + row);
      v = lane < stride ? __ldg(nbr + static_cast<int64_t>(row) * stride + lane) : -1;
    }
    return v;
  };
  auto prefetch_from = [&](int mine, int n, int first) {     // L2-prefetch neighbours [first, n) of a row
    const int total = (n - first) * per_row;
    for (int p0 = 0; p0 < total; p0 += 32) {
      const int p = p0 + lane;
      const int j = first + p / per_row, q = p % per_row;
      const int id = __shfl_sync(0xffffffffu, mine, j & 31);
      if (p < total) prefetch_l2(tbase + static_cast<size_t>(static_cast<uint32_t>(id)) * ld_bytes + q * 128);
    }
  };

  int n_next;
  int v_next = fetch_raw(gw, n_next);
  const int rows = live_rows(num_rows_dev, max_rows);
  for (int r = gw; r < rows; r += W) {
    const int n = min(n_next, stride);
    const int mine = lane < n ? v_next : -1;
    v_next = fetch_raw(r + W, n_next);
    const bool have_next = PFNEXT && (r + W < rows);
    const float inv = 1.0f / static_cast<float>(n);
    for (int cbase = 0; cbase < dim4; cbase += 32) {
      const int c4 = cbase + lane;
      const bool active = c4 < dim4;
      const char* col = tbase + 16 * c4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j0 = 0; j0 < n; j0 += KB) {
        float4 v[KB];
#pragma unroll
        for (int u = 0; u < KB; ++u) {
          const int got = __shfl_sync(0xffffffffu, mine, (j0 + u) & 31);
          const int id = (active && j0 + u < n) ? got : -1;
          v[u] = ldg_gather<0u, EF>(col + static_cast<size_t>(static_cast<uint32_t>(id)) * ld_bytes, id >= 0);
        }
        if (cbase == 0 && j0 == 0) {
          if (PFSELF && n > KB) prefetch_from(mine, n, KB);
          if (have_next) {
            const int nn = min(n_next, stride);
            prefetch_from(lane < nn ? v_next : -1, nn, 0);
          }
        }
#pragma unroll
        for (int u = 0; u < KB; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
      }
      if (active) {
        acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
        *reinterpret_cast<float4*>(out + static_cast<int64_t>(r) * ld_out + 4 * c4) = acc;
      }
    }
  }
}

// ---- E: non-persistent with KB loads per batch (one row per warp), optional PFSELF ----
template <int KB, int MINB, int PFSELF, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, MINB)
agg_e(const float* __restrict__ table, uint32_t ld_bytes, int dim4, const int32_t* __restrict__ nbr, int stride,
      const int32_t* __restrict__ cnt, const int32_t* __restrict__ num_rows_dev, int max_rows,
      float* __restrict__ out, int64_t ld_out) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (r >= max_rows) return;
  const int n_raw = __ldg(cnt + r);
  const int mine_raw = lane < stride ? __ldg(nbr + static_cast<int64_t>(r) * stride + lane) : -1;
  if (r >= live_rows(num_rows_dev, max_rows)) return;
  const int n = min(n_raw, stride);
  const float inv = 1.0f / static_cast<float>(n);
  const char* tbase = reinterpret_cast<const char*>(table);
  const int per_row = (16 * dim4 + 127) / 128;
  const int mine = lane < n ? mine_raw : -1;
  for (int cbase = 0; cbase < dim4; cbase += 32) {
    const int c4 = cbase + lane;
    const bool active = c4 < dim4;
    const char* col = tbase + 16 * c4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j0 = 0; j0 < n; j0 += KB) {
      float4 v[KB];
#pragma unroll
      for (int u = 0; u < KB; ++u) {
        const int got = __shfl_sync(0xffffffffu, mine, (j0 + u) & 31);
        const int id = (active && j0 + u < n) ? got : -1;
        v[u] = ldg_gather<0u, 0>(col + static_cast<size_t>(static_cast<uint32_t>(id)) * ld_bytes, id >= 0);
      }
      if (PFSELF && cbase == 0 && j0 == 0 && n > KB) {
        const int total = (n - KB) * per_row;
        for (int p0 = 0; p0 < total; p0 += 32) {
          const int p = p0 + lane;
          const int j = KB + p / per_row, q = p % per_row;
          const int id = __shfl_sync(0xffffffffu, mine, j & 31);
          if (p < total) prefetch_l2(tbase + static_cast<size_t>(static_cast<uint32_t>(id)) * ld_bytes + q * 128);
        }
      }
#pragma unroll
      for (int u = 0; u < KB; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    if (active) {
      acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
      *reinterpret_cast<float4*>(out + static_cast<int64_t>(r) * ld_out + 4 * c4) = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------
struct Frontier {
  int32_t *nbr, *cnt, *num_rows;
  float* out;
  int max_rows, rows;
  double bytes;
};

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static inline uint32_t rnd() {
  rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
  return static_cast<uint32_t>(rng_state >> 33);
}

static Frontier make_frontier(int max_rows, int rows, int stride, int n_nodes, int dim) {
  std::vector<int32_t> nbr(static_cast<size_t>(max_rows) * stride, -1), cnt(max_rows, 0);
  long long nnz = 0;
  for (int r = 0; r < rows; ++r) {
    int ids[32];
    for (int j = 0; j < stride; ++j) ids[j] = static_cast<int>((static_cast<uint64_t>(rnd()) * n_nodes) >> 31) % n_nodes;
    std::sort(ids, ids + stride);
    for (int j = 0; j < stride; ++j) nbr[static_cast<size_t>(r) * stride + j] = ids[j];
    cnt[r] = stride;
    nnz += stride;
  }
  Frontier f{};
  f.max_rows = max_rows; f.rows = rows;
  CK(cudaMalloc(&f.nbr, nbr.size() * 4)); CK(cudaMalloc(&f.cnt, cnt.size() * 4)); CK(cudaMalloc(&f.num_rows, 4));
  CK(cudaMalloc(&f.out, static_cast<size_t>(max_rows) * dim * 4));
  CK(cudaMemcpy(f.nbr, nbr.data(), nbr.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(f.cnt, cnt.data(), cnt.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(f.num_rows, &rows, 4, cudaMemcpyHostToDevice));
  f.bytes = static_cast<double>(nnz) * dim * 4 + static_cast<double>(rows) * dim * 4 + nnz * 4.0 + (rows + 1) * 4.0;
  return f;
}

__global__ void fill_table(float* t, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    t[i] = static_cast<float>((i * 2654435761ull) & 0xffff) * (1.0f / 65536.0f) - 0.5f;
}

template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kernel)(KArgs...), int grid, int block, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

struct Variant {
  const char* name;
  // launches the variant on a frontier
  void (*run)(const float* table, uint32_t ld_bytes, int dim4, const Frontier& f, int stride, int64_t ld_out, cudaStream_t st, int knob);
  int knob;
};

template <int MINB>
static void run_base(const float* t, uint32_t ldb, int d4, const Frontier& f, int stride, int64_t ldo, cudaStream_t st, int) {
  launch_pdl(agg_base<MINB>, (f.max_rows + kAggWarps - 1) / kAggWarps, kAggWarps * 32, st, t, ldb, d4, f.nbr, stride, f.cnt, f.num_rows, f.max_rows, f.out, ldo);
}
template <int LA, int PFG, int MINB, int KB = 12>
static void run_persist(const float* t, uint32_t ldb, int d4, const Frontier& f, int stride, int64_t ldo, cudaStream_t st, int ctas_per_sm) {
  int grid = 148 * ctas_per_sm;
  const int need = (f.max_rows + kAggWarps - 1) / kAggWarps;
  if (grid > need) grid = need;
  launch_pdl(agg_persist<LA, PFG, MINB, KB>, grid, kAggWarps * 32, st, t, ldb, d4, f.nbr, stride, f.cnt, f.num_rows, f.max_rows, f.out, ldo);
}
template <int PFG>
static void run_base_pf(const float* t, uint32_t ldb, int d4, const Frontier& f, int stride, int64_t ldo, cudaStream_t st, int ahead) {
  launch_pdl(agg_base_pf<PFG>, (f.max_rows + kAggWarps - 1) / kAggWarps, kAggWarps * 32, st, t, ldb, d4, f.nbr, stride, f.cnt, f.num_rows, f.max_rows, f.out, ldo, ahead);
}

// knob > 0: CTAs per SM; knob < 0: balanced grid for at most -knob CTAs per SM (every warp gets the same row count)
template <int KB, int MINB, int PFSELF, int PFNEXT, int EF>
static void run_d(const float* t, uint32_t ldb, int d4, const Frontier& f, int stride, int64_t ldo, cudaStream_t st, int knob) {
  const int need = (f.max_rows + kAggWarps - 1) / kAggWarps;
  int grid;
  if (knob > 0) grid = 148 * knob;
  else {
    const int wmax = 148 * (-knob) * kAggWarps;
    const int k = (f.max_rows + wmax - 1) / wmax;
    const int w = (f.max_rows + k - 1) / k;
    grid = (w + kAggWarps - 1) / kAggWarps;
  }
  if (grid > need) grid = need;
  launch_pdl(agg_d<KB, MINB, PFSELF, PFNEXT, EF>, grid, kAggWarps * 32, st, t, ldb, d4, f.nbr, stride, f.cnt, f.num_rows, f.max_rows, f.out, ldo);
}
template <int KB, int MINB, int PFSELF, int WARPS>
static void run_e(const float* t, uint32_t ldb, int d4, const Frontier& f, int stride, int64_t ldo, cudaStream_t st, int) {
  launch_pdl(agg_e<KB, MINB, PFSELF, WARPS>, (f.max_rows + WARPS - 1) / WARPS, WARPS * 32, st, t, ldb, d4, f.nbr, stride, f.cnt, f.num_rows, f.max_rows, f.out, ldo);
}

int main(int argc, char** argv) {
  const int n_nodes = 2449029, dim = 100, stride = 10;
  const int ld = (argc > 1) ? atoi(argv[1]) : 100;       // floats per table row (100 = shipped layout)
  const int dim4 = dim / 4;
  const uint32_t ld_bytes = ld * 4;
  if (argc > 2) {
    size_t g = 0;
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, atoi(argv[2]));
    cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    printf("L2 fetch granularity: requested %s -> %zu (%s)\n", argv[2], g, cudaGetErrorString(e));
  }
  const int only_top = argc > 3;
  const int carve = argc > 4 ? atoi(argv[4]) : -1;
  if (carve >= -1) {
    cudaFuncSetAttribute(agg_d<6, 5, 0, 0, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    cudaFuncSetAttribute(agg_base<3>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    cudaFuncSetAttribute(agg_persist<2, 0, 5, 6>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    printf("carveout %d\n", carve);
  }
  const int n_front = 16, max_rows = 11264, rows = 10900;
  const int big_rows = 87874;
  float* table;
  CK(cudaMalloc(&table, static_cast<size_t>(n_nodes) * ld * 4));
  fill_table<<<148 * 8, 256>>>(table, static_cast<size_t>(n_nodes) * ld);
  CK(cudaDeviceSynchronize());
  std::vector<Frontier> fr;
  for (int i = 0; i < n_front; ++i) fr.push_back(make_frontier(max_rows, rows, stride, n_nodes, dim));
  Frontier big = make_frontier(big_rows, big_rows, stride, n_nodes, dim);
  std::vector<float> ref(static_cast<size_t>(rows) * dim), got(static_cast<size_t>(rows) * dim);

  std::vector<Variant> vs = {
      {"A base minb3 (shipped)", run_base<3>, 0},
      {"B persist LA2 nopf 5cta KB6", run_persist<2, 0, 5, 6>, 5},
      {"D KB6 5cta", run_d<6, 5, 0, 0, 0>, 5},
      {"D KB5 5cta", run_d<5, 5, 0, 0, 0>, 5},
      {"D KB5 5cta balanced", run_d<5, 5, 0, 0, 0>, -5},
      {"D KB5 5cta pfself", run_d<5, 5, 1, 0, 0>, 5},
      {"D KB5 5cta pfnext", run_d<5, 5, 0, 1, 0>, 5},
      {"D KB5 5cta pfself+next", run_d<5, 5, 1, 1, 0>, 5},
      {"D KB5 6cta(minb6)", run_d<5, 6, 0, 0, 0>, 6},
      {"D KB5 6cta pfself", run_d<5, 6, 1, 0, 0>, 6},
      {"D KB4 6cta", run_d<4, 6, 0, 0, 0>, 6},
      {"D KB4 7cta", run_d<4, 7, 0, 0, 0>, 7},
      {"D KB3 8cta", run_d<3, 8, 0, 0, 0>, 8},
      {"D KB3 8cta pfself", run_d<3, 8, 1, 0, 0>, 8},
      {"D KB4 6cta pfself", run_d<4, 6, 1, 0, 0>, 6},
      {"D KB6 5cta pfself", run_d<6, 5, 1, 0, 0>, 5},
      {"D KB10 3cta", run_d<10, 3, 0, 0, 0>, 3},
      {"D KB10 4cta(minb4)", run_d<10, 4, 0, 0, 0>, 4},
      {"D KB5 4cta", run_d<5, 5, 0, 0, 0>, 4},
      {"D KB5 4cta pfself", run_d<5, 5, 1, 0, 0>, 4},
      {"E nonpersist KB5 minb5 8w", run_e<5, 5, 0, 8>, 0},
      {"E nonpersist KB5 minb5 8w pfself", run_e<5, 5, 1, 8>, 0},
      {"E nonpersist KB5 minb10 4w", run_e<5, 10, 0, 4>, 0},
      {"E nonpersist KB5 minb10 4w pfself", run_e<5, 10, 1, 4>, 0},
      {"E nonpersist KB4 minb12 4w", run_e<4, 12, 0, 4>, 0},
      {"E nonpersist KB10 minb6 4w", run_e<10, 6, 0, 4>, 0},
      {"E nonpersist KB5 minb20 2w", run_e<5, 20, 0, 2>, 0},
  };

  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  double bytes = 0;
  for (auto& f : fr) bytes += f.bytes;
  bytes /= n_front;
  printf("ld=%d floats  rows=%d  bytes/launch=%.1f MB  big: rows=%d bytes=%.1f MB\n", ld, rows, bytes / 1e6, big_rows, big.bytes / 1e6);
  for (size_t vi = 0; vi < vs.size(); ++vi) {
    auto& v = vs[vi];
    if (only_top && vi > 2) continue;
    for (int i = 0; i < 3; ++i) v.run(table, ld_bytes, dim4, fr[i], stride, dim, st, v.knob);
    CK(cudaStreamSynchronize(st));
    float best = 1e9f, sum = 0;
    const int reps = 5;
    for (int rep = 0; rep < reps; ++rep) {
      CK(cudaEventRecord(a, st));
      for (auto& f : fr) v.run(table, ld_bytes, dim4, f, stride, dim, st, v.knob);
      CK(cudaEventRecord(b, st));
      CK(cudaStreamSynchronize(st));
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      best = std::min(best, ms); sum += ms;
    }
    const double us_best = best * 1e3 / n_front, us_avg = sum / reps * 1e3 / n_front;
    // correctness against variant 0
    CK(cudaMemcpy(got.data(), fr[0].out, got.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    if (vi == 0) ref = got; else for (size_t i = 0; i < got.size(); ++i) bad += (got[i] != ref[i]);
    // saturating size
    for (int i = 0; i < 2; ++i) v.run(table, ld_bytes, dim4, big, stride, dim, st, v.knob);
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(a, st));
    for (int i = 0; i < 5; ++i) v.run(table, ld_bytes, dim4, big, stride, dim, st, v.knob);
    CK(cudaEventRecord(b, st));
    CK(cudaStreamSynchronize(st));
    float msb; CK(cudaEventElapsedTime(&msb, a, b));
    const double us_big = msb * 1e3 / 5;
    printf("%-36s small: %6.2f us avg %6.2f best  %5.0f GB/s (%.3f)   big: %7.2f us %5.0f GB/s (%.3f)  mismatches %zu\n", v.name,
           us_avg, us_best, bytes / us_avg / 1e3, bytes / us_avg / 1e3 / 6544.0, us_big, big.bytes / us_big / 1e3,
           big.bytes / us_big / 1e3 / 6544.0, bad);
    fflush(stdout);
  }
  return 0;
}
