// K2 unique + remap: LSD radix sort (8-bit digits) -> run-length heads -> binary-search remap.
// Replaces the Python set.union / dict(zip) of src/models.py:286-288 and the per-element
// dict lookups of :306 / :274.  Integer work, bit-exact by construction; output order is
// ascending node id (the canonical form of SURVEY.md §8a row A2).
//
// Two paths:
//   * one CTA (1024 threads) that keeps both key buffers in shared memory and does sort,
//     unique and remap in a single launch -- the minibatch frontier (|B|*(fanout+2) ids,
//     12K at b_sz 1024) fits, and a single launch is what the step latency wants;
//   * a multi-CTA path (histogram / scan / stable scatter per digit) for larger inputs.
#include <algorithm>
#include "common.cuh"

namespace gs {

constexpr uint32_t kInvalidKey = 0xFFFFFFFFu;   // (uint32_t)-1: list padding sorts last
constexpr int kSmallThreads = 1024;
constexpr int kSmallWarps = kSmallThreads / 32;
constexpr int kSmallMaxKeys = 24576;            // 2 * 96 KB key buffers + 16 KB counters < 227 KB
constexpr int kTileThreads = 256;
constexpr int kTileWarps = kTileThreads / 32;
constexpr int kTileKeys = 4096;                 // keys per CTA per pass in the multi-CTA path

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

__device__ __forceinline__ uint32_t combined_key(const int32_t* __restrict__ nodes, const int32_t* __restrict__ nbr,
                                                 int rows, int stride, int i) {
  if (i < rows) return static_cast<uint32_t>(nodes[i]);
  const int64_t j = static_cast<int64_t>(i) - rows;
  if (j < static_cast<int64_t>(rows) * stride) return static_cast<uint32_t>(nbr[j]);
  return kInvalidKey;
}

__device__ __forceinline__ int lower_bound_u32(const uint32_t* a, int n, uint32_t v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Stable counting of one warp's contiguous chunk: wcnt[d] += number of keys with digit d.
template <typename CounterT>
__device__ __forceinline__ void warp_digit_count(const uint32_t* keys, int chunk_begin, int chunk_len, int shift,
                                                 CounterT* wcnt, int lane) {
  for (int g = 0; g < chunk_len; g += 32) {
    const uint32_t d = (keys[chunk_begin + g + lane] >> shift) & 255u;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    if (lane == __ffs(peers) - 1) wcnt[d] = static_cast<CounterT>(wcnt[d] + __popc(peers));
    __syncwarp();
  }
}

// Stable scatter of the same chunk: key goes to base[d] + (running count of d in this warp).
template <typename CounterT, typename BaseFn>
__device__ __forceinline__ void warp_digit_scatter(const uint32_t* keys, int chunk_begin, int chunk_len, int shift,
                                                   CounterT* woff, BaseFn base, uint32_t* out, int lane) {
  for (int g = 0; g < chunk_len; g += 32) {
    const uint32_t key = keys[chunk_begin + g + lane];
    const uint32_t d = (key >> shift) & 255u;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const int rank = __popc(peers & lanemask_lt());
    const int64_t pos = base(d) + woff[d] + rank;
    __syncwarp();
    if (lane == __ffs(peers) - 1) woff[d] = static_cast<CounterT>(woff[d] + __popc(peers));
    __syncwarp();
    out[pos] = key;
  }
}

// ---------------------------------------------------------------------------------------
// single-CTA path
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSmallThreads, 1)
unique_small_kernel(const int32_t* __restrict__ nodes, const int32_t* __restrict__ num_rows_dev, int max_rows,
                    const int32_t* __restrict__ nbr, int stride, int passes, int cap_keys,
                    int32_t* __restrict__ uniq, int32_t* __restrict__ num_uniq_dev,
                    int32_t* __restrict__ nbr_idx, int32_t* __restrict__ self_idx) {
  pdl_sync();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* buf_a = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* buf_b = buf_a + cap_keys;
  uint16_t* wcnt = reinterpret_cast<uint16_t*>(buf_b + cap_keys);       // [kSmallWarps][256]
  __shared__ uint32_t s_base[256];
  __shared__ int s_warp_tot[kSmallWarps];
  __shared__ int s_total;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rows = live_rows(num_rows_dev, max_rows);
  const int m = rows * (stride + 1);
  const int m_pad = ((m + kSmallThreads - 1) / kSmallThreads) * kSmallThreads;   // <= cap_keys
  const int chunk = m_pad / kSmallWarps;                                         // multiple of 32

  for (int i = tid; i < m_pad; i += kSmallThreads) buf_a[i] = combined_key(nodes, nbr, rows, stride, i);
  uint32_t* in = buf_a;
  uint32_t* out = buf_b;
  for (int p = 0; p < passes; ++p) {
    const int shift = 8 * p;
    for (int i = tid; i < kSmallWarps * 256; i += kSmallThreads) wcnt[i] = 0;
    __syncthreads();
    warp_digit_count(in, warp * chunk, chunk, shift, wcnt + warp * 256, lane);
    __syncthreads();
    if (tid < 256) {                       // exclusive prefix over warps, per digit
      uint32_t run = 0;
      for (int w = 0; w < kSmallWarps; ++w) {
        const uint32_t c = wcnt[w * 256 + tid];
        wcnt[w * 256 + tid] = static_cast<uint16_t>(run);
        run += c;
      }
      s_base[tid] = run;                   // digit total
    }
    __syncthreads();
    if (warp == 0) {                       // exclusive prefix over the 256 digit totals
      uint32_t v[8], sum = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) { v[q] = s_base[lane * 8 + q]; sum += v[q]; }
      uint32_t incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      uint32_t run = incl - sum;
#pragma unroll
      for (int q = 0; q < 8; ++q) { s_base[lane * 8 + q] = run; run += v[q]; }
    }
    __syncthreads();
    warp_digit_scatter(in, warp * chunk, chunk, shift, wcnt + warp * 256,
                       [&](uint32_t d) { return static_cast<int64_t>(s_base[d]); }, out, lane);
    __syncthreads();
    uint32_t* t = in; in = out; out = t;
  }
  // ---- run-length heads -> compact ----
  const int per = m_pad / kSmallThreads;
  const int beg = tid * per;
  int heads = 0;
  for (int i = beg; i < beg + per; ++i) {
    const uint32_t k = in[i];
    heads += (k != kInvalidKey) && (i == 0 || in[i - 1] != k);
  }
  int incl = heads;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = s_warp_tot[lane];
    int wi = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    s_warp_tot[lane] = wi - v;
    if (lane == 31) s_total = wi;
  }
  __syncthreads();
  int w_at = s_warp_tot[warp] + incl - heads;
  for (int i = beg; i < beg + per; ++i) {
    const uint32_t k = in[i];
    if ((k != kInvalidKey) && (i == 0 || in[i - 1] != k)) {
      out[w_at] = k;
      uniq[w_at] = static_cast<int32_t>(k);
      ++w_at;
    }
  }
  __syncthreads();
  const int n_uniq = s_total;
  if (tid == 0) *num_uniq_dev = n_uniq;
  // ---- remap ----
  if (nbr_idx != nullptr) {
    const int64_t live = static_cast<int64_t>(rows) * stride, all = static_cast<int64_t>(max_rows) * stride;
    for (int64_t i = tid; i < all; i += kSmallThreads) {
      int32_t idx = -1;
      if (i < live) {
        const int32_t id = nbr[i];
        if (id >= 0) idx = lower_bound_u32(out, n_uniq, static_cast<uint32_t>(id));
      }
      nbr_idx[i] = idx;
    }
  }
  if (self_idx != nullptr) {
    for (int i = tid; i < max_rows; i += kSmallThreads)
      self_idx[i] = i < rows ? lower_bound_u32(out, n_uniq, static_cast<uint32_t>(nodes[i])) : -1;
  }
}

// ---------------------------------------------------------------------------------------
// multi-CTA path
// ---------------------------------------------------------------------------------------
__global__ void gather_keys_kernel(const int32_t* __restrict__ nodes, const int32_t* __restrict__ num_rows_dev,
                                   int max_rows, const int32_t* __restrict__ nbr, int stride, int cap_keys,
                                   uint32_t* __restrict__ keys) {
  pdl_sync();
  const int rows = live_rows(num_rows_dev, max_rows);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cap_keys; i += gridDim.x * blockDim.x)
    keys[i] = combined_key(nodes, nbr, rows, stride, i);
}

__global__ void __launch_bounds__(kTileThreads)
radix_hist_kernel(const uint32_t* __restrict__ keys, int shift, int nblk, uint32_t* __restrict__ ghist) {
  pdl_sync();
  __shared__ uint32_t hist[256];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t* tile = keys + static_cast<int64_t>(blockIdx.x) * kTileKeys;
  for (int i = threadIdx.x; i < kTileKeys; i += kTileThreads) atomicAdd(&hist[(tile[i] >> shift) & 255u], 1u);
  __syncthreads();
  ghist[threadIdx.x * nblk + blockIdx.x] = hist[threadIdx.x];
}

// in-place exclusive scan of n uint32 by one CTA; total (optional) to *total_out
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(uint32_t* __restrict__ data, int n, int32_t* total_out) {
  pdl_sync();
  __shared__ uint32_t s_part[1024];
  const int tid = threadIdx.x;
  const int per = (n + 1023) / 1024;
  const int beg = tid * per, end = min(n, beg + per);
  uint32_t sum = 0;
  for (int i = beg; i < end; ++i) sum += data[i];
  s_part[tid] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    uint32_t add = tid >= o ? s_part[tid - o] : 0;
    __syncthreads();
    s_part[tid] += add;
    __syncthreads();
  }
  uint32_t run = s_part[tid] - sum;
  for (int i = beg; i < end; ++i) { uint32_t v = data[i]; data[i] = run; run += v; }
  if (total_out != nullptr && tid == 1023) *total_out = static_cast<int32_t>(s_part[1023]);
}

__global__ void __launch_bounds__(kTileThreads)
radix_scatter_kernel(const uint32_t* __restrict__ keys, int shift, int nblk, const uint32_t* __restrict__ ghist,
                     uint32_t* __restrict__ out) {
  pdl_sync();
  __shared__ uint32_t s_tile[kTileKeys];
  __shared__ uint32_t wcnt[kTileWarps * 256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t* tile = keys + static_cast<int64_t>(blockIdx.x) * kTileKeys;
  for (int i = tid; i < kTileKeys; i += kTileThreads) s_tile[i] = tile[i];
  for (int i = tid; i < kTileWarps * 256; i += kTileThreads) wcnt[i] = 0;
  __syncthreads();
  constexpr int chunk = kTileKeys / kTileWarps;
  warp_digit_count(s_tile, warp * chunk, chunk, shift, wcnt + warp * 256, lane);
  __syncthreads();
  {
    uint32_t run = 0;                       // tid is the digit (256 threads)
    for (int w = 0; w < kTileWarps; ++w) {
      const uint32_t c = wcnt[w * 256 + tid];
      wcnt[w * 256 + tid] = run;
      run += c;
    }
  }
  __syncthreads();
  const int b = blockIdx.x;
  warp_digit_scatter(s_tile, warp * chunk, chunk, shift, wcnt + warp * 256,
                     [&](uint32_t d) { return static_cast<int64_t>(__ldg(&ghist[d * nblk + b])); }, out, lane);
}

__device__ __forceinline__ bool is_head(const uint32_t* __restrict__ keys, int64_t i) {
  const uint32_t k = keys[i];
  return (k != kInvalidKey) && (i == 0 || keys[i - 1] != k);
}

__global__ void __launch_bounds__(kTileThreads)
rle_count_kernel(const uint32_t* __restrict__ keys, uint32_t* __restrict__ bcount) {
  pdl_sync();
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kTileKeys;
  int c = 0;
  for (int i = threadIdx.x; i < kTileKeys; i += kTileThreads) c += is_head(keys, base + i);
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) bcount[blockIdx.x] = s_cnt;
}

__global__ void __launch_bounds__(kTileThreads)
rle_write_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ boffset, int32_t* __restrict__ uniq) {
  pdl_sync();
  __shared__ int s_scan[kTileThreads];
  const int tid = threadIdx.x;
  constexpr int per = kTileKeys / kTileThreads;
  const int64_t beg = static_cast<int64_t>(blockIdx.x) * kTileKeys + tid * per;
  int heads = 0;
  for (int i = 0; i < per; ++i) heads += is_head(keys, beg + i);
  s_scan[tid] = heads;
  __syncthreads();
  for (int o = 1; o < kTileThreads; o <<= 1) {
    int add = tid >= o ? s_scan[tid - o] : 0;
    __syncthreads();
    s_scan[tid] += add;
    __syncthreads();
  }
  int64_t w_at = static_cast<int64_t>(boffset[blockIdx.x]) + s_scan[tid] - heads;
  for (int i = 0; i < per; ++i)
    if (is_head(keys, beg + i)) uniq[w_at++] = static_cast<int32_t>(keys[beg + i]);
}

__global__ void remap_kernel(const int32_t* __restrict__ nodes, const int32_t* __restrict__ num_rows_dev, int max_rows,
                             const int32_t* __restrict__ nbr, int stride, const int32_t* __restrict__ uniq,
                             const int32_t* __restrict__ num_uniq_dev, int32_t* __restrict__ nbr_idx,
                             int32_t* __restrict__ self_idx) {
  pdl_sync();
  const int rows = live_rows(num_rows_dev, max_rows);
  const int n_uniq = *num_uniq_dev;
  const uint32_t* u = reinterpret_cast<const uint32_t*>(uniq);
  const int64_t live = static_cast<int64_t>(rows) * stride, all = static_cast<int64_t>(max_rows) * stride;
  const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (nbr_idx != nullptr) {
    for (int64_t i = t0; i < all; i += step) {
      int32_t idx = -1;
      if (i < live) {
        const int32_t id = nbr[i];
        if (id >= 0) idx = lower_bound_u32(u, n_uniq, static_cast<uint32_t>(id));
      }
      nbr_idx[i] = idx;
    }
  }
  if (self_idx != nullptr) {
    for (int64_t i = t0; i < max_rows; i += step)
      self_idx[i] = i < rows ? lower_bound_u32(u, n_uniq, static_cast<uint32_t>(nodes[i])) : -1;
  }
}


// ---------------------------------------------------------------------------------------
// bitmap path: unique + remap in O(M + N/32) memory-parallel work, no sort.
//   mark  : bitmap[id>>5] |= 1<<(id&31) for every input id            (M atomics, spread)
//   scan  : per-word exclusive popcount prefix (4096-word blocks + one block of block sums)
//   emit  : every set bit writes its id at its rank  -> uniq ascending
//   remap : rank(id) = prefix[word] + popc(bits below id)
//   clear : the touched words are zeroed again, so the bitmap is all-zero between calls
// Output is identical to the radix path (ascending unique ids), bit for bit.
// ---------------------------------------------------------------------------------------
constexpr int kBmBlockWords = 4096;   // words per scan block: 1024 threads x uint4

__device__ __forceinline__ int32_t combined_id(const int32_t* __restrict__ nodes, const int32_t* __restrict__ nbr,
                                               int rows, int stride, int64_t i) {
  if (i < rows) return nodes[i];
  return nbr[i - rows];
}

__global__ void __launch_bounds__(256)
bitmap_mark_kernel(const int32_t* __restrict__ nodes, const int32_t* __restrict__ num_rows_dev, int max_rows,
                   const int32_t* __restrict__ nbr, int stride, int64_t num_nodes, uint32_t* __restrict__ bitmap,
                   int clear) {
  pdl_sync();
  const int rows = live_rows(num_rows_dev, max_rows);
  const int64_t m = static_cast<int64_t>(rows) * (stride + 1);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < m;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int32_t id = combined_id(nodes, nbr, rows, stride, i);
    if (id < 0 || id >= num_nodes) continue;
    if (clear) bitmap[id >> 5] = 0u;
    else atomicOr(&bitmap[id >> 5], 1u << (id & 31));
  }
}

// per-word exclusive popcount prefix inside 4096-word blocks + the block sums; the LAST block to finish turns the
// block sums into exclusive prefixes and writes the number of set bits (= |unique|) -- no separate scan launch
__global__ void __launch_bounds__(1024)
bitmap_scan_kernel(const uint32_t* __restrict__ bitmap, int32_t* __restrict__ wprefix, uint32_t* __restrict__ block_sum,
                   int nblk, unsigned int* __restrict__ ticket, int32_t* __restrict__ num_uniq) {
  pdl_sync();
  __shared__ int s_warp[32];
  __shared__ bool s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t w0 = static_cast<int64_t>(blockIdx.x) * kBmBlockWords + 4 * tid;
  const uint4 b = *reinterpret_cast<const uint4*>(bitmap + w0);
  const int c0 = __popc(b.x), c1 = __popc(b.y), c2 = __popc(b.z), c3 = __popc(b.w);
  const int mine = c0 + c1 + c2 + c3;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = s_warp[lane], wi = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    s_warp[lane] = wi - v;
    if (lane == 31) block_sum[blockIdx.x] = static_cast<uint32_t>(wi);
  }
  __syncthreads();
  const int base = s_warp[warp] + incl - mine;
  *reinterpret_cast<int4*>(wprefix + w0) = make_int4(base, base + c0, base + c0 + c1, base + c0 + c1 + c2);
  // ---- last block: exclusive scan of the block sums (nblk <= 2048), in chunks of 32 by one warp ----
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == static_cast<unsigned int>(nblk) - 1u;
  }
  __syncthreads();
  if (!s_last || warp != 0) return;
  __threadfence();
  unsigned int carry = 0;
  for (int b0 = 0; b0 < nblk; b0 += 32) {
    const int i = b0 + lane;
    const unsigned int v = i < nblk ? *reinterpret_cast<volatile const uint32_t*>(block_sum + i) : 0u;
    unsigned int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (i < nblk) block_sum[i] = carry + inc - v;
    carry += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) {
    *num_uniq = static_cast<int32_t>(carry);
    *ticket = 0u;
  }
}

__global__ void __launch_bounds__(256)
bitmap_emit_remap_kernel(const int32_t* __restrict__ nodes, const int32_t* __restrict__ num_rows_dev, int max_rows,
                         const int32_t* __restrict__ nbr, int stride, int64_t num_nodes, int64_t words,
                         int emit_blocks, const uint32_t* __restrict__ bitmap, const int32_t* __restrict__ wprefix,
                         const uint32_t* __restrict__ block_sum, int32_t* __restrict__ uniq,
                         int32_t* __restrict__ nbr_idx, int32_t* __restrict__ self_idx) {
  pdl_sync();
  if (static_cast<int>(blockIdx.x) < emit_blocks) {            // ---- emit: one thread per bitmap word
    const int64_t w = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (w >= words) return;
    uint32_t bits = bitmap[w];
    if (bits == 0u) return;
    int at = static_cast<int>(block_sum[w / kBmBlockWords]) + wprefix[w];
    while (bits) {
      const int b = __ffs(bits) - 1;
      uniq[at++] = static_cast<int32_t>(w * 32 + b);
      bits &= bits - 1;
    }
    return;
  }
  // ---- remap: one thread per input slot
  const int rows = live_rows(num_rows_dev, max_rows);
  const int64_t live = static_cast<int64_t>(rows) * stride, all = static_cast<int64_t>(max_rows) * stride;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x - emit_blocks) * blockDim.x + threadIdx.x;
  const int64_t step = static_cast<int64_t>(gridDim.x - emit_blocks) * blockDim.x;
  auto rank_of = [&](int32_t id) -> int32_t {
    const int64_t w = id >> 5;
    return static_cast<int32_t>(block_sum[w / kBmBlockWords]) + wprefix[w] +
           __popc(bitmap[w] & ((1u << (id & 31)) - 1u));
  };
  if (nbr_idx != nullptr) {
    for (int64_t i = t0; i < all; i += step) {
      int32_t idx = -1;
      if (i < live) {
        const int32_t id = nbr[i];
        if (id >= 0 && id < num_nodes) idx = rank_of(id);
      }
      nbr_idx[i] = idx;
    }
  }
  if (self_idx != nullptr) {
    for (int64_t i = t0; i < max_rows; i += step) {
      int32_t idx = -1;
      if (i < rows) {
        const int32_t id = nodes[i];
        if (id >= 0 && id < num_nodes) idx = rank_of(id);
      }
      self_idx[i] = idx;
    }
  }
}

struct BitmapPlan {
  int64_t words, words_pad;
  int nblk;
  size_t off_prefix, off_bsum, off_ticket, total;
};

static BitmapPlan make_bitmap_plan(int64_t num_nodes) {
  BitmapPlan p{};
  p.words = (num_nodes + 31) / 32;
  p.nblk = static_cast<int>((p.words + kBmBlockWords - 1) / kBmBlockWords);
  if (p.nblk < 1) p.nblk = 1;
  p.words_pad = static_cast<int64_t>(p.nblk) * kBmBlockWords;
  size_t off = static_cast<size_t>(p.words_pad) * 4;
  p.off_prefix = off; off += static_cast<size_t>(p.words_pad) * 4;
  p.off_bsum = off;   off += static_cast<size_t>(p.nblk + 1) * 4;
  off = (off + 15) & ~static_cast<size_t>(15);
  p.off_ticket = off; off += 16;
  p.total = (off + 15) & ~static_cast<size_t>(15);
  return p;
}

struct UniquePlan {
  bool small;
  int cap_keys;       // padded key capacity
  int nblk;           // tiles in the multi-CTA path
  size_t off_keys_b, off_hist, off_bcount, total;
};

static UniquePlan make_plan(int max_rows, int stride) {
  UniquePlan p{};
  const int64_t m = static_cast<int64_t>(max_rows) * (stride + 1);
  if (m <= kSmallMaxKeys) {
    p.small = true;
    p.cap_keys = static_cast<int>(((m + kSmallThreads - 1) / kSmallThreads) * kSmallThreads);
    p.total = 16;
    return p;
  }
  p.small = false;
  p.nblk = static_cast<int>((m + kTileKeys - 1) / kTileKeys);
  p.cap_keys = p.nblk * kTileKeys;
  size_t off = static_cast<size_t>(p.cap_keys) * 4;
  p.off_keys_b = off; off += static_cast<size_t>(p.cap_keys) * 4;
  p.off_hist = off;   off += static_cast<size_t>(256) * p.nblk * 4;
  p.off_bcount = off; off += static_cast<size_t>(p.nblk + 1) * 4;
  p.total = off;
  return p;
}

}  // namespace gs

using namespace gs;

extern "C" size_t gs_unique_workspace_bytes(int32_t max_rows, int32_t stride) {
  if (max_rows < 0 || stride < 0) return 0;
  return make_plan(max_rows, stride).total;
}

extern "C" int gs_unique_remap(const int32_t* nodes, const int32_t* num_rows_dev, int32_t max_rows,
                               const int32_t* nbr, int32_t stride, int32_t id_bits,
                               int32_t* uniq, int32_t* num_uniq_dev, int32_t* nbr_idx, int32_t* self_idx,
                               void* workspace, size_t workspace_bytes, gs_stream_t stream) {
  if (!nodes || !uniq || !num_uniq_dev || max_rows < 0 || stride < 0) return GS_ERR_BAD_ARG;
  if (stride > 0 && !nbr) return GS_ERR_BAD_ARG;
  if (id_bits < 1 || id_bits > 32) return GS_ERR_BAD_ARG;
  if (static_cast<int64_t>(max_rows) * (stride + 1) >= (1ll << 31)) return GS_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  if (max_rows == 0) {
    cudaError_t e = cudaMemsetAsync(num_uniq_dev, 0, sizeof(int32_t), st);
    return e == cudaSuccess ? GS_OK : static_cast<int>(e);
  }
  const int passes = (id_bits + 7) / 8;
  const UniquePlan p = make_plan(max_rows, stride);
  if (p.small) {
    const size_t smem = static_cast<size_t>(p.cap_keys) * 8 + kSmallWarps * 256 * sizeof(uint16_t);
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(unique_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kSmallMaxKeys * 8 + kSmallWarps * 256 * (int)sizeof(uint16_t));
      if (e != cudaSuccess) return static_cast<int>(e);
      attr_set = true;
    }
    launch(unique_small_kernel, 1, kSmallThreads, smem, st, nodes, num_rows_dev, max_rows, nbr, stride, passes,
                                                        p.cap_keys, uniq, num_uniq_dev, nbr_idx, self_idx);
    return finish_launch();
  }
  if (!workspace || workspace_bytes < p.total) return GS_ERR_WORKSPACE;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  uint32_t* keys_a = reinterpret_cast<uint32_t*>(ws);
  uint32_t* keys_b = reinterpret_cast<uint32_t*>(ws + p.off_keys_b);
  uint32_t* ghist = reinterpret_cast<uint32_t*>(ws + p.off_hist);
  uint32_t* bcount = reinterpret_cast<uint32_t*>(ws + p.off_bcount);
  int launches = 0;
  launch(gather_keys_kernel, std::min(p.nblk * (kTileKeys / 256), 148 * 8), 256, 0, st, nodes, num_rows_dev, max_rows, nbr,
                                                                                stride, p.cap_keys, keys_a);
  ++launches;
  uint32_t* in = keys_a;
  uint32_t* out = keys_b;
  for (int pass = 0; pass < passes; ++pass) {
    launch(radix_hist_kernel, p.nblk, kTileThreads, 0, st, in, 8 * pass, p.nblk, ghist);
    launch(exclusive_scan_kernel, 1, 1024, 0, st, ghist, 256 * p.nblk, nullptr);
    launch(radix_scatter_kernel, p.nblk, kTileThreads, 0, st, in, 8 * pass, p.nblk, ghist, out);
    launches += 3;
    uint32_t* t = in; in = out; out = t;
  }
  launch(rle_count_kernel, p.nblk, kTileThreads, 0, st, in, bcount);
  launch(exclusive_scan_kernel, 1, 1024, 0, st, bcount, p.nblk, num_uniq_dev);
  launch(rle_write_kernel, p.nblk, kTileThreads, 0, st, in, bcount, uniq);
  launches += 3;
  if (nbr_idx || self_idx) {
    const int64_t work = static_cast<int64_t>(max_rows) * (stride > 0 ? stride : 1);
    const int blocks = static_cast<int>(std::min<int64_t>((work + 255) / 256, 148 * 16));
    launch(remap_kernel, blocks, 256, 0, st, nodes, num_rows_dev, max_rows, nbr, stride, uniq, num_uniq_dev, nbr_idx,
                                         self_idx);
    ++launches;
  }
  return finish_launch(launches);
}

extern "C" size_t gs_unique_bitmap_workspace_bytes(int64_t num_nodes) {
  if (num_nodes < 1) return 0;
  return make_bitmap_plan(num_nodes).total;
}

extern "C" int gs_unique_remap_bitmap_ex(const int32_t* nodes, const int32_t* num_rows_dev, int32_t max_rows,
                                         const int32_t* nbr, int32_t stride, int64_t num_nodes,
                                         int32_t* uniq, int32_t* num_uniq_dev, int32_t* nbr_idx, int32_t* self_idx,
                                         void* workspace, size_t workspace_bytes, int32_t flags, gs_stream_t stream) {
  if (!nodes || !uniq || !num_uniq_dev || max_rows < 0 || stride < 0 || num_nodes < 1) return GS_ERR_BAD_ARG;
  if (stride > 0 && !nbr) return GS_ERR_BAD_ARG;
  if (num_nodes > (1ll << 31)) return GS_ERR_UNSUPPORTED;
  const BitmapPlan p = make_bitmap_plan(num_nodes);
  if (p.nblk > 2048) return GS_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < p.total || !aligned16(workspace)) return GS_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  if (max_rows == 0) {
    cudaError_t e = cudaMemsetAsync(num_uniq_dev, 0, sizeof(int32_t), st);
    return e == cudaSuccess ? GS_OK : static_cast<int>(e);
  }
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  uint32_t* bitmap = reinterpret_cast<uint32_t*>(ws);
  int32_t* wprefix = reinterpret_cast<int32_t*>(ws + p.off_prefix);
  uint32_t* bsum = reinterpret_cast<uint32_t*>(ws + p.off_bsum);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(ws + p.off_ticket);
  const int64_t m = static_cast<int64_t>(max_rows) * (stride + 1);
  const int id_blocks = static_cast<int>(std::min<int64_t>((m + 255) / 256, 148 * 8));
  int launches = 2;
  if (!(flags & GS_UNIQUE_MARKED)) {
    launch(bitmap_mark_kernel, id_blocks, 256, 0, st, nodes, num_rows_dev, max_rows, nbr, stride, num_nodes, bitmap, 0);
    ++launches;
  }
  launch(bitmap_scan_kernel, p.nblk, 1024, 0, st, bitmap, wprefix, bsum, p.nblk, ticket, num_uniq_dev);
  const int emit_blocks = static_cast<int>((p.words + 255) / 256);
  const int64_t slots = static_cast<int64_t>(max_rows) * (stride > 0 ? stride : 1);
  const int remap_blocks = static_cast<int>(std::min<int64_t>((slots + 255) / 256, 148 * 8));
  launch(bitmap_emit_remap_kernel, emit_blocks + remap_blocks, 256, 0, st, nodes, num_rows_dev, max_rows, nbr, stride,
                                                                      num_nodes, p.words, emit_blocks, bitmap, wprefix,
                                                                      bsum, uniq, nbr_idx, self_idx);
  if (!(flags & GS_UNIQUE_LEAVE_MARKS)) {
    launch(bitmap_mark_kernel, id_blocks, 256, 0, st, nodes, num_rows_dev, max_rows, nbr, stride, num_nodes, bitmap, 1);
    ++launches;
  }
  return finish_launch(launches);
}

extern "C" int gs_unique_remap_bitmap(const int32_t* nodes, const int32_t* num_rows_dev, int32_t max_rows,
                                      const int32_t* nbr, int32_t stride, int64_t num_nodes,
                                      int32_t* uniq, int32_t* num_uniq_dev, int32_t* nbr_idx, int32_t* self_idx,
                                      void* workspace, size_t workspace_bytes, gs_stream_t stream) {
  return gs_unique_remap_bitmap_ex(nodes, num_rows_dev, max_rows, nbr, stride, num_nodes, uniq, num_uniq_dev, nbr_idx,
                                   self_idx, workspace, workspace_bytes, 0, stream);
}
