// K1 neighbour sampler, K5 random-walk positives and negative sampler.
// Integer / index work, HBM-latency bound; no tensor cores by design.
#include "common.cuh"

namespace gs {

// ---------------------------------------------------------------------------------------
// K1: one warp -- or half-warp, see the kernel -- per destination row, everything in registers.  Replaces
// src/models.py:279-285.
//   deg <  k : every neighbour (lane j reads col[beg+j])             (:282 else-branch)
//   deg >= k : k distinct uniform positions by Floyd's subset sampling: for i in 0..k-1,
//              j = deg-k+i, draw t in [0,j]; take t unless some earlier pick equals t, else
//              take j.  Every k-subset is equally likely, which is what random.sample gives.
//              Lane i owns pick i; the "already taken" test is one warp vote.
// The k (<= 32) ids are then ranked across lanes (rank = number of smaller valid ids), which
// sorts the row ascending without a scratch array, and the node's own id is dropped /
// inserted once according to self_mode (the `| {self}` of :285 and the `- {self}` of :298).
// ---------------------------------------------------------------------------------------
constexpr int kSampleWarps = 8;

// Optional work folded into a sampler launch (each saves a launch of the preparation chain, ~4 us apiece there):
//   queue / fetch_dst : the rows ARE the next batch of the device-side queue (gs_fetch_batch): row r's node is
//                       queue[(next % rows) * max_rows + r], copied to fetch_dst[r]; the last CTA advances `next`
//   mark              : K2's "mark" pass -- every id of the row (its node and the drawn neighbours) sets its bit
//   clear             : K2's "clear" pass of the PREVIOUS unique -- the rows of this launch are exactly the ids that
//                       unique emitted, so each row zeroes the bitmap word of its own node
//   prefetch          the rows the NEXT kernel of the chain will gather (layer 1: the drawn neighbours' and the node's
//                       own row of the feature table) are requested into L2 with one bulk prefetch each, so their DRAM
//                       fetch runs under the launch gap and the ramp of the aggregation kernel instead of inside it
struct SampleExtras {
  long long* queue;          // {address, rows, next, ticket}
  int32_t* fetch_dst;
  uint32_t* mark;
  uint32_t* clear;
  const char* prefetch_table;
  long long prefetch_ld_bytes;
  int prefetch_row_bytes;    // multiple of 16
};

// One prefetch.global.L2 per 128-byte line a row touches (LSU path: a few cycles each).  The bulk form
// (cp.async.bulk.prefetch.L2, one instruction per row) goes through the TMA unit at ~45 cycles per request and made the
// sampler 4 us slower at 120K rows (measured, profiles/r2_l2_prefetch.txt).
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, int bytes) {
#ifdef GS_PREFETCH_BULK
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
#else
  const char* c = static_cast<const char*>(p);
  const char* last = c + bytes - 1;
  for (const char* q = c; q < last; q += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
  asm volatile("prefetch.global.L2 [%0];" ::"l"(last));
#endif
}

// LANES = 32: one warp per row.  LANES = 16 (fan-out + self <= 16, the usual case): one HALF-warp per row, two rows per
// warp -- the kernel is a chain of three dependent memory round trips per row (node id -> rowptr -> col) with a few
// hundred cycles of warp votes in between, so what matters is that every row of a launch is resident at once: 11K rows
// are 1.2 waves of whole warps but 0.6 of a wave of half-warps.  A lane's pick index is its lane number inside the
// group, so the draws (Philox block (row, pick)) and the outputs are the same for both layouts.  All warp-level
// operations are executed by the whole warp (full mask); votes are cut down to the lane's group afterwards.
template <int LANES>
__global__ void __launch_bounds__(kSampleWarps * 32)
sample_neighbors_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t num_nodes,
                        const int32_t* __restrict__ nodes, const int32_t* __restrict__ num_rows_dev, int max_rows,
                        int k, int stride, int self_mode, uint64_t seed, uint64_t offset,
                        const int64_t* __restrict__ offset_dev,
                        int32_t* __restrict__ out_nbr, int32_t* __restrict__ out_cnt, const SampleExtras ex) {
  pdl_sync();
  constexpr int kRowsPerWarp = 32 / LANES;
  constexpr uint32_t kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int sl = lane % LANES;                                                  // lane inside the row's group = pick index
  const uint32_t gm = LANES == 32 ? kFull : (0xffffu << (lane & 16));           // the lanes of my group
  const int r = (blockIdx.x * kSampleWarps + (threadIdx.x >> 5)) * kRowsPerWarp + lane / LANES;
  const int rows = live_rows(num_rows_dev, max_rows);
  const int32_t* src = nodes;
  if (ex.queue != nullptr) {                      // fused gs_fetch_batch
    const int32_t* base = reinterpret_cast<const int32_t*>(ex.queue[0]);
    const long long q_rows = ex.queue[1], next = ex.queue[2];
    src = (base != nullptr && q_rows > 0) ? base + (next % q_rows) * max_rows : nullptr;
  }
  constexpr int32_t kNone = 0x7fffffff;
  const bool live = r < rows && r < max_rows;
  const int32_t me = (live && src != nullptr) ? __ldg(src + r) : -1;
  const bool ok = live && me >= 0 && me < num_nodes;
  if (live && ex.fetch_dst != nullptr && sl == 0) ex.fetch_dst[r] = me;
  int64_t beg = 0;
  uint32_t deg = 0;
  if (ok) {
    if (sl == 0) {
      if (ex.clear != nullptr) ex.clear[me >> 5] = 0u;
      if (ex.mark != nullptr) atomicOr(ex.mark + (me >> 5), 1u << (me & 31));
    }
    beg = __ldg(rowptr + me);
    deg = static_cast<uint32_t>(__ldg(rowptr + me + 1) - beg);
  }
  const bool floyd = ok && deg >= static_cast<uint32_t>(k);
  int32_t val = kNone;
  if (ok && !floyd && static_cast<uint32_t>(sl) < deg) val = __ldg(col + beg + sl);
  if (__any_sync(kFull, floyd)) {
    if (offset_dev != nullptr) offset += static_cast<uint64_t>(__ldg(offset_dev)) << 8;
    // Floyd's subset sampling.  Draw i comes from Philox block (row, i): lane i of the group computes its own (the k
    // draws in parallel instead of every lane running the same serial stream); only the "already taken?" resolution is
    // sequential, one shuffle + one vote per pick.  (Groups whose row takes every neighbour run along idly.)
    const uint32_t j_mine = deg - k + sl;                     // pick i ranges over [0, deg - k + i]
    uint32_t t_mine = 0;
    if (floyd && sl < k) {
      uint32_t draw[4];
      philox4x32_10(static_cast<uint32_t>(r), static_cast<uint32_t>(sl), static_cast<uint32_t>(offset),
                    static_cast<uint32_t>(offset >> 32), seed, draw);
      t_mine = __umulhi(draw[0], j_mine + 1);                 // uniform in [0, j]
    }
    uint32_t mypos = 0xffffffffu;
    for (int i = 0; i < k; ++i) {
      const uint32_t t = __shfl_sync(kFull, t_mine, i, LANES);
      const bool taken = (__ballot_sync(kFull, mypos == t) & gm) != 0u;
      if (sl == i) mypos = taken ? j_mine : t;
    }
    if (floyd && sl < k) val = __ldg(col + beg + mypos);
  }
  bool valid = val != kNone;
  if (self_mode != GS_SELF_KEEP && val == me) valid = false;
  if (self_mode == GS_SELF_ONCE && sl == k && ok) { val = me; valid = true; }   // k < LANES in this mode
  // The reference's rows are sets (src/dataCenter.py:33).  A CSR row that repeats an id (graphs generated on the
  // device) must behave the same way: a repeated id keeps only its lowest lane -- otherwise two lanes would share
  // a rank and leave a slot unwritten.  Invalid lanes get keys of their own, so they match nobody.
  const uint32_t same = __match_any_sync(kFull, valid ? static_cast<uint32_t>(val) : (0x80000000u | lane)) & gm;
  valid = valid && (same & ((1u << lane) - 1u)) == 0u;
  const int32_t key = valid ? val : kNone;
  int rank = 0;
  const int scan = k + (self_mode == GS_SELF_ONCE ? 1 : 0);
  for (int o = 0; o < scan; ++o) rank += (__shfl_sync(kFull, key, o, LANES) < key) ? 1 : 0;
  const int m = __popc(__ballot_sync(kFull, valid) & gm);
  if (r < max_rows) {
    int32_t* dst = out_nbr + static_cast<int64_t>(r) * stride;
    if (!live) {               // keep the padding region well defined for the consumers
      for (int j = sl; j < stride; j += LANES) dst[j] = -1;
      if (sl == 0) out_cnt[r] = 0;
    } else {
      if (valid) {
        dst[rank] = val;
        if (ex.mark != nullptr) atomicOr(ex.mark + (val >> 5), 1u << (val & 31));
        if (ex.prefetch_table != nullptr) prefetch_l2_bulk(ex.prefetch_table + val * ex.prefetch_ld_bytes, ex.prefetch_row_bytes);
      }
      if (ex.prefetch_table != nullptr && sl == LANES - 1 && ok)
        prefetch_l2_bulk(ex.prefetch_table + me * ex.prefetch_ld_bytes, ex.prefetch_row_bytes);
      for (int j = m + sl; j < stride; j += LANES) dst[j] = -1;
      if (sl == 0) out_cnt[r] = m;
    }
  }
  if (ex.queue != nullptr) {                      // the last CTA to get here advances the queue cursor
    __syncthreads();
    if (threadIdx.x == 0) {
      const long long next = ex.queue[2];
      __threadfence();
      const unsigned long long done = atomicAdd(reinterpret_cast<unsigned long long*>(ex.queue + 3), 1ULL);
      if (done == gridDim.x - 1) {
        ex.queue[2] = next + 1;
        ex.queue[3] = 0;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// K5a: positives.  src/models.py:169-186.  One thread per (seed, walk).
// ---------------------------------------------------------------------------------------
__global__ void random_walk_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                   int64_t num_nodes, const int32_t* __restrict__ seeds, int num_seeds, int n_walks,
                                   int walk_len, const uint8_t* __restrict__ is_train, uint64_t seed, uint64_t offset,
                                   const int64_t* __restrict__ offset_dev, int32_t* __restrict__ pos) {
  pdl_sync();
  if (offset_dev != nullptr) offset += static_cast<uint64_t>(__ldg(offset_dev)) << 8;     // step counter of a captured loop
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_seeds * n_walks) return;
  const int s = t / n_walks;
  const int32_t start = seeds[s];
  int32_t* dst = pos + static_cast<int64_t>(t) * walk_len;
  PhiloxStream rng(seed, offset, static_cast<uint32_t>(t));
  int32_t cur = start;
  bool dead = !(start >= 0 && start < num_nodes) || (rowptr[start + 1] == rowptr[start]);   // :171-172
  for (int step = 0; step < walk_len; ++step) {
    int32_t out = -1;
    if (!dead) {
      const int64_t beg = rowptr[cur];
      const uint32_t deg = static_cast<uint32_t>(rowptr[cur + 1] - beg);
      if (deg == 0) {
        dead = true;
      } else {
        const int32_t nxt = col[beg + rng.below(deg)];                                     // :178
        if (nxt != start && is_train[nxt]) out = nxt;                                      // :180
        cur = nxt;
      }
    }
    dst[step] = out;
  }
}

// ---------------------------------------------------------------------------------------
// K5b: negatives.  src/models.py:153-167.  One CTA per seed.
//   1. level-synchronous BFS for `hops` levels over the full adjacency, visited set kept as
//      a bitmap of num_nodes bits, frontiers as two id queues (all in `workspace`).
//   2. far = train nodes whose bit is clear; |far| counted by the CTA.
//   3. num_neg distinct ranks in [0,|far|) by Floyd's algorithm (thread 0), then the CTA
//      walks the train list again and emits the nodes holding those ranks.
// ---------------------------------------------------------------------------------------
constexpr int kNegThreads = 256;

__global__ void __launch_bounds__(kNegThreads)
negative_sample_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t num_nodes,
                       const int32_t* __restrict__ seeds, int hops, int num_neg,
                       const int32_t* __restrict__ train_nodes, int num_train, uint64_t seed, uint64_t offset,
                       const int64_t* __restrict__ offset_dev, int32_t* __restrict__ neg, int32_t* __restrict__ neg_cnt,
                       uint32_t* __restrict__ workspace, const uint8_t* __restrict__ is_train) {
  pdl_sync();
  if (offset_dev != nullptr) offset += static_cast<uint64_t>(__ldg(offset_dev)) << 8;
  const int s = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kNegThreads / 32;
  const int64_t words = (num_nodes + 31) / 32;
  const int64_t per_seed = words + 2 * num_nodes;
  uint32_t* bitmap = workspace + s * per_seed;
  int32_t* queue_a = reinterpret_cast<int32_t*>(bitmap + words);
  int32_t* queue_b = queue_a + num_nodes;
  __shared__ int s_cur, s_next, s_far, s_ball_train;
  __shared__ uint32_t s_rank[GS_MAX_FANOUT * 4];
  __shared__ int s_scan[kNegThreads];

  for (int64_t w = tid; w < words; w += kNegThreads) bitmap[w] = 0u;
  const int32_t me = seeds[s];
  if (tid == 0) { s_cur = 0; s_next = 0; s_ball_train = 0; }
  __syncthreads();
  if (tid == 0 && me >= 0 && me < num_nodes) {
    bitmap[me >> 5] |= 1u << (me & 31);
    queue_a[0] = me;
    s_cur = 1;
    if (is_train != nullptr && is_train[me]) s_ball_train = 1;
  }
  __syncthreads();
  int32_t* cur_q = queue_a;
  int32_t* next_q = queue_b;
  for (int hop = 0; hop < hops; ++hop) {
    const int n_cur = s_cur;
    if (n_cur == 0) break;
    for (int i = warp; i < n_cur; i += nwarps) {
      const int32_t v = cur_q[i];
      const int64_t beg = rowptr[v], end = rowptr[v + 1];
      for (int64_t e = beg + lane; e < end; e += 32) {
        const int32_t u = col[e];
        const uint32_t bit = 1u << (u & 31);
        const uint32_t old = atomicOr(&bitmap[u >> 5], bit);
        if (!(old & bit)) {
          next_q[atomicAdd(&s_next, 1)] = u;
          if (is_train != nullptr && is_train[u]) atomicAdd(&s_ball_train, 1);     // train nodes inside the ball
        }
      }
    }
    __syncthreads();
    if (tid == 0) { s_cur = s_next; s_next = 0; }
    int32_t* t = cur_q; cur_q = next_q; next_q = t;
    __syncthreads();
  }
  __threadfence_block();
  // ---- fast path (is_train given, plenty of far nodes): rejection sampling.  |far| = |train| - |train in ball| is known
  //      from the marking pass; 32 candidates per round are drawn uniformly from the train list (one Philox block per
  //      lane), those inside the ball or already taken are rejected, the rest are accepted in lane order -- the accepted
  //      sequence is that of sequential rejection sampling, i.e. a uniform num_neg-subset of the far set (:164), without
  //      walking the train list (twice) per seed as the exact path below does. ----
  if (is_train != nullptr && num_train - s_ball_train >= 4 * num_neg) {
    int32_t* dst = neg + static_cast<int64_t>(s) * num_neg;
    if (warp == 0) {
      int32_t* acc = reinterpret_cast<int32_t*>(s_rank);
      int have = 0;
      for (uint32_t round = 0; have < num_neg && round < 4096u; ++round) {
        uint32_t draw[4];
        philox4x32_10(static_cast<uint32_t>(s), round * 32u + lane, static_cast<uint32_t>(offset),
                      static_cast<uint32_t>(offset >> 32), seed, draw);
        const int32_t v = train_nodes[__umulhi(draw[0], static_cast<uint32_t>(num_train))];
        bool ok = !((bitmap[v >> 5] >> (v & 31)) & 1u);
        for (int j = 0; j < have; ++j) ok = ok && acc[j] != v;
        const uint32_t same = __match_any_sync(0xffffffffu, ok ? static_cast<uint32_t>(v) : (0x80000000u | lane));
        ok = ok && (same & ((1u << lane) - 1u)) == 0u;
        const uint32_t m = __ballot_sync(0xffffffffu, ok);
        const int pos = have + __popc(m & ((1u << lane) - 1u));
        if (ok && pos < num_neg) { acc[pos] = v; dst[pos] = v; }
        have = min(num_neg, have + __popc(m));
        __syncwarp();
      }
      for (int j = have + lane; j < num_neg; j += 32) dst[j] = -1;           // (only if the round cap was hit)
      if (lane == 0) neg_cnt[s] = have;
    }
    return;
  }
  // count far train nodes
  int mine = 0;
  for (int i = tid; i < num_train; i += kNegThreads) {
    const int32_t v = train_nodes[i];
    mine += !((bitmap[v >> 5] >> (v & 31)) & 1u);
  }
  if (tid == 0) s_far = 0;
  __syncthreads();
  mine = __reduce_add_sync(0xffffffffu, mine);
  if (lane == 0) atomicAdd(&s_far, mine);
  __syncthreads();
  const int far = s_far;
  const int take = far < num_neg ? far : num_neg;     // :164  (all of them when too few)
  if (tid == 0) {
    if (far > num_neg) {
      PhiloxStream rng(seed, offset, static_cast<uint32_t>(s));
      int m = 0;
      for (uint32_t j = far - num_neg; j < static_cast<uint32_t>(far); ++j) {
        uint32_t t = rng.below(j + 1);
        bool taken = false;
        for (int q = 0; q < m; ++q) taken |= (s_rank[q] == t);
        s_rank[m++] = taken ? j : t;
      }
      // ascending ranks so the output order is deterministic
      for (int a = 1; a < m; ++a) {
        uint32_t v = s_rank[a];
        int b = a - 1;
        while (b >= 0 && s_rank[b] > v) { s_rank[b + 1] = s_rank[b]; --b; }
        s_rank[b + 1] = v;
      }
    }
    neg_cnt[s] = take;
  }
  __syncthreads();
  // second walk over the train list in chunks of kNegThreads with a block scan of the
  // "is far" flags to recover each far node's rank.
  int32_t* dst = neg + static_cast<int64_t>(s) * num_neg;
  for (int i = tid; i < num_neg; i += kNegThreads) dst[i] = -1;
  __syncthreads();
  int base = 0;
  for (int chunk = 0; chunk < num_train; chunk += kNegThreads) {
    const int i = chunk + tid;
    int32_t v = -1;
    int flag = 0;
    if (i < num_train) {
      v = train_nodes[i];
      flag = !((bitmap[v >> 5] >> (v & 31)) & 1u);
    }
    s_scan[tid] = flag;
    __syncthreads();
    for (int o = 1; o < kNegThreads; o <<= 1) {          // Hillis-Steele inclusive scan
      int add = tid >= o ? s_scan[tid - o] : 0;
      __syncthreads();
      s_scan[tid] += add;
      __syncthreads();
    }
    const int rank = base + s_scan[tid] - flag;
    if (flag) {
      if (far <= num_neg) {
        dst[rank] = v;
      } else {
        int lo = 0, hi = num_neg;                        // binary search in the sorted rank list
        while (lo < hi) { int mid = (lo + hi) >> 1; if (s_rank[mid] < static_cast<uint32_t>(rank)) lo = mid + 1; else hi = mid; }
        if (lo < num_neg && s_rank[lo] == static_cast<uint32_t>(rank)) dst[lo] = v;
      }
    }
    base += s_scan[kNegThreads - 1];
    __syncthreads();
  }
}

// Batch queue of the device-resident train loop (src/utils.py:141-145: the batches of an epoch are
// slices of one shuffled id array, all known up front).  desc = {address of an int32 [rows x b_sz]
// array, rows, next}: copies row `next % rows` into dst and advances `next`, inside the step's
// graph, so consecutive graph replays need no host-side copy between them.
__global__ void __launch_bounds__(256)
fetch_batch_kernel(long long* __restrict__ desc, int b_sz, int32_t* __restrict__ dst) {
  pdl_sync();
  const int32_t* base = reinterpret_cast<const int32_t*>(desc[0]);
  const long long rows = desc[1], next = desc[2];
  if (base == nullptr || rows <= 0) return;
  const int32_t* src = base + (next % rows) * b_sz;
  for (int i = threadIdx.x; i < b_sz; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
  if (threadIdx.x == 0) desc[2] = next + 1;
}

}  // namespace gs

using namespace gs;

extern "C" int gs_fetch_batch(int64_t* queue_desc, int32_t b_sz, int32_t* dst, gs_stream_t stream) {
  if (!queue_desc || !dst || b_sz < 1) return GS_ERR_BAD_ARG;
  launch(fetch_batch_kernel, 1, 256, 0, as_stream(stream), reinterpret_cast<long long*>(queue_desc), b_sz, dst);
  return finish_launch();
}

extern "C" int gs_sample_neighbors_ex(const int64_t* rowptr, const int32_t* col, int64_t num_nodes,
                                      const int32_t* nodes, const int32_t* num_rows_dev, int32_t max_rows,
                                      int32_t k, int32_t stride, int32_t self_mode, uint64_t seed, uint64_t offset,
                                      const int64_t* offset_dev, int32_t* out_nbr, int32_t* out_cnt,
                                      int64_t* queue_desc, int32_t* fetch_dst, uint32_t* mark_bitmap,
                                      uint32_t* clear_bitmap, const void* prefetch_table, int64_t prefetch_ld_bytes,
                                      int32_t prefetch_row_bytes, gs_stream_t stream) {
  if (!rowptr || !col || !out_nbr || !out_cnt) return GS_ERR_BAD_ARG;
  if (!nodes && !queue_desc) return GS_ERR_BAD_ARG;
  if (queue_desc && num_rows_dev) return GS_ERR_BAD_ARG;          // a queued batch has exactly max_rows seeds
  if (mark_bitmap && clear_bitmap) return GS_ERR_BAD_ARG;         // marking and clearing words in one grid would race
  if (k < 1 || k > GS_MAX_FANOUT || max_rows < 0) return GS_ERR_BAD_ARG;
  if (self_mode == GS_SELF_ONCE && k > GS_MAX_FANOUT - 1) return GS_ERR_UNSUPPORTED;
  if (stride < k + (self_mode == GS_SELF_ONCE ? 1 : 0)) return GS_ERR_BAD_ARG;
  if (self_mode < GS_SELF_KEEP || self_mode > GS_SELF_ONCE) return GS_ERR_BAD_ARG;
  if (max_rows == 0) return GS_OK;
  if (prefetch_table && ((prefetch_row_bytes & 15) || prefetch_row_bytes < 16 || (prefetch_ld_bytes & 15) ||
                         (reinterpret_cast<uintptr_t>(prefetch_table) & 15)))
    return GS_ERR_ALIGNMENT;
  SampleExtras ex{reinterpret_cast<long long*>(queue_desc), fetch_dst, mark_bitmap, clear_bitmap,
                  static_cast<const char*>(prefetch_table), prefetch_ld_bytes, prefetch_row_bytes};
  if (k + (self_mode == GS_SELF_ONCE ? 1 : 0) <= 16) {            // two rows per warp
    const int per_cta = kSampleWarps * 2;
    launch(sample_neighbors_kernel<16>, (max_rows + per_cta - 1) / per_cta, kSampleWarps * 32, 0, as_stream(stream),
           rowptr, col, num_nodes, nodes, num_rows_dev, max_rows, k, stride, self_mode, seed, offset, offset_dev, out_nbr,
           out_cnt, ex);
  } else {
    launch(sample_neighbors_kernel<32>, (max_rows + kSampleWarps - 1) / kSampleWarps, kSampleWarps * 32, 0,
           as_stream(stream), rowptr, col, num_nodes, nodes, num_rows_dev, max_rows, k, stride, self_mode, seed, offset,
           offset_dev, out_nbr, out_cnt, ex);
  }
  return finish_launch();
}

extern "C" int gs_sample_neighbors(const int64_t* rowptr, const int32_t* col, int64_t num_nodes,
                                   const int32_t* nodes, const int32_t* num_rows_dev, int32_t max_rows,
                                   int32_t k, int32_t stride, int32_t self_mode, uint64_t seed, uint64_t offset,
                                   const int64_t* offset_dev, int32_t* out_nbr, int32_t* out_cnt, gs_stream_t stream) {
  if (!nodes) return GS_ERR_BAD_ARG;
  return gs_sample_neighbors_ex(rowptr, col, num_nodes, nodes, num_rows_dev, max_rows, k, stride, self_mode, seed, offset,
                                offset_dev, out_nbr, out_cnt, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, stream);
}

extern "C" int gs_random_walk_pos(const int64_t* rowptr, const int32_t* col, int64_t num_nodes,
                                  const int32_t* seeds, int32_t num_seeds, int32_t n_walks, int32_t walk_len,
                                  const uint8_t* is_train, uint64_t seed, uint64_t offset, const int64_t* offset_dev,
                                  int32_t* pos, gs_stream_t stream) {
  if (!rowptr || !col || !seeds || !is_train || !pos) return GS_ERR_BAD_ARG;
  if (num_seeds < 0 || n_walks < 1 || walk_len < 1) return GS_ERR_BAD_ARG;
  if (num_seeds == 0) return GS_OK;
  const int total = num_seeds * n_walks, threads = 128;
  launch(random_walk_kernel, (total + threads - 1) / threads, threads, 0, as_stream(stream), 
      rowptr, col, num_nodes, seeds, num_seeds, n_walks, walk_len, is_train, seed, offset, offset_dev, pos);
  return finish_launch();
}

extern "C" size_t gs_negative_workspace_bytes(int64_t num_nodes, int32_t num_seeds) {
  const int64_t words = (num_nodes + 31) / 32;
  return static_cast<size_t>(num_seeds) * static_cast<size_t>(words + 2 * num_nodes) * sizeof(uint32_t);
}

extern "C" int gs_negative_sample_ex(const int64_t* rowptr, const int32_t* col, int64_t num_nodes,
                                     const int32_t* seeds, int32_t num_seeds, int32_t hops, int32_t num_neg,
                                     const int32_t* train_nodes, int32_t num_train, const uint8_t* is_train,
                                     uint64_t seed, uint64_t offset, const int64_t* offset_dev, int32_t* neg,
                                     int32_t* neg_cnt, void* workspace, size_t workspace_bytes, gs_stream_t stream) {
  if (!rowptr || !col || !seeds || !train_nodes || !neg || !neg_cnt || !workspace) return GS_ERR_BAD_ARG;
  if (num_neg < 1 || num_neg > GS_MAX_FANOUT * 4 || hops < 0 || num_seeds < 0 || num_train < 0) return GS_ERR_BAD_ARG;
  if (workspace_bytes < gs_negative_workspace_bytes(num_nodes, num_seeds)) return GS_ERR_WORKSPACE;
  if (num_seeds == 0) return GS_OK;
  launch(negative_sample_kernel, num_seeds, kNegThreads, 0, as_stream(stream),
      rowptr, col, num_nodes, seeds, hops, num_neg, train_nodes, num_train, seed, offset, offset_dev, neg, neg_cnt,
      static_cast<uint32_t*>(workspace), is_train);
  return finish_launch();
}

extern "C" int gs_negative_sample(const int64_t* rowptr, const int32_t* col, int64_t num_nodes,
                                  const int32_t* seeds, int32_t num_seeds, int32_t hops, int32_t num_neg,
                                  const int32_t* train_nodes, int32_t num_train, uint64_t seed, uint64_t offset,
                                  const int64_t* offset_dev, int32_t* neg, int32_t* neg_cnt, void* workspace,
                                  size_t workspace_bytes, gs_stream_t stream) {
  return gs_negative_sample_ex(rowptr, col, num_nodes, seeds, num_seeds, hops, num_neg, train_nodes, num_train, nullptr,
                               seed, offset, offset_dev, neg, neg_cnt, workspace, workspace_bytes, stream);
}
