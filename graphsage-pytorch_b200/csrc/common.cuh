// Shared device/host helpers for the gsage_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gsage_b200.h"

namespace gs {

constexpr int kWarp = 32;
constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

extern int64_t g_launches;     // counted on the host at every kernel launch (gs_launch_count)

inline int finish_launch(int n = 1) {
  g_launches += n;
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? GS_OK : static_cast<int>(e);
}

inline cudaStream_t as_stream(gs_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch.  A step is ~20 short kernels in a dependency chain, so the
// launch latency + CTA ramp-up between two kernels is comparable to the kernels themselves.
// Every kernel starts with pdl_sync(): it lets the NEXT kernel of the stream be scheduled as
// soon as all CTAs of this one are resident (launch_dependents) and then waits until the
// PREVIOUS kernel has completed and flushed its writes (wait) before touching global memory.
// Both are no-ops for a launch without the attribute.  GS_PDL=0 disables the attribute.
bool pdl_enabled();
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
// The two halves on their own, for a kernel that reads some inputs BEFORE it waits.  Every kernel of the library lets
// its successor in at its own start, so a whole chain of launches can be resident and waiting: an early read is only
// safe for data whose producer is ordered before the chain by a FULL dependency (another stream's event, a host
// synchronisation) -- never for anything a kernel of the same stream wrote with programmatic launches in between.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Every kernel of the library asks for the maximum shared-memory carveout, whether it uses shared memory
// or not.  An SM can only hold CTAs that agree on its L1/shared split: a CTA that needs a different split
// waits until the SM has drained.  In the pipelined step the gathers of the preparation branch (no shared
// memory, default split) kept the classifier and GEMM CTAs of the training chain (100-200 KB) off every SM
// for the whole aggregation kernel -- 20-30 us per step (profiles/r1_timeline_*.txt).  The streaming kernels
// load with L1::no_allocate, so they lose nothing.
void prefer_max_smem(const void* kernel);      // api.cu; once per kernel function
// Exception: the HBM-bound gather kernels.  The L1 the max-shared split takes away is where their loads in
// flight land (measured: layer-1 aggregation 0.59 -> 0.52 of the HBM copy peak, 0.82 -> 0.69 at 88K rows), so
// they ask for it only when launched as background work of a two-branch step (gs_set_background).
void set_kernel_carveout(const void* kernel, bool max_shared);
bool background_launches();
bool early_reads();          // gs_set_early_reads: index lists / row counts may be read before the PDL wait

template <typename... KArgs, typename... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  prefer_max_smem(reinterpret_cast<const void*>(kernel));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);     // errors surface through finish_launch()
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ int live_rows(const int32_t* num_rows_dev, int max_rows) {
  if (num_rows_dev == nullptr) return max_rows;
  int n = __ldg(num_rows_dev);
  return n < max_rows ? n : max_rows;
}

// 128-bit read-only load that does not allocate in L1: gathered feature rows are touched
// once per kernel, L1 residency only evicts the index lists.
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// The same load under a predicate, yielding `fill` (a 32-bit pattern) in all four lanes when it is
// off.  Written as one asm block so the compiler sees no select between the load and its use: a
// batch of these issues back to back, which is the whole point of register staging.
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

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Philox4x32-10 (Salmon et al., SC'11).  ctr = 128-bit counter, key = 64-bit key.
__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifdef __CUDA_ARCH__
  uint32_t hi0 = __umulhi(M0, c[0]), hi1 = __umulhi(M1, c[2]);
#else
  uint32_t hi0 = static_cast<uint32_t>((static_cast<uint64_t>(M0) * c[0]) >> 32);
  uint32_t hi1 = static_cast<uint32_t>((static_cast<uint64_t>(M1) * c[2]) >> 32);
#endif
  uint32_t lo0 = M0 * c[0], lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint64_t key, uint32_t (&out)[4]) {
  uint32_t c[4] = {c0, c1, c2, c3};
  uint32_t k[2] = {static_cast<uint32_t>(key), static_cast<uint32_t>(key >> 32)};
#pragma unroll
  for (int i = 0; i < 10; ++i) philox_round(c, k);
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// Stream of 32-bit draws for one (row, purpose): block b of the stream is
// philox(ctr = {row, b, offset_lo, offset_hi}, key = seed).
struct PhiloxStream {
  uint32_t row, blk, off_lo, off_hi;
  uint64_t key;
  uint32_t buf[4];
  int have;
  __device__ PhiloxStream(uint64_t seed, uint64_t offset, uint32_t row_)
      : row(row_), blk(0), off_lo(static_cast<uint32_t>(offset)), off_hi(static_cast<uint32_t>(offset >> 32)),
        key(seed), have(0) {}
  __device__ __forceinline__ uint32_t next() {
    if (have == 0) {
      philox4x32_10(row, blk++, off_lo, off_hi, key, buf);
      have = 4;
    }
    return buf[4 - have--];
  }
  // uniform integer in [0, n)  (n >= 1), multiply-shift
  __device__ __forceinline__ uint32_t below(uint32_t n) { return __umulhi(next(), n); }
};

}  // namespace gs
