// Diagnostics: what a TMA tile::gather4 load needs from its tensor map (box rows 1 or 4?) and what it writes.
// usage: gather4_probe <box_rows>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap map, float* out, int* status, int r0, int r1, int r2, int r3, int c0) {
  __shared__ __align__(1024) float tile[8 * 32];
  __shared__ __align__(8) uint64_t bar;
  for (int i = threadIdx.x; i < 8 * 32; i += blockDim.x) tile[i] = -1.f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(4 * 128) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(&bar)) : "memory");
    int ok = 0;
    for (int spin = 0; spin < 2000000 && !ok; ++spin) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    *status = ok;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 8 * 32; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char** argv) {
  const int box_rows = argc > 1 ? atoi(argv[1]) : 1;
  const int R = 64, C = 100;
  float* h = (float*)malloc(R * C * 4);
  for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[r * C + c] = r * 1000.f + c;
  float *d, *out; int* st;
  cudaMalloc(&d, R * C * 4); cudaMalloc(&out, 8 * 32 * 4); cudaMalloc(&st, 4);
  cudaMemcpy(d, h, R * C * 4, cudaMemcpyHostToDevice); cudaMemset(st, 0, 4);
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn fn = (EncodeTiledFn)fnp;
  CUtensorMap map; memset(&map, 0, sizeof(map));
  const cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)R}; const cuuint64_t gstride[1] = {(cuuint64_t)C * 4};
  const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows}; const cuuint32_t estr[2] = {1u, 1u};
  CUresult e = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("box_rows %d: encode -> %d\n", box_rows, (int)e);
  if (e != CUDA_SUCCESS) return 0;
  for (int c0 = 0; c0 <= 96; c0 += 96) {
    probe<<<1, 128>>>(map, out, st, 5, 17, 3, 60, c0);
    cudaError_t ce = cudaDeviceSynchronize();
    int hs = -1; float ho[8 * 32];
    cudaMemcpy(&hs, st, 4, cudaMemcpyDeviceToHost); cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
    printf("c0 %d: sync %s, barrier completed %d\n", c0, cudaGetErrorString(ce), hs);
    for (int row = 0; row < 5; ++row) {   // un-swizzle: 16-byte piece j of row `row` sits at piece j ^ (row & 7)
      printf("  smem row %d:", row);
      for (int j = 0; j < 8; ++j) printf(" %.0f", ho[row * 32 + ((j ^ (row & 7)) << 2)]);
      printf("  | last of piece 0: %.0f\n", ho[row * 32 + ((0 ^ (row & 7)) << 2) + 3]);
    }
    if (ce != cudaSuccess) break;
  }
  return 0;
}
