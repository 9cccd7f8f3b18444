// Library-level entry points: version, error strings, launch counter.
#include "common.cuh"

#include <stdlib.h>

#include <mutex>

namespace gs {
int64_t g_launches = 0;
static int g_pdl_override = -1;          // gs_set_pdl: -1 = follow GS_PDL (default on), 0 = off, 1 = on
bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("GS_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl_override < 0 ? on == 1 : g_pdl_override == 1;
}
}

namespace gs {
static std::mutex g_carve_mu;
static const void* g_carve_seen[256];
static int g_carve_n = 0;
static int g_background = 0;            // gs_set_background
static int g_early_reads = 0;           // gs_set_early_reads

static bool carveout_enabled() {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("GS_MAX_SMEM_CARVEOUT");       // 0: leave the driver's default split (A/B measurements)
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  return enabled == 1;
}
static bool carve_seen_or_add(const void* kernel) {       // caller holds g_carve_mu
  for (int i = 0; i < g_carve_n; ++i)
    if (g_carve_seen[i] == kernel) return true;
  if (g_carve_n < 256) g_carve_seen[g_carve_n++] = kernel;
  return false;
}
void prefer_max_smem(const void* kernel) {
  std::lock_guard<std::mutex> lock(g_carve_mu);
  if (!carveout_enabled() || carve_seen_or_add(kernel)) return;
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}
void set_kernel_carveout(const void* kernel, bool max_shared) {
  std::lock_guard<std::mutex> lock(g_carve_mu);
  if (!carveout_enabled()) return;
  carve_seen_or_add(kernel);             // launch() leaves this kernel's split alone from now on
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                       max_shared ? cudaSharedmemCarveoutMaxShared : cudaSharedmemCarveoutDefault);
}
bool background_launches() { return g_background != 0; }
bool early_reads() { return g_early_reads != 0; }
}

extern "C" void gs_set_background(int32_t on) { gs::g_background = on ? 1 : 0; }

extern "C" void gs_set_early_reads(int32_t on) { gs::g_early_reads = on ? 1 : 0; }

extern "C" void gs_set_pdl(int32_t mode) { gs::g_pdl_override = mode < 0 ? -1 : (mode ? 1 : 0); }

// Timeline markers (diagnostics): one thread writes the GPU's global nanosecond timer.  A trainer built
// with native.timeline_begin() puts one behind every launch of a step; read back they give per-kernel
// completion times on each branch of the step graph (profiles/*timeline*), which ncu's serialised
// replay cannot.  Markers are ordinary stream work: they cost ~1-2 us each and disable PDL overlap.
namespace gs {
__global__ void stamp_kernel(unsigned long long* slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *slot = t;
}
}
extern "C" int gs_debug_stamp(uint64_t* slot, gs_stream_t stream) {
  if (!slot) return GS_ERR_BAD_ARG;
  gs::stamp_kernel<<<1, 1, 0, gs::as_stream(stream)>>>(reinterpret_cast<unsigned long long*>(slot));
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? GS_OK : static_cast<int>(e);
}

extern "C" int gs_version(void) { return GS_ABI_VERSION; }

extern "C" int64_t gs_launch_count(void) { return gs::g_launches; }
extern "C" void gs_launch_count_reset(void) { gs::g_launches = 0; }

extern "C" const char* gs_error_string(int code) {
  switch (code) {
    case GS_OK: return "ok";
    case GS_ERR_BAD_ARG: return "gsage_b200: bad argument (null pointer, negative size or inconsistent shape)";
    case GS_ERR_UNSUPPORTED: return "gsage_b200: unsupported configuration";
    case GS_ERR_WORKSPACE: return "gsage_b200: workspace missing or too small";
    case GS_ERR_ALIGNMENT: return "gsage_b200: pointer not 16-byte aligned or leading dimension not a multiple of 4";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "gsage_b200: unknown error";
}
