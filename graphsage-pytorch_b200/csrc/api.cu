// Library-level entry points: version, error strings, launch counter.
#include "common.cuh"

#include <stdlib.h>

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

extern "C" void gs_set_pdl(int32_t mode) { gs::g_pdl_override = mode < 0 ? -1 : (mode ? 1 : 0); }

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
