#include <stdarg.h>

#include "gg_common.cuh"

namespace gg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace gg

#ifndef GG_BUILD_ID_STR
#define GG_BUILD_ID_STR "unstamped"
#endif
// "GG_BUILD_ID=<digest of csrc/ + gg_b200.h>": genomic_gnn_b200/build.py finds the marker in the file to decide whether
// the library on disk was built from the sources next to it (a stale .so must never be loaded against a newer header)
static const char g_build_id[] = "GG_BUILD_ID=" GG_BUILD_ID_STR;

extern "C" int gg_abi_version(void) { return GG_ABI_VERSION; }
extern "C" const char* gg_build_id(void) { return g_build_id + 12; }
extern "C" const char* gg_last_error(void) { return gg::g_err; }
