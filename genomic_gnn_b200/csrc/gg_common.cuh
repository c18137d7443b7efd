// Shared helpers for the sm_100a kernels behind include/gg_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gg_b200.h"

namespace gg {

void set_error(const char* fmt, ...);
int sm_count();  // SMs of the current device (cached per device)

inline cudaStream_t as_stream(gg_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define GG_CUDA_CHECK(expr)                                                              \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      ::gg::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
      return GG_ERR_CUDA;                                                                \
    }                                                                                    \
  } while (0)

#define GG_REQUIRE(cond, code, msg)        \
  do {                                     \
    if (!(cond)) {                         \
      ::gg::set_error("%s (%s)", msg, #cond); \
      return code;                         \
    }                                      \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// carve typed sub-buffers out of a caller-owned workspace
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* p = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
  size_t used() const { return align_up(off, 256); }
};

}  // namespace gg
