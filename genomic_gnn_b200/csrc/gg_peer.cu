// One-shot exchanges over NVLink peer memory for the multi-GPU graph build (C1 count pieces,
// C2 packed KF edges, candidate counts).  Every rank owns one buffer of a symmetric allocation (same size, mapped into
// every rank's address space: torch.distributed._symmetric_memory hands out the peer pointers) plus a pad of 32-bit
// flags.  An exchange is
//     gg_peer_push    my segment -> the same offset in every peer's buffer (16-byte coalesced stores over NVLink),
//     gg_peer_signal  flag[channel][my rank] = sequence number in every peer's pad (after a system-scope fence),
//     gg_peer_wait    spin until flag[channel][r] has reached the sequence number for every r,
// three launches on the caller's stream and no host involvement -- against ~100 us of launch and protocol latency per
// NCCL collective, which is what a 2 ms build on 8 GPUs was spending most of its exchange time on.  The reference has
// no counterpart (single process, single device).
#include "gg_common.cuh"

namespace {

constexpr int kPushThreads = 512;

struct Peers {
  int n;
  void* ptr[GG_MAX_PEERS];
};

__global__ void __launch_bounds__(kPushThreads) push_kernel(const uint4* src, long long n_vec, Peers dst, int slices) {
  // blockIdx.x = peer * slices + slice: every peer's copy gets its own CTAs, so all NVLink ports are busy at once
  const int peer = blockIdx.x / slices, slice = blockIdx.x % slices;
  uint4* out = static_cast<uint4*>(dst.ptr[peer]);
  if (out == nullptr) return;
  for (long long i = slice * static_cast<long long>(kPushThreads) + threadIdx.x; i < n_vec;
       i += static_cast<long long>(slices) * kPushThreads)
    out[i] = __ldg(src + i);
}

__global__ void signal_kernel(Peers flags, int index, uint32_t value) {
  __threadfence_system();  // the pushes of the kernels before this one are performed before the flag becomes visible
  const int p = threadIdx.x;
  if (p < flags.n && flags.ptr[p] != nullptr) {
    uint32_t* f = static_cast<uint32_t*>(flags.ptr[p]) + index;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(value) : "memory");
  }
}

__global__ void wait_kernel(const uint32_t* flags, int first, int count, int skip, uint32_t value, long long* status) {
  const int r = threadIdx.x;
  if (r < count && r != skip) {
    const uint32_t* f = flags + first + r;
    bool ok = false;
    for (long long spin = 0; spin < (1ll << 26); ++spin) {   // ~10 s: a build is milliseconds
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (static_cast<int32_t>(v - value) >= 0) {
        ok = true;
        break;
      }
      __nanosleep(64);
    }
    if (!ok && status) atomicAdd(reinterpret_cast<unsigned long long*>(status), 1ull);  // a peer never arrived
  }
  __syncthreads();
  __threadfence_system();
}

}  // namespace

extern "C" int gg_peer_push(const void* src, int64_t n_bytes, const gg_peers_t* h_dst, gg_stream_t st) {
  GG_REQUIRE(h_dst != nullptr && h_dst->n >= 1 && h_dst->n <= GG_MAX_PEERS, GG_ERR_ARG, "1..16 peers expected");
  GG_REQUIRE(n_bytes >= 0 && n_bytes % 16 == 0, GG_ERR_ARG, "segment size must be a multiple of 16 bytes");
  if (n_bytes == 0) return GG_OK;
  GG_REQUIRE(src != nullptr && (reinterpret_cast<uintptr_t>(src) & 15) == 0, GG_ERR_ARG, "source must be 16-byte aligned");
  Peers p;
  p.n = h_dst->n;
  int live = 0;
  for (int i = 0; i < GG_MAX_PEERS; ++i) {
    p.ptr[i] = i < p.n ? h_dst->ptr[i] : nullptr;
    GG_REQUIRE((reinterpret_cast<uintptr_t>(p.ptr[i]) & 15) == 0, GG_ERR_ARG, "destinations must be 16-byte aligned");
    live += p.ptr[i] != nullptr;
  }
  if (live == 0) return GG_OK;
  const long long n_vec = n_bytes / 16;
  long long slices = (n_vec + kPushThreads * 4 - 1) / (kPushThreads * 4);  // >= 4 vectors per thread
  const long long cap = (static_cast<long long>(gg::sm_count()) * 4 + p.n - 1) / p.n;
  if (slices > cap) slices = cap;
  if (slices < 1) slices = 1;
  push_kernel<<<static_cast<unsigned>(p.n * slices), kPushThreads, 0, gg::as_stream(st)>>>(
      static_cast<const uint4*>(src), n_vec, p, static_cast<int>(slices));
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}

extern "C" int gg_peer_signal(const gg_peers_t* h_flags, int index, uint32_t value, gg_stream_t st) {
  GG_REQUIRE(h_flags != nullptr && h_flags->n >= 1 && h_flags->n <= GG_MAX_PEERS && index >= 0, GG_ERR_ARG,
             "1..16 peers expected");
  Peers p;
  p.n = h_flags->n;
  for (int i = 0; i < GG_MAX_PEERS; ++i) p.ptr[i] = i < p.n ? h_flags->ptr[i] : nullptr;
  signal_kernel<<<1, 32, 0, gg::as_stream(st)>>>(p, index, value);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}

extern "C" int gg_peer_wait(const uint32_t* flags, int first, int count, int skip, uint32_t value, int64_t* status,
                            gg_stream_t st) {
  GG_REQUIRE(flags != nullptr && first >= 0 && count >= 1 && count <= 32, GG_ERR_ARG, "1..32 flags expected");
  wait_kernel<<<1, 32, 0, gg::as_stream(st)>>>(flags, first, count, skip, value, reinterpret_cast<long long*>(status));
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}
