// Host-side helpers of the host-to-host build (no device code).
//
// The float32 node features are a 256-entry table look-up of the uint8 sub-k-mer counts (K3), four times their size.
// When the graph is handed to CPU-side loaders, moving the counts over PCIe and expanding them with the host cores
// WHILE the (much larger) edge lists are still crossing the bus takes the expansion off the critical path:
// 67 MB instead of 268 MB on the wire at k=8, small_k=5, and ~4.5 ms of 16 host threads hidden under 7.5 ms of DMA.
#include <thread>
#include <vector>

#include "gg_common.cuh"

extern "C" int gg_host_features_f32(const uint8_t* h_counts, int64_t n_elems, const float* h_value_lut, float* h_x,
                                    int n_threads) {
  GG_REQUIRE(n_elems >= 0 && n_threads >= 1 && n_threads <= 256, GG_ERR_ARG, "bad size / thread count");
  if (n_elems == 0) return GG_OK;
  GG_REQUIRE(h_counts && h_value_lut && h_x, GG_ERR_ARG, "null buffer");
  float lut[256];
  for (int i = 0; i < 256; ++i) lut[i] = h_value_lut[i];
  auto work = [&](int64_t lo, int64_t hi) {
    const uint8_t* s = h_counts;
    float* d = h_x;
    for (int64_t i = lo; i < hi; ++i) d[i] = lut[s[i]];
  };
  if (n_threads == 1 || n_elems < (1 << 16)) {
    work(0, n_elems);
    return GG_OK;
  }
  std::vector<std::thread> pool;
  pool.reserve(n_threads);
  for (int t = 0; t < n_threads; ++t) {
    // 64-byte aligned cuts: no two threads share a cache line of the output
    const int64_t lo = n_elems * t / n_threads / 16 * 16, hi = t + 1 == n_threads ? n_elems : n_elems * (t + 1) / n_threads / 16 * 16;
    pool.emplace_back(work, lo, hi);
  }
  for (auto& th : pool) th.join();
  return GG_OK;
}

// Asynchronous flavour: a native thread waits for the CUDA event that marks the arrival of the counts on the host and
// then runs the expansion, so that nothing depends on the caller's interpreter getting scheduled (a Python worker
// thread can wait up to a GIL switch interval, 5 ms, before it even starts).
namespace {
struct HostJob {
  std::thread th;
  int rc = GG_OK;
};
}  // namespace

extern "C" void* gg_host_features_f32_start(const uint8_t* h_counts, int64_t n_elems, const float* h_value_lut,
                                            float* h_x, int n_threads, int device, void* cuda_event) {
  if (n_elems < 0 || n_threads < 1 || n_threads > 256 || (n_elems > 0 && (!h_counts || !h_value_lut || !h_x))) {
    gg::set_error("gg_host_features_f32_start: bad argument");
    return nullptr;
  }
  HostJob* job = new HostJob();
  std::vector<float> lut(h_value_lut, h_value_lut + 256);
  job->th = std::thread([=]() {
    if (cuda_event) {
      if (cudaSetDevice(device) != cudaSuccess || cudaEventSynchronize(static_cast<cudaEvent_t>(cuda_event)) != cudaSuccess) {
        job->rc = GG_ERR_CUDA;
        return;
      }
    }
    job->rc = gg_host_features_f32(h_counts, n_elems, lut.data(), h_x, n_threads);
  });
  return job;
}

extern "C" int gg_host_job_wait(void* handle) {
  GG_REQUIRE(handle != nullptr, GG_ERR_ARG, "null job");
  HostJob* job = static_cast<HostJob*>(handle);
  job->th.join();
  const int rc = job->rc;
  delete job;
  if (rc != GG_OK) gg::set_error("host feature expansion failed (status %d)", rc);
  return rc;
}
