// K3: sub-k-mer frequency features of the graph nodes.
//
// Replaces node_embedding_initial(method="kmer_frequency") (reference
// gnn_utils.py:428-493): counts[i, c] = occurrences of s-mer c among the k-s+1
// overlapping s-mers of node i.  The integer counts are what the KF kernels
// consume; the float features the GNN sees are a table look-up of them.
#include "gg_common.cuh"

namespace {

// n_dev (nullable): the number of rows, still on the device (at most n; rows beyond it keep their zeros)
__global__ void __launch_bounds__(256) counts_kernel(const int64_t* node_codes, int64_t n, const int64_t* n_dev, int k,
                                                     int s, uint8_t* counts, int32_t* sqnorm, uint32_t* col_present) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  if (n_dev != nullptr && i >= *n_dev) {
    sqnorm[i] = 0;
    return;
  }
  const uint64_t code = static_cast<uint64_t>(node_codes[i]);
  const int64_t d = 1ll << (2 * s);
  const uint64_t smask = static_cast<uint64_t>(d - 1);
  uint8_t* row = counts + i * d;  // zeroed by the launcher; this thread is the row's only writer
  const int m = k - s + 1;
  for (int j = 0; j < m; ++j) {
    const uint32_t col = static_cast<uint32_t>((code >> (2 * (k - s - j))) & smask);
    row[col] += 1;
    col_present[col] = 1;
  }
  int sq = 0;  // sum over occurrences of the count of that column == sum of squares
  for (int j = 0; j < m; ++j) sq += row[(code >> (2 * (k - s - j))) & smask];
  sqnorm[i] = sq;
}

__global__ void __launch_bounds__(256) features_kernel(const uint8_t* counts, int64_t n, int d, const int32_t* col_keep,
                                                       int d_keep, const float* lut, float* x) {
  const int64_t total = n * d_keep;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t i = e / d_keep;
    const int j = static_cast<int>(e - i * d_keep);
    x[e] = lut[counts[i * d + col_keep[j]]];
  }
}

// every column kept (the usual case) and d a multiple of 4: four counts in, one float4 out, fully coalesced
__global__ void __launch_bounds__(256) features_all_kernel(const uint32_t* counts4, int64_t n_words, const float* lut,
                                                           float4* x) {
  __shared__ float s_lut[256];
  s_lut[threadIdx.x] = lut[threadIdx.x];
  __syncthreads();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  constexpr int kUnroll = 4;
  for (int64_t e0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e0 < n_words; e0 += stride * kUnroll) {
    uint32_t w[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t e = e0 + u * stride;
      w[u] = e < n_words ? __ldg(counts4 + e) : 0u;
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t e = e0 + u * stride;
      if (e < n_words)
        __stcs(x + e, make_float4(s_lut[w[u] & 0xFFu], s_lut[(w[u] >> 8) & 0xFFu], s_lut[(w[u] >> 16) & 0xFFu],
                                  s_lut[w[u] >> 24]));
    }
  }
}

}  // namespace

static int kf_counts_impl(const int64_t* node_codes, int64_t n, const int64_t* n_dev, int k, int s, uint8_t* counts,
                          int32_t* sqnorm, uint32_t* col_present, gg_stream_t st) {
  GG_REQUIRE(k >= 1 && k <= 31, GG_ERR_K, "k outside [1, 31]");
  GG_REQUIRE(s >= 1 && s <= k && s <= 8, GG_ERR_K, "small_k outside [1, min(k, 8)]");
  GG_REQUIRE(k - s + 1 <= 15, GG_ERR_K, "k - small_k + 1 must be <= 15 (uint8 counts, uint16 dmin)");
  GG_REQUIRE(n >= 0 && (n == 0 || (node_codes && counts && sqnorm && col_present)), GG_ERR_ARG, "null buffer");
  cudaStream_t stream = gg::as_stream(st);
  const int64_t d = 1ll << (2 * s);
  GG_CUDA_CHECK(cudaMemsetAsync(col_present, 0, d * sizeof(uint32_t), stream));
  if (n == 0) return GG_OK;
  GG_CUDA_CHECK(cudaMemsetAsync(counts, 0, static_cast<size_t>(n) * d, stream));
  counts_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(node_codes, n, n_dev, k, s, counts, sqnorm,
                                                                            col_present);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}

extern "C" int gg_kf_counts(const int64_t* node_codes, int64_t n, int k, int s, uint8_t* counts, int32_t* sqnorm,
                            uint32_t* col_present, gg_stream_t st) {
  return kf_counts_impl(node_codes, n, nullptr, k, s, counts, sqnorm, col_present, st);
}

extern "C" int gg_kf_counts_upto(const int64_t* node_codes, int64_t n_max, const int64_t* n_dev, int k, int s,
                                 uint8_t* counts, int32_t* sqnorm, uint32_t* col_present, gg_stream_t st) {
  GG_REQUIRE(n_dev != nullptr, GG_ERR_ARG, "null row count");
  return kf_counts_impl(node_codes, n_max, n_dev, k, s, counts, sqnorm, col_present, st);
}

extern "C" int gg_kf_features_f32(const uint8_t* counts, int64_t n, int d, const int32_t* col_keep, int d_keep,
                                  const float* value_lut, float* x, gg_stream_t st) {
  GG_REQUIRE(n >= 0 && d > 0 && d_keep >= 0 && d_keep <= d, GG_ERR_ARG, "bad shape");
  if (n == 0 || d_keep == 0) return GG_OK;
  GG_REQUIRE(counts && col_keep && value_lut && x, GG_ERR_ARG, "null buffer");
  const int64_t total = n * d_keep;
  const int64_t cap = static_cast<int64_t>(gg::sm_count()) * 16;
  if (d_keep == d && d % 4 == 0 && (reinterpret_cast<uintptr_t>(counts) & 3) == 0 &&
      (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const int64_t n_words = total / 4;
    int64_t g = (n_words + 256 * 4 - 1) / (256 * 4);
    if (g > cap) g = cap;
    features_all_kernel<<<static_cast<unsigned>(g), 256, 0, gg::as_stream(st)>>>(
        reinterpret_cast<const uint32_t*>(counts), n_words, value_lut, reinterpret_cast<float4*>(x));
    GG_CUDA_CHECK(cudaGetLastError());
    return GG_OK;
  }
  int64_t grid = (total + 255) / 256;
  if (grid > cap) grid = cap;
  features_kernel<<<static_cast<unsigned>(grid), 256, 0, gg::as_stream(st)>>>(counts, n, d, col_keep, d_keep,
                                                                             value_lut, x);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}
