// f2: reads -> k-mer index arrays (the `transform()` half of the representation API).
//
// Replaces the per-read Python loop of KmerSsgnn.transform / KmerSsgnnMb.transform (reference
// src/representations/kmer_ssgnn.py:100-160, word2index at :148-158): every window seq[i:i+k], i = 0, stride, ...
// becomes index_dict[kmer], windows holding an 'N' become index_dict["N"*k] (seq_to_kmers(..., inlcudeUNK=True),
// src/utils.py:28-30).  One warp per read, lanes over windows.
#include "gg_common.cuh"

namespace {

__global__ void __launch_bounds__(256) kmer_index_kernel(const uint8_t* stream, const int64_t* read_start,
                                                         const int64_t* read_len, int64_t n_reads, int k, int stride,
                                                         const int32_t* code_to_index, int64_t unk_index,
                                                         const int64_t* out_offsets, int64_t* out,
                                                         unsigned long long* status) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  unsigned long long missing = 0, bad = 0;
  for (int64_t r = warp; r < n_reads; r += n_warps) {
    const int64_t start = read_start[r], len = read_len[r];
    const int64_t n_win = len >= k ? (len - k) / stride + 1 : 0;
    int64_t* dst = out + out_offsets[r];
    for (int64_t w = lane; w < n_win; w += 32) {
      const uint8_t* p = stream + start + w * stride;
      uint32_t code = 0;
      bool has_n = false, illegal = false;
      for (int j = 0; j < k; ++j) {
        const uint8_t ch = __ldg(p + j);
        uint32_t c = 0;
        switch (ch) {
          case 'A': c = 0; break;
          case 'C': c = 1; break;
          case 'G': c = 2; break;
          case 'T': c = 3; break;
          case 'N': has_n = true; break;
          default: illegal = true; break;
        }
        code = (code << 2) | c;
      }
      int64_t idx = unk_index;
      if (illegal) {
        ++bad;
        idx = -1;
      } else if (!has_n) {
        idx = code_to_index[code];
        if (idx < 0) ++missing;
      }
      dst[w] = idx;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    missing += __shfl_xor_sync(0xFFFFFFFFu, missing, o);
    bad += __shfl_xor_sync(0xFFFFFFFFu, bad, o);
  }
  if (lane == 0) {
    if (bad) atomicAdd(status + 0, bad);
    if (missing) atomicAdd(status + 1, missing);
  }
}

}  // namespace

extern "C" int gg_kmer_index(const uint8_t* stream, const int64_t* read_start, const int64_t* read_len,
                             int64_t n_reads, int k, int stride, const int32_t* code_to_index, int64_t unk_index,
                             const int64_t* out_offsets, int64_t* out, int64_t* status, gg_stream_t st) {
  GG_REQUIRE(k >= 1 && k <= 15, GG_ERR_K, "k outside [1, 15]");
  GG_REQUIRE(stride >= 1 && n_reads >= 0, GG_ERR_ARG, "bad stride / read count");
  GG_REQUIRE(status, GG_ERR_ARG, "null status");
  cudaStream_t s = gg::as_stream(st);
  GG_CUDA_CHECK(cudaMemsetAsync(status, 0, 2 * sizeof(int64_t), s));
  if (n_reads == 0) return GG_OK;
  GG_REQUIRE(stream && read_start && read_len && code_to_index && out_offsets && out, GG_ERR_ARG, "null buffer");
  const int64_t warps_needed = n_reads;
  int64_t grid = (warps_needed * 32 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(gg::sm_count()) * 16;
  if (grid > cap) grid = cap;
  kmer_index_kernel<<<static_cast<unsigned>(grid), 256, 0, s>>>(stream, read_start, read_len, n_reads, k, stride,
                                                                code_to_index, unk_index, out_offsets, out,
                                                                reinterpret_cast<unsigned long long*>(status));
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}
