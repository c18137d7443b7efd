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

// ---------------------------------------------------------------------------------------------
// f4: pairs of consecutive surviving k-mers for stride > 1 (higher-stride DB graphs,
// reference kmer_ssgnn.py:215-226 -> weighted_directed_edges(..., stride=s, inlcudeUNK=False)).
// Windows sit at 0, s, 2s, ... of every read; windows holding an N are dropped before pairing
// (src/utils.py:28-31, 79-82).  A pair is no longer a (k+1)-mer, so every pair becomes one
// (u, v, pos) record; gg_db_build de-duplicates them (count, first position) exactly as it does
// the stride-1 bridge records.  One warp per read, 32 windows per step, the last clean window of
// a step is carried into the next one.
// ---------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256) stride_pairs_kernel(const uint8_t* stream, const int64_t* read_start,
                                                           const int64_t* read_len, int64_t n_reads, int k, int stride,
                                                           int64_t pos_base, gg_bridge_t* recs, int64_t capacity,
                                                           unsigned long long* status) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  unsigned long long bad = 0;
  for (int64_t r = warp; r < n_reads; r += n_warps) {
    const int64_t start = read_start[r], len = read_len[r];
    const int64_t n_win = len >= k ? (len - k) / stride + 1 : 0;
    bool have_prev = false;       // warp-uniform: a clean window is waiting for its successor
    uint32_t prev_code = 0;
    int64_t prev_pos = 0;
    for (int64_t w0 = 0; w0 < n_win; w0 += 32) {
      const int64_t w = w0 + lane;
      uint32_t code = 0;
      bool clean = w < n_win;
      if (clean) {
        const uint8_t* p = stream + start + w * stride;
        for (int j = 0; j < k; ++j) {
          const uint8_t ch = __ldg(p + j);
          uint32_t c = 0;
          switch (ch) {
            case 'A': c = 0; break;
            case 'C': c = 1; break;
            case 'G': c = 2; break;
            case 'T': c = 3; break;
            case 'N': clean = false; break;
            default: clean = false; ++bad; break;
          }
          code = (code << 2) | c;
        }
      }
      const uint32_t cmask = __ballot_sync(0xFFFFFFFFu, clean);
      if (cmask == 0) continue;
      // successor of lane i = next clean lane of this step; the carried window pairs with the first clean lane
      const uint32_t later = lane == 31 ? 0u : (cmask >> (lane + 1));
      const int succ = later ? lane + __ffs(later) : -1;
      const uint32_t succ_code = __shfl_sync(0xFFFFFFFFu, code, succ < 0 ? 0 : succ);
      const bool emits = clean && succ >= 0;
      const int first_clean = __ffs(cmask) - 1, last_clean = 31 - __clz(cmask);
      const bool carry_emits = have_prev && lane == first_clean;
      const int n_mine = (emits ? 1 : 0) + (carry_emits ? 1 : 0);
      int incl = n_mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += v;
      }
      const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
      unsigned long long base = 0;
      if (lane == 0 && total) base = atomicAdd(status + GG_ST_BRIDGES, static_cast<unsigned long long>(total));
      base = __shfl_sync(0xFFFFFFFFu, base, 0) + (incl - n_mine);
      if (carry_emits) {
        if (static_cast<int64_t>(base) < capacity) {
          gg_bridge_t rec;
          rec.u = prev_code;
          rec.v = code;
          rec.pos = static_cast<uint64_t>(prev_pos);
          recs[base] = rec;
        }
        ++base;
      }
      if (emits && static_cast<int64_t>(base) < capacity) {
        gg_bridge_t rec;
        rec.u = code;
        rec.v = succ_code;
        rec.pos = static_cast<uint64_t>(pos_base + start + w * stride);
        recs[base] = rec;
      }
      have_prev = true;
      prev_code = __shfl_sync(0xFFFFFFFFu, code, last_clean);
      prev_pos = pos_base + start + (w0 + last_clean) * stride;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xFFFFFFFFu, bad, o);
  if (lane == 0 && bad) atomicAdd(status + GG_ST_BAD_BYTES, bad);
}

}  // namespace

extern "C" int gg_count_pairs_stride(const uint8_t* stream, const int64_t* read_start, const int64_t* read_len,
                                     int64_t n_reads, int k, int stride, int64_t pos_base, gg_bridge_t* records,
                                     int64_t capacity, int64_t* status, gg_stream_t st) {
  GG_REQUIRE(k >= 1 && k <= GG_MAX_K, GG_ERR_K, "k outside [1, GG_MAX_K]");
  GG_REQUIRE(stride >= 1 && n_reads >= 0 && capacity >= 0, GG_ERR_ARG, "bad stride / sizes");
  GG_REQUIRE(status, GG_ERR_ARG, "null status");
  if (n_reads == 0) return GG_OK;
  GG_REQUIRE(stream && read_start && read_len && (capacity == 0 || records), GG_ERR_ARG, "null buffer");
  int64_t grid = (n_reads * 32 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(gg::sm_count()) * 16;
  if (grid > cap) grid = cap;
  stride_pairs_kernel<<<static_cast<unsigned>(grid), 256, 0, gg::as_stream(st)>>>(
      stream, read_start, read_len, n_reads, k, stride, pos_base, records, capacity,
      reinterpret_cast<unsigned long long*>(status));
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}
