// K1: 2-bit encode + rolling (k+1)-mer extraction + histogram, and the
// first-occurrence scan that defines dict / node order.
//
// Replaces seq_to_kmers + weighted_directed_edges (reference src/utils.py:15-31,
// 52-84; stride 1, inlcudeUNK=False).  One thread owns 16 consecutive stream
// positions (one aligned uint4) plus a 16-byte look-ahead, classifies the 32
// bytes with SWAR arithmetic (no table), keeps the 32 two-bit codes in a 64-bit
// register pair and slides a (k+1)-base window over it.  Counts go straight to
// the L2-resident 4^(k+1)-bin table with RED.ADD: the bin space (1 MB at k=8)
// exceeds one SM's shared memory and keys are near-uniform, so privatised
// shared-memory histograms cannot help -- see DESIGN.md "K1" for the
// atomic-throughput bound this kernel runs against.
#include "gg_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kPosPerThread = 16;
constexpr int kTile = kThreads * kPosPerThread;  // stream positions per CTA iteration
constexpr int kTilesPerGroup = 4;                // first-occurrence scan: 16 KB of stream per grab

enum Mode { kCount = 0, kFirst = 1 };

struct Args {
  const uint8_t* stream;
  int64_t n_bytes;
  int64_t fixed_len;
  int64_t pos_base;
  int k;
  uint32_t* hist;
  uint64_t* first_pos;
  gg_bridge_t* bridges;
  int64_t bridge_capacity;
  unsigned long long* status;
};

// bit 7 of every byte lane set iff that byte is non-zero (exact per lane)
__device__ __forceinline__ uint32_t nonzero_lanes(uint32_t v) {
  return (((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & 0x80808080u;
}

// gather bit 0 of the four byte lanes into a nibble, lane 0 -> bit 0
__device__ __forceinline__ uint32_t lanes_to_nibble(uint32_t f) { return ((f * 0x00204081u) >> 21) & 0xFu; }

// Four ASCII bytes -> 8 bits of 2-bit codes (first byte most significant) and
// per-byte flag nibbles (first byte = bit 0).
__device__ __forceinline__ void classify4(uint32_t x, uint32_t& code8, uint32_t& n4, uint32_t& sep4,
                                          uint32_t& bad4) {
  uint32_t c = (x >> 1) & 0x03030303u;  // A,C,T,G -> 0,1,2,3
  c ^= (c >> 1) & 0x01010101u;          // A,C,G,T -> 0,1,2,3
  const uint32_t c0 = c & 0x01010101u, c1 = (c >> 1) & 0x01010101u;
  // the only byte that has this code AND is a base: 'A'+{0,2,6,0x13}
  const uint32_t expect = 0x41414141u + ((c0 | c1) << 1) + ((c1 & ~c0) << 2) + (c0 & c1) * 0x11u;
  const uint32_t not_base = nonzero_lanes(x ^ expect) >> 7;
  const uint32_t not_n = nonzero_lanes(x ^ 0x4E4E4E4Eu) >> 7;
  const uint32_t not_sep = nonzero_lanes(x ^ 0x0A0A0A0Au) >> 7;
  code8 = (c * 0x40100401u) >> 24;
  n4 = lanes_to_nibble(not_n ^ 0x01010101u);
  sep4 = lanes_to_nibble(not_sep ^ 0x01010101u);
  bad4 = lanes_to_nibble(not_base & not_n & not_sep);
}

__device__ __forceinline__ void classify16(const uint4& w, uint32_t& codes, uint32_t& nmask, uint32_t& sepmask,
                                           uint32_t& badmask) {
  uint32_t c, n, s, b;
  classify4(w.x, c, n, s, b);
  codes = c << 24; nmask = n; sepmask = s; badmask = b;
  classify4(w.y, c, n, s, b);
  codes |= c << 16; nmask |= n << 4; sepmask |= s << 4; badmask |= b << 4;
  classify4(w.z, c, n, s, b);
  codes |= c << 8; nmask |= n << 8; sepmask |= s << 8; badmask |= b << 8;
  classify4(w.w, c, n, s, b);
  codes |= c; nmask |= n << 12; sepmask |= s << 12; badmask |= b << 12;
}

// bit j of the result set iff bits j .. j+w-1 of v are all set
__device__ __forceinline__ uint32_t erode(uint32_t v, int w) {
  if (w <= 0) return 0xFFFFFFFFu;
  uint32_t r = v;
  int len = 1;
  while (len * 2 <= w) {
    r &= r >> len;
    len *= 2;
  }
  if (w > len) r &= r >> (w - len);
  return r;
}

// After the k-mer at some position, the next window is broken by an N: find the
// first clean k-mer further along the same read (reference utils.py:31 deletes
// N-holding k-mers before pairing).  q = stream offset just after that N.
__device__ bool next_clean_kmer(const Args& a, int64_t q, uint32_t& v_out, bool& bad) {
  const uint32_t kmask = (a.k >= 16) ? 0xFFFFFFFFu : ((1u << (2 * a.k)) - 1u);
  uint32_t v = 0;
  int run = 0;
  for (; q < a.n_bytes; ++q) {
    if (a.fixed_len > 0 && (q % a.fixed_len) == 0) return false;
    const uint8_t ch = a.stream[q];
    uint32_t code;
    switch (ch) {
      case 'A': code = 0; break;
      case 'C': code = 1; break;
      case 'G': code = 2; break;
      case 'T': code = 3; break;
      case 'N': run = 0; v = 0; continue;
      case GG_SEP: if (a.fixed_len == 0) return false;  // fallthrough: SEP is not legal in fixed-length mode
      default: bad = true; return false;
    }
    v = ((v << 2) | code) & kmask;
    if (++run == a.k) {
      v_out = v;
      return true;
    }
  }
  return false;
}

template <int MODE>
__device__ __forceinline__ void touch(const Args& a, uint32_t key, uint64_t pos, int& newly_seen) {
  if (MODE == kCount) {
    atomicAdd(a.hist + key, 1u);
  } else {
    unsigned long long* slot = reinterpret_cast<unsigned long long*>(a.first_pos) + key;
    if (__ldcg(slot) > pos) {
      const unsigned long long old = atomicMin(slot, static_cast<unsigned long long>(pos));
      newly_seen += (old == GG_NO_POS);
    }
  }
}

template <int MODE>
__device__ __forceinline__ void process_chunk(const Args& a, int64_t b, int& newly_seen) {
  if (b >= a.n_bytes) return;
  const int k = a.k;
  const uint4 w0 = __ldg(reinterpret_cast<const uint4*>(a.stream + b));
  const uint4 w1 = __ldg(reinterpret_cast<const uint4*>(a.stream + b + 16));
  uint32_t hi, lo, n0, n1, s0, s1, b0, b1;
  classify16(w0, hi, n0, s0, b0);
  classify16(w1, lo, n1, s1, b1);
  uint32_t nmask = n0 | (n1 << 16), sepmask = s0 | (s1 << 16), badmask = b0 | (b1 << 16);

  const int64_t left = a.n_bytes - b;  // > 0
  const uint32_t beyond = left >= 32 ? 0u : ~((1u << left) - 1u);
  nmask &= ~beyond;
  badmask &= ~beyond;
  uint32_t hard = beyond, brk = 0;
  if (a.fixed_len > 0) {
    badmask |= sepmask & ~beyond;  // a separator byte is illegal inside fixed-length rows
    const int64_t len = a.fixed_len;
    int64_t j;  // first read start at or after b
    if (a.n_bytes < (1ll << 31)) {
      const uint32_t r = static_cast<uint32_t>(b) % static_cast<uint32_t>(len);
      j = r ? static_cast<int64_t>(len - r) : 0;
    } else {
      j = (len - (b % len)) % len;
    }
    for (; j < 32; j += len) brk |= 1u << j;
  } else {
    hard |= sepmask;
  }
  if (MODE == kCount) {
    const uint32_t own_bad = badmask & 0xFFFFu;
    if (own_bad) atomicAdd(a.status + GG_ST_BAD_BYTES, static_cast<unsigned long long>(__popc(own_bad)));
  }

  const uint32_t dirty = nmask | hard | badmask;
  const uint32_t inner = ~(brk >> 1);  // bit j clear iff a read starts at j+1
  const uint32_t pair_ok = erode(~dirty, k + 1) & erode(inner, k) & 0xFFFFu;
  const int key_shift = 32 - 2 * (k + 1);
  const uint64_t pos0 = static_cast<uint64_t>(a.pos_base + b);

#pragma unroll
  for (int j = 0; j < kPosPerThread; ++j) {
    if ((pair_ok >> j) & 1u) {
      const uint32_t win = __funnelshift_l(lo, hi, 2 * j);  // bases j .. j+15, first base on top
      touch<MODE>(a, win >> key_shift, pos0 + j, newly_seen);
    }
  }

  // rare: a clean k-mer whose successor window is broken by an N inside the same read
  uint32_t cand = erode(~dirty, k) & erode(inner, k - 1) & (nmask >> k) & ~(brk >> k) & 0xFFFFu;
  while (cand) {
    const int j = __ffs(cand) - 1;
    cand &= cand - 1;
    const uint32_t win = __funnelshift_l(lo, hi, 2 * j);
    const uint32_t u = win >> (32 - 2 * k);
    uint32_t v = 0;
    bool bad = false;
    if (next_clean_kmer(a, b + j + k + 1, v, bad)) {
      const uint32_t overlap_mask = (1u << (2 * (k - 1))) - 1u;
      if ((u & overlap_mask) == (v >> 2)) {
        touch<MODE>(a, (u << 2) | (v & 3u), pos0 + j, newly_seen);
      } else if (MODE == kCount) {
        const unsigned long long slot = atomicAdd(a.status + GG_ST_BRIDGES, 1ull);
        if (static_cast<int64_t>(slot) < a.bridge_capacity) {
          gg_bridge_t r;
          r.u = u;
          r.v = v;
          r.pos = pos0 + j;
          a.bridges[slot] = r;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads) count_kernel(Args a) {
  const int64_t n_tiles = (a.n_bytes + kTile - 1) / kTile;
  int unused = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
    process_chunk<kCount>(a, tile * kTile + static_cast<int64_t>(threadIdx.x) * kPosPerThread, unused);
}

// Ordered scan with early exit: groups of tiles are handed out in stream order;
// once every non-zero bin has been seen no new group is started.  Groups already
// in flight always finish, and every group handed out later lies further along
// the stream, so the minima are final when the kernel ends.
__global__ void __launch_bounds__(kThreads) first_kernel(Args a) {
  __shared__ long long s_group;
  const int64_t n_tiles = (a.n_bytes + kTile - 1) / kTile;
  const int64_t n_groups = (n_tiles + kTilesPerGroup - 1) / kTilesPerGroup;
  while (true) {
    if (threadIdx.x == 0) {
      const long long remaining = *reinterpret_cast<volatile long long*>(a.status + GG_ST_REMAINING);
      s_group = remaining > 0 ? static_cast<long long>(atomicAdd(a.status + GG_ST_CURSOR, 1ull)) : -1;
    }
    __syncthreads();
    const long long group = s_group;
    __syncthreads();
    if (group < 0 || group >= n_groups) break;
    int newly_seen = 0;
    for (int t = 0; t < kTilesPerGroup; ++t) {
      const int64_t tile = group * kTilesPerGroup + t;
      if (tile < n_tiles)
        process_chunk<kFirst>(a, tile * kTile + static_cast<int64_t>(threadIdx.x) * kPosPerThread, newly_seen);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) newly_seen += __shfl_xor_sync(0xFFFFFFFFu, newly_seen, o);
    if ((threadIdx.x & 31) == 0 && newly_seen)
      atomicAdd(a.status + GG_ST_REMAINING, static_cast<unsigned long long>(-static_cast<long long>(newly_seen)));
  }
}

// bins that are occupied / occupied but not yet located
__global__ void __launch_bounds__(256) pending_kernel(const uint32_t* hist, const uint64_t* first_pos, int64_t bins,
                                                      unsigned long long* status) {
  int nz = 0, pend = 0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < bins;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const bool occ = hist[i] != 0;
    nz += occ;
    pend += occ && first_pos[i] == GG_NO_POS;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nz += __shfl_xor_sync(0xFFFFFFFFu, nz, o);
    pend += __shfl_xor_sync(0xFFFFFFFFu, pend, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (nz) atomicAdd(status + GG_ST_NONZERO, static_cast<unsigned long long>(nz));
    if (pend) atomicAdd(status + GG_ST_REMAINING, static_cast<unsigned long long>(pend));
  }
}

// ---------------------------------------------------------------------------------------------
// Shared-memory counting path (k <= 8).
//
// Measured on B200 (scripts/micro/atomics.cu): random 32-bit atomic adds run at >= 1600 G/s into shared memory,
// 183 G/s into an L2-resident table and 51 G/s into a cluster-distributed (DSMEM) table.  The L2 path above is
// therefore capped at ~0.8 ms for 1.42e8 windows.  The 4^(k+1)-bin table does not fit one SM, so the bin space is
// cut into `n_ranges` slices of <= kSmemBins bins and the CTAs are dealt round-robin to the slices: CTA b counts
// only keys of slice b % n_ranges over slab b / n_ranges of the input, in a private shared-memory histogram, and adds
// its slice to the global table once at the end.  Every key is extracted n_ranges times, so the stream is first
// packed (pack_kernel: 2-bit codes + two flag bits per base, one pass over the ASCII bytes) and the counting
// passes only shift and mask.
// ---------------------------------------------------------------------------------------------
constexpr int kSmemBins = 52432;     // 4^9 / 5 rounded up to a multiple of 16: 209,728 bytes of shared memory
constexpr int kSmemThreads = 1024;
constexpr int kMaxRanges = 8;

struct Packed {
  uint32_t* codes;   // 16 bases per word, first base in the top bits
  uint16_t* bad;     // N, separator, illegal byte or beyond the end: the base cannot be part of a k-mer
  uint16_t* hard;    // separator, illegal byte or beyond the end: a k-mer search must stop here
  int64_t n_words;   // ceil(n_bytes / 16) + 1 all-bad tail word
};

__global__ void __launch_bounds__(256) pack_kernel(const uint8_t* stream, int64_t n_bytes, int64_t fixed_len,
                                                   Packed pk, unsigned long long* status) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= pk.n_words) return;
  const int64_t b = i * 16;
  if (b >= n_bytes) {
    pk.codes[i] = 0;
    pk.bad[i] = 0xFFFF;
    pk.hard[i] = 0xFFFF;
    return;
  }
  const uint4 w = __ldg(reinterpret_cast<const uint4*>(stream + b));
  uint32_t codes, nmask, sepmask, badmask;
  classify16(w, codes, nmask, sepmask, badmask);
  const int64_t left = n_bytes - b;
  const uint32_t beyond = left >= 16 ? 0u : (~((1u << left) - 1u) & 0xFFFFu);
  nmask &= ~beyond;
  badmask &= ~beyond;
  sepmask &= ~beyond;
  if (fixed_len > 0) {
    badmask |= sepmask;  // a separator byte is illegal inside fixed-length rows
    sepmask = 0;
  }
  if (badmask) atomicAdd(status + GG_ST_BAD_BYTES, static_cast<unsigned long long>(__popc(badmask)));
  const uint32_t hard = sepmask | badmask | beyond;
  pk.codes[i] = codes;
  pk.bad[i] = static_cast<uint16_t>(nmask | hard);
  pk.hard[i] = static_cast<uint16_t>(hard);
}

struct SmemArgs {
  Packed pk;
  int64_t n_bytes;
  int64_t fixed_len;
  int64_t pos_base;
  int k;
  int n_ranges;
  uint32_t bins_per_range;
  uint32_t* hist;
  gg_bridge_t* bridges;
  int64_t bridge_capacity;
  unsigned long long* status;
};

__device__ __forceinline__ uint32_t packed_base(const Packed& pk, int64_t q) {
  return (pk.codes[q >> 4] >> (30 - 2 * static_cast<int>(q & 15))) & 3u;
}

// packed-stream version of next_clean_kmer
__device__ bool next_clean_kmer_packed(const SmemArgs& a, int64_t q, uint32_t& v_out) {
  const uint32_t kmask = (1u << (2 * a.k)) - 1u;
  uint32_t v = 0;
  int run = 0;
  for (; q < a.n_bytes; ++q) {
    if (a.fixed_len > 0 && (q % a.fixed_len) == 0) return false;
    const uint32_t bit = 1u << (q & 15);
    if (a.pk.hard[q >> 4] & bit) return false;
    if (a.pk.bad[q >> 4] & bit) {  // an N
      run = 0;
      v = 0;
      continue;
    }
    v = ((v << 2) | packed_base(a.pk, q)) & kmask;
    if (++run == a.k) {
      v_out = v;
      return true;
    }
  }
  return false;
}

__global__ void __launch_bounds__(kSmemThreads, 1) count_smem_kernel(SmemArgs a) {
  extern __shared__ uint32_t h[];
  const int range = blockIdx.x % a.n_ranges;
  const int slab = blockIdx.x / a.n_ranges, n_slabs = gridDim.x / a.n_ranges;
  if (slab >= n_slabs) return;  // the few CTAs that do not complete a round of slices stay idle
  const uint32_t lo = static_cast<uint32_t>(range) * a.bins_per_range;
  for (uint32_t i = threadIdx.x; i < a.bins_per_range; i += kSmemThreads) h[i] = 0;
  __syncthreads();

  const int k = a.k;
  const int key_shift = 32 - 2 * (k + 1);
  const int64_t n_data_words = a.pk.n_words - 1;  // the tail word is padding
  const int64_t w_begin = n_data_words * slab / n_slabs, w_end = n_data_words * (slab + 1) / n_slabs;
  constexpr int kBatch = 4;  // words per thread per iteration: all loads are issued before any is consumed
  for (int64_t i0 = w_begin + threadIdx.x; i0 < w_end; i0 += static_cast<int64_t>(kSmemThreads) * kBatch) {
    uint32_t hi_[kBatch], lo_[kBatch], dirty_[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int64_t i = i0 + static_cast<int64_t>(u) * kSmemThreads;
      const bool in = i < w_end;
      hi_[u] = in ? __ldg(a.pk.codes + i) : 0u;
      lo_[u] = in ? __ldg(a.pk.codes + i + 1) : 0u;
      dirty_[u] = in ? (static_cast<uint32_t>(__ldg(a.pk.bad + i)) | (static_cast<uint32_t>(__ldg(a.pk.bad + i + 1)) << 16))
                     : 0xFFFFFFFFu;
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int64_t i = i0 + static_cast<int64_t>(u) * kSmemThreads;
      const uint32_t hi = hi_[u], lo_w = lo_[u], dirty = dirty_[u];
      if (dirty == 0xFFFFFFFFu) continue;
      const int64_t b = i * 16;
      uint32_t brk = 0;
      if (a.fixed_len > 0) {
        const int64_t len = a.fixed_len;
        int64_t j;
        if (a.n_bytes < (1ll << 31)) {
          const uint32_t r = static_cast<uint32_t>(b) % static_cast<uint32_t>(len);
          j = r ? static_cast<int64_t>(len - r) : 0;
        } else {
          j = (len - (b % len)) % len;
        }
        for (; j < 32; j += len) brk |= 1u << j;
      }
      const uint32_t inner = ~(brk >> 1);
      const uint32_t pair_ok = erode(~dirty, k + 1) & erode(inner, k) & 0xFFFFu;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t idx = (__funnelshift_l(lo_w, hi, 2 * j) >> key_shift) - lo;
        if (((pair_ok >> j) & 1u) && idx < a.bins_per_range) atomicAdd(&h[idx], 1u);
      }
      if (range == 0) {  // one slice owner also handles the rare N bridges of its slab
        const uint32_t hard = static_cast<uint32_t>(__ldg(a.pk.hard + i)) | (static_cast<uint32_t>(__ldg(a.pk.hard + i + 1)) << 16);
        const uint32_t nmask = dirty & ~hard;
        uint32_t cand = erode(~dirty, k) & erode(inner, k - 1) & (nmask >> k) & ~(brk >> k) & 0xFFFFu;
        while (cand) {
          const int j = __ffs(cand) - 1;
          cand &= cand - 1;
          const uint32_t uu = __funnelshift_l(lo_w, hi, 2 * j) >> (32 - 2 * k);
          uint32_t v = 0;
          if (next_clean_kmer_packed(a, b + j + k + 1, v)) {
            const uint32_t overlap_mask = (1u << (2 * (k - 1))) - 1u;
            if ((uu & overlap_mask) == (v >> 2)) {
              atomicAdd(a.hist + ((uu << 2) | (v & 3u)), 1u);
            } else {
              const unsigned long long slot = atomicAdd(a.status + GG_ST_BRIDGES, 1ull);
              if (static_cast<int64_t>(slot) < a.bridge_capacity) {
                gg_bridge_t r;
                r.u = uu;
                r.v = v;
                r.pos = static_cast<uint64_t>(a.pos_base + b + j);
                a.bridges[slot] = r;
              }
            }
          }
        }
      }
    }
  }
  __syncthreads();
  const uint32_t total_bins = 1u << (2 * (k + 1));
  for (uint32_t i = threadIdx.x; i < a.bins_per_range; i += kSmemThreads) {
    const uint32_t v = h[i];
    if (v && lo + i < total_bins) atomicAdd(a.hist + lo + i, v);
  }
}

int check_stream_args(const uint8_t* stream, int64_t n_bytes, int64_t fixed_len, int k) {
  GG_REQUIRE(k >= 1 && k <= GG_MAX_K, GG_ERR_K, "k outside [1, GG_MAX_K]");
  GG_REQUIRE(n_bytes >= 0 && fixed_len >= 0, GG_ERR_ARG, "negative size");
  GG_REQUIRE(n_bytes == 0 || stream != nullptr, GG_ERR_ARG, "null stream");
  GG_REQUIRE((reinterpret_cast<uintptr_t>(stream) & 15) == 0, GG_ERR_ARG, "stream must be 16-byte aligned");
  GG_REQUIRE(fixed_len == 0 || n_bytes % fixed_len == 0, GG_ERR_ARG, "n_bytes is not a multiple of fixed_len");
  return GG_OK;
}

}  // namespace

extern "C" size_t gg_stream_padded_bytes(int64_t n_bytes) {
  return n_bytes <= 0 ? 32 : (static_cast<size_t>(n_bytes) + 15) / 16 * 16 + 16;
}

extern "C" size_t gg_hist_bins(int k) { return (k < 0 || k > GG_MAX_K) ? 0 : static_cast<size_t>(1) << (2 * (k + 1)); }

extern "C" int gg_count_reset(int k, uint32_t* hist, uint64_t* first_pos, int64_t* status, gg_stream_t st) {
  GG_REQUIRE(k >= 1 && k <= GG_MAX_K, GG_ERR_K, "k outside [1, GG_MAX_K]");
  GG_REQUIRE(hist && status, GG_ERR_ARG, "null buffer");
  const size_t bins = gg_hist_bins(k);
  GG_CUDA_CHECK(cudaMemsetAsync(hist, 0, bins * sizeof(uint32_t), gg::as_stream(st)));
  if (first_pos) GG_CUDA_CHECK(cudaMemsetAsync(first_pos, 0xFF, bins * sizeof(uint64_t), gg::as_stream(st)));
  GG_CUDA_CHECK(cudaMemsetAsync(status, 0, GG_ST_WORDS * sizeof(int64_t), gg::as_stream(st)));
  return GG_OK;
}

extern "C" int gg_count_kp1mers(const uint8_t* stream, int64_t n_bytes, int64_t fixed_len, int k, int64_t pos_base,
                                uint32_t* hist, gg_bridge_t* bridges, int64_t bridge_capacity, int64_t* status,
                                gg_stream_t st) {
  if (int rc = check_stream_args(stream, n_bytes, fixed_len, k)) return rc;
  GG_REQUIRE(hist && status, GG_ERR_ARG, "null buffer");
  GG_REQUIRE(bridge_capacity >= 0 && (bridge_capacity == 0 || bridges), GG_ERR_ARG, "bridge list");
  if (n_bytes == 0) return GG_OK;
  Args a{stream, n_bytes, fixed_len, pos_base, k, hist, nullptr, bridges, bridge_capacity,
         reinterpret_cast<unsigned long long*>(status)};
  const int64_t n_tiles = (n_bytes + kTile - 1) / kTile;
  const int64_t resident = static_cast<int64_t>(gg::sm_count()) * (2048 / kThreads);
  const int grid = static_cast<int>(n_tiles < resident ? n_tiles : resident);
  count_kernel<<<grid, kThreads, 0, gg::as_stream(st)>>>(a);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}

extern "C" int gg_first_occurrence(const uint8_t* stream, int64_t n_bytes, int64_t fixed_len, int k,
                                   int64_t pos_base, const uint32_t* hist, uint64_t* first_pos, int64_t* status,
                                   gg_stream_t st) {
  if (int rc = check_stream_args(stream, n_bytes, fixed_len, k)) return rc;
  GG_REQUIRE(hist && first_pos && status, GG_ERR_ARG, "null buffer");
  if (n_bytes == 0) return GG_OK;
  cudaStream_t s = gg::as_stream(st);
  unsigned long long* stw = reinterpret_cast<unsigned long long*>(status);
  // status[CURSOR], [REMAINING], [NONZERO] are contiguous words 2..4
  GG_CUDA_CHECK(cudaMemsetAsync(stw + GG_ST_CURSOR, 0, 3 * sizeof(unsigned long long), s));
  const int64_t bins = static_cast<int64_t>(gg_hist_bins(k));
  const int pgrid = static_cast<int>((bins + 255) / 256 < 4 * gg::sm_count() ? (bins + 255) / 256 : 4 * gg::sm_count());
  pending_kernel<<<pgrid, 256, 0, s>>>(hist, first_pos, bins, stw);
  GG_CUDA_CHECK(cudaGetLastError());
  Args a{stream, n_bytes, fixed_len, pos_base, k, const_cast<uint32_t*>(hist), first_pos, nullptr, 0, stw};
  const int64_t n_tiles = (n_bytes + kTile - 1) / kTile;
  const int64_t n_groups = (n_tiles + kTilesPerGroup - 1) / kTilesPerGroup;
  // one CTA per SM: every CTA scans at least one group, so the floor of the early-exit scan is grid x 16 KB
  const int64_t resident = static_cast<int64_t>(gg::sm_count());
  const int grid = static_cast<int>(n_groups < resident ? n_groups : resident);
  first_kernel<<<grid, kThreads, 0, s>>>(a);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}

extern "C" size_t gg_count_smem_workspace_bytes(int64_t n_bytes, int k) {
  if (n_bytes < 0 || k < 1 || k > GG_MAX_K) return 0;
  const size_t bins = static_cast<size_t>(1) << (2 * (k + 1));
  if ((bins + kSmemBins - 1) / kSmemBins > kMaxRanges) return 0;  // table too large for this path: use gg_count_kp1mers
  const size_t n_words = (static_cast<size_t>(n_bytes) + 15) / 16 + 1;
  gg::Carver c(nullptr);
  c.take<uint32_t>(n_words);
  c.take<uint16_t>(n_words);
  c.take<uint16_t>(n_words);
  return c.used();
}

extern "C" int gg_count_kp1mers_smem(const uint8_t* stream, int64_t n_bytes, int64_t fixed_len, int k,
                                     int64_t pos_base, uint32_t* hist, gg_bridge_t* bridges, int64_t bridge_capacity,
                                     int64_t* status, void* workspace, size_t workspace_bytes, gg_stream_t st) {
  if (int rc = check_stream_args(stream, n_bytes, fixed_len, k)) return rc;
  GG_REQUIRE(hist && status, GG_ERR_ARG, "null buffer");
  GG_REQUIRE(bridge_capacity >= 0 && (bridge_capacity == 0 || bridges), GG_ERR_ARG, "bridge list");
  const size_t need = gg_count_smem_workspace_bytes(n_bytes, k);
  GG_REQUIRE(need != 0, GG_ERR_K, "(k+1)-mer table too large for the shared-memory path; use gg_count_kp1mers");
  GG_REQUIRE(workspace && workspace_bytes >= need, GG_ERR_WORKSPACE, "workspace too small");
  if (n_bytes == 0) return GG_OK;
  cudaStream_t s = gg::as_stream(st);
  const uint32_t bins = 1u << (2 * (k + 1));
  const int n_ranges = static_cast<int>((bins + kSmemBins - 1) / kSmemBins);
  const uint32_t bins_per_range = n_ranges == 1 ? bins : (bins + n_ranges - 1) / n_ranges;
  Packed pk;
  pk.n_words = (n_bytes + 15) / 16 + 1;
  gg::Carver c(workspace);
  pk.codes = c.take<uint32_t>(pk.n_words);
  pk.bad = c.take<uint16_t>(pk.n_words);
  pk.hard = c.take<uint16_t>(pk.n_words);
  unsigned long long* stw = reinterpret_cast<unsigned long long*>(status);
  pack_kernel<<<static_cast<unsigned>((pk.n_words + 255) / 256), 256, 0, s>>>(stream, n_bytes, fixed_len, pk, stw);
  GG_CUDA_CHECK(cudaGetLastError());
  SmemArgs a{pk, n_bytes, fixed_len, pos_base, k, n_ranges, bins_per_range, hist, bridges, bridge_capacity, stw};
  const size_t smem = static_cast<size_t>(bins_per_range) * sizeof(uint32_t);
  GG_CUDA_CHECK(cudaFuncSetAttribute(count_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
  int grid = gg::sm_count();
  if (grid < n_ranges) grid = n_ranges;
  count_smem_kernel<<<grid, kSmemThreads, smem, s>>>(a);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}
