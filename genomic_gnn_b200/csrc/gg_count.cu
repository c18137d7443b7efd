// K1: 2-bit encode + rolling (k+1)-mer extraction + histogram, and the
// first-occurrence scan that defines dict / node order.
//
// Replaces seq_to_kmers + weighted_directed_edges (reference src/utils.py:15-31,
// 52-84; stride 1, inlcudeUNK=False).  One thread owns 16 consecutive stream
// positions (one aligned uint4); a warp covers 31 such chunks and lane 31 only
// supplies the look-ahead of lane 30, so every byte is classified exactly once
// (SWAR arithmetic, no table; an exact all-ACGT test short-cuts the N /
// separator / illegal-byte masks for clean warps).  The 32 two-bit codes of a
// thread live in a register pair and a (k+1)-base window slides over them.
// Counting paths: shared-memory histograms behind a key partition (k <= 8, the
// default) or one RED.ADD per window into the L2-resident table (any k).
#include "gg_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kPosPerThread = 16;
constexpr int kOwnedLanes = 31;                         // lane 31 of a warp is look-ahead only
constexpr int kWarpSpan = kOwnedLanes * kPosPerThread;  // 496 stream positions per warp pass
constexpr int kTile = (kThreads / 32) * kWarpSpan;      // stream positions per CTA iteration
constexpr int kTilesPerGroup = 4;                       // first-occurrence scan: ~16 KB of stream per grab
constexpr uint32_t kFullWarp = 0xFFFFFFFFu;

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

// Exact "all 16 bytes are A/C/G/T" test.  A, C, G, T = 0x41, 0x43, 0x47, 0x54 share bits 7,6,5,3 = 0,1,0,0; among the
// bytes that do, (b4,b2,b1,b0) must be A 0001, C 0011, G 0111 or T 1100 (the formula is checked over all byte values
// by tests/test_abi.py::test_swar_acgt_formula; the kernels by the N / illegal-byte GPU parity tests).
__device__ __forceinline__ uint32_t acgt_low_bits(uint32_t x) {  // bit 0 of each byte lane
  const uint32_t s1 = x >> 1, s2 = x >> 2, s4 = x >> 4;
  const uint32_t g1 = x & (~s2 | s1);   // b0 & (!b2 | b1): A, C, G
  const uint32_t g2 = s2 & ~s1 & ~x;    // b2 & !b1 & !b0: T
  return (~s4 & g1) | (s4 & g2);
}
__device__ __forceinline__ bool all_acgt16(const uint4& w) {
  const uint32_t fixed = ((w.x & 0xE8E8E8E8u) ^ 0x40404040u) | ((w.y & 0xE8E8E8E8u) ^ 0x40404040u) |
                         ((w.z & 0xE8E8E8E8u) ^ 0x40404040u) | ((w.w & 0xE8E8E8E8u) ^ 0x40404040u);
  const uint32_t low = acgt_low_bits(w.x) & acgt_low_bits(w.y) & acgt_low_bits(w.z) & acgt_low_bits(w.w);
  return fixed == 0u && (low & 0x01010101u) == 0x01010101u;
}
__device__ __forceinline__ uint32_t codes4(uint32_t x) {
  uint32_t c = (x >> 1) & 0x03030303u;
  c ^= (c >> 1) & 0x01010101u;
  return (c * 0x40100401u) >> 24;
}
__device__ __forceinline__ uint32_t codes16(const uint4& w) {
  return (codes4(w.x) << 24) | (codes4(w.y) << 16) | (codes4(w.z) << 8) | codes4(w.w);
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

// Warp-cooperative load + classify.  Lane l classifies the 16 bytes at b (its chunk) and takes the look-ahead codes
// and flags from lane l+1; lanes 0..30 own their 16 positions, lane 31 owns none.  Must be called by all 32 lanes.
// hi/lo: the 32 two-bit codes, first base on top.  pair_ok bit j: a clean (k+1)-mer of one read starts at b+j.
// cand bit j: a clean k-mer starts at b+j whose successor window is broken by an N inside the same read (rare).
__device__ __forceinline__ uint32_t low_mask(int n) {  // n lowest bits set, n clamped to [0, 32]
  return n <= 0 ? 0u : (n >= 32 ? 0xFFFFFFFFu : (1u << n) - 1u);
}

// `phase` = b % fixed_len (callers keep it incrementally, see Walker; ignored when fixed_len == 0)
template <bool COUNT_BAD>
__device__ __forceinline__ void prepare_warp(const Args& a, int64_t b, uint32_t phase, int lane, uint32_t& hi,
                                             uint32_t& lo, uint32_t& pair_ok, uint32_t& cand) {
  const int k = a.k;
  uint4 w = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);  // past the end: masked by `beyond`
  if (b < a.n_bytes) w = __ldg(reinterpret_cast<const uint4*>(a.stream + b));
  const uint32_t own = lane < kOwnedLanes ? 0xFFFFu : 0u;
  const int64_t left = a.n_bytes - b;
  const int avail = left >= 32 ? 32 : (left <= 0 ? 0 : static_cast<int>(left));
  const uint32_t len32 = static_cast<uint32_t>(a.fixed_len);
  const int j0 = phase ? static_cast<int>(len32 - phase) : 0;  // first read start at or after b (fixed-length rows)
  if (__all_sync(kFullWarp, all_acgt16(w)) && (len32 == 0u || len32 >= 32u)) {
    // Clean warp: no N, separator or illegal byte in any of its bytes, so only the end of the stream and (fixed
    // rows, at most one start per 32 positions) read boundaries cut windows: a (k+1)-mer at j needs j+k < avail
    // and must not span a boundary, i.e. j is outside [j0-k, j0-1].
    hi = codes16(w);
    lo = __shfl_down_sync(kFullWarp, hi, 1);
    uint32_t ok = low_mask(avail - k);
    if (len32 != 0u) ok &= ~(low_mask(j0) & ~low_mask(j0 - k));
    pair_ok = ok & own;
    cand = 0;
    return;
  }
  uint32_t n16, s16, b16;
  classify16(w, hi, n16, s16, b16);
  lo = __shfl_down_sync(kFullWarp, hi, 1);
  const uint32_t ns = n16 | (s16 << 16);
  const uint32_t ns_next = __shfl_down_sync(kFullWarp, ns, 1), b_next = __shfl_down_sync(kFullWarp, b16, 1);
  uint32_t nmask = n16 | (ns_next << 16);
  const uint32_t sepmask = s16 | (ns_next & 0xFFFF0000u);
  uint32_t badmask = b16 | (b_next << 16);
  const uint32_t beyond = ~low_mask(avail);
  nmask &= ~beyond;
  badmask &= ~beyond;
  uint32_t hard = beyond, brk = 0;
  if (len32 != 0u) {
    badmask |= sepmask & ~beyond;  // a separator byte is illegal inside fixed-length rows
    for (int64_t j = j0; j < 32; j += a.fixed_len) brk |= 1u << j;
  } else {
    hard |= sepmask;
  }
  if (COUNT_BAD) {
    const uint32_t own_bad = badmask & own;
    if (own_bad) atomicAdd(a.status + GG_ST_BAD_BYTES, static_cast<unsigned long long>(__popc(own_bad)));
  }
  const uint32_t dirty = nmask | hard | badmask;
  const uint32_t inner = ~(brk >> 1);  // bit j clear iff a read starts at j+1
  pair_ok = erode(~dirty, k + 1) & erode(inner, k) & own;
  cand = erode(~dirty, k) & erode(inner, k - 1) & (nmask >> k) & ~(brk >> k) & own;
}

// stream offset of this lane's chunk in warp pass `pass` (passes are numbered over the whole grid)
__device__ __forceinline__ int64_t chunk_offset(int64_t pass, int lane) {
  return pass * kWarpSpan + static_cast<int64_t>(lane) * kPosPerThread;
}

// A lane's walk over its chunks (one per warp pass): the offset and offset % fixed_len, advanced without a division.
struct Walker {
  int64_t b, step;
  uint32_t phase, step_phase, len;
  __device__ __forceinline__ Walker(const Args& a, int64_t first_pass, int64_t pass_stride, int lane) {
    b = chunk_offset(first_pass, lane);
    step = pass_stride * kWarpSpan;
    len = static_cast<uint32_t>(a.fixed_len);
    phase = len ? static_cast<uint32_t>(b % a.fixed_len) : 0u;
    step_phase = len ? static_cast<uint32_t>(step % a.fixed_len) : 0u;
  }
  __device__ __forceinline__ void next() {
    b += step;
    phase += step_phase;
    if (phase >= len) phase -= len;
  }
};

__device__ __forceinline__ void red_global_inc_if(uint32_t* p, uint32_t pred) {
  asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %1, 0;\n @p red.global.add.u32 [%0], 1;\n}" ::"l"(p), "r"(pred) : "memory");
}
__device__ __forceinline__ void red_shared_inc_if(uint32_t saddr, uint32_t pred) {
  asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %1, 0;\n @p red.shared.add.u32 [%0], 1;\n}" ::"r"(saddr), "r"(pred) : "memory");
}

// rare: a clean k-mer whose successor window is broken by an N inside the same read
template <int MODE>
__device__ __forceinline__ void bridge_chunk(const Args& a, int64_t b, uint32_t hi, uint32_t lo, uint32_t cand,
                                             int& newly_seen) {
  const int k = a.k;
  const uint64_t pos0 = static_cast<uint64_t>(a.pos_base + b);
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

template <int MODE>
__device__ __forceinline__ void process_warp(const Args& a, int64_t b, uint32_t phase, int lane, int& newly_seen) {
  uint32_t hi, lo, pair_ok, cand;
  prepare_warp<MODE == kCount>(a, b, phase, lane, hi, lo, pair_ok, cand);
  const int key_shift = 32 - 2 * (a.k + 1);
  const uint64_t pos0 = static_cast<uint64_t>(a.pos_base + b);
  if (MODE == kCount) {
#pragma unroll
    for (int j = 0; j < kPosPerThread; ++j)  // bases j .. j+k, first base on top
      red_global_inc_if(a.hist + (__funnelshift_l(lo, hi, 2 * j) >> key_shift), (pair_ok >> j) & 1u);
  } else {
    // first-occurrence scan: all 16 slots are read first (independent loads in flight together); an atomic is only
    // needed where the stored offset is larger, and its return value only where the bin was still unseen
    unsigned long long* slots = reinterpret_cast<unsigned long long*>(a.first_pos);
    unsigned long long cur[kPosPerThread];
#pragma unroll
    for (int j = 0; j < kPosPerThread; ++j)
      cur[j] = ((pair_ok >> j) & 1u) ? __ldcg(slots + (__funnelshift_l(lo, hi, 2 * j) >> key_shift)) : 0ull;
#pragma unroll
    for (int j = 0; j < kPosPerThread; ++j) {
      if (cur[j] > pos0 + j) {  // never for invalid positions (0)
        unsigned long long* slot = slots + (__funnelshift_l(lo, hi, 2 * j) >> key_shift);
        if (cur[j] == GG_NO_POS) newly_seen += (atomicMin(slot, pos0 + j) == GG_NO_POS);
        else atomicMin(slot, pos0 + j);
      }
    }
  }
  if (cand) bridge_chunk<MODE>(a, b, hi, lo, cand, newly_seen);
}

__global__ void __launch_bounds__(kThreads) count_kernel(Args a) {
  const int64_t n_tiles = (a.n_bytes + kTile - 1) / kTile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kThreads / 32;
  Walker wk(a, static_cast<int64_t>(blockIdx.x) * kWarps + warp, static_cast<int64_t>(gridDim.x) * kWarps, lane);
  int unused = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, wk.next())
    process_warp<kCount>(a, wk.b, wk.phase, lane, unused);
}

// Ordered scan with early exit: groups of tiles are handed out in stream order;
// once every non-zero bin has been seen no new group is started.  Groups already
// in flight always finish, and every group handed out later lies further along
// the stream, so the minima are final when the kernel ends.
__global__ void __launch_bounds__(kThreads) first_kernel(Args a) {
  __shared__ long long s_group;
  const int64_t n_tiles = (a.n_bytes + kTile - 1) / kTile;
  const int64_t n_groups = (n_tiles + kTilesPerGroup - 1) / kTilesPerGroup;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
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
      if (tile < n_tiles) {
        const int64_t b = chunk_offset(tile * (kThreads / 32) + warp, lane);
        const uint32_t phase = a.fixed_len ? static_cast<uint32_t>(b % a.fixed_len) : 0u;
        process_warp<kFirst>(a, b, phase, lane, newly_seen);
      }
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

// ---------------------------------------------------------------------------------------------
// Partitioned counting path (k <= 8) -- the default.
//
// Shared-memory atomics are ~10x faster than L2 atomics, but only 32,768 uint32 bins (128 KB) of the 4^(k+1)-bin
// table fit one SM beside the rest.  So the keys are first PARTITIONED by their top bits into n_buckets = bins /
// 32,768 streams of 15-bit keys (part_kernel: one pass over the ASCII bytes, each key is extracted once, grouped
// per 4096-position tile in shared memory and written out as aligned 16-byte groups; order inside a bucket does
// not matter to a histogram, so a bucket is just a bump-allocated array), then each bucket is counted with
// full-warp shared-memory atomics (count_bucket_kernel) and added to the global table once per CTA.  Every key costs
// 2 + 2 bytes of extra HBM traffic and exactly one shared-memory atomic; the sliced path above re-extracts every
// key once per slice and issues its atomics with 1/n_slices of the lanes active.
// For tables of <= 16,384 bins (k <= 6) there is one bucket and the partition is skipped: count_direct_kernel.
// ---------------------------------------------------------------------------------------------
constexpr int kLowBits = 15;
constexpr uint32_t kBucketBins = 1u << kLowBits;
constexpr int kMaxBuckets = 8;
constexpr int kPartThreads = 256;
constexpr int kPartWarps = kPartThreads / 32;
constexpr int kPartTile = kPartWarps * kWarpSpan;         // 3968 stream positions
constexpr int kStageKeys = kPartTile + 8 * kMaxBuckets;   // every bucket's run is padded to a multiple of 8 keys
constexpr uint16_t kPadKey = 0xFFFF;                      // bit 15 set: not a key

struct PartArgs {
  Args a;
  int bucket_bits;             // 1..3: the key's top bits select the bucket
  uint16_t* keys;              // [1 << bucket_bits][capacity]
  int64_t capacity;            // keys per bucket, multiple of 8
  unsigned long long* cursor;  // [kMaxBuckets] keys written per bucket (always a multiple of 8)
};

// bits 0, 2, 4, ... of x -> bits 0..15
__device__ __forceinline__ uint32_t compress_even_bits(uint32_t x) {
  x &= 0x55555555u;
  x = (x | (x >> 1)) & 0x33333333u;
  x = (x | (x >> 2)) & 0x0F0F0F0Fu;
  x = (x | (x >> 4)) & 0x00FF00FFu;
  x = (x | (x >> 8)) & 0x0000FFFFu;
  return x;
}

// Inside the kernel a key's bucket is always named by the top THREE bits of its window ("virtual bucket" q, with the
// bits below bucket_bits forced to zero), so that the bookkeeping is the same for 2, 4 and 8 buckets.
__global__ void __launch_bounds__(kPartThreads) part_kernel(PartArgs p) {
  __shared__ __align__(16) uint16_t stage[kStageKeys + 8];  // [kStageKeys] = spare slot for invalid positions
  __shared__ uint32_t cur[kMaxBuckets][kPartThreads];               // staging slot of this thread's next key of bucket q
  __shared__ __align__(16) uint16_t warp_tot[kPartWarps][kMaxBuckets];
  __shared__ __align__(16) uint16_t warp_off[kPartWarps][kMaxBuckets];  // first staging slot of (warp, q)
  __shared__ uint32_t bkt_start[kMaxBuckets + 1];                   // first staging slot of q; [8] = slots used
  __shared__ unsigned long long bkt_gbase[kMaxBuckets];             // reserved offset in the bucket's global array
  const Args& a = p.a;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int key_shift = 32 - 2 * (a.k + 1);
  const int vshift = 3 - p.bucket_bits;
  const uint32_t vmask = (7u >> vshift) << vshift;
  const uint32_t plane1 = p.bucket_bits >= 2 ? 0xFFFFu : 0u, plane0 = p.bucket_bits >= 3 ? 0xFFFFu : 0u;
  const int64_t n_tiles = (a.n_bytes + kPartTile - 1) / kPartTile;
  Walker wk(a, static_cast<int64_t>(blockIdx.x) * kPartWarps + warp, static_cast<int64_t>(gridDim.x) * kPartWarps, lane);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, wk.next()) {
    const int64_t b = wk.b;
    uint32_t hi, lo, pair_ok, cand;
    prepare_warp<true>(a, b, wk.phase, lane, hi, lo, pair_ok, cand);
    // bit planes of the windows' top three bits over the 16 owned positions (bit j = position j): base j holds bits
    // 31-2j and 30-2j of hi; the third bit is the top bit of the next base
    const uint32_t t2 = __brev(compress_even_bits(hi >> 1)) >> 16;
    const uint32_t t1 = (__brev(compress_even_bits(hi)) >> 16) & plane1;
    const uint32_t t0 = ((t2 >> 1) | ((lo >> 31) << 15)) & plane0;
    const uint32_t a1 = pair_ok & t2, a0 = pair_ok & ~t2;
    const uint32_t b3 = a1 & t1, b2 = a1 & ~t1, b1 = a0 & t1, b0 = a0 & ~t1;
    // keys of this thread per virtual bucket, two 16-bit fields per register: r[i] = (bucket 2i, bucket 2i+1)
    uint32_t r0 = __popc(b0 & ~t0) | (__popc(b0 & t0) << 16), r1 = __popc(b1 & ~t0) | (__popc(b1 & t0) << 16);
    uint32_t r2 = __popc(b2 & ~t0) | (__popc(b2 & t0) << 16), r3 = __popc(b3 & ~t0) | (__popc(b3 & t0) << 16);
    const uint32_t o0 = r0, o1 = r1, o2 = r2, o3 = r3;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {  // inclusive warp scan (a warp holds at most 496 keys)
      const uint32_t u0 = __shfl_up_sync(kFullWarp, r0, o), u1 = __shfl_up_sync(kFullWarp, r1, o);
      const uint32_t u2 = __shfl_up_sync(kFullWarp, r2, o), u3 = __shfl_up_sync(kFullWarp, r3, o);
      if (lane >= o) {
        r0 += u0;
        r1 += u1;
        r2 += u2;
        r3 += u3;
      }
    }
    if (lane == 31) *reinterpret_cast<uint4*>(&warp_tot[warp][0]) = make_uint4(r0, r1, r2, r3);
    __syncthreads();
    if (threadIdx.x < kMaxBuckets) {
      const int q = threadIdx.x;
      uint32_t run = 0, pre[kPartWarps];
#pragma unroll
      for (int w = 0; w < kPartWarps; ++w) {
        pre[w] = run;
        run += warp_tot[w][q];
      }
      const uint32_t padded = (run + 7u) & ~7u;
      uint32_t incl = padded;
#pragma unroll
      for (int o = 1; o < kMaxBuckets; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFu, incl, o);
        if (q >= o) incl += t;
      }
      const uint32_t start = incl - padded;
      bkt_start[q] = start;
      if (q == kMaxBuckets - 1) bkt_start[kMaxBuckets] = incl;
      for (uint32_t i = run; i < padded; ++i) stage[start + i] = kPadKey;
#pragma unroll
      for (int w = 0; w < kPartWarps; ++w) warp_off[w][q] = static_cast<uint16_t>(start + pre[w]);
      bkt_gbase[q] = padded ? atomicAdd(p.cursor + (q >> vshift), static_cast<unsigned long long>(padded)) : 0ull;
    }
    __syncthreads();
    {  // this thread's first staging slot per bucket = warp's slot + keys of earlier lanes (packed 16-bit adds)
      const uint4 wo = *reinterpret_cast<const uint4*>(&warp_off[warp][0]);
      const uint32_t c0 = wo.x + (r0 - o0), c1 = wo.y + (r1 - o1), c2 = wo.z + (r2 - o2), c3 = wo.w + (r3 - o3);
      cur[0][threadIdx.x] = c0 & 0xFFFFu; cur[1][threadIdx.x] = c0 >> 16;
      cur[2][threadIdx.x] = c1 & 0xFFFFu; cur[3][threadIdx.x] = c1 >> 16;
      cur[4][threadIdx.x] = c2 & 0xFFFFu; cur[5][threadIdx.x] = c2 >> 16;
      cur[6][threadIdx.x] = c3 & 0xFFFFu; cur[7][threadIdx.x] = c3 >> 16;
    }
#pragma unroll
    for (int j = 0; j < kPosPerThread; ++j) {  // branch-free: an invalid position bumps nothing and writes the spare slot
      const uint32_t win = __funnelshift_l(lo, hi, 2 * j);
      const uint32_t valid = (pair_ok >> j) & 1u;
      uint32_t* c = &cur[(win >> 29) & vmask][threadIdx.x];
      const uint32_t slot = *c;
      *c = slot + valid;
      stage[valid ? slot : static_cast<uint32_t>(kStageKeys)] = static_cast<uint16_t>((win >> key_shift) & (kBucketBins - 1u));
    }
    __syncthreads();
    {  // copy-out: the warp(s) of virtual bucket q stream its run (a multiple of 8 keys) as aligned 16-byte groups; 512-thread CTAs were measured slower (0.40 vs 0.35 ms)
      static_assert(kPartWarps % kMaxBuckets == 0, "whole warps per virtual bucket");
      constexpr int kWarpsPerBucket = kPartWarps / kMaxBuckets;
      const int q = warp / kWarpsPerBucket, sub = warp % kWarpsPerBucket;
      const uint32_t first = bkt_start[q], n_groups = (bkt_start[q + 1] - first) >> 3;
      const unsigned long long g0 = bkt_gbase[q];
      uint16_t* dst = p.keys + static_cast<int64_t>(q >> vshift) * p.capacity;
      for (uint32_t g = sub * 32 + lane; g < n_groups; g += 32 * kWarpsPerBucket) {
        const unsigned long long off = g0 + (g << 3);
        if (off + 8 <= static_cast<unsigned long long>(p.capacity))
          *reinterpret_cast<uint4*>(dst + off) = *reinterpret_cast<const uint4*>(stage + first + (g << 3));
      }
    }
    if (cand) {
      int unused = 0;
      bridge_chunk<kCount>(a, b, hi, lo, cand, unused);
    }
    __syncthreads();  // stage / warp_tot are reused by the next tile
  }
}

constexpr int kBucketThreads = 1024;

__global__ void __launch_bounds__(kBucketThreads, 1) count_bucket_kernel(const uint16_t* keys, int64_t capacity,
                                                                        const unsigned long long* cursor,
                                                                        int n_buckets, uint32_t* hist) {
  extern __shared__ uint32_t h[];
  const int bucket = blockIdx.x % n_buckets;
  const int slab = blockIdx.x / n_buckets, n_slabs = gridDim.x / n_buckets;
  if (slab >= n_slabs) return;
  unsigned long long filled = cursor[bucket];
  if (filled > static_cast<unsigned long long>(capacity)) filled = static_cast<unsigned long long>(capacity);
  const int64_t n_groups = static_cast<int64_t>(filled >> 3);
  const int64_t g_begin = n_groups * slab / n_slabs, g_end = n_groups * (slab + 1) / n_slabs;
  if (g_begin >= g_end) return;
  for (uint32_t i = threadIdx.x; i < kBucketBins; i += kBucketThreads) h[i] = 0;
  __syncthreads();
  const uint4* src = reinterpret_cast<const uint4*>(keys + static_cast<int64_t>(bucket) * capacity);
  constexpr int kBatch = 4;  // all loads of a batch are issued before any key is counted
  for (int64_t g0 = g_begin + threadIdx.x; g0 < g_end; g0 += static_cast<int64_t>(kBucketThreads) * kBatch) {
    uint4 v[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int64_t g = g0 + static_cast<int64_t>(u) * kBucketThreads;
      v[u] = g < g_end ? __ldcs(src + g) : make_uint4(~0u, ~0u, ~0u, ~0u);
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint32_t k0 = w[t] & 0xFFFFu, k1 = w[t] >> 16;
        if (k0 < kBucketBins) atomicAdd(&h[k0], 1u);
        if (k1 < kBucketBins) atomicAdd(&h[k1], 1u);
      }
    }
  }
  __syncthreads();
  uint32_t* out = hist + static_cast<int64_t>(bucket) * kBucketBins;
  for (uint32_t i = threadIdx.x; i < kBucketBins; i += kBucketThreads) {
    const uint32_t c = h[i];
    if (c) atomicAdd(out + i, c);
  }
}

// tables that fit one SM whole (k <= 6): extract and count in one kernel
__global__ void __launch_bounds__(kBucketThreads, 1) count_direct_kernel(Args a, uint32_t bins) {
  extern __shared__ uint32_t h[];
  for (uint32_t i = threadIdx.x; i < bins; i += kBucketThreads) h[i] = 0;
  __syncthreads();
  constexpr int kWarps = kBucketThreads / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int key_shift = 32 - 2 * (a.k + 1);
  const int64_t n_tiles = (a.n_bytes + kWarps * kWarpSpan - 1) / (kWarps * kWarpSpan);
  const uint32_t h_addr = static_cast<uint32_t>(__cvta_generic_to_shared(h));
  Walker wk(a, static_cast<int64_t>(blockIdx.x) * kWarps + warp, static_cast<int64_t>(gridDim.x) * kWarps, lane);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, wk.next()) {
    const int64_t b = wk.b;
    uint32_t hi, lo, pair_ok, cand;
    prepare_warp<true>(a, b, wk.phase, lane, hi, lo, pair_ok, cand);
#pragma unroll
    for (int j = 0; j < kPosPerThread; ++j)
      red_shared_inc_if(h_addr + 4u * (__funnelshift_l(lo, hi, 2 * j) >> key_shift), (pair_ok >> j) & 1u);
    if (cand) {
      int unused = 0;
      bridge_chunk<kCount>(a, b, hi, lo, cand, unused);
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < bins; i += kBucketThreads) {
    const uint32_t c = h[i];
    if (c) atomicAdd(a.hist + i, c);
  }
}

// Multi-GPU: every rank's (hist, first_pos) piece has been all-gathered; sum the histograms and take the earliest
// first occurrence (GG_NO_POS is the largest unsigned value).
__global__ void __launch_bounds__(256) merge_kernel(const unsigned long long* pieces, int world, long long stride,
                                                    long long bins, uint32_t* hist_out, unsigned long long* first_out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < bins;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint32_t sum = 0;
    unsigned long long first = GG_NO_POS;
    for (int r = 0; r < world; ++r) {
      const unsigned long long* piece = pieces + r * stride;
      sum += reinterpret_cast<const uint32_t*>(piece)[i];
      const unsigned long long f = piece[bins / 2 + i];
      first = f < first ? f : first;
    }
    hist_out[i] = sum;
    first_out[i] = first;
  }
}

int check_stream_args(const uint8_t* stream, int64_t n_bytes, int64_t fixed_len, int k) {
  GG_REQUIRE(k >= 1 && k <= GG_MAX_K, GG_ERR_K, "k outside [1, GG_MAX_K]");
  GG_REQUIRE(n_bytes >= 0 && fixed_len >= 0 && fixed_len < (1ll << 31), GG_ERR_ARG, "negative size / row too long");
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

static int part_buckets(int k) {
  const int bits = 2 * (k + 1);
  return bits <= kLowBits ? 1 : (bits - kLowBits > 3 ? 0 : 1 << (bits - kLowBits));
}

static int64_t part_capacity(int64_t n_bytes) {
  const int64_t n_tiles = (n_bytes + kPartTile - 1) / kPartTile;
  return (n_bytes + 8 * n_tiles + 7) / 8 * 8;  // every key in one bucket + that bucket's per-tile padding
}

extern "C" size_t gg_count_part_workspace_bytes(int64_t n_bytes, int k) {
  if (n_bytes < 0 || k < 1 || k > GG_MAX_K) return 0;
  const int nb = part_buckets(k);
  if (nb == 0) return 0;  // table too large for this path: use gg_count_kp1mers
  gg::Carver c(nullptr);
  c.take<unsigned long long>(kMaxBuckets);
  if (nb > 1) c.take<uint16_t>(static_cast<size_t>(nb) * part_capacity(n_bytes));
  return c.used();
}

extern "C" int gg_count_kp1mers_part(const uint8_t* stream, int64_t n_bytes, int64_t fixed_len, int k,
                                     int64_t pos_base, uint32_t* hist, gg_bridge_t* bridges, int64_t bridge_capacity,
                                     int64_t* status, void* workspace, size_t workspace_bytes, gg_stream_t st) {
  if (int rc = check_stream_args(stream, n_bytes, fixed_len, k)) return rc;
  GG_REQUIRE(hist && status, GG_ERR_ARG, "null buffer");
  GG_REQUIRE(bridge_capacity >= 0 && (bridge_capacity == 0 || bridges), GG_ERR_ARG, "bridge list");
  const size_t need = gg_count_part_workspace_bytes(n_bytes, k);
  GG_REQUIRE(need != 0, GG_ERR_K, "(k+1)-mer table too large for the partitioned path; use gg_count_kp1mers");
  GG_REQUIRE(workspace && workspace_bytes >= need, GG_ERR_WORKSPACE, "workspace too small");
  if (n_bytes == 0) return GG_OK;
  cudaStream_t s = gg::as_stream(st);
  Args a{stream, n_bytes, fixed_len, pos_base, k, hist, nullptr, bridges, bridge_capacity,
         reinterpret_cast<unsigned long long*>(status)};
  const int nb = part_buckets(k);
  const int sms = gg::sm_count();
  if (nb == 1) {
    const uint32_t bins = 1u << (2 * (k + 1));
    const size_t smem = static_cast<size_t>(bins) * sizeof(uint32_t);
    GG_CUDA_CHECK(cudaFuncSetAttribute(count_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
    const int64_t span = static_cast<int64_t>(kBucketThreads / 32) * kWarpSpan;
    const int64_t n_tiles = (n_bytes + span - 1) / span;
    count_direct_kernel<<<static_cast<int>(n_tiles < sms ? n_tiles : sms), kBucketThreads, smem, s>>>(a, bins);
    GG_CUDA_CHECK(cudaGetLastError());
    return GG_OK;
  }
  gg::Carver c(workspace);
  int bucket_bits = 0;
  while ((1 << bucket_bits) < nb) ++bucket_bits;
  PartArgs p{a, bucket_bits, nullptr, part_capacity(n_bytes), c.take<unsigned long long>(kMaxBuckets)};
  p.keys = c.take<uint16_t>(static_cast<size_t>(nb) * p.capacity);
  GG_CUDA_CHECK(cudaMemsetAsync(p.cursor, 0, kMaxBuckets * sizeof(unsigned long long), s));
  const int64_t n_tiles = (n_bytes + kPartTile - 1) / kPartTile;
  const int64_t resident = static_cast<int64_t>(sms) * (2048 / kPartThreads);
  part_kernel<<<static_cast<int>(n_tiles < resident ? n_tiles : resident), kPartThreads, 0, s>>>(p);
  GG_CUDA_CHECK(cudaGetLastError());
  const size_t smem = static_cast<size_t>(kBucketBins) * sizeof(uint32_t);
  GG_CUDA_CHECK(cudaFuncSetAttribute(count_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
  int grid = sms / nb * nb;
  if (grid < nb) grid = nb;
  count_bucket_kernel<<<grid, kBucketThreads, smem, s>>>(p.keys, p.capacity, p.cursor, nb, hist);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}

/* pieces: `world` buffers of piece_stride 64-bit words each, laid out [hist: bins uint32][first_pos: bins uint64][...] */
extern "C" int gg_count_merge(const uint64_t* pieces, int world, int64_t piece_stride, int k, uint32_t* hist_out,
                              uint64_t* first_out, gg_stream_t st) {
  GG_REQUIRE(k >= 1 && k <= GG_MAX_K, GG_ERR_K, "k outside [1, GG_MAX_K]");
  const long long bins = static_cast<long long>(gg_hist_bins(k));
  GG_REQUIRE(world >= 1 && piece_stride >= bins / 2 + bins, GG_ERR_ARG, "bad piece layout");
  GG_REQUIRE(pieces && hist_out && first_out, GG_ERR_ARG, "null buffer");
  long long grid = (bins + 255) / 256;
  const long long cap = static_cast<long long>(gg::sm_count()) * 8;
  if (grid > cap) grid = cap;
  merge_kernel<<<static_cast<unsigned>(grid), 256, 0, gg::as_stream(st)>>>(
      reinterpret_cast<const unsigned long long*>(pieces), world, piece_stride, bins, hist_out,
      reinterpret_cast<unsigned long long*>(first_out));
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}
