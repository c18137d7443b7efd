// K4' -- KF similarity edges for NARROW rows (small_k <= 2: rows of <= 16 counts) through duplicate-row classes.
//
// Replaces the scoring half of cosine_similarity_to_edge_index_weight (reference gnn_utils.py:240-322) for the
// small_k where many nodes share one feature row: at k=8, small_k=2 the 65,536 nodes have only 17,444 distinct
// rows (k=10: 95,924 of 1,048,576), so scoring every node pair repeats each dot product ~14x (k=10: ~120x).
// Here the rows are grouped into classes (radix sort of a mix of the 64-bit packed row), the CLASS pairs are scored once
// (dp4a, exact `dot >= dmin[sq_a * sq_b]` test, symmetric bit matrix), and every class then expands its
// neighbour classes into a shared-memory bitmap over node space from which its member rows count (pass 1) and
// write (pass 2) their edges directly in the reference's emission order -- block pair (I, J >= I), row, column --
// as the packed 8-byte words gg_kf_expand consumes.  Results are bit-identical to gg_kf_emit + gg_kf_place.
#include <cub/cub.cuh>

#include "gg_common.cuh"

namespace {

constexpr int kBlock = GG_KF_BLOCK;
constexpr int kSegWords = kBlock / 32;  // bitmap words per block of 1024 columns
constexpr int kThreads = 256;
constexpr int kMaxBitmapNodes = 1 << 20;  // 128 KB of shared memory (k <= 10)

// ---------------------------------------------------------------------------- classes
// Rows are grouped by sorting a 32-bit mix of the exact 64-bit row key (4 radix passes instead of 8) and cutting the
// sorted sequence wherever the EXACT keys of neighbours differ.  Two different rows that collide in the mix may
// interleave and split a class in two; that costs a little de-duplication, never correctness (every class still
// holds identical rows only, and equal-row classes are simply neighbours with cosine 1).
__device__ __forceinline__ uint32_t mix32(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return static_cast<uint32_t>(k);
}

__global__ void __launch_bounds__(kThreads) row_key_kernel(const uint8_t* counts, int pitch, int words, long long n,
                                                           unsigned long long* keys, uint32_t* hkeys, int32_t* idx,
                                                           unsigned long long* status) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const uint32_t* row = reinterpret_cast<const uint32_t*>(counts + i * pitch);
  unsigned long long key = 0;
  uint32_t over = 0, sq = 0;
  for (int w = 0; w < words; ++w) {
    const uint32_t v = __ldg(row + w);
    over |= v & 0xF0F0F0F0u;
    const uint32_t nib = (v & 0xFu) | ((v >> 4) & 0xF0u) | ((v >> 8) & 0xF00u) | ((v >> 12) & 0xF000u);
    key |= static_cast<unsigned long long>(nib) << (16 * w);
    sq = __dp4a(v & 0x0F0F0F0Fu, v & 0x0F0F0F0Fu, sq);
  }
  if (over) atomicAdd(status + 1, 1ull);  // a count above 15 does not fit the packed key: caller must not use this path
  keys[i] = key;
  // sort key: squared norm (<= 225 for rows this path accepts) above 24 bits of the mix -- classes of equal norm end up
  // next to each other, which lets the class-pair histogram count long runs of columns with one norm product
  hkeys[i] = ((sq > 255u ? 255u : sq) << 24) | (mix32(key) >> 8);
  idx[i] = static_cast<int32_t>(i);
}

__global__ void __launch_bounds__(kThreads) head_kernel(const unsigned long long* keys, const int32_t* sidx,
                                                        long long n, int32_t* head) {
  const long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (p < n) head[p] = (p == 0 || keys[sidx[p]] != keys[sidx[p - 1]]) ? 1 : 0;
}

__global__ void __launch_bounds__(kThreads) class_start_kernel(const int32_t* head, const int32_t* cid_incl,
                                                               long long n, int32_t* class_start,
                                                               unsigned long long* status) {
  const long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (p >= n) return;
  const int32_t c = cid_incl[p] - 1;
  if (head[p]) class_start[c] = static_cast<int32_t>(p);
  if (p == n - 1) {
    class_start[c + 1] = static_cast<int32_t>(n);
    status[0] = static_cast<unsigned long long>(c + 1);
  }
}

// ---------------------------------------------------------------------------- class-pair scoring
constexpr int kScoreThreads = 128;
constexpr int kScoreRows = 4;                             // row classes per thread (register blocking)
constexpr int kScoreA = kScoreThreads * kScoreRows;       // row classes per CTA
constexpr int kScoreB = 256;                              // column classes per shared-memory tile

struct ScoreArgs {
  const uint8_t* counts;
  int pitch, words;
  const int32_t* sqnorm;
  const int32_t* members;
  const int32_t* class_start;
  int n_classes;
  const uint16_t* dmin;
  int sq_max;
  uint32_t* adj;
  int adj_words;
  // HIST: class_hist[dot * (p_max + 1) + sq_a * sq_b] += number of NODE pairs s < t the passing class pair stands for
  // (|A| |B|, or |A| (|A| - 1) / 2 inside one class), accumulated for row classes [hist_begin, hist_end) only so that
  // ranks can split the class-pair triangle and sum their tables
  unsigned long long* class_hist;
  int p_max, hist_begin, hist_end;
  // TIES: node pairs of the cut-off class (dot^2 * tie_den == tie_num * P) whose smaller node lies at or beyond
  // row_lo / row_hi: tie_out[0], tie_out[1] (their difference = the members of the class this row shard emits)
  long long tie_num, tie_den;
  int row_lo, row_hi;
  unsigned long long* tie_out;
};

__device__ __forceinline__ void hist_flush_row(const ScoreArgs& a, uint32_t* acc, int r, unsigned long long p,
                                               unsigned long long wa) {
  for (int dot = 0; dot < 16; ++dot) {
    uint32_t* slot = acc + (dot * 4 + r) * 128 + threadIdx.x;   // [kHistDots][kScoreRows][kScoreThreads]
    const uint32_t v = *slot;
    if (v) {
      atomicAdd(a.class_hist + static_cast<unsigned long long>(dot) * (a.p_max + 1) + p, wa * v);
      *slot = 0u;
    }
  }
}

// members of class c with node id >= row (members ascend inside a class)
__device__ __forceinline__ int members_at_or_after(const int32_t* members, int m0, int m1, int row) {
  int lo = m0, hi = m1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(members + mid) < row) lo = mid + 1;
    else hi = mid;
  }
  return m1 - lo;
}

// words == 16 marks 16-byte rows at a 16-byte pitch in a 16-byte aligned matrix: one vector load
__device__ __forceinline__ uint4 load_row(const uint8_t* counts, int pitch, int words, int node) {
  if (words == 16) return __ldg(reinterpret_cast<const uint4*>(counts) + node);
  const uint32_t* r = reinterpret_cast<const uint32_t*>(counts + static_cast<size_t>(node) * pitch);
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  v.x = __ldg(r);
  if (words > 1) v.y = __ldg(r + 1);
  if (words > 2) v.z = __ldg(r + 2);
  if (words > 3) v.w = __ldg(r + 3);
  return v;
}

__device__ __forceinline__ uint32_t dot16(const uint4& a, const uint4& b) {
  return __dp4a(a.x, b.x, __dp4a(a.y, b.y, __dp4a(a.z, b.z, __dp4a(a.w, b.w, 0u))));
}

// Upper triangle of class pairs (b >= a, the diagonal included: two members of one class have cosine 1); a passing
// pair sets both adj[a][b] and adj[b][a].  Every thread keeps four row classes in registers and walks the column
// tile in shared memory (one broadcast 16-byte load per column class serves 4 x 128 pairs); the per-row threshold
// column thr[sq_b][row] makes the exact test one conflict-free shared-memory byte load, because all threads look at
// the same b and hence the same sq_b.
constexpr int kHistDots = 16;  // dots counted in the private table (larger ones are rare at small_k <= 2: direct atomics)

struct ScoreArgs;
__device__ __forceinline__ void hist_flush_row(const ScoreArgs& a, uint32_t* acc, int r, unsigned long long p,
                                               unsigned long long wa);

// EXTRA: 0 plain, 1 HIST, 2 TIES, 3 COUNT (see ScoreArgs; COUNT: tie_out[0] += node pairs of ALL passing class pairs)
template <int EXTRA>
__global__ void __launch_bounds__(kScoreThreads) score_kernel(ScoreArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint4* brow = reinterpret_cast<uint4*>(smem);
  uint32_t* boff = reinterpret_cast<uint32_t*>(smem + kScoreB * sizeof(uint4));  // sq_b * kScoreA
  uint8_t* thr = smem + kScoreB * (sizeof(uint4) + sizeof(uint32_t));            // [sq_max + 1][kScoreA]
  // EXTRA: per column class of the tile, its size (HIST) or its members at/after row_lo and row_hi (TIES)
  uint32_t* bw0 = reinterpret_cast<uint32_t*>(thr + static_cast<size_t>(a.sq_max + 1) * kScoreA);
  uint32_t* bw1 = bw0 + kScoreB;
  // HIST: private counters acc[dot][row][thread] of column-class sizes for the current run of columns with one squared
  // norm (classes are sorted by it): a run costs one atomic per (row, dot) instead of one per class pair
  uint32_t* acc = bw1 + kScoreB;
  uint32_t run_sq[kScoreRows] = {0u, 0u, 0u, 0u};
  if (EXTRA == 1) {
    for (int i = threadIdx.x; i < kHistDots * kScoreA; i += kScoreThreads) acc[i] = 0u;
  }
  const int a0 = blockIdx.y * kScoreA, b0 = blockIdx.x * kScoreB;
  if (b0 + kScoreB <= a0) return;  // every column class of this tile lies before every row class
  uint4 ra[kScoreRows];
  int sqa[kScoreRows];
  uint32_t aw0[kScoreRows], aw1[kScoreRows];
  unsigned long long tie0 = 0, tie1 = 0;
#pragma unroll
  for (int r = 0; r < kScoreRows; ++r) {
    const int ca = a0 + r * kScoreThreads + threadIdx.x;
    ra[r] = make_uint4(0u, 0u, 0u, 0u);
    sqa[r] = 0;
    aw0[r] = aw1[r] = 0;
    if (ca < a.n_classes) {
      const int m0 = __ldg(a.class_start + ca), m1 = __ldg(a.class_start + ca + 1);
      const int rep = __ldg(a.members + m0);
      ra[r] = load_row(a.counts, a.pitch, a.words, rep);
      sqa[r] = __ldg(a.sqnorm + rep);
      if (EXTRA == 1 || EXTRA == 3) aw0[r] = static_cast<uint32_t>(m1 - m0);
      if (EXTRA == 2) {
        aw0[r] = static_cast<uint32_t>(members_at_or_after(a.members, m0, m1, a.row_lo));
        aw1[r] = static_cast<uint32_t>(members_at_or_after(a.members, m0, m1, a.row_hi));
      }
    }
  }
  for (int sq = 0; sq <= a.sq_max; ++sq) {
#pragma unroll
    for (int r = 0; r < kScoreRows; ++r) {
      const uint32_t d = __ldg(a.dmin + sqa[r] * sq);  // padding rows: sq_a = 0 -> dmin[0] = never
      thr[sq * kScoreA + r * kScoreThreads + threadIdx.x] = static_cast<uint8_t>(d > 255u ? 255u : d);  // dot <= 225
    }
  }
  for (int t = threadIdx.x; t < kScoreB; t += kScoreThreads) {
    const int cb = b0 + t;
    uint4 rb = make_uint4(0u, 0u, 0u, 0u);
    int sqb = 0;
    uint32_t w0 = 0, w1 = 0;
    if (cb < a.n_classes) {
      const int m0 = __ldg(a.class_start + cb), m1 = __ldg(a.class_start + cb + 1);
      const int rep = __ldg(a.members + m0);
      rb = load_row(a.counts, a.pitch, a.words, rep);
      sqb = __ldg(a.sqnorm + rep);
      if (EXTRA == 1 || EXTRA == 3) w0 = static_cast<uint32_t>(m1 - m0);
      if (EXTRA == 2) {
        w0 = static_cast<uint32_t>(members_at_or_after(a.members, m0, m1, a.row_lo));
        w1 = static_cast<uint32_t>(members_at_or_after(a.members, m0, m1, a.row_hi));
      }
    }
    brow[t] = rb;
    boff[t] = static_cast<uint32_t>(sqb * kScoreA);
    if (EXTRA) {
      bw0[t] = w0;
      bw1[t] = w1;
    }
  }
  __syncthreads();
  const uint8_t* my_thr = thr + threadIdx.x;
  // hits are collected as one 32-bit mask per row and 32-column chunk and written once per chunk: a per-pair branch
  // would be taken by almost every warp iteration (128 pairs per warp and step, ~0.5 % of them hits)
  for (int t0 = 0; t0 < kScoreB; t0 += 32) {
    uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
#pragma unroll 8
    for (int tt = 0; tt < 32; ++tt) {
      const uint4 rb = brow[t0 + tt];
      const uint8_t* th = my_thr + boff[t0 + tt];
      const uint32_t d0 = dot16(ra[0], rb), d1 = dot16(ra[1], rb), d2 = dot16(ra[2], rb), d3 = dot16(ra[3], rb);
      m0 |= static_cast<uint32_t>(d0 >= th[0]) << tt;
      m1 |= static_cast<uint32_t>(d1 >= th[kScoreThreads]) << tt;
      m2 |= static_cast<uint32_t>(d2 >= th[2 * kScoreThreads]) << tt;
      m3 |= static_cast<uint32_t>(d3 >= th[3 * kScoreThreads]) << tt;
    }
    if (m0 | m1 | m2 | m3) {
      const int cb0 = b0 + t0;
      const uint32_t mm[kScoreRows] = {m0, m1, m2, m3};
#pragma unroll
      for (int r = 0; r < kScoreRows; ++r) {
        const int ca = a0 + r * kScoreThreads + threadIdx.x;
        uint32_t m = mm[r];
        // upper triangle only (cb >= ca): the mirrored bit is set below, from this side
        if (cb0 + 31 < ca) m = 0u;
        else if (cb0 < ca) m &= ~((1u << (ca - cb0)) - 1u);
        if (!m) continue;
        atomicOr(a.adj + static_cast<size_t>(ca) * a.adj_words + (cb0 >> 5), m);
        const bool count_row = EXTRA == 2 || (EXTRA == 1 && ca >= a.hist_begin && ca < a.hist_end);
        for (; m; m &= m - 1) {
          const int tb = t0 + __ffs(m) - 1, cb = b0 + tb;
          if (cb != ca) atomicOr(a.adj + static_cast<size_t>(cb) * a.adj_words + (ca >> 5), 1u << (ca & 31));
          if (EXTRA == 3) {
            const unsigned long long wa = aw0[r];
            tie0 += cb == ca ? wa * (wa - 1ull) / 2ull : wa * bw0[tb];
          } else if (EXTRA && count_row) {
            const unsigned long long dot = dot16(ra[r], brow[tb]);
            const unsigned long long p = static_cast<unsigned long long>(sqa[r]) * (boff[tb] / kScoreA);
            if (EXTRA == 1) {
              const unsigned long long wa = aw0[r];
              const uint32_t sqb = boff[tb] / kScoreA;
              if (cb == ca || dot >= kHistDots) {
                const unsigned long long w = cb == ca ? wa * (wa - 1ull) / 2ull : wa * bw0[tb];
                if (w) atomicAdd(a.class_hist + dot * static_cast<unsigned long long>(a.p_max + 1) + p, w);
              } else {
                if (sqb != run_sq[r]) {
                  hist_flush_row(a, acc, r, static_cast<unsigned long long>(sqa[r]) * run_sq[r], wa);
                  run_sq[r] = sqb;
                }
                acc[(static_cast<int>(dot) * kScoreRows + r) * kScoreThreads + threadIdx.x] += bw0[tb];
              }
            } else if (static_cast<long long>(dot * dot) * a.tie_den == a.tie_num * static_cast<long long>(p)) {
              const unsigned long long x0 = aw0[r], x1 = aw1[r];
              tie0 += cb == ca ? x0 * (x0 - 1ull) / 2ull : x0 * bw0[tb];
              tie1 += cb == ca ? x1 * (x1 - 1ull) / 2ull : x1 * bw1[tb];
            }
          }
        }
      }
    }
  }
  if (EXTRA == 1) {
#pragma unroll
    for (int r = 0; r < kScoreRows; ++r)
      hist_flush_row(a, acc, r, static_cast<unsigned long long>(sqa[r]) * run_sq[r], aw0[r]);
  }
  if (EXTRA == 2 || EXTRA == 3) {  // one pair of atomics per warp
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      tie0 += __shfl_xor_sync(0xFFFFFFFFu, tie0, o);
      tie1 += __shfl_xor_sync(0xFFFFFFFFu, tie1, o);
    }
    if ((threadIdx.x & 31) == 0 && (tie0 | tie1)) {
      atomicAdd(a.tie_out, tie0);
      atomicAdd(a.tie_out + 1, tie1);
    }
  }
}

// ---------------------------------------------------------------------------- per-class expansion
struct RowsArgs {
  const uint32_t* adj;
  int adj_words;
  const int32_t* members;
  const int32_t* class_start;
  int n_classes;
  int n_jblocks;
  int ib0, ib1;
  uint32_t* rowtot;                     // pass 1: candidates of (block pair, row), layout of gg_kf_rowcnt_elems
  const unsigned long long* offsets;    // pass 2: exclusive prefix sum of rowtot
  const uint8_t* counts;
  int pitch, words;
  uint2* ordered;
  long long capacity;
  const long long* skip_flag;  // pass 1, nullable: non-zero = the caller will not use the counts, do nothing
};

__host__ __device__ __forceinline__ long long rowcnt_word(int i_local, int n_jblocks, int j, int r) {
  return (static_cast<long long>(i_local) * n_jblocks + j) * kBlock + r;
}

// strict upper triangle inside a diagonal block: keep block-local columns > r of the 32-column chunk at c0
__device__ __forceinline__ uint32_t diag_mask(int r, int c0) {
  const int first = r + 1 - c0;
  return first <= 0 ? 0xFFFFFFFFu : (first >= 32 ? 0u : ~((1u << first) - 1u));
}

// The per-class kernels are latency-bound (dependent global loads, block barriers, phases in which few threads have
// work).  The counting pass does little per class, so classes in flight matter most: 128-thread CTAs (0.17 -> 0.15 ms);
// the placing pass has an edge copy-out that wants wide CTAs: 256 threads (128: 0.29 -> 0.33 ms).
// BIG: bitmaps of more than 64 KB (k >= 10) leave room for one CTA per SM only, which then wants all the threads it can
// get and wide windows (k=10: 16 windows of 64 blocks per class became 4 of 256; count 23 -> 10 ms, place 37 -> 18 ms).
template <bool PLACE, bool BIG>
struct RowsCfg {
  static constexpr int kRowThreads = BIG ? 1024 : (PLACE ? 256 : 128);
  static constexpr int kWinBlocks = BIG ? 256 : 64;       // column blocks per window
  static constexpr int kMemberChunk = PLACE || BIG ? 256 : 128;  // members of one class per sweep (shared-memory scratch)
  static constexpr int kListCap = PLACE || BIG ? 2048 : 1024;    // work-list entries (>= 1024: a column block always fits)
  static constexpr int kWordsPerThread = kWinBlocks * kSegWords / kRowThreads;
  static constexpr int kLanesPerBlock = kSegWords / kWordsPerThread;  // neighbouring lanes that share a column block
  static_assert(kWordsPerThread % 4 == 0 && kSegWords % kWordsPerThread == 0 && kWordsPerThread <= 32,
                "a thread's words lie inside one column block and load as uint4");
};
constexpr size_t kBigBitmapBytes = 64 << 10;

// One class per CTA iteration (classes are handed out through an atomic cursor: their sizes vary widely).
//   bm    bitmap over node space: bit j set iff node j belongs to a neighbour class of this class
//   cj    set bits per block of 1024 columns (pass 1)
//   list  first the neighbour classes (so that the bitmap is filled with all lanes busy); in pass 2 the candidate
//         columns of a window of column blocks in ascending order, (j | dot << 24, rank of j inside its block) --
//         all member rows of the class have the same candidates and the same dots, only their offsets differ
//   memi / dcnt / mbase  per member row i: node id (-1: outside the row shard); set bits of its own (diagonal) block
//         at columns <= i, i.e. the candidates the strict upper triangle drops -- all of them precede every kept bit
//         of that block, so a kept bit's rank is its rank inside the block minus dcnt; index of (block row, column
//         block 0, row) in `offsets`
template <bool PLACE, bool BIG>
__global__ void __launch_bounds__(RowsCfg<PLACE, BIG>::kRowThreads) class_rows_kernel(RowsArgs a, unsigned int* cursor) {
  using Cfg = RowsCfg<PLACE, BIG>;
  constexpr int kRowThreads = Cfg::kRowThreads, kMemberChunk = Cfg::kMemberChunk, kListCap = Cfg::kListCap,
                kWordsPerThread = Cfg::kWordsPerThread, kLanesPerBlock = Cfg::kLanesPerBlock,
                kWinBlocks = Cfg::kWinBlocks;
  extern __shared__ __align__(16) uint32_t sm[];
  __shared__ int s_class;
  __shared__ unsigned int s_n, s_warp[kRowThreads / 32];
  uint32_t* bm = sm;  // padded to whole windows
  const int bm_padded = (a.n_jblocks + kWinBlocks - 1) / kWinBlocks * kWinBlocks * kSegWords;
  uint32_t* cj = bm + bm_padded;
  uint2* list = reinterpret_cast<uint2*>(cj + ((a.n_jblocks + 3) & ~3));
  uint32_t* nlist = reinterpret_cast<uint32_t*>(list);  // neighbour classes, 2 * kListCap slots
  long long* mbase = reinterpret_cast<long long*>(list + kListCap);
  int* memi = reinterpret_cast<int*>(mbase + kMemberChunk);
  uint32_t* dcnt = reinterpret_cast<uint32_t*>(memi + kMemberChunk);
  constexpr unsigned int kNbCap = 2 * kListCap;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row_lo = a.ib0 * kBlock, row_hi = a.ib1 * kBlock;
  if (!PLACE && a.skip_flag != nullptr && *a.skip_flag != 0) return;  // uniform: nothing has been written yet
  while (true) {
    if (threadIdx.x == 0) {
      s_class = static_cast<int>(atomicAdd(cursor, 1u));
      s_n = 0u;
    }
    __syncthreads();
    const int c = s_class;
    if (c >= a.n_classes) break;
    const int m0 = __ldg(a.class_start + c), m1 = __ldg(a.class_start + c + 1);
    // members ascend: skip classes without a row in this shard (uniform across the CTA).  The first member at or
    // after row_lo decides: shards of a few row blocks (top-n chunks) hold a member of only a few percent of the classes
    if (a.ib0 > 0 || a.ib1 < a.n_jblocks) {  // a shard that holds every row needs no search (two dependent loads per class)
      const int first_in = m1 - members_at_or_after(a.members, m0, m1, row_lo);
      if (first_in >= m1 || __ldg(a.members + first_in) >= row_hi) {
        __syncthreads();  // s_class is rewritten at the top of the loop
        continue;
      }
    }
    for (int w = threadIdx.x; w < bm_padded / 4; w += kRowThreads) reinterpret_cast<uint4*>(bm)[w] = make_uint4(0u, 0u, 0u, 0u);
    // neighbour classes of this class: compact the set bits of its adjacency row into the work list ...
    const uint32_t* arow = a.adj + static_cast<size_t>(c) * a.adj_words;
    uint32_t spill[2] = {0u, 0u};  // a word whose bits did not fit the list (dense adjacency rows)
    int spill_w[2] = {0, 0}, n_spill = 0;
    for (int w = threadIdx.x; w < a.adj_words; w += kRowThreads) {
      uint32_t bits = __ldg(arow + w);
      if (!bits) continue;
      unsigned int pos = atomicAdd(&s_n, static_cast<unsigned int>(__popc(bits)));
      for (; bits && pos < kNbCap; bits &= bits - 1) nlist[pos++] = static_cast<uint32_t>(w * 32 + __ffs(bits) - 1);
      if (bits) {
        if (n_spill < 2) {
          spill[n_spill] = bits;
          spill_w[n_spill] = w;
          ++n_spill;
        } else {
          n_spill = 3;  // more than two spilled words per thread: handled by the slow sweep below
        }
      }
    }
    __syncthreads();  // the bitmap is zero and the list complete
    // ... and set the bits of their members
    {
      const unsigned int n_nb = s_n < kNbCap ? s_n : kNbCap;
      for (unsigned int e = threadIdx.x; e < n_nb; e += kRowThreads) {
        const int b = static_cast<int>(nlist[e]);
        const int end = __ldg(a.class_start + b + 1);
        for (int m = __ldg(a.class_start + b); m < end; ++m) {
          const int j = __ldg(a.members + m);
          atomicOr(bm + (j >> 5), 1u << (j & 31));
        }
      }
      if (n_spill == 3) {  // re-walk this thread's words: idempotent, bits already listed are simply set twice
        for (int w = threadIdx.x; w < a.adj_words; w += kRowThreads) {
          for (uint32_t bits = __ldg(arow + w); bits; bits &= bits - 1) {
            const int b = w * 32 + __ffs(bits) - 1;
            const int end = __ldg(a.class_start + b + 1);
            for (int m = __ldg(a.class_start + b); m < end; ++m) {
              const int j = __ldg(a.members + m);
              atomicOr(bm + (j >> 5), 1u << (j & 31));
            }
          }
        }
      } else {
        for (int q = 0; q < n_spill; ++q) {
          for (uint32_t bits = spill[q]; bits; bits &= bits - 1) {
            const int b = spill_w[q] * 32 + __ffs(bits) - 1;
            const int end = __ldg(a.class_start + b + 1);
            for (int m = __ldg(a.class_start + b); m < end; ++m) {
              const int j = __ldg(a.members + m);
              atomicOr(bm + (j >> 5), 1u << (j & 31));
            }
          }
        }
      }
    }
    __syncthreads();
    if (!PLACE) {
      // set bits per column block: a thread sums its consecutive words, neighbouring lanes make one block
      for (int w0 = 0; w0 < bm_padded; w0 += kRowThreads * kWordsPerThread) {
        const int w = w0 + threadIdx.x * kWordsPerThread;
        int t = 0;
#pragma unroll
        for (int q = 0; q < kWordsPerThread; q += 4) {
          const uint4 x = *reinterpret_cast<const uint4*>(bm + w + q);
          t += __popc(x.x) + __popc(x.y) + __popc(x.z) + __popc(x.w);
        }
#pragma unroll
        for (int o = 1; o < kLanesPerBlock; o <<= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
        const int jb = w >> 5;
        if ((lane & (kLanesPerBlock - 1)) == 0 && jb < a.n_jblocks) cj[jb] = static_cast<uint32_t>(t);
      }
    }
    const uint4 rc = load_row(a.counts, a.pitch, a.words, __ldg(a.members + m0));  // the row all members share
    for (int mc = m0; mc < m1; mc += kMemberChunk) {
      const int mcount = (m1 - mc) < kMemberChunk ? (m1 - mc) : kMemberChunk;
      for (int mi = warp; mi < mcount; mi += kRowThreads / 32) {  // one warp per member: lane <-> word of its block
        const int i = __ldg(a.members + mc + mi), ib = i >> 10, r = i & (kBlock - 1);
        const uint32_t t = __reduce_add_sync(0xFFFFFFFFu, __popc(bm[ib * kSegWords + lane] & ~diag_mask(r, lane * 32)));
        if (lane == 0) {
          dcnt[mi] = t;
          const bool mine = ib >= a.ib0 && ib < a.ib1;
          memi[mi] = mine ? i : -1;
          mbase[mi] = mine ? rowcnt_word(ib - a.ib0, a.n_jblocks, 0, r) : 0;
        }
      }
      __syncthreads();
      if (!PLACE) {
        const int n_tasks = mcount * a.n_jblocks;
        for (int t = threadIdx.x; t < n_tasks; t += kRowThreads) {
          const int mi = t / a.n_jblocks, jb = t - mi * a.n_jblocks;
          const int i = memi[mi];
          if (i < 0) continue;
          const int ib = i >> 10;
          if (jb < ib) continue;
          const uint32_t cnt = cj[jb] - (jb == ib ? dcnt[mi] : 0u);
          if (cnt) a.rowtot[mbase[mi] + static_cast<long long>(jb) * kBlock] = cnt;
        }
        __syncthreads();
      } else {
        // members ascend: column blocks before the first member's block hold no edge of this sweep
        for (int jw = (__ldg(a.members + mc) >> 10) / kWinBlocks * kWinBlocks; jw < a.n_jblocks; jw += kWinBlocks) {
          // window of 64 column blocks: every thread owns 8 or 16 consecutive bitmap words (all inside one block)
          const int w = jw * kSegWords + threadIdx.x * kWordsPerThread;
          int mine = 0;
          uint32_t nzq = 0;  // which of its words hold a bit at all (most do not)
#pragma unroll
          for (int q = 0; q < kWordsPerThread; q += 4) {
            const uint4 x = *reinterpret_cast<const uint4*>(bm + w + q);
            mine += __popc(x.x) + __popc(x.y) + __popc(x.z) + __popc(x.w);
            nzq |= (static_cast<uint32_t>(x.x != 0u) | (static_cast<uint32_t>(x.y != 0u) << 1) |
                    (static_cast<uint32_t>(x.z != 0u) << 2) | (static_cast<uint32_t>(x.w != 0u) << 3)) << q;
          }
          int incl = mine;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += v;
          }
          if (lane == 31) s_warp[warp] = static_cast<unsigned int>(incl);
          // rank inside the column block: the lanes of a block are neighbours
          const int excl_warp = incl - mine;
          const uint32_t rank0 = static_cast<uint32_t>(excl_warp - __shfl_sync(0xFFFFFFFFu, excl_warp, lane & ~(kLanesPerBlock - 1)));
          __syncthreads();
          unsigned int pre = static_cast<unsigned int>(excl_warp), total = 0;
#pragma unroll
          for (int q = 0; q < kRowThreads / 32; ++q) {
            const unsigned int t = s_warp[q];
            if (q < warp) pre += t;
            total += t;
          }
          // members whose block lies inside or before this window (ascending order: a prefix of the sweep)
          int n_mem = 0;
          for (int q0 = 0; q0 < mcount; q0 += kRowThreads) {
            const int mi = q0 + threadIdx.x;
            n_mem += __syncthreads_count(mi < mcount && (__ldg(a.members + mc + mi) >> 10) < jw + kWinBlocks);
          }
          for (unsigned int lo = 0; lo < total; lo += kListCap) {  // usually one pass: the window's candidates fit
            const unsigned int n_ent = total - lo < kListCap ? total - lo : kListCap;
            if (mine && pre < lo + n_ent && pre + mine > lo) {  // few lanes are busy here: keep the loop body tiny
              unsigned int idx = pre;
              uint32_t rank = rank0;
              for (uint32_t nz = nzq; nz; nz &= nz - 1) {  // ascending words, empty ones skipped
                const int wq = w + __ffs(nz) - 1;
                for (uint32_t word = bm[wq]; word; word &= word - 1, ++idx, ++rank)
                  if (idx >= lo && idx < lo + n_ent)
                    list[idx - lo] = make_uint2(static_cast<uint32_t>(wq * 32 + __ffs(word) - 1), rank);
              }
            }
            __syncthreads();
            for (unsigned int e = threadIdx.x; e < n_ent; e += kRowThreads)  // all lanes busy: the dots
              list[e].x |= dot16(rc, load_row(a.counts, a.pitch, a.words, static_cast<int>(list[e].x))) << 24;
            __syncthreads();
            // every thread takes one candidate column and walks the member rows (uniform loop, broadcast reads)
            for (unsigned int e = threadIdx.x; e < n_ent; e += kRowThreads) {
              const uint2 ent = list[e];
              const int j = static_cast<int>(ent.x & 0xFFFFFFu), jb = j >> 10;
              const uint32_t hi = ent.x & 0xFF000000u;
              const unsigned long long* off_j = a.offsets + static_cast<long long>(jb) * kBlock;
              for (int mi = 0; mi < n_mem; ++mi) {
                const int i = memi[mi];
                if (i < 0 || j <= i) continue;
                const unsigned long long pos = __ldg(off_j + mbase[mi]) + ent.y - (jb == (i >> 10) ? dcnt[mi] : 0u);
                if (static_cast<long long>(pos) < a.capacity)
                  a.ordered[pos] = make_uint2(static_cast<uint32_t>(i) | hi, static_cast<uint32_t>(j));
              }
            }
            __syncthreads();
          }
        }
      }
    }
  }
}

struct ToU64 {
  __host__ __device__ unsigned long long operator()(uint32_t v) const { return v; }
};

int check_rows(int64_t n, int d_valid, int pitch) {
  GG_REQUIRE(n >= 0 && n <= kMaxBitmapNodes, GG_ERR_ARG, "n outside [0, 2^20] for the duplicate-row path");
  GG_REQUIRE(d_valid >= 4 && d_valid <= 16 && d_valid % 4 == 0, GG_ERR_ARG, "duplicate-row path needs 4..16 columns");
  GG_REQUIRE(pitch >= d_valid && pitch % 4 == 0, GG_ERR_ARG, "pitch must be a multiple of 4 and >= d_valid");
  return GG_OK;
}

int row_words(const uint8_t* counts, int pitch, int d_valid) {
  return (d_valid == 16 && pitch == 16 && (reinterpret_cast<uintptr_t>(counts) & 15) == 0) ? 16 : d_valid / 4;
}

size_t sort_temp_bytes(long long n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, static_cast<const uint32_t*>(nullptr),
                                  static_cast<uint32_t*>(nullptr), static_cast<const int32_t*>(nullptr),
                                  static_cast<int32_t*>(nullptr), n);
  return bytes;
}
size_t isum_temp_bytes(long long n) {
  size_t bytes = 0;
  cub::DeviceScan::InclusiveSum(nullptr, bytes, static_cast<const int32_t*>(nullptr), static_cast<int32_t*>(nullptr), n);
  return bytes;
}
size_t esum_temp_bytes(long long n) {
  size_t bytes = 0;
  cub::TransformInputIterator<unsigned long long, ToU64, const uint32_t*> it(nullptr, ToU64());
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, static_cast<unsigned long long*>(nullptr), n);
  return bytes;
}

int rows_grid(int n_classes, size_t smem);

bool rows_big(int n_jblocks) { return static_cast<size_t>(n_jblocks) * kSegWords * sizeof(uint32_t) > kBigBitmapBytes; }

template <bool PLACE, bool BIG>
size_t rows_smem_of(int n_jblocks) {
  using Cfg = RowsCfg<PLACE, BIG>;
  const size_t padded = (static_cast<size_t>(n_jblocks) + Cfg::kWinBlocks - 1) / Cfg::kWinBlocks * Cfg::kWinBlocks * kSegWords;
  return (padded + ((static_cast<size_t>(n_jblocks) + 3) & ~static_cast<size_t>(3))) * sizeof(uint32_t) +
         Cfg::kListCap * sizeof(uint2) + Cfg::kMemberChunk * (sizeof(long long) + sizeof(int) + sizeof(uint32_t));
}

template <bool PLACE, bool BIG>
int launch_rows(const RowsArgs& a, unsigned int* cursor, cudaStream_t s) {
  const size_t smem = rows_smem_of<PLACE, BIG>(a.n_jblocks);
  GG_CUDA_CHECK(cudaFuncSetAttribute(class_rows_kernel<PLACE, BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
  class_rows_kernel<PLACE, BIG><<<rows_grid(a.n_classes, smem), RowsCfg<PLACE, BIG>::kRowThreads, smem, s>>>(a, cursor);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}

int rows_grid(int n_classes, size_t smem) {
  size_t per_sm = (220u << 10) / (smem + 1024);
  if (per_sm > 12) per_sm = 12;
  if (per_sm < 1) per_sm = 1;
  const int cap = gg::sm_count() * static_cast<int>(per_sm);
  return n_classes < cap ? (n_classes < 1 ? 1 : n_classes) : cap;
}

}  // namespace

extern "C" size_t gg_kf_dedup_classes_workspace_bytes(int64_t n) {
  if (n <= 0) return 256;
  gg::Carver c(nullptr);
  c.take<unsigned long long>(n);
  c.take<uint32_t>(n);
  c.take<uint32_t>(n);
  c.take<int32_t>(n);
  c.take<int32_t>(n);
  c.take<int32_t>(n);
  const size_t a = sort_temp_bytes(n), b = isum_temp_bytes(n);
  c.take<char>(a > b ? a : b);
  return c.used();
}

extern "C" int gg_kf_dedup_classes(const uint8_t* counts, int pitch, int d_valid, int64_t n, int32_t* members,
                                   int32_t* class_start, int64_t* dd_status, void* workspace, size_t workspace_bytes,
                                   gg_stream_t st) {
  if (int rc = check_rows(n, d_valid, pitch)) return rc;
  GG_REQUIRE(dd_status, GG_ERR_ARG, "null status");
  cudaStream_t s = gg::as_stream(st);
  GG_CUDA_CHECK(cudaMemsetAsync(dd_status, 0, 2 * sizeof(int64_t), s));
  if (n == 0) return GG_OK;
  GG_REQUIRE(counts && members && class_start && workspace, GG_ERR_ARG, "null buffer");
  GG_REQUIRE((reinterpret_cast<uintptr_t>(counts) & 3) == 0, GG_ERR_ARG, "counts must be 4-byte aligned");
  GG_REQUIRE(workspace_bytes >= gg_kf_dedup_classes_workspace_bytes(n), GG_ERR_WORKSPACE, "workspace too small");
  gg::Carver c(workspace);
  unsigned long long* keys = c.take<unsigned long long>(n);
  uint32_t* hkeys = c.take<uint32_t>(n);
  uint32_t* shkeys = c.take<uint32_t>(n);
  int32_t* idx = c.take<int32_t>(n);
  int32_t* head = c.take<int32_t>(n);
  int32_t* cid = c.take<int32_t>(n);
  const size_t ta = sort_temp_bytes(n), tb = isum_temp_bytes(n);
  size_t temp_bytes = ta > tb ? ta : tb;
  void* temp = c.take<char>(temp_bytes);
  unsigned long long* stw = reinterpret_cast<unsigned long long*>(dd_status);
  const unsigned grid = static_cast<unsigned>((n + kThreads - 1) / kThreads);
  row_key_kernel<<<grid, kThreads, 0, s>>>(counts, pitch, d_valid / 4, n, keys, hkeys, idx, stw);
  GG_CUDA_CHECK(cudaGetLastError());
  // stable LSD radix sort: members of one class stay in ascending node order
  GG_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, hkeys, shkeys, idx, members, static_cast<int>(n), 0,
                                                32, s));
  head_kernel<<<grid, kThreads, 0, s>>>(keys, members, n, head);
  GG_CUDA_CHECK(cudaGetLastError());
  temp_bytes = ta > tb ? ta : tb;
  GG_CUDA_CHECK(cub::DeviceScan::InclusiveSum(temp, temp_bytes, head, cid, static_cast<int>(n), s));
  class_start_kernel<<<grid, kThreads, 0, s>>>(head, cid, n, class_start, stw);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}

extern "C" size_t gg_kf_dedup_adj_words(int64_t n_classes) {
  if (n_classes <= 0) return 0;
  return static_cast<size_t>(n_classes) * static_cast<size_t>((n_classes + 31) / 32);
}

extern "C" int gg_kf_dedup_score(const uint8_t* counts, int pitch, int d_valid, const int32_t* sqnorm,
                                 const int32_t* members, const int32_t* class_start, int64_t n_classes,
                                 const uint16_t* dmin, int32_t p_max, uint32_t* adj, const gg_kf_dedup_extra_t* extra,
                                 gg_stream_t st) {
  GG_REQUIRE(n_classes >= 0 && n_classes <= kMaxBitmapNodes, GG_ERR_ARG, "bad class count");
  GG_REQUIRE(d_valid >= 4 && d_valid <= 16 && d_valid % 4 == 0 && pitch >= d_valid && pitch % 4 == 0, GG_ERR_ARG,
             "duplicate-row path needs 4..16 columns");
  if (n_classes == 0) return GG_OK;
  GG_REQUIRE(counts && sqnorm && members && class_start && dmin && adj, GG_ERR_ARG, "null buffer");
  int sq_max = 0;  // p_max = sq_max^2 by construction (core.threshold_tables)
  while ((sq_max + 1) * (sq_max + 1) <= p_max) ++sq_max;
  GG_REQUIRE(sq_max * sq_max == p_max && sq_max <= 225, GG_ERR_ARG, "p_max must be a square <= 225^2");
  cudaStream_t s = gg::as_stream(st);
  const int adj_words = static_cast<int>((n_classes + 31) / 32);
  GG_CUDA_CHECK(cudaMemsetAsync(adj, 0, gg_kf_dedup_adj_words(n_classes) * sizeof(uint32_t), s));
  ScoreArgs a{};
  a.counts = counts;
  a.pitch = pitch;
  a.words = row_words(counts, pitch, d_valid);
  a.sqnorm = sqnorm;
  a.members = members;
  a.class_start = class_start;
  a.n_classes = static_cast<int>(n_classes);
  a.dmin = dmin;
  a.sq_max = sq_max;
  a.adj = adj;
  a.adj_words = adj_words;
  a.p_max = p_max;
  int mode = 0;
  if (extra != nullptr && extra->class_hist != nullptr) {
    GG_REQUIRE(extra->tie_out == nullptr, GG_ERR_ARG, "class_hist and tie_out are separate passes");
    GG_REQUIRE(extra->hist_class_begin >= 0 && extra->hist_class_begin <= extra->hist_class_end, GG_ERR_ARG,
               "bad class range for the histogram");
    mode = 1;
    a.class_hist = reinterpret_cast<unsigned long long*>(extra->class_hist);
    a.hist_begin = static_cast<int>(extra->hist_class_begin);
    a.hist_end = static_cast<int>(extra->hist_class_end < n_classes ? extra->hist_class_end : n_classes);
  } else if (extra != nullptr && extra->tie_out != nullptr && extra->tie_num == 0 && extra->tie_den == 0) {
    mode = 3;  // COUNT
    a.tie_out = reinterpret_cast<unsigned long long*>(extra->tie_out);
  } else if (extra != nullptr && extra->tie_out != nullptr) {
    GG_REQUIRE(extra->tie_num > 0 && extra->tie_den > 0 && extra->tie_num <= (1 << 16) && extra->tie_den <= (1ll << 32),
               GG_ERR_ARG, "cut-off class must be dot^2 / P with dot <= 255");
    GG_REQUIRE(extra->row_lo >= 0 && extra->row_lo <= extra->row_hi, GG_ERR_ARG, "bad row shard");
    mode = 2;
    a.tie_num = extra->tie_num;
    a.tie_den = extra->tie_den;
    a.row_lo = static_cast<int>(extra->row_lo);
    a.row_hi = static_cast<int>(extra->row_hi);
    a.tie_out = reinterpret_cast<unsigned long long*>(extra->tie_out);
  }
  const size_t smem = kScoreB * (sizeof(uint4) + sizeof(uint32_t)) + static_cast<size_t>(sq_max + 1) * kScoreA +
                      2 * kScoreB * sizeof(uint32_t) + (mode == 1 ? kHistDots * kScoreA * sizeof(uint32_t) : 0);
  auto kernel = mode == 0 ? score_kernel<0> : (mode == 1 ? score_kernel<1> : (mode == 2 ? score_kernel<2> : score_kernel<3>));
  GG_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  dim3 grid(static_cast<unsigned>((n_classes + kScoreB - 1) / kScoreB),
            static_cast<unsigned>((n_classes + kScoreA - 1) / kScoreA));
  kernel<<<grid, kScoreThreads, smem, s>>>(a);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}

static size_t rowcnt_elems(int64_t n, int ib0, int ib1) {
  if (n <= 0 || ib1 <= ib0) return 0;
  return static_cast<size_t>(ib1 - ib0) * static_cast<size_t>((n + kBlock - 1) / kBlock) * kBlock;
}

extern "C" size_t gg_kf_dedup_count_workspace_bytes(int64_t n, int iblock_begin, int iblock_end) {
  const size_t elems = rowcnt_elems(n, iblock_begin, iblock_end) + 1;
  gg::Carver c(nullptr);
  c.take<uint32_t>(elems);
  c.take<unsigned int>(4);  // class cursors of the two passes
  c.take<char>(esum_temp_bytes(static_cast<long long>(elems)));
  return c.used();
}

static int fill_rows_args(RowsArgs& a, int64_t n, int ib0, int ib1, const uint32_t* adj, const int32_t* members,
                          const int32_t* class_start, int64_t n_classes) {
  GG_REQUIRE(n >= 0 && n <= kMaxBitmapNodes, GG_ERR_ARG, "n outside [0, 2^20] for the duplicate-row path");
  const int nb = static_cast<int>((n + kBlock - 1) / kBlock);
  GG_REQUIRE(ib0 >= 0 && ib0 <= ib1 && ib1 <= nb, GG_ERR_ARG, "row-block shard outside [0, ceil(n/1024)]");
  GG_REQUIRE(n_classes >= 0 && n_classes <= n, GG_ERR_ARG, "bad class count");
  a.adj = adj;
  a.adj_words = static_cast<int>((n_classes + 31) / 32);
  a.members = members;
  a.class_start = class_start;
  a.n_classes = static_cast<int>(n_classes);
  a.n_jblocks = nb;
  a.ib0 = ib0;
  a.ib1 = ib1;
  return GG_OK;
}

/* offsets: [gg_kf_rowcnt_elems(n, ib0, ib1) + 1]; the last entry is the number of edges of the shard */
extern "C" int gg_kf_dedup_count(int64_t n, int iblock_begin, int iblock_end, const uint32_t* adj,
                                 const int32_t* members, const int32_t* class_start, int64_t n_classes,
                                 uint64_t* offsets, const int64_t* skip_flag, void* workspace, size_t workspace_bytes,
                                 gg_stream_t st) {
  RowsArgs a{};
  a.skip_flag = reinterpret_cast<const long long*>(skip_flag);
  if (int rc = fill_rows_args(a, n, iblock_begin, iblock_end, adj, members, class_start, n_classes)) return rc;
  GG_REQUIRE(offsets && workspace, GG_ERR_ARG, "null buffer");
  GG_REQUIRE(workspace_bytes >= gg_kf_dedup_count_workspace_bytes(n, iblock_begin, iblock_end), GG_ERR_WORKSPACE,
             "workspace too small");
  cudaStream_t s = gg::as_stream(st);
  const size_t elems = rowcnt_elems(n, iblock_begin, iblock_end) + 1;
  gg::Carver c(workspace);
  a.rowtot = c.take<uint32_t>(elems);
  unsigned int* cursor = c.take<unsigned int>(4);
  size_t temp_bytes = esum_temp_bytes(static_cast<long long>(elems));
  void* temp = c.take<char>(temp_bytes);
  GG_CUDA_CHECK(cudaMemsetAsync(a.rowtot, 0, elems * sizeof(uint32_t), s));
  if (n_classes > 0 && elems > 1) {
    GG_REQUIRE(adj && members && class_start, GG_ERR_ARG, "null buffer");
    GG_CUDA_CHECK(cudaMemsetAsync(cursor, 0, 4 * sizeof(unsigned int), s));
    if (int rc = rows_big(a.n_jblocks) ? launch_rows<false, true>(a, cursor, s) : launch_rows<false, false>(a, cursor, s))
      return rc;
  }
  cub::TransformInputIterator<unsigned long long, ToU64, const uint32_t*> it(a.rowtot, ToU64());
  GG_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, it, reinterpret_cast<unsigned long long*>(offsets),
                                              static_cast<long long>(elems), s));
  return GG_OK;
}

extern "C" int gg_kf_dedup_place(int64_t n, int iblock_begin, int iblock_end, const uint32_t* adj,
                                 const int32_t* members, const int32_t* class_start, int64_t n_classes,
                                 const uint64_t* offsets, const uint8_t* counts, int pitch, int d_valid,
                                 uint64_t* ordered, int64_t capacity, void* workspace, size_t workspace_bytes,
                                 gg_stream_t st) {
  RowsArgs a{};
  if (int rc = fill_rows_args(a, n, iblock_begin, iblock_end, adj, members, class_start, n_classes)) return rc;
  if (int rc = check_rows(n, d_valid, pitch)) return rc;
  if (n_classes == 0 || capacity <= 0 || rowcnt_elems(n, iblock_begin, iblock_end) == 0) return GG_OK;
  GG_REQUIRE(adj && members && class_start && offsets && counts && ordered && workspace, GG_ERR_ARG, "null buffer");
  GG_REQUIRE(workspace_bytes >= 256, GG_ERR_WORKSPACE, "workspace too small (256 bytes)");
  unsigned int* cursor = static_cast<unsigned int*>(workspace);
  GG_CUDA_CHECK(cudaMemsetAsync(cursor, 0, 4 * sizeof(unsigned int), gg::as_stream(st)));
  a.offsets = reinterpret_cast<const unsigned long long*>(offsets);
  a.counts = counts;
  a.pitch = pitch;
  a.words = row_words(counts, pitch, d_valid);
  a.ordered = reinterpret_cast<uint2*>(ordered);
  a.capacity = capacity;
  if (int rc = rows_big(a.n_jblocks) ? launch_rows<true, true>(a, cursor, gg::as_stream(st))
                                     : launch_rows<true, false>(a, cursor, gg::as_stream(st)))
    return rc;
  return GG_OK;
}
