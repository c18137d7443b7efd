// f4b: per-row k-nearest-neighbour KF edges (the --representation_ss_faiss_ann mode).
//
// Replaces kf_faiss_to_edge_index_weight (reference gnn_common/gnn_utils.py:122-237): every node keeps its k most
// similar nodes (k = int(edges_keep_top_k * N), gnn_tasks_miniBatch/dataloader.py:113), by inner product of the
// L2-normalised rows ("IP" = cosine) or by squared Euclidean distance of the raw frequency rows ("L2", turned into
// 1 - d / d_max afterwards), self matches and similarities <= 1e-4 removed.  The reference searches an approximate
// FAISS index (IVFFlat by default); this is the exact search that index approximates.
//
// Rows are small non-negative integer counts, so both orders depend on (dot, sqnorm_j) only and take few distinct
// values per row.  The host ranks the (sqnorm class, dot) pairs once (rank 0 = most similar); the device then needs
// two passes over the Gram matrix and never sorts by value:
//   pass 1  hist[i][rank] = number of columns j != i of that rank (per-CTA shared-memory histograms when they fit),
//   plan    per row: the rank where the k-th neighbour falls, how many of that tie class fit, the row's first
//           output slot (exclusive scan) and one cursor per (row, rank),
//   pass 2  the selected (i, j) of every row, in column order: a warp owns its 16 rows for all columns, so the slot of
//           a pair is a running per-row count (kept in registers) plus a ballot prefix -- no atomics,
//   order   a warp per row moves its (at most k) pairs to the (row, rank) cursors: grouped by source row, sorted by
//           similarity, ties by ascending column -- deterministic.
// The Gram tiles come from mma.sync m16n8k32 (u8 x u8 -> s32): a CTA keeps 128 rows in shared memory and streams
// all columns past them in 128-column tiles (register-prefetched).  This is the first correct path for this mode;
// folding the per-row selection into the tcgen05 kernel of gg_kf_mma.cu is the next step.
#include <cub/cub.cuh>

#include "gg_common.cuh"
#include "gg_knn_tc.cuh"

#include <atomic>

namespace {

std::atomic<int> g_knn_impl{0};  // gg_knn_set_impl: 0 auto, 1 mma.sync kernel, 2 tcgen05 kernel (error when it cannot serve)

constexpr int kRows = 128;     // rows per CTA: 8 warps x 16
constexpr int kCols = 128;     // columns per tile
constexpr int kThreads = 256;
constexpr int kChunk = 128;    // bytes of a row staged per column-tile load
constexpr int kMaxPitch = 1024;
constexpr unsigned kNoRank = 0xFFFFu;
constexpr size_t kSmemLimit = 227 * 1024;

enum { kHistGlobal = 0, kHistSmem16 = 1, kHistSmem32 = 2 };

struct KnnArgs {
  const uint8_t* counts;
  int pitch;                 // bytes per row
  int kpad;                  // pitch rounded up to a multiple of 32 (one MMA k-step)
  int chunk;                 // min(kpad, kChunk)
  int seg_shift;             // log2 of the load units per staged row
  int64_t n;
  int64_t row_begin, row_end;  // the shard's rows (row_begin is a multiple of 128); per-row arrays are shard-local
  const int32_t* sqnorm;
  const uint8_t* sqid;       // [sq_max + 1] sqnorm -> class index, 0xFF when no row has it
  int sq_max;
  const uint16_t* rank_tab;  // [n_sq][dot_max + 1] -> rank, 0xFFFF = never a neighbour
  int n_sq, dot_max, n_ranks;
  int hist_mode;             // pass 1: where the histograms are accumulated
  uint32_t* hist;            // [n][n_ranks] pass 1 output
  const int32_t* cut;        // [n] rank of the class the k-th neighbour falls in (n_ranks: fewer than k candidates)
  const int32_t* tie_left;   // [n] members of the cut class to take (the first in column order)
  const int64_t* row_off;    // [n + 1]
  uint64_t* recs;            // pass 2 output: column | dot << 32 | rank << 48, rows at row_off, column order
};

__device__ __forceinline__ void mma_u8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                       uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// 128 rows x `chunk` bytes of the count matrix, as per-thread registers (loaded early, stored to shared memory late).
// Unit = 16 bytes (pitch >= 16) or one word (pitch 4 / 8); a row holds 1 << seg_shift units, a thread at most 4.
struct Prefetch {
  uint4 v[4];
  int32_t sq;
};

__device__ __forceinline__ void prefetch_rows(const KnnArgs& a, Prefetch& p, int64_t row0, int kb0) {
  const bool wide = a.pitch >= 16;
  const int ub = wide ? 4 : 2;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = threadIdx.x + j * kThreads;
    const int r = i >> a.seg_shift, b = kb0 + ((i & ((1 << a.seg_shift) - 1)) << ub);
    const int64_t row = row0 + r;
    p.v[j] = make_uint4(0, 0, 0, 0);
    if (r < kRows && row < a.n && b < a.pitch) {
      const uint8_t* src = a.counts + row * a.pitch + b;
      if (wide) p.v[j] = __ldg(reinterpret_cast<const uint4*>(src));
      else p.v[j].x = __ldg(reinterpret_cast<const uint32_t*>(src));
    }
  }
}

__device__ __forceinline__ void store_rows(const KnnArgs& a, const Prefetch& p, uint8_t* dst, int stride) {
  const bool wide = a.pitch >= 16;
  const int ub = wide ? 4 : 2;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = threadIdx.x + j * kThreads;
    const int r = i >> a.seg_shift, b = (i & ((1 << a.seg_shift) - 1)) << ub;
    if (r < kRows) {
      if (wide) *reinterpret_cast<uint4*>(dst + r * stride + b) = p.v[j];
      else *reinterpret_cast<uint32_t*>(dst + r * stride + b) = p.v[j].x;
    }
  }
}

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const uint8_t* p) {
  const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

// the CTA's 128 rows, all kpad bytes (once per CTA: plain loop)
__device__ __forceinline__ void stage_panel(const KnnArgs& a, uint8_t* dst, int stride, int64_t row0) {
  const int words = a.kpad >> 2;
  for (int i = threadIdx.x; i < kRows * words; i += kThreads) {
    const int r = i / words, wd = i - r * words;
    const int64_t row = row0 + r;
    const int b = wd << 2;
    uint32_t v = 0;
    if (row < a.row_end && b < a.pitch) v = __ldg(reinterpret_cast<const uint32_t*>(a.counts + row * a.pitch + b));
    *reinterpret_cast<uint32_t*>(dst + r * stride + b) = v;
  }
}

template <int PASS, int HIST>
__global__ void __launch_bounds__(kThreads, 1) knn_kernel(KnnArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int pstride = a.kpad + 16, cstride = a.chunk + 16;
  uint8_t* panel = smem;                                        // [128][kpad + 16]
  uint8_t* stage = panel + kRows * pstride;                     // [128][chunk + 16]
  uint16_t* rank_tab = reinterpret_cast<uint16_t*>(stage + kCols * cstride);
  const int tab_elems = (a.n_sq + 1) * (a.dot_max + 1);  // + one all-"never" row for columns past the end
  uint8_t* sqid = reinterpret_cast<uint8_t*>(rank_tab + ((tab_elems + 1) & ~1));     // [256]
  uint8_t* colcls = sqid + 256;                                                       // [128]
  uint32_t* shist = reinterpret_cast<uint32_t*>(colcls + kCols);                      // pass 1, smem modes

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int64_t row0 = a.row_begin + static_cast<int64_t>(blockIdx.x) * kRows;
  const int64_t loc0 = static_cast<int64_t>(blockIdx.x) * kRows;  // shard-local index of the CTA's first row

  for (int i = threadIdx.x; i < tab_elems; i += kThreads)
    rank_tab[i] = i < a.n_sq * (a.dot_max + 1) ? a.rank_tab[i] : static_cast<uint16_t>(kNoRank);
  for (int i = threadIdx.x; i < 256; i += kThreads) sqid[i] = i <= a.sq_max ? a.sqid[i] : 0xFF;
  int hist_words = 0;
  if (PASS == 1 && HIST != kHistGlobal) {
    hist_words = HIST == kHistSmem16 ? kRows * ((a.n_ranks + 1) >> 1) : kRows * a.n_ranks;
    for (int i = threadIdx.x; i < hist_words; i += kThreads) shist[i] = 0;
  }
  stage_panel(a, panel, pstride, row0);

  // the two rows this lane sees in every accumulator fragment
  const int lrow[2] = {warp * 16 + g, warp * 16 + g + 8};
  const int64_t grow[2] = {row0 + lrow[0], row0 + lrow[1]};
  const int64_t loc[2] = {loc0 + lrow[0], loc0 + lrow[1]};
  int32_t cut[2] = {-1, -1}, budget[2] = {0, 0};
  int64_t rbase[2] = {0, 0};
  int pos[2] = {0, 0}, seen[2] = {0, 0};
  if (PASS == 2) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (grow[h] < a.row_end) {
        cut[h] = a.cut[loc[h]];
        budget[h] = a.tie_left[loc[h]];
        rbase[h] = a.row_off[loc[h]];
      }
  }
  const bool rowok[2] = {grow[0] < a.row_end, grow[1] < a.row_end};
  const int hwords = HIST == kHistSmem16 ? (a.n_ranks + 1) >> 1 : a.n_ranks;
  const int hbase[2] = {lrow[0] * hwords, lrow[1] * hwords};
  const unsigned lower = (0xFu << (4 * g)) & ((1u << lane) - 1u);  // lanes of my row group with a smaller q
  const unsigned group = 0xFu << (4 * g);

  // ldmatrix addresses: lane l supplies row (l & 7) of 8x16-byte matrix (l >> 3).
  // A: matrices (rows 0-7, k 0-15), (rows 8-15, k 0-15), (rows 0-7, k 16-31), (rows 8-15, k 16-31) = a0..a3.
  // B: (n-tile nt, k 0-15), (nt, k 16-31), (nt + 1, k 0-15), (nt + 1, k 16-31) = b0, b1 of two n-tiles.
  const int mi = lane >> 3, mr = lane & 7;
  const uint8_t* a_ld = panel + (warp * 16 + mr + (mi & 1) * 8) * pstride + (mi >> 1) * 16;
  const uint8_t* b_ld = stage + ((mi >> 1) * 8 + mr) * cstride + (mi & 1) * 16;
  const int n_chunks = a.kpad / a.chunk;
  const int64_t n_tiles = (a.n + kCols - 1) / kCols;

  Prefetch pre;
  prefetch_rows(a, pre, 0, 0);
  pre.sq = (threadIdx.x < kCols && threadIdx.x < a.n) ? a.sqnorm[threadIdx.x] : -1;

  for (int64_t tile = 0; tile < n_tiles; ++tile) {
    const int64_t col0 = tile * kCols;
    int acc[16][4];
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0;

    for (int c = 0; c < n_chunks; ++c) {
      __syncthreads();  // the previous stage (and colcls) has been consumed
      store_rows(a, pre, stage, cstride);
      if (c == 0 && threadIdx.x < kCols) {
        const unsigned cls = pre.sq >= 0 ? sqid[pre.sq] : 0xFFu;
        colcls[threadIdx.x] = cls == 0xFFu ? a.n_sq : cls;  // the all-"never" row of rank_tab
      }
      __syncthreads();
      {  // next stage's global loads fly during this stage's MMAs
        const bool last_chunk = c + 1 == n_chunks;
        const int64_t ncol0 = last_chunk ? col0 + kCols : col0;
        if (ncol0 < a.n) {
          prefetch_rows(a, pre, ncol0, last_chunk ? 0 : (c + 1) * a.chunk);
          if (last_chunk && threadIdx.x < kCols)
            pre.sq = ncol0 + threadIdx.x < a.n ? a.sqnorm[ncol0 + threadIdx.x] : -1;
        }
      }
      for (int ks = 0; ks < a.chunk; ks += 32) {
        uint32_t af[4];
        ldsm_x4(af, a_ld + c * a.chunk + ks);
#pragma unroll
        for (int nt = 0; nt < 16; nt += 2) {
          uint32_t bf[4];
          ldsm_x4(bf, b_ld + nt * 8 * cstride + ks);
          mma_u8(acc[nt], af[0], af[1], af[2], af[3], bf[0], bf[1]);
          mma_u8(acc[nt + 1], af[0], af[1], af[2], af[3], bf[2], bf[3]);
        }
      }
    }

    // epilogue: fragment element e of n-tile nt is (row grow[e >> 1], column col0 + nt * 8 + 2 q + (e & 1)).
    // Branch-free rank lookup: columns past the end sit in an all-"never" class; the self pair only exists in the
    // diagonal tile (row block = column tile).
    const bool diag = col0 == row0;
    const int tstride = a.dot_max + 1;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      const int lc0 = nt * 8 + 2 * q;
      const int cb[2] = {colcls[lc0] * tstride, colcls[lc0 + 1] * tstride};
      unsigned rk[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const unsigned r = rank_tab[cb[e & 1] + min(acc[nt][e], a.dot_max)];
        const bool dead = !rowok[e >> 1] || (diag && lc0 + (e & 1) == lrow[e >> 1]);
        rk[e] = dead ? kNoRank : r;
      }
      if (PASS == 1) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (rk[e] == kNoRank) continue;
          if (HIST == kHistSmem32)
            atomicAdd(shist + hbase[e >> 1] + rk[e], 1u);
          else if (HIST == kHistSmem16)
            atomicAdd(shist + hbase[e >> 1] + (rk[e] >> 1), 1u << ((rk[e] & 1) * 16));
          else
            atomicAdd(a.hist + loc[e >> 1] * a.n_ranks + rk[e], 1u);
        }
      } else {
        // every lane of a row group keeps the same running state of its two rows (pos: pairs written, seen: members
        // of the cut class met), advanced by ballot totals -- no shared state, no atomics
        bool sel[4], tie[4];
        bool any = false, anytie = false;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          sel[e] = static_cast<int>(rk[e]) <= cut[e >> 1];  // 0xFFFF is above every cut
          tie[e] = static_cast<int>(rk[e]) == cut[e >> 1];
          any |= sel[e];
          anytie |= tie[e];
        }
        if (!__any_sync(0xFFFFFFFFu, any)) continue;
        if (__any_sync(0xFFFFFFFFu, anytie)) {
          // members of the cut class are taken in column order while the row's tie budget lasts
          unsigned bt[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) bt[e] = __ballot_sync(0xFFFFFFFFu, tie[e]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int h = e >> 1;
            const int before = __popc(bt[2 * h] & lower) + __popc(bt[2 * h + 1] & lower) + ((e & 1) && tie[e - (e & 1)]);
            if (tie[e] && seen[h] + before >= budget[h]) sel[e] = false;
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) seen[h] += __popc(bt[2 * h] & group) + __popc(bt[2 * h + 1] & group);
        }
        unsigned ba[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) ba[e] = __ballot_sync(0xFFFFFFFFu, sel[e]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (!sel[e]) continue;
          const int h = e >> 1;
          const int before = __popc(ba[2 * h] & lower) + __popc(ba[2 * h + 1] & lower) + ((e & 1) && sel[e - (e & 1)]);
          a.recs[rbase[h] + pos[h] + before] = static_cast<uint64_t>(col0 + lc0 + (e & 1)) |
                                               static_cast<uint64_t>(acc[nt][e]) << 32 |
                                               static_cast<uint64_t>(rk[e]) << 48;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) pos[h] += __popc(ba[2 * h] & group) + __popc(ba[2 * h + 1] & group);
      }
    }
  }

  if (PASS == 1 && HIST != kHistGlobal) {
    __syncthreads();
    const int rows_here = a.row_end - row0 < kRows ? static_cast<int>(a.row_end - row0) : kRows;
    const int total = rows_here * a.n_ranks;
    uint32_t* out = a.hist + loc0 * a.n_ranks;
    if (HIST == kHistSmem32) {
      for (int i = threadIdx.x; i < total; i += kThreads) out[i] = shist[i];
    } else {
      const int wpr = (a.n_ranks + 1) >> 1;
      for (int i = threadIdx.x; i < total; i += kThreads) {
        const int r = i / a.n_ranks, j = i - r * a.n_ranks;
        out[i] = (shist[r * wpr + (j >> 1)] >> ((j & 1) * 16)) & 0xFFFFu;
      }
    }
  }
}

// ---- order: column-ordered pairs of a row -> (row, rank) cursors ------------------------------------------------------
struct OrderArgs {
  const uint64_t* recs;
  const int64_t* row_off;
  int64_t n;          // rows of the shard
  int64_t row_begin;
  int n_ranks;
  uint32_t* cursors;  // [n][n_ranks] first output slot of every (row, rank); advanced here
  const int32_t* sqnorm;
  int metric;
  const int32_t* qmax;
  int64_t* src;
  int64_t* dst;
  float* w;
};

__global__ void __launch_bounds__(256) order_kernel(OrderArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  if (row >= a.n) return;
  const int64_t base = a.row_off[row];
  const int cnt = static_cast<int>(a.row_off[row + 1] - base);
  if (cnt == 0) return;
  const int sqi = a.sqnorm[a.row_begin + row];
  const float qmax_f = a.metric == 1 ? static_cast<float>(*a.qmax) : 1.0f;
  uint32_t* cur = a.cursors + row * a.n_ranks;
  for (int i0 = 0; i0 < cnt; i0 += 32) {
    const int i = i0 + lane;
    const bool valid = i < cnt;
    const uint64_t rec = valid ? a.recs[base + i] : ~0ull;
    const unsigned rank = static_cast<unsigned>(rec >> 48);  // 0xFFFF on idle lanes: a group of its own
    const unsigned same = __match_any_sync(0xFFFFFFFFu, rank);
    const int leader = __ffs(same) - 1;
    unsigned first = 0;
    if (valid && lane == leader) first = atomicAdd(cur + rank, static_cast<unsigned>(__popc(same)));
    first = __shfl_sync(0xFFFFFFFFu, first, leader);
    if (valid) {
      const int64_t slot = static_cast<int64_t>(first) + __popc(same & ((1u << lane) - 1u));
      const int64_t col = static_cast<int64_t>(rec & 0xFFFFFFFFull);
      const int dot = static_cast<int>((rec >> 32) & 0xFFFFu);
      const int sqj = a.sqnorm[col];
      float wv;
      if (a.metric == 0) {
        wv = fminf(static_cast<float>(static_cast<double>(dot) / sqrt(static_cast<double>(sqi) * sqj)), 1.0f);
      } else {
        wv = 1.0f - static_cast<float>(sqi + sqj - 2 * dot) / qmax_f;
      }
      a.src[slot] = a.row_begin + row;
      a.dst[slot] = col;
      a.w[slot] = wv;
    }
    __syncwarp();
  }
}

// The same for the staging layout of the tcgen05 kernel (gg_knn_tc.cuh): a row's pairs come as three column thirds, each
// in column order, each holding at most the row's tie budget of cut-rank members; walking the thirds in order and
// keeping cut-rank members while the budget lasts reproduces the column-order rule over the whole row.
struct Order3Args {
  OrderArgs o;
  const uint32_t* third_cnt;
  const int32_t* cut;
  const int32_t* tie_left;
};

__global__ void __launch_bounds__(256) order3_kernel(Order3Args a3) {
  const OrderArgs& a = a3.o;
  const int lane = threadIdx.x & 31;
  const int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  if (row >= a.n) return;
  const int64_t base = a.row_off[row];
  const int64_t cnt_row = a.row_off[row + 1] - base;
  if (cnt_row == 0) return;
  const int sqi = a.sqnorm[a.row_begin + row];
  const float qmax_f = a.metric == 1 ? static_cast<float>(*a.qmax) : 1.0f;
  uint32_t* cur = a.cursors + row * a.n_ranks;
  const unsigned cut = static_cast<unsigned>(a3.cut[row]);
  const int budget = a3.tie_left[row];
  int seen = 0;
  for (int g = 0; g < 3; ++g) {
    const int cnt = static_cast<int>(a3.third_cnt[3 * row + g]);
    const uint64_t* src = a.recs + 3 * base + g * cnt_row;
    for (int i0 = 0; i0 < cnt; i0 += 32) {
      const int i = i0 + lane;
      const bool valid = i < cnt;
      const uint64_t rec = valid ? src[i] : ~0ull;
      const unsigned rank = static_cast<unsigned>(rec >> 48);
      const bool is_tie = valid && rank == cut;
      const unsigned tb = __ballot_sync(0xFFFFFFFFu, is_tie);
      const bool keep = valid && (!is_tie || seen + __popc(tb & ((1u << lane) - 1u)) < budget);
      seen += __popc(tb);
      const unsigned key = keep ? rank : 0x10000u + lane;  // dropped and idle lanes: groups of their own
      const unsigned same = __match_any_sync(0xFFFFFFFFu, key);
      const int leader = __ffs(same) - 1;
      unsigned first = 0;
      if (keep && lane == leader) first = atomicAdd(cur + rank, static_cast<unsigned>(__popc(same)));
      first = __shfl_sync(0xFFFFFFFFu, first, leader);
      if (keep) {
        const int64_t slot = static_cast<int64_t>(first) + __popc(same & ((1u << lane) - 1u));
        const int64_t col = static_cast<int64_t>(rec & 0xFFFFFFFFull);
        const int dot = static_cast<int>((rec >> 32) & 0xFFFFu);
        const int sqj = a.sqnorm[col];
        float wv;
        if (a.metric == 0) {
          wv = fminf(static_cast<float>(static_cast<double>(dot) / sqrt(static_cast<double>(sqi) * sqj)), 1.0f);
        } else {
          wv = 1.0f - static_cast<float>(sqi + sqj - 2 * dot) / qmax_f;
        }
        a.src[slot] = a.row_begin + row;
        a.dst[slot] = col;
        a.w[slot] = wv;
      }
      __syncwarp();
    }
  }
}

// ---- plan: cut rank, tie budget, rows' output counts -------------------------------------------------------------
__global__ void cut_kernel(const uint32_t* hist, int64_t n, int n_ranks, int k_nn, int metric, const int32_t* sqnorm,
                           const int32_t* rank_q, int32_t* cut, int32_t* tie_left, uint32_t* before, int32_t* qmax) {
  const int64_t row = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (row >= n) return;
  const uint32_t* h = hist + row * n_ranks;
  uint32_t cum = 0;
  int c = n_ranks, left = 0, last = -1;
  for (int r = 0; r < n_ranks; ++r) {
    const uint32_t x = h[r];
    if (x == 0) continue;
    last = r;
    if (cum + x >= static_cast<uint32_t>(k_nn)) {
      c = r;
      left = k_nn - static_cast<int>(cum);
      break;
    }
    cum += x;
  }
  cut[row] = c;
  tie_left[row] = left;
  before[row] = cum;
  if (metric == 1 && last >= 0) atomicMax(qmax, sqnorm[row] + rank_q[last]);  // `last` is the worst class kept
}

// L2: neighbours at the largest kept distance get similarity 0 and are dropped (gnn_utils.py:214-224); in integer
// units 1 - d / d_max > 1e-4 is d < d_max.  They can only sit in the row's worst kept class.
__global__ void rows_kernel(int64_t n, int n_ranks, int metric, const int32_t* sqnorm, const int32_t* rank_q,
                            const int32_t* qmax, int32_t* cut, int32_t* tie_left, const uint32_t* before,
                            const uint32_t* hist, uint32_t* rowcnt) {
  const int64_t row = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (row >= n) return;
  int c = cut[row];
  uint32_t cnt = before[row] + static_cast<uint32_t>(tie_left[row]);
  if (metric == 1) {
    if (c < n_ranks) {
      if (sqnorm[row] + rank_q[c] >= *qmax) {  // the whole cut class is at the global maximum
        cnt = before[row];
        tie_left[row] = 0;
      }
    } else {
      // fewer than k candidates: every class was kept; drop a last class sitting at the maximum
      int last = -1;
      for (int r = n_ranks - 1; r >= 0; --r)
        if (hist[row * n_ranks + r]) {
          last = r;
          break;
        }
      if (last >= 0 && sqnorm[row] + rank_q[last] >= *qmax) {
        cnt = before[row] - hist[row * n_ranks + last];
        cut[row] = last;
        tie_left[row] = 0;
      }
    }
  }
  rowcnt[row] = cnt;
}

struct ToI64 {
  __host__ __device__ int64_t operator()(uint32_t x) const { return static_cast<int64_t>(x); }
};

// hist[row][r] := first output slot of rank r in row `row`
__global__ void cursor_kernel(int64_t n, int n_ranks, const int64_t* row_off, uint32_t* hist) {
  const int64_t row = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (row >= n) return;
  uint32_t* h = hist + row * n_ranks;
  uint32_t at = static_cast<uint32_t>(row_off[row]);
  for (int r = 0; r < n_ranks; ++r) {
    const uint32_t x = h[r];
    h[r] = at;
    at += x;
  }
}

size_t knn_base_smem(int kpad, int chunk, int n_sq, int dot_max) {
  const int tab_elems = (n_sq + 1) * (dot_max + 1);
  return static_cast<size_t>(kRows) * (kpad + 16) + static_cast<size_t>(kCols) * (chunk + 16) +
         2 * static_cast<size_t>((tab_elems + 1) & ~1) + 256 + kCols;
}

size_t hist_smem(int mode, int n_ranks) {
  if (mode == kHistSmem32) return static_cast<size_t>(kRows) * n_ranks * 4;
  if (mode == kHistSmem16) return static_cast<size_t>(kRows) * ((n_ranks + 1) / 2) * 4;
  return 0;
}

int fill_args(KnnArgs& a, const uint8_t* counts, int pitch, int64_t n, int64_t row_begin, int64_t row_end,
              const int32_t* sqnorm, const uint8_t* sqid, int sq_max, const uint16_t* rank_tab, int n_sq, int dot_max,
              int n_ranks) {
  GG_REQUIRE(n >= 1 && n < (int64_t(1) << 31), GG_ERR_ARG, "n outside [1, 2^31)");
  GG_REQUIRE(row_begin >= 0 && row_begin < row_end && row_end <= n && row_begin % kRows == 0, GG_ERR_ARG,
             "row shard must be a non-empty [multiple of 128, <= n) range");
  GG_REQUIRE(pitch == 4 || pitch == 8 || (pitch >= 16 && pitch % 16 == 0 && pitch <= kMaxPitch), GG_ERR_ARG,
             "pitch must be 4, 8 or a multiple of 16 up to 1024");
  GG_REQUIRE(counts && sqnorm && sqid && rank_tab, GG_ERR_ARG, "null argument");
  GG_REQUIRE(n_sq >= 1 && n_sq <= 254 && sq_max >= 0 && sq_max <= 254 && dot_max >= 0, GG_ERR_ARG, "bad class table");
  GG_REQUIRE(n_ranks >= 1 && n_ranks < 0xFFFF, GG_ERR_ARG, "n_ranks outside [1, 65535)");
  GG_REQUIRE(static_cast<uint64_t>(row_end - row_begin) * static_cast<uint64_t>(n_ranks) < (uint64_t(1) << 40),
             GG_ERR_ARG, "histogram too large");
  a = KnnArgs{};
  a.counts = counts;
  a.pitch = pitch;
  a.kpad = (pitch + 31) / 32 * 32;
  a.chunk = a.kpad < kChunk ? a.kpad : kChunk;
  GG_REQUIRE(a.kpad % a.chunk == 0, GG_ERR_ARG, "row pitch above 128 must be a multiple of 128");
  GG_REQUIRE(a.chunk == 32 || a.chunk == 64 || a.chunk == 128, GG_ERR_ARG, "row pitch of 65..127 bytes");
  const int units = pitch >= 16 ? a.chunk / 16 : a.chunk / 4;  // 2, 4 or 8 per staged row
  a.seg_shift = units == 2 ? 1 : units == 4 ? 2 : 3;
  a.n = n;
  a.row_begin = row_begin;
  a.row_end = row_end;
  a.sqnorm = sqnorm;
  a.sqid = sqid;
  a.sq_max = sq_max;
  a.rank_tab = rank_tab;
  a.n_sq = n_sq;
  a.dot_max = dot_max;
  a.n_ranks = n_ranks;
  GG_REQUIRE(knn_base_smem(a.kpad, a.chunk, n_sq, dot_max) <= kSmemLimit, GG_ERR_ARG,
             "tables do not fit shared memory");
  return GG_OK;
}

template <int PASS, int HIST>
int launch_knn(const KnnArgs& a, size_t smem, cudaStream_t s) {
  GG_CUDA_CHECK(cudaFuncSetAttribute(knn_kernel<PASS, HIST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
  const unsigned grid = static_cast<unsigned>((a.row_end - a.row_begin + kRows - 1) / kRows);
  knn_kernel<PASS, HIST><<<grid, kThreads, smem, s>>>(a);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}

// =====================================================================================================================
// Sparse path: an inverted-index join instead of a Gram matrix.
//
// A row of 4^s sub-k-mer counts has at most k - s + 1 non-zero entries (4 of 1024 at k = 8, s = 5), so two rows have a
// non-zero dot only when they share a column: 1.5 % of the pairs at s = 5.  Both passes therefore never form the zero
// dots: a CTA takes one row i, walks the posting lists of the row's non-zero columns (rows that hold the column,
// ascending, with their counts) and accumulates dot(i, j) into a byte per column j in shared memory (N bytes: up to
// ~200,000 nodes).  Pass 1 then visits the touched j once more, takes each accumulated dot out with an atomic AND (the
// first visitor of a j gets it, later ones get zero: a j reached through two columns is counted once, and the
// accumulator is clean again for the next row), and counts ranks; zero-dot members of a column class are its size minus
// the row's non-zero count for that class, as on the tensor path.  Pass 2 accumulates the same way and then scans the
// accumulator in COLUMN order (every thread a contiguous column range, two sweeps around a block scan: count, then
// write), which is the order the tie rule needs; zero-dot columns are selected through the per-column class array.
// The plan kernels and the reorder kernel are shared with the dense paths.
// =====================================================================================================================
constexpr int kSpMaxNnz = 16;        // non-zero entries per row the index holds (rows of k-mer counts: <= 15)
constexpr int kSpThreads = 256;
constexpr int kSpMaxNodes = 200000;  // accumulator bytes + tables must fit one CTA's shared memory

constexpr int kKnnMaxSq = 64;

struct SparseIndex {
  // row side: entries of row i at rnz[i * kSpMaxNnz ...], column | count << 16
  uint32_t* rnz;
  uint8_t* rnz_n;
  // column side (posting lists): rows of column c, ascending, at post[col_start[c] .. col_start[c + 1]), row | count << 24
  uint32_t* col_start;   // [d + 1]
  uint32_t* post;        // [n * kSpMaxNnz]
  uint8_t* colcls;       // [n + 4] norm class of every node (n_sq: none)
  uint32_t* colcount;    // [kKnnMaxSq] nodes per class
  unsigned int* cursor;  // [4] row cursors of the two passes
  // build-time scratch: (column, row | count << 24) pairs in row order -> stable radix sort by column
  uint32_t* row_start;   // [n + 1]
  uint16_t *key_in, *key_out;
  uint32_t* val_in;
  void* cub_temp;
  size_t cub_bytes;
  size_t total;
};

struct U8ToU32 {
  __host__ __device__ uint32_t operator()(uint8_t v) const { return v; }
};

size_t sparse_cub_bytes(int64_t n) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, static_cast<const uint16_t*>(nullptr), static_cast<uint16_t*>(nullptr),
                                  static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr),
                                  static_cast<int>(n * kSpMaxNnz));
  cub::TransformInputIterator<uint32_t, U8ToU32, const uint8_t*> it(nullptr, U8ToU32());
  cub::DeviceScan::ExclusiveSum(nullptr, b, it, static_cast<uint32_t*>(nullptr), static_cast<int>(n + 1));
  return a > b ? a : b;
}

SparseIndex sparse_carve(void* ws, int64_t n, int d) {
  gg::Carver c(ws);
  SparseIndex x;
  const size_t cap = static_cast<size_t>(n) * kSpMaxNnz;
  x.rnz = c.take<uint32_t>(cap);
  x.rnz_n = c.take<uint8_t>(n + 1);
  x.col_start = c.take<uint32_t>(d + 1);
  x.post = c.take<uint32_t>(cap);
  x.colcls = c.take<uint8_t>(n + 4);
  x.colcount = c.take<uint32_t>(kKnnMaxSq);
  x.cursor = c.take<unsigned int>(4);
  x.row_start = c.take<uint32_t>(n + 1);
  x.key_in = c.take<uint16_t>(cap);
  x.key_out = c.take<uint16_t>(cap);
  x.val_in = c.take<uint32_t>(cap);
  x.cub_bytes = sparse_cub_bytes(n);
  x.cub_temp = c.take<char>(x.cub_bytes);
  x.total = c.used();
  return x;
}

// one warp per row: non-zero entries of the row (ballot-compacted, ascending column), column histogram, norm class
__global__ void __launch_bounds__(256) sparse_rows_kernel(const uint8_t* counts, int pitch, int64_t n, const int32_t* sqnorm,
                                                          const uint8_t* sqid, int sq_max, int n_sq, SparseIndex x,
                                                          unsigned long long* status) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  if (row >= n) return;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(counts + row * pitch);
  int found = 0;
  for (int w0 = 0; w0 < pitch / 4; w0 += 32) {
    const int w = w0 + lane;
    const uint32_t v = w < pitch / 4 ? __ldg(src + w) : 0u;
    for (uint32_t any = __ballot_sync(0xFFFFFFFFu, v != 0u); any; any &= any - 1) {
      const int l = __ffs(any) - 1;
      const uint32_t word = __shfl_sync(0xFFFFFFFFu, v, l);
      if (lane == 0) {
        for (int b = 0; b < 4; ++b) {
          const uint32_t cnt = (word >> (8 * b)) & 0xFFu;
          if (!cnt) continue;
          const uint32_t col = static_cast<uint32_t>(4 * (w0 + l) + b);
          if (found < kSpMaxNnz) {
            x.rnz[row * kSpMaxNnz + found] = col | (cnt << 16);
            atomicAdd(x.col_start + col + 1, 1u);   // histogram, turned into offsets by the scan
          }
          ++found;
        }
      }
    }
  }
  if (lane == 0) {
    x.rnz_n[row] = static_cast<uint8_t>(found < kSpMaxNnz ? found : kSpMaxNnz);
    if (found > kSpMaxNnz) atomicMax(status, static_cast<unsigned long long>(found));  // the caller must not use this path
    const int sq = sqnorm[row];
    uint32_t cls = 0xFFu;
    if (sq >= 0 && sq <= sq_max) cls = sqid[sq];
    x.colcls[row] = static_cast<uint8_t>(cls == 0xFFu ? n_sq : cls);
    if (cls != 0xFFu) atomicAdd(x.colcount + cls, 1u);
  }
}

// col_start[c + 1] holds the size of column c: inclusive scan in place (one CTA; d <= 1024)
__global__ void __launch_bounds__(1024) sparse_scan_kernel(uint32_t* col_start, int d) {
  __shared__ uint32_t buf[1024];
  const int t = threadIdx.x;
  uint32_t v = t < d ? col_start[t + 1] : 0u;
  buf[t] = v;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const uint32_t u = t >= o ? buf[t - o] : 0u;
    __syncthreads();
    buf[t] += u;
    __syncthreads();
  }
  if (t < d) col_start[t + 1] = buf[t];
  if (t == 0) col_start[0] = 0u;
}

// (column, row | count << 24) pairs in row order; unused slots keep the all-ones key and sort to the end
__global__ void __launch_bounds__(256) sparse_pairs_kernel(int64_t n, SparseIndex x) {
  const int64_t row = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (row >= n) return;
  const int cnt = x.rnz_n[row];
  const uint32_t at = x.row_start[row];
  for (int e = 0; e < cnt; ++e) {
    const uint32_t ent = x.rnz[row * kSpMaxNnz + e];
    x.key_in[at + e] = static_cast<uint16_t>(ent & 0xFFFFu);
    x.val_in[at + e] = static_cast<uint32_t>(row) | ((ent >> 16) << 24);
  }
}

struct SparseArgs {
  SparseIndex x;
  int64_t n, row_begin, row_end;
  const int32_t* sqnorm;
  const uint16_t* rank_tab;   // [n_sq][dot_max + 1]
  int n_sq, dot_max, n_ranks;
  uint32_t* hist;             // pass 1 out
  const int32_t* cut;
  const int32_t* tie_left;
  const int64_t* row_off;
  uint64_t* recs;
  int per_log2;               // log2 of the accumulator words a thread of pass 2 scans (sp_phys)
};

// The accumulator holds one byte per node, four to a word.  Pass 2 lets thread t scan the words [t << per_log2,
// (t + 1) << per_log2) in ascending order, which would put all 32 lanes of a warp on one bank; skewing the layout by one
// word per thread range (word w lives at w + (w >> per_log2)) makes lane t hit bank (t + step) mod 32 instead.
__device__ __forceinline__ int sp_phys(int w, int per_log2) { return w + (w >> per_log2); }
// bit 7 of every non-zero byte
__device__ __forceinline__ uint32_t nonzero_bytes_sp(uint32_t w) { return (((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w) & 0x80808080u; }

// accumulate dot(i, j) for every j that shares a column with row i: acc is one byte per node (dots are <= 255)
__device__ __forceinline__ void sparse_accumulate(const SparseArgs& a, int64_t row, uint32_t* acc, uint32_t* lists,
                                                  int& n_lists, int& total) {
  // lists[3 e], [3 e + 1], [3 e + 2] = first posting entry, exclusive prefix of the lengths, this row's count
  const int cnt = a.x.rnz_n[row];
  if (threadIdx.x < 32) {
    uint32_t len = 0;
    if (static_cast<int>(threadIdx.x) < cnt) {
      const uint32_t ent = a.x.rnz[row * kSpMaxNnz + threadIdx.x];
      const uint32_t col = ent & 0xFFFFu;
      const uint32_t first = a.x.col_start[col];
      len = a.x.col_start[col + 1] - first;
      lists[3 * threadIdx.x] = first;
      lists[3 * threadIdx.x + 2] = ent >> 16;
    }
    uint32_t incl = len;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (static_cast<int>(threadIdx.x) >= o) incl += u;
    }
    if (static_cast<int>(threadIdx.x) < cnt) lists[3 * threadIdx.x + 1] = incl - len;
    if (threadIdx.x == 31) {
      lists[3 * kSpMaxNnz] = static_cast<uint32_t>(cnt);
      lists[3 * kSpMaxNnz + 1] = incl;
    }
  }
  __syncthreads();
  n_lists = static_cast<int>(lists[3 * kSpMaxNnz]);
  total = static_cast<int>(lists[3 * kSpMaxNnz + 1]);
  for (int t = threadIdx.x; t < total; t += kSpThreads) {
    int e = 0;
    while (e + 1 < n_lists && lists[3 * (e + 1) + 1] <= static_cast<uint32_t>(t)) ++e;
    const uint32_t p = __ldg(a.x.post + lists[3 * e] + (t - lists[3 * e + 1]));
    const uint32_t j = p & 0xFFFFFFu;
    atomicAdd(acc + sp_phys(static_cast<int>(j >> 2), a.per_log2), ((p >> 24) * lists[3 * e + 2]) << (8 * (j & 3u)));
  }
  __syncthreads();
}

template <int PASS>
__global__ void __launch_bounds__(kSpThreads) knn_sparse_kernel(SparseArgs a, unsigned int* row_cursor) {
  extern __shared__ __align__(16) uint32_t sp_smem[];
  const int acc_words = static_cast<int>((a.n + 3) / 4);
  const int acc_alloc = (sp_phys(acc_words, a.per_log2) + 4) & ~3;
  uint32_t* acc = sp_smem;                                   // one byte per node, skewed (sp_phys)
  uint32_t* lists = acc + acc_alloc;                         // [3 * kSpMaxNnz + 2]
  uint32_t* shist = lists + 3 * kSpMaxNnz + 4;               // pass 1: [n_ranks] + [n_sq] non-zero counts per class
  uint32_t* scan = lists + 3 * kSpMaxNnz + 4;                // pass 2: [2 * kSpThreads] block scan
  __shared__ int s_row;
  const int tstride = a.dot_max + 1;
  for (int i = threadIdx.x; i < acc_alloc; i += kSpThreads) acc[i] = 0u;
  if (PASS == 1)
    for (int i = threadIdx.x; i < a.n_ranks + a.n_sq; i += kSpThreads) shist[i] = 0u;
  __syncthreads();
  while (true) {
    if (threadIdx.x == 0) s_row = static_cast<int>(atomicAdd(row_cursor, 1u));
    __syncthreads();
    const int64_t row = a.row_begin + s_row;
    __syncthreads();
    if (row >= a.row_end) break;
    int n_lists, total;
    sparse_accumulate(a, row, acc, lists, n_lists, total);
    const int64_t loc = row - a.row_begin;
    if (PASS == 1) {
      uint32_t* nzc = shist + a.n_ranks;
      for (int t = threadIdx.x; t < total; t += kSpThreads) {
        int e = 0;
        while (e + 1 < n_lists && lists[3 * (e + 1) + 1] <= static_cast<uint32_t>(t)) ++e;
        const uint32_t j = __ldg(a.x.post + lists[3 * e] + (t - lists[3 * e + 1])) & 0xFFFFFFu;
        const int sh = 8 * (j & 3u);
        const uint32_t old = atomicAnd(acc + sp_phys(static_cast<int>(j >> 2), a.per_log2), ~(0xFFu << sh));   // the first visitor of j takes the dot out
        const uint32_t dot = (old >> sh) & 0xFFu;
        if (!dot) continue;
        const uint32_t cls = a.x.colcls[j];
        if (cls >= static_cast<uint32_t>(a.n_sq)) continue;
        atomicAdd(nzc + cls, 1u);            // the self pair too: it is not a zero-dot column
        if (static_cast<int64_t>(j) == row) continue;
        const uint32_t r = a.rank_tab[cls * tstride + min(dot, static_cast<uint32_t>(a.dot_max))];
        if (r != kNoRank) atomicAdd(shist + r, 1u);
      }
      __syncthreads();
      if (threadIdx.x < a.n_sq) {   // zero-dot members of every class
        const int c = threadIdx.x;
        const uint32_t r0 = a.rank_tab[c * tstride];
        if (r0 != kNoRank) {
          const int sq_row = a.sqnorm[row];
          const bool own = sq_row == 0 && a.x.colcls[row] == c;   // an all-zero row meets itself with dot 0
          atomicAdd(shist + r0, a.x.colcount[c] - nzc[c] - (own ? 1u : 0u));
        }
      }
      __syncthreads();
      for (int i = threadIdx.x; i < a.n_ranks; i += kSpThreads) {
        a.hist[loc * a.n_ranks + i] = shist[i];
        shist[i] = 0u;
      }
      if (threadIdx.x < a.n_sq) nzc[threadIdx.x] = 0u;
      __syncthreads();
    } else {
      // column-ordered selection: thread t owns the words [t * per, (t + 1) * per) of the accumulator
      const int cut = a.cut[loc], budget = a.tie_left[loc];
      uint32_t zsel = 0, ztie = 0;   // classes whose zero-dot columns rank above / at the cut (n_sq <= 32 here)
      for (int c = 0; c < a.n_sq; ++c) {
        const int r0 = a.rank_tab[c * tstride];
        zsel |= static_cast<uint32_t>(r0 < cut) << c;
        ztie |= static_cast<uint32_t>(r0 == cut) << c;
      }
      const bool zeros = (zsel | ztie) != 0u;
      const int per = 1 << a.per_log2;
      const int w_lo = min(acc_words, static_cast<int>(threadIdx.x) * per), w_hi = min(acc_words, w_lo + per);
      uint32_t* my_acc = acc + threadIdx.x;   // sp_phys(w) = w + threadIdx.x inside this thread's range
      // A lane's range holds a handful of non-zero words, but some lane of the warp has one at nearly every step: walking
      // the range word by word would run the per-column code 64 times per sweep with one or two lanes busy.  So the
      // range is looked at once, 64 words at a time, for a bit mask of the words that matter (all of them when zero-dot
      // columns can be selected), and both sweeps walk the set bits.
      int n_sel = 0, n_tie = 0;
      unsigned long long masks[4] = {0ull, 0ull, 0ull, 0ull};
      for (int pass2 = 0; pass2 < 2; ++pass2) {
        int sel_before = 0, tie_before = 0, tie_room = 0;
        long long out = 0;
        if (pass2) {
          sel_before = static_cast<int>(scan[threadIdx.x]);
          tie_before = static_cast<int>(scan[kSpThreads + threadIdx.x]);
          tie_room = budget - tie_before;   // members of the cut rank this thread may still take, in column order
          out = a.row_off[loc] + sel_before + (tie_before < budget ? tie_before : budget);
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {   // a thread's range is at most 256 words: four masks, kept for the second sweep
          const int w0 = w_lo + 64 * ch;
          if (w0 >= w_hi) continue;
          const int span = min(64, w_hi - w0);
          if (!pass2) {
            unsigned long long m0 = 0;
            if (zeros) {
              m0 = span == 64 ? ~0ull : ((1ull << span) - 1ull);
            } else {
              for (int i = 0; i < span; ++i) m0 |= static_cast<unsigned long long>(my_acc[w0 + i] != 0u) << i;
            }
            masks[ch] = m0;
          }
          for (unsigned long long mask = masks[ch]; mask; mask &= mask - 1) {
            const int w = w0 + __ffsll(static_cast<long long>(mask)) - 1;
            const uint32_t word = my_acc[w];
            if (pass2 && word) my_acc[w] = 0u;   // clean for the next row
            const uint32_t cls4 = *reinterpret_cast<const uint32_t*>(a.x.colcls + 4 * w);
            // the columns of the word that matter: non-zero dots, or all four when zero dots can be selected
            for (uint32_t bm = zeros ? 0x80808080u : nonzero_bytes_sp(word); bm; bm &= bm - 1) {
              const int b = (__ffs(bm) - 1) >> 3;
              const int64_t j = 4ll * w + b;
              if (j >= a.n || j == row) continue;
              const uint32_t dot = (word >> (8 * b)) & 0xFFu, cls = (cls4 >> (8 * b)) & 0xFFu;
              if (cls >= static_cast<uint32_t>(a.n_sq)) continue;
              const int r = a.rank_tab[cls * tstride + min(dot, static_cast<uint32_t>(a.dot_max))];
              if (!pass2) {
                n_sel += r < cut;
                n_tie += r == cut;
              } else {
                bool take = r < cut;
                if (r == cut && tie_room > 0) {
                  take = true;
                  --tie_room;
                }
                if (take)
                  a.recs[out++] = static_cast<uint64_t>(j) | static_cast<uint64_t>(dot) << 32 | static_cast<uint64_t>(r) << 48;
              }
            }
          }
        }
        if (pass2) break;
        scan[threadIdx.x] = static_cast<uint32_t>(n_sel);
        scan[kSpThreads + threadIdx.x] = static_cast<uint32_t>(n_tie);
        __syncthreads();
        if (threadIdx.x < 32) {   // exclusive scans of both arrays by one warp
          uint32_t run_s = 0, run_t = 0;
          for (int i0 = 0; i0 < kSpThreads; i0 += 32) {
            const uint32_t vs = scan[i0 + threadIdx.x], vt = scan[kSpThreads + i0 + threadIdx.x];
            uint32_t is = vs, it = vt;
            for (int o = 1; o < 32; o <<= 1) {
              const uint32_t us = __shfl_up_sync(0xFFFFFFFFu, is, o), ut = __shfl_up_sync(0xFFFFFFFFu, it, o);
              if (static_cast<int>(threadIdx.x) >= o) {
                is += us;
                it += ut;
              }
            }
            scan[i0 + threadIdx.x] = run_s + is - vs;
            scan[kSpThreads + i0 + threadIdx.x] = run_t + it - vt;
            run_s += __shfl_sync(0xFFFFFFFFu, is, 31);
            run_t += __shfl_sync(0xFFFFFFFFu, it, 31);
          }
        }
        __syncthreads();
      }
      __syncthreads();
    }
  }
}

// ---- the KF threshold edges of sparse wide rows through the same join (gg_kf_emit_sparse) ----------------------------
// cos(i, j) > threshold needs dot(i, j) >= dmin[sq_i sq_j] >= 1, so only pairs that share a column can be edges.  One CTA
// per row i accumulates dot(i, j) for the j > i of its posting lists, then every thread scans whole 128-column tiles of
// the accumulator in column order (its range is at least one tile, so a candidate's rank inside its tile is a running
// count of the thread): sweep 1 tests, clears what fails, counts per tile (rowcnt) and per thread; a warp scan + one
// atomic reserves the record slots; sweep 2 writes the records {s, t, rank in tile, dot} and clears.  Same outputs as
// gg_kf_emit: unordered records + per-(block pair, row, tile) counts for gg_kf_place.
struct KfSparseArgs {
  SparseIndex x;
  int64_t n;
  int ib0, ib1, n_jblocks, per_log2;
  const int32_t* sqnorm;
  const uint16_t* dmin;
  uint8_t* rowcnt;
  uint4* recs;
  long long capacity;
  unsigned long long* status;
};

__global__ void __launch_bounds__(kSpThreads) kf_sparse_kernel(KfSparseArgs a, unsigned int* row_cursor) {
  extern __shared__ __align__(16) uint32_t sp_smem[];
  const int acc_words = static_cast<int>((a.n + 3) / 4);
  const int acc_alloc = (sp_phys(acc_words, a.per_log2) + 4) & ~3;
  uint32_t* acc = sp_smem;
  uint32_t* lists = acc + acc_alloc;
  __shared__ int s_row;
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < acc_alloc; i += kSpThreads) acc[i] = 0u;
  __syncthreads();
  const int64_t row_lo = static_cast<int64_t>(a.ib0) * GG_KF_BLOCK;
  const int64_t row_hi = min(a.n, static_cast<int64_t>(a.ib1) * GG_KF_BLOCK);
  while (true) {
    if (threadIdx.x == 0) s_row = static_cast<int>(atomicAdd(row_cursor, 1u));
    __syncthreads();
    const int64_t row = row_lo + s_row;
    __syncthreads();
    if (row >= row_hi) break;
    // posting lists of the row's columns (as sparse_accumulate), entries j > row only
    const int cnt = a.x.rnz_n[row];
    if (threadIdx.x < 32) {
      uint32_t len = 0;
      if (lane < cnt) {
        const uint32_t ent = a.x.rnz[row * kSpMaxNnz + lane];
        const uint32_t col = ent & 0xFFFFu;
        const uint32_t first = a.x.col_start[col];
        len = a.x.col_start[col + 1] - first;
        lists[3 * lane] = first;
        lists[3 * lane + 2] = ent >> 16;
      }
      uint32_t incl = len;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += u;
      }
      if (lane < cnt) lists[3 * lane + 1] = incl - len;
      if (lane == 31) {
        lists[3 * kSpMaxNnz] = static_cast<uint32_t>(cnt);
        lists[3 * kSpMaxNnz + 1] = incl;
      }
    }
    __syncthreads();
    const int n_lists = static_cast<int>(lists[3 * kSpMaxNnz]), total = static_cast<int>(lists[3 * kSpMaxNnz + 1]);
    for (int t = threadIdx.x; t < total; t += kSpThreads) {
      int e = 0;
      while (e + 1 < n_lists && lists[3 * (e + 1) + 1] <= static_cast<uint32_t>(t)) ++e;
      const uint32_t p = __ldg(a.x.post + lists[3 * e] + (t - lists[3 * e + 1]));
      const uint32_t j = p & 0xFFFFFFu;
      if (static_cast<int64_t>(j) <= row) continue;
      atomicAdd(acc + sp_phys(static_cast<int>(j >> 2), a.per_log2), ((p >> 24) * lists[3 * e + 2]) << (8 * (j & 3u)));
    }
    __syncthreads();
    const int per = 1 << a.per_log2;
    const int w_lo = min(acc_words, static_cast<int>(threadIdx.x) * per), w_hi = min(acc_words, w_lo + per);
    uint32_t* my_acc = acc + threadIdx.x;
    const uint32_t sq_i = static_cast<uint32_t>(a.sqnorm[row]);
    const int i_local = static_cast<int>(row >> 10) - a.ib0, r = static_cast<int>(row & (GG_KF_BLOCK - 1));
    // sweep 1: exact test per non-zero column; what fails is cleared; per-tile counts go to rowcnt
    int mine = 0;
    for (int t0 = w_lo; t0 < w_hi; t0 += 32) {   // one 128-column tile = 32 words
      unsigned int mask = 0;
      const int span = min(32, w_hi - t0);
      for (int i = 0; i < span; ++i) mask |= static_cast<unsigned int>(my_acc[t0 + i] != 0u) << i;
      int tile_cnt = 0;
      for (; mask; mask &= mask - 1) {
        const int w = t0 + __ffs(mask) - 1;
        uint32_t word = my_acc[w];
        for (uint32_t bm = nonzero_bytes_sp(word); bm; bm &= bm - 1) {
          const int b = (__ffs(bm) - 1) >> 3;
          const uint32_t dot = (word >> (8 * b)) & 0xFFu;
          const uint32_t sq_j = static_cast<uint32_t>(__ldg(a.sqnorm + 4 * w + b));
          if (dot >= __ldg(a.dmin + sq_i * sq_j)) ++tile_cnt;
          else word &= ~(0xFFu << (8 * b));
        }
        my_acc[w] = word;
      }
      if (tile_cnt) {
        const int tile_g = t0 >> 5;   // global tile index: column block = tile_g / 8, tile inside it = tile_g % 8
        a.rowcnt[((static_cast<long long>(i_local) * a.n_jblocks + (tile_g >> 3)) * GG_KF_BLOCK + r) * 8 + (tile_g & 7)] =
            static_cast<uint8_t>(tile_cnt);
        mine += tile_cnt;
      }
    }
    // record slots: one atomic per warp
    int incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += v;
    }
    const int warp_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    unsigned long long base = 0;
    if (warp_total) {
      if (lane == 31) base = atomicAdd(a.status, static_cast<unsigned long long>(warp_total));
      base = __shfl_sync(0xFFFFFFFFu, base, 31) + (incl - mine);
    }
    // sweep 2: what is left in the accumulator are the candidates
    if (mine) {
      for (int t0 = w_lo; t0 < w_hi; t0 += 32) {
        unsigned int mask = 0;
        const int span = min(32, w_hi - t0);
        for (int i = 0; i < span; ++i) mask |= static_cast<unsigned int>(my_acc[t0 + i] != 0u) << i;
        uint32_t rank = 0;
        for (; mask; mask &= mask - 1) {
          const int w = t0 + __ffs(mask) - 1;
          const uint32_t word = my_acc[w];
          my_acc[w] = 0u;
          for (uint32_t bm = nonzero_bytes_sp(word); bm; bm &= bm - 1) {
            const int b = (__ffs(bm) - 1) >> 3;
            if (static_cast<long long>(base) < a.capacity)
              a.recs[base] = make_uint4(static_cast<uint32_t>(row), static_cast<uint32_t>(4 * w + b), rank,
                                        (word >> (8 * b)) & 0xFFu);
            ++base;
            ++rank;
          }
        }
      }
    }
    __syncthreads();
  }
}

int sparse_per_log2(int64_t n) {
  const int64_t acc_words = (n + 3) / 4;
  int lg = 0;
  while ((static_cast<int64_t>(kSpThreads) << lg) < acc_words) ++lg;
  return lg;
}

size_t sparse_smem(int64_t n, int n_ranks, int n_sq, int pass) {
  const size_t acc_words = (static_cast<size_t>(n) + 3) / 4;
  const size_t acc_alloc = (acc_words + (acc_words >> sparse_per_log2(n)) + 4) & ~static_cast<size_t>(3);
  const size_t tail = pass == 1 ? static_cast<size_t>(n_ranks + n_sq) : 2 * kSpThreads;
  return (acc_alloc + 3 * kSpMaxNnz + 4 + tail) * sizeof(uint32_t);
}

// the tcgen05 CTA-pair Gram serves rows of 128..1024 bytes (small_k >= 4) with at most 32 norm classes
int use_tc(const KnnArgs& a, int pass, bool& yes) {
  const int impl = g_knn_impl.load();
  yes = impl != 1 && ggkf::knn_tc_eligible(a.pitch, a.n, a.n_sq, a.dot_max, a.n_ranks, pass);
  GG_REQUIRE(impl != 2 || yes, GG_ERR_ARG, "the tcgen05 kNN kernel cannot serve this shape (row pitch, classes, shared memory)");
  return GG_OK;
}

ggkf::KnnTcArgs tc_args(const KnnArgs& a) {
  ggkf::KnnTcArgs t{};
  t.counts = a.counts;
  t.d = a.pitch;
  t.n = a.n;
  t.row_begin = a.row_begin;
  t.row_end = a.row_end;
  t.sqnorm = a.sqnorm;
  t.sqid = a.sqid;
  t.sq_max = a.sq_max;
  t.rank_tab = a.rank_tab;
  t.n_sq = a.n_sq;
  t.dot_max = a.dot_max;
  t.n_ranks = a.n_ranks;
  t.hist16 = a.n <= 65536 ? 1 : 0;
  return t;
}

}  // namespace

extern "C" int gg_knn_set_impl(int impl) {
  GG_REQUIRE(impl >= 0 && impl <= 3, GG_ERR_ARG, "impl must be 0 (auto), 1 (mma.sync), 2 (tcgen05) or 3 (sparse)");
  g_knn_impl.store(impl);
  return GG_OK;
}

/* pass 1: hist[row_end - row_begin][n_ranks] of the row shard.  hist_mode: -1 auto, 0 global atomics, 1 shared memory
 * 16-bit (n <= 65536), 2 shared memory 32-bit; a forced shared-memory mode that does not fit is an error. */
extern "C" int gg_knn_count(const uint8_t* counts, int pitch, int64_t n, int64_t row_begin, int64_t row_end,
                            const int32_t* sqnorm, const uint8_t* sqid, int sq_max, const uint16_t* rank_tab, int n_sq,
                            int dot_max, int n_ranks, int hist_mode, uint32_t* hist, gg_stream_t st) {
  KnnArgs a;
  const int rc =
      fill_args(a, counts, pitch, n, row_begin, row_end, sqnorm, sqid, sq_max, rank_tab, n_sq, dot_max, n_ranks);
  if (rc != GG_OK) return rc;
  GG_REQUIRE(hist, GG_ERR_ARG, "null histogram");
  GG_REQUIRE(hist_mode >= -1 && hist_mode <= 2, GG_ERR_ARG, "hist_mode must be -1, 0, 1 or 2");
  if (hist_mode < 0) {  // a forced histogram mode names the mma.sync kernel
    bool tc = false;
    if (int rc2 = use_tc(a, 1, tc)) return rc2;
    if (tc) {
      ggkf::KnnTcArgs t = tc_args(a);
      t.hist = hist;
      return ggkf::knn_tc_launch(t, 1, nullptr, gg::as_stream(st));
    }
  }
  const size_t base = knn_base_smem(a.kpad, a.chunk, n_sq, dot_max);
  const bool fits32 = base + hist_smem(kHistSmem32, n_ranks) <= kSmemLimit;
  const bool fits16 = n <= 65536 && base + hist_smem(kHistSmem16, n_ranks) <= kSmemLimit;
  if (hist_mode < 0) hist_mode = fits32 ? kHistSmem32 : fits16 ? kHistSmem16 : kHistGlobal;
  GG_REQUIRE(hist_mode != kHistSmem32 || fits32, GG_ERR_ARG, "32-bit shared-memory histograms do not fit");
  GG_REQUIRE(hist_mode != kHistSmem16 || fits16, GG_ERR_ARG, "16-bit shared-memory histograms do not fit");
  a.hist_mode = hist_mode;
  a.hist = hist;
  cudaStream_t s = gg::as_stream(st);
  if (hist_mode == kHistGlobal)
    GG_CUDA_CHECK(cudaMemsetAsync(hist, 0, static_cast<size_t>(row_end - row_begin) * n_ranks * sizeof(uint32_t), s));
  const size_t smem = base + hist_smem(hist_mode, n_ranks);
  if (hist_mode == kHistSmem32) return launch_knn<1, kHistSmem32>(a, smem, s);
  if (hist_mode == kHistSmem16) return launch_knn<1, kHistSmem16>(a, smem, s);
  return launch_knn<1, kHistGlobal>(a, smem, s);
}

extern "C" size_t gg_knn_plan_workspace_bytes(int64_t n) {
  size_t scan = 0;
  cub::TransformInputIterator<int64_t, ToI64, const uint32_t*> it(nullptr, ToI64());
  cub::DeviceScan::ExclusiveSum(nullptr, scan, it, static_cast<int64_t*>(nullptr), static_cast<int>(n + 1));
  return gg::align_up(scan, 256) + 2 * gg::align_up(static_cast<size_t>(n + 1) * 4, 256) + 256;
}

/* plan for the rows [row_begin, row_end) (per-row arrays are shard-local, m = row_end - row_begin entries): cut[m],
 * tie_left[m], row_offsets[m + 1] (the last entry = number of edges of the shard), qmax (device int32[1], L2 only: the
 * largest kept squared distance in integer units); turns hist into the per-(row, rank) cursors of gg_knn_write.
 * phase 0: everything.  Row-sharded builds need the GLOBAL qmax: phase 1 stops after the cut ranks and the shard's
 * qmax, the caller maximises qmax over the shards, phase 2 finishes (same workspace, untouched in between). */
extern "C" int gg_knn_plan(int64_t n, int64_t row_begin, int64_t row_end, int n_ranks, int k_nn, int metric,
                           const int32_t* sqnorm, const int32_t* rank_q, uint32_t* hist, int32_t* cut, int32_t* tie_left,
                           int64_t* row_offsets, int32_t* qmax, int phase, void* workspace, size_t workspace_bytes,
                           gg_stream_t st) {
  GG_REQUIRE(n >= 1 && n < (int64_t(1) << 31), GG_ERR_ARG, "n outside [1, 2^31)");
  GG_REQUIRE(row_begin >= 0 && row_begin < row_end && row_end <= n, GG_ERR_ARG, "empty or out-of-range row shard");
  GG_REQUIRE(n_ranks >= 1 && n_ranks < 0xFFFF, GG_ERR_ARG, "n_ranks outside [1, 65535)");
  GG_REQUIRE(k_nn >= 1 && k_nn < n, GG_ERR_ARG, "k must be in [1, n - 1]");
  GG_REQUIRE(metric == 0 || metric == 1, GG_ERR_ARG, "metric must be 0 (IP) or 1 (L2)");
  GG_REQUIRE(phase >= 0 && phase <= 2, GG_ERR_ARG, "phase must be 0, 1 or 2");
  GG_REQUIRE(hist && cut && tie_left && row_offsets && qmax && sqnorm && workspace, GG_ERR_ARG, "null argument");
  GG_REQUIRE(metric == 0 || rank_q, GG_ERR_ARG, "L2 needs rank_q");
  const int64_t m = row_end - row_begin;
  GG_REQUIRE(workspace_bytes >= gg_knn_plan_workspace_bytes(m), GG_ERR_ARG, "workspace too small");
  cudaStream_t s = gg::as_stream(st);
  gg::Carver cv(workspace);
  uint32_t* before = cv.take<uint32_t>(m + 1);
  uint32_t* rowcnt = cv.take<uint32_t>(m + 1);
  void* temp = cv.take<char>(0);
  size_t temp_bytes = workspace_bytes - cv.used();
  const unsigned grid = static_cast<unsigned>((m + 127) / 128);
  const int32_t* sq = sqnorm + row_begin;
  if (phase != 2) {
    GG_CUDA_CHECK(cudaMemsetAsync(qmax, 0, sizeof(int32_t), s));
    cut_kernel<<<grid, 128, 0, s>>>(hist, m, n_ranks, k_nn, metric, sq, rank_q, cut, tie_left, before, qmax);
    GG_CUDA_CHECK(cudaGetLastError());
  }
  if (phase != 1) {
    GG_CUDA_CHECK(cudaMemsetAsync(rowcnt + m, 0, sizeof(uint32_t), s));
    rows_kernel<<<grid, 128, 0, s>>>(m, n_ranks, metric, sq, rank_q, qmax, cut, tie_left, before, hist, rowcnt);
    GG_CUDA_CHECK(cudaGetLastError());
    cub::TransformInputIterator<int64_t, ToI64, const uint32_t*> it(rowcnt, ToI64());
    GG_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, it, row_offsets, static_cast<int>(m + 1), s));
    cursor_kernel<<<grid, 128, 0, s>>>(m, n_ranks, row_offsets, hist);
    GG_CUDA_CHECK(cudaGetLastError());
  }
  return GG_OK;
}

/* 64-bit words of scratch gg_knn_write needs for a shard of n_rows rows and n_edges edges (the kernel that will serve
 * the shape decides: the tcgen05 kernel stages every row three times over, plus three counters per row) */
extern "C" size_t gg_knn_write_scratch_words(int pitch, int64_t n, int n_sq, int dot_max, int n_ranks, int64_t n_rows,
                                             int64_t n_edges) {
  if (n_rows < 0 || n_edges < 0) return 0;
  const bool tc = g_knn_impl.load() != 1 && ggkf::knn_tc_eligible(pitch, n, n_sq, dot_max, n_ranks, 2);
  return tc ? 3 * static_cast<size_t>(n_edges) + (3 * static_cast<size_t>(n_rows) + 1) / 2 : static_cast<size_t>(n_edges);
}

/* pass 2 + order for the row shard: edge_src / edge_dst / weight[row_offsets[m]], grouped by source row, most similar
 * first.  recs: scratch of gg_knn_write_scratch_words() 64-bit words (`recs_words` of them). */
extern "C" int gg_knn_write(const uint8_t* counts, int pitch, int64_t n, int64_t row_begin, int64_t row_end,
                            const int32_t* sqnorm, const uint8_t* sqid, int sq_max, const uint16_t* rank_tab, int n_sq,
                            int dot_max, int n_ranks, int metric,
                            uint32_t* cursors, const int32_t* cut, const int32_t* tie_left,
                            const int64_t* row_offsets, const int32_t* qmax, uint64_t* recs, int64_t recs_words,
                            int64_t* edge_src, int64_t* edge_dst, float* weight, int64_t capacity, gg_stream_t st) {
  KnnArgs a;
  const int rc =
      fill_args(a, counts, pitch, n, row_begin, row_end, sqnorm, sqid, sq_max, rank_tab, n_sq, dot_max, n_ranks);
  if (rc != GG_OK) return rc;
  GG_REQUIRE(metric == 0 || metric == 1, GG_ERR_ARG, "metric must be 0 (IP) or 1 (L2)");
  GG_REQUIRE(cursors && cut && tie_left && row_offsets && qmax, GG_ERR_ARG, "null plan");
  GG_REQUIRE(capacity >= 0 && (capacity == 0 || (recs && edge_src && edge_dst && weight)), GG_ERR_ARG, "null output");
  GG_REQUIRE(capacity < (int64_t(1) << 32), GG_ERR_ARG, "more than 2^32 - 1 edges");
  if (capacity == 0) return GG_OK;
  a.cut = cut;
  a.tie_left = tie_left;
  a.row_off = row_offsets;
  a.recs = recs;
  cudaStream_t s = gg::as_stream(st);
  bool tc = false;
  if (int rc3 = use_tc(a, 2, tc)) return rc3;
  const int64_t m = row_end - row_begin;
  GG_REQUIRE(recs_words >= static_cast<int64_t>(gg_knn_write_scratch_words(pitch, n, n_sq, dot_max, n_ranks, m, capacity)),
             GG_ERR_WORKSPACE, "recs scratch smaller than gg_knn_write_scratch_words()");
  OrderArgs o{recs, row_offsets, m, row_begin, n_ranks, cursors, sqnorm, metric, qmax, edge_src, edge_dst, weight};
  const unsigned grid = static_cast<unsigned>((m * 32 + 255) / 256);
  if (tc) {
    ggkf::KnnTcArgs t = tc_args(a);
    t.cut = cut;
    t.tie_left = tie_left;
    t.row_off = row_offsets;
    t.recs = recs;
    t.third_cnt = reinterpret_cast<uint32_t*>(recs + 3 * capacity);
    if (int rc2 = ggkf::knn_tc_launch(t, 2, nullptr, s)) return rc2;
    Order3Args o3{o, t.third_cnt, cut, tie_left};
    order3_kernel<<<grid, 256, 0, s>>>(o3);
    GG_CUDA_CHECK(cudaGetLastError());
    return GG_OK;
  }
  if (int rc2 = launch_knn<2, kHistGlobal>(a, knn_base_smem(a.kpad, a.chunk, n_sq, dot_max), s)) return rc2;
  order_kernel<<<grid, 256, 0, s>>>(o);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}

/* ---- sparse (inverted-index) flavour of the two passes: rows with at most 16 non-zero counts, up to 200,000 nodes ------ */
extern "C" int gg_knn_sparse_eligible(int pitch, int64_t n, int n_sq, int n_ranks) {
  if (pitch < 64 || pitch > 1024 || pitch % 4 != 0 || n < 2 || n > kSpMaxNodes || n_sq < 1 || n_sq > 32) return 0;
  const size_t worst = sparse_smem(n, n_ranks, n_sq, 1) > sparse_smem(n, n_ranks, n_sq, 2) ? sparse_smem(n, n_ranks, n_sq, 1)
                                                                                             : sparse_smem(n, n_ranks, n_sq, 2);
  return worst <= kSmemLimit ? 1 : 0;
}

extern "C" size_t gg_knn_sparse_index_bytes(int64_t n, int pitch) {
  if (n < 1 || pitch < 4 || pitch > 1024) return 0;
  return sparse_carve(nullptr, n, pitch).total;
}

/* status (device int64[2]): [0] the largest non-zero count of a row when it exceeds 16 (the index is then unusable) */
extern "C" int gg_knn_sparse_index(const uint8_t* counts, int pitch, int64_t n, const int32_t* sqnorm, const uint8_t* sqid,
                                   int sq_max, int n_sq, void* index, size_t index_bytes, int64_t* status, gg_stream_t st) {
  GG_REQUIRE(counts && sqnorm && sqid && index && status, GG_ERR_ARG, "null argument");
  GG_REQUIRE(n >= 1 && n <= kSpMaxNodes && pitch >= 4 && pitch <= 1024 && pitch % 4 == 0, GG_ERR_ARG, "bad shape");
  GG_REQUIRE(n_sq >= 1 && n_sq <= 32 && sq_max >= 0 && sq_max <= 254, GG_ERR_ARG, "bad class table");
  SparseIndex x = sparse_carve(index, n, pitch);
  GG_REQUIRE(index_bytes >= x.total, GG_ERR_WORKSPACE, "index workspace too small");
  cudaStream_t s = gg::as_stream(st);
  GG_CUDA_CHECK(cudaMemsetAsync(status, 0, 2 * sizeof(int64_t), s));
  GG_CUDA_CHECK(cudaMemsetAsync(x.col_start, 0, (pitch + 1) * sizeof(uint32_t), s));
  GG_CUDA_CHECK(cudaMemsetAsync(x.colcount, 0, kKnnMaxSq * sizeof(uint32_t), s));
  GG_CUDA_CHECK(cudaMemsetAsync(x.key_in, 0xFF, static_cast<size_t>(n) * kSpMaxNnz * sizeof(uint16_t), s));
  GG_CUDA_CHECK(cudaMemsetAsync(x.rnz_n + n, 0, 1, s));
  const unsigned wgrid = static_cast<unsigned>((n * 32 + 255) / 256), grid = static_cast<unsigned>((n + 255) / 256);
  sparse_rows_kernel<<<wgrid, 256, 0, s>>>(counts, pitch, n, sqnorm, sqid, sq_max, n_sq, x,
                                           reinterpret_cast<unsigned long long*>(status));
  GG_CUDA_CHECK(cudaGetLastError());
  sparse_scan_kernel<<<1, 1024, 0, s>>>(x.col_start, pitch);
  GG_CUDA_CHECK(cudaGetLastError());
  size_t tb = x.cub_bytes;
  cub::TransformInputIterator<uint32_t, U8ToU32, const uint8_t*> it(x.rnz_n, U8ToU32());
  GG_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(x.cub_temp, tb, it, x.row_start, static_cast<int>(n + 1), s));
  sparse_pairs_kernel<<<grid, 256, 0, s>>>(n, x);
  GG_CUDA_CHECK(cudaGetLastError());
  tb = x.cub_bytes;
  int bits = 1;
  while ((1 << bits) < pitch) ++bits;   // pads (0xFFFF) still sort behind every real column on one more bit
  GG_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(x.cub_temp, tb, x.key_in, x.key_out, x.val_in, x.post,
                                                static_cast<int>(n * kSpMaxNnz), 0, bits + 1, s));
  return GG_OK;
}

static int sparse_args(SparseArgs& a, void* index, size_t index_bytes, int pitch, int64_t n, int64_t row_begin,
                       int64_t row_end, const int32_t* sqnorm, const uint16_t* rank_tab, int n_sq, int dot_max,
                       int n_ranks) {
  GG_REQUIRE(index && sqnorm && rank_tab, GG_ERR_ARG, "null argument");
  GG_REQUIRE(n >= 2 && n <= kSpMaxNodes && pitch >= 4 && pitch <= 1024, GG_ERR_ARG, "bad shape");
  GG_REQUIRE(row_begin >= 0 && row_begin < row_end && row_end <= n, GG_ERR_ARG, "bad row shard");
  GG_REQUIRE(n_sq >= 1 && n_sq <= 32 && dot_max >= 0 && n_ranks >= 1 && n_ranks < 0xFFFF, GG_ERR_ARG, "bad class table");
  a = SparseArgs{};
  a.x = sparse_carve(index, n, pitch);
  GG_REQUIRE(index_bytes >= a.x.total, GG_ERR_WORKSPACE, "index workspace too small");
  a.n = n;
  a.row_begin = row_begin;
  a.row_end = row_end;
  a.sqnorm = sqnorm;
  a.rank_tab = rank_tab;
  a.n_sq = n_sq;
  a.dot_max = dot_max;
  a.n_ranks = n_ranks;
  a.per_log2 = sparse_per_log2(n);
  return GG_OK;
}

template <int PASS>
static int launch_sparse(const SparseArgs& a, cudaStream_t s) {
  const size_t smem = sparse_smem(a.n, a.n_ranks, a.n_sq, PASS);
  GG_REQUIRE(smem <= kSmemLimit, GG_ERR_ARG, "accumulator does not fit shared memory");
  GG_CUDA_CHECK(cudaFuncSetAttribute(knn_sparse_kernel<PASS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
  int per_sm = 0;
  GG_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, knn_sparse_kernel<PASS>, kSpThreads, smem));
  if (per_sm < 1) per_sm = 1;
  const int64_t rows = a.row_end - a.row_begin;
  const int64_t cap = static_cast<int64_t>(gg::sm_count()) * per_sm;
  unsigned int* cursor = a.x.cursor + (PASS - 1);
  GG_CUDA_CHECK(cudaMemsetAsync(cursor, 0, sizeof(unsigned int), s));
  knn_sparse_kernel<PASS><<<static_cast<unsigned>(rows < cap ? rows : cap), kSpThreads, smem, s>>>(a, cursor);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}

extern "C" int gg_knn_sparse_count(void* index, size_t index_bytes, int pitch, int64_t n, int64_t row_begin,
                                   int64_t row_end, const int32_t* sqnorm, const uint16_t* rank_tab, int n_sq, int dot_max,
                                   int n_ranks, uint32_t* hist, gg_stream_t st) {
  SparseArgs a;
  if (int rc = sparse_args(a, index, index_bytes, pitch, n, row_begin, row_end, sqnorm, rank_tab, n_sq, dot_max, n_ranks))
    return rc;
  GG_REQUIRE(hist, GG_ERR_ARG, "null histogram");
  a.hist = hist;
  return launch_sparse<1>(a, gg::as_stream(st));
}

extern "C" int gg_knn_sparse_write(void* index, size_t index_bytes, int pitch, int64_t n, int64_t row_begin,
                                   int64_t row_end, const int32_t* sqnorm, const uint16_t* rank_tab, int n_sq, int dot_max,
                                   int n_ranks, int metric, uint32_t* cursors, const int32_t* cut, const int32_t* tie_left,
                                   const int64_t* row_offsets, const int32_t* qmax, uint64_t* recs, int64_t* edge_src,
                                   int64_t* edge_dst, float* weight, int64_t capacity, gg_stream_t st) {
  SparseArgs a;
  if (int rc = sparse_args(a, index, index_bytes, pitch, n, row_begin, row_end, sqnorm, rank_tab, n_sq, dot_max, n_ranks))
    return rc;
  GG_REQUIRE(metric == 0 || metric == 1, GG_ERR_ARG, "metric must be 0 (IP) or 1 (L2)");
  GG_REQUIRE(cursors && cut && tie_left && row_offsets && qmax, GG_ERR_ARG, "null plan");
  GG_REQUIRE(capacity >= 0 && (capacity == 0 || (recs && edge_src && edge_dst && weight)), GG_ERR_ARG, "null output");
  GG_REQUIRE(capacity < (int64_t(1) << 32), GG_ERR_ARG, "more than 2^32 - 1 edges");
  if (capacity == 0) return GG_OK;
  a.cut = cut;
  a.tie_left = tie_left;
  a.row_off = row_offsets;
  a.recs = recs;
  cudaStream_t s = gg::as_stream(st);
  if (int rc = launch_sparse<2>(a, s)) return rc;
  const int64_t m = row_end - row_begin;
  OrderArgs o{recs, row_offsets, m, row_begin, n_ranks, cursors, sqnorm, metric, qmax, edge_src, edge_dst, weight};
  order_kernel<<<static_cast<unsigned>((m * 32 + 255) / 256), 256, 0, s>>>(o);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}

/* KF threshold edges of sparse wide rows (an index of gg_knn_sparse_index over the same counts): the outputs of gg_kf_emit
 * -- unordered records + per-(block pair, row, tile) counts, kf_status[0] = candidates found (may exceed rec_capacity:
 * score again with more room) -- without forming the zero dots.  dmin as for gg_kf_emit (p_max + 1 entries). */
extern "C" int gg_kf_emit_sparse(void* index, size_t index_bytes, int pitch, int64_t n, int iblock_begin, int iblock_end,
                                 const int32_t* sqnorm, const uint16_t* dmin, gg_kf_rec_t* recs, int64_t rec_capacity,
                                 uint64_t* rowcnt, int64_t* kf_status, gg_stream_t st) {
  GG_REQUIRE(index && sqnorm && dmin && rowcnt && kf_status, GG_ERR_ARG, "null argument");
  GG_REQUIRE(n >= 2 && n <= kSpMaxNodes && pitch >= 4 && pitch <= 1024, GG_ERR_ARG, "bad shape");
  const int nb = static_cast<int>((n + GG_KF_BLOCK - 1) / GG_KF_BLOCK);
  GG_REQUIRE(iblock_begin >= 0 && iblock_begin <= iblock_end && iblock_end <= nb, GG_ERR_ARG, "bad row-block shard");
  GG_REQUIRE(rec_capacity >= 0 && (rec_capacity == 0 || recs), GG_ERR_ARG, "record buffer");
  cudaStream_t s = gg::as_stream(st);
  GG_CUDA_CHECK(cudaMemsetAsync(kf_status, 0, 4 * sizeof(int64_t), s));
  const size_t rc_elems = gg_kf_rowcnt_elems(n, iblock_begin, iblock_end);
  if (rc_elems == 0) return GG_OK;
  GG_CUDA_CHECK(cudaMemsetAsync(rowcnt, 0, rc_elems * sizeof(uint64_t), s));
  KfSparseArgs a{};
  a.x = sparse_carve(index, n, pitch);
  GG_REQUIRE(index_bytes >= a.x.total, GG_ERR_WORKSPACE, "index workspace too small");
  a.n = n;
  a.ib0 = iblock_begin;
  a.ib1 = iblock_end;
  a.n_jblocks = nb;
  a.per_log2 = sparse_per_log2(n) < 5 ? 5 : sparse_per_log2(n);   // a thread scans whole 128-column tiles
  a.sqnorm = sqnorm;
  a.dmin = dmin;
  a.rowcnt = reinterpret_cast<uint8_t*>(rowcnt);
  a.recs = reinterpret_cast<uint4*>(recs);
  a.capacity = rec_capacity;
  a.status = reinterpret_cast<unsigned long long*>(kf_status);
  const size_t acc_words = (static_cast<size_t>(n) + 3) / 4;
  const size_t smem = (((acc_words + (acc_words >> a.per_log2) + 4) & ~static_cast<size_t>(3)) + 3 * kSpMaxNnz + 4) *
                      sizeof(uint32_t);
  GG_REQUIRE(smem <= kSmemLimit, GG_ERR_ARG, "accumulator does not fit shared memory");
  GG_CUDA_CHECK(cudaFuncSetAttribute(kf_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int per_sm = 0;
  GG_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kf_sparse_kernel, kSpThreads, smem));
  if (per_sm < 1) per_sm = 1;
  const int64_t row_lo = static_cast<int64_t>(iblock_begin) * GG_KF_BLOCK;
  const int64_t row_hi = n < static_cast<int64_t>(iblock_end) * GG_KF_BLOCK ? n : static_cast<int64_t>(iblock_end) * GG_KF_BLOCK;
  const int64_t rows = row_hi - row_lo;
  if (rows <= 0) return GG_OK;
  const int64_t cap = static_cast<int64_t>(gg::sm_count()) * per_sm;
  unsigned int* cursor = a.x.cursor + 2;
  GG_CUDA_CHECK(cudaMemsetAsync(cursor, 0, sizeof(unsigned int), s));
  kf_sparse_kernel<<<static_cast<unsigned>(rows < cap ? rows : cap), kSpThreads, smem, s>>>(a, cursor);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}
