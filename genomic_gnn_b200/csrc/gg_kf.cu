// K4-K6: KF cosine-similarity edges -- SIMT scoring kernels, ordered
// compaction and the global top-n select.
//
// Replaces cosine_similarity_to_edge_index_weight (reference
// gnn_utils.py:240-322).  Feature rows are small integer counts, so
// cos(i,j) = dot / sqrt(sq_i sq_j) with integer dot and the strict test
// cos > threshold is evaluated exactly as dot >= dmin[sq_i * sq_j].
//
//   kf_dp4a_kernel<W>   D = 4W <= 64 columns: the row lives in registers, the
//                       1024-column panel of block J in shared memory; 4 dp4a
//                       per pair at D = 16.  ALU-issue bound (K = 16 is too
//                       shallow for the tensor pipe to matter).
//   kf_generic_kernel   any D that is a multiple of 64: register-tiled 128x32
//                       scores per pass, dp4a over 64-byte slabs.  Validation
//                       and fall-back path for the tcgen05 kernel (gg_kf_mma.cu).
#include <cub/cub.cuh>

#include "gg_kf_emit.cuh"

namespace ggkf {
int emit_tcgen05(const KfArgs& a, int p_max, cudaStream_t s);  // gg_kf_mma.cu
int emit_tcgen05_pair(const KfArgs& a, cudaStream_t s);  // gg_kf_mma.cu (cta_group::2)
}

namespace {

using namespace ggkf;

struct Unit {
  int i_block, i_local, j_block, strip;
  bool active;
};

__device__ __forceinline__ Unit decode_unit(const KfArgs& a) {
  Unit u;
  const long long id = blockIdx.x;
  u.strip = static_cast<int>(id % kStripsPerBlock);
  const long long rest = id / kStripsPerBlock;
  u.j_block = static_cast<int>(rest % a.n_jblocks);
  u.i_local = static_cast<int>(rest / a.n_jblocks);
  u.i_block = a.ib0 + u.i_local;
  u.active = u.j_block >= u.i_block && static_cast<long long>(u.i_block) * kBlock + u.strip * kStrip < a.n;
  return u;
}

__device__ __forceinline__ bool fast_pass(uint32_t dot, uint32_t scale_i, uint32_t sq_j, int shift) {
  return ((dot * dot) << shift) > scale_i * sq_j;
}

// ---------------------------------------------------------------------------- D <= 64
// Exact threshold without arithmetic: for the row a thread owns, the smallest passing dot depends only on the
// column's squared norm, so each CTA tabulates thr_tab[sq][row] = dmin[sq_row * sq] once (<= 226 x 128 bytes) and
// the per-pair test is one LDS.U8 + one compare -- no conservative pre-filter, no refine pass.
using DeepStage = WarpStageT<256>;  // ~20 records per (row strip, column block) warp-chunk: flush every ~12 chunks

template <int W>
__global__ void __launch_bounds__(kStrip) kf_dp4a_kernel(KfArgs a, int sq_max) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* colw = reinterpret_cast<uint32_t*>(smem_raw);                         // [kBlock][W]
  uint16_t* sqoff = reinterpret_cast<uint16_t*>(colw + kBlock * W);               // [kBlock] sq_col * kStrip
  DeepStage* stages = reinterpret_cast<DeepStage*>(sqoff + kBlock);               // [4]
  uint8_t* thr_tab = reinterpret_cast<uint8_t*>(stages + kStrip / 32);            // [sq_max + 1][kStrip]

  const Unit un = decode_unit(a);
  if (!un.active) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  DeepStage& ws = stages[warp];
  if (lane == 0) ws.fill = 0;
  __syncwarp();

  const long long col0 = static_cast<long long>(un.j_block) * kBlock;
  const int cols_in_block = static_cast<int>(min(static_cast<long long>(kBlock), a.n - col0));
  for (int c = tid; c < kBlock; c += kStrip) {
    const bool ok = c < cols_in_block;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.counts + (col0 + c) * a.d);
#pragma unroll
    for (int w = 0; w < W; ++w) colw[c * W + w] = ok ? __ldg(src + w) : 0u;
    sqoff[c] = ok ? static_cast<uint16_t>(__ldg(a.sqnorm + col0 + c) * kStrip) : 0;
  }

  const int r = un.strip * kStrip + tid;  // block-local row
  const long long s = static_cast<long long>(un.i_block) * kBlock + r;
  const bool row_ok = s < a.n;
  uint32_t row[W];
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.counts + (row_ok ? s : 0) * a.d);
#pragma unroll
    for (int w = 0; w < W; ++w) row[w] = row_ok ? __ldg(src + w) : 0u;
    const uint32_t sq_i = row_ok ? static_cast<uint32_t>(__ldg(a.sqnorm + s)) : 0u;
    for (int q = 0; q <= sq_max; ++q) {
      const uint32_t need = row_ok ? __ldg(a.dmin + sq_i * q) : 255u;  // dmin[0] is "never"
      thr_tab[q * kStrip + tid] = static_cast<uint8_t>(need > 255u ? 255u : need);
    }
  }
  __syncthreads();
  const uint8_t* my_thr = thr_tab + tid;

  const bool diag = un.i_block == un.j_block;
  uint32_t tile_cnt = 0;
  // diagonal block: columns <= r contribute nothing, skip whole tiles below the strip
  const int c_begin = diag ? un.strip * kStrip : 0;
  for (int c0 = c_begin; c0 < cols_in_block; c0 += 32) {
    uint32_t mask = 0;
#pragma unroll
    for (int b = 0; b < 32; ++b) {
      const uint32_t* cw = colw + (c0 + b) * W;
      uint32_t dot = 0;
#pragma unroll
      for (int w = 0; w < W; ++w) dot = __dp4a(row[w], cw[w], dot);
      mask |= static_cast<uint32_t>(dot >= my_thr[sqoff[c0 + b]]) << b;
    }
    mask &= chunk_col_mask(cols_in_block, c0);
    if (diag) mask &= chunk_diag_mask(r, c0);
    // narrow rows: the records carry no dot, K6 recomputes it from the two short rows
    if (__any_sync(0xFFFFFFFFu, mask != 0))  // (the dot field of these records is never read)
      stage_commit(ws, a.out, lane, mask, static_cast<uint32_t>(s), static_cast<uint32_t>(col0 + c0), tile_cnt);
    if (((c0 + 32) & (kTile - 1)) == 0 || c0 + 32 >= cols_in_block) {  // end of a 128-column tile
      if (row_ok) a.rowcnt[rowcnt_index(un.i_local, a.n_jblocks, un.j_block, r, c0 / kTile)] = static_cast<uint8_t>(tile_cnt);
      tile_cnt = 0;
    }
  }
  stage_flush(ws, a.out, lane);
}

// ---------------------------------------------------------------------------- any D % 64 == 0
constexpr int kSlabWords = 16;  // 64 bytes of a row per pass

__global__ void __launch_bounds__(kStrip) kf_generic_kernel(KfArgs a) {
  __shared__ __align__(16) uint32_t colw[32 * kSlabWords];
  __shared__ uint16_t sqn[32];
  __shared__ WarpStage stages[kStrip / 32];

  const Unit un = decode_unit(a);
  if (!un.active) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  WarpStage& ws = stages[warp];
  if (lane == 0) ws.fill = 0;
  __syncwarp();

  const long long col0 = static_cast<long long>(un.j_block) * kBlock;
  const int cols_in_block = static_cast<int>(min(static_cast<long long>(kBlock), a.n - col0));
  const int r = un.strip * kStrip + tid;
  const long long s = static_cast<long long>(un.i_block) * kBlock + r;
  const bool row_ok = s < a.n;
  uint32_t sq_i = 0, scale_i = 0;
  if (row_ok) {
    sq_i = static_cast<uint32_t>(__ldg(a.sqnorm + s));
    scale_i = __ldg(a.fast_scale + sq_i);
  }
  const uint4* row_ptr = reinterpret_cast<const uint4*>(a.counts + (row_ok ? s : 0) * a.d);
  const int n_slabs = a.d / (4 * kSlabWords);
  const bool diag = un.i_block == un.j_block;
  uint32_t tile_cnt = 0;
  const int c_begin = diag ? un.strip * kStrip : 0;

  for (int c0 = c_begin; c0 < cols_in_block; c0 += 32) {
    uint32_t acc[32];
#pragma unroll
    for (int b = 0; b < 32; ++b) acc[b] = 0;
    for (int slab = 0; slab < n_slabs; ++slab) {
      __syncthreads();
      {  // 128 threads x one uint4 = 32 columns x 64 bytes
        const int c = tid >> 2, part = tid & 3;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (c0 + c < cols_in_block)
          v = __ldg(reinterpret_cast<const uint4*>(a.counts + (col0 + c0 + c) * a.d) + slab * 4 + part);
        reinterpret_cast<uint4*>(colw)[c * 4 + part] = v;
        if (slab == 0 && tid < 32)
          sqn[tid] = (c0 + tid < cols_in_block) ? static_cast<uint16_t>(__ldg(a.sqnorm + col0 + c0 + tid)) : 0;
      }
      __syncthreads();
      uint32_t row[kSlabWords];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 v = row_ok ? __ldg(row_ptr + slab * 4 + q) : make_uint4(0, 0, 0, 0);
        row[4 * q] = v.x; row[4 * q + 1] = v.y; row[4 * q + 2] = v.z; row[4 * q + 3] = v.w;
      }
#pragma unroll
      for (int b = 0; b < 32; ++b) {
#pragma unroll
        for (int w = 0; w < kSlabWords; ++w) acc[b] = __dp4a(row[w], colw[b * kSlabWords + w], acc[b]);
      }
    }
    uint32_t mask = 0;
#pragma unroll
    for (int b = 0; b < 32; ++b) mask |= static_cast<uint32_t>(fast_pass(acc[b], scale_i, sqn[b], a.fast_shift)) << b;
    mask &= chunk_col_mask(cols_in_block, c0);
    if (diag) mask &= chunk_diag_mask(r, c0);
    if (!row_ok) mask = 0;
    if (__any_sync(0xFFFFFFFFu, mask != 0)) {
      uint32_t exact = 0;
#pragma unroll
      for (int b = 0; b < 32; ++b) {
        if ((mask >> b) & 1u) {
          if (acc[b] >= __ldg(a.dmin + sq_i * sqn[b])) {
            exact |= 1u << b;
            stage_set_dot(ws, b, lane, acc[b]);
          }
        }
      }
      stage_commit(ws, a.out, lane, exact, static_cast<uint32_t>(s), static_cast<uint32_t>(col0 + c0), tile_cnt);
    }
    if (((c0 + 32) & (kTile - 1)) == 0 || c0 + 32 >= cols_in_block) {
      if (row_ok) a.rowcnt[rowcnt_index(un.i_local, a.n_jblocks, un.j_block, r, c0 / kTile)] = static_cast<uint8_t>(tile_cnt);
      tile_cnt = 0;
    }
  }
  stage_flush(ws, a.out, lane);
}

template <int W>
int launch_dp4a(const KfArgs& a, long long units, int sq_max, cudaStream_t s) {
  const size_t smem = static_cast<size_t>(kBlock) * W * 4 + kBlock * 2 + (kStrip / 32) * sizeof(DeepStage) +
                      static_cast<size_t>(sq_max + 1) * kStrip;
  GG_CUDA_CHECK(cudaFuncSetAttribute(kf_dp4a_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
  kf_dp4a_kernel<W><<<static_cast<unsigned>(units), kStrip, smem, s>>>(a, sq_max);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}

// ---------------------------------------------------------------------------- ordered compaction
struct RowTotal {  // candidates of one (block pair, row): sum of its 8 packed tile counts
  __host__ __device__ unsigned long long operator()(unsigned long long v) const { return byte_sum(v); }
};

// Two steps instead of one scatter: a record's five output fields would otherwise be five scattered sector
// writes.  place_kernel scatters ONE packed 8-byte word per record into emission order, expand_kernel then
// streams that array and writes every output array fully coalesced.
__global__ void __launch_bounds__(256) place_kernel(const gg_kf_rec_t* recs, long long n_recs,
                                                    const unsigned long long* offsets,
                                                    const unsigned long long* rowcnt, int ib0, int n_jblocks,
                                                    const uint8_t* counts, int pitch, uint2* ordered, int words) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n_recs;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 rc = __ldg(reinterpret_cast<const uint4*>(recs) + i);
    if (rc.x == kNoRec) continue;  // padding of a reserved slot block (block_commit)
    const uint32_t s = rc.x, t = rc.y;
    const long long word = rowcnt_word(static_cast<int>(s / kBlock) - ib0, n_jblocks, static_cast<int>(t / kBlock),
                                       static_cast<int>(s % kBlock));
    const int tile = static_cast<int>((t % kBlock) / kTile);
    const unsigned long long earlier_tiles = __ldg(rowcnt + word) & ((1ull << (8 * tile)) - 1ull);
    const unsigned long long pos = __ldg(offsets + word) + byte_sum(earlier_tiles) + rc.z;
    uint32_t dot = rc.w;
    if (counts) {  // narrow rows: the scoring kernels do not carry the dot
      const uint32_t* ra = reinterpret_cast<const uint32_t*>(counts + static_cast<size_t>(s) * pitch);
      const uint32_t* rb = reinterpret_cast<const uint32_t*>(counts + static_cast<size_t>(t) * pitch);
      dot = 0;
      for (int wd = 0; wd < words; ++wd) dot = __dp4a(__ldg(ra + wd), __ldg(rb + wd), dot);
    }
    ordered[pos] = make_uint2(s | (dot << 24), t);  // s, t < 2^24 (check_params), dot <= 225
  }
}

// Up to kMaxSegs source segments (one per rank after a multi-GPU all-gather of the packed words, each padded to the
// same stride) are read as one logical array.
constexpr int kMaxSegs = 16;
struct Segs {
  int n;
  long long src[kMaxSegs];      // first word of segment r in `ordered`
  long long dst[kMaxSegs + 1];  // first output index of segment r; dst[n] = total
};

__global__ void __launch_bounds__(256) expand_kernel(const uint2* ordered, Segs segs, const int32_t* sqnorm,
                                                     int64_t* src, int64_t* dst, double* w64, float* w32,
                                                     uint32_t* dot_out) {
  const long long n_out = segs.dst[segs.n];
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n_out;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int r = 0;
    while (r + 1 < segs.n && i >= segs.dst[r + 1]) ++r;
    const uint2 e = __ldg(ordered + segs.src[r] + (i - segs.dst[r]));
    const uint32_t s = e.x & 0xFFFFFFu, dot = e.x >> 24, t = e.y;
    src[i] = s;
    dst[i] = t;
    const double p = static_cast<double>(static_cast<long long>(__ldg(sqnorm + s)) * static_cast<long long>(__ldg(sqnorm + t)));
    const double w = static_cast<double>(dot) / sqrt(p);
    if (w64) w64[i] = w;
    if (w32) w32[i] = static_cast<float>(w);
    if (dot_out) dot_out[i] = dot;
  }
}

// ---------------------------------------------------------------------------- global top-n
__device__ __forceinline__ uint32_t class_key(const int64_t* src, const int64_t* dst, const uint32_t* dot,
                                              const int32_t* sqnorm, int32_t p_max, long long i) {
  const uint32_t p = static_cast<uint32_t>(sqnorm[src[i]]) * static_cast<uint32_t>(sqnorm[dst[i]]);
  return dot[i] * static_cast<uint32_t>(p_max + 1) + p;
}

__global__ void __launch_bounds__(256) class_hist_kernel(const int64_t* src, const int64_t* dst, const uint32_t* dot,
                                                         long long n, const int32_t* sqnorm, int32_t p_max,
                                                         unsigned long long* hist) {
  for (long long base = blockIdx.x * static_cast<long long>(blockDim.x); base < n;
       base += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long i = base + threadIdx.x;
    const bool ok = i < n;
    const uint32_t key = ok ? class_key(src, dst, dot, sqnorm, p_max, i) : 0xFFFFFFFFu;
    // few distinct classes: one atomic per class per warp
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
    if (ok && (__ffs(peers) - 1) == (threadIdx.x & 31)) atomicAdd(hist + key, static_cast<unsigned long long>(__popc(peers)));
  }
}

__global__ void __launch_bounds__(256) tie_flag_kernel(const int64_t* src, const int64_t* dst, const uint32_t* dot,
                                                       long long n, const int32_t* sqnorm, int32_t p_max,
                                                       const uint8_t* action, uint8_t* act_out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    act_out[i] = action[class_key(src, dst, dot, sqnorm, p_max, i)];
}

struct IsTie {
  __host__ __device__ unsigned long long operator()(uint8_t a) const { return a == 2 ? 1ull : 0ull; }
};

__global__ void __launch_bounds__(256) keep_flag_kernel(const uint8_t* act, const unsigned long long* tie_rank,
                                                        long long n, long long tie_budget, uint8_t* keep) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    keep[i] = act[i] == 1 || (act[i] == 2 && static_cast<long long>(tie_rank[i]) < tie_budget);
}

struct IsKept {
  __host__ __device__ unsigned long long operator()(uint8_t a) const { return a ? 1ull : 0ull; }
};

__global__ void __launch_bounds__(256) select_scatter_kernel(const uint8_t* keep, const unsigned long long* pos,
                                                             long long n, const int64_t* src, const int64_t* dst,
                                                             const double* w64, int64_t* o_src, int64_t* o_dst,
                                                             double* o_w64, float* o_w32, int64_t* out_count) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (keep[i]) {
      const unsigned long long p = pos[i];
      o_src[p] = src[i];
      o_dst[p] = dst[i];
      if (o_w64) o_w64[p] = w64[i];
      if (o_w32) o_w32[p] = static_cast<float>(w64[i]);
    }
    if (i == n - 1) *out_count = static_cast<int64_t>(pos[i] + (keep[i] ? 1 : 0));
  }
}

// ---- the same on PACKED words (s | dot << 24, t): the form the placement kernels write and the exchanges move
__device__ __forceinline__ uint32_t packed_class_key(uint2 e, const int32_t* sqnorm, int32_t p_max) {
  const uint32_t s = e.x & 0xFFFFFFu, dot = e.x >> 24;
  const uint32_t p = static_cast<uint32_t>(__ldg(sqnorm + s)) * static_cast<uint32_t>(__ldg(sqnorm + e.y));
  return dot * static_cast<uint32_t>(p_max + 1) + p;
}

__global__ void __launch_bounds__(256) class_hist_packed_kernel(const uint2* ordered, long long n, const int32_t* sqnorm,
                                                                int32_t p_max, unsigned long long* hist) {
  for (long long base = blockIdx.x * static_cast<long long>(blockDim.x); base < n;
       base += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long i = base + threadIdx.x;
    const bool ok = i < n;
    const uint32_t key = ok ? packed_class_key(__ldg(ordered + i), sqnorm, p_max) : 0xFFFFFFFFu;
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
    if (ok && (__ffs(peers) - 1) == (threadIdx.x & 31)) atomicAdd(hist + key, static_cast<unsigned long long>(__popc(peers)));
  }
}

// Spread rule for the cut-off class: of its T members, numbered g = 0 .. T-1 in GLOBAL emission order, exactly B are
// kept -- those where floor((g + 1) r / 2^64) > floor(g r / 2^64) with r = ceil(B 2^64 / T).  The rule needs nothing
// but g, so any sharding of the emission order (ranks, row-block chunks) selects the same members, and every shard
// keeps its fair share of them.  Members kept among the first g: floor(g r / 2^64).
__host__ __device__ __forceinline__ unsigned long long ties_kept_before(unsigned long long g, unsigned long long r) {
#ifdef __CUDA_ARCH__
  return __umul64hi(g, r);
#else
  return static_cast<unsigned long long>((static_cast<unsigned __int128>(g) * r) >> 64);
#endif
}

// One scan serves the whole selection: its element is (1 << 32) for an edge above the cut-off class, 1 for a member of
// the class, 0 for anything to drop; the exclusive prefix of word i is then (edges above before i) << 32 | (members
// before i), and because the rule above depends on the member NUMBER only, the output slot of word i follows from
// the prefix alone: above + ties_kept_before(g0 + members) - ties_kept_before(g0).
struct ClassOf {
  const uint2* ordered;
  const int32_t* sqnorm;
  const uint8_t* action;
  int32_t p_max;
  __host__ __device__ unsigned long long operator()(long long i) const {
#ifdef __CUDA_ARCH__
    const uint8_t a = action[packed_class_key(__ldg(ordered + i), sqnorm, p_max)];
    return a == 1 ? (1ull << 32) : (a == 2 ? 1ull : 0ull);
#else
    return 0ull;
#endif
  }
};

__global__ void __launch_bounds__(256) packed_scatter_kernel(const uint2* ordered, long long n, const int32_t* sqnorm,
                                                             int32_t p_max, const uint8_t* action,
                                                             const unsigned long long* prefix, unsigned long long r,
                                                             const long long* cursors, long long out_capacity,
                                                             int64_t* o_src, int64_t* o_dst, double* o_w64, float* o_w32,
                                                             long long* totals) {
  const unsigned long long g0 = static_cast<unsigned long long>(cursors[0]);
  const long long out0 = cursors[1];
  const unsigned long long kept0 = ties_kept_before(g0, r);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint2 e = __ldg(ordered + i);
    const uint32_t s = e.x & 0xFFFFFFu, dot = e.x >> 24, t = e.y;
    const long long p = static_cast<long long>(__ldg(sqnorm + s)) * static_cast<long long>(__ldg(sqnorm + t));
    const uint8_t act = action[dot * static_cast<uint32_t>(p_max + 1) + static_cast<uint32_t>(p)];
    const unsigned long long pre = prefix[i];
    const unsigned long long above = pre >> 32, g = g0 + (pre & 0xFFFFFFFFull);
    const bool kept = act == 1 || (act == 2 && ties_kept_before(g + 1ull, r) > ties_kept_before(g, r));
    if (kept) {
      const long long o = out0 + static_cast<long long>(above + (ties_kept_before(g, r) - kept0));
      if (o < out_capacity) {
        o_src[o] = s;
        o_dst[o] = t;
        const double w = static_cast<double>(dot) / sqrt(static_cast<double>(p));
        if (o_w64) o_w64[o] = w;
        if (o_w32) o_w32[o] = static_cast<float>(w);
      }
    }
    if (i == n - 1) {  // totals of this call, applied to the cursors by the kernel behind this one
      const unsigned long long members = (pre & 0xFFFFFFFFull) + (act == 2 ? 1ull : 0ull);
      totals[0] = static_cast<long long>(members);
      totals[1] = static_cast<long long>(above + (act == 1 ? 1ull : 0ull) + (ties_kept_before(g0 + members, r) - kept0));
    }
  }
}

__global__ void advance_cursors_kernel(long long* cursors, const long long* totals) {
  cursors[0] += totals[0];
  cursors[1] += totals[1];
}

inline int grid_for(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = static_cast<long long>(gg::sm_count()) * 8;
  if (g > cap) g = cap;
  return g < 1 ? 1 : static_cast<int>(g);
}

template <typename Op, typename In>
size_t scan_temp_bytes(long long n) {
  size_t bytes = 0;
  cub::TransformInputIterator<unsigned long long, Op, const In*> it(nullptr, Op());
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, static_cast<unsigned long long*>(nullptr), n);
  return bytes;
}

int check_params(const gg_kf_params_t* p) {
  GG_REQUIRE(p != nullptr, GG_ERR_ARG, "null params");
  GG_REQUIRE(p->n >= 0 && p->n <= (1ll << 24), GG_ERR_ARG, "n outside [0, 2^24]");
  GG_REQUIRE(p->d >= 4 && p->d % 4 == 0, GG_ERR_ARG, "d must be a positive multiple of 4");
  const int n_blocks = static_cast<int>((p->n + GG_KF_BLOCK - 1) / GG_KF_BLOCK);
  GG_REQUIRE(p->iblock_begin >= 0 && p->iblock_begin <= p->iblock_end && p->iblock_end <= n_blocks, GG_ERR_ARG,
             "row-block shard outside [0, ceil(n/1024)]");
  GG_REQUIRE(p->p_max >= 0 && p->fast_shift >= 0 && p->fast_shift < 32, GG_ERR_ARG, "bad threshold tables");
  return GG_OK;
}

}  // namespace

extern "C" size_t gg_kf_rowcnt_elems(int64_t n, int iblock_begin, int iblock_end) {
  if (n <= 0 || iblock_end <= iblock_begin) return 0;
  const size_t nj = static_cast<size_t>((n + GG_KF_BLOCK - 1) / GG_KF_BLOCK);
  return static_cast<size_t>(iblock_end - iblock_begin) * nj * GG_KF_BLOCK;
}

extern "C" int gg_kf_emit(const gg_kf_params_t* p, const uint8_t* counts, const int32_t* sqnorm, const uint16_t* dmin,
                          const uint32_t* fast_scale, gg_kf_rec_t* recs, int64_t rec_capacity, uint64_t* rowcnt,
                          int64_t* kf_status, uint64_t* class_hist, gg_stream_t st) {
  if (int rc = check_params(p)) return rc;
  GG_REQUIRE(kf_status, GG_ERR_ARG, "null status");
  cudaStream_t s = gg::as_stream(st);
  GG_CUDA_CHECK(cudaMemsetAsync(kf_status, 0, 4 * sizeof(int64_t), s));
  const size_t rc_elems = gg_kf_rowcnt_elems(p->n, p->iblock_begin, p->iblock_end);
  if (rc_elems == 0) return GG_OK;
  GG_REQUIRE(counts && sqnorm && dmin && fast_scale && rowcnt, GG_ERR_ARG, "null buffer");
  GG_REQUIRE(rec_capacity >= 0 && (rec_capacity == 0 || recs), GG_ERR_ARG, "record buffer");
  GG_REQUIRE((reinterpret_cast<uintptr_t>(counts) & 15) == 0 && (reinterpret_cast<uintptr_t>(recs) & 15) == 0,
             GG_ERR_ARG, "counts / recs must be 16-byte aligned");
  GG_CUDA_CHECK(cudaMemsetAsync(rowcnt, 0, rc_elems * sizeof(uint64_t), s));

  KfArgs a;
  a.counts = counts;
  a.sqnorm = sqnorm;
  a.dmin = dmin;
  a.fast_scale = fast_scale;
  a.n = p->n;
  a.d = p->d;
  a.ib0 = p->iblock_begin;
  a.ib1 = p->iblock_end;
  a.n_jblocks = static_cast<int>((p->n + GG_KF_BLOCK - 1) / GG_KF_BLOCK);
  a.fast_shift = p->fast_shift;
  a.rowcnt = reinterpret_cast<uint8_t*>(rowcnt);
  a.out.recs = recs;
  a.out.capacity = rec_capacity;
  a.out.status = reinterpret_cast<unsigned long long*>(kf_status);
  a.class_hist = reinterpret_cast<unsigned long long*>(class_hist);
  a.hist_m = p->hist_m;
  a.p_max = p->p_max;
  a.interleave = 0;

  int impl = p->impl;
  const bool mma_ok = p->d == 32 || p->d == 64 || (p->d % 128 == 0 && p->d <= 1024);
  // >= 128-byte rows: CTA-pair kernel; 32/64-byte rows: single-CTA kernel with the narrow epilogue (1.12 ms vs
  // 1.42 ms for the dp4a kernel at k=8, D=16); anything else: SIMT kernels
  const bool dp4a_ok = p->d == 4 || p->d == 16 || p->d == 64;
  if (impl == GG_KF_IMPL_AUTO)
    impl = (p->d % 128 == 0 && p->d <= 1024) ? GG_KF_IMPL_TCGEN05_PAIR
           : (mma_ok ? GG_KF_IMPL_TCGEN05 : (dp4a_ok ? GG_KF_IMPL_DP4A : GG_KF_IMPL_NAIVE));
  const long long units = static_cast<long long>(a.ib1 - a.ib0) * a.n_jblocks * kStripsPerBlock;
  GG_REQUIRE(units < (1ll << 31), GG_ERR_ARG, "too many work units for one launch; shard the row blocks");
  GG_REQUIRE(class_hist == nullptr || impl == GG_KF_IMPL_TCGEN05_PAIR, GG_ERR_ARG,
             "the in-kernel score-class histogram exists for the CTA-pair tensor path only");
  GG_REQUIRE(p->hist_m >= 0, GG_ERR_ARG, "negative hist_m");
  switch (impl) {
    case GG_KF_IMPL_DP4A: {
      int sq_max = 0;  // p_max = sq_max^2 by construction (core.threshold_tables)
      while ((sq_max + 1) * (sq_max + 1) <= p->p_max) ++sq_max;
      GG_REQUIRE(sq_max * sq_max == p->p_max && sq_max <= 255, GG_ERR_ARG, "p_max must be a square <= 255^2");
      if (p->d == 4) return launch_dp4a<1>(a, units, sq_max, s);
      if (p->d == 16) return launch_dp4a<4>(a, units, sq_max, s);
      if (p->d == 64) return launch_dp4a<16>(a, units, sq_max, s);
      gg::set_error("dp4a KF kernel supports d in {4, 16, 64}, got %d", p->d);
      return GG_ERR_ARG;
    }
    case GG_KF_IMPL_TCGEN05:
      return ggkf::emit_tcgen05(a, p->p_max, s);
    case GG_KF_IMPL_TCGEN05_PAIR:
      return ggkf::emit_tcgen05_pair(a, s);
    case GG_KF_IMPL_NAIVE:
      GG_REQUIRE(p->d % 64 == 0, GG_ERR_ARG, "generic KF kernel needs d % 64 == 0");
      kf_generic_kernel<<<static_cast<unsigned>(units), kStrip, 0, s>>>(a);
      GG_CUDA_CHECK(cudaGetLastError());
      return GG_OK;
    default:
      gg::set_error("unknown KF impl %d", impl);
      return GG_ERR_ARG;
  }
}

extern "C" size_t gg_kf_place_workspace_bytes(int64_t n, int iblock_begin, int iblock_end) {
  const size_t elems = gg_kf_rowcnt_elems(n, iblock_begin, iblock_end);
  gg::Carver c(nullptr);
  c.take<unsigned long long>(elems ? elems : 1);
  c.take<char>(scan_temp_bytes<RowTotal, unsigned long long>(static_cast<long long>(elems ? elems : 1)));
  return c.used();
}

extern "C" int gg_kf_place(const gg_kf_params_t* p, const gg_kf_rec_t* recs, int64_t n_recs, const uint64_t* rowcnt,
                           const uint8_t* counts, uint64_t* ordered, void* workspace, size_t workspace_bytes,
                           gg_stream_t st) {
  if (int rc = check_params(p)) return rc;
  GG_REQUIRE(n_recs >= 0, GG_ERR_ARG, "negative record count");
  if (n_recs == 0) return GG_OK;
  GG_REQUIRE(recs && rowcnt && ordered && workspace, GG_ERR_ARG, "null buffer");
  GG_REQUIRE(workspace_bytes >= gg_kf_place_workspace_bytes(p->n, p->iblock_begin, p->iblock_end), GG_ERR_WORKSPACE,
             "workspace too small");
  cudaStream_t s = gg::as_stream(st);
  const long long elems = static_cast<long long>(gg_kf_rowcnt_elems(p->n, p->iblock_begin, p->iblock_end));
  gg::Carver c(workspace);
  unsigned long long* offsets = c.take<unsigned long long>(elems);
  size_t tb = scan_temp_bytes<RowTotal, unsigned long long>(elems);
  void* temp = c.take<char>(tb);
  const unsigned long long* words = reinterpret_cast<const unsigned long long*>(rowcnt);
  cub::TransformInputIterator<unsigned long long, RowTotal, const unsigned long long*> it(words, RowTotal());
  GG_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(temp, tb, it, offsets, elems, s));
  const int n_jblocks = static_cast<int>((p->n + GG_KF_BLOCK - 1) / GG_KF_BLOCK);
  // rows of <= 64 bytes: records carry no dot, recompute it here (counts must be the matrix the emit step scored)
  const bool recompute = p->d <= 64;
  GG_REQUIRE(!recompute || counts != nullptr, GG_ERR_ARG, "counts is required to finalise narrow-row records");
  place_kernel<<<grid_for(n_recs), 256, 0, s>>>(recs, n_recs, offsets, words, p->iblock_begin, n_jblocks,
                                                recompute ? counts : nullptr, p->d, reinterpret_cast<uint2*>(ordered),
                                                (p->d_valid > 0 && p->d_valid < p->d ? p->d_valid : p->d) / 4);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}

extern "C" int gg_kf_expand(const uint64_t* ordered, int n_segments, const int64_t* h_seg_first, const int64_t* h_seg_size,
                            const int32_t* sqnorm, int64_t* edge_src, int64_t* edge_dst, double* weight64,
                            float* weight32, uint32_t* dot_out, gg_stream_t st) {
  GG_REQUIRE(n_segments >= 1 && n_segments <= kMaxSegs && h_seg_first && h_seg_size, GG_ERR_ARG,
             "1..16 segments expected");
  Segs segs;
  segs.n = n_segments;
  long long total = 0;
  for (int r = 0; r < n_segments; ++r) {
    GG_REQUIRE(h_seg_size[r] >= 0 && h_seg_first[r] >= 0, GG_ERR_ARG, "negative segment");
    segs.src[r] = h_seg_first[r];
    segs.dst[r] = total;
    total += h_seg_size[r];
  }
  segs.dst[n_segments] = total;
  if (total == 0) return GG_OK;
  GG_REQUIRE(ordered && sqnorm && edge_src && edge_dst, GG_ERR_ARG, "null buffer");
  expand_kernel<<<grid_for(total), 256, 0, gg::as_stream(st)>>>(reinterpret_cast<const uint2*>(ordered), segs, sqnorm,
                                                               edge_src, edge_dst, weight64, weight32, dot_out);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}

extern "C" size_t gg_kf_finalize_workspace_bytes(int64_t n, int iblock_begin, int iblock_end, int64_t n_recs) {
  gg::Carver c(nullptr);
  c.take<char>(gg_kf_place_workspace_bytes(n, iblock_begin, iblock_end));
  c.take<uint2>(n_recs > 0 ? n_recs : 1);
  return c.used();
}

extern "C" int gg_kf_finalize(const gg_kf_params_t* p, const gg_kf_rec_t* recs, int64_t n_recs, const uint64_t* rowcnt,
                              const int32_t* sqnorm, const uint8_t* counts, int64_t* edge_src, int64_t* edge_dst,
                              double* weight64, float* weight32, uint32_t* dot_out, void* workspace,
                              size_t workspace_bytes, gg_stream_t st) {
  if (int rc = check_params(p)) return rc;
  GG_REQUIRE(n_recs >= 0, GG_ERR_ARG, "negative record count");
  if (n_recs == 0) return GG_OK;
  GG_REQUIRE(workspace && workspace_bytes >= gg_kf_finalize_workspace_bytes(p->n, p->iblock_begin, p->iblock_end, n_recs),
             GG_ERR_WORKSPACE, "workspace too small");
  gg::Carver c(workspace);
  const size_t place_bytes = gg_kf_place_workspace_bytes(p->n, p->iblock_begin, p->iblock_end);
  void* place_ws = c.take<char>(place_bytes);
  uint64_t* ordered = reinterpret_cast<uint64_t*>(c.take<uint2>(n_recs));
  if (int rc = gg_kf_place(p, recs, n_recs, rowcnt, counts, ordered, place_ws, place_bytes, st)) return rc;
  const int64_t first = 0, size = n_recs;
  return gg_kf_expand(ordered, 1, &first, &size, sqnorm, edge_src, edge_dst, weight64, weight32, dot_out, st);
}

extern "C" int gg_kf_class_hist(const int64_t* edge_src, const int64_t* edge_dst, const uint32_t* dot, int64_t n_edges,
                                const int32_t* sqnorm, int32_t p_max, unsigned long long* class_hist, gg_stream_t st) {
  GG_REQUIRE(n_edges >= 0 && p_max >= 0, GG_ERR_ARG, "bad size");
  if (n_edges == 0) return GG_OK;
  GG_REQUIRE(edge_src && edge_dst && dot && sqnorm && class_hist, GG_ERR_ARG, "null buffer");
  class_hist_kernel<<<grid_for(n_edges), 256, 0, gg::as_stream(st)>>>(edge_src, edge_dst, dot, n_edges, sqnorm, p_max,
                                                                     class_hist);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}

extern "C" int gg_kf_class_hist_packed(const uint64_t* ordered, int64_t n_edges, const int32_t* sqnorm, int32_t p_max,
                                       uint64_t* class_hist, gg_stream_t st) {
  GG_REQUIRE(n_edges >= 0 && p_max >= 0, GG_ERR_ARG, "bad size");
  if (n_edges == 0) return GG_OK;
  GG_REQUIRE(ordered && sqnorm && class_hist, GG_ERR_ARG, "null buffer");
  class_hist_packed_kernel<<<grid_for(n_edges), 256, 0, gg::as_stream(st)>>>(
      reinterpret_cast<const uint2*>(ordered), n_edges, sqnorm, p_max, reinterpret_cast<unsigned long long*>(class_hist));
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}

namespace {
size_t class_scan_temp_bytes(long long n) {
  size_t bytes = 0;
  cub::TransformInputIterator<unsigned long long, ClassOf, cub::CountingInputIterator<long long>> it(
      cub::CountingInputIterator<long long>(0), ClassOf{nullptr, nullptr, nullptr, 0});
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, static_cast<unsigned long long*>(nullptr), n);
  return bytes;
}
}  // namespace

extern "C" size_t gg_kf_select_packed_workspace_bytes(int64_t n_edges) {
  const long long n = n_edges > 0 ? n_edges : 1;
  gg::Carver c(nullptr);
  c.take<unsigned long long>(n);
  c.take<long long>(2);
  c.take<char>(class_scan_temp_bytes(n));
  return c.used();
}

extern "C" int gg_kf_select_packed(const uint64_t* ordered, int64_t n_edges, const int32_t* sqnorm, int32_t p_max,
                                   const uint8_t* class_action, uint64_t tie_ratio, int64_t* cursors,
                                   int64_t* out_src, int64_t* out_dst, double* out_w64, float* out_w32,
                                   int64_t out_capacity, void* workspace, size_t workspace_bytes, gg_stream_t st) {
  GG_REQUIRE(n_edges >= 0 && n_edges < (1ll << 31) && p_max >= 0 && out_capacity >= 0, GG_ERR_ARG,
             "bad size (at most 2^31 - 1 packed words per call)");
  if (n_edges == 0) return GG_OK;
  GG_REQUIRE(ordered && sqnorm && class_action && cursors && out_src && out_dst && workspace, GG_ERR_ARG, "null buffer");
  GG_REQUIRE(workspace_bytes >= gg_kf_select_packed_workspace_bytes(n_edges), GG_ERR_WORKSPACE, "workspace too small");
  cudaStream_t s = gg::as_stream(st);
  gg::Carver c(workspace);
  unsigned long long* prefix = c.take<unsigned long long>(n_edges);
  long long* totals = c.take<long long>(2);
  size_t temp_bytes = class_scan_temp_bytes(n_edges);
  void* temp = c.take<char>(temp_bytes);
  const uint2* words = reinterpret_cast<const uint2*>(ordered);
  {
    cub::TransformInputIterator<unsigned long long, ClassOf, cub::CountingInputIterator<long long>> it(
        cub::CountingInputIterator<long long>(0), ClassOf{words, sqnorm, class_action, p_max});
    GG_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, it, prefix, n_edges, s));
  }
  packed_scatter_kernel<<<grid_for(n_edges), 256, 0, s>>>(words, n_edges, sqnorm, p_max, class_action, prefix,
                                                          static_cast<unsigned long long>(tie_ratio),
                                                          reinterpret_cast<const long long*>(cursors), out_capacity,
                                                          out_src, out_dst, out_w64, out_w32, totals);
  GG_CUDA_CHECK(cudaGetLastError());
  advance_cursors_kernel<<<1, 1, 0, s>>>(reinterpret_cast<long long*>(cursors), totals);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}

extern "C" size_t gg_kf_select_workspace_bytes(int64_t n_edges) {
  const long long n = n_edges > 0 ? n_edges : 1;
  gg::Carver c(nullptr);
  c.take<uint8_t>(n);
  c.take<uint8_t>(n);
  c.take<unsigned long long>(n);
  c.take<unsigned long long>(n);
  c.take<char>(scan_temp_bytes<IsTie, uint8_t>(n));
  return c.used();
}

extern "C" int gg_kf_select(const int64_t* edge_src, const int64_t* edge_dst, const double* weight64,
                            const uint32_t* dot, int64_t n_edges, const int32_t* sqnorm, int32_t p_max,
                            const uint8_t* class_action, int64_t tie_budget, int64_t* out_src, int64_t* out_dst,
                            double* out_w64, float* out_w32, int64_t* out_count, void* workspace,
                            size_t workspace_bytes, gg_stream_t st) {
  GG_REQUIRE(n_edges >= 0 && p_max >= 0 && out_count, GG_ERR_ARG, "bad size");
  cudaStream_t s = gg::as_stream(st);
  GG_CUDA_CHECK(cudaMemsetAsync(out_count, 0, sizeof(int64_t), s));
  if (n_edges == 0) return GG_OK;
  GG_REQUIRE(edge_src && edge_dst && weight64 && dot && sqnorm && class_action && out_src && out_dst && workspace,
             GG_ERR_ARG, "null buffer");
  GG_REQUIRE(workspace_bytes >= gg_kf_select_workspace_bytes(n_edges), GG_ERR_WORKSPACE, "workspace too small");
  gg::Carver c(workspace);
  uint8_t* act = c.take<uint8_t>(n_edges);
  uint8_t* keep = c.take<uint8_t>(n_edges);
  unsigned long long* tie_rank = c.take<unsigned long long>(n_edges);
  unsigned long long* pos = c.take<unsigned long long>(n_edges);
  size_t tb = scan_temp_bytes<IsTie, uint8_t>(n_edges);
  void* temp = c.take<char>(tb);
  const int g = grid_for(n_edges);
  tie_flag_kernel<<<g, 256, 0, s>>>(edge_src, edge_dst, dot, n_edges, sqnorm, p_max, class_action, act);
  {
    cub::TransformInputIterator<unsigned long long, IsTie, const uint8_t*> it(act, IsTie());
    size_t b = tb;
    GG_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(temp, b, it, tie_rank, n_edges, s));
  }
  keep_flag_kernel<<<g, 256, 0, s>>>(act, tie_rank, n_edges, tie_budget, keep);
  {
    cub::TransformInputIterator<unsigned long long, IsKept, const uint8_t*> it(keep, IsKept());
    size_t b = tb;
    GG_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(temp, b, it, pos, n_edges, s));
  }
  select_scatter_kernel<<<g, 256, 0, s>>>(keep, pos, n_edges, edge_src, edge_dst, weight64, out_src, out_dst, out_w64,
                                          out_w32, out_count);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}
