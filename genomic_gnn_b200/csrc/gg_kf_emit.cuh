// Candidate emission shared by every KF scoring kernel (SIMT dp4a, SIMT
// generic, tcgen05).  One thread owns one row s of a 128-row strip and walks
// the 128 columns of one tile in increasing t, so the rank of a candidate
// inside (row, column tile) is a running per-thread counter; that rank plus a
// prefix sum over the per-(block pair, row, tile) counts reproduces the
// reference's emission order (gnn_utils.py:262-301) without sorting, and lets
// any warp score any tile independently.
#pragma once
#include "gg_common.cuh"

namespace ggkf {

constexpr int kBlock = GG_KF_BLOCK;  // rows / columns of a reference block
constexpr int kStrip = 128;          // rows per CTA (one per thread)
constexpr int kStripsPerBlock = kBlock / kStrip;
constexpr int kTile = 128;           // columns per rank-counting tile
constexpr int kTilesPerBlock = kBlock / kTile;
constexpr int kStageCap = 64;        // default: records staged per warp before a coalesced flush

struct EmitOut {
  gg_kf_rec_t* recs;
  long long capacity;
  unsigned long long* status;  // [0] candidates (also the global write cursor, except for block_commit: [1])
};

constexpr uint32_t kNoRec = 0xFFFFFFFFu;  // gg_kf_rec_t::s of a padding record (block_commit); placement skips it

struct KfArgs {
  const uint8_t* counts;
  const int32_t* sqnorm;
  const uint16_t* dmin;
  const uint32_t* fast_scale;
  long long n;
  int d;
  int ib0, ib1;
  int n_jblocks;
  int fast_shift;
  uint8_t* rowcnt;  // one 64-bit word per (block pair, row): byte t = candidates of that row in column tile t
  EmitOut out;
  // Optional exact score-class histogram of the candidates (global top-n BEFORE emission, gnn_utils.py:311-317):
  // class_hist[dot * (p_max + 1) + sq_s * sq_t] += 1 per candidate.  hist_m > 0 promises that every row sums to
  // hist_m (rows made by gg_kf_counts: m = k - small_k + 1), which lets the tensor kernel keep a compact
  // shared-memory histogram keyed by ((sq - m) / 2, dot); 0 = no such promise.
  unsigned long long* class_hist;
  int hist_m;
  int p_max;
  int interleave;  // CTA-pair kernel: strip-pair rows dealt round-robin (serpentine) to the CTA pairs, see pair_walk_begin
};

// CAP trades shared memory for global-cursor traffic: every flush is one returning atomic on a single address,
// whose round trip (~1 us under load) stalls the warp, so kernels that emit many records use a deep stage.
template <int CAP>
struct WarpStageT {
  static constexpr int kCap = CAP;
  uint4 rec[CAP];
  uint32_t dotw[8][32];  // [column bit / 4][lane]: dots of the current 32-column chunk, one byte each (dot <= 225)
  int fill;
  int pad[3];
};
using WarpStage = WarpStageT<kStageCap>;

template <class WS>
__device__ __forceinline__ void stage_set_dot(WS& ws, int b, int lane, uint32_t dot) {
  reinterpret_cast<uint8_t*>(&ws.dotw[b >> 2][lane])[b & 3] = static_cast<uint8_t>(dot);
}
template <class WS>
__device__ __forceinline__ uint32_t stage_get_dot(const WS& ws, int b, int lane) {
  return (ws.dotw[b >> 2][lane] >> (8 * (b & 3))) & 0xFFu;
}

template <class WS>
__device__ __forceinline__ void stage_flush(WS& ws, const EmitOut& out, int lane) {
  const int fill = ws.fill;  // warp-uniform
  if (fill == 0) return;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(out.status, static_cast<unsigned long long>(fill));
  base = __shfl_sync(0xFFFFFFFFu, base, 0);
  uint4* dst = reinterpret_cast<uint4*>(out.recs);
  for (int i = lane; i < fill; i += 32)
    if (static_cast<long long>(base + i) < out.capacity) dst[base + i] = ws.rec[i];
  __syncwarp();
  if (lane == 0) ws.fill = 0;
  __syncwarp();
}

// Every lane has marked its exact candidates of the current chunk in `exact`
// (bit b = column t0 + b) and left their dots in ws.dotw (stage_set_dot); turn them into
// records.  Must be called by the whole warp.
template <class WS>
__device__ __forceinline__ void stage_commit(WS& ws, const EmitOut& out, int lane, uint32_t exact, uint32_t s,
                                             uint32_t t0, uint32_t& tile_cnt) {
  const int n_l = __popc(exact);
  int incl = n_l;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += v;
  }
  const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
  if (total == 0) return;
  const int off = incl - n_l;
  if (total > WS::kCap) {  // dense chunk: bypass the stage
    stage_flush(ws, out, lane);
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(out.status, static_cast<unsigned long long>(total));
    base = __shfl_sync(0xFFFFFFFFu, base, 0) + off;
    uint4* dst = reinterpret_cast<uint4*>(out.recs);
    int i = 0;
    for (uint32_t m = exact; m; m &= m - 1, ++i) {
      const int b = __ffs(m) - 1;
      if (static_cast<long long>(base + i) < out.capacity)
        dst[base + i] = make_uint4(s, t0 + b, tile_cnt + i, stage_get_dot(ws, b, lane));
    }
  } else {
    if (ws.fill + total > WS::kCap) stage_flush(ws, out, lane);
    const int fill = ws.fill;
    int i = 0;
    for (uint32_t m = exact; m; m &= m - 1, ++i) {
      const int b = __ffs(m) - 1;
      ws.rec[fill + off + i] = make_uint4(s, t0 + b, tile_cnt + i, stage_get_dot(ws, b, lane));
    }
    __syncwarp();
    if (lane == 0) ws.fill = fill + total;
    __syncwarp();
  }
  tile_cnt += n_l;
}

// Whole-tile variant for narrow feature rows: every lane holds the exact candidate masks of the four 32-column
// chunks of one 128-column tile (m[c] bit b = column t0 + 32 c + b).  One warp scan per tile instead of per chunk;
// the records carry no dot (the finalize step recomputes it from the two short rows).
template <class WS>
__device__ __forceinline__ void stage_commit_tile(WS& ws, const EmitOut& out, int lane, const uint32_t (&m)[4],
                                                  uint32_t s, uint32_t t0, uint32_t& tile_cnt) {
  const int n_l = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
  int incl = n_l;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += v;
  }
  const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
  if (total == 0) return;
  const int off = incl - n_l;
  const bool direct = total > WS::kCap;
  if (direct || ws.fill + total > WS::kCap) stage_flush(ws, out, lane);
  unsigned long long base = 0;
  if (direct) {
    if (lane == 0) base = atomicAdd(out.status, static_cast<unsigned long long>(total));
    base = __shfl_sync(0xFFFFFFFFu, base, 0) + off;
  }
  const int fill = ws.fill;
  uint4* dst = reinterpret_cast<uint4*>(out.recs);
  int i = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    for (uint32_t mm = m[c]; mm; mm &= mm - 1, ++i) {
      const uint4 rec = make_uint4(s, t0 + 32 * c + (__ffs(mm) - 1), tile_cnt + i, 0u);
      if (direct) {
        if (static_cast<long long>(base + i) < out.capacity) dst[base + i] = rec;
      } else {
        ws.rec[fill + off + i] = rec;
      }
    }
  }
  if (!direct) {
    __syncwarp();
    if (lane == 0) ws.fill = fill + total;
    __syncwarp();
  }
  tile_cnt += n_l;
}

// ---- block commit (CTA-pair tensor kernel): records go straight to global memory, into slot blocks a warp reserves
// from the cursor status[1] -- 32 slots at first, doubling to 1024 -- so a dense shard (top-n re-emission at threshold
// 0: ~35 candidates per warp and 32-column chunk) costs one returning atomic per ~1000 records instead of one per
// chunk on a single address.  Unused slots of a block are padded with kNoRec records; the exact candidate count is
// kept per thread and added to status[0] once per warp.  After a reservation lands beyond the capacity nothing more
// is reserved or written (the caller scores again with the exact size, or -- global top-n -- never needed the records).
struct BlockCursor {
  unsigned long long base;
  int used, size;
  bool dead;
  unsigned long long found;  // per thread
};

__device__ __forceinline__ void block_pad(const BlockCursor& bc, const EmitOut& out, int lane) {
  uint4* dst = reinterpret_cast<uint4*>(out.recs);
  for (int i = bc.used + lane; i < bc.size; i += 32)
    if (static_cast<long long>(bc.base + i) < out.capacity) dst[bc.base + i] = make_uint4(kNoRec, 0u, 0u, 0u);
}

template <class WS>
__device__ __forceinline__ void block_commit(BlockCursor& bc, const WS& ws, const EmitOut& out, int lane, uint32_t exact,
                                             uint32_t s, uint32_t t0, uint32_t& tile_cnt) {
  const int n_l = __popc(exact);
  bc.found += n_l;
  if (bc.dead) {
    tile_cnt += n_l;
    return;
  }
  int incl = n_l;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += v;
  }
  const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
  if (total == 0) return;
  if (bc.used + total > bc.size) {
    block_pad(bc, out, lane);
    int next = bc.size * 2 < 32 ? 32 : (bc.size * 2 > 1024 ? 1024 : bc.size * 2);
    if (next < total) next = (total + 31) & ~31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(out.status + 1, static_cast<unsigned long long>(next));
    bc.base = __shfl_sync(0xFFFFFFFFu, base, 0);
    bc.used = 0;
    bc.size = next;
    if (static_cast<long long>(bc.base) >= out.capacity) {  // uniform: the record buffer is full, count only from now on
      bc.dead = true;
      bc.size = 0;
      tile_cnt += n_l;
      return;
    }
  }
  uint4* dst = reinterpret_cast<uint4*>(out.recs);
  const unsigned long long pos = bc.base + bc.used + (incl - n_l);
  int i = 0;
  for (uint32_t m = exact; m; m &= m - 1, ++i) {
    const int b = __ffs(m) - 1;
    if (static_cast<long long>(pos + i) < out.capacity)
      dst[pos + i] = make_uint4(s, t0 + b, tile_cnt + i, stage_get_dot(ws, b, lane));
  }
  bc.used += total;
  tile_cnt += n_l;
}

__device__ __forceinline__ void block_finish(BlockCursor& bc, const EmitOut& out, int lane) {
  if (!bc.dead) block_pad(bc, out, lane);
  unsigned long long f = bc.found;
#pragma unroll
  for (int o = 16; o; o >>= 1) f += __shfl_xor_sync(0xFFFFFFFFu, f, o);
  if (lane == 0 && f) atomicAdd(out.status, f);
}

// valid-column bits of the 32-column chunk starting at block-local column c0
__device__ __forceinline__ uint32_t chunk_col_mask(int cols_in_block, int c0) {
  const int left = cols_in_block - c0;
  return left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << left) - 1u));
}

// strict upper triangle inside a diagonal block: keep block-local columns > r
__device__ __forceinline__ uint32_t chunk_diag_mask(int r, int c0) {
  const int first = r + 1 - c0;  // first kept bit
  return first <= 0 ? 0xFFFFFFFFu : (first >= 32 ? 0u : ~((1u << first) - 1u));
}

// Counts are laid out so that a plain prefix sum walks (I, J, row, tile) = the reference's emission order.
// A (row, tile) count is <= 128, so the 8 tiles of a row pack into the bytes of one 64-bit word.
static_assert(kTilesPerBlock == 8 && kTile <= 255, "tile counts are packed as 8 bytes per row");
__host__ __device__ __forceinline__ long long rowcnt_word(int i_local, int n_jblocks, int j, int r) {
  return (static_cast<long long>(i_local) * n_jblocks + j) * kBlock + r;
}
__device__ __forceinline__ long long rowcnt_index(int i_local, int n_jblocks, int j, int r, int tile) {
  return rowcnt_word(i_local, n_jblocks, j, r) * kTilesPerBlock + tile;
}
__host__ __device__ __forceinline__ unsigned long long byte_sum(unsigned long long x) {
  const unsigned long long pair = (x & 0x00FF00FF00FF00FFull) + ((x >> 8) & 0x00FF00FF00FF00FFull);
  return (pair * 0x0001000100010001ull) >> 48;
}

}  // namespace ggkf
