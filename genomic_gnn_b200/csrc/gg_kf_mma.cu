// K4 (tensor path): KF Gram tiles on the 5th-generation tensor cores with the
// exact-threshold epilogue fused behind them, so the N x N score matrix never
// leaves the SM.
//
// Replaces the blocked X_i @ X_j.T of cosine_similarity_to_edge_index_weight
// (reference gnn_utils.py:262-279) for feature widths D >= 128 (small_k >= 4).
//
//   * operands are the uint8 sub-k-mer count rows as they sit in HBM: A = B =
//     counts[N, D] (K-major both), fetched by TMA (one tensor map, 128 rows x
//     128 bytes boxes, 128-byte swizzle), multiplied with
//     tcgen05.mma.kind::i8 into int32 accumulators in TMEM -- exact integers,
//     no conversion pass, no rounding;
//   * a CTA is persistent (one per SM) and walks (128-row strip, column block J)
//     units of the upper triangle; the strip's A rows (<= 128 KB) stay resident
//     in shared memory across all its column tiles, B tiles stream through a
//     multi-stage mbarrier ring;
//   * warp roles: warps 0-11 are three epilogue groups of 128 threads (TMEM ->
//     registers -> integer threshold test -> staged candidate records) that take
//     the accumulator tiles round-robin -- one group alone is issue-latency
//     bound at ~2.6x the MMA time of a tile --, warp 12 is the TMA producer,
//     warp 13 the MMA issuer and TMEM owner; four 128-column accumulators keep
//     the tensor pipe ahead of the epilogues.
#include <cuda.h>

#include "gg_kf_emit.cuh"
#include "gg_knn_tc.cuh"

namespace ggkf {
namespace {

constexpr int kTileN = 128;           // columns per accumulator tile
constexpr int kTileM = kStrip;        // 128 rows
constexpr int kUmmaK = 32;            // K per tcgen05.mma.kind::i8
constexpr int kMaxKBlock = 128;       // widest TMA box / smem stage row: one 128-byte swizzle row
constexpr int kAccStages = 4;         // 4 x 128 TMEM columns
constexpr int kTmemCols = 512;
constexpr int kMaxBStages = 8;
constexpr int kEpiGroups = 3;
constexpr int kGroupThreads = 128;
constexpr int kEpiThreads = kEpiGroups * kGroupThreads;
constexpr int kProducerWarp = kEpiThreads / 32, kMmaWarp = kProducerWarp + 1;
constexpr int kThreads = kEpiThreads + 64;
static_assert(kTileN == kTile, "rank tiles and accumulator tiles must coincide");
constexpr int kNarrowStageCap = 512;  // narrow rows emit ~40 records per warp-tile and have shared memory to spare
constexpr int kMaxD = 1024;           // A strip must fit shared memory next to the B ring

struct __align__(8) Barriers {
  unsigned long long a_full, a_empty;
  unsigned long long b_full[kMaxBStages], b_empty[kMaxBStages];
  unsigned long long acc_full[kAccStages], acc_empty[kAccStages];
  uint32_t tmem_base;
  int error;
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done;
}
// bounded spin: a protocol bug must surface as an error, never as a hung GPU
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity, int* error_flag) {
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 28); ++spin)
    if (mbar_try_wait(bar, parity)) return;
  *error_flag = 1;
  __trap();
}
// Whole-warp wait whose exit is a warp vote: control flow after it is provably
// uniform, so loop state, descriptors and addresses stay in uniform registers.
__device__ __forceinline__ void mbar_wait_warp(unsigned long long* bar, uint32_t parity, int* error_flag) {
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 28); ++spin)
    if (__all_sync(0xFFFFFFFFu, mbar_try_wait(bar, parity))) return;
  *error_flag = 1;
  __trap();
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, unsigned long long* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Geometry of one K block.  Row pitch in shared memory = kb_bytes (32, 64 or 128) with the matching TMA / UMMA
// swizzle mode, so that a feature width of 16..64 bytes is a single block and wider rows are split into
// 128-byte blocks.  K-major operand descriptor: 8-row groups are 8*kb_bytes apart (SBO), the leading-dimension
// field is the canonical 1 for swizzled K-major layouts, descriptor version 1 (Blackwell).
struct Geom {
  int kb_bytes;      // bytes of K per TMA box row / per stage row
  int n_kblocks;     // row pitch of `counts` / kb_bytes
  int stage_bytes;   // 128 rows * kb_bytes
  int mma_per_kb;    // kb_bytes / 32
  uint32_t desc_hi;  // SBO, version, layout type
};
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) {
  return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16);
}
// kind::i8, uint8 x uint8 -> s32, both operands K-major, M = 128, N = 128
constexpr uint32_t kIdesc = (2u << 4) | (0u << 7) | (0u << 10) | ((kTileN >> 3) << 17) | ((kTileM >> 4) << 24);

// ---------------------------------------------------------------- work walk
// Units (strip, J >= I) of this shard flattened strip-major; every CTA takes a
// contiguous range so that its A strip changes only a handful of times.
struct Walk {
  int strip, i_block, j_block;  // global strip index, its row block, column block
  long long left;               // units still to do
};

__device__ __forceinline__ long long units_before_block(int i_local, int ib0, int n_j) {
  // sum over row blocks ib0 .. ib0+i_local-1 of 8 * (n_j - I)
  const long long a = i_local;
  return kStripsPerBlock * (a * (n_j - ib0) - a * (a - 1) / 2);
}

__device__ Walk walk_begin(const KfArgs& a, long long first_unit, long long n_units) {
  Walk w;
  int il = 0;
  const int n_il = a.ib1 - a.ib0;
  while (il + 1 < n_il && units_before_block(il + 1, a.ib0, a.n_jblocks) <= first_unit) ++il;
  long long rem = first_unit - units_before_block(il, a.ib0, a.n_jblocks);
  w.i_block = a.ib0 + il;
  const int per_strip = a.n_jblocks - w.i_block;
  w.strip = w.i_block * kStripsPerBlock + static_cast<int>(rem / per_strip);
  w.j_block = w.i_block + static_cast<int>(rem % per_strip);
  w.left = n_units;
  return w;
}

__device__ __forceinline__ void walk_next(const KfArgs& a, Walk& w) {
  --w.left;
  if (++w.j_block == a.n_jblocks) {
    ++w.strip;
    w.i_block = w.strip / kStripsPerBlock;
    w.j_block = w.i_block;
  }
}

// first column tile of block J that can hold a pair s < t for this strip
__device__ __forceinline__ int first_tile(const Walk& w) {
  return w.i_block == w.j_block ? (w.strip % kStripsPerBlock) : 0;
}
// a strip / column tile that lies entirely beyond the last node has nothing to score
__device__ __forceinline__ int last_tile(const KfArgs& a, const Walk& w) {
  const long long cols = a.n - static_cast<long long>(w.j_block) * kBlock;
  const long long t = (cols + kTileN - 1) / kTileN;
  return static_cast<int>(t < kTilesPerBlock ? t : kTilesPerBlock);
}

// ---------------------------------------------------------------- kernel
template <int KB_BYTES>
__global__ void __launch_bounds__(kThreads, 1)
kf_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap, KfArgs a, int n_bstages, long long total_units,
                  int* error_flag, Geom g, int sq_max) {
  // Narrow rows (one K block): a tile is a single MMA, the epilogue is everything.  It then tests dot >= per-row
  // threshold table (exact, one LDS + one compare per score), keeps only bit masks, and commits once per tile.
  constexpr bool kNarrow = KB_BYTES < kMaxKBlock;
  using Stage = WarpStageT<(KB_BYTES < kMaxKBlock) ? kNarrowStageCap : kStageCap>;
  extern __shared__ unsigned char smem_raw[];
  // swizzled TMA boxes and UMMA descriptors assume (up to) 1024-byte aligned tiles
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int n_kblocks = g.n_kblocks;
  constexpr int kKBlock = KB_BYTES;
  constexpr int kStageBytes = kTileN * KB_BYTES;
  unsigned char* smem_a = smem;
  unsigned char* smem_b = smem_a + static_cast<size_t>(n_kblocks) * kStageBytes;
  uint16_t* sqn = reinterpret_cast<uint16_t*>(smem_b + static_cast<size_t>(n_bstages) * kStageBytes);
  Stage* stages = reinterpret_cast<Stage*>(sqn + kEpiGroups * kBlock);  // sqn: one panel per epilogue group
  Barriers* bars = reinterpret_cast<Barriers*>(stages + kEpiThreads / 32);
  uint8_t* thr_tabs = reinterpret_cast<uint8_t*>(bars + 1);  // narrow only: [group][sq_max + 1][128]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&bars->a_full, 1);
    mbar_init(&bars->a_empty, 1);
    for (int i = 0; i < n_bstages; ++i) {
      mbar_init(&bars->b_full[i], 1);
      mbar_init(&bars->b_empty[i], 1);
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&bars->acc_full[i], 1);
      mbar_init(&bars->acc_empty[i], kGroupThreads / 32);
    }
    bars->error = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // This CTA owns all 512 columns of its SM's tensor memory, so the allocation starts at lane 0 / column 0;
  // using the literal keeps every TMEM address in uniform registers.
  if (bars->tmem_base != 0) {
    if (tid == 0) *error_flag = 2;
    __trap();
  }
  constexpr uint32_t tmem_base = 0;

  // contiguous unit range of this CTA
  const long long base_units = total_units / gridDim.x, extra = total_units % gridDim.x;
  const long long u0 = base_units * blockIdx.x + (blockIdx.x < extra ? blockIdx.x : extra);
  const long long n_units = base_units + (blockIdx.x < extra ? 1 : 0);

  // The producer and MMA warps run their loops warp-uniformly (all 32 lanes wait and
  // count; one lane issues): addresses, descriptors and phases then live in
  // uniform registers.  A lane-0-only loop costs an ELECT + R2UR.BROADCAST per
  // operand and made the single issuing thread the bottleneck (profiles/r1).
  if (warp == kProducerWarp) {
    // ===================== TMA producer =====================
    if (n_units > 0) {
      if (elect_one()) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
      Walk w = walk_begin(a, u0, n_units);
      int cur_strip = -1;
      uint32_t a_round = 0, stage = 0, round = 0;
      for (; w.left > 0; walk_next(a, w)) {
        if (w.strip != cur_strip) {
          if (a_round > 0) mbar_wait_warp(&bars->a_empty, (a_round - 1) & 1, &bars->error);  // MMAs on the old strip done
          if (elect_one()) {
            mbar_expect_tx(&bars->a_full, static_cast<uint32_t>(n_kblocks) * kStageBytes);
            for (int kb = 0; kb < n_kblocks; ++kb)
              tma_load_2d(smem_a + static_cast<size_t>(kb) * kStageBytes, &tmap, &bars->a_full, kb * kKBlock,
                          w.strip * kTileM);
          }
          __syncwarp();
          cur_strip = w.strip;
          ++a_round;
        }
        const int t_end = last_tile(a, w);
        for (int ct = first_tile(w); ct < t_end; ++ct) {
          const int col = w.j_block * kBlock + ct * kTileN;
          for (int kb = 0; kb < n_kblocks; ++kb) {
            if (round > 0) mbar_wait_warp(&bars->b_empty[stage], (round - 1) & 1, &bars->error);
            if (elect_one()) {
              mbar_expect_tx(&bars->b_full[stage], kStageBytes);
              tma_load_2d(smem_b + static_cast<size_t>(stage) * kStageBytes, &tmap, &bars->b_full[stage],
                          kb * kKBlock, col);
            }
            __syncwarp();
            if (++stage == static_cast<uint32_t>(n_bstages)) {
              stage = 0;
              ++round;
            }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    if (n_units > 0) {
      Walk w = walk_begin(a, u0, n_units);
      int cur_strip = -1;
      uint32_t a_round = 0, stage = 0, round = 0, acc = 0, acc_round = 0;
      const uint32_t desc_hi = g.desc_hi;
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem_a));
      const uint32_t b_lo0 = umma_desc_lo(smem_u32(smem_b));
      constexpr int mma_per_kb = KB_BYTES / kUmmaK;
      for (; w.left > 0; walk_next(a, w)) {
        if (w.strip != cur_strip) {
          if (cur_strip >= 0 && elect_one()) tc_commit(&bars->a_empty);  // fires when every MMA issued so far retired
          __syncwarp();
          mbar_wait_warp(&bars->a_full, a_round & 1, &bars->error);
          tc_fence_after();
          cur_strip = w.strip;
          ++a_round;
        }
        const int t_end = last_tile(a, w);
        for (int ct = first_tile(w); ct < t_end; ++ct) {
          if (acc_round > 0) mbar_wait_warp(&bars->acc_empty[acc], (acc_round - 1) & 1, &bars->error);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * kTileN;
          for (int kb = 0; kb < n_kblocks; ++kb) {
            mbar_wait_warp(&bars->b_full[stage], round & 1, &bars->error);
            tc_fence_after();
            // descriptor low words: start address in 16-byte units (the tiles are 1024-byte aligned, < 256 KB)
            const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(kb) * (kStageBytes >> 4);
            const uint32_t b_lo = b_lo0 + stage * (kStageBytes >> 4);
            if (elect_one()) {
#pragma unroll
              for (int k4 = 0; k4 < mma_per_kb; ++k4) {
                const uint64_t da = (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + k4 * (kUmmaK >> 4));
                const uint64_t db = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + k4 * (kUmmaK >> 4));
                tc_mma_i8(tmem_d, da, db, kIdesc, (kb | k4) != 0);
              }
              tc_commit(&bars->b_empty[stage]);  // smem slot reusable once these MMAs have read it
              if (kb == n_kblocks - 1) tc_commit(&bars->acc_full[acc]);
            }
            __syncwarp();
            if (++stage == static_cast<uint32_t>(n_bstages)) {
              stage = 0;
              ++round;
            }
          }
          if (++acc == kAccStages) {
            acc = 0;
            ++acc_round;
          }
        }
      }
    }
  } else {
    // ===================== epilogue groups (warp % 4 selects the TMEM lane quarter) =====================
    Stage& ws = stages[warp];
    if (lane == 0) ws.fill = 0;
    __syncwarp();
    const int group = warp / 4;
    const int gt = tid % kGroupThreads;  // row of the strip == TMEM lane
    uint16_t* my_sqn = sqn + group * kBlock;
    uint8_t* my_thr_w = thr_tabs + static_cast<size_t>(group) * (sq_max + 1) * kGroupThreads + gt;
    const uint8_t* my_thr = my_thr_w;
    int tab_strip = -1;
    if (n_units > 0) {
      Walk w = walk_begin(a, u0, n_units);
      uint32_t tile_seq = 0;  // position in this CTA's tile sequence (same enumeration as the MMA warp)
      for (; w.left > 0; walk_next(a, w)) {
        const long long col0 = static_cast<long long>(w.j_block) * kBlock;
        const int cols_in_block = static_cast<int>(min(static_cast<long long>(kBlock), a.n - col0));
        const int r = (w.strip % kStripsPerBlock) * kStrip + gt;  // block-local row
        const long long s = static_cast<long long>(w.strip) * kStrip + gt;
        const bool row_ok = s < a.n;
        const bool diag = w.i_block == w.j_block;
        uint32_t sq_i = 0, scale_i = 0;
        bool panel_loaded = false;
        const int t_end = last_tile(a, w);
        for (int ct = first_tile(w); ct < t_end; ++ct, ++tile_seq) {
          if (tile_seq % kEpiGroups != static_cast<uint32_t>(group)) continue;
          if (!panel_loaded) {  // first tile this group takes in the unit: fetch the unit's column norms
            asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kGroupThreads) : "memory");
            for (int c = gt; c < kBlock; c += kGroupThreads) {
              const uint32_t sqc = c < cols_in_block ? static_cast<uint32_t>(__ldg(a.sqnorm + col0 + c)) : 0u;
              my_sqn[c] = static_cast<uint16_t>(kNarrow ? sqc * kGroupThreads : sqc);  // narrow: offset into thr table
            }
            asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kGroupThreads) : "memory");
            if (row_ok) {
              sq_i = static_cast<uint32_t>(__ldg(a.sqnorm + s));
              scale_i = __ldg(a.fast_scale + sq_i);
            }
            if (kNarrow && w.strip != tab_strip) {  // this thread's column of the table: no one else reads it
              for (int q = 0; q <= sq_max; ++q) {
                const uint32_t need = row_ok ? __ldg(a.dmin + sq_i * q) : 255u;
                my_thr_w[q * kGroupThreads] = static_cast<uint8_t>(need > 255u ? 255u : need);
              }
              tab_strip = w.strip;
            }
            panel_loaded = true;
          }
          const uint32_t acc = tile_seq % kAccStages;
          mbar_wait(&bars->acc_full[acc], (tile_seq / kAccStages) & 1, &bars->error);
          tc_fence_after();
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + acc * kTileN;
          uint32_t tile_cnt = 0;
          if constexpr (kNarrow) {
            uint32_t m[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int c0 = ct * kTileN + 32 * q;
              uint32_t v[32];
              tc_ld32(taddr + 32 * q, v);
              uint32_t mask = 0;
#pragma unroll
              for (int b = 0; b < 32; ++b) mask |= static_cast<uint32_t>(v[b] >= my_thr[my_sqn[c0 + b]]) << b;
              mask &= chunk_col_mask(cols_in_block, c0);
              if (diag) mask &= chunk_diag_mask(r, c0);
              m[q] = mask;
            }
            // the accumulator is free as soon as the masks are in registers
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->acc_empty[acc]);
            stage_commit_tile(ws, a.out, lane, m, static_cast<uint32_t>(s),
                              static_cast<uint32_t>(col0 + ct * kTileN), tile_cnt);
            if (row_ok) a.rowcnt[rowcnt_index(w.i_block - a.ib0, a.n_jblocks, w.j_block, r, ct)] =
                static_cast<uint8_t>(tile_cnt);
            continue;
          }
#pragma unroll 1
          for (int cc = 0; cc < kTileN; cc += 32) {
            const int c0 = ct * kTileN + cc;  // block-local column of bit 0
            uint32_t v[32];
            tc_ld32(taddr + cc, v);
            uint32_t mask = 0;
#pragma unroll
            for (int b = 0; b < 32; ++b)
              mask |= static_cast<uint32_t>(((v[b] * v[b]) << a.fast_shift) > scale_i * my_sqn[c0 + b]) << b;
            mask &= chunk_col_mask(cols_in_block, c0);
            if (diag) mask &= chunk_diag_mask(r, c0);
            if (!row_ok) mask = 0;
            if (__any_sync(0xFFFFFFFFu, mask != 0)) {
              // rare path: park the chunk's dots (one byte each) so that the survivors can be indexed dynamically
#pragma unroll
              for (int j = 0; j < 8; ++j)
                ws.dotw[j][lane] = v[4 * j] | (v[4 * j + 1] << 8) | (v[4 * j + 2] << 16) | (v[4 * j + 3] << 24);
              uint32_t exact = 0;
              for (uint32_t m2 = mask; m2; m2 &= m2 - 1) {
                const int b = __ffs(m2) - 1;
                if (stage_get_dot(ws, b, lane) >= __ldg(a.dmin + sq_i * my_sqn[c0 + b])) exact |= 1u << b;
              }
              stage_commit(ws, a.out, lane, exact, static_cast<uint32_t>(s), static_cast<uint32_t>(col0 + c0),
                           tile_cnt);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->acc_empty[acc]);
          if (row_ok) a.rowcnt[rowcnt_index(w.i_block - a.ib0, a.n_jblocks, w.j_block, r, ct)] = static_cast<uint8_t>(tile_cnt);
        }
      }
      stage_flush(ws, a.out, lane);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
  if (tid == 0 && bars->error) *error_flag = 1;
}


// ======================================================================================
// CTA-pair variant (cta_group::2): two CTAs of a cluster score a 256-row x 256-column tile
// together.  Each CTA keeps its own 128-row strip of A resident and streams only HALF of
// every B tile (128 of the 256 rows); the leader CTA's single thread issues
// tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 32), which feeds both SMs' tensor cores
// and writes 128 x 256 accumulators into each CTA's own TMEM.  Operand traffic per MMA and
// TMA fill traffic per SM are both halved against the single-CTA kernel, whose SS-mode
// 128x128x32 MMAs saturate the 128 B/clk of shared-memory bandwidth (profiles/r1).
//
// Barrier placement: b_full / a_full / acc_empty live in the LEADER (the only waiter is its
// MMA warp; the peer's TMA transactions and epilogue arrivals reach them through the
// cluster address space); b_empty / a_empty / acc_full are multicast by tcgen05.commit to
// the same offset in both CTAs.
// ======================================================================================
constexpr int kPairTileN = 256;
constexpr int kPairAccStages = 2;
constexpr int kPairPanel = kBlock + 16;  // per epilogue group: column norms of a block + per-tile minima
constexpr int kPairsPerBlock = kStripsPerBlock / 2;
constexpr int kPairTilesPerBlock = kBlock / kPairTileN;
constexpr uint32_t kIdescPair = (2u << 4) | ((kPairTileN >> 3) << 17) | ((256u >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(unsigned long long* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair_leader_only(unsigned long long* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(1))
      : "memory");
}
__device__ __forceinline__ void tc_mma_i8_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Two ways to hand the (strip pair, column block) units of a shard to the CTA pairs:
//   contiguous  every CTA pair takes a contiguous range of the strip-pair-major unit sequence (few A reloads, perfect
//               balance; the way to split a shard of FEWER strip pairs than CTA pairs, e.g. one chunk of a top-n
//               re-emission, and the rows of the last partial wave below);
//   interleaved whole strip-pair rows dealt round-robin, serpentine from wave to wave so that the row lengths even out.
//               All CTA pairs then sit in the same wave of <= 74 neighbouring strip pairs and walk their column blocks
//               J = I, I + 1, ... side by side, a window of a few dozen column blocks apart: the B tiles they stream
//               are shared through L2 instead of being fetched from HBM once per strip pair.  That is what decides
//               the speed once the matrix outgrows L2 (k >= 9: 268 MB and up; at k = 10 the contiguous order moved
//               2.1 TB of B tiles per pass and ran at the HBM rate, not the tensor rate).
struct PairWalk {
  int pair, i_block, j_block;  // global strip-pair index (strips 2*pair, 2*pair+1), its row block, column block
  long long left;              // units still to do (both phases)
  // phase 0: whole rows of the full waves, dealt serpentine; phase 1: a contiguous range of the remaining rows' units
  int wave, n_waves, slot, n_slots;
  long long tail_first, tail_units;
};
// units of the shard's strip pairs [ib0 * kPairsPerBlock, sp), strip-pair-major
__device__ __forceinline__ long long pair_units_before(const KfArgs& a, int sp) {
  const long long il = sp / kPairsPerBlock - a.ib0;
  return kPairsPerBlock * (il * (a.n_jblocks - a.ib0) - il * (il - 1) / 2) +
         static_cast<long long>(sp % kPairsPerBlock) * (a.n_jblocks - sp / kPairsPerBlock);
}
__device__ __forceinline__ int serpentine_pair(const KfArgs& a, int wave, int slot, int n_slots) {
  return a.ib0 * kPairsPerBlock + wave * n_slots + ((wave & 1) ? n_slots - 1 - slot : slot);
}
// position the walk on unit `u` (strip-pair-major index inside the shard)
__device__ void pair_walk_seek(const KfArgs& a, PairWalk& w, long long u) {
  int lo = a.ib0 * kPairsPerBlock, hi = a.ib1 * kPairsPerBlock;  // last strip pair whose first unit is <= u
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (pair_units_before(a, mid) <= u) lo = mid;
    else hi = mid;
  }
  w.pair = lo;
  w.i_block = lo / kPairsPerBlock;
  w.j_block = w.i_block + static_cast<int>(u - pair_units_before(a, lo));
}
__device__ PairWalk pair_walk_begin(const KfArgs& a, int slot, int n_slots) {
  PairWalk w;
  w.slot = slot;
  w.n_slots = n_slots;
  w.wave = 0;
  const int sp0 = a.ib0 * kPairsPerBlock, sp1 = a.ib1 * kPairsPerBlock;
  w.n_waves = a.interleave ? (sp1 - sp0) / n_slots : 0;
  w.left = 0;
  for (int wave = 0; wave < w.n_waves; ++wave)
    w.left += a.n_jblocks - serpentine_pair(a, wave, slot, n_slots) / kPairsPerBlock;
  // what the full waves leave over, split evenly unit by unit
  const long long tail0 = pair_units_before(a, sp0 + w.n_waves * n_slots), tail_all = pair_units_before(a, sp1) - tail0;
  const long long base = tail_all / n_slots, extra = tail_all % n_slots;
  w.tail_first = tail0 + base * slot + (slot < extra ? slot : extra);
  w.tail_units = base + (slot < extra ? 1 : 0);
  w.left += w.tail_units;
  if (w.n_waves > 0) {
    w.pair = serpentine_pair(a, 0, slot, n_slots);
    w.i_block = w.pair / kPairsPerBlock;
    w.j_block = w.i_block;
  } else if (w.tail_units > 0) {
    pair_walk_seek(a, w, w.tail_first);
  }
  return w;
}
__device__ __forceinline__ void pair_walk_next(const KfArgs& a, PairWalk& w) {
  --w.left;
  if (++w.j_block < a.n_jblocks) return;
  if (w.wave < w.n_waves) {          // phase 0: next row of the serpentine, or over to the tail
    if (++w.wave < w.n_waves) {
      w.pair = serpentine_pair(a, w.wave, w.slot, w.n_slots);
      w.i_block = w.pair / kPairsPerBlock;
      w.j_block = w.i_block;
    } else if (w.left > 0) {
      pair_walk_seek(a, w, w.tail_first);
    }
    return;
  }
  ++w.pair;                            // phase 1: contiguous
  w.i_block = w.pair / kPairsPerBlock;
  w.j_block = w.i_block;
}
__device__ __forceinline__ int pair_first_tile(const PairWalk& w) {
  return w.i_block == w.j_block ? (w.pair % kPairsPerBlock) : 0;  // 256-column tile holding the pair's diagonal
}
__device__ __forceinline__ int pair_last_tile(const KfArgs& a, const PairWalk& w) {
  const long long cols = a.n - static_cast<long long>(w.j_block) * kBlock;
  const long long t = (cols + kPairTileN - 1) / kPairTileN;
  return static_cast<int>(t < kPairTilesPerBlock ? t : kPairTilesPerBlock);
}

// Score-class histogram inside the epilogue (HIST): a candidate's class is (dot, sq_s * sq_t).  When every row sums to
// m (KfArgs::hist_m) a squared norm is m + 2 * id with id < kHistSq for m <= 6, so classes with dot < kHistDots fit a
// compact per-CTA table [id_s][id_t][dot] of 32-bit counters in shared memory (8 KB); the by far most frequent class
// of a thread -- dot 1 against an all-distinct column (sq_t = m) -- is counted in a register and folded in once per
// unit, because 20 lanes of a warp hitting one shared-memory word serialise.  Everything else (larger dots, rows that
// break the promise) goes to the global table directly.  The table is added to the global one when the CTA ends.
constexpr int kHistSq = 16, kHistDots = 8, kHistBins = kHistSq * kHistSq * kHistDots;
struct PairStage {       // per epilogue warp: the dots of the current 32-column chunk (stage_get_dot); records go
  uint32_t dotw[8][32];  // straight to global memory through block_commit
};

template <bool HIST>
__global__ void __launch_bounds__(kThreads, 1)
kf_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmap, KfArgs a, int n_bstages, long long total_units,
                       int* error_flag, Geom g) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int kKBlock = 128;
  constexpr int kStageBytes = kTileM * kKBlock;  // each CTA stages 128 rows of A or of its half of B
  const int n_kblocks = g.n_kblocks;
  unsigned char* smem_a = smem;
  unsigned char* smem_b = smem_a + static_cast<size_t>(n_kblocks) * kStageBytes;
  uint16_t* sqn = reinterpret_cast<uint16_t*>(smem_b + static_cast<size_t>(n_bstages) * kStageBytes);
  PairStage* stages = reinterpret_cast<PairStage*>(sqn + kEpiGroups * kPairPanel);
  Barriers* bars = reinterpret_cast<Barriers*>(stages + kEpiThreads / 32);
  uint32_t* sh_hist = reinterpret_cast<uint32_t*>(bars + 1);  // HIST only: [kHistBins]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t hist_m = HIST ? static_cast<uint32_t>(a.hist_m) : 0u;
  const bool hist_compact = HIST && hist_m >= 1 && hist_m * hist_m - hist_m <= 2u * (kHistSq - 1);
  if (HIST) {
    for (int i = tid; i < kHistBins; i += kThreads) sh_hist[i] = 0u;
  }

  if (tid == 0) {
    mbar_init(&bars->a_full, 1);
    mbar_init(&bars->a_empty, 1);
    for (int i = 0; i < n_bstages; ++i) {
      mbar_init(&bars->b_full[i], 1);
      mbar_init(&bars->b_empty[i], 1);
    }
    for (int i = 0; i < kPairAccStages; ++i) {
      mbar_init(&bars->acc_full[i], 1);
      mbar_init(&bars->acc_empty[i], 2 * 2 * (kGroupThreads / 32));  // 2 CTAs x 2 half tiles x 4 warps
    }
    bars->error = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before anyone signals across the pair
  tc_fence_after();
  if (bars->tmem_base != 0) {
    if (tid == 0) *error_flag = 2;
    __trap();
  }
  constexpr uint32_t tmem_base = 0;

  const uint32_t n_pairs_grid = gridDim.x / 2, pair_id = blockIdx.x / 2;
  const long long n_units = total_units;  // > 0 (checked by the launcher); every role derives its own share in pair_walk_begin

  if (warp == kProducerWarp) {
    // ===================== TMA producer (both CTAs; transactions complete on the leader's barriers) ==========
    if (n_units > 0) {
      if (elect_one()) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
      const uint32_t a_full_leader = map_to_cta(smem_u32(&bars->a_full), 0);
      PairWalk w = pair_walk_begin(a, static_cast<int>(pair_id), static_cast<int>(n_pairs_grid));
      int cur_pair = -1;
      uint32_t a_round = 0, stage = 0, round = 0;
      for (; w.left > 0; pair_walk_next(a, w)) {
        if (w.pair != cur_pair) {
          if (a_round > 0) mbar_wait_warp(&bars->a_empty, (a_round - 1) & 1, &bars->error);
          if (elect_one()) {
            if (leader) mbar_expect_tx(&bars->a_full, 2u * static_cast<uint32_t>(n_kblocks) * kStageBytes);
            for (int kb = 0; kb < n_kblocks; ++kb)
              tma_load_2d_pair(smem_a + static_cast<size_t>(kb) * kStageBytes, &tmap, a_full_leader, kb * kKBlock,
                               (2 * w.pair + static_cast<int>(rank)) * kTileM);
          }
          __syncwarp();
          cur_pair = w.pair;
          ++a_round;
        }
        const int t_end = pair_last_tile(a, w);
        for (int ct = pair_first_tile(w); ct < t_end; ++ct) {
          const int col = w.j_block * kBlock + ct * kPairTileN + static_cast<int>(rank) * kTileM;  // this CTA's half of B
          for (int kb = 0; kb < n_kblocks; ++kb) {
            if (round > 0) mbar_wait_warp(&bars->b_empty[stage], (round - 1) & 1, &bars->error);
            if (elect_one()) {
              if (leader) mbar_expect_tx(&bars->b_full[stage], 2u * kStageBytes);
              tma_load_2d_pair(smem_b + static_cast<size_t>(stage) * kStageBytes, &tmap,
                               map_to_cta(smem_u32(&bars->b_full[stage]), 0), kb * kKBlock, col);
            }
            __syncwarp();
            if (++stage == static_cast<uint32_t>(n_bstages)) {
              stage = 0;
              ++round;
            }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && n_units > 0) {
      PairWalk w = pair_walk_begin(a, static_cast<int>(pair_id), static_cast<int>(n_pairs_grid));
      int cur_pair = -1;
      uint32_t a_round = 0, stage = 0, round = 0, acc = 0, acc_round = 0;
      const uint32_t desc_hi = g.desc_hi;
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem_a));
      const uint32_t b_lo0 = umma_desc_lo(smem_u32(smem_b));
      for (; w.left > 0; pair_walk_next(a, w)) {
        if (w.pair != cur_pair) {
          if (cur_pair >= 0 && elect_one()) tc_commit_pair(&bars->a_empty);
          __syncwarp();
          mbar_wait_warp(&bars->a_full, a_round & 1, &bars->error);
          tc_fence_after();
          cur_pair = w.pair;
          ++a_round;
        }
        const int t_end = pair_last_tile(a, w);
        for (int ct = pair_first_tile(w); ct < t_end; ++ct) {
          if (acc_round > 0) mbar_wait_warp(&bars->acc_empty[acc], (acc_round - 1) & 1, &bars->error);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * kPairTileN;
          for (int kb = 0; kb < n_kblocks; ++kb) {
            mbar_wait_warp(&bars->b_full[stage], round & 1, &bars->error);
            tc_fence_after();
            const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(kb) * (kStageBytes >> 4);
            const uint32_t b_lo = b_lo0 + stage * (kStageBytes >> 4);
            if (elect_one()) {
#pragma unroll
              for (int k4 = 0; k4 < kKBlock / kUmmaK; ++k4) {
                const uint64_t da = (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + k4 * (kUmmaK >> 4));
                const uint64_t db = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + k4 * (kUmmaK >> 4));
                tc_mma_i8_pair(tmem_d, da, db, kIdescPair, (kb | k4) != 0);
              }
              tc_commit_pair(&bars->b_empty[stage]);
              if (kb == n_kblocks - 1) tc_commit_pair(&bars->acc_full[acc]);
            }
            __syncwarp();
            if (++stage == static_cast<uint32_t>(n_bstages)) {
              stage = 0;
              ++round;
            }
          }
          if (++acc == kPairAccStages) {
            acc = 0;
            ++acc_round;
          }
        }
      }
    }
  } else {
    // ===================== epilogue groups: 128-column half tiles round-robin =====================
    PairStage& ws = stages[warp];
    BlockCursor bc{0ull, 0, 0, false, 0ull};
    const int group = warp / 4;
    const int gt = tid % kGroupThreads;
    uint16_t* my_sqn = sqn + group * kPairPanel;  // [kBlock] column norms + [8] smallest norm per 128-column tile
    if (n_units > 0) {
      PairWalk w = pair_walk_begin(a, static_cast<int>(pair_id), static_cast<int>(n_pairs_grid));
      uint32_t tile_seq = 0;
      // HIST: this thread's compact row slot (kHistBins: the row breaks the promise) and its register-counted class
      uint32_t h_row = kHistBins, h_mode = 0xFFFFFFFFu, h_priv = 0;
      for (; w.left > 0; pair_walk_next(a, w)) {
        const long long col0 = static_cast<long long>(w.j_block) * kBlock;
        const int cols_in_block = static_cast<int>(min(static_cast<long long>(kBlock), a.n - col0));
        const int strip = 2 * w.pair + static_cast<int>(rank);
        const int strip_local = strip % kStripsPerBlock;
        const int r = strip_local * kStrip + gt;
        const long long s = static_cast<long long>(strip) * kStrip + gt;
        const bool row_ok = s < a.n;
        const bool diag = w.i_block == w.j_block;
        uint32_t sq_i = 0, scale_i = 0;
        bool panel_loaded = false;
        const int t_end = pair_last_tile(a, w);
        for (int ct = pair_first_tile(w); ct < t_end; ++ct, ++tile_seq) {
          const uint32_t acc = tile_seq % kPairAccStages;
          const uint32_t acc_parity = (tile_seq / kPairAccStages) & 1;
          for (int half = 0; half < 2; ++half) {
            if ((2 * tile_seq + half) % kEpiGroups != static_cast<uint32_t>(group)) continue;
            const int ct128 = 2 * ct + half;  // rank tile inside the block
            // nothing to score: the half tile lies below this strip's diagonal or beyond the last node
            const bool skip = (diag && ct128 < strip_local) || ct128 * kTile >= cols_in_block;
            mbar_wait(&bars->acc_full[acc], acc_parity, &bars->error);
            tc_fence_after();
            if (!skip) {
              if (!panel_loaded) {
                asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kGroupThreads) : "memory");
                for (int c = gt; c < kBlock; c += kGroupThreads)
                  my_sqn[c] = c < cols_in_block ? static_cast<uint16_t>(__ldg(a.sqnorm + col0 + c)) : 0;
                asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kGroupThreads) : "memory");
                {  // smallest squared norm per 128-column tile (real columns only): warp q of the group takes tiles 2q, 2q+1
                  const int q = warp & 3;
#pragma unroll
                  for (int t2 = 0; t2 < 2; ++t2) {
                    const int tile = 2 * q + t2;
                    uint32_t mn = 0xFFFFu;
#pragma unroll
                    for (int e = 0; e < kTile / 32; ++e) {
                      const int c = tile * kTile + e * 32 + lane;
                      const uint32_t x = my_sqn[c];
                      if (c < cols_in_block && x < mn) mn = x;
                    }
                    mn = __reduce_min_sync(0xFFFFFFFFu, mn);
                    if (lane == 0) my_sqn[kBlock + tile] = static_cast<uint16_t>(mn);
                  }
                }
                asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kGroupThreads) : "memory");
                if (row_ok) {
                  sq_i = static_cast<uint32_t>(__ldg(a.sqnorm + s));
                  scale_i = __ldg(a.fast_scale + sq_i);
                }
                if (HIST) {
                  const uint32_t id = (sq_i - hist_m) >> 1;
                  const bool fits = hist_compact && row_ok && sq_i >= hist_m && !((sq_i - hist_m) & 1u) && id < kHistSq;
                  h_row = fits ? id * kHistSq * kHistDots : static_cast<uint32_t>(kHistBins);
                  h_mode = fits ? h_row + 1u : 0xFFFFFFFFu;  // (id_t = 0, dot = 1)
                }
                panel_loaded = true;
              }
              const uint32_t taddr =
                  tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + acc * kPairTileN + half * kTile;
              // First-level filter, one compare per score: dmin grows with the norm product, so no column of this
              // tile can pass below dmin[sq_i * smallest column norm of the tile].  (A tile without a real column
              // keeps the 0xFFFF sentinel and is skipped above.)
              const uint32_t tile_min_sq = my_sqn[kBlock + ct128];
              const uint32_t tmin = (row_ok && tile_min_sq != 0xFFFFu) ? __ldg(a.dmin + sq_i * tile_min_sq) : 0xFFFFu;
              uint32_t tile_cnt = 0;
#pragma unroll 1
              for (int cc = 0; cc < kTile; cc += 32) {
                const int c0 = ct128 * kTile + cc;
                uint32_t v[32];
                tc_ld32(taddr + cc, v);
                bool hot = false;
#pragma unroll
                for (int b = 0; b < 32; ++b) hot |= v[b] >= tmin;
                if (!__any_sync(0xFFFFFFFFu, hot)) continue;
                // Second level: the same bound per column (3 instructions a score; the tighter pre-filter through
                // fast_scale cost 8 and, once candidates are dense -- a top-n re-emission at threshold 0 passes the
                // first level in every chunk --, made the epilogue the kernel: 574 warp-instructions per 32 x 32
                // scores, tensor pipe 34 % active).  The exact test below sorts out what the bound lets through.
                uint32_t mask = 0;
#pragma unroll
                for (int b = 0; b < 32; ++b) mask |= static_cast<uint32_t>(v[b] >= tmin) << b;
                mask &= chunk_col_mask(cols_in_block, c0);
                if (diag) mask &= chunk_diag_mask(r, c0);
                if (!row_ok) mask = 0;
                if (__any_sync(0xFFFFFFFFu, mask != 0)) {
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    ws.dotw[j][lane] = v[4 * j] | (v[4 * j + 1] << 8) | (v[4 * j + 2] << 16) | (v[4 * j + 3] << 24);
                  uint32_t exact = 0;
                  for (uint32_t m2 = mask; m2; m2 &= m2 - 1) {
                    const int b = __ffs(m2) - 1;
                    const uint32_t dot = stage_get_dot(ws, b, lane), sq_j = my_sqn[c0 + b];
                    // a column with the tile's smallest norm (most of them: all-distinct rows) has dmin == tmin
                    if (sq_j == tile_min_sq || dot >= __ldg(a.dmin + sq_i * sq_j)) {
                      exact |= 1u << b;
                      if (HIST) {
                        const uint32_t idj = (sq_j - hist_m) >> 1;
                        if (h_row < kHistBins && dot < kHistDots && sq_j >= hist_m && !((sq_j - hist_m) & 1u) && idj < kHistSq) {
                          const uint32_t key = h_row + idj * kHistDots + dot;
                          if (key == h_mode) ++h_priv;
                          else atomicAdd(sh_hist + key, 1u);
                        } else {
                          atomicAdd(a.class_hist + static_cast<size_t>(dot) * (a.p_max + 1) + sq_i * sq_j, 1ull);
                        }
                      }
                    }
                  }
                  block_commit(bc, ws, a.out, lane, exact, static_cast<uint32_t>(s), static_cast<uint32_t>(col0 + c0),
                               tile_cnt);
                }
              }
              if (row_ok) a.rowcnt[rowcnt_index(w.i_block - a.ib0, a.n_jblocks, w.j_block, r, ct128)] =
                  static_cast<uint8_t>(tile_cnt);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(&bars->acc_empty[acc]), 0));
          }
        }
        if (HIST && h_priv) {  // the row (and with it h_mode) may change with the next unit
          atomicAdd(sh_hist + h_mode, h_priv);
          h_priv = 0;
        }
      }
      block_finish(bc, a.out, lane);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (HIST) {  // fold this CTA's compact table into the global (dot, P) table
    for (int i = tid; i < kHistBins; i += kThreads) {
      const uint32_t c = sh_hist[i];
      if (!c) continue;
      const uint32_t dot = i % kHistDots, idj = (i / kHistDots) % kHistSq, idi = i / (kHistDots * kHistSq);
      const uint32_t p = (hist_m + 2u * idi) * (hist_m + 2u * idj);
      atomicAdd(a.class_hist + static_cast<size_t>(dot) * (a.p_max + 1) + p, static_cast<unsigned long long>(c));
    }
  }
  cluster_sync_all();  // the peer may still be signalling this CTA's barriers / reading the pair's operands
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
  if (tid == 0 && bars->error) *error_flag = 1;
}


// ======================================================================================
// Tensor-pipe ceiling probe (bench.py): the CTA-pair kernel's MMA stream with nothing around it -- resident operands
// (an 8 x 128-byte-K-block A strip and one B stage per CTA, filled once), no TMA refills, no epilogue.  One leader
// thread per CTA pair issues `iters` tiles of n_kblocks x 4 tcgen05.mma.cta_group::2.kind::i8 (M = 256, N = 256,
// K = 32) into two alternating TMEM stages.  Timed from the host with CUDA events; the rate is the denominator the
// kind::i8 roofline fraction of kf_tcgen05_pair_kernel is quoted against, measured at the clocks the bench runs at.
// ======================================================================================
__global__ void __launch_bounds__(128, 1) tensor_probe_kernel(int iters, int n_kblocks, uint32_t desc_hi, int* error_flag) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int kStageBytes = kTileM * 128;
  unsigned char* smem_a = smem;
  unsigned char* smem_b = smem_a + static_cast<size_t>(n_kblocks) * kStageBytes;
  Barriers* bars = reinterpret_cast<Barriers*>(smem_b + kStageBytes);
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool leader = cluster_ctarank() == 0;
  for (int i = tid; i < (n_kblocks + 1) * kStageBytes / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = (static_cast<uint32_t>(i) * 2654435761u >> 7) & 0x03030303u;  // counts 0..3
  if (tid == 0) {
    mbar_init(&bars->a_full, 1);
    bars->error = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy fills -> visible to the tensor core
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  if (warp == 0 && leader) {
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem_a)), b_lo = umma_desc_lo(smem_u32(smem_b));
    for (int it = 0; it < iters; ++it) {
      const uint32_t tmem_d = tmem_base + (it & 1) * kPairTileN;
      for (int kb = 0; kb < n_kblocks; ++kb) {
        const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(kb) * (kStageBytes >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k4 = 0; k4 < 128 / kUmmaK; ++k4) {
            const uint64_t da = (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + k4 * (kUmmaK >> 4));
            const uint64_t db = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + k4 * (kUmmaK >> 4));
            tc_mma_i8_pair(tmem_d, da, db, kIdescPair, (kb | k4) != 0);
          }
        }
        __syncwarp();
      }
    }
    if (elect_one()) tc_commit_pair(&bars->a_full);  // fires in both CTAs once every MMA above has retired
    __syncwarp();
  }
  mbar_wait(&bars->a_full, 0, &bars->error);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
  if (tid == 0 && bars->error) *error_flag = 1;
}


// ======================================================================================
// Per-row kNN KF edges (the faiss_ann mode, reference gnn_utils.py:122-237) on the same CTA-pair Gram: the FULL matrix
// this time -- every row against every column -- because a row's neighbours lie on both sides of the diagonal.
// A CTA pair owns strip pairs of 256 rows (A resident) and streams all columns past them; a thread of an epilogue
// group is one row, as above.  Ranks come from the host table rank_tab[class of sq_j][dot] (gg_knn.cu).
//   pass 1  per-row rank histograms in shared memory.  Only non-zero dots cost per-pair work (1.5 % of the pairs of
//           1024-column rows): the 32 dots of a chunk are packed four to a word with byte permutes, a lane looks only at
//           its non-zero words, and the zero-dot members of a column class are its size minus the row's non-zero count
//           for that class, added when the strip is finished.  All three epilogue groups, half tiles round-robin.
//   pass 2  the selected pairs of every row in COLUMN order (ties of the cut rank are taken in that order while the
//           row's budget lasts).  The columns are cut into three contiguous thirds, the tiles of the thirds interleaved
//           on the tensor pipe, and epilogue group g walks third g in order: a row's running position and tie count
//           are registers of its thread in that group.  Every third keeps the ties its own budget allows; the reorder
//           kernel (gg_knn.cu) reads the thirds in order and applies the budget across them.  Zero-dot columns are
//           selected through per-chunk class masks built once per half tile.
// ======================================================================================
constexpr uint32_t kKnnNoRank = 0xFFFFu;

__device__ __forceinline__ void bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// low bytes of four accumulators -> one word (dots are <= 255)
__device__ __forceinline__ uint32_t pack4(uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3) {
  return __byte_perm(__byte_perm(v0, v1, 0x0040), __byte_perm(v2, v3, 0x0040), 0x5410);
}
// bit 7 of every non-zero byte
__device__ __forceinline__ uint32_t nonzero_bytes(uint32_t w) {
  return (((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w) & 0x80808080u;
}
// the j-th of eight registers (j is not a compile-time constant: a select tree instead of local memory)
__device__ __forceinline__ uint32_t pick8(const uint32_t (&w)[8], int j) {
  const uint32_t x0 = (j & 1) ? w[1] : w[0], x1 = (j & 1) ? w[3] : w[2], x2 = (j & 1) ? w[5] : w[4], x3 = (j & 1) ? w[7] : w[6];
  const uint32_t y0 = (j & 2) ? x1 : x0, y1 = (j & 2) ? x3 : x2;
  return (j & 4) ? y1 : y0;
}

// Pass 2 deals 128-column blocks, not tiles: half h of the issue sequence (tile h / 2, CTA h % 2 of the pair loads it)
// is the next block of column third h % 3, so that with two accumulator tiles in flight all three epilogue groups --
// group g owns the halves with h % 3 == g -- have work.  The two halves of a B tile need not be neighbours: each CTA
// of the pair loads its own.
struct KnnBlocks {
  int n_blk, third;  // 128-column blocks of the matrix, blocks per third
  __device__ __forceinline__ int tiles() const { return (3 * third + 1) / 2; }
  // first column of half h; beyond the last column (TMA zero fill, nothing to score) when the third has no such block
  __device__ __forceinline__ long long col(int h, long long n) const {
    const int i = h / 3, blk = (h % 3) * third + i;
    return (i < third && blk < n_blk) ? static_cast<long long>(blk) * kTile : (n + kTile - 1) / kTile * kTile;
  }
};

template <int PASS>
__global__ void __launch_bounds__(kThreads, 1)
knn_tc_pair_kernel(const __grid_constant__ CUtensorMap tmap, KnnTcArgs a, int n_bstages, int* error_flag, Geom g) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int kKBlock = 128;
  constexpr int kStageBytes = kTileM * kKBlock;
  const int n_kblocks = g.n_kblocks;
  unsigned char* smem_a = smem;
  unsigned char* smem_b = smem_a + static_cast<size_t>(n_kblocks) * kStageBytes;
  uint8_t* panel = smem_b + static_cast<size_t>(n_bstages) * kStageBytes;         // [kEpiGroups][kBlock] column classes
  Barriers* bars = reinterpret_cast<Barriers*>(panel + kEpiGroups * kBlock);
  const int tstride = a.dot_max + 1;
  const int tab_elems = (a.n_sq + 1) * tstride;                                    // + an all-"never" class
  uint16_t* rank_tab = reinterpret_cast<uint16_t*>(bars + 1);
  uint32_t* colcount = reinterpret_cast<uint32_t*>(rank_tab + ((tab_elems + 1) & ~1));  // [kKnnTcMaxSq]
  // pass 1: nz [128][n_sq], shist [128][hwords].  pass 2: class masks [kEpiGroups][2][4][kKnnTcMaxSq], then the packed
  // dots of the current chunk per epilogue warp (PairStage)
  uint32_t* nz = colcount + kKnnTcMaxSq;
  uint32_t* shist = nz + kTileM * a.n_sq;
  uint32_t* clsmask = colcount + kKnnTcMaxSq;
  PairStage* stages = reinterpret_cast<PairStage*>(clsmask + kEpiGroups * 2 * 4 * kKnnTcMaxSq);
  const int hwords = a.hist16 ? (a.n_ranks + 1) >> 1 : a.n_ranks;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  for (int i = tid; i < tab_elems; i += kThreads)
    rank_tab[i] = i < a.n_sq * tstride ? a.rank_tab[i] : static_cast<uint16_t>(kKnnNoRank);
  if (tid < kKnnTcMaxSq) colcount[tid] = 0u;
  if (PASS == 1) {
    for (int i = tid; i < kTileM * (a.n_sq + hwords); i += kThreads) nz[i] = 0u;
  }
  if (tid == 0) {
    mbar_init(&bars->a_full, 1);
    mbar_init(&bars->a_empty, 1);
    for (int i = 0; i < n_bstages; ++i) {
      mbar_init(&bars->b_full[i], 1);
      mbar_init(&bars->b_empty[i], 1);
    }
    for (int i = 0; i < kPairAccStages; ++i) {
      mbar_init(&bars->acc_full[i], 1);
      // pass 1: every epilogue warp of both CTAs takes part in every tile; pass 2: 2 CTAs x 2 half tiles x 4 warps
      mbar_init(&bars->acc_empty[i], PASS == 1 ? 2 * (kEpiThreads / 32) : 2 * 2 * (kGroupThreads / 32));
    }
    bars->error = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (PASS == 1) {  // columns per class (real columns only): a warp counts 32 columns per step with one ballot per class
    uint32_t mine = 0;
    for (long long j0 = static_cast<long long>(warp) * 32; j0 < a.n; j0 += (kThreads / 32) * 32) {
      const long long j = j0 + lane;
      uint32_t cls = 0xFFu;
      if (j < a.n) {
        const int sq = __ldg(a.sqnorm + j);
        if (sq >= 0 && sq <= a.sq_max) cls = __ldg(a.sqid + sq);
      }
      for (int c = 0; c < a.n_sq; ++c) {
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, cls == static_cast<uint32_t>(c));
        if (lane == c) mine += __popc(m);
      }
    }
    if (lane < a.n_sq && mine) atomicAdd(colcount + lane, mine);
    __syncthreads();
  }
  cluster_sync_all();  // both CTAs' barriers are initialised before anyone signals across the pair
  tc_fence_after();
  if (bars->tmem_base != 0) {
    if (tid == 0) *error_flag = 2;
    __trap();
  }
  constexpr uint32_t tmem_base = 0;

  const int n_pairs_grid = static_cast<int>(gridDim.x / 2), pair_id = static_cast<int>(blockIdx.x / 2);
  const int n_sp = static_cast<int>((a.row_end - a.row_begin + 2 * kTileM - 1) / (2 * kTileM));  // strip pairs of the shard
  KnnBlocks blocks;
  blocks.n_blk = static_cast<int>((a.n + kTile - 1) / kTile);
  blocks.third = (blocks.n_blk + 2) / 3;
  // pass 1 walks the 256-column tiles in plain order
  const int n_slots = PASS == 1 ? static_cast<int>((a.n + kPairTileN - 1) / kPairTileN) : blocks.tiles();

  if (warp == kProducerWarp) {
    // ===================== TMA producer (both CTAs; transactions complete on the leader's barriers) ==========
    if (elect_one()) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
    const uint32_t a_full_leader = map_to_cta(smem_u32(&bars->a_full), 0);
    uint32_t a_round = 0, stage = 0, round = 0;
    for (int sp = pair_id; sp < n_sp; sp += n_pairs_grid) {
      if (a_round > 0) mbar_wait_warp(&bars->a_empty, (a_round - 1) & 1, &bars->error);
      if (elect_one()) {
        if (leader) mbar_expect_tx(&bars->a_full, 2u * static_cast<uint32_t>(n_kblocks) * kStageBytes);
        const int row = static_cast<int>(a.row_begin) + (2 * sp + static_cast<int>(rank)) * kTileM;
        for (int kb = 0; kb < n_kblocks; ++kb)
          tma_load_2d_pair(smem_a + static_cast<size_t>(kb) * kStageBytes, &tmap, a_full_leader, kb * kKBlock, row);
      }
      __syncwarp();
      ++a_round;
      for (int slot = 0; slot < n_slots; ++slot) {
        // this CTA's half of B
        const int col = PASS == 1 ? slot * kPairTileN + static_cast<int>(rank) * kTileM
                                  : static_cast<int>(blocks.col(2 * slot + static_cast<int>(rank), a.n));
        for (int kb = 0; kb < n_kblocks; ++kb) {
          if (round > 0) mbar_wait_warp(&bars->b_empty[stage], (round - 1) & 1, &bars->error);
          if (elect_one()) {
            if (leader) mbar_expect_tx(&bars->b_full[stage], 2u * kStageBytes);
            tma_load_2d_pair(smem_b + static_cast<size_t>(stage) * kStageBytes, &tmap,
                             map_to_cta(smem_u32(&bars->b_full[stage]), 0), kb * kKBlock, col);
          }
          __syncwarp();
          if (++stage == static_cast<uint32_t>(n_bstages)) {
            stage = 0;
            ++round;
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      uint32_t a_round = 0, stage = 0, round = 0, acc = 0, acc_round = 0;
      const uint32_t desc_hi = g.desc_hi;
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem_a));
      const uint32_t b_lo0 = umma_desc_lo(smem_u32(smem_b));
      for (int sp = pair_id; sp < n_sp; sp += n_pairs_grid) {
        if (a_round > 0 && elect_one()) tc_commit_pair(&bars->a_empty);
        __syncwarp();
        mbar_wait_warp(&bars->a_full, a_round & 1, &bars->error);
        tc_fence_after();
        ++a_round;
        for (int slot = 0; slot < n_slots; ++slot) {
          if (acc_round > 0) mbar_wait_warp(&bars->acc_empty[acc], (acc_round - 1) & 1, &bars->error);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * kPairTileN;
          for (int kb = 0; kb < n_kblocks; ++kb) {
            mbar_wait_warp(&bars->b_full[stage], round & 1, &bars->error);
            tc_fence_after();
            const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(kb) * (kStageBytes >> 4);
            const uint32_t b_lo = b_lo0 + stage * (kStageBytes >> 4);
            if (elect_one()) {
#pragma unroll
              for (int k4 = 0; k4 < kKBlock / kUmmaK; ++k4) {
                const uint64_t da = (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + k4 * (kUmmaK >> 4));
                const uint64_t db = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + k4 * (kUmmaK >> 4));
                tc_mma_i8_pair(tmem_d, da, db, kIdescPair, (kb | k4) != 0);
              }
              tc_commit_pair(&bars->b_empty[stage]);
              if (kb == n_kblocks - 1) tc_commit_pair(&bars->acc_full[acc]);
            }
            __syncwarp();
            if (++stage == static_cast<uint32_t>(n_bstages)) {
              stage = 0;
              ++round;
            }
          }
          if (++acc == kPairAccStages) {
            acc = 0;
            ++acc_round;
          }
        }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int group = warp / 4;
    const int gt = tid % kGroupThreads;  // row of the strip = TMEM lane
    uint8_t* my_panel = panel + group * kBlock;
    uint32_t tile_seq = 0;               // tiles issued so far (every role counts the same sequence)
    for (int sp = pair_id; sp < n_sp; sp += n_pairs_grid) {
      const long long row = a.row_begin + static_cast<long long>(2 * sp + static_cast<int>(rank)) * kTileM + gt;
      const bool row_ok = row < a.row_end;
      const long long loc = row - a.row_begin;
      // pass 2: the row's plan and this group's running state
      int cut = -1, budget = 0, seen = 0;
      long long out_at = 0, out_first = 0;
      uint32_t zsel = 0, ztie = 0;  // classes whose zero-dot columns rank above / at the cut
      if (PASS == 2 && row_ok) {
        cut = __ldg(a.cut + loc);
        budget = __ldg(a.tie_left + loc);
        const long long r0 = __ldg(a.row_off + loc), r1 = __ldg(a.row_off + loc + 1);
        out_first = out_at = 3 * r0 + group * (r1 - r0);   // this group's third of the row's staging space
        for (int c = 0; c < a.n_sq; ++c) {
          const int rk0 = rank_tab[c * tstride];
          zsel |= static_cast<uint32_t>(rk0 < cut) << c;   // 0xFFFF is above every cut
          ztie |= static_cast<uint32_t>(rk0 == cut) << c;
        }
      }
      int panel_block = -1;
      uint32_t mask_buf = 0;  // pass 2: class-mask buffer of this group's next half tile
      for (int slot = 0; slot < n_slots; ++slot) {
        const uint32_t acc = tile_seq % kPairAccStages;
        const uint32_t acc_parity = (tile_seq / kPairAccStages) & 1;
        const uint32_t my_seq = tile_seq++;
        for (int half = 0; half < 2; ++half) {
          // pass 2: group g owns the halves of column third g.  pass 1: every group works on both halves of every tile
          // (the eight 32-column chunks of a tile are dealt round-robin below): with only two accumulator tiles the
          // time one tile spends in the epilogue, not the epilogue's throughput, paces the tensor pipe
          if (PASS == 2 && static_cast<uint32_t>(2 * slot + half) % kEpiGroups != static_cast<uint32_t>(group)) continue;
          // first column of the half tile
          const long long c0 = PASS == 1 ? static_cast<long long>(slot) * kPairTileN + half * kTile
                                         : blocks.col(2 * slot + half, a.n);
          const int jb = static_cast<int>(c0 / kBlock);
          if (jb != panel_block && c0 < a.n) {  // column classes of the next 1024 columns (uniform per group)
            bar_sync(2 + group, kGroupThreads);
            for (int c = gt; c < kBlock; c += kGroupThreads) {
              const long long j = static_cast<long long>(jb) * kBlock + c;
              uint32_t cls = 0xFFu;
              if (j < a.n) {
                const int sq = __ldg(a.sqnorm + j);
                if (sq >= 0 && sq <= a.sq_max) cls = __ldg(a.sqid + sq);
              }
              my_panel[c] = static_cast<uint8_t>(cls == 0xFFu ? a.n_sq : cls);
            }
            bar_sync(2 + group, kGroupThreads);
            panel_block = jb;
          }
          const int pc0 = static_cast<int>(c0 - static_cast<long long>(jb) * kBlock);  // panel index of the half tile
          uint32_t* my_masks = clsmask + (group * 2 + mask_buf) * 4 * kKnnTcMaxSq;
          if (PASS == 2 && c0 < a.n) {
            // class masks of the four 32-column chunks: warp q of the group makes chunk q.  (Two buffers in turn: a
            // warp that is still reading the masks of the previous half tile reads the other one.)
            mask_buf ^= 1u;
            const int q = warp & 3;
            const uint32_t cls = my_panel[pc0 + q * 32 + lane];
            uint32_t mine = 0;
            for (int c = 0; c < a.n_sq; ++c) {
              const uint32_t m = __ballot_sync(0xFFFFFFFFu, cls == static_cast<uint32_t>(c));
              if (lane == c) mine = m;
            }
            my_masks[q * kKnnTcMaxSq + lane] = mine;
            bar_sync(2 + group, kGroupThreads);
          }
          mbar_wait(&bars->acc_full[acc], acc_parity, &bars->error);
          tc_fence_after();
          if (c0 < a.n) {
            const uint32_t taddr =
                tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + acc * kPairTileN + half * kTile;
#pragma unroll 1
            for (int cc = 0; cc < kTile; cc += 32) {
              if (PASS == 1 && (my_seq + 4 * half + (cc >> 5)) % kEpiGroups != static_cast<uint32_t>(group)) continue;
              uint32_t v[32];
              tc_ld32(taddr + cc, v);
              uint32_t w[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) w[j] = pack4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              uint32_t wmask = 0;  // non-zero words
#pragma unroll
              for (int j = 0; j < 8; ++j) wmask |= static_cast<uint32_t>(w[j] != 0u) << j;
              if (!row_ok) wmask = 0;
              const long long cbase = c0 + cc;            // global column of bit 0
              const long long self = row - cbase;          // bit of the row's own column, if inside the chunk
              if (PASS == 1) {
                for (; wmask; wmask &= wmask - 1) {
                  const int j = __ffs(wmask) - 1;
                  const uint32_t word = pick8(w, j);
                  for (uint32_t t = nonzero_bytes(word); t; t &= t - 1) {
                    const int byte = (__ffs(t) - 1) >> 3, b = 4 * j + byte;
                    const uint32_t cls = my_panel[pc0 + cc + b];
                    atomicAdd(nz + gt * a.n_sq + cls, 1u);   // the self pair too: it is not a zero-dot column
                    if (b == self) continue;
                    const uint32_t r = rank_tab[cls * tstride + ((word >> (8 * byte)) & 0xFFu)];
                    if (r == kKnnNoRank) continue;
                    if (a.hist16) atomicAdd(shist + gt * hwords + (r >> 1), 1u << ((r & 1u) * 16));
                    else atomicAdd(shist + gt * hwords + r, 1u);
                  }
                }
              } else {
                PairStage& ws = stages[warp];
                if (__any_sync(0xFFFFFFFFu, wmask != 0u)) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) ws.dotw[j][lane] = w[j];
                }
                uint32_t sel = 0, tie = 0, nzmask = 0;
                for (; wmask; wmask &= wmask - 1) {
                  const int j = __ffs(wmask) - 1;
                  const uint32_t word = ws.dotw[j][lane];
                  for (uint32_t t = nonzero_bytes(word); t; t &= t - 1) {
                    const int byte = (__ffs(t) - 1) >> 3, b = 4 * j + byte;
                    const uint32_t cls = my_panel[pc0 + cc + b];
                    const int r = rank_tab[cls * tstride + ((word >> (8 * byte)) & 0xFFu)];
                    nzmask |= 1u << b;
                    sel |= static_cast<uint32_t>(r < cut) << b;
                    tie |= static_cast<uint32_t>(r == cut) << b;
                  }
                }
                if (zsel | ztie) {
                  uint32_t zs = 0, zt = 0;
                  for (uint32_t zm = zsel; zm; zm &= zm - 1) zs |= my_masks[(cc >> 5) * kKnnTcMaxSq + __ffs(zm) - 1];
                  for (uint32_t zm = ztie; zm; zm &= zm - 1) zt |= my_masks[(cc >> 5) * kKnnTcMaxSq + __ffs(zm) - 1];
                  sel |= zs & ~nzmask;
                  tie |= zt & ~nzmask;
                }
                if (self >= 0 && self < 32) {
                  sel &= ~(1u << self);
                  tie &= ~(1u << self);
                }
                if (tie) {  // members of the cut rank are taken in column order while the row's budget lasts
                  const int t = __popc(tie), room = budget - seen;
                  seen += t;
                  if (room <= 0) {
                    tie = 0;
                  } else {
                    for (int x = t - room; x > 0; --x) tie &= ~(0x80000000u >> __clz(tie));
                  }
                }
                for (uint32_t m2 = sel | tie; m2; m2 &= m2 - 1) {
                  const int b = __ffs(m2) - 1;
                  const uint32_t cls = my_panel[pc0 + cc + b];
                  const uint32_t dot = ((nzmask >> b) & 1u) ? stage_get_dot(ws, b, lane) : 0u;
                  const uint32_t r = rank_tab[cls * tstride + dot];
                  a.recs[out_at++] = static_cast<uint64_t>(cbase + b) | static_cast<uint64_t>(dot) << 32 |
                                     static_cast<uint64_t>(r) << 48;
                }
                __syncwarp();  // the staged dots are rewritten by the next chunk
              }
            }
          }
          if (PASS == 1 && half == 0) continue;   // pass 1: one arrival per warp and tile, after both halves
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(&bars->acc_empty[acc]), 0));
        }
      }
      if (PASS == 2 && row_ok) a.third_cnt[3 * loc + group] = static_cast<uint32_t>(out_at - out_first);
      if (PASS == 1) {
        // the strip is finished: zero-dot members of every class, then the rows go out and the tables are cleared
        bar_sync(1, kEpiThreads);
        if (tid < kTileM && row_ok) {   // group 0: thread <-> row
          // an all-zero row meets itself with dot 0: that column is not a neighbour
          const int sq_row = __ldg(a.sqnorm + row);
          const uint32_t own = sq_row == 0 ? __ldg(a.sqid) : 0xFFu;
          for (int c = 0; c < a.n_sq; ++c) {
            const uint32_t r0 = rank_tab[c * tstride];
            if (r0 == kKnnNoRank) continue;
            const uint32_t z = colcount[c] - nz[tid * a.n_sq + c] - (static_cast<uint32_t>(c) == own ? 1u : 0u);
            if (a.hist16) shist[tid * hwords + (r0 >> 1)] += z << ((r0 & 1u) * 16);
            else shist[tid * hwords + r0] += z;
          }
        }
        bar_sync(1, kEpiThreads);
        {
          const long long first = a.row_begin + static_cast<long long>(2 * sp + static_cast<int>(rank)) * kTileM;
          const long long left = a.row_end - first;
          const int rows_here = left >= kTileM ? kTileM : (left > 0 ? static_cast<int>(left) : 0);
          uint32_t* out = a.hist + (first - a.row_begin) * a.n_ranks;
          for (int i = tid; i < rows_here * a.n_ranks; i += kEpiThreads) {
            const int r = i / a.n_ranks, j = i - r * a.n_ranks;
            out[i] = a.hist16 ? (shist[r * hwords + (j >> 1)] >> ((j & 1) * 16)) & 0xFFFFu : shist[r * hwords + j];
          }
        }
        bar_sync(1, kEpiThreads);
        for (int i = tid; i < kTileM * (a.n_sq + hwords); i += kEpiThreads) nz[i] = 0u;
        bar_sync(1, kEpiThreads);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be signalling this CTA's barriers / reading the pair's operands
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
  if (tid == 0 && bars->error) *error_flag = 1;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace

int emit_tcgen05_pair(const KfArgs& a, cudaStream_t s);

int emit_tcgen05(const KfArgs& a, int p_max, cudaStream_t s) {
  // a.d is the row pitch of `counts` in bytes: 32 or 64 (one K block), or a multiple of 128 up to kMaxD
  GG_REQUIRE(a.d == 32 || a.d == 64 || (a.d % kMaxKBlock == 0 && a.d >= kMaxKBlock && a.d <= kMaxD), GG_ERR_ARG,
             "tcgen05 KF kernel needs a row pitch of 32, 64 or a multiple of 128 up to 1024 bytes");
  EncodeTiledFn encode = encode_tiled();
  GG_REQUIRE(encode != nullptr, GG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");

  Geom g;
  g.kb_bytes = a.d < kMaxKBlock ? a.d : kMaxKBlock;
  g.n_kblocks = a.d / g.kb_bytes;
  g.stage_bytes = kTileN * g.kb_bytes;
  g.mma_per_kb = g.kb_bytes / kUmmaK;
  const CUtensorMapSwizzle swz = g.kb_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                 : (g.kb_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
  const uint32_t layout_type = g.kb_bytes == 32 ? 6u : (g.kb_bytes == 64 ? 4u : 2u);  // UMMA SWIZZLE_32B/64B/128B
  g.desc_hi = static_cast<uint32_t>((8 * g.kb_bytes) >> 4) | (1u << 14) | (layout_type << 29);

  CUtensorMap tmap;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(a.d), static_cast<cuuint64_t>(a.n)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(a.d)};  // row pitch in bytes
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(g.kb_bytes), kTileM};
  const cuuint32_t elem[2] = {1, 1};
  const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(a.counts), dims, strides, box,
                            elem, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    gg::set_error("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return GG_ERR_CUDA;
  }

  int sq_max = 0;  // p_max = sq_max^2 (core.threshold_tables); the narrow epilogue tabulates dmin per row
  while (static_cast<long long>(sq_max + 1) * (sq_max + 1) <= p_max) ++sq_max;
  const bool narrow = g.kb_bytes < kMaxKBlock;
  GG_REQUIRE(!narrow || (static_cast<long long>(sq_max) * sq_max == p_max && sq_max <= 255 && sq_max * kGroupThreads < 65536),
             GG_ERR_ARG, "p_max must be a square with sq_max <= 255 for narrow rows");
  const size_t fixed = static_cast<size_t>(g.n_kblocks) * g.stage_bytes + kEpiGroups * kBlock * sizeof(uint16_t) +
                       (kEpiThreads / 32) * (narrow ? sizeof(WarpStageT<kNarrowStageCap>) : sizeof(WarpStage)) +
                       sizeof(Barriers) + (narrow ? static_cast<size_t>(kEpiGroups) * (sq_max + 1) * kGroupThreads : 0) +
                       1024 /* base alignment slack */;
  int dev = 0, max_smem = 0;
  GG_CUDA_CHECK(cudaGetDevice(&dev));
  GG_CUDA_CHECK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  int n_bstages = static_cast<int>((static_cast<size_t>(max_smem) - fixed) / g.stage_bytes);
  if (n_bstages > kMaxBStages) n_bstages = kMaxBStages;
  GG_REQUIRE(n_bstages >= 2, GG_ERR_ARG, "not enough shared memory for the B ring");
  const size_t smem_bytes = fixed + static_cast<size_t>(n_bstages) * g.stage_bytes;
  auto kernel = g.kb_bytes == 32 ? kf_tcgen05_kernel<32> : (g.kb_bytes == 64 ? kf_tcgen05_kernel<64> : kf_tcgen05_kernel<128>);
  GG_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes)));

  // units of this shard: 8 strips per row block, n_j - I column blocks each
  long long total_units = 0;
  for (int ib = a.ib0; ib < a.ib1; ++ib) total_units += static_cast<long long>(kStripsPerBlock) * (a.n_jblocks - ib);
  if (total_units == 0) return GG_OK;
  const long long sms = gg::sm_count();
  const unsigned grid = static_cast<unsigned>(total_units < sms ? total_units : sms);
  int* error_flag = reinterpret_cast<int*>(a.out.status + 2);  // kf_status[2]
  kernel<<<grid, kThreads, smem_bytes, s>>>(tmap, a, n_bstages, total_units, error_flag, g, sq_max);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}

int emit_tcgen05_pair(const KfArgs& a, cudaStream_t s) {
  GG_REQUIRE(a.d % kMaxKBlock == 0 && a.d >= kMaxKBlock && a.d <= kMaxD, GG_ERR_ARG,
             "CTA-pair tcgen05 KF kernel needs a row pitch that is a multiple of 128 up to 1024 bytes");
  EncodeTiledFn encode = encode_tiled();
  GG_REQUIRE(encode != nullptr, GG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  Geom g;
  g.kb_bytes = kMaxKBlock;
  g.n_kblocks = a.d / g.kb_bytes;
  g.stage_bytes = kTileM * g.kb_bytes;
  g.mma_per_kb = g.kb_bytes / kUmmaK;
  g.desc_hi = static_cast<uint32_t>((8 * g.kb_bytes) >> 4) | (1u << 14) | (2u << 29);
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(a.d), static_cast<cuuint64_t>(a.n)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(a.d)};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(g.kb_bytes), kTileM};
  const cuuint32_t elem[2] = {1, 1};
  const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(a.counts), dims, strides, box,
                            elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    gg::set_error("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return GG_ERR_CUDA;
  }
  const bool hist = a.class_hist != nullptr;
  const size_t fixed = static_cast<size_t>(g.n_kblocks) * g.stage_bytes + kEpiGroups * kPairPanel * sizeof(uint16_t) +
                       (kEpiThreads / 32) * sizeof(PairStage) + sizeof(Barriers) +
                       (hist ? kHistBins * sizeof(uint32_t) : 0) + 1024;
  int dev = 0, max_smem = 0;
  GG_CUDA_CHECK(cudaGetDevice(&dev));
  GG_CUDA_CHECK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  int n_bstages = static_cast<int>((static_cast<size_t>(max_smem) - fixed) / g.stage_bytes);
  if (n_bstages > kMaxBStages) n_bstages = kMaxBStages;
  GG_REQUIRE(n_bstages >= 2, GG_ERR_ARG, "not enough shared memory for the B ring");
  const size_t smem_bytes = fixed + static_cast<size_t>(n_bstages) * g.stage_bytes;
  auto kernel = hist ? kf_tcgen05_pair_kernel<true> : kf_tcgen05_pair_kernel<false>;
  GG_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes)));
  long long total_units = 0;  // (strip pair, column block) units of this shard
  for (int ib = a.ib0; ib < a.ib1; ++ib) total_units += static_cast<long long>(kPairsPerBlock) * (a.n_jblocks - ib);
  if (total_units == 0) return GG_OK;
  const long long pairs = gg::sm_count() / 2;
  const unsigned grid = 2u * static_cast<unsigned>(total_units < pairs ? total_units : pairs);
  // A matrix that does not fit L2 and enough strip-pair rows for every CTA pair to take whole ones: deal the rows
  // wave by wave (shared B window in L2); the rows of the last, partial wave are split unit by unit as before.
  KfArgs args = a;
  args.interleave = (static_cast<double>(a.n) * a.d > 96e6 &&
                     static_cast<long long>(a.ib1 - a.ib0) * kPairsPerBlock >= 2 * pairs) ? 1 : 0;
  int* error_flag = reinterpret_cast<int*>(a.out.status + 2);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  GG_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, tmap, args, n_bstages, total_units, error_flag, g));
  return GG_OK;
}


// ---- per-row kNN on the CTA-pair Gram ------------------------------------------------------------------
__device__ int g_knn_tc_error;  // set before a trap (mbarrier time-out, TMEM base): the launch then fails as a whole

static size_t knn_tc_fixed_smem(const KnnTcArgs& a, int pass) {
  const int n_kblocks = a.d / kMaxKBlock;
  const int tab_elems = (a.n_sq + 1) * (a.dot_max + 1);
  const int hwords = a.hist16 ? (a.n_ranks + 1) / 2 : a.n_ranks;
  size_t bytes = static_cast<size_t>(n_kblocks) * kTileM * kMaxKBlock + kEpiGroups * kBlock + sizeof(Barriers) +
                 2 * static_cast<size_t>((tab_elems + 1) & ~1) + kKnnTcMaxSq * sizeof(uint32_t) + 1024;
  if (pass == 1) bytes += static_cast<size_t>(kTileM) * (a.n_sq + hwords) * sizeof(uint32_t);
  else bytes += kEpiGroups * 2 * 4 * kKnnTcMaxSq * sizeof(uint32_t) + (kEpiThreads / 32) * sizeof(PairStage);
  return bytes;
}

bool knn_tc_eligible(int pitch, long long n, int n_sq, int dot_max, int n_ranks, int pass) {
  if (pitch % kMaxKBlock != 0 || pitch < kMaxKBlock || pitch > kMaxD) return false;
  if (n_sq < 1 || n_sq > kKnnTcMaxSq || dot_max > 255 || n >= (1ll << 31)) return false;
  if (encode_tiled() == nullptr) return false;
  KnnTcArgs a{};
  a.d = pitch;
  a.n = n;
  a.n_sq = n_sq;
  a.dot_max = dot_max;
  a.n_ranks = n_ranks;
  a.hist16 = n <= 65536 ? 1 : 0;
  int dev = 0, max_smem = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
    return false;
  // at least three B stages: with fewer the tensor pipe waits for every TMA round trip
  return knn_tc_fixed_smem(a, pass) + 3 * static_cast<size_t>(kTileM) * kMaxKBlock <= static_cast<size_t>(max_smem);
}

int knn_tc_launch(const KnnTcArgs& a, int pass, int* error_flag, cudaStream_t s) {
  if (error_flag == nullptr) GG_CUDA_CHECK(cudaGetSymbolAddress(reinterpret_cast<void**>(&error_flag), g_knn_tc_error));
  GG_REQUIRE(pass == 1 || pass == 2, GG_ERR_ARG, "pass must be 1 or 2");
  GG_REQUIRE(a.d % kMaxKBlock == 0 && a.d >= kMaxKBlock && a.d <= kMaxD, GG_ERR_ARG,
             "tcgen05 kNN kernel needs a row pitch that is a multiple of 128 up to 1024 bytes");
  GG_REQUIRE(a.row_begin % kTileM == 0 && a.row_begin < a.row_end && a.row_end <= a.n, GG_ERR_ARG, "bad row shard");
  EncodeTiledFn encode = encode_tiled();
  GG_REQUIRE(encode != nullptr, GG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  Geom g;
  g.kb_bytes = kMaxKBlock;
  g.n_kblocks = a.d / g.kb_bytes;
  g.stage_bytes = kTileM * g.kb_bytes;
  g.mma_per_kb = g.kb_bytes / kUmmaK;
  g.desc_hi = static_cast<uint32_t>((8 * g.kb_bytes) >> 4) | (1u << 14) | (2u << 29);
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(a.d), static_cast<cuuint64_t>(a.n)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(a.d)};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(g.kb_bytes), kTileM};
  const cuuint32_t elem[2] = {1, 1};
  const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(a.counts), dims, strides, box,
                            elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    gg::set_error("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return GG_ERR_CUDA;
  }
  const size_t fixed = knn_tc_fixed_smem(a, pass);
  int dev = 0, max_smem = 0;
  GG_CUDA_CHECK(cudaGetDevice(&dev));
  GG_CUDA_CHECK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  GG_REQUIRE(fixed + 2 * static_cast<size_t>(g.stage_bytes) <= static_cast<size_t>(max_smem), GG_ERR_ARG,
             "not enough shared memory for the B ring");
  int n_bstages = static_cast<int>((static_cast<size_t>(max_smem) - fixed) / g.stage_bytes);
  if (n_bstages > kMaxBStages) n_bstages = kMaxBStages;
  const size_t smem_bytes = fixed + static_cast<size_t>(n_bstages) * g.stage_bytes;
  auto kernel = pass == 1 ? knn_tc_pair_kernel<1> : knn_tc_pair_kernel<2>;
  GG_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes)));
  const long long n_sp = (a.row_end - a.row_begin + 2 * kTileM - 1) / (2 * kTileM);
  const long long pairs = gg::sm_count() / 2;
  const unsigned grid = 2u * static_cast<unsigned>(n_sp < pairs ? n_sp : pairs);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  GG_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, tmap, a, n_bstages, error_flag, g));
  return GG_OK;
}


// ops issued by one launch: pairs x iters x n_kblocks x 4 MMAs x 2 x 256 x 256 x 32
int tensor_probe(int iters, int n_kblocks, int* error_flag, double* h_ops, cudaStream_t s) {
  GG_REQUIRE(iters >= 1 && n_kblocks >= 1 && n_kblocks <= 8 && error_flag && h_ops, GG_ERR_ARG, "bad probe arguments");
  const size_t smem_bytes = static_cast<size_t>(n_kblocks + 1) * kTileM * 128 + sizeof(Barriers) + 1024;
  GG_CUDA_CHECK(cudaFuncSetAttribute(tensor_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem_bytes)));
  const unsigned pairs = static_cast<unsigned>(gg::sm_count() / 2);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const uint32_t desc_hi = static_cast<uint32_t>((8 * 128) >> 4) | (1u << 14) | (2u << 29);
  GG_CUDA_CHECK(cudaLaunchKernelEx(&cfg, tensor_probe_kernel, iters, n_kblocks, desc_hi, error_flag));
  *h_ops = static_cast<double>(pairs) * iters * n_kblocks * 4.0 * 2.0 * 256.0 * 256.0 * 32.0;
  return GG_OK;
}

}  // namespace ggkf
