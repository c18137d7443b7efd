// Per-row kNN KF edges on the tcgen05 CTA-pair Gram (gg_kf_mma.cu): arguments shared with gg_knn.cu.
#pragma once
#include "gg_common.cuh"

namespace ggkf {

constexpr int kKnnTcMaxSq = 32;  // squared-norm classes the per-row class masks hold

struct KnnTcArgs {
  const uint8_t* counts;
  int d;                       // row pitch in bytes: a multiple of 128 up to 1024
  long long n;
  long long row_begin, row_end;  // the shard's rows (row_begin a multiple of 128)
  const int32_t* sqnorm;
  const uint8_t* sqid;         // [sq_max + 1] squared norm -> class (0xFF: none)
  int sq_max;
  const uint16_t* rank_tab;    // [n_sq][dot_max + 1] -> rank (0xFFFF: never a neighbour)
  int n_sq, dot_max, n_ranks;
  uint32_t* hist;              // pass 1: [rows][n_ranks]
  int hist16;                  // pass 1: 16-bit shared-memory counters (n <= 65536)
  const int32_t* cut;          // pass 2: the plan of gg_knn_plan
  const int32_t* tie_left;
  const int64_t* row_off;
  // pass 2 output: column | dot << 32 | rank << 48.  A row's staging space is 3 x (its edge count) words at
  // 3 * row_off[row]; epilogue group g fills third g with the pairs it selects from the g-th third of the columns, in
  // column order, and third_cnt[3 * row + g] says how many.  Every third keeps at most tie_left members of the cut rank.
  uint64_t* recs;
  uint32_t* third_cnt;
};

// true when the tensor path can serve this shape (gg_knn.cu falls back to its mma.sync kernel otherwise)
bool knn_tc_eligible(int pitch, long long n, int n_sq, int dot_max, int n_ranks, int pass);
int knn_tc_launch(const KnnTcArgs& a, int pass, int* error_flag, cudaStream_t s);

}  // namespace ggkf
