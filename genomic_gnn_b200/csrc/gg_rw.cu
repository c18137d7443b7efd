// f3: biased random walks on the De Bruijn graph and their skip-gram windows -> positive node pairs.
//
// Replaces sample_biased_random_walks (reference gnn_common/gnn_utils.py:514-600): for every start node and walk
// round, a walk of up to walk_length steps over the directed, weighted graph -- the first step draws a successor in
// proportion to the edge weights, later steps use the node2vec bias (weight / p back to the previous node, weight to a
// successor that has an edge to the previous node, weight / q otherwise) -- then for every position i of the walk a
// window half-width shrink ~ U{1..window_size} and the pairs (walk[i], walk[j]), j in [i - shrink, i + shrink], j != i.
// The reference draws from Python's Mersenne Twister, so parity is distributional (transition frequencies, window
// law, structure of the pair list), not bit-wise; here every walk owns a Philox stream keyed by (seed, walk index), so a
// run is reproducible and the counting pass and the writing pass replay identical walks.
//
// One thread per walk.  De Bruijn rows have at most 4 successors (plus rare bridges), so the row scans are short.
#include <curand_kernel.h>

#include "gg_common.cuh"

namespace {

constexpr int kMaxWalk = 32;  // walk_length <= 32: the walk lives in registers / local memory

struct RwArgs {
  const int64_t* row_ptr;
  const int64_t* col;
  const float* w;
  int64_t n;
  const int64_t* start;
  int64_t n_start;
  int num_walks, walk_length, window;
  double inv_p, inv_q;
  unsigned long long seed;
  const int64_t* offsets;  // write pass: first pair slot of every walk
  int64_t* counts;         // count pass: pairs of every walk
  int64_t* src;
  int64_t* dst;
  int64_t capacity;
  int64_t* walks;          // optional [n_walks][walk_length + 1], -1 padded
};

__device__ __forceinline__ bool has_edge(const RwArgs& a, int64_t u, int64_t v) {
  for (int64_t e = a.row_ptr[u]; e < a.row_ptr[u + 1]; ++e)
    if (a.col[e] == v) return true;
  return false;
}

template <bool WRITE>
__global__ void __launch_bounds__(128) rw_kernel(RwArgs a) {
  const int64_t n_walks = static_cast<int64_t>(a.num_walks) * a.n_start;
  const int64_t wid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (wid >= n_walks) return;
  curandStatePhilox4_32_10_t rng;
  curand_init(a.seed, static_cast<unsigned long long>(wid), 0ull, &rng);
  int64_t walk[kMaxWalk + 1];
  int len = 1;
  walk[0] = a.start[wid % a.n_start];  // round r visits the start nodes in the same (shuffled) order
  int64_t prev = -1;
  for (int step = 0; step < a.walk_length; ++step) {
    const int64_t cur = walk[len - 1];
    const int64_t e0 = a.row_ptr[cur], e1 = a.row_ptr[cur + 1];
    if (e1 == e0) break;  // dead end: the walk stops (gnn_utils.py:585-586)
    double total = 0.0;
    for (int64_t e = e0; e < e1; ++e) {
      double x = static_cast<double>(a.w[e]);
      if (prev >= 0) {
        const int64_t nb = a.col[e];
        if (nb == prev) x *= a.inv_p;
        else if (!has_edge(a, nb, prev)) x *= a.inv_q;
      }
      total += x;
    }
    // random.choices: first successor whose cumulative weight exceeds u * total
    const double u = curand_uniform_double(&rng);  // (0, 1]
    const double target = (1.0 - u) * total;       // [0, total)
    double cum = 0.0;
    int64_t next = a.col[e1 - 1];
    for (int64_t e = e0; e < e1; ++e) {
      double x = static_cast<double>(a.w[e]);
      if (prev >= 0) {
        const int64_t nb = a.col[e];
        if (nb == prev) x *= a.inv_p;
        else if (!has_edge(a, nb, prev)) x *= a.inv_q;
      }
      cum += x;
      if (cum > target) {
        next = a.col[e];
        break;
      }
    }
    walk[len++] = next;
    prev = cur;
  }
  if (WRITE && a.walks) {
    for (int i = 0; i <= a.walk_length; ++i) a.walks[wid * (a.walk_length + 1) + i] = i < len ? walk[i] : -1;
  }
  int64_t pos = WRITE ? a.offsets[wid] : 0;
  int64_t count = 0;
  for (int i = 0; i < len; ++i) {
    const int shrink = 1 + static_cast<int>(curand(&rng) % static_cast<unsigned>(a.window));  // randint(1, window)
    const int lo = i - shrink < 0 ? 0 : i - shrink, hi = i + shrink >= len ? len - 1 : i + shrink;
    for (int j = lo; j <= hi; ++j) {
      if (j == i) continue;
      if (WRITE) {
        if (pos < a.capacity) {
          a.src[pos] = walk[i];
          a.dst[pos] = walk[j];
        }
        ++pos;
      } else {
        ++count;
      }
    }
  }
  if (!WRITE) a.counts[wid] = count;
}

}  // namespace

/* mode 0: pair_counts[w] = pairs of walk w (w = round * n_start + start index);
 * mode 1: pair_offsets[w] = first output slot of walk w (exclusive prefix sum of the counts), pairs are written to
 *         pairs_src / pairs_dst, and -- when walks_out is not null -- the walks themselves ([n_walks][walk_length+1]). */
extern "C" int gg_rw_pairs(const int64_t* row_ptr, const int64_t* col, const float* weight, int64_t n,
                           const int64_t* start_nodes, int64_t n_start, int num_walks, int walk_length,
                           int window_size, double p, double q, uint64_t seed, int mode, int64_t* pair_counts,
                           const int64_t* pair_offsets, int64_t* pairs_src, int64_t* pairs_dst, int64_t capacity,
                           int64_t* walks_out, gg_stream_t st) {
  GG_REQUIRE(n >= 0 && n_start >= 0 && num_walks >= 0, GG_ERR_ARG, "negative size");
  GG_REQUIRE(walk_length >= 0 && walk_length <= kMaxWalk, GG_ERR_ARG, "walk_length outside [0, 32]");
  GG_REQUIRE(window_size >= 1, GG_ERR_ARG, "window_size must be >= 1");
  GG_REQUIRE(p > 0.0 && q > 0.0, GG_ERR_ARG, "p and q must be positive");
  GG_REQUIRE(mode == 0 || mode == 1, GG_ERR_ARG, "mode must be 0 (count) or 1 (write)");
  const int64_t n_walks = static_cast<int64_t>(num_walks) * n_start;
  if (n_walks == 0) return GG_OK;
  GG_REQUIRE(row_ptr && col && weight && start_nodes, GG_ERR_ARG, "null graph");
  GG_REQUIRE(mode == 0 ? pair_counts != nullptr : (pair_offsets && (capacity == 0 || (pairs_src && pairs_dst))),
             GG_ERR_ARG, "null output");
  RwArgs a{row_ptr, col, weight, n, start_nodes, n_start, num_walks, walk_length, window_size, 1.0 / p, 1.0 / q,
           static_cast<unsigned long long>(seed), pair_offsets, pair_counts, pairs_src, pairs_dst, capacity, walks_out};
  const unsigned grid = static_cast<unsigned>((n_walks + 127) / 128);
  if (mode == 0) rw_kernel<false><<<grid, 128, 0, gg::as_stream(st)>>>(a);
  else rw_kernel<true><<<grid, 128, 0, gg::as_stream(st)>>>(a);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}
