// K2: De Bruijn graph assembly on the device.
//
// Replaces build_deBruijn_graph (reference src/utils.py:87-165) followed by
// torch_geometric.utils.from_networkx (gnn_tasks/utils.py:37): out-count
// normalisation in float64, node numbering by first occurrence, edges grouped
// by source in node order and ordered inside a source by first occurrence of
// the pair.  The edge "universe" is every (k+1)-mer bin plus the de-duplicated
// bridge records; absent entries get an all-ones sort key and fall off the end,
// so no host round trip is needed to learn N or E.  The two orderings are
// radix sorts (CUB, header-only part of the CUDA toolkit) over at most
// 4^(k+1) keys -- latency-bound utility work, not a roofline kernel.
#include <cub/cub.cuh>

#include "gg_common.cuh"

namespace {

constexpr uint64_t kInf = GG_NO_POS;

struct Universe {
  const uint32_t* hist;
  const uint64_t* first_pos;
  int64_t bins;
  // de-duplicated bridges: slot i valid iff b_cnt[i] > 0
  const uint64_t* b_key;  // (u << 26) | v, sorted
  const uint64_t* b_pos;
  const uint32_t* b_cnt;
  int64_t n_bridges;
  int k;
  __device__ __forceinline__ int64_t size() const { return bins + n_bridges; }
  __device__ __forceinline__ bool get(int64_t e, uint32_t& u, uint32_t& v, uint64_t& cnt, uint64_t& pos) const {
    if (e < bins) {
      const uint32_t c = hist[e];
      if (c == 0) return false;
      u = static_cast<uint32_t>(e >> 2);
      v = static_cast<uint32_t>(e & ((1ll << (2 * k)) - 1));
      cnt = c;
      pos = first_pos[e];
      return true;
    }
    const int64_t i = e - bins;
    if (b_cnt[i] == 0) return false;
    u = static_cast<uint32_t>(b_key[i] >> 26);
    v = static_cast<uint32_t>(b_key[i] & ((1ull << 26) - 1));
    cnt = b_cnt[i];
    pos = b_pos[i];
    return true;
  }
};

#define GRID_STRIDE(i, n) \
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < (n); i += static_cast<int64_t>(gridDim.x) * blockDim.x)

__global__ void bridge_keys_kernel(const gg_bridge_t* br, int64_t n, uint64_t* key, uint64_t* pos) {
  GRID_STRIDE(i, n) {
    key[i] = (static_cast<uint64_t>(br[i].u) << 26) | br[i].v;
    pos[i] = br[i].pos;
  }
}

// run heads of the sorted bridge keys absorb their run: count and earliest position
__global__ void bridge_dedupe_kernel(const uint64_t* key, const uint64_t* pos, int64_t n, uint32_t* cnt_out,
                                     uint64_t* pos_out) {
  GRID_STRIDE(i, n) {
    if (i > 0 && key[i - 1] == key[i]) {
      cnt_out[i] = 0;
      continue;
    }
    uint32_t c = 0;
    uint64_t p = kInf;
    for (int64_t j = i; j < n && key[j] == key[i]; ++j) {
      ++c;
      p = min(p, pos[j]);
    }
    cnt_out[i] = c;
    pos_out[i] = p;
  }
}

__global__ void out_count_kernel(Universe un, unsigned long long* out_count) {
  GRID_STRIDE(e, un.size()) {
    uint32_t u, v;
    uint64_t cnt, pos;
    if (un.get(e, u, v, cnt, pos)) atomicAdd(out_count + u, static_cast<unsigned long long>(cnt));
  }
}

__device__ __forceinline__ double edge_weight(uint64_t cnt, unsigned long long tot, int normalise) {
  return normalise ? static_cast<double>(cnt) / static_cast<double>(tot) : static_cast<double>(cnt);
}

__global__ void node_key_kernel(Universe un, const unsigned long long* out_count, int normalise, double threshold,
                                unsigned long long* node_key, uint8_t* kept) {
  GRID_STRIDE(e, un.size()) {
    uint32_t u, v;
    uint64_t cnt, pos;
    bool keep = un.get(e, u, v, cnt, pos);
    if (keep && threshold >= 0.0) keep = edge_weight(cnt, out_count[u], normalise) > threshold;
    kept[e] = keep;
    if (keep) {
      atomicMin(node_key + u, static_cast<unsigned long long>(2 * pos));
      atomicMin(node_key + v, static_cast<unsigned long long>(2 * pos + 1));
    }
  }
}

__global__ void node_sort_input_kernel(const unsigned long long* node_key, int64_t n_codes, int all_kmers,
                                       uint64_t* key, uint32_t* val) {
  GRID_STRIDE(c, n_codes) {
    key[c] = all_kmers ? static_cast<uint64_t>(c) : node_key[c];
    val[c] = static_cast<uint32_t>(c);
  }
}

// `deg` (nullable): kept out-edges of node i among the four (k+1)-mer bins of its code -- the input of the sort-free
// edge ordering (edge_write_grouped_kernel)
__global__ void node_rank_kernel(const uint64_t* key, const uint32_t* val, int64_t n_codes, int64_t* node_codes,
                                 int32_t* code_to_node, int64_t* out_n, const uint8_t* kept, int32_t* deg) {
  GRID_STRIDE(i, n_codes) {
    const bool valid = key[i] != kInf;
    code_to_node[val[i]] = valid ? static_cast<int32_t>(i) : -1;
    if (valid) {
      node_codes[i] = val[i];
      if (i + 1 == n_codes || key[i + 1] == kInf) *out_n = i + 1;
    }
    if (deg) deg[i] = valid ? __popc(reinterpret_cast<const uint32_t*>(kept)[val[i]]) : 0;  // kept[] holds 0 / 1 bytes
  }
}

__global__ void edge_sort_input_kernel(Universe un, const uint8_t* kept, const int32_t* code_to_node, int pos_bits,
                                       uint64_t* key, uint32_t* val) {
  GRID_STRIDE(e, un.size()) {
    uint64_t kk = kInf;
    if (kept[e]) {
      uint32_t u, v;
      uint64_t cnt, pos;
      un.get(e, u, v, cnt, pos);
      kk = (static_cast<uint64_t>(code_to_node[u]) << pos_bits) | pos;
    }
    key[e] = kk;
    val[e] = static_cast<uint32_t>(e);
  }
}

__global__ void edge_write_kernel(Universe un, const uint64_t* key, const uint32_t* val,
                                  const unsigned long long* out_count, const int32_t* code_to_node, int normalise,
                                  int64_t* src, int64_t* dst, double* w64, float* w32, int64_t* out_e) {
  const int64_t n = un.size();
  GRID_STRIDE(i, n) {
    if (key[i] == kInf) continue;
    uint32_t u, v;
    uint64_t cnt, pos;
    un.get(val[i], u, v, cnt, pos);
    const double w = edge_weight(cnt, out_count[u], normalise);
    src[i] = code_to_node[u];
    dst[i] = code_to_node[v];
    if (w64) w64[i] = w;
    if (w32) w32[i] = static_cast<float>(w);
    if (i + 1 == n || key[i + 1] == kInf) *out_e = i + 1;
  }
}

// Edge order without a sort (streams without bridge records): an edge is a kept (k+1)-mer bin 4u + b, so a source has
// at most four of them.  PyG order = sources in node order, a source's edges by first occurrence of the pair: node i
// writes its <= 4 edges, ordered by position with a 4-element network, at the exclusive prefix sum of the degrees.
__global__ void edge_write_grouped_kernel(Universe un, const uint8_t* kept, const int64_t* node_codes,
                                          const int64_t* out_n, const int32_t* offset,
                                          const unsigned long long* out_count, const int32_t* code_to_node,
                                          int normalise, int64_t n_codes, int64_t* src, int64_t* dst, double* w64,
                                          float* w32, int64_t* out_e) {
  const int64_t n_nodes = *out_n;
  GRID_STRIDE(i, n_codes) {
    if (i >= n_nodes) continue;
    const uint32_t u = static_cast<uint32_t>(node_codes[i]);
    uint64_t pos[4];
    int b[4] = {0, 1, 2, 3};
#pragma unroll
    for (int x = 0; x < 4; ++x) pos[x] = kept[4ll * u + x] ? un.first_pos[4ll * u + x] : kInf;
#define GG_CSWAP(x, y)                        \
  if (pos[y] < pos[x]) {                      \
    const uint64_t tp = pos[x];               \
    pos[x] = pos[y];                          \
    pos[y] = tp;                              \
    const int tb = b[x];                      \
    b[x] = b[y];                              \
    b[y] = tb;                                \
  }
    GG_CSWAP(0, 1) GG_CSWAP(2, 3) GG_CSWAP(0, 2) GG_CSWAP(1, 3) GG_CSWAP(1, 2)
#undef GG_CSWAP
    const int64_t at = offset[i];
    const unsigned long long tot = out_count[u];
    int n_out = 0;
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      if (pos[x] == kInf) continue;
      const int64_t e = 4ll * u + b[x];
      const double w = edge_weight(un.hist[e], tot, normalise);
      src[at + n_out] = i;
      dst[at + n_out] = code_to_node[e & ((1ll << (2 * un.k)) - 1)];
      if (w64) w64[at + n_out] = w;
      if (w32) w32[at + n_out] = static_cast<float>(w);
      ++n_out;
    }
    if (i + 1 == n_nodes) *out_e = at + n_out;
  }
}

size_t degree_scan_temp_bytes(int64_t n) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, static_cast<const int32_t*>(nullptr), static_cast<int32_t*>(nullptr), n);
  return bytes;
}

inline int grid_for(int64_t n) {
  int64_t g = (n + 255) / 256;
  const int64_t cap = static_cast<int64_t>(gg::sm_count()) * 8;
  if (g > cap) g = cap;
  return g < 1 ? 1 : static_cast<int>(g);
}

size_t sort_temp_bytes(int64_t n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, static_cast<const uint64_t*>(nullptr),
                                  static_cast<uint64_t*>(nullptr), static_cast<const uint32_t*>(nullptr),
                                  static_cast<uint32_t*>(nullptr), static_cast<int64_t>(n));
  size_t bytes2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes2, static_cast<const uint64_t*>(nullptr),
                                  static_cast<uint64_t*>(nullptr), static_cast<const uint64_t*>(nullptr),
                                  static_cast<uint64_t*>(nullptr), static_cast<int64_t>(n));
  return bytes > bytes2 ? bytes : bytes2;
}

struct Layout {
  unsigned long long* out_count;
  unsigned long long* node_key;
  uint8_t* kept;
  uint64_t *key_in, *key_out;
  uint32_t *val_in, *val_out;
  uint64_t *b_key_in, *b_key, *b_pos_in, *b_pos_sorted, *b_pos;
  uint32_t* b_cnt;
  int32_t *deg, *deg_offset;
  void* cub_temp;
  size_t cub_bytes;
  size_t total;
};

Layout carve(void* ws, int k, int64_t n_bridges) {
  const int64_t n_codes = 1ll << (2 * k);
  const int64_t uni = (1ll << (2 * (k + 1))) + n_bridges;
  gg::Carver c(ws);
  Layout l;
  l.out_count = c.take<unsigned long long>(n_codes);
  l.node_key = c.take<unsigned long long>(n_codes);
  l.kept = c.take<uint8_t>(uni);
  l.key_in = c.take<uint64_t>(uni);
  l.key_out = c.take<uint64_t>(uni);
  l.val_in = c.take<uint32_t>(uni);
  l.val_out = c.take<uint32_t>(uni);
  const int64_t nb = n_bridges > 0 ? n_bridges : 1;
  l.b_key_in = c.take<uint64_t>(nb);
  l.b_key = c.take<uint64_t>(nb);
  l.b_pos_in = c.take<uint64_t>(nb);
  l.b_pos_sorted = c.take<uint64_t>(nb);
  l.b_pos = c.take<uint64_t>(nb);
  l.b_cnt = c.take<uint32_t>(nb);
  l.deg = c.take<int32_t>(n_codes);
  l.deg_offset = c.take<int32_t>(n_codes);
  {
    const size_t a = sort_temp_bytes(uni), b = degree_scan_temp_bytes(n_codes);
    l.cub_bytes = a > b ? a : b;
  }
  l.cub_temp = c.take<char>(l.cub_bytes);
  l.total = c.used();
  return l;
}

}  // namespace

extern "C" size_t gg_db_workspace_bytes(int k, int64_t n_bridges) {
  if (k < 1 || k > GG_MAX_K || n_bridges < 0) return 0;
  return carve(nullptr, k, n_bridges).total;
}

extern "C" int gg_db_build(const uint32_t* hist, const uint64_t* first_pos, const gg_bridge_t* bridges,
                           int64_t n_bridges, int k, int normalise, int create_all_kmers,
                           double edge_weight_threshold, int64_t* node_codes, int32_t* code_to_node,
                           int64_t* edge_src, int64_t* edge_dst, double* weight64, float* weight32,
                           int64_t edge_capacity, int64_t* out_ne, void* workspace, size_t workspace_bytes, int pos_bits,
                           gg_stream_t st) {
  GG_REQUIRE(k >= 1 && k <= GG_MAX_K, GG_ERR_K, "k outside [1, GG_MAX_K]");
  GG_REQUIRE(pos_bits >= 1 && pos_bits <= 40, GG_ERR_ARG, "pos_bits outside [1, 40]");
  // the edge sort key is (node rank << pos_bits) | first position, plus one bit for the all-ones "dropped" key
  GG_REQUIRE(pos_bits + 2 * k + 1 <= 64, GG_ERR_ARG, "pos_bits + 2k + 1 must fit the 64-bit edge sort key");
  GG_REQUIRE(hist && first_pos && node_codes && code_to_node && edge_src && edge_dst && out_ne && workspace,
             GG_ERR_ARG, "null buffer");
  GG_REQUIRE(n_bridges >= 0 && (n_bridges == 0 || bridges), GG_ERR_ARG, "bridge list");
  const int64_t n_codes = 1ll << (2 * k);
  const int64_t bins = n_codes * 4;
  const int64_t uni = bins + n_bridges;
  GG_REQUIRE(edge_capacity >= uni, GG_ERR_ARG, "edge_capacity must be >= 4^(k+1) + n_bridges");
  Layout l = carve(workspace, k, n_bridges);
  GG_REQUIRE(workspace_bytes >= l.total, GG_ERR_WORKSPACE, "workspace too small");
  cudaStream_t s = gg::as_stream(st);

  GG_CUDA_CHECK(cudaMemsetAsync(l.out_count, 0, n_codes * sizeof(unsigned long long), s));
  GG_CUDA_CHECK(cudaMemsetAsync(l.node_key, 0xFF, n_codes * sizeof(unsigned long long), s));
  GG_CUDA_CHECK(cudaMemsetAsync(out_ne, 0, 2 * sizeof(int64_t), s));

  if (n_bridges > 0) {
    bridge_keys_kernel<<<grid_for(n_bridges), 256, 0, s>>>(bridges, n_bridges, l.b_key_in, l.b_pos_in);
    size_t tb = l.cub_bytes;
    GG_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(l.cub_temp, tb, l.b_key_in, l.b_key, l.b_pos_in, l.b_pos_sorted,
                                                  n_bridges, 0, 52, s));
    bridge_dedupe_kernel<<<grid_for(n_bridges), 256, 0, s>>>(l.b_key, l.b_pos_sorted, n_bridges, l.b_cnt, l.b_pos);
  }
  Universe un{hist, first_pos, bins, l.b_key, l.b_pos, l.b_cnt, n_bridges, k};

  out_count_kernel<<<grid_for(uni), 256, 0, s>>>(un, l.out_count);
  node_key_kernel<<<grid_for(uni), 256, 0, s>>>(un, l.out_count, normalise, edge_weight_threshold, l.node_key, l.kept);

  node_sort_input_kernel<<<grid_for(n_codes), 256, 0, s>>>(l.node_key, n_codes, create_all_kmers, l.key_in, l.val_in);
  {
    size_t tb = l.cub_bytes;
    // node keys are 2*pos(+1) < 2^(pos_bits + 1) (or the code < 4^k); the all-ones "absent" key still sorts last on
    // one more bit than any real key has
    const int node_bits = create_all_kmers ? 2 * k + 1 : pos_bits + 2;
    GG_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(l.cub_temp, tb, l.key_in, l.key_out, l.val_in, l.val_out, n_codes,
                                                  0, node_bits, s));
  }
  const bool grouped = n_bridges == 0;  // every edge is a (k+1)-mer bin: order them per source, no sort
  node_rank_kernel<<<grid_for(n_codes), 256, 0, s>>>(l.key_out, l.val_out, n_codes, node_codes, code_to_node, out_ne,
                                                     l.kept, grouped ? l.deg : nullptr);
  if (grouped) {
    size_t tb = l.cub_bytes;
    GG_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(l.cub_temp, tb, l.deg, l.deg_offset, static_cast<int>(n_codes), s));
    edge_write_grouped_kernel<<<grid_for(n_codes), 256, 0, s>>>(un, l.kept, node_codes, out_ne, l.deg_offset, l.out_count,
                                                                code_to_node, normalise, n_codes, edge_src, edge_dst,
                                                                weight64, weight32, out_ne + 1);
    GG_CUDA_CHECK(cudaGetLastError());
    return GG_OK;
  }

  edge_sort_input_kernel<<<grid_for(uni), 256, 0, s>>>(un, l.kept, code_to_node, pos_bits, l.key_in, l.val_in);
  {
    size_t tb = l.cub_bytes;
    // key = node rank (< 4^k) above pos_bits of first position; one more bit keeps the all-ones key of dropped edges last
    GG_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(l.cub_temp, tb, l.key_in, l.key_out, l.val_in, l.val_out, uni, 0,
                                                  pos_bits + 2 * k + 1, s));
  }
  edge_write_kernel<<<grid_for(uni), 256, 0, s>>>(un, l.key_out, l.val_out, l.out_count, code_to_node, normalise,
                                                  edge_src, edge_dst, weight64, weight32, out_ne + 1);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}

// ---------------------------------------------------------------------------------------------
// Quantile pruning of a built graph (reference src/utils.py:151-160): threshold =
// np.quantile(weights, 1 - keep_top_k) with numpy's default linear interpolation, edges with
// weight < threshold are removed, nodes stay.  The interpolation repeats numpy's _lerp without
// fused multiply-adds so that the threshold is the same float64.
// ---------------------------------------------------------------------------------------------
namespace {

__global__ void quantile_kernel(const double* sorted, int64_t n, double q, double* out) {
  const double vi = static_cast<double>(n - 1) * q;
  int64_t lo = static_cast<int64_t>(floor(vi));
  if (lo < 0) lo = 0;
  if (lo > n - 1) lo = n - 1;
  const int64_t hi = lo + 1 > n - 1 ? n - 1 : lo + 1;
  const double t = __dsub_rn(vi, static_cast<double>(lo));
  const double a = sorted[lo], b = sorted[hi];
  const double d = __dsub_rn(b, a);
  *out = t >= 0.5 ? __dsub_rn(b, __dmul_rn(d, __dsub_rn(1.0, t))) : __dadd_rn(a, __dmul_rn(d, t));
}

__global__ void prune_flag_kernel(const double* w, int64_t n, const double* thr, uint8_t* keep) {
  GRID_STRIDE(i, n) keep[i] = !(w[i] < *thr);
}

struct KeepToU64 {
  __host__ __device__ unsigned long long operator()(uint8_t v) const { return v ? 1ull : 0ull; }
};

__global__ void prune_scatter_kernel(const uint8_t* keep, const unsigned long long* pos, int64_t n, const int64_t* src,
                                     const int64_t* dst, const double* w, int64_t* o_src, int64_t* o_dst, double* o_w64,
                                     float* o_w32, int64_t* out_count) {
  GRID_STRIDE(i, n) {
    if (keep[i]) {
      const unsigned long long p = pos[i];
      o_src[p] = src[i];
      o_dst[p] = dst[i];
      if (o_w64) o_w64[p] = w[i];
      if (o_w32) o_w32[p] = static_cast<float>(w[i]);
    }
    if (i == n - 1) *out_count = static_cast<int64_t>(pos[i] + (keep[i] ? 1 : 0));
  }
}

struct PruneLayout {
  double* sorted;
  double* thr;
  uint8_t* keep;
  unsigned long long* pos;
  void* temp;
  size_t temp_bytes, total;
};

PruneLayout prune_carve(void* ws, int64_t n) {
  gg::Carver c(ws);
  PruneLayout l;
  const int64_t m = n > 0 ? n : 1;
  l.sorted = c.take<double>(m);
  l.thr = c.take<double>(1);
  l.keep = c.take<uint8_t>(m);
  l.pos = c.take<unsigned long long>(m);
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, a, static_cast<const double*>(nullptr), static_cast<double*>(nullptr), m);
  cub::TransformInputIterator<unsigned long long, KeepToU64, const uint8_t*> it(nullptr, KeepToU64());
  cub::DeviceScan::ExclusiveSum(nullptr, b, it, static_cast<unsigned long long*>(nullptr), m);
  l.temp_bytes = a > b ? a : b;
  l.temp = c.take<char>(l.temp_bytes);
  l.total = c.used();
  return l;
}

}  // namespace

extern "C" size_t gg_db_prune_workspace_bytes(int64_t n_edges) {
  return n_edges < 0 ? 0 : prune_carve(nullptr, n_edges).total;
}

extern "C" int gg_db_prune_quantile(const int64_t* edge_src, const int64_t* edge_dst, const double* weight64,
                                    int64_t n_edges, double quantile, int64_t* out_src, int64_t* out_dst,
                                    double* out_w64, float* out_w32, int64_t* out_count, void* workspace,
                                    size_t workspace_bytes, gg_stream_t st) {
  GG_REQUIRE(n_edges >= 0 && out_count, GG_ERR_ARG, "bad size");
  cudaStream_t s = gg::as_stream(st);
  GG_CUDA_CHECK(cudaMemsetAsync(out_count, 0, sizeof(int64_t), s));
  if (n_edges == 0) return GG_OK;
  GG_REQUIRE(edge_src && edge_dst && weight64 && out_src && out_dst && workspace, GG_ERR_ARG, "null buffer");
  PruneLayout l = prune_carve(workspace, n_edges);
  GG_REQUIRE(workspace_bytes >= l.total, GG_ERR_WORKSPACE, "workspace too small");
  size_t tb = l.temp_bytes;
  GG_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(l.temp, tb, weight64, l.sorted, n_edges, 0, 64, s));
  quantile_kernel<<<1, 1, 0, s>>>(l.sorted, n_edges, quantile, l.thr);
  prune_flag_kernel<<<grid_for(n_edges), 256, 0, s>>>(weight64, n_edges, l.thr, l.keep);
  cub::TransformInputIterator<unsigned long long, KeepToU64, const uint8_t*> it(l.keep, KeepToU64());
  tb = l.temp_bytes;
  GG_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(l.temp, tb, it, l.pos, n_edges, s));
  prune_scatter_kernel<<<grid_for(n_edges), 256, 0, s>>>(l.keep, l.pos, n_edges, edge_src, edge_dst, weight64, out_src,
                                                         out_dst, out_w64, out_w32, out_count);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}
