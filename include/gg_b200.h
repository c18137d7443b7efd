/*
 * gg_b200.h -- C ABI of the B200-native k-mer graph-construction hot path.
 *
 * This is the drop-in boundary for the graph-construction path of
 * ratschlab/genomic-gnn.  The reference is pure Python and has no FFI of its
 * own; each entry point below states which reference function it replaces
 * (paths relative to the reference tree).  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with "h_";
 *   - every call is asynchronous on the caller's stream (gg_stream_t is a
 *     cudaStream_t) and allocates nothing: workspaces are caller-owned and
 *     sized by the *_workspace_bytes() queries;
 *   - return value: GG_OK or a negative GG_ERR_* (argument / launch errors
 *     only -- data-dependent results such as "bad byte seen" or "capacity
 *     exceeded" are written to the device status words documented per call,
 *     because the host does not synchronise inside these calls);
 *   - no global state: calls are re-entrant, one stream per call.
 *
 * k-mer code: base-4 value of the string with A,C,G,T = 0,1,2,3 and the first
 * base most significant (reference: itertools.product(ALPHABET, repeat=k),
 * src/utils.py:125 and gnn_utils.py:458-459).
 */
#ifndef GG_B200_H
#define GG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* gg_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define GG_API __attribute__((visibility("default")))
#else
#define GG_API
#endif

#define GG_ABI_VERSION 3
#define GG_OK 0
#define GG_ERR_ARG (-1)      /* null pointer, negative size, misaligned buffer */
#define GG_ERR_K (-2)        /* k / small_k outside the supported range */
#define GG_ERR_CUDA (-5)     /* a CUDA runtime call failed; see gg_last_error() */
#define GG_ERR_WORKSPACE (-6)/* workspace smaller than *_workspace_bytes() */

#define GG_SEP 10            /* '\n': hard read boundary in a separator-delimited stream */
#define GG_MAX_K 13          /* dense (k+1)-mer histogram: 4^(k+1) uint32 bins */
#define GG_KF_BLOCK 1024     /* reference chunk_size (gnn_utils.py:241); fixes edge order */
#define GG_NO_POS 0xFFFFFFFFFFFFFFFFull

GG_API int gg_abi_version(void);
GG_API const char* gg_last_error(void); /* thread-local, valid until the next failing call */
GG_API const char* gg_build_id(void);   /* digest of the sources this library was compiled from (build.py) */

/* ------------------------------------------------------------------------
 * K1  counting.  Replaces seq_to_kmers + weighted_directed_edges
 *     (src/utils.py:15-31, 52-84) for stride 1, inlcudeUNK=False.
 *
 * Input: ASCII bytes.  fixed_len > 0: reads are back-to-back rows of exactly
 * fixed_len bytes; fixed_len == 0: every read is followed by one GG_SEP byte.
 * Bytes allowed: A C G T N (and GG_SEP when fixed_len == 0); anything else is
 * counted in status[GG_ST_BAD_BYTES] (the reference would fail later with a
 * KeyError at gnn_utils.py:471).  The buffer must be 16-byte aligned and
 * readable up to gg_stream_padded_bytes(n_bytes); padding content is ignored.
 *
 * A pair of consecutive clean k-mers is one (k+1)-mer: hist[(u << 2) | last
 * base of v] += 1.  k-mers holding an N are deleted before pairing
 * (utils.py:31), so the k-mer before an N run pairs with the first clean k-mer
 * after it: such "bridge" pairs that are not (k-1)-overlapping go to the
 * bridges list, overlapping ones are folded into hist.
 * ------------------------------------------------------------------------ */
typedef struct {
  uint32_t u, v; /* k-mer codes */
  uint64_t pos;  /* global stream offset of u's first base */
} gg_bridge_t;

enum {
  GG_ST_BAD_BYTES = 0, /* bytes outside the alphabet */
  GG_ST_BRIDGES = 1,   /* bridge records found (may exceed capacity: then re-run larger) */
  GG_ST_CURSOR = 2,    /* internal: chunk cursor of gg_first_occurrence */
  GG_ST_REMAINING = 3, /* internal: bins whose first occurrence is still unknown */
  GG_ST_NONZERO = 4,   /* non-zero (k+1)-mer bins */
  GG_ST_WORDS = 8
};

GG_API size_t gg_stream_padded_bytes(int64_t n_bytes);
GG_API size_t gg_hist_bins(int k); /* 4^(k+1) */

/* zero hist, set first_pos to GG_NO_POS, zero status[GG_ST_WORDS] */
GG_API int gg_count_reset(int k, uint32_t* hist, uint64_t* first_pos, int64_t* status, gg_stream_t st);

/* accumulate one shard of reads into hist (may be called repeatedly). pos_base is
 * the global stream offset of stream[0] (multi-GPU / multi-call first-occurrence keys). */
GG_API int gg_count_kp1mers(const uint8_t* stream, int64_t n_bytes, int64_t fixed_len, int k, int64_t pos_base,
                     uint32_t* hist, gg_bridge_t* bridges, int64_t bridge_capacity, int64_t* status,
                     gg_stream_t st);

/* Same contract and result as gg_count_kp1mers, for tables of up to 8 x 52,432 bins (k <= 8): the stream is packed
 * to 2-bit codes once, then counted in per-CTA shared-memory histograms over slices of the bin space (shared-memory
 * atomics sustain ~10x the rate of L2 atomics on B200).  gg_count_smem_workspace_bytes() returns 0 when k is too
 * large for this path. */
GG_API size_t gg_count_smem_workspace_bytes(int64_t n_bytes, int k);
GG_API int gg_count_kp1mers_smem(const uint8_t* stream, int64_t n_bytes, int64_t fixed_len, int k, int64_t pos_base,
                          uint32_t* hist, gg_bridge_t* bridges, int64_t bridge_capacity, int64_t* status,
                          void* workspace, size_t workspace_bytes, gg_stream_t st);

/* Same contract and result as gg_count_kp1mers, for tables of up to 8 x 32,768 bins (k <= 8) -- the fastest path:
 * the keys are partitioned by their top bits into <= 8 streams of 15-bit keys (one pass over the ASCII bytes), then
 * every stream is counted with full-warp shared-memory atomics; tables of <= 16,384 bins (k <= 6) are counted in
 * shared memory directly.  gg_count_part_workspace_bytes() returns 0 when k is too large for this path. */
GG_API size_t gg_count_part_workspace_bytes(int64_t n_bytes, int k);
GG_API int gg_count_kp1mers_part(const uint8_t* stream, int64_t n_bytes, int64_t fixed_len, int k, int64_t pos_base,
                          uint32_t* hist, gg_bridge_t* bridges, int64_t bridge_capacity, int64_t* status,
                          void* workspace, size_t workspace_bytes, gg_stream_t st);

/* first_pos[e] = min global offset at which (k+1)-mer e starts.  Scans the shard
 * in order and stops as soon as every non-zero bin of `hist` has been seen
 * (hist must already hold this shard's counts).  Defines dict insertion order
 * (utils.py:82) and, through it, node order (utils.py:131-143). */
GG_API int gg_first_occurrence(const uint8_t* stream, int64_t n_bytes, int64_t fixed_len, int k, int64_t pos_base,
                        const uint32_t* hist, uint64_t* first_pos, int64_t* status, gg_stream_t st);

/* Multi-GPU merge of per-rank K1 results after ONE all-gather: pieces holds `world` buffers of piece_stride 64-bit
 * words, each laid out [hist: 4^(k+1) uint32][first_pos: 4^(k+1) uint64][anything]; hist_out = sum of the histograms,
 * first_out = earliest first occurrence. */
GG_API int gg_count_merge(const uint64_t* pieces, int world, int64_t piece_stride, int k, uint32_t* hist_out,
                   uint64_t* first_out, gg_stream_t st);

/* ------------------------------------------------------------------------
 * K2  De Bruijn graph.  Replaces build_deBruijn_graph (src/utils.py:87-165)
 *     + torch_geometric.utils.from_networkx (gnn_tasks/utils.py:37).
 *
 * Nodes in first-occurrence order (or lexicographic when create_all_kmers);
 * edges grouped by source in node order, inside a source by first occurrence
 * of the pair; weight = count / out-count(source) in float64 (normalise) or
 * the raw count.  edge_weight_threshold < 0 disables the strict `>` filter
 * of utils.py:139-141.  out_ne[0] = N, out_ne[1] = E.  pos_bits: every first_pos / bridge position is < 2^pos_bits
 * (bounds the radix sorts; 40 is always safe).
 * ------------------------------------------------------------------------ */
GG_API size_t gg_db_workspace_bytes(int k, int64_t n_bridges);
GG_API int gg_db_build(const uint32_t* hist, const uint64_t* first_pos, const gg_bridge_t* bridges, int64_t n_bridges,
                int k, int normalise, int create_all_kmers, double edge_weight_threshold,
                int64_t* node_codes, int32_t* code_to_node, int64_t* edge_src, int64_t* edge_dst,
                double* weight64, float* weight32, int64_t edge_capacity, int64_t* out_ne,
                void* workspace, size_t workspace_bytes, int pos_bits, gg_stream_t st);

/* ------------------------------------------------------------------------
 * K3  node features.  Replaces node_embedding_initial(method="kmer_frequency")
 *     (gnn_utils.py:428-493).  counts[i, c] = occurrences of s-mer c in node i;
 *     sqnorm[i] = sum_c counts^2; col_present[c] != 0 iff some node has it
 *     (gnn_utils.py:474-481 drops the others).  gg_kf_features_f32 gathers the
 *     kept columns through a 256-entry value table (float32(c * (1/(k-s+1)))).
 * ------------------------------------------------------------------------ */
GG_API int gg_kf_counts(const int64_t* node_codes, int64_t n, int k, int s, uint8_t* counts, int32_t* sqnorm,
                 uint32_t* col_present, gg_stream_t st);
/* The same with the number of nodes still on the device (*n_dev <= n_max, e.g. out_ne[0] of gg_db_build queued just
 * before): buffers are sized for n_max rows, rows [*n_dev, n_max) of counts / sqnorm are zero.  Lets a build read N, E,
 * K1's status words and the column-presence flags in ONE device read. */
GG_API int gg_kf_counts_upto(const int64_t* node_codes, int64_t n_max, const int64_t* n_dev, int k, int s, uint8_t* counts,
                      int32_t* sqnorm, uint32_t* col_present, gg_stream_t st);
GG_API int gg_kf_features_f32(const uint8_t* counts, int64_t n, int d, const int32_t* col_keep, int d_keep,
                       const float* value_lut, float* x, gg_stream_t st);

/* ------------------------------------------------------------------------
 * K4-K6  KF similarity edges.  Replaces cosine_similarity_to_edge_index_weight
 *     (gnn_utils.py:240-322).
 *
 * cos(i,j) = dot / sqrt(sqnorm_i * sqnorm_j) with integer dot: the strict test
 * cos > threshold is evaluated exactly as dot >= dmin[sqnorm_i * sqnorm_j]
 * (dmin built on the host from the float64 threshold as a rational).
 *
 * gg_kf_emit scores every pair s < t with s in the row blocks
 * [iblock_begin, iblock_end) (GG_KF_BLOCK rows each; multi-GPU sharding unit),
 * writes one record per candidate (unordered) and, per (block pair, row), the
 * candidate counts of its eight 128-column tiles packed as the bytes of one
 * 64-bit word (rowcnt, gg_kf_rowcnt_elems() words).
 * gg_kf_finalize turns the counts into offsets and scatters the records into
 * the reference's emission order: block (I, J >= I), then row, then column.
 * ------------------------------------------------------------------------ */
typedef struct {
  uint32_t s, t;
  uint32_t rank; /* rank of t among the candidates of row s inside its 128-column tile */
  uint32_t dot;  /* only filled for rows wider than 64 bytes; narrower rows are re-multiplied by gg_kf_finalize */
} gg_kf_rec_t;

enum { GG_KF_IMPL_AUTO = 0, GG_KF_IMPL_DP4A = 1, GG_KF_IMPL_TCGEN05 = 2, GG_KF_IMPL_NAIVE = 3, GG_KF_IMPL_TCGEN05_PAIR = 4 };

typedef struct {
  int64_t n;            /* nodes */
  int32_t d;            /* columns of counts (4^s, multiple of 4) */
  int32_t iblock_begin; /* row-block shard */
  int32_t iblock_end;
  int32_t p_max;        /* dmin has p_max + 1 entries */
  int32_t fast_shift;   /* fast pre-filter: (dot*dot) << fast_shift > fast_scale[sqnorm_i] * sqnorm_j */
  int32_t impl;         /* GG_KF_IMPL_* */
  int32_t d_valid;      /* bytes of a row that can be non-zero (<= d; rows may be zero-padded to the 32-byte pitch
                           the tensor path needs); 0 means d */
  int32_t hist_m;       /* > 0: promise that every row of counts sums to hist_m (rows made by gg_kf_counts:
                           k - small_k + 1), which lets gg_kf_emit keep its score-class histogram in shared memory;
                           0: no promise (the histogram is still exact, only slower) */
} gg_kf_params_t;

/* kf_status (device int64[4]): [0] candidates found, exact (may exceed rec_capacity: re-run larger), [1] record slots
 * used when that is more than [0] (GG_KF_IMPL_TCGEN05_PAIR reserves slots in blocks and pads unused ones with records
 * whose s is 0xFFFFFFFF: pass max([0], [1]) as n_recs to gg_kf_place, which skips them, and size `ordered` by [0]; once
 * a reservation lands beyond rec_capacity nothing more is reserved or written), [2] pipeline error, [3] reserved.
 * class_hist (nullable; GG_KF_IMPL_TCGEN05_PAIR only): uint64[(sqnorm_max + 1) * (p_max + 1)], NOT zeroed by the call;
 * class_hist[dot * (p_max + 1) + sqnorm_s * sqnorm_t] += 1 for every candidate, whether or not its record fitted
 * rec_capacity -- the exact score classes of the global top-n (gnn_utils.py:311-317) BEFORE anything is emitted, so
 * that a binding top-n never needs the candidates materialised (k = 10, threshold 0: 5.5e11 of them). */
GG_API size_t gg_kf_rowcnt_elems(int64_t n, int iblock_begin, int iblock_end);
GG_API int gg_kf_emit(const gg_kf_params_t* h_params, const uint8_t* counts, const int32_t* sqnorm,
               const uint16_t* dmin, const uint32_t* fast_scale /* [max sqnorm + 1] */,
               gg_kf_rec_t* recs, int64_t rec_capacity, uint64_t* rowcnt, int64_t* kf_status, uint64_t* class_hist,
               gg_stream_t st);

/* gg_kf_finalize = gg_kf_place (records -> one packed 64-bit word per edge, in emission order: s | dot << 24 in the low
 * half, t in the high half) followed by gg_kf_expand (packed words -> edge_index / weights, coalesced).  Multi-GPU
 * builds call the halves separately: the packed words (8 bytes per edge instead of 20) are what is all-gathered, and
 * gg_kf_expand then reads the gathered buffer as up to 16 segments, segment r holding h_seg_size[r] words from word
 * h_seg_first[r] on. */
GG_API size_t gg_kf_place_workspace_bytes(int64_t n, int iblock_begin, int iblock_end);
GG_API int gg_kf_place(const gg_kf_params_t* h_params, const gg_kf_rec_t* recs, int64_t n_recs, const uint64_t* rowcnt,
                const uint8_t* counts, uint64_t* ordered, void* workspace, size_t workspace_bytes, gg_stream_t st);
GG_API int gg_kf_expand(const uint64_t* ordered, int n_segments, const int64_t* h_seg_first, const int64_t* h_seg_size,
                 const int32_t* sqnorm, int64_t* edge_src, int64_t* edge_dst, double* weight64, float* weight32,
                 uint32_t* dot_out, gg_stream_t st);
/* ------------------------------------------------------------------------
 * K4'  duplicate-row path for narrow rows (d_valid <= 16 columns, every count <= 15: small_k <= 2).
 *
 * Many nodes share one feature row there (k=8, small_k=2: 17,444 distinct rows among 65,536 nodes; k=10: 95,924
 * among 1,048,576), so the rows are grouped into classes, the CLASS pairs are scored once, and every class expands
 * its neighbour classes into a bitmap over node space from which its member rows write their edges directly in
 * emission order.  Produces exactly the packed words gg_kf_emit + gg_kf_place produce; n <= 2^20.
 *
 *   gg_kf_dedup_classes  members[n] = node ids grouped by class (ascending inside a class), class_start[n + 1],
 *                        dd_status (device int64[2]): [0] number of classes, [1] rows holding a count > 15
 *                        (then this path must not be used).  The host reads [0] to size `adj`.
 *   gg_kf_dedup_score    adj[c][w]: bit matrix of class pairs whose cosine exceeds the threshold (symmetric; the
 *                        diagonal is set because two members of one class have cosine 1), gg_kf_dedup_adj_words()
 *                        uint32 words; same exact test as gg_kf_emit (dmin / p_max of gg_kf_params_t).
 *   gg_kf_dedup_count    offsets[gg_kf_rowcnt_elems() + 1]: exclusive prefix sum of the candidates per
 *                        (block pair, row) of the shard; the LAST entry is the shard's number of edges.
 *   gg_kf_dedup_place    ordered[e] = s | dot << 24, t for every edge of the shard, in emission order
 *                        (the input of gg_kf_expand).
 * ------------------------------------------------------------------------ */
GG_API size_t gg_kf_dedup_classes_workspace_bytes(int64_t n);
GG_API int gg_kf_dedup_classes(const uint8_t* counts, int pitch, int d_valid, int64_t n, int32_t* members,
                        int32_t* class_start, int64_t* dd_status, void* workspace, size_t workspace_bytes,
                        gg_stream_t st);
GG_API size_t gg_kf_dedup_adj_words(int64_t n_classes);
/* Optional by-products of a scoring pass (host struct, nullable; at most one of class_hist / tie_out set):
 *   class_hist  uint64[(sqnorm_max + 1) * (p_max + 1)], not zeroed: for every passing class pair (A, B) with A in
 *               [hist_class_begin, hist_class_end), class_hist[dot * (p_max + 1) + sq_A * sq_B] += the node pairs
 *               s < t it stands for (|A| |B|; |A| (|A| - 1) / 2 for A = B) -- the score classes of the global top-n
 *               before anything is emitted; ranks split the class range and sum their tables.
 *   tie_out     uint64[2], not zeroed: node pairs of the class dot^2 / P == tie_num / tie_den (the cut-off class)
 *               whose smaller node is >= row_lo ([0]) and >= row_hi ([1]); [0] - [1] = the members of that class the
 *               row shard [row_lo, row_hi) emits.  With tie_num == tie_den == 0: tie_out[0] += the node pairs of ALL
 *               passing class pairs, i.e. the number of candidates of the whole matrix. */
typedef struct {
  uint64_t* class_hist;
  int64_t hist_class_begin, hist_class_end;
  uint64_t* tie_out;
  int64_t tie_num, tie_den;
  int64_t row_lo, row_hi;
} gg_kf_dedup_extra_t;
GG_API int gg_kf_dedup_score(const uint8_t* counts, int pitch, int d_valid, const int32_t* sqnorm, const int32_t* members,
                      const int32_t* class_start, int64_t n_classes, const uint16_t* dmin, int32_t p_max,
                      uint32_t* adj, const gg_kf_dedup_extra_t* h_extra, gg_stream_t st);
GG_API size_t gg_kf_dedup_count_workspace_bytes(int64_t n, int iblock_begin, int iblock_end);
GG_API int gg_kf_dedup_count(int64_t n, int iblock_begin, int iblock_end, const uint32_t* adj, const int32_t* members,
                      const int32_t* class_start, int64_t n_classes, uint64_t* offsets,
                      const int64_t* skip_flag /* nullable device word: non-zero = leave every count at zero (the caller
                                                  already knows it will cut by top-n and not use them) */,
                      void* workspace, size_t workspace_bytes, gg_stream_t st);
GG_API int gg_kf_dedup_place(int64_t n, int iblock_begin, int iblock_end, const uint32_t* adj, const int32_t* members,
                      const int32_t* class_start, int64_t n_classes, const uint64_t* offsets, const uint8_t* counts,
                      int pitch, int d_valid, uint64_t* ordered, int64_t capacity, void* workspace /* >= 256 bytes */,
                      size_t workspace_bytes, gg_stream_t st);

GG_API size_t gg_kf_finalize_workspace_bytes(int64_t n, int iblock_begin, int iblock_end, int64_t n_recs);
GG_API int gg_kf_finalize(const gg_kf_params_t* h_params, const gg_kf_rec_t* recs, int64_t n_recs, const uint64_t* rowcnt,
                   const int32_t* sqnorm, const uint8_t* counts /* the matrix gg_kf_emit scored; used when d <= 64 */,
                   int64_t* edge_src, int64_t* edge_dst, double* weight64, float* weight32,
                   uint32_t* dot_out /* nullable: dot per edge, for top-n */, void* workspace, size_t workspace_bytes,
                   gg_stream_t st);

/* K5  global top-n (gnn_utils.py:311-317).  Histogram of exact score classes
 * (dot, P = sqnorm_s * sqnorm_t) over finalised edges, then a stable select:
 * class_action[dot * (p_max+1) + P] = 1 keep, 2 cut-off class (keep the first
 * `tie_budget` in emission order), 0 drop.  out_count[0] = edges kept. */
GG_API int gg_kf_class_hist(const int64_t* edge_src, const int64_t* edge_dst, const uint32_t* dot, int64_t n_edges,
                     const int32_t* sqnorm, int32_t p_max, unsigned long long* class_hist, gg_stream_t st);
/* The same two steps on PACKED words (the output of gg_kf_place / gg_kf_dedup_place), for top-n selection without
 * ever expanding what is dropped, one chunk of the emission order at a time:
 *   gg_kf_class_hist_packed  class_hist[dot * (p_max + 1) + P] += 1 per word (not zeroed by the call)
 *   gg_kf_select_packed      keeps words of class_action 1 and, of the cut-off class (action 2), member number g
 *                            (GLOBAL emission order, g = cursors[0] + rank inside this call) iff
 *                            floor((g + 1) r / 2^64) > floor(g r / 2^64), r = tie_ratio = ceil(B 2^64 / T) for B of T
 *                            members to keep: the choice depends on g alone, so ranks and chunks need no
 *                            coordination beyond their starting g.  Kept words are expanded to out_src / out_dst /
 *                            out_w64 (nullable) / out_w32 (nullable) from index cursors[1] on (nothing is written at
 *                            or beyond out_capacity); the call then advances cursors[0] by the members of the cut-off
 *                            class it saw and cursors[1] by the edges it kept (device int64[2], caller-initialised).
 *                            At most 2^31 - 1 words per call. */
GG_API int gg_kf_class_hist_packed(const uint64_t* ordered, int64_t n_edges, const int32_t* sqnorm, int32_t p_max,
                            uint64_t* class_hist, gg_stream_t st);
GG_API size_t gg_kf_select_packed_workspace_bytes(int64_t n_edges);
GG_API int gg_kf_select_packed(const uint64_t* ordered, int64_t n_edges, const int32_t* sqnorm, int32_t p_max,
                        const uint8_t* class_action, uint64_t tie_ratio, int64_t* cursors, int64_t* out_src,
                        int64_t* out_dst, double* out_w64, float* out_w32, int64_t out_capacity, void* workspace,
                        size_t workspace_bytes, gg_stream_t st);
GG_API size_t gg_kf_select_workspace_bytes(int64_t n_edges);
GG_API int gg_kf_select(const int64_t* edge_src, const int64_t* edge_dst, const double* weight64, const uint32_t* dot,
                 int64_t n_edges, const int32_t* sqnorm, int32_t p_max, const uint8_t* class_action,
                 int64_t tie_budget, int64_t* out_src, int64_t* out_dst, double* out_w64, float* out_w32,
                 int64_t* out_count, void* workspace, size_t workspace_bytes, gg_stream_t st);

/* ------------------------------------------------------------------------
 * f2  reads -> k-mer index arrays.  Replaces word2index inside KmerSsgnn.transform
 *     (src/representations/kmer_ssgnn.py:148-158; seq_to_kmers(..., inlcudeUNK=True), src/utils.py:28-30):
 *     out[out_offsets[r] + w] = code_to_index[code of read r's window w] for windows at 0, stride, 2*stride, ...;
 *     windows holding an 'N' give unk_index.  status[0] = windows with a byte outside ACGTN (written as -1),
 *     status[1] = windows whose k-mer has code_to_index < 0 (the caller's "missing k-mers" rebuild trigger).
 * ------------------------------------------------------------------------ */
GG_API int gg_kmer_index(const uint8_t* stream, const int64_t* read_start, const int64_t* read_len, int64_t n_reads,
                  int k, int stride, const int32_t* code_to_index, int64_t unk_index, const int64_t* out_offsets,
                  int64_t* out, int64_t* status, gg_stream_t st);

/* ------------------------------------------------------------------------
 * f4  higher-stride De Bruijn graphs (kmer_ssgnn.py:215-226).
 *     gg_count_pairs_stride: weighted_directed_edges(..., stride = s > 1, inlcudeUNK=False): one (u, v, pos) record
 *     per pair of consecutive surviving k-mers at offsets 0, s, 2s, ... (status[GG_ST_BRIDGES] = records found);
 *     feed them to gg_db_build as `bridges` with an all-zero histogram.
 *     gg_db_prune_quantile: keep_top_k of build_deBruijn_graph (utils.py:151-160): drops edges with weight <
 *     np.quantile(weights, quantile) (linear interpolation, numpy's _lerp), order kept, out_count[0] = edges left.
 * ------------------------------------------------------------------------ */
GG_API int gg_count_pairs_stride(const uint8_t* stream, const int64_t* read_start, const int64_t* read_len,
                          int64_t n_reads, int k, int stride, int64_t pos_base, gg_bridge_t* records, int64_t capacity,
                          int64_t* status, gg_stream_t st);
GG_API size_t gg_db_prune_workspace_bytes(int64_t n_edges);
GG_API int gg_db_prune_quantile(const int64_t* edge_src, const int64_t* edge_dst, const double* weight64, int64_t n_edges,
                         double quantile, int64_t* out_src, int64_t* out_dst, double* out_w64, float* out_w32,
                         int64_t* out_count, void* workspace, size_t workspace_bytes, gg_stream_t st);

/* Host-side half of gg_kf_features_f32 for the host-to-host build: h_x[i] = h_value_lut[h_counts[i]] on n_threads host
 * threads (HOST pointers).  Lets the counts cross PCIe instead of the four times larger float32 features; valid when
 * every column is kept (d_keep == d). */
GG_API int gg_host_features_f32(const uint8_t* h_counts, int64_t n_elems, const float* h_value_lut, float* h_x,
                         int n_threads);
/* Asynchronous flavour: returns a job handle (null on a bad argument) at once; a native thread first waits for
 * `cuda_event` (a cudaEvent_t of device `device` marking the arrival of h_counts on the host; may be null) and then
 * expands.  The buffers must stay valid until gg_host_job_wait(handle), which joins and returns the job's status. */
GG_API void* gg_host_features_f32_start(const uint8_t* h_counts, int64_t n_elems, const float* h_value_lut, float* h_x,
                                 int n_threads, int device, void* cuda_event);
GG_API int gg_host_job_wait(void* handle);

/* ------------------------------------------------------------------------
 * f3  biased random walks + skip-gram windows -> positive node pairs.  Replaces sample_biased_random_walks
 *     (gnn_common/gnn_utils.py:514-600; called per epoch / per batch by the sampling tasks).
 *
 * Graph: CSR over the directed weighted De Bruijn graph (row_ptr[n + 1], col[E], weight[E]; successors in edge
 * order).  Walk w = round * n_start + s starts at start_nodes[s]; first step ~ edge weights, later steps node2vec
 * (weight / p back to the previous node, weight if the successor has an edge to the previous node, weight / q
 * otherwise); a node without successors ends the walk.  For every position i a half-width ~ U{1..window_size} and the
 * pairs (walk[i], walk[j]), j != i inside the window.  Every walk owns a Philox stream keyed by (seed, w): mode 0
 * (pair_counts[w]) and mode 1 (pairs written from pair_offsets[w], the exclusive prefix sum of the counts) replay the
 * same walks.  walks_out (mode 1, nullable): [n_walks][walk_length + 1], -1 padded.  walk_length <= 32.
 * ------------------------------------------------------------------------ */
GG_API int gg_rw_pairs(const int64_t* row_ptr, const int64_t* col, const float* weight, int64_t n,
                const int64_t* start_nodes, int64_t n_start, int num_walks, int walk_length, int window_size,
                double p, double q, uint64_t seed, int mode, int64_t* pair_counts, const int64_t* pair_offsets,
                int64_t* pairs_src, int64_t* pairs_dst, int64_t capacity, int64_t* walks_out, gg_stream_t st);

/* ------------------------------------------------------------------------
 * f1  sparse aggregation of the GCN layers trained on the graph (gnn_common/gnn_models.py:58-100 uses PyG's GCNConv:
 *     out[i] = sum over edges (j -> i) of norm_ji * h[j]).  y = A x, A in CSR by target node (col = source node,
 *     val = normalised weight), x / y dense float32 [n_rows, d], d <= 128.  Fixed summation order (reproducible).
 * ------------------------------------------------------------------------ */
GG_API int gg_spmm_csr_f32(const int64_t* row_ptr, const int32_t* col, const float* val, int64_t n_rows, const float* x,
                    int d, float* y, gg_stream_t st);

/* ------------------------------------------------------------------------
 * f4b  per-row k-nearest-neighbour KF edges: the --representation_ss_faiss_ann mode.  Replaces
 *      kf_faiss_to_edge_index_weight (gnn_common/gnn_utils.py:122-237; k = int(edges_keep_top_k * N),
 *      gnn_tasks_miniBatch/dataloader.py:108-121) by the EXACT search its FAISS index approximates.
 *
 * metric 0 ("IP"): cosine of the count rows; metric 1 ("L2"): squared Euclidean distance, returned as
 * 1 - d / (largest kept d) with the neighbours at that largest distance dropped (similarity <= 1e-4).  Self matches
 * are excluded; "IP" neighbours with a zero dot are never returned.
 * Tables from the host: sqid[sq_max + 1] maps a squared norm to its class (0xFF: no such row), rank_tab[n_sq]
 * [dot_max + 1] ranks the (class of column j, dot) pairs by similarity (0 = most similar, 0xFFFF = never a
 * neighbour; equal similarity <=> equal rank), rank_q[n_ranks] (L2) = sqnorm_j - 2 dot of a rank.
 *   gg_knn_count  hist[n][n_ranks] = columns j != i of every rank, per row
 *   gg_knn_plan   cut[n] (rank holding the k-th neighbour), tie_left[n] (members of that rank that still fit),
 *                 row_offsets[n + 1] (LAST entry = number of edges), qmax; hist becomes the cursors of gg_knn_write
 *   gg_knn_write  second pass (the selected pairs of every row, in column order, into `recs`) + reorder: edges
 *                 grouped by source row, most similar first, ties by ascending column.
 * Rows are independent: every call works on the row shard [row_begin, row_end) (row_begin a multiple of 128; the
 * multi-GPU unit) against all n columns, with shard-local per-row arrays (m = row_end - row_begin entries) and
 * shard-local output slots; "L2" needs the maximum of qmax over the shards between the two phases of gg_knn_plan.
 * pitch: bytes per row of counts: 4, 8, or a multiple of 16 (of 128 above 128) up to 1024.
 * ------------------------------------------------------------------------ */
/* Which Gram kernel serves both passes: 0 = auto (the tcgen05 CTA-pair kernel for rows of 128..1024 bytes with at most
 * 32 norm classes, else the mma.sync kernel), 1 = mma.sync only, 2 = tcgen05 or an error, 3 = as 0 (names the sparse
 * flavour below for gg_knn_sparse_eligible).  Process-wide. */
GG_API int gg_knn_set_impl(int impl);
GG_API int gg_knn_count(const uint8_t* counts, int pitch, int64_t n, int64_t row_begin, int64_t row_end,
                 const int32_t* sqnorm, const uint8_t* sqid, int sq_max, const uint16_t* rank_tab, int n_sq, int dot_max,
                 int n_ranks, int hist_mode /* -1 auto; 0 global atomics, 1 / 2 per-CTA shared memory, 16- / 32-bit bins */,
                 uint32_t* hist, gg_stream_t st);
GG_API size_t gg_knn_plan_workspace_bytes(int64_t n_rows);
GG_API int gg_knn_plan(int64_t n, int64_t row_begin, int64_t row_end, int n_ranks, int k_nn, int metric,
                const int32_t* sqnorm, const int32_t* rank_q, uint32_t* hist, int32_t* cut, int32_t* tie_left,
                int64_t* row_offsets, int32_t* qmax,
                int phase /* 0 all; 1 cut ranks + the shard's qmax; 2 finish with the global qmax in place */,
                void* workspace, size_t workspace_bytes, gg_stream_t st);
/* 64-bit words of `recs` scratch for a shard of n_rows rows / n_edges edges (the tcgen05 kernel stages more) */
GG_API size_t gg_knn_write_scratch_words(int pitch, int64_t n, int n_sq, int dot_max, int n_ranks, int64_t n_rows,
                                  int64_t n_edges);
GG_API int gg_knn_write(const uint8_t* counts, int pitch, int64_t n, int64_t row_begin, int64_t row_end,
                 const int32_t* sqnorm, const uint8_t* sqid, int sq_max, const uint16_t* rank_tab, int n_sq, int dot_max,
                 int n_ranks, int metric, uint32_t* cursors, const int32_t* cut, const int32_t* tie_left,
                 const int64_t* row_offsets, const int32_t* qmax,
                 uint64_t* recs /* scratch */, int64_t recs_words /* >= gg_knn_write_scratch_words() */,
                 int64_t* edge_src, int64_t* edge_dst, float* weight, int64_t capacity /* = row_offsets[m] */,
                 gg_stream_t st);

/* Sparse flavour of the two passes (rows with at most 16 non-zero counts -- rows of sub-k-mer counts have at most
 * k - small_k + 1 --, 64..1024 columns, up to 200,000 nodes): an inverted-index join that never forms the zero dots.
 * gg_knn_sparse_index builds row lists + posting lists + norm classes into a caller-owned workspace
 * (gg_knn_sparse_index_bytes); idx_status (device int64[2]): [0] != 0 = a row holds more than 16 non-zero counts, the
 * index must not be used.  gg_knn_sparse_count / gg_knn_sparse_write replace gg_knn_count / gg_knn_write between the same
 * gg_knn_plan calls (recs: row_offsets[m] words); results are identical. */
/* The KF THRESHOLD edges (gg_kf_emit's job) of the same sparse wide rows through the same index: a pair can only be an
 * edge when it shares a column.  Outputs as gg_kf_emit (records + rowcnt + kf_status[0]); feed them to gg_kf_place. */
GG_API int gg_kf_emit_sparse(void* index, size_t index_bytes, int pitch, int64_t n, int iblock_begin, int iblock_end,
                      const int32_t* sqnorm, const uint16_t* dmin, gg_kf_rec_t* recs, int64_t rec_capacity, uint64_t* rowcnt,
                      int64_t* kf_status, gg_stream_t st);
GG_API int gg_knn_sparse_eligible(int pitch, int64_t n, int n_sq, int n_ranks);
GG_API size_t gg_knn_sparse_index_bytes(int64_t n, int pitch);
GG_API int gg_knn_sparse_index(const uint8_t* counts, int pitch, int64_t n, const int32_t* sqnorm, const uint8_t* sqid,
                        int sq_max, int n_sq, void* index, size_t index_bytes, int64_t* idx_status, gg_stream_t st);
GG_API int gg_knn_sparse_count(void* index, size_t index_bytes, int pitch, int64_t n, int64_t row_begin, int64_t row_end,
                        const int32_t* sqnorm, const uint16_t* rank_tab, int n_sq, int dot_max, int n_ranks, uint32_t* hist,
                        gg_stream_t st);
GG_API int gg_knn_sparse_write(void* index, size_t index_bytes, int pitch, int64_t n, int64_t row_begin, int64_t row_end,
                        const int32_t* sqnorm, const uint16_t* rank_tab, int n_sq, int dot_max, int n_ranks, int metric,
                        uint32_t* cursors, const int32_t* cut, const int32_t* tie_left, const int64_t* row_offsets,
                        const int32_t* qmax, uint64_t* recs, int64_t* edge_src, int64_t* edge_dst, float* weight,
                        int64_t capacity, gg_stream_t st);

/* ------------------------------------------------------------------------
 * Multi-GPU exchanges over NVLink peer memory (no reference counterpart: the reference is single-device).
 * Every rank owns one buffer of a symmetric allocation that all ranks of the box have mapped, plus a pad of 32-bit
 * flags; the HOST arrays below hold, per rank, the address of that rank's buffer (+ the common offset) / flag pad in
 * the caller's address space (null = skip that rank, e.g. the caller itself).
 *   gg_peer_push    copy n_bytes (multiple of 16) from src to h_dst->ptr[r] for every r -- my segment into the same
 *                   place of every peer's buffer;
 *   gg_peer_signal  after everything earlier on the stream: flag pad of every peer, word `index`, = value
 *                   (system-scope release);
 *   gg_peer_wait    return (on the stream) once flags[first + r] has reached `value` for every r in [0, count) but
 *                   `skip` (serial numbers: compared as signed differences); a peer that never arrives after 2^26
 *                   polls (~10 s) increments status[0] instead of hanging the device.
 * ------------------------------------------------------------------------ */
#define GG_MAX_PEERS 16
typedef struct {
  int n;
  void* ptr[GG_MAX_PEERS];
} gg_peers_t;
GG_API int gg_peer_push(const void* src, int64_t n_bytes, const gg_peers_t* h_dst, gg_stream_t st);
GG_API int gg_peer_signal(const gg_peers_t* h_flags, int index, uint32_t value, gg_stream_t st);
GG_API int gg_peer_wait(const uint32_t* flags, int first, int count, int skip, uint32_t value, int64_t* status,
                 gg_stream_t st);

/* ------------------------------------------------------------------------
 * Hardware-ceiling probes (measurement only; no reference counterpart).  One kernel launch per call on the caller's
 * stream; *h_updates / *h_ops (HOST out-parameters) = operations that launch performs, to be divided by the launch's
 * CUDA-event time.
 *   gg_probe_smem_atomics  K1's secondary ceiling: uniform 32-bit shared-memory atomic adds into a `bins`-entry table
 *                          per CTA (1024 threads, one CTA per SM) -- the update pattern of the bucket-count kernel.
 *                          scratch: >= number of SMs uint32.
 *   gg_probe_tensor_i8     K4's ceiling: the CTA-pair kernel's tcgen05.mma.cta_group::2.kind::i8 stream (M = 256,
 *                          N = 256, K = 32; uint8 operands resident in shared memory, n_kblocks x 128-byte K blocks
 *                          per tile) with no TMA and no epilogue.  *h_ops counts 2 ops per multiply-accumulate.
 *                          error_flag: device int32, set to 1 on a barrier time-out.
 * ------------------------------------------------------------------------ */
GG_API int gg_probe_smem_atomics(int bins, int iters, uint32_t* scratch, double* h_updates, gg_stream_t st);
GG_API int gg_probe_tensor_i8(int iters, int n_kblocks, int32_t* error_flag, double* h_ops, gg_stream_t st);

#ifdef __cplusplus
}
#endif
#endif /* GG_B200_H */
