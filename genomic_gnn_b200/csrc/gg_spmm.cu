// f1: sparse aggregation of the GCN layers that consume the graph (reference gnn_common/gnn_models.py:58-100 builds
// them from torch_geometric's GCNConv, whose propagate step is  out[i] = sum over edges (j -> i) of norm_ji * h[j]).
//
// y = A x with A in CSR by TARGET node (row_ptr[n + 1], col = source node, val = normalised edge weight) and x, y dense
// row-major float32 [n, d].  One warp per row; a lane owns the features lane, lane + 32, ... (d <= 128: at most four
// accumulators), the row's edges are walked sequentially, so every x row is read as one coalesced segment and the
// summation order is fixed (the kernel is bit-reproducible).  HBM/L2-bound: 4 (d + 3) bytes per edge.  The backward pass
// with respect to x is the same kernel on the transposed CSR.
#include "gg_common.cuh"

namespace {

constexpr int kMaxAcc = 4;  // d <= 128

__global__ void __launch_bounds__(256) spmm_kernel(const int64_t* row_ptr, const int32_t* col, const float* val,
                                                   long long n_rows, const float* x, int d, float* y) {
  const long long row = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  float acc[kMaxAcc] = {0.f, 0.f, 0.f, 0.f};
  const long long e0 = row_ptr[row], e1 = row_ptr[row + 1];
  for (long long e = e0; e < e1; ++e) {
    const float w = __ldg(val + e);
    const float* src = x + static_cast<long long>(__ldg(col + e)) * d;
#pragma unroll
    for (int q = 0; q < kMaxAcc; ++q) {
      const int f = lane + 32 * q;
      if (f < d) acc[q] = fmaf(w, __ldg(src + f), acc[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < kMaxAcc; ++q) {
    const int f = lane + 32 * q;
    if (f < d) y[row * d + f] = acc[q];
  }
}

}  // namespace

extern "C" int gg_spmm_csr_f32(const int64_t* row_ptr, const int32_t* col, const float* val, int64_t n_rows,
                               const float* x, int d, float* y, gg_stream_t st) {
  GG_REQUIRE(n_rows >= 0 && d >= 1 && d <= 32 * kMaxAcc, GG_ERR_ARG, "d outside [1, 128]");
  if (n_rows == 0) return GG_OK;
  GG_REQUIRE(row_ptr && x && y, GG_ERR_ARG, "null buffer");
  const long long threads = n_rows * 32;
  spmm_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, gg::as_stream(st)>>>(row_ptr, col, val, n_rows, x, d,
                                                                                          y);
  GG_CUDA_CHECK(cudaGetLastError());
  return GG_OK;
}
