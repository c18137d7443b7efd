// Hardware-ceiling probes behind bench.py's secondary rooflines.  Each call launches ONE kernel on the caller's stream
// and reports (through a host out-parameter) how many operations that launch performs; the caller times it with CUDA
// events.  They replace no reference code: they give the measured denominators for the two stages whose bound is not
// HBM bandwidth -- K1 counting (shared-memory atomic rate) and the K4 Gram (kind::i8 tensor-pipe rate).
#include "gg_common.cuh"

namespace ggkf {
int tensor_probe(int iters, int n_kblocks, int* error_flag, double* h_ops, cudaStream_t s);  // gg_kf_mma.cu
}

namespace {

__device__ __forceinline__ unsigned mix(unsigned x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}

// the update pattern of count_bucket_kernel: 1024 threads, a 32,768-bin uint32 table per CTA, one ATOMS per key,
// keys uniform over the table
__global__ void __launch_bounds__(1024) smem_atomic_probe_kernel(unsigned* out, int bins, int iters) {
  extern __shared__ unsigned table[];
  for (int i = threadIdx.x; i < bins; i += blockDim.x) table[i] = 0;
  __syncthreads();
  unsigned x = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned mask = static_cast<unsigned>(bins - 1);
  for (int i = 0; i < iters; ++i) {
    x = mix(x + i);
    atomicAdd(&table[x & mask], 1u);
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = table[0];
}

}  // namespace

extern "C" int gg_probe_smem_atomics(int bins, int iters, uint32_t* scratch, double* h_updates, gg_stream_t st) {
  GG_REQUIRE(bins >= 32 && bins <= 32768 && (bins & (bins - 1)) == 0 && iters >= 1 && scratch && h_updates, GG_ERR_ARG,
             "bins must be a power of two in [32, 32768]");
  const size_t smem = static_cast<size_t>(bins) * sizeof(unsigned);
  GG_CUDA_CHECK(cudaFuncSetAttribute(smem_atomic_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
  const int sms = gg::sm_count();
  smem_atomic_probe_kernel<<<sms, 1024, smem, gg::as_stream(st)>>>(scratch, bins, iters);
  GG_CUDA_CHECK(cudaGetLastError());
  *h_updates = static_cast<double>(sms) * 1024.0 * iters;
  return GG_OK;
}

extern "C" int gg_probe_tensor_i8(int iters, int n_kblocks, int32_t* error_flag, double* h_ops, gg_stream_t st) {
  return ggkf::tensor_probe(iters, n_kblocks, error_flag, h_ops, gg::as_stream(st));
}
