// Micro-benchmark: random 32-bit atomic adds into a cluster-distributed shared-memory table (DSMEM).
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int CLUSTER, int BINS_PER_CTA>
__global__ void dsmem_atomics(unsigned* out, int iters) {
  extern __shared__ unsigned h[];
  cg::cluster_group cluster = cg::this_cluster();
  for (int i = threadIdx.x; i < BINS_PER_CTA; i += blockDim.x) h[i] = 0;
  cluster.sync();
  unsigned* remote[CLUSTER];
#pragma unroll
  for (int r = 0; r < CLUSTER; ++r) remote[r] = cluster.map_shared_rank(h, r);
  unsigned x = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = 0; i < iters; ++i) {
    x = hash(x + i);
    const unsigned key = x % (CLUSTER * BINS_PER_CTA);
    unsigned* base = remote[0];
#pragma unroll
    for (int r = 1; r < CLUSTER; ++r) if (key / BINS_PER_CTA == r) base = remote[r];
    atomicAdd(base + key % BINS_PER_CTA, 1u);
  }
  cluster.sync();
  if (threadIdx.x == 0) out[blockIdx.x] = h[0];
}

template <int CLUSTER, int BINS>
void run(const char* name, int sms, unsigned* out) {
  const int iters = 1024, threads = 1024;
  auto kern = dsmem_atomics<CLUSTER, BINS>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BINS * 4);
  if (CLUSTER > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  const int grid = sms / CLUSTER * CLUSTER;
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = BINS * 4;
  cudaLaunchAttribute attr[1]; attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaLaunchKernelEx(&cfg, kern, out, iters); cudaDeviceSynchronize();
  cudaEventRecord(a); cudaLaunchKernelEx(&cfg, kern, out, iters); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double n = double(grid) * threads * iters;
  printf("%-40s grid %3d  %8.3f ms  %8.1f G updates/s (%s)\n", name, grid, ms, n / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned* out; cudaMalloc(&out, 1 << 20);
  run<1, 32768>("cluster 1  (local only, 128 KB)", sms, out);
  run<2, 32768>("cluster 2  x 128 KB", sms, out);
  run<4, 32768>("cluster 4  x 128 KB", sms, out);
  run<8, 32768>("cluster 8  x 128 KB = 1 MB (4^9 bins)", sms, out);
  run<16, 16384>("cluster 16 x  64 KB = 1 MB (4^9 bins)", sms, out);
  return 0;
}
