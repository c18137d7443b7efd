// Micro-benchmark: throughput of random-address 32-bit atomic adds in shared memory vs global (L2) on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
template <int BINS>
__global__ void smem_atomics(unsigned* out, int iters) {
  extern __shared__ unsigned h[];
  for (int i = threadIdx.x; i < BINS; i += blockDim.x) h[i] = 0;
  __syncthreads();
  unsigned x = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = 0; i < iters; ++i) { x = hash(x + i); atomicAdd(&h[x % BINS], 1u); }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = h[0];
}
__global__ void gmem_atomics(unsigned* table, unsigned bins, int iters) {
  unsigned x = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = 0; i < iters; ++i) { x = hash(x + i); atomicAdd(&table[x % bins], 1u); }
}
__global__ void hash_only(unsigned* out, int iters) {
  unsigned x = blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
  for (int i = 0; i < iters; ++i) { x = hash(x + i); acc += x % 32768; }
  if (acc == 12345) out[0] = acc;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned *out, *table; cudaMalloc(&out, 1 << 20); cudaMalloc(&table, 1 << 22); cudaMemset(table, 0, 1 << 22);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 4096, threads = 1024;
  auto run = [&](const char* name, auto launch, double n_updates) {
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-28s %8.3f ms  %8.1f G updates/s  (%s)\n", name, ms, n_updates / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  };
  const double n = double(sms) * threads * iters;
  run("hash only (ALU floor)", [&] { hash_only<<<sms, threads>>>(out, iters); }, n);
  cudaFuncSetAttribute(smem_atomics<32768>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4);
  run("smem atomics 32768 bins", [&] { smem_atomics<32768><<<sms, threads, 32768 * 4>>>(out, iters); }, n);
  run("smem atomics 4096 bins", [&] { smem_atomics<4096><<<sms, threads, 4096 * 4>>>(out, iters); }, n);
  run("smem atomics 256 bins", [&] { smem_atomics<256><<<sms, threads, 256 * 4>>>(out, iters); }, n);
  run("L2 atomics 262144 bins", [&] { gmem_atomics<<<sms * 2, threads>>>(table, 262144, iters / 2); }, n);
  run("L2 atomics 1048576 bins", [&] { gmem_atomics<<<sms * 2, threads>>>(table, 1048576, iters / 2); }, n);
  run("L2 atomics 16384 bins", [&] { gmem_atomics<<<sms * 2, threads>>>(table, 16384, iters / 2); }, n);
  return 0;
}
