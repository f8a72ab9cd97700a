// Micro-benchmark: DRAM fetch granularity on B200.  Reads `bytes` (16/32/64) at the start of every
// `stride`-byte slot of a buffer much larger than L2; run under
//   ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum ./dram_gran
// and divide dram bytes by the number of slots.  (Evidence for the pyramid subtile size)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void rd(const uint4* __restrict__ buf, long long slots, int stride16, int n16, unsigned* sink) {
  unsigned acc = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < slots; i += (long long)gridDim.x * blockDim.x) {
    // scatter slots so consecutive threads do not hit consecutive slots
    long long j = (i * 7919) % slots;
    for (int k = 0; k < n16; ++k) {
      uint4 v;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(buf + j * stride16 + k));
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}
int main() {
  const size_t bytes = 8ull << 30;
  uint4* buf; unsigned* sink;
  cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4);
  cudaMemset(buf, 1, bytes);
  const int strides[] = {64, 128, 256, 512, 1024};
  const int sizes[] = {16, 32, 64};
  for (int s : strides) for (int z : sizes) {
    if (z > s) continue;
    long long slots = (long long)(bytes / s);
    if (slots > (1ll << 24)) slots = 1ll << 24;  // 16M accesses
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    rd<<<148 * 16, 256>>>(buf, slots, s / 16, z / 16, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("stride %4d read %2d B: %lld slots, %.3f ms, useful %.1f GB/s\n", s, z, slots, ms, slots * (double)z / ms / 1e6);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
