// Latency / issue rate of the legacy warp-level mma.sync.m16n8k16 (f16 in, f32 accumulate) on sm_100a, the
// instruction behind the decoder's and the BiLSTM's small mat-vecs.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/hmma_bench.cu -o /tmp/hb && /tmp/hb
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int CHAINS>
__global__ void k(long long* out, float* sink, int n) {
  uint32_t a[4] = {threadIdx.x, 2, 3, 4}, b[2] = {5, threadIdx.x};
  float d[CHAINS][4] = {};
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) mma(d[c], a, b);
  const long long t1 = clock64();
  float s = 0;
  for (int c = 0; c < CHAINS; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
  if (s == 1234.5f) *sink = s;
  if (threadIdx.x == 0) out[0] = (t1 - t0);
}
int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 8); cudaMalloc(&sink, 4);
  const int n = 256;
  for (int warps : {1, 4, 8, 16}) {
    long long h1, h3, h6;
    k<1><<<1, 32 * warps>>>(out, sink, n); cudaMemcpy(&h1, out, 8, cudaMemcpyDeviceToHost);
    k<3><<<1, 32 * warps>>>(out, sink, n); cudaMemcpy(&h3, out, 8, cudaMemcpyDeviceToHost);
    k<6><<<1, 32 * warps>>>(out, sink, n); cudaMemcpy(&h6, out, 8, cudaMemcpyDeviceToHost);
    printf("%2d warps: dependent chain %.1f cycles per mma; 3 chains %.1f per mma; 6 chains %.1f per mma (per warp)\n", warps,
           (double)h1 / n, (double)h3 / (3 * n), (double)h6 / (6 * n));
  }
  return 0;
}
