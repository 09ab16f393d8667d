// What does publishing n tagged 8-byte words cost the publishing CTA?  (the decoder's epilogues)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/tagged_store_bench.cu -o /tmp/tsb && /tmp/tsb
#include <cstdio>
#include <cuda_runtime.h>
template <int V>
__device__ __forceinline__ void st8(unsigned long long* dst, unsigned long long v) {
  if (V == 0) asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst), "l"(v) : "memory");
  if (V == 1) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(dst), "l"(v) : "memory");
  if (V == 2) asm volatile("st.global.cg.u64 [%0], %1;" ::"l"(dst), "l"(v) : "memory");
  if (V == 3) asm volatile("st.global.u64 [%0], %1;" ::"l"(dst), "l"(v) : "memory");
}
template <int V>
__global__ void __launch_bounds__(512, 1) k(unsigned long long* buf, int n, int stride, int reps, long long* out) {
  long long total = 0;
  for (int r = 0; r < reps; ++r) {
    __syncthreads();
    const long long t0 = clock64();
    if ((int)threadIdx.x < n) st8<V>(buf + (size_t)blockIdx.x * 65536 + (threadIdx.x / 28) * stride + threadIdx.x % 28, ((unsigned long long)r << 32) | threadIdx.x);
    __syncthreads();
    const long long t1 = clock64();
    if (r > 0) total += t1 - t0;
    for (int i = 0; i < 200; ++i) __nanosleep(20);   // let the stores drain
  }
  if (threadIdx.x == 0) out[blockIdx.x] = total / (reps - 1);
}
int main() {
  unsigned long long* buf; long long* out;
  cudaMalloc(&buf, 148ull * 65536 * 8); cudaMalloc(&out, 8 * 148);
  const char* names[4] = {"st.volatile", "st.relaxed.gpu", "st.global.cg (weak)", "st.global (weak)"};
  for (int ctas : {1, 148})
    for (int n : {48, 224, 448}) {
      long long h[4];
      k<0><<<ctas, 512>>>(buf, n, 300, 20, out); cudaMemcpy(&h[0], out, 8, cudaMemcpyDeviceToHost);
      k<1><<<ctas, 512>>>(buf, n, 300, 20, out); cudaMemcpy(&h[1], out, 8, cudaMemcpyDeviceToHost);
      k<2><<<ctas, 512>>>(buf, n, 300, 20, out); cudaMemcpy(&h[2], out, 8, cudaMemcpyDeviceToHost);
      k<3><<<ctas, 512>>>(buf, n, 300, 20, out); cudaMemcpy(&h[3], out, 8, cudaMemcpyDeviceToHost);
      printf("%3d CTAs, %3d words each (runs of 28 at stride 300):", ctas, n);
      for (int v = 0; v < 4; ++v) printf("  %s %lld", names[v], h[v]);
      printf(" cycles\n");
    }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
