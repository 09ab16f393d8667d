// How fast can a CTA sweep a vector set that other CTAs have published?  148 CTAs x 512 threads each read the SAME
// n 16-byte pieces from L2 (all CTAs at once, as the decoder's consumers do) with four load flavours.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/tagged_load_bench.cu -o /tmp/tlb && /tmp/tlb
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
constexpr int THREADS = 512;

template <int V>
__device__ __forceinline__ void load16(const unsigned long long* src, unsigned long long& a, unsigned long long& b) {
  if (V == 0) asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(src) : "memory");
  if (V == 1) asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(src) : "memory");
  if (V == 2) asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(src) : "memory");
  if (V == 3) {
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(a) : "l"(src) : "memory");
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(b) : "l"(src + 1) : "memory");
  }
  if (V == 4) {
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a) : "l"(src) : "memory");
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(b) : "l"(src + 1) : "memory");
  }
  if (V == 5) asm volatile("ld.global.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(src) : "memory");
  if (V == 6) asm volatile("ld.global.cv.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(src) : "memory");
  if (V == 7) asm volatile("ld.global.lu.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(src) : "memory");
  if (V == 8) asm volatile("ld.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(src) : "memory");
}

template <int V>
__global__ void __launch_bounds__(THREADS, 1) sweep(const unsigned long long* data, int pieces, int iters, long long* cycles,
                                                    unsigned long long* sink) {
  __shared__ float2 stage[4096];
  unsigned long long acc = 0;
  long long t0 = 0;
  for (int it = 0; it < iters + 2; ++it) {
    if (it == 2) t0 = clock64();
    for (int i0 = threadIdx.x; i0 < pieces; i0 += 5 * THREADS) {
      unsigned long long a[5], b[5];
#pragma unroll
      for (int j = 0; j < 5; ++j)
        if (i0 + j * THREADS < pieces) load16<V>(data + 2 * (size_t)(i0 + j * THREADS), a[j], b[j]);
#pragma unroll
      for (int j = 0; j < 5; ++j)
        if (i0 + j * THREADS < pieces) {
          stage[(i0 + j * THREADS) & 4095] = make_float2(__uint_as_float((unsigned)a[j]), __uint_as_float((unsigned)b[j]));
          acc += a[j] >> 32;
        }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) cycles[blockIdx.x] = (clock64() - t0) / iters;
  if (acc == 0x123456789ull) *sink = acc + (unsigned long long)stage[5].x;
}

// latency of ONE dependent load per iteration (thread 0 only): the next address depends on the loaded value
template <int V>
__global__ void chase(const unsigned long long* data, int iters, long long* cycles, unsigned long long* sink) {
  if (threadIdx.x != 0) return;
  unsigned long long off = 0, a, b;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    load16<V>(data + off, a, b);
    off = (a & 1) * 2 + ((off + 1024) & 65535);      // a is 0: walks through 64 K words = 512 KB, always an L2 hit
  }
  cycles[0] = (clock64() - t0) / iters;
  if (off == 0x12345) *sink = off;
}
template <int V>
void run_chase(const char* name) {
  unsigned long long *data, *sink; long long* cyc;
  CK(cudaMalloc(&data, 1 << 20)); CK(cudaMemset(data, 0, 1 << 20)); CK(cudaMalloc(&sink, 8)); CK(cudaMalloc(&cyc, 8));
  chase<V><<<1, 32>>>(data, 2000, cyc, sink); CK(cudaDeviceSynchronize());
  chase<V><<<1, 32>>>(data, 2000, cyc, sink); CK(cudaDeviceSynchronize());
  long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  printf("dependent-load latency, %-30s %lld cycles\n", name, h); fflush(stdout);
  cudaFree(data); cudaFree(sink); cudaFree(cyc);
}
template <int V>
void run(const char* name, int sms, int pieces, int ctas) {
  unsigned long long *data, *sink;
  long long* cyc;
  CK(cudaMalloc(&data, 1 << 20));
  CK(cudaMemset(data, 0, 1 << 20));
  CK(cudaMalloc(&sink, 8));
  CK(cudaMalloc(&cyc, 8 * 256));
  int iters = 500;
  sweep<V><<<ctas, THREADS>>>(data, pieces, iters, cyc, sink);
  CK(cudaDeviceSynchronize());
  sweep<V><<<ctas, THREADS>>>(data, pieces, iters, cyc, sink);
  CK(cudaDeviceSynchronize());
  long long h[256];
  CK(cudaMemcpy(h, cyc, 8 * ctas, cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-28s %5d pieces (%3d KB) x %3d CTAs: %6lld cycles per sweep (slowest CTA)\n", name, pieces, pieces * 16 / 1024, ctas, mx);
  fflush(stdout);
  cudaFree(data); cudaFree(sink); cudaFree(cyc);
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  run_chase<0>("ld.volatile.v2.u64"); run_chase<1>("ld.relaxed.gpu.v2.u64"); run_chase<2>("ld.global.cg.v2.u64");
  run_chase<5>("ld.global.L1::no_allocate.v2"); run_chase<6>("ld.global.cv.v2.u64"); run_chase<7>("ld.global.lu.v2.u64");
  run_chase<8>("ld.global.v2.u64 (L1)");
  for (int ctas : {1, sms})
    for (int pieces : {150, 1200}) {
      run<0>("ld.volatile.v2.u64", sms, pieces, ctas);
      run<1>("ld.relaxed.gpu.v2.u64", sms, pieces, ctas);
      run<2>("ld.global.cg.v2.u64", sms, pieces, ctas);
      run<3>("2 x ld.volatile.u64", sms, pieces, ctas);
      run<4>("2 x ld.relaxed.gpu.u64", sms, pieces, ctas);
      run<5>("ld.global.L1::no_allocate.v2", sms, pieces, ctas);
      run<6>("ld.global.cv.v2.u64", sms, pieces, ctas);
      run<7>("ld.global.lu.v2.u64", sms, pieces, ctas);
      run<8>("ld.global.v2.u64 (L1, stale!)", sms, pieces, ctas);
    }
  return 0;
}
