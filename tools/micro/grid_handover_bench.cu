// Microbenchmark behind the decoder's hand-over design (profiles/r2_grid_handover_bench.txt): cycles per grid-wide
// hand-over on a full cooperative grid (1 CTA of 512 threads per SM) for several barrier forms, with and without
// a payload (every CTA writes 24 floats before, and reads the whole 8 x 1200 float vector set after, like a
// decoder mat-vec phase).  Build + run on the GPU box:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/grid_handover_bench.cu -o /tmp/ghb && /tmp/ghb
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int THREADS = 512;

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

// V0: the decoder's current barrier
__device__ __forceinline__ void bar_v0(unsigned* counter, unsigned& target, unsigned n) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += n;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned seen;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory"); } while (seen < target);
  }
  __syncthreads();
}
// V1: fence + relaxed add, relaxed polling, one fence at the end
__device__ __forceinline__ void bar_v1(unsigned* counter, unsigned& target, unsigned n) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += n;
    __threadfence();
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned seen;
    do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory"); } while (seen < target);
    __threadfence();
  }
  __syncthreads();
}
// V2: one flag per CTA (no atomics): st.release own flag, warp 0 polls all flags
__device__ __forceinline__ void bar_v2(unsigned* flags, unsigned& target, unsigned n) {
  __syncthreads();
  target += 1;
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x), "r"(target) : "memory");
    bool ok;
    do {
      ok = true;
      for (unsigned i = threadIdx.x; i < n; i += 32) {
        unsigned seen;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(flags + i) : "memory");
        ok = ok && seen >= target;
      }
      ok = __all_sync(0xffffffffu, ok);
    } while (!ok);
    __threadfence();
  }
  __syncthreads();
}
// V3: like V0, but the arrival is issued by the thread right after ITS OWN stores + a CTA barrier that only orders
// (bar.arrive / bar.sync split: the poller does not wait for the syncthreads of the writers)
__device__ __forceinline__ void bar_v3(unsigned* counter, unsigned& target, unsigned n) {
  // writers: fence own stores, then arrive on named barrier 1; thread 0 syncs on it, then adds
  __threadfence();
  if (threadIdx.x < 32) {
    asm volatile("bar.sync 1, %0;" ::"r"(THREADS) : "memory");
  } else {
    asm volatile("bar.arrive 1, %0;" ::"r"(THREADS) : "memory");
  }
  if (threadIdx.x == 0) {
    target += n;
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned seen;
    do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory"); } while (seen < target);
    __threadfence();
  }
  __syncthreads();
}

// V4: no fences at all (floor of "atomic + poll"; NOT a correct hand-over for weak stores)
__device__ __forceinline__ void bar_v4(unsigned* counter, unsigned& target, unsigned n) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += n;
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned seen;
    do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory"); } while (seen < target);
  }
  __syncthreads();
}
// V5: release on the arrival, relaxed spin, no acquire fence (the payload is read with L2-only loads afterwards)
__device__ __forceinline__ void bar_v5(unsigned* counter, unsigned& target, unsigned n) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += n;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned seen;
    do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory"); } while (seen < target);
  }
  __syncthreads();
}
// V6: every writer thread fences its own stores (in parallel), then plain barrier without a fence in thread 0
__device__ __forceinline__ void bar_v6(unsigned* counter, unsigned& target, unsigned n, bool wrote) {
  if (wrote) __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    target += n;
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned seen;
    do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory"); } while (seen < target);
  }
  __syncthreads();
}
// V7: volatile (ld.volatile) spin like the decoder's tagged hand-over, atomicAdd arrival, no fences
__device__ __forceinline__ void bar_v7(unsigned* counter, unsigned& target, unsigned n) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += n;
    atomicAdd(counter, 1u);
    while (*reinterpret_cast<volatile unsigned*>(counter) < target) {}
  }
  __syncthreads();
}

// one-way latency: CTA 0 and CTA 1 bounce an 8-byte (value, tag) word
__global__ void pingpong(unsigned long long* slots, int iters, long long* cycles) {
  if (threadIdx.x != 0 || blockIdx.x > 1) return;
  volatile unsigned long long* mine = slots + blockIdx.x * 16;
  volatile unsigned long long* other = slots + (1 - blockIdx.x) * 16;
  long long t0 = clock64();
  for (int it = 1; it <= iters; ++it) {
    if (blockIdx.x == 0) {
      *other = ((unsigned long long)it << 32) | 7u;
      while ((unsigned)(*mine >> 32) != (unsigned)it) {}
    } else {
      while ((unsigned)(*mine >> 32) != (unsigned)it) {}
      *other = ((unsigned long long)it << 32) | 7u;
    }
  }
  if (blockIdx.x == 0) cycles[0] = (clock64() - t0) / iters;
}

template <int V>
__global__ void __launch_bounds__(THREADS, 1) bench(unsigned* sync, float* data, int iters, int payload_floats, int write_floats,
                                                    long long* cycles, float* sink) {
  extern __shared__ float4 stage[];
  unsigned target = 0;
  const unsigned n = gridDim.x;
  float acc = 0.f;
  long long t0 = 0;
  for (int it = 0; it < iters + 4; ++it) {
    if (it == 4) t0 = clock64();
    float* buf = data + (size_t)(it & 1) * payload_floats;
    if (write_floats > 0 && threadIdx.x < write_floats) {
      const int i = blockIdx.x * write_floats + threadIdx.x;
      if (i < payload_floats) buf[i] = (float)it;
    }
    if (V == 0) bar_v0(sync, target, n);
    if (V == 1) bar_v1(sync, target, n);
    if (V == 2) bar_v2(sync, target, n);
    if (V == 3) bar_v3(sync, target, n);
    if (V == 4) bar_v4(sync, target, n);
    if (V == 5) bar_v5(sync, target, n);
    if (V == 6) bar_v6(sync, target, n, write_floats > 0 && threadIdx.x < write_floats);
    if (V == 7) bar_v7(sync, target, n);
    if (write_floats > 0) {
      for (int i = threadIdx.x; i < payload_floats / 4; i += THREADS) cp_async16(&stage[i], buf + 4 * i);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      const float4 v = stage[threadIdx.x % (payload_floats / 4)];
      // check: every element written by an in-range CTA must carry this iteration's value
      if ((int)(threadIdx.x % (payload_floats / 4)) * 4 < (int)n * write_floats && v.x != (float)it) acc = -1.f;
      if (acc >= 0.f) acc += 1.f;
    }
  }
  if (threadIdx.x == 0) cycles[blockIdx.x] = (clock64() - t0) / iters;
  if (acc == 12345.678f) *sink = acc;
  if (__syncthreads_or(acc < 0.f) && threadIdx.x == 0) cycles[blockIdx.x] = -1;      // stale data seen
}

template <int V>
void run(const char* name, int sms, int payload, int wr) {
  unsigned* sync;
  float *data, *sink;
  long long* cyc;
  CK(cudaMalloc(&sync, 4096));
  CK(cudaMemset(sync, 0, 4096));
  CK(cudaMalloc(&data, 2 * 65536 * 4));
  CK(cudaMemset(data, 0, 2 * 65536 * 4));
  CK(cudaMalloc(&sink, 4));
  CK(cudaMalloc(&cyc, 8 * 256));
  int iters = 2000;
  int pl = payload, w = wr;
  void* args[] = {&sync, &data, &iters, &pl, &w, &cyc, &sink};
  size_t smem = 160 * 1024;
  CK(cudaFuncSetAttribute(bench<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  CK(cudaLaunchCooperativeKernel((void*)bench<V>, dim3(sms), dim3(THREADS), args, smem, 0));
  CK(cudaDeviceSynchronize());
  CK(cudaMemset(sync, 0, 4096));
  cudaEventRecord(a);
  CK(cudaLaunchCooperativeKernel((void*)bench<V>, dim3(sms), dim3(THREADS), args, smem, 0));
  cudaEventRecord(b);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  long long h[256];
  CK(cudaMemcpy(h, cyc, 8 * sms, cudaMemcpyDeviceToHost));
  bool stale = false;
  for (int i = 0; i < sms; ++i) stale |= h[i] < 0;
  printf("%-34s payload %5d floats: %7.3f us / hand-over, %6lld cycles (CTA 0)%s\n", name, wr ? payload : 0,
         ms * 1e3 / (iters + 4), h[0], stale ? "  STALE DATA" : "");
  fflush(stdout);
  cudaFree(sync); cudaFree(data); cudaFree(sink); cudaFree(cyc);
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  printf("%d SMs, %d threads per CTA\n", sms, THREADS);
  {
    unsigned long long* slots;
    long long* cyc;
    CK(cudaMalloc(&slots, 4096));
    CK(cudaMemset(slots, 0, 4096));
    CK(cudaMalloc(&cyc, 64));
    pingpong<<<2, 32>>>(slots, 2000, cyc);
    CK(cudaDeviceSynchronize());
    long long h;
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("ping-pong of a tagged 8-byte word between two CTAs: %lld cycles per round trip\n", h);
    fflush(stdout);
  }
  for (int pass = 0; pass < 3; ++pass) {
    const int payload = pass == 0 ? 0 : pass == 1 ? 8 * 1200 : 32 * 1200;
    const int wr = pass == 0 ? 0 : (payload + sms - 1) / sms;
    run<0>("red.release + ld.acquire spin", sms, payload ? payload : 1024, wr);
    run<1>("fence + relaxed red/ld + fence", sms, payload ? payload : 1024, wr);
    run<2>("flag per CTA, warp polls", sms, payload ? payload : 1024, wr);
    run<3>("per-thread fence, bar.arrive split", sms, payload ? payload : 1024, wr);
    run<4>("no fences (floor, unsafe)", sms, payload ? payload : 1024, wr);
    run<5>("red.release + relaxed spin", sms, payload ? payload : 1024, wr);
    run<6>("writers fence, relaxed red/spin", sms, payload ? payload : 1024, wr);
    run<7>("atomicAdd + volatile spin (unsafe)", sms, payload ? payload : 1024, wr);
  }
  return 0;
}
