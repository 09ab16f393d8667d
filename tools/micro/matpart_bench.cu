// Where do the cycles of one decoder mat-vec piece go?  Replica of mat_part()'s inner loop (ldmatrix of the hi/lo
// weight fragments, fp32 inputs split into half pairs, 3 mma.sync per step) with the steps timed one by one.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/matpart_bench.cu -o /tmp/mp && /tmp/mp
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
constexpr int KS = 1224, XSTRIDE = 1224;
struct Smem {
  __half w[2][12][KS];
  float xs[16][XSTRIDE];
  float part[16][16][16];
};
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split_half2(float2 x, uint32_t& hi, uint32_t& lo) {
  const float hx = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
  const float hy = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
  const __half2 h = __floats2half2_rn(hx, hy);
  const __half2 l = __floats2half2_rn(x.x - hx, x.y - hy);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// ~JUNK_KB KB of straight-line code nobody else executes: evicts the loop from the instruction caches
template <int N>
__device__ __forceinline__ float junk(float x) {
#pragma unroll
  for (int i = 0; i < N; ++i) x = fmaf(x, 1.0001f + 1e-6f * i, 0.5f + i);
  return x;
}
template <int MODE>   // 0 full, 1 no split (raw bits), 2 no ldmatrix, 3 no mma, 4 full + junk between repetitions
__global__ void __launch_bounds__(512, 1) k(long long* out, float* sink, int total, int n_warps, int reps) {
  extern __shared__ __align__(16) unsigned char raw[];
  Smem& sm = *reinterpret_cast<Smem*>(raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < (int)(sizeof(Smem) / 4); i += 512) reinterpret_cast<float*>(raw)[i] = 0.001f * (i & 63);
  __syncthreads();
  long long t_mma = 0, t_red = 0;
  float res = 0.f;
  for (int rep = 0; rep < reps; ++rep) {
    if (MODE >= 4) {
      res = MODE == 4 ? junk<2048>(res) : MODE == 5 ? junk<8192>(res) : junk<16384>(res);
      __syncthreads();
    }
    const long long t0 = clock64();
    if (warp < n_warps) {
      int arow = (lane & 7) + ((lane >> 3) & 1) * 8;
      if (arow >= 12) arow = 0;
      const uint32_t a_off = (uint32_t)(arow * KS + (lane >> 4) * 8) * 2;
      const uint32_t a_hi = (uint32_t)__cvta_generic_to_shared(&sm.w[0][0][0]) + a_off;
      const uint32_t a_lo = (uint32_t)__cvta_generic_to_shared(&sm.w[1][0][0]) + a_off;
      const float* xrow0 = &sm.xs[lane >> 2][2 * (lane & 3)];
      float acc[3][4] = {};
#pragma unroll
      for (int it = 0; it < 5; ++it) {
        const int ls = warp + it * n_warps;
        if (ls < total) {
          const int koff = 16 * ls, xo = 16 * ls;
          uint32_t ah[4] = {1, 2, 3, 4}, al[4] = {5, 6, 7, 8}, bh[2], bl[2];
          if (MODE != 2 || MODE >= 4) {
            ldmatrix_x4(ah, a_hi + koff * 2);
            ldmatrix_x4(al, a_lo + koff * 2);
          }
          const float2 x0 = *reinterpret_cast<const float2*>(xrow0 + xo), x1 = *reinterpret_cast<const float2*>(xrow0 + xo + 8);
          if (MODE != 1 || MODE >= 4) {
            split_half2(x0, bh[0], bl[0]);
            split_half2(x1, bh[1], bl[1]);
          } else {
            bh[0] = __float_as_uint(x0.x); bl[0] = __float_as_uint(x0.y); bh[1] = __float_as_uint(x1.x); bl[1] = __float_as_uint(x1.y);
          }
          if (MODE != 3 || MODE >= 4) {
            mma_f16(acc[0], ah, bh);
            mma_f16(acc[1], al, bh);
            mma_f16(acc[2], ah, bl);
          } else {
            acc[0][0] += __uint_as_float(ah[0] ^ bh[0] ^ al[1] ^ bl[1] ^ ah[2] ^ al[3] ^ bh[1] ^ bl[0]);
          }
        }
      }
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = acc[0][j] + (acc[1][j] + acc[2][j]);
      *reinterpret_cast<float2*>(&sm.part[warp][lane >> 2][2 * (lane & 3)]) = make_float2(v[0], v[1]);
      *reinterpret_cast<float2*>(&sm.part[warp][(lane >> 2) + 8][2 * (lane & 3)]) = make_float2(v[2], v[3]);
    }
    const long long t1 = clock64();
    __syncthreads();
    float a = 0.f;
    if (tid < 256 && (tid & 15) < 8 && (tid >> 4) < 12)
      for (int w = 0; w < n_warps; ++w) a += sm.part[w][tid >> 4][tid & 15];
    res += a;
    __syncthreads();
    const long long t2 = clock64();
    if (rep > 0) { t_mma += t1 - t0; t_red += t2 - t1; }
  }
  if (res == 1234.5f) *sink = res;
  if (tid == 0) { out[0] = t_mma / (reps - 1); out[1] = t_red / (reps - 1); }
}
int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 16); cudaMalloc(&sink, 4);
  const char* names[7] = {"full", "no split", "no ldmatrix", "no mma", "full, 32 KB of other code between", "full, 128 KB between", "full, 256 KB between"};
  auto run = [&](auto kern, int mode, int total, int nw) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    kern<<<1, 512, sizeof(Smem)>>>(out, sink, total, nw, 50);
    long long h[2];
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%-36s %2d steps on %2d warps: products %5lld cycles (warp 0), barrier + reduction %5lld\n", names[mode], total, nw, h[0], h[1]);
  };
  for (int total : {19, 38})
    for (int nw : {4, 8, 16}) {
      run(k<0>, 0, total, nw);
      run(k<1>, 1, total, nw);
      run(k<2>, 2, total, nw);
      run(k<3>, 3, total, nw);
      run(k<4>, 4, total, nw);
      run(k<5>, 5, total, nw);
      run(k<6>, 6, total, nw);
    }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
