// Generic fp32 (FFMA) implicit-GEMM Conv1d / Linear for channels-last activations.
//
// out[b, t, n] = epilogue( sum_{src, tap, c} in_src[b, t + tap*dil - center, c] * W[(src,tap,c), n] + bias[n] )
//
// This is the exact-fp32 anchor of the path: every Conv1d / Linear the reference
// runs through cuDNN/cuBLAS (reference src/common/layers.py:40-71 wrappers, the WN
// convolutions src/waveglow/glow.py:156-164, the upsampler glow.py:253) maps onto
// this one kernel with a different source list and epilogue.  128x128x8 tiles,
// 256 threads, 8x8 register micro-tiles, double-buffered shared memory.
#include "fac_common.cuh"

namespace fac {
namespace {

constexpr int BM = 128, BN = 128, BK = 8, NTHREADS = 256, APAD = 4;

struct SrcDev {
  const float* ptr;
  long long bs, rs, cs;
  int C, taps, dil, center, rows;
};

struct GemmParams {
  SrcDev src[2];
  int n_src;
  const float* w;
  const float* bias;
  int T, N, N_pad, tiles_per_batch, total_kblocks;
  long long w_zs, out_zs;
  fac_conv_epilogue epi;
};

struct KIter {
  int s, tap, c0;
};

__device__ __forceinline__ void advance(KIter& it, const GemmParams& p) {
  it.c0 += BK;
  if (it.c0 >= p.src[it.s].C) {
    it.c0 = 0;
    if (++it.tap >= p.src[it.s].taps) {
      it.tap = 0;
      ++it.s;
    }
  }
}

// Stage one 128 x 8 activation tile into registers (one float4 per thread).
__device__ __forceinline__ float4 load_a(const GemmParams& p, const KIter& it, int b, int t0, int tid) {
  const SrcDev& s = p.src[it.s];
  const int shift = it.tap * s.dil - s.center;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s.cs == 1) {  // channels-last: 4 consecutive channels of one row
    const int m = tid >> 1, kq = tid & 1;
    const int r = t0 + m + shift;
    if (r >= 0 && r < s.rows)
      v = __ldg(reinterpret_cast<const float4*>(s.ptr + b * s.bs + (long long)r * s.rs + it.c0 + kq * 4));
  } else {  // channel-major (B, C, T) input: 4 consecutive rows of one channel
    const int k = tid >> 5, m4 = (tid & 31) * 4;
    const float* base = s.ptr + b * s.bs + (long long)(it.c0 + k) * s.cs;
    const int r = t0 + m4 + shift;
    float e[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rr = r + j;
      e[j] = (rr >= 0 && rr < s.rows) ? __ldg(base + (long long)rr * s.rs) : 0.f;
    }
    v = make_float4(e[0], e[1], e[2], e[3]);
  }
  return v;
}

__device__ __forceinline__ void store_a(float (*As)[BM + APAD], const GemmParams& p, const KIter& it, int tid,
                                        const float4& v) {
  if (p.src[it.s].cs == 1) {
    const int m = tid >> 1, kq = tid & 1;
    As[kq * 4 + 0][m] = v.x;
    As[kq * 4 + 1][m] = v.y;
    As[kq * 4 + 2][m] = v.z;
    As[kq * 4 + 3][m] = v.w;
  } else {
    const int k = tid >> 5, m4 = (tid & 31) * 4;
    *reinterpret_cast<float4*>(&As[k][m4]) = v;
  }
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == FAC_ACT_RELU) return fmaxf(v, 0.f);
  if (act == FAC_ACT_TANH) return tanhf(v);
  return v;
}

__global__ void __launch_bounds__(NTHREADS, 2) conv_gemm_f32_kernel(const GemmParams p) {
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int b = blockIdx.x / p.tiles_per_batch;
  const int t0 = (blockIdx.x % p.tiles_per_batch) * BM;
  const int n0 = blockIdx.y * BN;
  const float* __restrict__ w = p.w + blockIdx.z * p.w_zs;

  const int ty = tid >> 4, tx = tid & 15;
  const int bk = tid >> 5, bn4 = (tid & 31) * 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  KIter it{0, 0, 0};
  float4 ra = load_a(p, it, b, t0, tid);
  float4 rb = __ldg(reinterpret_cast<const float4*>(w + (long long)bk * p.N_pad + n0 + bn4));
  store_a(As[0], p, it, tid, ra);
  *reinterpret_cast<float4*>(&Bs[0][bk][bn4]) = rb;
  __syncthreads();

  int cur = 0;
  for (int kb = 0; kb < p.total_kblocks; ++kb) {
    const bool has_next = kb + 1 < p.total_kblocks;
    KIter nit = it;
    if (has_next) {
      advance(nit, p);
      ra = load_a(p, nit, b, t0, tid);
      rb = __ldg(reinterpret_cast<const float4*>(w + (long long)((kb + 1) * BK + bk) * p.N_pad + n0 + bn4));
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (has_next) {
      store_a(As[cur ^ 1], p, nit, tid, ra);
      *reinterpret_cast<float4*>(&Bs[cur ^ 1][bk][bn4]) = rb;
      it = nit;
    }
    __syncthreads();
    cur ^= 1;
  }

  // ---- epilogue -----------------------------------------------------------
  const fac_conv_epilogue& e = p.epi;
  const long long zoff = blockIdx.z * p.out_zs;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4));
    const int t = t0 + m;
    if (t >= p.T) continue;
    const bool dead_row = e.kind == FAC_EPI_LINEAR && e.row_lengths != nullptr && t >= __ldg(e.row_lengths + b);
    const long long row = zoff + b * e.out_batch_stride + (long long)t * e.out_row_stride;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + h * 64 + tx * 4;
      if (n >= p.N) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = acc[i][h * 4 + j] + (p.bias ? __ldg(p.bias + n + j) : 0.f);

      if (e.kind == FAC_EPI_GATE) {
        // columns (n, n+1) and (n+2, n+3) are (tanh, sigmoid) pairs of channels n/2, n/2+1
        float2 o;
        o.x = tanhf(v[0]) * sigmoidf_exact(v[1]);
        o.y = tanhf(v[2]) * sigmoidf_exact(v[3]);
        *reinterpret_cast<float2*>(e.out + row + (n >> 1)) = o;
      } else if (e.kind == FAC_EPI_RES_SKIP) {
        if (n < e.n_split) {
          float4* dst = reinterpret_cast<float4*>(e.out + row + n);
          float4 o = *dst;
          o.x += v[0]; o.y += v[1]; o.z += v[2]; o.w += v[3];
          *dst = o;
        } else {
          float4* dst = reinterpret_cast<float4*>(e.out2 + row + (n - e.n_split));
          float4 o = make_float4(v[0], v[1], v[2], v[3]);
          if (e.accumulate_out2) {
            const float4 old = *dst;
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
          }
          *dst = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (n + j < p.N) {
            float o = apply_act(v[j], e.act);
            if (e.mask) o *= __ldg(e.mask + row + n + j);
            if (e.residual) o += __ldg(e.residual + row + n + j);
            e.out[row + n + j] = dead_row ? 0.f : o;
          }
        }
      }
    }
  }
}

}  // namespace

int launch_conv_gemm_f32(const fac_conv_src* srcs, int n_srcs, const float* w_packed, const float* bias, int B,
                         int T_out, int N, const fac_conv_epilogue* epi, int n_phases, long long w_phase_stride,
                         long long out_phase_stride, cudaStream_t stream) {
  FAC_REQUIRE(n_srcs >= 1 && n_srcs <= 2, "conv_gemm: n_srcs must be 1 or 2 (got %d)", n_srcs);
  FAC_REQUIRE(B > 0 && T_out > 0 && N > 0, "conv_gemm: empty problem B=%d T=%d N=%d", B, T_out, N);
  FAC_REQUIRE(epi && epi->out, "conv_gemm: missing epilogue/output");
  GemmParams p{};
  p.n_src = n_srcs;
  int kblocks = 0;
  for (int s = 0; s < n_srcs; ++s) {
    const fac_conv_src& in = srcs[s];
    FAC_REQUIRE(in.ptr != nullptr, "conv_gemm: source %d is NULL", s);
    FAC_REQUIRE(in.channels > 0 && in.channels % BK == 0, "conv_gemm: source %d channels %d not a multiple of %d", s,
                in.channels, BK);
    FAC_REQUIRE(in.taps >= 1, "conv_gemm: source %d has %d taps", s, in.taps);
    if (in.ch_stride == 1)
      FAC_REQUIRE(in.row_stride % 4 == 0 && in.batch_stride % 4 == 0 && ((size_t)in.ptr & 15) == 0,
                  "conv_gemm: channels-last source %d must be 16-byte aligned per row", s);
    p.src[s] = SrcDev{in.ptr, in.batch_stride, in.row_stride, in.ch_stride, in.channels, in.taps, in.dilation,
                      in.center, in.rows};
    kblocks += in.taps * in.channels / BK;
  }
  p.w = w_packed;
  p.bias = bias;
  p.T = T_out;
  p.N = N;
  p.N_pad = round_up(N, BN);
  p.tiles_per_batch = ceil_div(T_out, BM);
  p.total_kblocks = kblocks;
  p.w_zs = w_phase_stride;
  p.out_zs = out_phase_stride;
  p.epi = *epi;
  if (epi->kind == FAC_EPI_GATE)
    FAC_REQUIRE(N % 4 == 0 && epi->out_row_stride % 2 == 0, "conv_gemm: gate epilogue needs N %% 4 == 0");
  if (epi->kind == FAC_EPI_RES_SKIP)
    FAC_REQUIRE(N % 4 == 0 && epi->n_split % 4 == 0 && epi->out_row_stride % 4 == 0 && epi->out2 != nullptr,
                "conv_gemm: res/skip epilogue needs 4-aligned widths and out2");
  dim3 grid(B * p.tiles_per_batch, p.N_pad / BN, n_phases > 0 ? n_phases : 1);
  conv_gemm_f32_kernel<<<grid, NTHREADS, 0, stream>>>(p);
  count_launch();
  return check_launch("conv_gemm_f32_kernel");
}

}  // namespace fac
