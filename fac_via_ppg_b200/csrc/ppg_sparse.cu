// Pruned-posteriorgram input of the PPG->Mel encoder (the step before the path is
// reference src/common/data_utils.py:55-59 get_ppg -> src/ppg/compute_ppg.py:186-202).
//
// A phonetic posteriorgram frame is a softmax over 5816 senones and is almost entirely tail: the first prenet
// layer (reference src/common/model.py:124-135, a bias-free Linear 5816 -> 600) is then a sum of a few dozen
// weight columns, yet the dense form streams 23 KB per frame and runs a K = 5816 GEMM.  Two kernels:
//   ppg_sparsify_kernel   dense channel-major (B, D, T) -> per frame the entries above a threshold as
//                         (index, value) lists of fixed capacity k, in ascending channel order (deterministic);
//                         a frame with more than k survivors is COUNTED, never silently truncated
//   prenet0_sparse_kernel out[b, t, :] = relu(sum_j val_j * W0^T[idx_j, :]) * dropout mask, exact fp32 FMA in list
//                         order, written as fp32 and/or as the fp16 hi/lo operand copies of the next tensor-core GEMM
// The lists can also come straight from the host (k * 8 bytes per frame instead of 23 KB over PCIe).
#include "fac_common.cuh"
#include <cuda_fp16.h>
#include <stdint.h>

namespace fac {
namespace {

constexpr int SP_WARPS = 8;          // channel slices per CTA
constexpr int SP_THREADS = 32 * SP_WARPS;
constexpr int SP_KMAX = 64;

// CTA = 32 consecutive frames of one utterance (lane = frame: coalesced 128-byte rows of the channel-major
// input) x 8 channel slices (warp).  Survivors are buffered per thread in shared memory, then the 8 slices of a
// frame are concatenated in slice order.
__global__ void __launch_bounds__(SP_THREADS) ppg_sparsify_kernel(const float* __restrict__ ppg, int* __restrict__ idx,
                                                                  float* __restrict__ val, int* __restrict__ overflow,
                                                                  int D, int T, int k, float threshold) {
  extern __shared__ unsigned char sp_smem[];
  int* buf_i = reinterpret_cast<int*>(sp_smem);                         // [warp][k][lane]
  float* buf_v = reinterpret_cast<float*>(buf_i + SP_WARPS * k * 32);
  __shared__ int counts[SP_WARPS][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y, t = blockIdx.x * 32 + lane;
  const int per = (D + SP_WARPS - 1) / SP_WARPS, d0 = warp * per, d1 = min(D, d0 + per);
  const float* src = ppg + (long long)b * D * T + t;
  int n = 0;
  if (t < T) {
#pragma unroll 8
    for (int d = d0; d < d1; ++d) {
      const float v = __ldg(src + (long long)d * T);
      if (v > threshold) {
        if (n < k) {
          buf_i[(warp * k + n) * 32 + lane] = d;
          buf_v[(warp * k + n) * 32 + lane] = v;
        }
        ++n;                         // keeps counting: the total tells whether the frame fits
      }
    }
  }
  counts[warp][lane] = n;
  __syncthreads();
  if (t >= T) return;
  int before = 0, total = 0;
  for (int w = 0; w < SP_WARPS; ++w) {
    const int c = counts[w][lane];
    if (w < warp) before += c;
    total += c;
  }
  const long long row = ((long long)b * T + t) * k;
  const int mine = min(n, k);
  for (int j = 0; j < mine && before + j < k; ++j) {
    idx[row + before + j] = buf_i[(warp * k + j) * 32 + lane];
    val[row + before + j] = buf_v[(warp * k + j) * 32 + lane];
  }
  if (warp == SP_WARPS - 1) {
    for (int j = min(total, k); j < k; ++j) {         // padding: index 0 with weight 0 contributes nothing
      idx[row + j] = 0;
      val[row + j] = 0.f;
    }
    if (total > k) atomicAdd(overflow, 1);
  }
}

// One CTA per frame group: thread = 4 consecutive output channels (float4 of a weight row), loop over the list.
constexpr int P0_FRAMES = 4;
__global__ void __launch_bounds__(160) prenet0_sparse_kernel(const int* __restrict__ idx, const float* __restrict__ val,
                                                             const float* __restrict__ w_t, const float* __restrict__ mask,
                                                             const int* __restrict__ row_lengths, float* __restrict__ out,
                                                             __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                                                             long long n_rows, int T, int k, int D, int E, int w_ld,
                                                             int out_ld, int pad) {
  __shared__ int s_idx[P0_FRAMES][SP_KMAX];
  __shared__ float s_val[P0_FRAMES][SP_KMAX];
  const long long r0 = (long long)blockIdx.x * P0_FRAMES;
  for (int i = threadIdx.x; i < P0_FRAMES * k; i += blockDim.x) {
    const int f = i / k, j = i - f * k;
    const long long r = r0 + f;
    int id = 0;
    float v = 0.f;
    if (r < n_rows) {
      id = __ldg(idx + r * k + j);
      v = __ldg(val + r * k + j);
      if (id < 0 || id >= D) {       // a corrupt list must not read outside the weight matrix
        id = 0;
        v = 0.f;
      }
    }
    s_idx[f][j] = id;
    s_val[f][j] = v;
  }
  __syncthreads();
  const int c = threadIdx.x * 4;
  if (c >= pad) return;
  for (int f = 0; f < P0_FRAMES; ++f) {
    const long long r = r0 + f;
    if (r >= n_rows) break;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < E) {
      for (int j = 0; j < k; ++j) {
        const float v = s_val[f][j];
        if (v == 0.f) continue;      // padding entries
        const float4 w = __ldg(reinterpret_cast<const float4*>(w_t + (long long)s_idx[f][j] * w_ld + c));
        acc.x = fmaf(v, w.x, acc.x);
        acc.y = fmaf(v, w.y, acc.y);
        acc.z = fmaf(v, w.z, acc.z);
        acc.w = fmaf(v, w.w, acc.w);
      }
      const float4 m = mask ? __ldg(reinterpret_cast<const float4*>(mask + r * E + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
      acc.x = fmaxf(acc.x, 0.f) * m.x;
      acc.y = fmaxf(acc.y, 0.f) * m.y;
      acc.z = fmaxf(acc.z, 0.f) * m.z;
      acc.w = fmaxf(acc.w, 0.f) * m.w;
      if (row_lengths != nullptr && (int)(r % T) >= __ldg(row_lengths + (int)(r / T))) acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (out) *reinterpret_cast<float4*>(out + r * out_ld + c) = acc;
    }
    if (out_hi) {                    // operand copies of the next GEMM; channels [E, pad) are exact zeros
      const float a[4] = {acc.x, acc.y, acc.z, acc.w};
      __half h[4], l[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float v = fminf(fmaxf(a[i], -65504.f), 65504.f);
        h[i] = __float2half_rn(v);
        l[i] = __float2half_rn(v - __half2float(h[i]));
      }
      *reinterpret_cast<uint2*>(out_hi + r * pad + c) = *reinterpret_cast<const uint2*>(h);
      if (out_lo) *reinterpret_cast<uint2*>(out_lo + r * pad + c) = *reinterpret_cast<const uint2*>(l);
    }
  }
}

}  // namespace

int ppg_sparsify(const float* ppg, int* idx, float* val, int* overflow, int B, int D, int T, int k, float threshold,
                 cudaStream_t st) {
  FAC_REQUIRE(ppg && idx && val && overflow, "ppg_sparsify: NULL argument");
  FAC_REQUIRE(B > 0 && D > 0 && T > 0, "ppg_sparsify: empty problem");
  FAC_REQUIRE(k > 0 && k <= SP_KMAX, "ppg_sparsify: list capacity k must be in [1, %d] (got %d)", SP_KMAX, k);
  const size_t smem = (size_t)SP_WARPS * k * 32 * 8;
  cudaError_t e = cudaFuncSetAttribute(ppg_sparsify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("ppg_sparsify: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return 2;
  }
  dim3 grid((unsigned)ceil_div(T, 32), (unsigned)B);
  ppg_sparsify_kernel<<<grid, SP_THREADS, smem, st>>>(ppg, idx, val, overflow, D, T, k, threshold);
  count_launch();
  return check_launch("ppg_sparsify_kernel");
}

int prenet0_sparse(const int* idx, const float* val, const float* w_t, int w_ld, const float* mask,
                   const int* row_lengths, float* out, int out_ld, void* out_hi, void* out_lo, int B, int T, int k, int D,
                   int E, int pad, cudaStream_t st) {
  FAC_REQUIRE(idx && val && w_t && (out || out_hi), "prenet0_sparse: NULL argument");
  FAC_REQUIRE(B > 0 && T > 0 && k > 0 && k <= SP_KMAX, "prenet0_sparse: bad sizes (k must be in [1, %d])", SP_KMAX);
  FAC_REQUIRE(E % 4 == 0 && pad % 4 == 0 && pad >= E && pad <= 640 && w_ld % 4 == 0 && w_ld >= E,
              "prenet0_sparse: E %d / pad %d / w_ld %d must be multiples of 4 with E <= pad <= 640", E, pad, w_ld);
  FAC_REQUIRE(out == nullptr || (out_ld % 4 == 0 && out_ld >= E), "prenet0_sparse: out_ld");
  const long long n_rows = (long long)B * T;
  prenet0_sparse_kernel<<<(unsigned)((n_rows + P0_FRAMES - 1) / P0_FRAMES), 160, 0, st>>>(
      idx, val, w_t, mask, row_lengths, out, reinterpret_cast<__half*>(out_hi), reinterpret_cast<__half*>(out_lo), n_rows,
      T, k, D, E, w_ld, out_ld, pad);
  count_launch();
  return check_launch("prenet0_sparse_kernel");
}

}  // namespace fac
