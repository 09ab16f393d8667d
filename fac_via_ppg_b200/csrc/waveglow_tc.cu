// WaveGlow WN layer on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), reference
// src/waveglow/glow.py:158-174.
//
// Each layer is two implicit GEMMs over 128-column time tiles:
//   G1: pre[128 x 2C]  = [x(t-d) | x(t) | x(t+d) | spect(t)] (K = 3C + n_cond) * W1^T, fused gate
//                         tanh(.)*sigmoid(.) epilogue  -> acts
//   G2: rs [128 x 2C|C] = acts (K = C) * W2^T, fused residual / skip epilogue -> x, skip
// Operands are bf16 in channels-last HBM layout and reach shared memory through TMA (the
// dilated taps are just shifted box coordinates; rows outside [0, T) are zero-filled by the
// TMA unit, which is exactly Conv1d's zero padding).  Accumulation is fp32 in tensor memory:
// one 128 x 512 tile fills the 512 TMEM columns.  Precision modes:
//   nsplit = 1  plain bf16 operands                                  (1 UMMA per product)
//   nsplit = 2  split-bf16: v = hi + lo, a*w ~= ah*wh + al*wh + ah*wl (3 UMMAs per product),
//               ~2^-16 relative operand error, fp32 accumulate: matches the fp32 reference to
//               ~1e-5 RMS on the waveform (tests/test_waveglow_tc_gpu.py).
// Warp roles (192 threads, persistent over tiles, 1 CTA / SM): warp 0 = TMA producer,
// warp 1 = UMMA issuer (one elected lane), warps 2-5 = epilogue (one TMEM lane quarter each).
#include "fac_common.cuh"
#include "tc_common.cuh"

namespace fac {
namespace {

using namespace tc;

constexpr int TC_BM = 128;       // time rows per tile (UMMA M)
constexpr int TC_BK = 16;        // K per pipeline stage = one UMMA K step; 32-byte rows, SWIZZLE_32B
constexpr int TC_ROWB = TC_BK * 2;
constexpr int TC_STAGES = 5;
constexpr int TC_NMAX = 512;     // accumulator columns per tile (all of TMEM)
constexpr int TC_NHALF = 256;    // N of one UMMA
constexpr int TC_THREADS = 192;
constexpr int A_BYTES = TC_BM * TC_ROWB;       // 4 KB
constexpr int W_BYTES = TC_NMAX * TC_ROWB;     // 16 KB (two halves of 8 KB)
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;   // hi + lo of both operands = 40 KB
constexpr int TC_SMEM = TC_STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;

enum { TC_GATE = 0, TC_RES_SKIP = 1 };

struct TcSrc {
  int channels, taps, dilation, center;
};

struct TcParams {
  int n_src;
  TcSrc src[2];
  int n_total, k_steps, T, B, tiles_per_batch, n_tiles, nsplit, mode, C;
  const float* bias;
  __nv_bfloat16* acts_hi;
  __nv_bfloat16* acts_lo;
  float* x;
  __nv_bfloat16* x_hi;
  __nv_bfloat16* x_lo;
  float* skip;
  int n_split_cols, accumulate_skip;
};

__device__ __forceinline__ float ex2_approx(float v) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
// tanh(a) * sigmoid(b) = (e^{2a} - 1) / ((e^{2a} + 1) (1 + e^{-b})), ~3e-7 absolute error.
__device__ __forceinline__ float gate_act(float a, float b) {
  a = fminf(fmaxf(a, -15.f), 15.f);
  const float ea = ex2_approx(a * 2.8853900817779268f);
  const float eb = ex2_approx(b * -1.4426950408889634f);
  return __fdividef(ea - 1.f, (ea + 1.f) * (1.f + eb));
}

__global__ void __launch_bounds__(TC_THREADS, 1)
wn_gemm_tc_kernel(const __grid_constant__ CUtensorMap a0_hi, const __grid_constant__ CUtensorMap a0_lo,
                  const __grid_constant__ CUtensorMap a1_hi, const __grid_constant__ CUtensorMap a1_lo,
                  const __grid_constant__ CUtensorMap w_hi, const __grid_constant__ CUtensorMap w_lo,
                  const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + TC_STAGES * STAGE_BYTES);
  uint64_t* empty = full + TC_STAGES;
  uint64_t* tmem_full = empty + TC_STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_halves = (p.n_total + TC_NHALF - 1) / TC_NHALF;
  const int n_half_cols = p.n_total < TC_NHALF ? p.n_total : TC_NHALF;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 4);
    fence_barrier_init();
    tma_prefetch_desc(&a0_hi);
    tma_prefetch_desc(&w_hi);
  }
  if (warp == 1) tmem_alloc(tmem_slot, TC_NMAX);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      const uint32_t stage_bytes =
          (uint32_t)(p.nsplit * (A_BYTES + n_halves * n_half_cols * TC_ROWB));
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int b = tile / p.tiles_per_batch;
        const int t0 = (tile % p.tiles_per_batch) * TC_BM;
        int ks = 0;
        for (int s = 0; s < p.n_src; ++s) {
          const CUtensorMap* mh = s == 0 ? &a0_hi : &a1_hi;
          const CUtensorMap* ml = s == 0 ? &a0_lo : &a1_lo;
          for (int tap = 0; tap < p.src[s].taps; ++tap) {
            const int row0 = t0 + tap * p.src[s].dilation - p.src[s].center;
            for (int c0 = 0; c0 < p.src[s].channels; c0 += TC_BK, ++ks) {
              mbar_wait(&empty[stage], phase ^ 1);
              uint8_t* st = smem + stage * STAGE_BYTES;
              mbar_arrive_expect_tx(&full[stage], stage_bytes);
              tma_load_3d(st, mh, &full[stage], c0, row0, b);
              if (p.nsplit == 2) tma_load_3d(st + A_BYTES, ml, &full[stage], c0, row0, b);
              for (int h = 0; h < n_halves; ++h) {
                tma_load_2d(st + 2 * A_BYTES + h * (W_BYTES / 2), &w_hi, &full[stage], ks * TC_BK, h * TC_NHALF);
                if (p.nsplit == 2)
                  tma_load_2d(st + 2 * A_BYTES + W_BYTES + h * (W_BYTES / 2), &w_lo, &full[stage], ks * TC_BK,
                              h * TC_NHALF);
              }
              if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== UMMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(TC_BM, n_half_cols);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        mbar_wait(tmem_empty, (uint32_t)((it & 1) ^ 1));   // epilogue has drained the previous tile
        tc_fence_after();
        for (int ks = 0; ks < p.k_steps; ++ks) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t a_h = make_smem_desc(st, TC_ROWB);
          const uint64_t a_l = make_smem_desc(st + A_BYTES, TC_ROWB);
          for (int h = 0; h < n_halves; ++h) {
            const uint32_t d = tmem_base + h * TC_NHALF;
            const uint64_t w_h = make_smem_desc(st + 2 * A_BYTES + h * (W_BYTES / 2), TC_ROWB);
            umma_bf16(d, a_h, w_h, idesc, ks > 0 ? 1u : 0u);
            if (p.nsplit == 2) {
              const uint64_t w_l = make_smem_desc(st + 2 * A_BYTES + W_BYTES + h * (W_BYTES / 2), TC_ROWB);
              umma_bf16(d, a_l, w_h, idesc, 1u);
              umma_bf16(d, a_h, w_l, idesc, 1u);
            }
          }
          umma_commit(&empty[stage]);   // frees the smem slot when these UMMAs have read it
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tmem_full);         // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================================================== epilogue (4 warps, one TMEM lane quarter each)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int b = tile / p.tiles_per_batch;
      const int t = (tile % p.tiles_per_batch) * TC_BM + row;
      const bool valid = t < p.T;
      const long long col = (long long)b * p.T + t;
      mbar_wait(tmem_full, (uint32_t)(it & 1));
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int n0 = 0; n0 < p.n_total; n0 += 32) {
        uint32_t r[32];
        tmem_ld32(trow + n0, r);
        tmem_ld_wait();
        if (!valid) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
          v[j + 0] = __uint_as_float(r[j + 0]) + bv.x;
          v[j + 1] = __uint_as_float(r[j + 1]) + bv.y;
          v[j + 2] = __uint_as_float(r[j + 2]) + bv.z;
          v[j + 3] = __uint_as_float(r[j + 3]) + bv.w;
        }
        if (p.mode == TC_GATE) {
          // columns (2c, 2c+1) hold the tanh / sigmoid pre-activations of channel c (glow.py:33-40)
          float g[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) g[j] = gate_act(v[2 * j], v[2 * j + 1]);
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) split2(g[2 * j], g[2 * j + 1], hi[j], lo[j]);
          const long long off = col * p.C + (n0 >> 1);
          uint4* dh = reinterpret_cast<uint4*>(p.acts_hi + off);
          dh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          dh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
          if (p.nsplit == 2) {
            uint4* dl = reinterpret_cast<uint4*>(p.acts_lo + off);
            dl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            dl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
          }
        } else if (n0 < p.n_split_cols) {
          // residual stream: x += rs[:, :C] (glow.py:166) + refresh the bf16 operand copies
          const long long off = col * p.C + n0;
          float4* xp = reinterpret_cast<float4*>(p.x + off);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o = xp[j];
            o.x += v[4 * j + 0]; o.y += v[4 * j + 1]; o.z += v[4 * j + 2]; o.w += v[4 * j + 3];
            xp[j] = o;
            v[4 * j + 0] = o.x; v[4 * j + 1] = o.y; v[4 * j + 2] = o.z; v[4 * j + 3] = o.w;
          }
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
          uint4* dh = reinterpret_cast<uint4*>(p.x_hi + off);
#pragma unroll
          for (int j = 0; j < 4; ++j) dh[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
          if (p.nsplit == 2) {
            uint4* dl = reinterpret_cast<uint4*>(p.x_lo + off);
#pragma unroll
            for (int j = 0; j < 4; ++j) dl[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
          }
        } else {
          // skip sum (glow.py:167-174)
          float4* sp = reinterpret_cast<float4*>(p.skip + col * p.C + (n0 - p.n_split_cols));
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o = make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            if (p.accumulate_skip) {
              const float4 old = sp[j];
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            sp[j] = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, TC_NMAX);
}

// fp32 -> bf16 (hi, lo) split of a flat array.
__global__ void split_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
  uint32_t h0, l0, h1, l1;
  split2(v.x, v.y, h0, l0);
  split2(v.z, v.w, h1, l1);
  reinterpret_cast<uint2*>(hi)[i] = make_uint2(h0, h1);
  if (lo) reinterpret_cast<uint2*>(lo)[i] = make_uint2(l0, l1);
}

// x = start(audio_0) (glow.py:156) in fp32 plus its bf16 operand copies.
__global__ void wn_start_tc_kernel(const float* __restrict__ audio, const float* __restrict__ w,
                                   const float* __restrict__ bias, float* __restrict__ x,
                                   __nv_bfloat16* __restrict__ x_hi, __nv_bfloat16* __restrict__ x_lo,
                                   long long n_cols, int C, int n_group, int off, int n_half) {
  const int c4 = C >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_cols * c4) return;
  const long long col = idx / c4;
  const int c = (int)(idx % c4) * 4;
  float4 o = __ldg(reinterpret_cast<const float4*>(bias + c));
  for (int j = 0; j < n_half; ++j) {
    const float a = __ldg(audio + col * n_group + off + j);
    const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (long long)j * C + c));
    o.x = fmaf(a, wv.x, o.x);
    o.y = fmaf(a, wv.y, o.y);
    o.z = fmaf(a, wv.z, o.z);
    o.w = fmaf(a, wv.w, o.w);
  }
  *reinterpret_cast<float4*>(x + col * C + c) = o;
  uint32_t h0, l0, h1, l1;
  split2(o.x, o.y, h0, l0);
  split2(o.z, o.w, h1, l1);
  *reinterpret_cast<uint2*>(x_hi + col * C + c) = make_uint2(h0, h1);
  if (x_lo) *reinterpret_cast<uint2*>(x_lo + col * C + c) = make_uint2(l0, l1);
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    return reinterpret_cast<EncodeTiledFn>(ptr);
  }();
  return fn;
}

// (B, T, C) channels-last bf16 activation: box = 16 channels x 128 rows of one utterance.
int make_act_map(CUtensorMap* m, const void* ptr, int B, int T, int C) {
  EncodeTiledFn fn = encode_fn();
  FAC_REQUIRE(fn != nullptr, "tensor-core path: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)T * C * 2};
  cuuint32_t box[3] = {TC_BK, TC_BM, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FAC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation %dx%dx%d) failed: %d", B, T, C, (int)r);
  return 0;
}
// (N, K) row-major bf16 weight: box = 16 k x min(256, N) rows.
int make_weight_map(CUtensorMap* m, const void* ptr, int N, int K) {
  EncodeTiledFn fn = encode_fn();
  FAC_REQUIRE(fn != nullptr, "tensor-core path: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {TC_BK, (cuuint32_t)(N < TC_NHALF ? N : TC_NHALF)};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FAC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weight %dx%d) failed: %d", N, K, (int)r);
  return 0;
}

int sm_count() {
  static int sms = [] {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  return sms;
}

int launch_tc(const CUtensorMap maps[6], const TcParams& p, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wn_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    if (e != cudaSuccess) {
      set_error("wn_gemm_tc: cannot reserve %d bytes of shared memory: %s", TC_SMEM, cudaGetErrorString(e));
      return 2;
    }
    attr_set = true;
  }
  const int grid = p.n_tiles < sm_count() ? p.n_tiles : sm_count();
  wn_gemm_tc_kernel<<<grid, TC_THREADS, TC_SMEM, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], p);
  count_launch();
  return check_launch("wn_gemm_tc_kernel");
}

}  // namespace

int wg_check_model(const fac_wg_model* m);
int wg_upsample_squeeze(const fac_wg_model* m, const float* mel_cl, float* spect, int B, int F, cudaStream_t st);
int wg_end(const fac_wg_model* m, int flow, const float* skip, float* audio, int B, int Tg, cudaStream_t st);

static int tc_check(const fac_wg_model* m, const fac_wg_tc_weights* w, const fac_wg_tc_workspace* ws, int nsplit) {
  if (int rc = wg_check_model(m)) return rc;
  FAC_REQUIRE(w && ws, "tensor-core path: NULL weights/workspace");
  FAC_REQUIRE(nsplit == 1 || nsplit == 2, "tensor-core path: nsplit must be 1 (bf16) or 2 (split-bf16), got %d", nsplit);
  const int C = m->n_channels, n_cond = m->n_mel * m->n_group;
  FAC_REQUIRE(C % 16 == 0 && n_cond % 16 == 0 && 2 * C <= TC_NMAX,
              "tensor-core path: needs n_channels %% 16 == 0, n_cond %% 16 == 0 and 2*n_channels <= %d", TC_NMAX);
  FAC_REQUIRE(ws->spect_hi && ws->x && ws->x_hi && ws->acts_hi && ws->skip, "tensor-core path: workspace incomplete");
  if (nsplit == 2) FAC_REQUIRE(ws->spect_lo && ws->x_lo && ws->acts_lo, "tensor-core path: lo buffers missing");
  return 0;
}

int wg_tc_prepare_spect(const fac_wg_model* m, const fac_wg_tc_workspace* ws, const float* mel_cl, int B, int F,
                        int nsplit, cudaStream_t st) {
  FAC_REQUIRE(ws && ws->spect_f32, "tensor-core path: spect_f32 scratch missing");
  if (int rc = wg_upsample_squeeze(m, mel_cl, ws->spect_f32, B, F, st)) return rc;
  const long long n4 = (long long)B * F * (m->hop / m->n_group) * m->n_mel * m->n_group / 4;
  split_bf16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(
      ws->spect_f32, reinterpret_cast<__nv_bfloat16*>(ws->spect_hi),
      nsplit == 2 ? reinterpret_cast<__nv_bfloat16*>(ws->spect_lo) : nullptr, n4);
  count_launch();
  return check_launch("split_bf16_kernel");
}

int wg_tc_start(const fac_wg_model* m, int flow, const float* audio, const fac_wg_tc_workspace* ws, int B, int Tg,
                int nsplit, cudaStream_t st) {
  const fac_wg_flow& f = m->flows[flow];
  const long long n_cols = (long long)B * Tg;
  const long long total = n_cols * (m->n_channels / 4);
  wn_start_tc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      audio, f.start_w, f.start_b, ws->x, reinterpret_cast<__nv_bfloat16*>(ws->x_hi),
      nsplit == 2 ? reinterpret_cast<__nv_bfloat16*>(ws->x_lo) : nullptr, n_cols, m->n_channels, m->n_group,
      m->n_group - f.n_rem, f.n_half);
  count_launch();
  return check_launch("wn_start_tc_kernel");
}

int wg_tc_layer(const fac_wg_model* m, const fac_wg_tc_weights* w, int flow, int layer, const fac_wg_tc_workspace* ws,
                int B, int Tg, int nsplit, cudaStream_t st) {
  if (int rc = tc_check(m, w, ws, nsplit)) return rc;
  FAC_REQUIRE(flow >= 0 && flow < m->n_flows && layer >= 0 && layer < m->n_layers, "wn_layer_tc: index out of range");
  const fac_wg_flow& f = m->flows[flow];
  const fac_wg_tc_flow& wf = w->flows[flow];
  const int C = m->n_channels, n_cond = m->n_mel * m->n_group, ks = m->kernel_size;
  const int dil = 1 << layer;
  const bool last = layer == m->n_layers - 1;
  CUtensorMap maps[6];
  TcParams p{};
  p.T = Tg;
  p.B = B;
  p.tiles_per_batch = ceil_div(Tg, TC_BM);
  p.n_tiles = B * p.tiles_per_batch;
  p.nsplit = nsplit;
  p.C = C;
  // ---- G1: [x taps | spect] -> gate -> acts
  const int K1 = ks * C + n_cond;
  if (int rc = make_act_map(&maps[0], ws->x_hi, B, Tg, C)) return rc;
  if (int rc = make_act_map(&maps[1], nsplit == 2 ? ws->x_lo : ws->x_hi, B, Tg, C)) return rc;
  if (int rc = make_act_map(&maps[2], ws->spect_hi, B, Tg, n_cond)) return rc;
  if (int rc = make_act_map(&maps[3], nsplit == 2 ? ws->spect_lo : ws->spect_hi, B, Tg, n_cond)) return rc;
  if (int rc = make_weight_map(&maps[4], wf.w1_hi[layer], 2 * C, K1)) return rc;
  if (int rc = make_weight_map(&maps[5], nsplit == 2 ? wf.w1_lo[layer] : wf.w1_hi[layer], 2 * C, K1)) return rc;
  p.n_src = 2;
  p.src[0] = TcSrc{C, ks, dil, dil * (ks - 1) / 2};
  p.src[1] = TcSrc{n_cond, 1, 0, 0};
  p.n_total = 2 * C;
  p.k_steps = K1 / TC_BK;
  p.mode = TC_GATE;
  p.bias = f.in_cond_b[layer];
  p.acts_hi = reinterpret_cast<__nv_bfloat16*>(ws->acts_hi);
  p.acts_lo = reinterpret_cast<__nv_bfloat16*>(ws->acts_lo);
  if (int rc = launch_tc(maps, p, st)) return rc;
  // ---- G2: acts -> res/skip
  const int n_rs = last ? C : 2 * C;
  if (int rc = make_act_map(&maps[0], ws->acts_hi, B, Tg, C)) return rc;
  if (int rc = make_act_map(&maps[1], nsplit == 2 ? ws->acts_lo : ws->acts_hi, B, Tg, C)) return rc;
  maps[2] = maps[0];
  maps[3] = maps[1];
  if (int rc = make_weight_map(&maps[4], wf.w2_hi[layer], n_rs, C)) return rc;
  if (int rc = make_weight_map(&maps[5], nsplit == 2 ? wf.w2_lo[layer] : wf.w2_hi[layer], n_rs, C)) return rc;
  p.n_src = 1;
  p.src[0] = TcSrc{C, 1, 0, 0};
  p.n_total = n_rs;
  p.k_steps = C / TC_BK;
  p.mode = TC_RES_SKIP;
  p.bias = f.res_skip_b[layer];
  p.x = ws->x;
  p.x_hi = reinterpret_cast<__nv_bfloat16*>(ws->x_hi);
  p.x_lo = reinterpret_cast<__nv_bfloat16*>(ws->x_lo);
  p.skip = ws->skip;
  p.n_split_cols = last ? 0 : C;
  p.accumulate_skip = layer > 0;
  return launch_tc(maps, p, st);
}

int wg_infer_tc(const fac_wg_model* m, const fac_wg_tc_weights* w, const float* mel_cl, float* audio,
                const fac_wg_tc_workspace* ws, int B, int F, int nsplit, cudaStream_t st) {
  if (int rc = tc_check(m, w, ws, nsplit)) return rc;
  FAC_REQUIRE(mel_cl && audio, "waveglow_infer_tc: NULL argument");
  const int Tg = F * (m->hop / m->n_group);
  if (int rc = wg_tc_prepare_spect(m, ws, mel_cl, B, F, nsplit, st)) return rc;
  for (int k = m->n_flows - 1; k >= 0; --k) {
    if (int rc = wg_tc_start(m, k, audio, ws, B, Tg, nsplit, st)) return rc;
    for (int i = 0; i < m->n_layers; ++i)
      if (int rc = wg_tc_layer(m, w, k, i, ws, B, Tg, nsplit, st)) return rc;
    if (int rc = wg_end(m, k, ws->skip, audio, B, Tg, st)) return rc;
  }
  return 0;
}

}  // namespace fac
