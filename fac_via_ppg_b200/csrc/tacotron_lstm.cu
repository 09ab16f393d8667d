// Encoder BiLSTM recurrence (reference src/common/model.py:211-213, 246-247: nn.LSTM, 1 layer,
// bidirectional, batch_first) as a thread-block-cluster kernel.
//
// The input projection x W_ih^T + b_ih + b_hh of both directions is one GEMM (xp, (B, T, 2*4H)); this
// kernel runs the sequential part h_t = f(xp_t + W_hh h_{t-1}).  One cluster of 8 CTAs owns one
// direction and up to 8 utterances: W_hh (4H x H = 1.44 MB fp32 for H = 300) is split by hidden unit
// across the 8 CTAs' shared memory and stays resident for all T steps as IEEE-half hi/lo pairs
// (w * 2^8 = hi + lo to ~2^-22).  Each step a CTA computes the 4 gates of its units for the cluster's
// utterances on the tensor cores (mma.sync m16n8k16, split-fp16: hi*hi + lo*hi + hi*lo, fp32
// accumulate; warp = (K quarter, group of m-tiles), the quarters meet in shared memory in fp32),
// updates c (registers) and h, and pushes its slice of h into all 8 CTAs' next-step buffer through
// distributed shared memory with stores that complete on the receiver's mbarrier (no cluster barrier per step:
// measured 1.3 k of the 5.5 k cycles of a step).  No HBM traffic per step besides
// reading xp (prefetched one step ahead) and writing h.
#include "fac_common.cuh"
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace cg = cooperative_groups;

namespace fac {
namespace {

constexpr int LSTM_CLUSTER = 8;
constexpr int LSTM_THREADS = 512;
constexpr int LSTM_NB = 8;            // utterance slots per cluster = one mma n-tile
constexpr int LSTM_KQ = 4;            // K quarters (warp & 3)
constexpr int LSTM_MG = 4;            // groups of m-tiles (warp >> 2), <= 3 m-tiles each
constexpr float LSTM_W_SCALE = 256.f;

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split_half2(float2 x, uint32_t& hi, uint32_t& lo) {
  // hi = x with the low 13 mantissa bits cleared (exactly representable in half: 11 significant bits), lo = the
  // exact remainder rounded to half.  Type conversions run at 16 lanes per cycle per SM, an eighth of the FP32
  // rate: this form needs two of them per pair of values instead of six (round, convert back, round again).
  const float hx = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
  const float hy = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
  const __half2 h = __floats2half2_rn(hx, hy);
  const __half2 l = __floats2half2_rn(x.x - hx, x.y - hy);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// ---- hand-over of h between the CTAs of a cluster without a cluster barrier: every remote store carries its own
// completion (st.async ... mbarrier::complete_tx) into the RECEIVING CTA's mbarrier of the buffer it fills, which
// that CTA arms with the byte count of a whole h_t; the CTA waits for the phase before it reads the buffer.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, int rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr),
               "r"(__float_as_uint(v)), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void mbar_init1(uint32_t bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arm(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
      ::"r"(bar), "r"(parity)
      : "memory");
}

struct LstmGeom {
  int upc;        // hidden units per CTA
  int rows;       // 4 * upc gate rows (unit-major: row = unit * 4 + gate)
  int m_tiles;    // ceil(rows / 16)
  int k_steps;    // ceil(H / 16)
  int ks;         // halfs per weight row  (16 * k_steps + 8: ldmatrix rows hit distinct banks)
  int hs;         // floats per h row      (16 * k_steps + 8)
  size_t off_wlo, off_part, off_h, off_bar, bytes;
};
inline LstmGeom lstm_geom(int H) {
  LstmGeom g;
  g.upc = ceil_div(H, LSTM_CLUSTER);
  g.rows = 4 * g.upc;
  g.m_tiles = ceil_div(g.rows, 16);
  g.k_steps = ceil_div(H, 16);
  g.ks = 16 * g.k_steps + 8;
  g.hs = 16 * g.k_steps + 8;
  const size_t w_bytes = (size_t)g.rows * g.ks * sizeof(__half);
  g.off_wlo = w_bytes;
  g.off_part = 2 * w_bytes;
  g.off_h = g.off_part + (size_t)LSTM_KQ * g.m_tiles * 16 * LSTM_NB * sizeof(float);
  g.off_bar = g.off_h + (size_t)2 * LSTM_NB * g.hs * sizeof(float);     // two mbarriers, one per h buffer
  g.bytes = g.off_bar + 2 * sizeof(unsigned long long);
  return g;
}

__global__ void __cluster_dims__(LSTM_CLUSTER, 1, 1) __launch_bounds__(LSTM_THREADS, 1)
    bilstm_cluster_kernel(const float* __restrict__ xp, const float* __restrict__ w_hh, float* __restrict__ out,
                          const int* __restrict__ lengths, int B, int T, int H, int nb_per_cluster, const LstmGeom geo,
                          long long* prof) {
  extern __shared__ __align__(16) unsigned char smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cluster_id = blockIdx.x / LSTM_CLUSTER;
  const int dir = cluster_id & 1;
  const int n0 = (cluster_id >> 1) * nb_per_cluster;   // first utterance of this cluster
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int upc = geo.upc, u0 = rank * upc;
  const int nu = max(0, min(H, u0 + upc) - u0);          // hidden units owned by this CTA
  const int ks = geo.ks, hs = geo.hs;

  __half* w_hi = reinterpret_cast<__half*>(smem);
  __half* w_lo = reinterpret_cast<__half*>(smem + geo.off_wlo);
  float* part = reinterpret_cast<float*>(smem + geo.off_part);   // [KQ][m_tiles*16][NB]
  float* h_buf = reinterpret_cast<float*>(smem + geo.off_h);     // [2][NB][hs]
  const uint32_t h_bar = smem_u32(smem + geo.off_bar);           // [2] one per h buffer
  const int prow = geo.m_tiles * 16;
  if (tid == 0) {
    mbar_init1(h_bar);
    mbar_init1(h_bar + 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // resident weights: row = unit * 4 + gate, zero K padding and zero rows for units past the end
  const float* w_dir = w_hh + (long long)dir * 4 * H * H;
  for (int i = tid; i < geo.rows * ks; i += LSTM_THREADS) {
    const int r = i / ks, k = i - r * ks;
    const int u = r >> 2, g = r & 3;
    float w = 0.f;
    if (u < nu && k < H) w = __ldg(w_dir + (long long)(g * H + u0 + u) * H + k) * LSTM_W_SCALE;
    const __half h = __float2half_rn(w);
    w_hi[i] = h;
    w_lo[i] = __float2half_rn(w - __half2float(h));
  }
  for (int i = tid; i < 2 * LSTM_NB * hs; i += LSTM_THREADS) h_buf[i] = 0.f;
  cluster.sync();

  // ---- warp job: K quarter kq, m-tiles [mt0, mt0 + nmt)
  const int kq = warp & (LSTM_KQ - 1), mg = warp >> 2;
  const int kbase = geo.k_steps / LSTM_KQ, kextra = geo.k_steps % LSTM_KQ;
  const int my_steps = kbase + (kq < kextra ? 1 : 0), k_first = kq * kbase + min(kq, kextra);
  const int mbase = geo.m_tiles / LSTM_MG, mextra = geo.m_tiles % LSTM_MG;
  const int nmt = mbase + (mg < mextra ? 1 : 0), mt0 = mg * mbase + min(mg, mextra);
  // ldmatrix: lane l addresses row (l & 7) + 8 * ((l >> 3) & 1) of the 8 x 8 block at k offset 8 * (l >> 4)
  uint32_t a_off[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    int r = (mt0 + j) * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
    if (j >= nmt || r >= geo.rows) r = 0;
    a_off[j] = (uint32_t)(r * ks + (lane >> 4) * 8 + k_first * 16) * 2;
  }
  const uint32_t whi_s = (uint32_t)__cvta_generic_to_shared(w_hi), wlo_s = (uint32_t)__cvta_generic_to_shared(w_lo);

  // ---- epilogue thread = (unit eu, utterance slot en); c lives in a register for the whole sequence
  const int eu = tid % upc, en = tid / upc;
  const bool e_act = tid < upc * LSTM_NB && eu < nu && en < nb_per_cluster && n0 + en < B;
  // ragged batch: this utterance has my_len valid rows; its reverse direction starts at row my_len - 1 and the
  // steps beyond its length neither read xp nor write out (their state is never consumed)
  const int my_len = e_act ? (lengths != nullptr ? min(max(__ldg(lengths + n0 + en), 0), T) : T) : 0;
  const long long xp_row = 2LL * 4 * H;
  auto load_xp = [&](int step, float (&dst)[4]) {
    const int tt = dir == 0 ? step : my_len - 1 - step;
#pragma unroll
    for (int g = 0; g < 4; ++g)
      dst[g] = (e_act && step < my_len)
                   ? __ldg(xp + ((long long)(n0 + en) * T + tt) * xp_row + (long long)dir * 4 * H + g * H + u0 + eu)
                   : 0.f;
  };
  float xv_cur[4], xv_nxt[4], c_reg = 0.f;
  load_xp(0, xv_cur);

  int cur = 0;
  long long pa[4] = {0, 0, 0, 0}, pt = clock64();
  auto pmark = [&](int i) {
    if (prof != nullptr && tid == 0) {
      const long long now = clock64();
      pa[i] += now - pt;
      pt = now;
    }
  };
  const uint32_t h_bytes = (uint32_t)(H * nb_per_cluster) * sizeof(float);    // what all CTAs together send per step
  for (int step = 0; step < T; ++step) {
    load_xp(step + 1, xv_nxt);
    // this step fills buffer cur ^ 1 everywhere: arm its barrier here; then wait for h_{t-1} (buffer cur, filled
    // during the previous step; the zero state of step 0 needs no wait)
    if (tid == 0) mbar_arm(h_bar + 8 * (cur ^ 1), h_bytes);
    if (step > 0) mbar_wait_parity(h_bar + 8 * cur, (uint32_t)((step - 1) >> 1) & 1u);
    // ---- gate pre-activations of this CTA's units: [rows x K] x [K x NB] on the tensor cores
    {
      const float* xrow = h_buf + (cur * LSTM_NB + (lane >> 2)) * hs + k_first * 16 + 2 * (lane & 3);
      float acc[3][3][4];
#pragma unroll
      for (int i = 0; i < 36; ++i) (&acc[0][0][0])[i] = 0.f;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        if (i < my_steps) {
          uint32_t bh[2], bl[2];
          split_half2(*reinterpret_cast<const float2*>(xrow + i * 16), bh[0], bl[0]);
          split_half2(*reinterpret_cast<const float2*>(xrow + i * 16 + 8), bh[1], bl[1]);
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            if (j < nmt) {
              uint32_t ah[4], al[4];
              ldmatrix_x4(ah, whi_s + a_off[j] + i * 32);
              ldmatrix_x4(al, wlo_s + a_off[j] + i * 32);
              mma_f16(acc[j][0], ah, bh);
              mma_f16(acc[j][1], al, bh);
              mma_f16(acc[j][2], ah, bl);
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (j < nmt) {
          float v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) v[q] = acc[j][0][q] + (acc[j][1][q] + acc[j][2][q]);
          float* dst = part + ((long long)kq * prow + (mt0 + j) * 16 + (lane >> 2)) * LSTM_NB + 2 * (lane & 3);
          *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
          *reinterpret_cast<float2*>(dst + 8 * LSTM_NB) = make_float2(v[2], v[3]);
        }
      }
    }
    __syncthreads();
    pmark(0);
    // ---- cell update (gate order i, f, g, o) and hand-over of h_t
    float hval = 0.f;
    if (e_act) {
      float gv[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < LSTM_KQ; ++q) a += part[((long long)q * prow + eu * 4 + g) * LSTM_NB + en];
        gv[g] = a * (1.0f / LSTM_W_SCALE) + xv_cur[g];
      }
      c_reg = sigmoidf_fast(gv[1]) * c_reg + sigmoidf_fast(gv[0]) * tanhf_fast(gv[2]);
      hval = sigmoidf_fast(gv[3]) * tanhf_fast(c_reg);
      const int tt = dir == 0 ? step : my_len - 1 - step;
      if (step < my_len) out[((long long)(n0 + en) * T + tt) * (2 * H) + dir * H + u0 + eu] = hval;
    }
    if (tid < upc * LSTM_NB && eu < nu && en < nb_per_cluster) {
      const uint32_t slot = smem_u32(h_buf + ((cur ^ 1) * LSTM_NB + en) * hs + u0 + eu), bar = h_bar + 8 * (cur ^ 1);
#pragma unroll
      for (int r = 0; r < LSTM_CLUSTER; ++r) st_async_f32(mapa_u32(slot, r), hval, mapa_u32(bar, r));
    }
    pmark(1);
    // no cluster barrier: a CTA can be at most one step ahead of the slowest one (it needs everybody's h_t), and
    // what it then overwrites remotely (buffer cur of step t + 1 = this step's cur ^ 1 ... of step t - 1) has been
    // read by its owner before that owner sent the h_t the writer waited for
    pmark(2);
    cur ^= 1;
#pragma unroll
    for (int g = 0; g < 4; ++g) xv_cur[g] = xv_nxt[g];
  }
  cluster.sync();        // nobody leaves while a peer may still store into this CTA
  if (prof != nullptr && tid == 0)
    for (int i = 0; i < 3; ++i) prof[blockIdx.x * 4 + i] = pa[i];
}

long long* g_lstm_prof = nullptr;

// Clusters of 8 CTAs (1 CTA per SM) that can be resident at once: the GPC boundaries leave fewer than
// SMs / 8 (measured on B200: 16 clusters requested -> a second wave, 2x the time).
int max_resident_clusters(size_t smem) {
  static int cached_on[FAC_MAX_DEVICES];
  static size_t cached_smem_on[FAC_MAX_DEVICES] = {};
  const int slot = current_device_slot();
  int& cached = cached_on[slot];
  size_t& cached_smem = cached_smem_on[slot];
  if (cached > 0 && cached_smem == smem) return cached;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(LSTM_CLUSTER * 32);
  cfg.blockDim = dim3(LSTM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = LSTM_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, bilstm_cluster_kernel, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = 1;
  }
  cached = n;
  cached_smem = smem;
  return n;
}

}  // namespace

void lstm_set_prof(long long* p) { g_lstm_prof = p; }

int lstm_bidir(const float* xp, const float* w_hh, float* out, const int* lengths, int B, int T, int H,
               cudaStream_t st) {
  FAC_REQUIRE(xp && w_hh && out, "bilstm: NULL argument");
  FAC_REQUIRE(B > 0 && T > 0, "bilstm: empty problem B=%d T=%d", B, T);
  FAC_REQUIRE(H > 0 && H % 4 == 0, "bilstm: hidden size %d must be a positive multiple of 4", H);
  const LstmGeom geo = lstm_geom(H);
  FAC_REQUIRE(geo.m_tiles <= 3 * LSTM_MG && geo.k_steps <= 5 * LSTM_KQ && geo.upc * LSTM_NB <= LSTM_THREADS &&
                  geo.bytes <= 227 * 1024,
              "bilstm: hidden size %d unsupported (needs %zu bytes of shared memory per CTA)", H, geo.bytes);
  cudaError_t e = cudaFuncSetAttribute(bilstm_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)geo.bytes);
  if (e != cudaSuccess) {
    set_error("bilstm: cannot reserve %zu bytes of shared memory: %s", geo.bytes, cudaGetErrorString(e));
    return 2;
  }
  // one wave: the fewest utterances per cluster (<= 8) whose cluster count is resident at once
  const int resident = max_resident_clusters(geo.bytes);
  int nb = 1;
  while (nb < LSTM_NB && 2 * ceil_div(B, nb) > resident) ++nb;
  const int groups = ceil_div(B, nb);
  bilstm_cluster_kernel<<<2 * groups * LSTM_CLUSTER, LSTM_THREADS, geo.bytes, st>>>(xp, w_hh, out, lengths, B, T, H,
                                                                                    nb, geo, g_lstm_prof);
  count_launch();
  return check_launch("bilstm_cluster_kernel");
}

}  // namespace fac
