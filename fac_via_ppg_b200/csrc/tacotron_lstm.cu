// Encoder BiLSTM recurrence (reference src/common/model.py:211-213, 246-247: nn.LSTM, 1 layer,
// bidirectional, batch_first) as a thread-block-cluster kernel.
//
// The input projection x W_ih^T + b_ih + b_hh of both directions is one fac_conv_gemm_f32 call
// (xp, (B, T, 2*4H)); this kernel runs the sequential part h_t = f(xp_t + W_hh h_{t-1}).
// One cluster of 8 CTAs owns one direction and up to NB utterances: W_hh (4H x H fp32 = 1.44 MB
// for H = 300) is split by hidden unit across the 8 CTAs' shared memory and stays resident for
// all T steps; each step every CTA computes the 4 gates of its units, updates c/h, and pushes
// its slice of h into all 8 CTAs' next-step buffer through distributed shared memory, followed
// by one cluster barrier.  No HBM traffic per step besides reading xp and writing h.
#include "fac_common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace fac {
namespace {

constexpr int LSTM_CLUSTER = 8;
constexpr int LSTM_THREADS = 512;

// Sum 16 per-lane values across the warp with 16 shuffles (halving butterfly): afterwards lane l holds
// the total of value index (l >> 1) & 15 (lanes l and l^1 agree).
__device__ __forceinline__ float warp_reduce16(float (&v)[16], int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool up = lane & 16;
    const float send = up ? v[i] : v[i + 8];
    const float keep = up ? v[i + 8] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool up = lane & 8;
    const float send = up ? v[i] : v[i + 4];
    const float keep = up ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool up = lane & 4;
    const float send = up ? v[i] : v[i + 2];
    const float keep = up ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    const bool up = lane & 2;
    const float send = up ? v[0] : v[1];
    const float keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

constexpr int LSTM_MAXK4 = 3;   // ceil((H/4) / 32) for H <= 384

// NB is 1, 2 or 4 utterances per cluster.  Each warp owns whole hidden units: for a unit it computes the
// 4 gate rows x NB utterances with h_{t-1} held in registers (loaded once per step), reduces the 16
// partial sums with a halving butterfly, updates c/h in-warp and pushes h to all 8 CTAs of the cluster.
template <int NB>
__global__ void __cluster_dims__(LSTM_CLUSTER, 1, 1) __launch_bounds__(LSTM_THREADS, 1)
    bilstm_cluster_kernel(const float* __restrict__ xp, const float* __restrict__ w_hh, float* __restrict__ out,
                          int B, int T, int H, int upc, long long* prof) {
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cluster_id = blockIdx.x / LSTM_CLUSTER;
  const int dir = cluster_id & 1;
  const int n0 = (cluster_id >> 1) * NB;  // first utterance of this cluster
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int u0 = rank * upc;
  const int nu = max(0, min(H, u0 + upc) - u0);  // hidden units owned by this CTA
  const int H4 = H >> 2;

  float* w_s = smem;                        // [upc][4][H]  unit-major: the 4 gate rows of a unit are adjacent
  float* h_buf = w_s + 4 * upc * H;         // [2][4][H]    (utterance slots beyond NB stay zero)
  float* c_s = h_buf + 2 * 4 * H;           // [upc][4]     cell state

  const float* w_dir = w_hh + (long long)dir * 4 * H * H;
  for (int i = tid; i < nu * 4 * H; i += LSTM_THREADS) {
    const int u = i / (4 * H), rem = i - u * 4 * H;
    const int g = rem / H, k = rem - g * H;
    w_s[i] = __ldg(w_dir + (long long)(g * H + u0 + u) * H + k);
  }
  for (int i = tid; i < 2 * 4 * H; i += LSTM_THREADS) h_buf[i] = 0.f;
  for (int i = tid; i < upc * 4; i += LSTM_THREADS) c_s[i] = 0.f;
  cluster.sync();

  const long long xp_row = 2LL * 4 * H;
  constexpr int UPW = 3;                         // units per warp: ceil(38 / 16)
  // lane -> (gate, utterance) of the input projection it fetches; software-pipelined one step ahead so
  // the global-load latency never sits on the recurrence's critical path
  const int xn = lane & 3, xg = (lane >> 2) & 3;
  const bool x_lane = lane < 16 && xn < NB && n0 + xn < B;
  auto load_xp = [&](int step, float (&dst)[UPW]) {
    const int tt = dir == 0 ? step : T - 1 - step;
#pragma unroll
    for (int j = 0; j < UPW; ++j) {
      const int u = warp + j * (LSTM_THREADS / 32);
      dst[j] = (x_lane && u < nu && step < T)
                   ? __ldg(xp + ((long long)(n0 + xn) * T + tt) * xp_row + (long long)dir * 4 * H + xg * H + u0 + u)
                   : 0.f;
    }
  };
  float xv_cur[UPW], xv_nxt[UPW];
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
  for (int step = 0; step < T; ++step) {
    load_xp(step + 1, xv_nxt);
    // write h_{t-1} (complete in h_buf[cur] after the previous cluster barrier) to global now: the
    // stores drain during this step instead of stalling the next barrier's release
    if (step > 0) {
      const int tp = dir == 0 ? step - 1 : T - step;
      for (int i = tid; i < nu * NB; i += LSTM_THREADS) {
        const int n = i / nu, u = i - n * nu;
        if (n0 + n < B)
          out[((long long)(n0 + n) * T + tp) * (2 * H) + dir * H + u0 + u] = h_buf[cur * 4 * H + n * H + u0 + u];
      }
    }
    // h_{t-1} of the 4 utterance slots into registers: float4 index k4 = lane + 32*i
    float4 hreg[4][LSTM_MAXK4];
    const float* hcur = h_buf + cur * 4 * H;
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int i = 0; i < LSTM_MAXK4; ++i) {
        const int k4 = lane + 32 * i;
        hreg[n][i] = (n < NB && k4 < H4) ? *reinterpret_cast<const float4*>(hcur + n * H + 4 * k4)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    pmark(0);
#pragma unroll
    for (int j = 0; j < UPW; ++j) {
      const int u = warp + j * (LSTM_THREADS / 32);
      if (u >= nu) break;
      const float xv = xv_cur[j];
      float acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.f;
      const float* wu = w_s + (long long)u * 4 * H;
#pragma unroll
      for (int i = 0; i < LSTM_MAXK4; ++i) {
        const int k4 = lane + 32 * i;
        if (k4 < H4) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 wv = *reinterpret_cast<const float4*>(wu + g * H + 4 * k4);
#pragma unroll
            for (int n = 0; n < NB; ++n)
              acc[g * 4 + n] = fmaf(wv.x, hreg[n][i].x, fmaf(wv.y, hreg[n][i].y,
                                    fmaf(wv.z, hreg[n][i].z, fmaf(wv.w, hreg[n][i].w, acc[g * 4 + n]))));
          }
        }
      }
      const float tot = warp_reduce16(acc, lane);     // lane 2*(g*4+n) (and +1) holds gate g of utterance n
      // gather the 4 gates of utterance n = lane & 3 (every lane participates in the shuffles)
      const int n = lane & 3;
      const float gi = __shfl_sync(0xffffffffu, tot, 2 * (0 * 4 + n)) + __shfl_sync(0xffffffffu, xv, 0 * 4 + n);
      const float gf = __shfl_sync(0xffffffffu, tot, 2 * (1 * 4 + n)) + __shfl_sync(0xffffffffu, xv, 1 * 4 + n);
      const float gg = __shfl_sync(0xffffffffu, tot, 2 * (2 * 4 + n)) + __shfl_sync(0xffffffffu, xv, 2 * 4 + n);
      const float go = __shfl_sync(0xffffffffu, tot, 2 * (3 * 4 + n)) + __shfl_sync(0xffffffffu, xv, 3 * 4 + n);
      // every group of 4 lanes now holds the same (n-indexed) gates: lane = r*4 + n serves cluster rank r
      float hval = 0.f;
      if (n < NB && n0 + n < B) {
        const float c_old = c_s[u * 4 + n];
        const float c_new = sigmoidf_fast(gf) * c_old + sigmoidf_fast(gi) * tanhf_fast(gg);
        hval = sigmoidf_fast(go) * tanhf_fast(c_new);
        __syncwarp(__activemask());
        if (lane < 4) c_s[u * 4 + n] = c_new;
      }
      if (n < NB) {
        float* slot = h_buf + (cur ^ 1) * 4 * H + n * H + u0 + u;
        *cluster.map_shared_rank(slot, lane >> 2) = hval;     // 8 ranks x 4 utterance slots = 32 lanes
      }
      __syncwarp();
    }
    pmark(1);
    cluster.sync();  // release/acquire: every CTA sees the complete h_t before step t+1
    pmark(2);
    cur ^= 1;
#pragma unroll
    for (int j = 0; j < UPW; ++j) xv_cur[j] = xv_nxt[j];
  }
  if (prof != nullptr && tid == 0)
    for (int i = 0; i < 3; ++i) prof[blockIdx.x * 4 + i] = pa[i];
  // the last step's h
  {
    const int tp = dir == 0 ? T - 1 : 0;
    for (int i = tid; i < nu * NB; i += LSTM_THREADS) {
      const int n = i / nu, u = i - n * nu;
      if (n0 + n < B)
        out[((long long)(n0 + n) * T + tp) * (2 * H) + dir * H + u0 + u] = h_buf[cur * 4 * H + n * H + u0 + u];
    }
  }
}

long long* g_lstm_prof = nullptr;

// Clusters of 8 CTAs (1 CTA per SM) that can be resident at once: the GPC boundaries leave fewer than
// SMs / 8 (measured on B200: 16 clusters requested -> a second wave, 2x the time).
template <int NB>
int max_resident_clusters(size_t smem) {
  static int cached = -1;
  if (cached >= 0) return cached;
  cudaFuncSetAttribute(bilstm_cluster_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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
  if (cudaOccupancyMaxActiveClusters(&n, bilstm_cluster_kernel<NB>, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = 1;
  }
  cached = n;
  return n;
}

template <int NB>
int launch_bilstm(const float* xp, const float* w_hh, float* out, int B, int T, int H, cudaStream_t st) {
  const int upc = ceil_div(H, LSTM_CLUSTER);
  const size_t smem = (size_t)(4 * upc * H + 2 * 4 * H + 4 * upc) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(bilstm_cluster_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) {
    set_error("bilstm: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return 2;
  }
  const int groups = ceil_div(B, NB);
  bilstm_cluster_kernel<NB><<<2 * groups * LSTM_CLUSTER, LSTM_THREADS, smem, st>>>(xp, w_hh, out, B, T, H, upc,
                                                                                   g_lstm_prof);
  count_launch();
  return check_launch("bilstm_cluster_kernel");
}

}  // namespace

void lstm_set_prof(long long* p) { g_lstm_prof = p; }

int lstm_bidir(const float* xp, const float* w_hh, float* out, int B, int T, int H, cudaStream_t st) {
  FAC_REQUIRE(xp && w_hh && out, "bilstm: NULL argument");
  FAC_REQUIRE(B > 0 && T > 0, "bilstm: empty problem B=%d T=%d", B, T);
  FAC_REQUIRE(ceil_div(ceil_div(H, LSTM_CLUSTER), LSTM_THREADS / 32) <= 3, "bilstm: hidden size %d needs more than 3 units per warp", H);
  FAC_REQUIRE(H > 0 && H % 4 == 0 && H <= 128 * LSTM_MAXK4, "bilstm: hidden size %d unsupported (multiple of 4, max %d)",
              H, 128 * LSTM_MAXK4);
  // one wave: the fewest utterances per cluster whose cluster count is resident at once
  const size_t smem = (size_t)(4 * ceil_div(H, LSTM_CLUSTER) * H + 2 * 4 * H + 4 * ceil_div(H, LSTM_CLUSTER)) * sizeof(float);
  if (2 * B <= max_resident_clusters<1>(smem)) return launch_bilstm<1>(xp, w_hh, out, B, T, H, st);
  if (2 * ceil_div(B, 2) <= max_resident_clusters<2>(smem)) return launch_bilstm<2>(xp, w_hh, out, B, T, H, st);
  return launch_bilstm<4>(xp, w_hh, out, B, T, H, st);
}

}  // namespace fac
