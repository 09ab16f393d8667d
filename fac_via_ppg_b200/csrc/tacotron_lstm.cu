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
constexpr int LSTM_THREADS = 256;
constexpr int LSTM_MAXK = 10;  // ceil(H / 32) for H <= 320

template <int NB>
__global__ void __cluster_dims__(LSTM_CLUSTER, 1, 1) __launch_bounds__(LSTM_THREADS, 1)
    bilstm_cluster_kernel(const float* __restrict__ xp, const float* __restrict__ w_hh, float* __restrict__ out,
                          int B, int T, int H, int upc) {
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cluster_id = blockIdx.x / LSTM_CLUSTER;
  const int dir = cluster_id & 1;
  const int n0 = (cluster_id >> 1) * NB;  // first utterance of this cluster
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int u0 = rank * upc;
  const int nu = max(0, min(H, u0 + upc) - u0);  // hidden units owned by this CTA
  const int n_rows = 4 * nu;

  float* w_s = smem;                        // [4*upc][H]   row q = g*nu + u  <->  W_hh[g*H + u0 + u][:]
  float* h_buf = w_s + 4 * upc * H;         // [2][NB][H]
  float* g_s = h_buf + 2 * NB * H;          // [4*upc][NB] recurrent part of the gates

  const float* w_dir = w_hh + (long long)dir * 4 * H * H;
  for (int i = tid; i < n_rows * H; i += LSTM_THREADS) {
    const int q = i / H, k = i - q * H;
    const int g = q / nu, u = q - g * nu;
    w_s[i] = __ldg(w_dir + (long long)(g * H + u0 + u) * H + k);
  }
  for (int i = tid; i < 2 * NB * H; i += LSTM_THREADS) h_buf[i] = 0.f;
  cluster.sync();

  // cell-update role: thread (u, n) keeps c in a register for the whole sequence
  const int cu = tid / NB, cn = tid - cu * NB;
  const bool cell_thread = cu < nu && (n0 + cn) < B;
  float c_state = 0.f;
  const long long xp_row = 2LL * 4 * H;

  int cur = 0;
  for (int step = 0; step < T; ++step) {
    const int tt = dir == 0 ? step : T - 1 - step;
    float xg[4] = {0.f, 0.f, 0.f, 0.f};
    if (cell_thread) {
      const float* xrow = xp + ((long long)(n0 + cn) * T + tt) * xp_row + (long long)dir * 4 * H + u0 + cu;
#pragma unroll
      for (int g = 0; g < 4; ++g) xg[g] = __ldg(xrow + g * H);
    }
    // h_{t-1} of the NB utterances into registers, k = lane + 32*i
    float hreg[NB][LSTM_MAXK];
    const float* hcur = h_buf + cur * NB * H;
#pragma unroll
    for (int n = 0; n < NB; ++n)
#pragma unroll
      for (int i = 0; i < LSTM_MAXK; ++i) {
        const int k = lane + 32 * i;
        hreg[n][i] = k < H ? hcur[n * H + k] : 0.f;
      }
    for (int q = warp; q < n_rows; q += LSTM_THREADS / 32) {
      float acc[NB];
#pragma unroll
      for (int n = 0; n < NB; ++n) acc[n] = 0.f;
      const float* wrow = w_s + q * H;
#pragma unroll
      for (int i = 0; i < LSTM_MAXK; ++i) {
        const int k = lane + 32 * i;
        const float wv = k < H ? wrow[k] : 0.f;
#pragma unroll
        for (int n = 0; n < NB; ++n) acc[n] = fmaf(wv, hreg[n][i], acc[n]);
      }
#pragma unroll
      for (int n = 0; n < NB; ++n) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], s);
      }
      if (lane == 0) {
#pragma unroll
        for (int n = 0; n < NB; ++n) g_s[q * NB + n] = acc[n];
      }
    }
    __syncthreads();
    if (cu < nu) {
      float hval = 0.f;
      if (cell_thread) {
        const float gi = g_s[(0 * nu + cu) * NB + cn] + xg[0];
        const float gf = g_s[(1 * nu + cu) * NB + cn] + xg[1];
        const float gg = g_s[(2 * nu + cu) * NB + cn] + xg[2];
        const float go = g_s[(3 * nu + cu) * NB + cn] + xg[3];
        c_state = sigmoidf_exact(gf) * c_state + sigmoidf_exact(gi) * tanhf(gg);
        hval = sigmoidf_exact(go) * tanhf(c_state);
        out[((long long)(n0 + cn) * T + tt) * (2 * H) + dir * H + u0 + cu] = hval;
      }
      if (cn < NB) {
        float* slot = h_buf + (cur ^ 1) * NB * H + cn * H + u0 + cu;
#pragma unroll
        for (int r = 0; r < LSTM_CLUSTER; ++r) *cluster.map_shared_rank(slot, r) = hval;
      }
    }
    cluster.sync();  // release/acquire: every CTA sees the complete h_t before step t+1
    cur ^= 1;
  }
}

template <int NB>
int launch_bilstm(const float* xp, const float* w_hh, float* out, int B, int T, int H, cudaStream_t st) {
  const int upc = ceil_div(H, LSTM_CLUSTER);
  const size_t smem = (size_t)(4 * upc * H + 2 * NB * H + 4 * upc * NB) * sizeof(float);
  FAC_REQUIRE(upc * NB <= LSTM_THREADS, "bilstm: %d units x %d utterances exceed the block", upc, NB);
  cudaError_t e = cudaFuncSetAttribute(bilstm_cluster_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) {
    set_error("bilstm: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return 2;
  }
  const int groups = ceil_div(B, NB);
  bilstm_cluster_kernel<NB><<<2 * groups * LSTM_CLUSTER, LSTM_THREADS, smem, st>>>(xp, w_hh, out, B, T, H, upc);
  count_launch();
  return check_launch("bilstm_cluster_kernel");
}

}  // namespace

int lstm_bidir(const float* xp, const float* w_hh, float* out, int B, int T, int H, cudaStream_t st) {
  FAC_REQUIRE(xp && w_hh && out, "bilstm: NULL argument");
  FAC_REQUIRE(B > 0 && T > 0, "bilstm: empty problem B=%d T=%d", B, T);
  FAC_REQUIRE(H > 0 && H <= 32 * LSTM_MAXK, "bilstm: hidden size %d unsupported (max %d)", H, 32 * LSTM_MAXK);
  // fill the machine: 18 clusters of 8 CTAs fit 148 SMs; more utterances per cluster beyond that
  if (2 * B <= 18) return launch_bilstm<1>(xp, w_hh, out, B, T, H, st);
  if (B <= 18) return launch_bilstm<2>(xp, w_hh, out, B, T, H, st);
  return launch_bilstm<4>(xp, w_hh, out, B, T, H, st);
}

}  // namespace fac
