// Autoregressive decoder of the PPG->Mel model as ONE persistent cooperative kernel
// (reference src/common/model.py:489-535 Decoder.inference, :387-442 decode, :100-121 Attention,
// :56-60 LocationLayer, :124-135 Prenet, src/common/utils.py:46-78 window mask).
//
// The reference spends ~40 kernel launches and >= 3 device->host syncs per output frame; here the
// whole loop (up to max_steps frames, all B utterances in lock step) is a single launch:
//   * the two LSTMCells (1200 x 1200 fp32 each, 11.5 MB together) are split by hidden unit across
//     the shared memory of all CTAs (one CTA per SM) and stay resident for the whole sequence;
//   * per step: [attention LSTM | all CTAs] -> grid barrier -> [location-sensitive attention over the
//     +-window only | one CTA per utterance] -> barrier -> [decoder LSTM | all CTAs] -> barrier ->
//     [mel projection + stop gate + prenet of the next step | one CTA per utterance] -> barrier;
//   * the window mask of utils.py:46-78 sets every energy outside [t-w, t+w] to -inf, i.e. those
//     softmax weights are exactly 0, so only the <= 2w+1 window positions are ever evaluated;
//   * the stop decision (sigmoid(gate) > threshold) is taken on the device.
#include "fac_common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace fac {

struct DecParams {
  fac_taco_decoder_weights w;
  const float* memory;        // (B, T_in, E)
  const float* pmem;          // (B, T_in, A)
  const int* lengths;         // [B]
  const unsigned char* drop;  // (max_steps, 2, B, P) in {0, 1}
  fac_taco_decoder_state s;
  float* mel;    // (B, max_steps, M)
  float* gate;   // (B, max_steps)
  float* align;  // (B, max_steps, T_in) pre-zeroed, or NULL
  int B, T_in, max_steps, window;
  float gate_threshold;
};

namespace {

constexpr int DEC_THREADS = 256;
constexpr int DEC_WARPS = DEC_THREADS / 32;
constexpr int R = 300;    // attention_rnn_dim == decoder_rnn_dim == prenet_dim
constexpr int E = 600;    // encoder_embedding_dim
constexpr int A = 150;    // attention_dim
constexpr int M = 80;     // n_acoustic_feat_dims
constexpr int NF = 32;    // attention_location_n_filters
constexpr int KF = 31;    // attention_location_kernel_size
constexpr int KIN = R + E + R;  // 1200: LSTMCell input | hidden concatenation
constexpr int MAXU = 3;         // hidden units per CTA (needs >= 100 CTAs)
constexpr int CHUNK = 8;        // utterances staged per LSTM pass
constexpr int MAXW = 64;        // max window positions (2*window+1 <= 64)

struct Smem {
  float w_att[MAXU * 4][KIN];
  float w_dec[MAXU * 4][KIN];
  float w_loc[2 * KF][NF];   // [c*KF + k][f]
  float w_ld[NF][A];         // location_dense transposed
  float v[A];
  alignas(16) union {
    float in[CHUNK][KIN];    // LSTM phases
    struct {                 // attention / projection phases
      float hq[R];
      float part[DEC_WARPS][R];
      float pq[A];
      float cat[2][MAXW + KF - 1 + 2];
      float loc[MAXW][NF];
      float e[MAXW];
      float hc[R + E];
      float melv[M];
      float p1[R];
    } a;
  } u;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

// One LSTMCell for the units [u0, u0+nu) of this CTA and all B utterances.
// in = [x0 (R) | ctx (E) | h_prev (R)]; writes h_next / c.
__device__ void lstm_phase(Smem& sm, const float (*w_s)[KIN], const float* __restrict__ bias, const float* x0,
                           const float* ctx, const float* h_prev, float* h_next, float* c, int B, int u0, int nu) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int n0 = 0; n0 < B; n0 += CHUNK) {
    const int nb = min(CHUNK, B - n0);
    __syncthreads();
    for (int i = tid; i < nb * KIN; i += DEC_THREADS) {
      const int n = i / KIN, k = i - n * KIN;
      const int b = n0 + n;
      float v;
      // written by other CTAs in the previous phase: read through L2 (.cg), never a stale L1 line
      if (k < R) v = __ldcg(x0 + b * R + k);
      else if (k < R + E) v = __ldcg(ctx + b * E + (k - R));
      else v = __ldcg(h_prev + b * R + (k - R - E));
      sm.u.in[n][k] = v;
    }
    __syncthreads();
    // warp tile: the 4 gate rows of one unit x 4 utterances
    const int n_tiles = nu * ((nb + 3) / 4);
    for (int tile = warp; tile < n_tiles; tile += DEC_WARPS) {
      const int u = tile % nu, ng = (tile / nu) * 4;
      float acc[4][4];
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int n = 0; n < 4; ++n) acc[g][n] = 0.f;
      for (int k4 = lane; k4 < KIN / 4; k4 += 32) {
        float4 wv[4], xv[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) wv[g] = *reinterpret_cast<const float4*>(&w_s[g * nu + u][k4 * 4]);
#pragma unroll
        for (int n = 0; n < 4; ++n)
          xv[n] = (ng + n < nb) ? *reinterpret_cast<const float4*>(&sm.u.in[ng + n][k4 * 4])
                                : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int n = 0; n < 4; ++n)
            acc[g][n] = fmaf(wv[g].x, xv[n].x,
                             fmaf(wv[g].y, xv[n].y, fmaf(wv[g].z, xv[n].z, fmaf(wv[g].w, xv[n].w, acc[g][n]))));
      }
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int n = 0; n < 4; ++n) acc[g][n] = warp_sum(acc[g][n]);
      if (lane < 4 && ng + lane < nb) {
        const int b = n0 + ng + lane, j = u0 + u;
        float gv[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float a = lane == 0 ? acc[g][0] : lane == 1 ? acc[g][1] : lane == 2 ? acc[g][2] : acc[g][3];
          gv[g] = a + __ldg(bias + g * R + j);
        }
        const float cn = sigmoidf_exact(gv[1]) * c[b * R + j] + sigmoidf_exact(gv[0]) * tanhf(gv[2]);
        c[b * R + j] = cn;
        h_next[b * R + j] = sigmoidf_exact(gv[3]) * tanhf(cn);
      }
    }
  }
}

// out[j] (j < n_out) = sum_k wt[k][j] * x[k], k < n_in; wt is (n_in, n_out) row-major in global memory.
// K is split across the 8 warps, lanes run over j (coalesced), partials reduced through shared memory.
__device__ void matvec_t(Smem& sm, const float* __restrict__ wt, const float* x_s, int n_in, int n_out, float* out_s) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kper = (n_in + DEC_WARPS - 1) / DEC_WARPS;
  const int k0 = warp * kper, k1 = min(n_in, k0 + kper);
  for (int j = lane; j < n_out; j += 32) {
    float acc = 0.f;
    for (int k = k0; k < k1; ++k) acc = fmaf(__ldg(wt + (long long)k * n_out + j), x_s[k], acc);
    sm.u.a.part[warp][j] = acc;
  }
  __syncthreads();
  for (int j = tid; j < n_out; j += DEC_THREADS) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < DEC_WARPS; ++w) acc += sm.u.a.part[w][j];
    out_s[j] = acc;
  }
  __syncthreads();
}

__device__ __forceinline__ void window_bounds(int t, int window, int len, int& start, int& end) {
  // src/common/utils.py:70-74
  const int max_idx = len - 1;
  start = min(max(0, t - window), max_idx);
  end = min(t + window, max_idx);
}

__device__ void attention_phase(Smem& sm, const DecParams& p, int b, int t, const float* h_att) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int len = p.lengths[b];
  int start, end;
  window_bounds(t, p.window, len, start, end);
  const int nw = end - start + 1;
  __syncthreads();
  for (int i = tid; i < R; i += DEC_THREADS) sm.u.a.hq[i] = __ldcg(h_att + b * R + i);
  // previous / cumulative weights around the window (zero outside the sequence: conv padding)
  const int c0 = start - (KF - 1) / 2, ncat = nw + KF - 1;
  float* wprev = p.s.w_prev + (long long)b * p.T_in;
  float* wcum = p.s.w_cum + (long long)b * p.T_in;
  for (int i = tid; i < 2 * ncat; i += DEC_THREADS) {
    const int c = i / ncat, q = i - c * ncat, pos = c0 + q;
    float v = 0.f;
    if (pos >= 0 && pos < p.T_in) v = c == 0 ? wprev[pos] : wcum[pos];
    sm.u.a.cat[c][q] = v;
  }
  __syncthreads();
  matvec_t(sm, p.w.wq_t, sm.u.a.hq, R, A, sm.u.a.pq);   // query_layer (model.py:92)
  // location_conv (model.py:57): loc[q][f] = sum_{c,k} w[f][c][k] * cat[c][q + k]
  for (int i = tid; i < nw * NF; i += DEC_THREADS) {
    const int q = i / NF, f = i - q * NF;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c)
      for (int k = 0; k < KF; ++k) acc = fmaf(sm.w_loc[c * KF + k][f], sm.u.a.cat[c][q + k], acc);
    sm.u.a.loc[q][f] = acc;
  }
  __syncthreads();
  // energies (model.py:94-97): e[q] = v . tanh(pq + location_dense(loc[q]) + processed_memory[q])
  for (int q = warp; q < nw; q += DEC_WARPS) {
    const float* pm = p.pmem + ((long long)b * p.T_in + start + q) * A;
    float part = 0.f;
    for (int a = lane; a < A; a += 32) {
      float pa = 0.f;
#pragma unroll
      for (int f = 0; f < NF; ++f) pa = fmaf(sm.w_ld[f][a], sm.u.a.loc[q][f], pa);
      part = fmaf(sm.v[a], tanhf(sm.u.a.pq[a] + pa + __ldg(pm + a)), part);
    }
    part = warp_sum(part);
    if (lane == 0) sm.u.a.e[q] = part;
  }
  __syncthreads();
  // softmax over the window (everything else is -inf -> weight 0, model.py:114-117)
  if (warp == 0) {
    float mx = -INFINITY;
    for (int q = lane; q < nw; q += 32) mx = fmaxf(mx, sm.u.a.e[q]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    float sum = 0.f;
    for (int q = lane; q < nw; q += 32) {
      const float ex = expf(sm.u.a.e[q] - mx);
      sm.u.a.e[q] = ex;
      sum += ex;
    }
    sum = warp_sum(sum);
    for (int q = lane; q < nw; q += 32) sm.u.a.e[q] = sm.u.a.e[q] / sum;
  }
  __syncthreads();
  // context (model.py:118): ctx = sum_q w[q] * memory[start + q]
  for (int c = tid; c < E; c += DEC_THREADS) {
    const float* mrow = p.memory + ((long long)b * p.T_in + start) * E + c;
    float acc = 0.f;
    for (int q = 0; q < nw; ++q) acc = fmaf(sm.u.a.e[q], __ldg(mrow + (long long)q * E), acc);
    p.s.ctx[b * E + c] = acc;
  }
  // new attention_weights (zero outside the window), cumulative weights (model.py:424), alignments
  int ostart = 0, oend = -1;
  if (t > 0) window_bounds(t - 1, p.window, len, ostart, oend);
  for (int pos = ostart + tid; pos <= oend; pos += DEC_THREADS)
    if (pos < start || pos > end) wprev[pos] = 0.f;
  for (int q = tid; q < nw; q += DEC_THREADS) {
    const float wv = sm.u.a.e[q];
    wprev[start + q] = wv;
    wcum[start + q] += wv;
    if (p.align) p.align[((long long)b * p.max_steps + t) * p.T_in + start + q] = wv;
  }
}

__device__ void output_phase(Smem& sm, const DecParams& p, int b, int t, const float* h_dec) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __syncthreads();
  for (int i = tid; i < R + E; i += DEC_THREADS)
    sm.u.a.hc[i] = i < R ? __ldcg(h_dec + b * R + i) : __ldcg(p.s.ctx + b * E + (i - R));
  __syncthreads();
  // linear_projection (rows 0..M-1) and gate_layer (row M), model.py:436-441
  for (int r = warp; r <= M; r += DEC_WARPS) {
    const float* wr = p.w.w_proj + (long long)r * (R + E);
    float acc = 0.f;
    for (int k = lane; k < R + E; k += 32) acc = fmaf(__ldg(wr + k), sm.u.a.hc[k], acc);
    acc = warp_sum(acc) + __ldg(p.w.b_proj + r);
    if (lane == 0) {
      if (r < M) {
        sm.u.a.melv[r] = acc;
        p.mel[((long long)b * p.max_steps + t) * M + r] = acc;
      } else {
        p.gate[(long long)b * p.max_steps + t] = acc;
        // stop test (model.py:524), per utterance
        if (p.s.out_len[b] == 0) {
          if (sigmoidf_exact(acc) > p.gate_threshold) {
            p.s.out_len[b] = t + 1;
            atomicAdd(p.s.done, 1);
          } else if (t + 1 == p.max_steps) {
            p.s.out_len[b] = p.max_steps;   // model.py:526-528 "Reached max decoder steps"
            atomicAdd(p.s.done + 1, 1);
          }
        }
      }
    }
  }
  __syncthreads();
  if (t + 1 >= p.max_steps) return;
  // prenet of the next step (model.py:507, 132-135): dropout p=0.5 is always on -> mask * 2
  const unsigned char* d1 = p.drop + (((long long)(t + 1) * 2 + 0) * p.B + b) * R;
  const unsigned char* d2 = p.drop + (((long long)(t + 1) * 2 + 1) * p.B + b) * R;
  for (int j = tid; j < R; j += DEC_THREADS) {
    float acc = 0.f;
    for (int k = 0; k < M; ++k) acc = fmaf(__ldg(p.w.w_pre1_t + k * R + j), sm.u.a.melv[k], acc);
    sm.u.a.p1[j] = fmaxf(acc, 0.f) * (2.0f * (float)d1[j]);
  }
  __syncthreads();
  matvec_t(sm, p.w.w_pre2_t, sm.u.a.p1, R, R, sm.u.a.hq);
  for (int j = tid; j < R; j += DEC_THREADS) p.s.pre[b * R + j] = fmaxf(sm.u.a.hq[j], 0.f) * (2.0f * (float)d2[j]);
}

__global__ void __launch_bounds__(DEC_THREADS, 1) taco_decoder_kernel(const DecParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x;
  const int G = gridDim.x;
  const int u0 = (int)((long long)blockIdx.x * R / G);
  const int u1 = (int)((long long)(blockIdx.x + 1) * R / G);
  const int nu = u1 - u0;

  // resident weights: the LSTM rows of this CTA's hidden units + the small attention tensors
  for (int i = tid; i < nu * 4 * KIN; i += DEC_THREADS) {
    const int q = i / KIN, k = i - q * KIN;
    const int g = q / nu, u = q - g * nu;
    sm.w_att[q][k] = __ldg(p.w.w_att + (long long)(g * R + u0 + u) * KIN + k);
    sm.w_dec[q][k] = __ldg(p.w.w_dec + (long long)(g * R + u0 + u) * KIN + k);
  }
  for (int i = tid; i < 2 * KF * NF; i += DEC_THREADS) {
    const int ck = i / NF, f = i - ck * NF;            // w_loc is (NF, 2, KF)
    sm.w_loc[ck][f] = __ldg(p.w.w_loc + f * 2 * KF + ck);
  }
  for (int i = tid; i < NF * A; i += DEC_THREADS) sm.w_ld[i / A][i % A] = __ldg(p.w.w_ld_t + i);
  for (int i = tid; i < A; i += DEC_THREADS) sm.v[i] = __ldg(p.w.v + i);
  __syncthreads();

  int cur = 0;
  for (int t = 0; t < p.max_steps; ++t) {
    float* h_att_cur = p.s.h_att + cur * p.B * R;
    float* h_att_nxt = p.s.h_att + (cur ^ 1) * p.B * R;
    float* h_dec_cur = p.s.h_dec + cur * p.B * R;
    float* h_dec_nxt = p.s.h_dec + (cur ^ 1) * p.B * R;
    // attention_rnn (model.py:400-402): input [prenet | context], hidden h_att
    lstm_phase(sm, sm.w_att, p.w.b_att, p.s.pre, p.s.ctx, h_att_cur, h_att_nxt, p.s.c_att, p.B, u0, nu);
    grid.sync();
    for (int b = blockIdx.x; b < p.B; b += G) attention_phase(sm, p, b, t, h_att_nxt);
    grid.sync();
    // decoder_rnn (model.py:425-428): input [h_att | context], hidden h_dec
    lstm_phase(sm, sm.w_dec, p.w.b_dec, h_att_nxt, p.s.ctx, h_dec_cur, h_dec_nxt, p.s.c_dec, p.B, u0, nu);
    grid.sync();
    for (int b = blockIdx.x; b < p.B; b += G) output_phase(sm, p, b, t, h_dec_nxt);
    grid.sync();
    cur ^= 1;
    // every utterance has fired its stop gate (or hit max_steps): uniform exit
    const volatile int* done = p.s.done;
    if (done[0] + done[1] >= p.B) {
      if (blockIdx.x == 0 && tid == 0) p.s.done[2] = t + 1;
      break;
    }
  }
}

}  // namespace

int taco_decoder_run(const fac_taco_decoder_weights* w, const float* memory, const float* pmem, const int* lengths,
                     const unsigned char* drop, const fac_taco_decoder_state* s, float* mel, float* gate, float* align,
                     int B, int T_in, int max_steps, int window, float gate_threshold, cudaStream_t st) {
  FAC_REQUIRE(w && memory && pmem && lengths && drop && s && mel && gate, "taco_decoder: NULL argument");
  FAC_REQUIRE(B > 0 && T_in > 0 && max_steps > 0, "taco_decoder: empty problem");
  FAC_REQUIRE(window >= 0 && 2 * window + 1 <= MAXW, "taco_decoder: attention window %d unsupported (max %d)", window,
              (MAXW - 1) / 2);
  int dev = 0, sms = 0, coop = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  FAC_REQUIRE(coop, "taco_decoder: device lacks cooperative launch");
  FAC_REQUIRE(sms * MAXU >= R, "taco_decoder: needs >= %d SMs, device has %d", (R + MAXU - 1) / MAXU, sms);
  const size_t smem = sizeof(Smem);
  cudaError_t e = cudaFuncSetAttribute(taco_decoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("taco_decoder: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return 2;
  }
  DecParams p{};
  p.w = *w;
  p.memory = memory; p.pmem = pmem; p.lengths = lengths; p.drop = drop;
  p.s = *s;
  p.mel = mel; p.gate = gate; p.align = align;
  p.B = B; p.T_in = T_in; p.max_steps = max_steps; p.window = window; p.gate_threshold = gate_threshold;
  void* args[] = {&p};
  e = cudaLaunchCooperativeKernel((void*)taco_decoder_kernel, dim3(sms), dim3(DEC_THREADS), args, smem, st);
  count_launch();
  if (e != cudaSuccess) {
    set_error("taco_decoder: cooperative launch failed: %s", cudaGetErrorString(e));
    return 2;
  }
  return check_launch("taco_decoder_kernel");
}

}  // namespace fac
