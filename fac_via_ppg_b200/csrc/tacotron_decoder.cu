// Autoregressive decoder of the PPG->Mel model as ONE persistent cooperative kernel
// (reference src/common/model.py:489-535 Decoder.inference, :387-442 decode, :100-121 Attention,
// :56-60 LocationLayer, :124-135 Prenet, src/common/utils.py:46-78 window mask).
//
// The reference spends ~40 kernel launches and >= 3 device->host syncs per output frame; here the
// whole loop (up to max_steps frames, all B utterances in lock step) is a single launch.
//   * EVERY weight matrix of the step is split by output row across the shared memory of all CTAs
//     (one CTA per SM) and stays resident for the whole sequence: the two LSTMCells (11.5 MB fp32),
//     the query layer, [mel projection | stop gate | prenet layer 0 composed with the projection]
//     and prenet layer 1.  Per step the only global traffic is the small state vectors (L2 resident).
//   * A step is six batched mat-vec phases separated by a lightweight grid barrier (one atomic +
//     acquire spin): attention LSTM | query | location-sensitive attention (one CTA per utterance,
//     only the <= 2w+1 window positions: everything outside [t-w, t+w] is masked to -inf by
//     utils.py:46-78, i.e. has softmax weight exactly 0) | decoder LSTM | projection+gate+prenet0 |
//     prenet1.  prenet0 has no bias or nonlinearity between it and the projection
//     (model.py:132-135, 436-438), so W_pre0 (W_proj hc + b) is evaluated as one composed matrix.
//   * the stop decision (sigmoid(gate) > threshold) is taken on the device.
#include "fac_common.cuh"

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

constexpr int DEC_THREADS = 512;
constexpr int DEC_WARPS = DEC_THREADS / 32;
constexpr int R = 300;    // attention_rnn_dim == decoder_rnn_dim == prenet_dim
constexpr int E = 600;    // encoder_embedding_dim
constexpr int A = 150;    // attention_dim
constexpr int M = 80;     // n_acoustic_feat_dims
constexpr int NF = 32;    // attention_location_n_filters
constexpr int KF = 31;    // attention_location_kernel_size
constexpr int KIN = R + E + R;  // 1200: LSTMCell input | hidden concatenation
constexpr int KHC = R + E;      // 900: [h_dec | context]
constexpr int NPP = M + 1 + R;  // 381 rows: mel projection, gate, composed prenet layer 0
constexpr int MAXU = 3;         // hidden units per CTA (needs >= 100 CTAs)
constexpr int MAXQ = 2;         // query rows per CTA
constexpr int MAXPP = 4;        // projection rows per CTA
constexpr int MAXP2 = 3;        // prenet-1 rows per CTA
constexpr int CHUNK = 8;        // utterances staged per pass
constexpr int MAXW = 64;        // max window positions (2*window+1 <= 64)
constexpr int KP = 2;           // K split of a mat-vec tile across warps (partials meet in shared memory)
constexpr int MAXTILES = 16;    // (rows or units) x utterance groups of 4 per staged chunk

struct Smem {
  float w_att[MAXU * 4][KIN];
  float w_dec[MAXU * 4][KIN];
  float w_q[MAXQ][R];
  float w_pp[MAXPP][KHC];
  float w_p2[MAXP2][R];
  float w_loc[2 * KF][NF];   // [c*KF + k][f]
  float w_ld[NF][A];         // location_dense transposed
  float v[A];
  float part[KP][MAXTILES][16];   // per-tile partial sums of the K halves
  alignas(16) union {
    float in[CHUNK][KIN];    // batched mat-vec phases
    struct {                 // attention phase
      float pq[A];
      float cat[2][MAXW + KF - 1 + 2];
      float loc[MAXW][NF];
      float e[MAXW];
    } a;
  } u;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

// Grid-wide barrier: monotonically increasing arrival counter (zeroed by the host), release/acquire.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
    __threadfence();
  }
  __syncthreads();
}

struct Seg {           // one piece of a concatenated input vector: utterance b reads ptr[b*stride + k]
  const float* ptr;
  int len, stride;     // multiples of 4 floats
};

// Stage the concatenated inputs of utterances [n0, n0+nb) into shared memory with 128-bit loads that
// bypass L1 (.cg): the vectors were written by other CTAs in the previous phase.
template <int NSEG>
__device__ __forceinline__ void stage_inputs(Smem& sm, const Seg (&segs)[NSEG], int n0, int nb) {
  __syncthreads();
  int base = 0;
#pragma unroll
  for (int s = 0; s < NSEG; ++s) {
    const int l4 = segs[s].len >> 2;
    for (int i = threadIdx.x; i < nb * l4; i += DEC_THREADS) {
      const int n = i / l4, k4 = i - n * l4;
      const float4 v = __ldcg(reinterpret_cast<const float4*>(segs[s].ptr + (long long)(n0 + n) * segs[s].stride) + k4);
      *reinterpret_cast<float4*>(&sm.u.in[n][base + 4 * k4]) = v;
    }
    base += segs[s].len;
  }
  __syncthreads();
}

// Sum 16 per-lane values across the warp with 16 shuffles (halving butterfly): afterwards lane l holds
// the total of value index ((l>>4)&1)*8 + ((l>>3)&1)*4 + ((l>>2)&1)*2 + ((l>>1)&1) (lanes l and l^1 agree).
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

// One LSTMCell for the units [u0, u0+nu) of this CTA and all B utterances.  Warp tile = the 4 gate rows
// of one unit x 4 utterances x one K half; the halves meet in shared memory before the cell update.
__device__ void lstm_phase(Smem& sm, const float (*w_s)[KIN], const float* __restrict__ bias, const Seg (&segs)[3],
                           float* h_next, float* c, int B, int u0, int nu) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int K4 = KIN / 4, K4H = (K4 + KP - 1) / KP;
  for (int n0 = 0; n0 < B; n0 += CHUNK) {
    const int nb = min(CHUNK, B - n0);
    stage_inputs(sm, segs, n0, nb);
    const int n_groups = (nb + 3) / 4;
    const int n_tiles = nu * n_groups;
    for (int job = warp; job < n_tiles * KP; job += DEC_WARPS) {
      const int tile = job % n_tiles, kp = job / n_tiles;
      const int u = tile % nu, ng = (tile / nu) * 4;
      float acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.f;
      for (int k4 = kp * K4H + lane; k4 < min(K4, (kp + 1) * K4H); k4 += 32) {
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
            acc[g * 4 + n] = fmaf(wv[g].x, xv[n].x, fmaf(wv[g].y, xv[n].y,
                                  fmaf(wv[g].z, xv[n].z, fmaf(wv[g].w, xv[n].w, acc[g * 4 + n]))));
      }
      const float total = warp_reduce16(acc, lane);
      if ((lane & 1) == 0) sm.part[kp][tile][lane >> 1] = total;   // value index (gate*4 + utterance) = lane>>1
    }
    __syncthreads();
    // cell update: one thread per (unit, utterance)
    for (int i = tid; i < nu * nb; i += DEC_THREADS) {
      const int u = i % nu, n = i / nu;
      const int tile = (n >> 2) * nu + u, j = u0 + u, b = n0 + n;
      float gv[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float a = __ldg(bias + g * R + j);
#pragma unroll
        for (int kp = 0; kp < KP; ++kp) a += sm.part[kp][tile][g * 4 + (n & 3)];
        gv[g] = a;
      }
      const float cn = sigmoidf_fast(gv[1]) * c[b * R + j] + sigmoidf_fast(gv[0]) * tanhf_fast(gv[2]);
      c[b * R + j] = cn;
      h_next[b * R + j] = sigmoidf_fast(gv[3]) * tanhf_fast(cn);
    }
  }
}

// Batched mat-vec over this CTA's resident rows: for every owned row r (global index row0 + r) and every
// utterance b, epi(row0 + r, b, dot(w_s[r], in_b)).  KROW % (4*KP) == 0.  Warp tile = one row x 4
// utterances x one K half.
template <int KROW, int NSEG, typename Epi>
__device__ void rows_phase(Smem& sm, const float (*w_s)[KROW], int row0, int nr, const Seg (&segs)[NSEG], int B,
                           Epi epi) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int K4 = KROW / 4, K4H = (K4 + KP - 1) / KP;
  for (int n0 = 0; n0 < B; n0 += CHUNK) {
    const int nb = min(CHUNK, B - n0);
    stage_inputs(sm, segs, n0, nb);
    const int n_groups = (nb + 3) / 4;
    const int n_tiles = nr * n_groups;
    for (int job = warp; job < n_tiles * KP; job += DEC_WARPS) {
      const int tile = job % n_tiles, kp = job / n_tiles;
      const int r = tile % nr, ng = (tile / nr) * 4;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k4 = kp * K4H + lane; k4 < min(K4, (kp + 1) * K4H); k4 += 32) {
        const float4 wv = *reinterpret_cast<const float4*>(&w_s[r][k4 * 4]);
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          if (ng + n < nb) {
            const float4 xv = *reinterpret_cast<const float4*>(&sm.u.in[ng + n][k4 * 4]);
            acc[n] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[n]))));
          }
        }
      }
#pragma unroll
      for (int n = 0; n < 4; ++n) acc[n] = warp_sum(acc[n]);
      if (lane < 4) sm.part[kp][tile][lane] = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
    }
    __syncthreads();
    for (int i = tid; i < nr * nb; i += DEC_THREADS) {
      const int r = i % nr, n = i / nr;
      const int tile = (n >> 2) * nr + r;
      float a = 0.f;
#pragma unroll
      for (int kp = 0; kp < KP; ++kp) a += sm.part[kp][tile][n & 3];
      epi(row0 + r, n0 + n, a);
    }
  }
}

__device__ __forceinline__ void window_bounds(int t, int window, int len, int& start, int& end) {
  // src/common/utils.py:70-74
  const int max_idx = len - 1;
  start = min(max(0, t - window), max_idx);
  end = min(t + window, max_idx);
}

__device__ void attention_phase(Smem& sm, const DecParams& p, int b, int t) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int len = p.lengths[b];
  int start, end;
  window_bounds(t, p.window, len, start, end);
  const int nw = end - start + 1;
  __syncthreads();
  for (int i = tid; i < A; i += DEC_THREADS) sm.u.a.pq[i] = __ldcg(p.s.pq + b * A + i);   // query_layer (model.py:92)
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

__device__ __forceinline__ void row_range(int n_rows, int& r0, int& nr) {
  r0 = (int)((long long)blockIdx.x * n_rows / gridDim.x);
  nr = (int)((long long)(blockIdx.x + 1) * n_rows / gridDim.x) - r0;
}

__global__ void __launch_bounds__(DEC_THREADS, 1) taco_decoder_kernel(const DecParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x;
  const int G = gridDim.x;
  int u0, nu, q0, nq, pp0, npp, p20, np2;
  row_range(R, u0, nu);
  row_range(A, q0, nq);
  row_range(NPP, pp0, npp);
  row_range(R, p20, np2);

  // ---- resident weights: this CTA's rows of every matrix of the step + the small attention tensors
  for (int i = tid; i < nu * 4 * KIN; i += DEC_THREADS) {
    const int q = i / KIN, k = i - q * KIN;
    const int g = q / nu, u = q - g * nu;
    sm.w_att[q][k] = __ldg(p.w.w_att + (long long)(g * R + u0 + u) * KIN + k);
    sm.w_dec[q][k] = __ldg(p.w.w_dec + (long long)(g * R + u0 + u) * KIN + k);
  }
  for (int i = tid; i < nq * R; i += DEC_THREADS) sm.w_q[i / R][i % R] = __ldg(p.w.wq + (long long)q0 * R + i);
  for (int i = tid; i < npp * KHC; i += DEC_THREADS)
    sm.w_pp[i / KHC][i % KHC] = __ldg(p.w.w_pp + (long long)pp0 * KHC + i);
  for (int i = tid; i < np2 * R; i += DEC_THREADS) sm.w_p2[i / R][i % R] = __ldg(p.w.w_pre2 + (long long)p20 * R + i);
  for (int i = tid; i < 2 * KF * NF; i += DEC_THREADS) {
    const int ck = i / NF, f = i - ck * NF;            // w_loc is (NF, 2, KF)
    sm.w_loc[ck][f] = __ldg(p.w.w_loc + f * 2 * KF + ck);
  }
  for (int i = tid; i < NF * A; i += DEC_THREADS) sm.w_ld[i / A][i % A] = __ldg(p.w.w_ld_t + i);
  for (int i = tid; i < A; i += DEC_THREADS) sm.v[i] = __ldg(p.w.v + i);
  __syncthreads();

  unsigned int* bar = reinterpret_cast<unsigned int*>(p.s.done + 3);
  unsigned int bar_target = 0;
  int cur = 0;
  for (int t = 0; t < p.max_steps; ++t) {
    float* h_att_cur = p.s.h_att + cur * p.B * R;
    float* h_att_nxt = p.s.h_att + (cur ^ 1) * p.B * R;
    float* h_dec_cur = p.s.h_dec + cur * p.B * R;
    float* h_dec_nxt = p.s.h_dec + (cur ^ 1) * p.B * R;
    // (1) attention_rnn (model.py:400-402): input [prenet | context], hidden h_att
    {
      const Seg segs[3] = {{p.s.pre, R, R}, {p.s.ctx, E, E}, {h_att_cur, R, R}};
      lstm_phase(sm, sm.w_att, p.w.b_att, segs, h_att_nxt, p.s.c_att, p.B, u0, nu);
    }
    grid_barrier(bar, bar_target);
    // (2) query_layer (model.py:92): pq = W_q h_att
    {
      const Seg segs[1] = {{h_att_nxt, R, R}};
      float* pq = p.s.pq;
      rows_phase<R>(sm, sm.w_q, q0, nq, segs, p.B, [pq](int row, int b, float v) { pq[b * A + row] = v; });
    }
    grid_barrier(bar, bar_target);
    // (3) location-sensitive attention, one CTA per utterance
    for (int b = blockIdx.x; b < p.B; b += G) attention_phase(sm, p, b, t);
    grid_barrier(bar, bar_target);
    // (4) decoder_rnn (model.py:425-428): input [h_att | context], hidden h_dec
    {
      const Seg segs[3] = {{h_att_nxt, R, R}, {p.s.ctx, E, E}, {h_dec_cur, R, R}};
      lstm_phase(sm, sm.w_dec, p.w.b_dec, segs, h_dec_nxt, p.s.c_dec, p.B, u0, nu);
    }
    grid_barrier(bar, bar_target);
    // (5) [linear_projection | gate_layer | prenet layer 0 o projection] on hc = [h_dec | context]
    //     (model.py:436-441, 507, 132-135)
    {
      const Seg segs[2] = {{h_dec_nxt, R, R}, {p.s.ctx, E, E}};
      const DecParams* pp = &p;
      const int tt = t;
      rows_phase<KHC>(sm, sm.w_pp, pp0, npp, segs, p.B, [pp, tt](int row, int b, float v) {
        const DecParams& q = *pp;
        v += __ldg(q.w.b_pp + row);
        if (row < M) {
          q.mel[((long long)b * q.max_steps + tt) * M + row] = v;
        } else if (row == M) {
          q.gate[(long long)b * q.max_steps + tt] = v;
          if (q.s.out_len[b] == 0) {          // stop test (model.py:524), per utterance
            if (sigmoidf_exact(v) > q.gate_threshold) {
              q.s.out_len[b] = tt + 1;
              atomicAdd(q.s.done, 1);
            } else if (tt + 1 == q.max_steps) {
              q.s.out_len[b] = q.max_steps;     // model.py:526-528 "Reached max decoder steps"
              atomicAdd(q.s.done + 1, 1);
            }
          }
        } else if (tt + 1 < q.max_steps) {
          // prenet layer 0 of the NEXT step; dropout p = 0.5 is always on -> mask * 2
          const int j = row - M - 1;
          const unsigned char d = q.drop[(((long long)(tt + 1) * 2 + 0) * q.B + b) * R + j];
          q.s.p1[b * R + j] = fmaxf(v, 0.f) * (2.0f * (float)d);
        }
      });
    }
    grid_barrier(bar, bar_target);
    // (6) prenet layer 1 of the next step
    if (t + 1 < p.max_steps) {
      const Seg segs[1] = {{p.s.p1, R, R}};
      const DecParams* pp = &p;
      const int tt = t;
      rows_phase<R>(sm, sm.w_p2, p20, np2, segs, p.B, [pp, tt](int row, int b, float v) {
        const DecParams& q = *pp;
        const unsigned char d = q.drop[(((long long)(tt + 1) * 2 + 1) * q.B + b) * R + row];
        q.s.pre[b * R + row] = fmaxf(v, 0.f) * (2.0f * (float)d);
      });
    }
    grid_barrier(bar, bar_target);
    cur ^= 1;
    // every utterance has fired its stop gate (or hit max_steps): uniform exit
    const volatile int* done = p.s.done;
    if (done[0] + done[1] >= p.B) {
      if (blockIdx.x == 0 && tid == 0) p.s.done[2] = t + 1;
      break;
    }
  }
}

__global__ void __launch_bounds__(DEC_THREADS, 1) grid_barrier_selftest_kernel(unsigned int* counter, int iters) {
  unsigned int target = 0;
  for (int i = 0; i < iters; ++i) grid_barrier(counter, target);
}

}  // namespace

// Diagnostic: `iters` back-to-back grid barriers on a full cooperative grid (1 CTA / SM); the caller
// times the launch to get the per-barrier cost that bounds the decoder's step latency.
int selftest_grid_barrier(unsigned int* counter, int iters, cudaStream_t st) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  void* args[] = {&counter, &iters};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)grid_barrier_selftest_kernel, dim3(sms), dim3(DEC_THREADS), args,
                                              0, st);
  count_launch();
  if (e != cudaSuccess) {
    set_error("selftest_grid_barrier: %s", cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

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
  FAC_REQUIRE(sms * MAXU >= R && sms * MAXQ >= A && sms * (MAXPP - 1) >= NPP,
              "taco_decoder: needs >= %d SMs, device has %d", (NPP + MAXPP - 2) / (MAXPP - 1), sms);
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
  // cooperative launch = co-residency guarantee for the grid barrier
  e = cudaLaunchCooperativeKernel((void*)taco_decoder_kernel, dim3(sms), dim3(DEC_THREADS), args, smem, st);
  count_launch();
  if (e != cudaSuccess) {
    set_error("taco_decoder: cooperative launch failed: %s", cudaGetErrorString(e));
    return 2;
  }
  return check_launch("taco_decoder_kernel");
}

}  // namespace fac
