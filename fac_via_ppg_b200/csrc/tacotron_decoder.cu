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
  long long* prof;            // optional [grid][16] cycle counters: phase body / barrier wait x 5, total
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
constexpr int MAXPP = 3;        // projection rows per CTA (needs >= 127 CTAs)
constexpr int MAXP2 = 3;        // prenet-1 rows per CTA
constexpr int CHUNK = 8;        // utterances staged per pass (double-buffered)
constexpr int MAXW = 64;        // max window positions (2*window+1 <= 64)
constexpr int CTXP = 3;         // q-range split of the context sum

struct AttScratch {             // attention phase; shares storage with the staged inputs
  float w_loc[2 * KF][NF];      // [c*KF + k][f]   (reloaded from L2 every step: 8 KB)
  float w_ld[NF][A];            // location_dense transposed (19 KB)
  float pq[A + 2];
  float cat[2][MAXW + KF - 1 + 2];
  float loc[MAXW][NF];
  float e[MAXW];
  alignas(16) float ctxp[CTXP][E];
};

struct Smem {
  float w_att[MAXU * 4][KIN];
  float w_dec[MAXU * 4][KIN];
  float w_pp[MAXPP][KHC];
  float w_p2[MAXP2][R];
  float v[A + 2];
  float b_att[MAXU * 4], b_dec[MAXU * 4], b_pp[4];
  float part[2][DEC_WARPS][16];   // per-job partial sums (one job per warp per chunk), double-buffered
  alignas(16) union {
    float in[2][CHUNK][KIN];      // staged input vectors of a chunk of utterances, double-buffered
    AttScratch a;
  } u;
};
static_assert(sizeof(Smem) <= 227 * 1024, "decoder shared memory");

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

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct Seg {           // one piece of a concatenated input vector: utterance b reads ptr[b*stride + k]
  const float* ptr;
  int len, stride;     // multiples of 4 floats
};

// Start the asynchronous copy (L2 -> shared memory, bypassing L1: the vectors were written by other CTAs
// in the previous phase) of the concatenated inputs of utterances [n0, n0+nb) into one staging buffer.
template <int NSEG>
__device__ __forceinline__ void stage_async(float (*buf)[KIN], const Seg (&segs)[NSEG], int n0, int nb) {
  int base = 0;
#pragma unroll
  for (int s = 0; s < NSEG; ++s) {
    const int l4 = segs[s].len >> 2;
    for (int i = threadIdx.x; i < nb * l4; i += DEC_THREADS) {
      const int n = i / l4, k4 = i - n * l4;
      cp_async16(&buf[n][base + 4 * k4], segs[s].ptr + (long long)(n0 + n) * segs[s].stride + 4 * k4);
    }
    base += segs[s].len;
  }
  cp_async_commit();
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

// One batched mat-vec phase over this CTA's resident rows.  The rows are grouped in `n_rt` row tiles of
// (up to) 4 rows -- the 4 gate rows of one LSTM unit, or the CTA's projection / prenet rows -- and the B
// utterances in chunks of CHUNK whose input vectors are copied into shared memory asynchronously, one
// chunk ahead of the arithmetic.  Inside a chunk one warp owns one job = (row tile, 4 utterances, K part):
// 16 accumulators per lane over its K slice, reduced across the lanes with a halving butterfly; the K
// parts meet in shared memory and `epi(chunk buffer, first utterance, utterances, n_tiles, kparts)` finishes.
//   row(rt, r): shared-memory pointer of row r of row tile rt (any valid row when r is past the end)
//   pre(n0, nb): called right after the copies are in flight (loads the epilogue wants early)
template <int NSEG, typename RowFn, typename PreFn, typename EpiFn>
__device__ __forceinline__ void matvec_phase(Smem& sm, const Seg (&segs)[NSEG], int K4, int n_rt, int B, RowFn row,
                                             PreFn pre, EpiFn epi) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_chunks = (B + CHUNK - 1) / CHUNK;
  stage_async(sm.u.in[0], segs, 0, min(CHUNK, B));
  for (int ci = 0; ci < n_chunks; ++ci) {
    const int n0 = ci * CHUNK, nb = min(CHUNK, B - n0), buf = ci & 1;
    pre(n0, nb);
    if (ci + 1 < n_chunks) {
      stage_async(sm.u.in[buf ^ 1], segs, n0 + CHUNK, min(CHUNK, B - n0 - CHUNK));
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int n_groups = (nb + 3) >> 2;
    const int n_tiles = n_rt * n_groups;                 // <= 6
    const int kparts = DEC_WARPS / n_tiles;              // K split so that every warp has one job
    const int tile = warp % n_tiles, kp = warp / n_tiles;
    if (kp < kparts) {
      const int rt = tile % n_rt, ng = (tile / n_rt) * 4;
      const float4* w0 = reinterpret_cast<const float4*>(row(rt, 0));
      const float4* w1 = reinterpret_cast<const float4*>(row(rt, 1));
      const float4* w2 = reinterpret_cast<const float4*>(row(rt, 2));
      const float4* w3 = reinterpret_cast<const float4*>(row(rt, 3));
      const float (*xin)[KIN] = sm.u.in[buf];
      // utterances past the end of the chunk re-read the last valid one (results ignored)
      const float4* x0 = reinterpret_cast<const float4*>(xin[min(ng + 0, nb - 1)]);
      const float4* x1 = reinterpret_cast<const float4*>(xin[min(ng + 1, nb - 1)]);
      const float4* x2 = reinterpret_cast<const float4*>(xin[min(ng + 2, nb - 1)]);
      const float4* x3 = reinterpret_cast<const float4*>(xin[min(ng + 3, nb - 1)]);
      float acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.f;
      const int k_end = (kp + 1) * K4 / kparts;
      for (int k4 = kp * K4 / kparts + lane; k4 < k_end; k4 += 32) {
        const float4 wv[4] = {w0[k4], w1[k4], w2[k4], w3[k4]};
        const float4 xv[4] = {x0[k4], x1[k4], x2[k4], x3[k4]};
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int n = 0; n < 4; ++n)
            acc[g * 4 + n] = fmaf(wv[g].x, xv[n].x, fmaf(wv[g].y, xv[n].y,
                                  fmaf(wv[g].z, xv[n].z, fmaf(wv[g].w, xv[n].w, acc[g * 4 + n]))));
      }
      const float total = warp_reduce16(acc, lane);
      if ((lane & 1) == 0) sm.part[buf][warp][lane >> 1] = total;   // value index (row*4 + utterance) = lane>>1
    }
    __syncthreads();
    epi(buf, n0, nb, n_tiles, kparts);
  }
}

// sum over the K parts of value (row r, utterance n of the chunk) of row tile rt
__device__ __forceinline__ float part_sum(const Smem& sm, int buf, int n_rt, int n_tiles, int kparts, int rt, int r,
                                          int n) {
  const int tile = (n >> 2) * n_rt + rt;
  float a = 0.f;
  for (int kp = 0; kp < kparts; ++kp) a += sm.part[buf][kp * n_tiles + tile][r * 4 + (n & 3)];
  return a;
}

// One LSTMCell (model.py:400-402 / 425-428) for the units [u0, u0+nu) of this CTA and all B utterances.
__device__ __forceinline__ void lstm_phase(Smem& sm, const float (*w_s)[KIN], const float* bias_s, const Seg (&segs)[3],
                                           float* h_next, float* c, int B, int u0, int nu) {
  const int tid = threadIdx.x;
  float c_old = 0.f;
  matvec_phase(
      sm, segs, KIN / 4, nu, B, [&](int rt, int g) { return &w_s[g * nu + rt][0]; },
      [&](int n0, int nb) {
        if (tid < nu * nb) c_old = c[(n0 + tid / nu) * R + u0 + tid % nu];   // only this thread ever touches it
      },
      [&](int buf, int n0, int nb, int n_tiles, int kparts) {
        if (tid < nu * nb) {                        // cell update: one thread per (unit, utterance)
          const int u = tid % nu, n = tid / nu, j = u0 + u, b = n0 + n;
          float gv[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) gv[g] = bias_s[g * nu + u] + part_sum(sm, buf, nu, n_tiles, kparts, u, g, n);
          const float cn = sigmoidf_fast(gv[1]) * c_old + sigmoidf_fast(gv[0]) * tanhf_fast(gv[2]);
          c[b * R + j] = cn;
          h_next[b * R + j] = sigmoidf_fast(gv[3]) * tanhf_fast(cn);
        }
      });
}

__device__ __forceinline__ void window_bounds(int t, int window, int len, int& start, int& end) {
  // src/common/utils.py:70-74
  const int max_idx = len - 1;
  start = min(max(0, t - window), max_idx);
  end = min(t + window, max_idx);
}

// Location-sensitive attention of utterance b at step t (model.py:100-121, 56-60, 78-98), including the
// query projection W_q h_att (model.py:92), whose 180 KB matrix is streamed from L2 by this one CTA.
__device__ void attention_phase(Smem& sm, const DecParams& p, const float* h_att, int b, int t) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  AttScratch& s = sm.u.a;
  const int len = p.lengths[b];
  int start, end;
  window_bounds(t, p.window, len, start, end);
  const int nw = end - start + 1;
  // ---- everything that does not depend on this step's arithmetic is requested up front
  for (int i = tid; i < 2 * KF * NF / 4; i += DEC_THREADS) cp_async16(&s.w_loc[0][0] + 4 * i, p.w.w_loc + 4 * i);
  for (int i = tid; i < NF * A / 4; i += DEC_THREADS) cp_async16(&s.w_ld[0][0] + 4 * i, p.w.w_ld_t + 4 * i);
  cp_async_commit();
  // processed_memory of the window: warp q-strided positions, lane-strided channels (5 per lane)
  float pm[4][5];
#pragma unroll
  for (int qi = 0; qi < 4; ++qi) {
    const int q = warp + qi * DEC_WARPS;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int a = lane + 32 * j;
      pm[qi][j] = (q < nw && a < A) ? __ldg(p.pmem + ((long long)b * p.T_in + start + q) * A + a) : 0.f;
    }
  }
  // previous / cumulative weights around the window (zero outside the sequence: conv padding)
  const int c0 = start - (KF - 1) / 2, ncat = nw + KF - 1;
  float* wprev = p.s.w_prev + (long long)b * p.T_in;
  float* wcum = p.s.w_cum + (long long)b * p.T_in;
  for (int i = tid; i < 2 * ncat; i += DEC_THREADS) {
    const int c = i / ncat, q = i - c * ncat, pos = c0 + q;
    float v = 0.f;
    if (pos >= 0 && pos < p.T_in) v = c == 0 ? wprev[pos] : wcum[pos];
    s.cat[c][q] = v;
  }
  // ---- query projection (model.py:92): pq = W_q h_att; warp per row, 75 float4 per row over the lanes
  {
    const float4* h4 = reinterpret_cast<const float4*>(h_att + b * R);
    const float4 hz = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 hv0 = __ldcg(h4 + lane), hv1 = __ldcg(h4 + lane + 32), hv2 = lane < 11 ? __ldcg(h4 + lane + 64) : hz;
    for (int r = warp; r < A; r += DEC_WARPS) {
      const float4* w4 = reinterpret_cast<const float4*>(p.w.wq + (long long)r * R);
      const float4 a0 = __ldg(w4 + lane), a1 = __ldg(w4 + lane + 32), a2 = lane < 11 ? __ldg(w4 + lane + 64) : hz;
      float acc = a0.x * hv0.x + a0.y * hv0.y + a0.z * hv0.z + a0.w * hv0.w;
      acc += a1.x * hv1.x + a1.y * hv1.y + a1.z * hv1.z + a1.w * hv1.w;
      acc += a2.x * hv2.x + a2.y * hv2.y + a2.z * hv2.z + a2.w * hv2.w;
      acc = warp_sum(acc);
      if (lane == 0) s.pq[r] = acc;
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  // ---- location_conv (model.py:57): loc[q][f] = sum_{c,k} w[f][c][k] * cat[c][q + k]
  for (int i = tid; i < nw * NF; i += DEC_THREADS) {
    const int q = i / NF, f = i - q * NF;
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
    for (int k = 0; k < KF; ++k) {
      acc0 = fmaf(s.w_loc[k][f], s.cat[0][q + k], acc0);
      acc1 = fmaf(s.w_loc[KF + k][f], s.cat[1][q + k], acc1);
    }
    s.loc[q][f] = acc0 + acc1;
  }
  __syncthreads();
  // ---- energies (model.py:94-97): e[q] = v . tanh(pq + location_dense(loc[q]) + processed_memory[q])
#pragma unroll
  for (int qi = 0; qi < 4; ++qi) {
    const int q = warp + qi * DEC_WARPS;
    if (q < nw) {
      float part = 0.f;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int a = lane + 32 * j;
        if (a < A) {
          float pa = 0.f;
#pragma unroll
          for (int f = 0; f < NF; ++f) pa = fmaf(s.w_ld[f][a], s.loc[q][f], pa);
          part = fmaf(sm.v[a], tanhf(s.pq[a] + pa + pm[qi][j]), part);
        }
      }
      part = warp_sum(part);
      if (lane == 0) s.e[q] = part;
    }
  }
  __syncthreads();
  // ---- softmax over the window (everything else is -inf -> weight 0, model.py:114-117); every warp
  //      evaluates it redundantly so that no further block barrier is needed before the context sum
  float mx, inv_sum;
  {
    const float e0 = lane < nw ? s.e[lane] : -INFINITY, e1 = lane + 32 < nw ? s.e[lane + 32] : -INFINITY;
    mx = fmaxf(e0, e1);
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, sft));
    const float x0 = lane < nw ? expf(e0 - mx) : 0.f, x1 = lane + 32 < nw ? expf(e1 - mx) : 0.f;
    inv_sum = 1.0f / warp_sum(x0 + x1);
  }
  // ---- context (model.py:118): ctx = sum_q w[q] * memory[start + q]; thread = (float4 column, q part)
  {
    const int c4 = tid % (E / 4), part = tid / (E / 4);
    if (part < CTXP) {
      const int per = (nw + CTXP - 1) / CTXP, q_lo = part * per, q_hi = min(nw, q_lo + per);
      const float4* mrow = reinterpret_cast<const float4*>(p.memory + ((long long)b * p.T_in + start) * E) + c4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
      for (int q = q_lo; q < q_hi; ++q) {
        const float4 m = __ldg(mrow + (long long)q * (E / 4));
        const float w = expf(s.e[q] - mx) * inv_sum;
        acc.x = fmaf(w, m.x, acc.x);
        acc.y = fmaf(w, m.y, acc.y);
        acc.z = fmaf(w, m.z, acc.z);
        acc.w = fmaf(w, m.w, acc.w);
      }
      *reinterpret_cast<float4*>(&s.ctxp[part][4 * c4]) = acc;
    }
  }
  // new attention_weights (zero outside the window), cumulative weights (model.py:424), alignments
  int ostart = 0, oend = -1;
  if (t > 0) window_bounds(t - 1, p.window, len, ostart, oend);
  for (int pos = ostart + tid; pos <= oend; pos += DEC_THREADS)
    if (pos < start || pos > end) wprev[pos] = 0.f;
  if (warp == 0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int q = lane + 32 * h;
      if (q < nw) {
        const float wv = expf(s.e[q] - mx) * inv_sum;
        wprev[start + q] = wv;
        wcum[start + q] += wv;
        if (p.align) p.align[((long long)b * p.max_steps + t) * p.T_in + start + q] = wv;
      }
    }
  }
  __syncthreads();
  for (int c = tid; c < E; c += DEC_THREADS) p.s.ctx[b * E + c] = s.ctxp[0][c] + s.ctxp[1][c] + s.ctxp[2][c];
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
  int u0, nu, pp0, npp, p20, np2;
  row_range(R, u0, nu);
  row_range(NPP, pp0, npp);
  row_range(R, p20, np2);

  // ---- resident weights: this CTA's rows of every matrix of the step
  for (int i = tid; i < nu * 4 * KIN; i += DEC_THREADS) {
    const int q = i / KIN, k = i - q * KIN;
    const int g = q / nu, u = q - g * nu;
    sm.w_att[q][k] = __ldg(p.w.w_att + (long long)(g * R + u0 + u) * KIN + k);
    sm.w_dec[q][k] = __ldg(p.w.w_dec + (long long)(g * R + u0 + u) * KIN + k);
  }
  for (int i = tid; i < nu * 4; i += DEC_THREADS) {
    const int g = i / nu, u = i - g * nu;
    sm.b_att[i] = __ldg(p.w.b_att + g * R + u0 + u);
    sm.b_dec[i] = __ldg(p.w.b_dec + g * R + u0 + u);
  }
  for (int i = tid; i < npp * KHC; i += DEC_THREADS)
    sm.w_pp[i / KHC][i % KHC] = __ldg(p.w.w_pp + (long long)pp0 * KHC + i);
  for (int i = tid; i < npp; i += DEC_THREADS) sm.b_pp[i] = __ldg(p.w.b_pp + pp0 + i);
  for (int i = tid; i < np2 * R; i += DEC_THREADS) sm.w_p2[i / R][i % R] = __ldg(p.w.w_pre2 + (long long)p20 * R + i);
  for (int i = tid; i < A; i += DEC_THREADS) sm.v[i] = __ldg(p.w.v + i);
  __syncthreads();

  unsigned int* bar = reinterpret_cast<unsigned int*>(p.s.done + 3);
  unsigned int bar_target = 0;
  int cur = 0;
  long long prof_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const long long prof_t0 = clock64();
  long long prof_prev = prof_t0;
  auto mark = [&](int slot) {          // thread 0's cycles since the previous mark go to `slot`
    if (p.prof != nullptr && tid == 0) {
      const long long now = clock64();
      prof_acc[slot] += now - prof_prev;
      prof_prev = now;
    }
  };
  for (int t = 0; t < p.max_steps; ++t) {
    float* h_att_cur = p.s.h_att + cur * p.B * R;
    float* h_att_nxt = p.s.h_att + (cur ^ 1) * p.B * R;
    float* h_dec_cur = p.s.h_dec + cur * p.B * R;
    float* h_dec_nxt = p.s.h_dec + (cur ^ 1) * p.B * R;
    // (1) attention_rnn (model.py:400-402): input [prenet | context], hidden h_att
    {
      const Seg segs[3] = {{p.s.pre, R, R}, {p.s.ctx, E, E}, {h_att_cur, R, R}};
      lstm_phase(sm, sm.w_att, sm.b_att, segs, h_att_nxt, p.s.c_att, p.B, u0, nu);
    }
    mark(0);
    grid_barrier(bar, bar_target);
    mark(1);
    // (2) query projection + location-sensitive attention, one CTA per utterance
    for (int b = blockIdx.x; b < p.B; b += G) {
      attention_phase(sm, p, h_att_nxt, b, t);
      __syncthreads();      // the scratch is reused (next utterance / staged inputs)
    }
    mark(2);
    grid_barrier(bar, bar_target);
    mark(3);
    // (3) decoder_rnn (model.py:425-428): input [h_att | context], hidden h_dec
    {
      const Seg segs[3] = {{h_att_nxt, R, R}, {p.s.ctx, E, E}, {h_dec_cur, R, R}};
      lstm_phase(sm, sm.w_dec, sm.b_dec, segs, h_dec_nxt, p.s.c_dec, p.B, u0, nu);
    }
    mark(4);
    grid_barrier(bar, bar_target);
    mark(5);
    // (4) [linear_projection | gate_layer | prenet layer 0 o projection] on hc = [h_dec | context]
    //     (model.py:436-441, 507, 132-135)
    {
      const Seg segs[2] = {{h_dec_nxt, R, R}, {p.s.ctx, E, E}};
      matvec_phase(
          sm, segs, KHC / 4, 1, p.B, [&](int, int r) { return &sm.w_pp[min(r, npp - 1)][0]; }, [](int, int) {},
          [&](int buf, int n0, int nb, int n_tiles, int kparts) {
            if (tid < npp * nb) {
              const int r = tid % npp, n = tid / npp, row = pp0 + r, b = n0 + n;
              const float v = sm.b_pp[r] + part_sum(sm, buf, 1, n_tiles, kparts, 0, r, n);
              if (row < M) {
                p.mel[((long long)b * p.max_steps + t) * M + row] = v;
              } else if (row == M) {
                p.gate[(long long)b * p.max_steps + t] = v;
                if (p.s.out_len[b] == 0) {          // stop test (model.py:524), per utterance
                  if (sigmoidf_exact(v) > p.gate_threshold) {
                    p.s.out_len[b] = t + 1;
                    atomicAdd(p.s.done, 1);
                  } else if (t + 1 == p.max_steps) {
                    p.s.out_len[b] = p.max_steps;     // model.py:526-528 "Reached max decoder steps"
                    atomicAdd(p.s.done + 1, 1);
                  }
                }
              } else if (t + 1 < p.max_steps) {
                // prenet layer 0 of the NEXT step; dropout p = 0.5 is always on -> mask * 2
                const int j = row - M - 1;
                const unsigned char d = p.drop[(((long long)(t + 1) * 2 + 0) * p.B + b) * R + j];
                p.s.p1[b * R + j] = fmaxf(v, 0.f) * (2.0f * (float)d);
              }
            }
          });
    }
    mark(6);
    grid_barrier(bar, bar_target);
    mark(7);
    // (5) prenet layer 1 of the next step
    if (t + 1 < p.max_steps) {
      const Seg segs[1] = {{p.s.p1, R, R}};
      matvec_phase(
          sm, segs, R / 4, 1, p.B, [&](int, int r) { return &sm.w_p2[min(r, np2 - 1)][0]; }, [](int, int) {},
          [&](int buf, int n0, int nb, int n_tiles, int kparts) {
            if (tid < np2 * nb) {
              const int r = tid % np2, n = tid / np2, row = p20 + r, b = n0 + n;
              const float v = part_sum(sm, buf, 1, n_tiles, kparts, 0, r, n);
              const unsigned char d = p.drop[(((long long)(t + 1) * 2 + 1) * p.B + b) * R + row];
              p.s.pre[b * R + row] = fmaxf(v, 0.f) * (2.0f * (float)d);
            }
          });
    }
    mark(8);
    grid_barrier(bar, bar_target);
    mark(9);
    cur ^= 1;
    // every utterance has fired its stop gate (or hit max_steps): uniform exit
    const volatile int* done = p.s.done;
    if (done[0] + done[1] >= p.B) {
      if (blockIdx.x == 0 && tid == 0) p.s.done[2] = t + 1;
      break;
    }
  }
  if (p.prof != nullptr && tid == 0) {
#pragma unroll
    for (int i = 0; i < 10; ++i) p.prof[blockIdx.x * 16 + i] = prof_acc[i];
    p.prof[blockIdx.x * 16 + 10] = clock64() - prof_t0;
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

static long long* g_dec_prof = nullptr;
void taco_set_prof(long long* p) { g_dec_prof = p; }

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
  FAC_REQUIRE(sms * MAXU >= R && sms * MAXPP >= NPP && sms * MAXP2 >= R,
              "taco_decoder: needs >= %d SMs, device has %d", (NPP + MAXPP - 1) / MAXPP, sms);
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
  p.prof = g_dec_prof;
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
