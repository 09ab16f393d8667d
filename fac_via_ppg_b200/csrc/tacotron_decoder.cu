// Autoregressive decoder of the PPG->Mel model as ONE persistent cooperative kernel
// (reference src/common/model.py:489-535 Decoder.inference, :387-442 decode, :100-121 Attention,
// :56-60 LocationLayer, :124-135 Prenet, src/common/utils.py:46-78 window mask).
//
// The reference spends ~40 kernel launches and >= 3 device->host syncs per output frame; here the
// whole loop (up to max_steps frames, all B utterances in lock step) is a single launch of one CTA per
// SM, and the CTAs are specialised:
//   * MATRIX CTAs (all but B of them) keep EVERY weight matrix of the step split by output row across
//     their shared memory for the whole sequence: the two LSTMCells (11.5 MB fp32),
//     [mel projection | stop gate | prenet layer 0 composed with the projection] and prenet layer 1.
//     prenet0 has no bias or nonlinearity between it and the projection (model.py:132-135, 436-438),
//     so W_pre0 (W_proj hc + b) is evaluated as one composed matrix.  A step is four batched mat-vec
//     phases: attention LSTM | decoder LSTM | projection+gate+prenet0 | prenet1.
//   * ATTENTION CTAs (one per utterance) own the location-sensitive attention of their utterance.
//     The query matrix W_q (180 KB) is resident in their shared memory, and everything that does not
//     depend on this step's attention-LSTM output -- location conv of the previous/cumulative weights,
//     location_dense, processed_memory, the encoder rows of the window (held in registers) -- is
//     prepared while the matrix CTAs run the other three phases.  On the critical path remain
//     W_q h, 41 x 150 tanh, a 41-way softmax and the 41 x 600 context sum.  Only the <= 2w+1 window
//     positions are evaluated: everything outside [t-w, t+w] is masked to -inf by utils.py:46-78,
//     i.e. has softmax weight exactly 0.
//   * phases hand over through lightweight grid barriers (one release-add + acquire spin); the two
//     barriers between matrix-only phases do not involve the attention CTAs.
//   * the stop decision (sigmoid(gate) > threshold) is taken on the device.
#include "fac_common.cuh"
#include <cuda_fp16.h>
#include <stdint.h>

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
constexpr int MIN_MATRIX_CTAS = 100;
constexpr int MAXU = 3;         // hidden units per matrix CTA (>= 100 matrix CTAs)
constexpr int MAXPP = 4;        // projection rows per matrix CTA
constexpr int MAXP2 = 3;        // prenet-1 rows per matrix CTA
constexpr int CHUNK = 8;        // utterances per staging buffer = one mma n-tile
constexpr int PASS = 2 * CHUNK; // utterances per arithmetic pass (both staging buffers)
constexpr int MAXW = 48;        // max window positions (2*window+1 <= 48)
constexpr int KS_LSTM = KIN + 8;        // halfs per resident LSTM weight row (+8: ldmatrix rows hit distinct banks)
constexpr int KP_PP = 912, KS_PP = KP_PP + 8;   // projection K padded to whole k16 steps
constexpr int KP_P2 = 304, KS_P2 = KP_P2 + 8;   // prenet-1 K padded
constexpr int XS = KIN + 8;             // floats per staged input row
constexpr float W_SCALE = 256.f;        // resident weights are stored times 2^8 so that their fp16 lo parts stay normal
constexpr int CTXP = 3;         // q-range split of the context sum
constexpr int QPP = MAXW / CTXP;  // window positions per part

struct MatSmem {                  // matrix CTAs
  // resident weights as IEEE-half hi/lo pairs (w * 2^8 = hi + lo to ~2^-22): the operands of mma.sync
  __half w_att[2][MAXU * 4][KS_LSTM];
  __half w_dec[2][MAXU * 4][KS_LSTM];
  __half w_pp[2][MAXPP][KS_PP];
  __half w_p2[2][MAXP2][KS_P2];
  float b_att[MAXU * 4], b_dec[MAXU * 4], b_pp[MAXPP];
  alignas(16) float part[DEC_WARPS][16][PASS];   // per-warp partial 16 x 16 tiles (K split over the warps)
  float sums[16][PASS];
  int n_done;
  unsigned int prof[16];
  alignas(16) float in[2][CHUNK][XS];   // staged input vectors of a chunk of utterances, double-buffered
};
struct AttSmem {                  // attention CTAs
  float wq[A][R];                 // query_layer weight, resident
  float S[MAXW][A];               // exp(2 (location_dense(location_conv(.)) + processed_memory)) of the coming window
  float v2[A + 2];                // -2 v
  float upq[A + 2];               // exp(2 W_q h_att)
  alignas(16) float h[R];         // attention_hidden of this step
  float e[MAXW];
  float wts[DEC_WARPS][MAXW];     // softmax weights, one copy per warp
  float cat[2][MAXW + KF - 1 + 2];
  int n_done;
  unsigned int prof[16];
  alignas(16) union {
    float loc[MAXW][NF];          // preparation
    float ctxp[CTXP][E];          // critical path
  } x;
};
static_assert(sizeof(MatSmem) <= 227 * 1024 && sizeof(AttSmem) <= 227 * 1024, "decoder shared memory");

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

// Barrier over `n` CTAs: monotonically increasing arrival counter (zeroed by the host).  The arrival is a
// release (every write of this CTA that bar.sync ordered before it is visible to whoever acquires the
// final count), the spin an acquire.  (Measured: several staggered pollers per CTA are slower, not faster.)
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& target, unsigned int n) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += n;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
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

// Staging = asynchronous copies (L2 -> shared memory, bypassing L1: the vectors were written by other CTAs in
// an earlier phase) of the concatenated input vectors; `mask` selects the segments (bit s = segment s).
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr)
               : "memory");
}
// D (16 x 8, fp32) += A (16 x 16, row-major halfs) * B (16 x 8, halfs; lane holds two k-pairs of one column)
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// fp32 pair -> IEEE-half hi/lo pairs (x = hi + lo to ~2^-22 relative)
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
// resident weight element: w * 2^8 -> hi/lo halfs
__device__ __forceinline__ void put_weight(__half* hi, __half* lo, float w) {
  w *= W_SCALE;
  const __half h = __float2half_rn(w);
  *hi = h;
  *lo = __float2half_rn(w - __half2float(h));
}

struct Prof {                 // thread 0's cycles between consecutive marks, per slot (accumulators in shared memory)
  unsigned int* acc;          // [16]: slots 0..9, total, 11..13 = fetch wait / arithmetic / epilogue of the mat-vec phases
  long long prev;
  bool on;
  __device__ __forceinline__ void init(bool enabled, unsigned int* smem_acc) {
    on = enabled && threadIdx.x == 0;
    acc = smem_acc;
    if (on) {
      for (int i = 0; i < 16; ++i) acc[i] = 0;
      prev = clock64();
    }
  }
  template <int SLOT>
  __device__ __forceinline__ void mark() {
    if (on) {
      const long long now = clock64();
      acc[SLOT] += (unsigned int)(now - prev);
      if (SLOT < 10) acc[10] += (unsigned int)(now - prev);
      prev = now;
    }
  }
  // sub-interval: does not move `prev` (the enclosing phase slot still gets the whole interval)
  template <int SLOT>
  __device__ __forceinline__ void sub(long long& from) {
    if (on) {
      const long long now = clock64();
      acc[SLOT] += (unsigned int)(now - from);
      from = now;
    }
  }
  __device__ __forceinline__ long long now() const { return on ? clock64() : 0; }
  __device__ __forceinline__ void flush(long long* out) {
    if (on)
      for (int i = 0; i < 16; ++i) out[blockIdx.x * 16 + i] = acc[i];
  }
};

// One batched mat-vec phase over a matrix CTA's resident rows (<= 16: the 4 gate rows of its LSTM units, or its
// projection / prenet rows) on the tensor cores.  The B utterances go in passes of 16 (two mma n-tiles = the two
// staging buffers); the fp32 input vectors of the next pass are copied into shared memory asynchronously behind
// the reduction and the epilogue of the current one.
// The K range is split over the 16 warps in whole k16 steps (a fixed split: the summation order of an output
// never depends on B); per step a warp loads the hi and lo weight fragments with ldmatrix, splits its slice
// of the inputs into half hi/lo pairs in registers and issues the three products hi*hi + lo*hi + hi*lo
// (fp32-grade: ~2^-21 relative).  The 16 partial 16 x 8 tiles meet in shared memory in fp32.
//   pre(n0, nb): called right after the copies are in flight (global loads the epilogue wants early)
//   epi(n0, nb): reads sm.sums[row][utterance] (scaled by W_SCALE)
//   early: segments of the first chunk that matvec_prefetch() already requested BEFORE the grid barrier
//          (vectors that were complete one barrier earlier), so that only the fresh segment is fetched after it
template <int NSEG>
__device__ __forceinline__ void stage_pass(MatSmem& sm, const Seg (&segs)[NSEG], int n0, int B, unsigned int mask) {
  // utterances [n0, n0 + 16) -> staging buffers 0 and 1, one cp.async group
  int base = 0;
  const int nb = min(PASS, B - n0);
#pragma unroll
  for (int s = 0; s < NSEG; ++s) {
    if (mask >> s & 1) {
      const int l4 = segs[s].len >> 2;
      for (int i = threadIdx.x; i < nb * l4; i += DEC_THREADS) {
        const int n = i / l4, k4 = i - n * l4;
        cp_async16(&sm.in[n >> 3][n & 7][base + 4 * k4], segs[s].ptr + (long long)(n0 + n) * segs[s].stride + 4 * k4);
      }
    }
    base += segs[s].len;
  }
  cp_async_commit();
}
template <int NSEG>
__device__ __forceinline__ void matvec_prefetch(MatSmem& sm, const Seg (&segs)[NSEG], int B, unsigned int early) {
  stage_pass(sm, segs, 0, B, early);
}
template <int NSEG, typename PreFn, typename EpiFn>
__device__ __forceinline__ void matvec_phase(MatSmem& sm, Prof& prof, const Seg (&segs)[NSEG], unsigned int early,
                                             const __half* w_hi, const __half* w_lo, int ks, int n_rows, int k_steps,
                                             int B, PreFn pre, EpiFn epi) {
  long long tp = prof.now();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int base = k_steps / DEC_WARPS, extra = k_steps % DEC_WARPS;
  const int my_steps = base + (warp < extra ? 1 : 0), my_first = warp * base + min(warp, extra);
  // ldmatrix: lane l addresses row (l & 7) + 8 * ((l >> 3) & 1) of the 8 x 8 block at k offset 8 * (l >> 4)
  int arow = (lane & 7) + ((lane >> 3) & 1) * 8;
  if (arow >= n_rows) arow = 0;                       // rows past the end: any valid row, results ignored
  const uint32_t a_off = (uint32_t)(arow * ks + (lane >> 4) * 8 + my_first * 16) * 2;
  const uint32_t a_hi = (uint32_t)__cvta_generic_to_shared(w_hi) + a_off;
  const uint32_t a_lo = (uint32_t)__cvta_generic_to_shared(w_lo) + a_off;
  stage_pass(sm, segs, 0, B, ~early);
  for (int n0 = 0; n0 < B; n0 += PASS) {
    const int nb = min(PASS, B - n0);
    pre(n0, nb);
    cp_async_wait<0>();
    __syncthreads();
    prof.sub<11>(tp);
    {
      // B fragments: column (utterance) lane >> 2 of each n-tile, k pairs 2 * (lane & 3) and + 8
      const int xo = my_first * 16 + 2 * (lane & 3);
      const float* xrow0 = &sm.in[0][min(lane >> 2, nb - 1)][xo];
      const float* xrow1 = &sm.in[1][min(lane >> 2, max(nb - 9, 0))][xo];
      const bool two = nb > CHUNK;
      // independent accumulation chains (hi*hi, lo*hi, hi*lo per n-tile), added in a fixed order at the end
      float acc[2][3][4];
#pragma unroll
      for (int i = 0; i < 24; ++i) (&acc[0][0][0])[i] = 0.f;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        if (i < my_steps) {
          uint32_t ah[4], al[4], bh[2], bl[2];
          ldmatrix_x4(ah, a_hi + i * 32);
          ldmatrix_x4(al, a_lo + i * 32);
          split_half2(*reinterpret_cast<const float2*>(xrow0 + i * 16), bh[0], bl[0]);
          split_half2(*reinterpret_cast<const float2*>(xrow0 + i * 16 + 8), bh[1], bl[1]);
          mma_f16(acc[0][0], ah, bh);
          mma_f16(acc[0][1], al, bh);
          mma_f16(acc[0][2], ah, bl);
          if (two) {
            split_half2(*reinterpret_cast<const float2*>(xrow1 + i * 16), bh[0], bl[0]);
            split_half2(*reinterpret_cast<const float2*>(xrow1 + i * 16 + 8), bh[1], bl[1]);
            mma_f16(acc[1][0], ah, bh);
            mma_f16(acc[1][1], al, bh);
            mma_f16(acc[1][2], ah, bl);
          }
        }
      }
      // accumulator layout: rows lane >> 2 and + 8, columns 2 * (lane & 3) + {0, 1} of the n-tile
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        if (nt == 0 || two) {
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = acc[nt][0][j] + (acc[nt][1][j] + acc[nt][2][j]);
          *reinterpret_cast<float2*>(&sm.part[warp][lane >> 2][nt * 8 + 2 * (lane & 3)]) = make_float2(v[0], v[1]);
          *reinterpret_cast<float2*>(&sm.part[warp][(lane >> 2) + 8][nt * 8 + 2 * (lane & 3)]) = make_float2(v[2], v[3]);
        }
      }
    }
    long long tq = tp;
    prof.sub<14>(tq);
    __syncthreads();
    prof.sub<15>(tq);
    // the staging buffers are free again: fetch the next pass behind the reduction and the epilogue
    if (n0 + PASS < B) stage_pass(sm, segs, n0 + PASS, B, ~0u);
    if (tid < 16 * PASS) {
      const int r = tid >> 4, c = tid & 15;
      if (c < nb) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < DEC_WARPS; ++w) a += sm.part[w][r][c];
        sm.sums[r][c] = a * (1.0f / W_SCALE);
      }
    }
    __syncthreads();
    prof.sub<12>(tp);
    epi(n0, nb);
    prof.sub<13>(tp);
  }
}

// One LSTMCell (model.py:400-402 / 425-428) for the units [u0, u0+nu) of this CTA and all B utterances.
__device__ __forceinline__ void lstm_phase(MatSmem& sm, Prof& prof, const __half (*w_s)[MAXU * 4][KS_LSTM],
                                           const float* bias_s, const Seg (&segs)[3], unsigned int early,
                                           float* h_next, float* c, int B, int u0, int nu,
                                           unsigned long long* h_tag = nullptr, unsigned int tag = 0) {
  const int tid = threadIdx.x;
  float c_old = 0.f;
  matvec_phase(
      sm, prof, segs, early, &w_s[0][0][0], &w_s[1][0][0], KS_LSTM, 4 * nu, KIN / 16, B,
      [&](int n0, int nb) {
        if (tid < nu * nb) c_old = c[(n0 + tid / nu) * R + u0 + tid % nu];   // only this thread ever touches it
      },
      [&](int n0, int nb) {
        if (tid < nu * nb) {                        // cell update: one thread per (unit, utterance)
          const int u = tid % nu, n = tid / nu, j = u0 + u, b = n0 + n;
          float gv[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) gv[g] = bias_s[g * nu + u] + sm.sums[g * nu + u][n];
          const float cn = sigmoidf_fast(gv[1]) * c_old + sigmoidf_fast(gv[0]) * tanhf_fast(gv[2]);
          c[b * R + j] = cn;
          const float h = sigmoidf_fast(gv[3]) * tanhf_fast(cn);
          h_next[b * R + j] = h;
          if (h_tag != nullptr)      // (value, tag) in ONE 8-byte store: whoever reads the tag has the value
            *reinterpret_cast<volatile unsigned long long*>(h_tag + b * R + j) =
                ((unsigned long long)tag << 32) | __float_as_uint(h);
        }
      });
}

__device__ __forceinline__ void window_bounds(int t, int window, int len, int& start, int& end) {
  // src/common/utils.py:70-74
  const int max_idx = len - 1;
  start = min(max(0, t - window), max_idx);
  end = min(t + window, max_idx);
}

__device__ __forceinline__ void row_range(int idx, int n_ctas, int n_rows, int& r0, int& nr) {
  r0 = (int)((long long)idx * n_rows / n_ctas);
  nr = (int)((long long)(idx + 1) * n_rows / n_ctas) - r0;
}

// ------------------------------------------------------------------------------------------ matrix CTAs
__device__ void matrix_role(const DecParams& p, MatSmem& sm, int mi, int GM) {
  const int tid = threadIdx.x;
  const unsigned int G = gridDim.x;
  int u0, nu, pp0, npp, p20, np2;
  row_range(mi, GM, R, u0, nu);
  row_range(mi, GM, NPP, pp0, npp);
  row_range(mi, GM, R, p20, np2);

  // ---- resident weights: this CTA's rows of every matrix of the step, as half hi/lo pairs; zero K padding
  for (int i = tid; i < nu * 4 * KS_LSTM; i += DEC_THREADS) {
    const int q = i / KS_LSTM, k = i - q * KS_LSTM;
    const int g = q / nu, u = q - g * nu;
    const long long src = (long long)(g * R + u0 + u) * KIN + k;
    put_weight(&sm.w_att[0][q][k], &sm.w_att[1][q][k], k < KIN ? __ldg(p.w.w_att + src) : 0.f);
    put_weight(&sm.w_dec[0][q][k], &sm.w_dec[1][q][k], k < KIN ? __ldg(p.w.w_dec + src) : 0.f);
  }
  for (int i = tid; i < nu * 4; i += DEC_THREADS) {
    const int g = i / nu, u = i - g * nu;
    sm.b_att[i] = __ldg(p.w.b_att + g * R + u0 + u);
    sm.b_dec[i] = __ldg(p.w.b_dec + g * R + u0 + u);
  }
  for (int i = tid; i < npp * KS_PP; i += DEC_THREADS) {
    const int q = i / KS_PP, k = i - q * KS_PP;
    put_weight(&sm.w_pp[0][q][k], &sm.w_pp[1][q][k], k < KHC ? __ldg(p.w.w_pp + (long long)(pp0 + q) * KHC + k) : 0.f);
  }
  for (int i = tid; i < npp; i += DEC_THREADS) sm.b_pp[i] = __ldg(p.w.b_pp + pp0 + i);
  for (int i = tid; i < np2 * KS_P2; i += DEC_THREADS) {
    const int q = i / KS_P2, k = i - q * KS_P2;
    put_weight(&sm.w_p2[0][q][k], &sm.w_p2[1][q][k], k < R ? __ldg(p.w.w_pre2 + (long long)(p20 + q) * R + k) : 0.f);
  }
  // the K padding of the staged inputs must hold finite values (it meets zero weights)
  for (int i = tid; i < 2 * CHUNK * XS; i += DEC_THREADS) (&sm.in[0][0][0])[i] = 0.f;
  __syncthreads();

  unsigned int* bar_all = reinterpret_cast<unsigned int*>(p.s.done + 3);
  unsigned int* bar_mat = reinterpret_cast<unsigned int*>(p.s.done + 4);
  unsigned int target_all = 0, target_mat = 0;
  int cur = 0;
  Prof prof;
  prof.init(p.prof != nullptr, sm.prof);
  for (int t = 0; t < p.max_steps; ++t) {
    float* h_att_cur = p.s.h_att + cur * p.B * R;
    float* h_att_nxt = p.s.h_att + (cur ^ 1) * p.B * R;
    float* h_dec_cur = p.s.h_dec + cur * p.B * R;
    float* h_dec_nxt = p.s.h_dec + (cur ^ 1) * p.B * R;
    // (1) attention_rnn (model.py:400-402): input [prenet | context], hidden h_att.  The context and the
    //     hidden state were requested before the barrier that ended the previous step (see below).
    {
      const Seg segs[3] = {{p.s.pre, R, R}, {p.s.ctx, E, E}, {h_att_cur, R, R}};
      lstm_phase(sm, prof, sm.w_att, sm.b_att, segs, t > 0 ? 6u : 0u, h_att_nxt, p.s.c_att, p.B, u0, nu,
                 reinterpret_cast<unsigned long long*>(p.s.h_tag), (unsigned int)(t + 1));
    }
    prof.mark<0>();
    // the attention CTAs pick h_att(t) up from the tagged copy without a barrier; this one (matrix CTAs only,
    // hidden behind the attention) orders the plain copy for the prefetch below
    grid_barrier(bar_mat, target_mat, GM);
    prof.mark<1>();
    // (2) attention CTAs at work; meanwhile fetch the two decoder_rnn inputs that are already complete
    const Seg segs_dec[3] = {{h_att_nxt, R, R}, {p.s.ctx, E, E}, {h_dec_cur, R, R}};
    matvec_prefetch(sm, segs_dec, p.B, 5u);
    grid_barrier(bar_all, target_all, G);      // context(t) complete
    prof.mark<3>();
    // (3) decoder_rnn (model.py:425-428): input [h_att | context], hidden h_dec
    lstm_phase(sm, prof, sm.w_dec, sm.b_dec, segs_dec, 5u, h_dec_nxt, p.s.c_dec, p.B, u0, nu);
    const Seg segs_pp[2] = {{h_dec_nxt, R, R}, {p.s.ctx, E, E}};
    matvec_prefetch(sm, segs_pp, p.B, 2u);     // the context, ahead of the barrier
    prof.mark<4>();
    grid_barrier(bar_mat, target_mat, GM);
    prof.mark<5>();
    // (4) [linear_projection | gate_layer | prenet layer 0 o projection] on hc = [h_dec | context]
    //     (model.py:436-441, 507, 132-135)
    {
      const Seg (&segs)[2] = segs_pp;
      const bool more = t + 1 < p.max_steps;
      unsigned char drop0 = 0;
      matvec_phase(
          sm, prof, segs, 2u, &sm.w_pp[0][0][0], &sm.w_pp[1][0][0], KS_PP, npp, KP_PP / 16, p.B,
          [&](int n0, int nb) {     // the dropout mask byte comes from DRAM: ask for it before the arithmetic
            if (tid < npp * nb) {
              const int row = pp0 + tid % npp;
              if (row > M && more)
                drop0 = p.drop[(((long long)(t + 1) * 2 + 0) * p.B + n0 + tid / npp) * R + row - M - 1];
            }
          },
          [&](int n0, int nb) {
            if (tid < npp * nb) {
              const int r = tid % npp, n = tid / npp, row = pp0 + r, b = n0 + n;
              const float v = sm.b_pp[r] + sm.sums[r][n];
              if (row < M) {
                p.mel[((long long)b * p.max_steps + t) * M + row] = v;
              } else if (row == M) {
                p.gate[(long long)b * p.max_steps + t] = v;
                if (p.s.out_len[b] == 0) {          // stop test (model.py:524), per utterance
                  if (sigmoidf_exact(v) > p.gate_threshold) {
                    p.s.out_len[b] = t + 1;
                    atomicAdd(p.s.done, 1);
                  } else if (!more) {
                    p.s.out_len[b] = p.max_steps;     // model.py:526-528 "Reached max decoder steps"
                    atomicAdd(p.s.done + 1, 1);
                  }
                }
              } else if (more) {
                // prenet layer 0 of the NEXT step; dropout p = 0.5 is always on -> mask * 2
                p.s.p1[b * R + row - M - 1] = fmaxf(v, 0.f) * (2.0f * (float)drop0);
              }
            }
          });
    }
    prof.mark<6>();
    grid_barrier(bar_mat, target_mat, GM);
    prof.mark<7>();
    // (5) prenet layer 1 of the next step; the stop counters are final since the last barrier
    if (tid == DEC_THREADS - 1) {
      const volatile int* done = p.s.done;
      sm.n_done = done[0] + done[1];
    }
    if (t + 1 < p.max_steps) {
      const Seg segs[1] = {{p.s.p1, R, R}};
      unsigned char drop1 = 0;
      matvec_phase(
          sm, prof, segs, 0u, &sm.w_p2[0][0][0], &sm.w_p2[1][0][0], KS_P2, np2, KP_P2 / 16, p.B,
          [&](int n0, int nb) {
            if (tid < np2 * nb)
              drop1 = p.drop[(((long long)(t + 1) * 2 + 1) * p.B + n0 + tid / np2) * R + p20 + tid % np2];
          },
          [&](int n0, int nb) {
            if (tid < np2 * nb) {
              const int r = tid % np2, n = tid / np2;
              const float v = sm.sums[r][n];
              p.s.pre[(n0 + n) * R + p20 + r] = fmaxf(v, 0.f) * (2.0f * (float)drop1);
            }
          });
    }
    prof.mark<8>();
    {
      // the next attention_rnn's context and hidden state are complete: request them ahead of the barrier
      // (a group left in flight by the last step is drained by the exit of the kernel)
      const Seg segs[3] = {{p.s.pre, R, R}, {p.s.ctx, E, E}, {h_att_nxt, R, R}};
      matvec_prefetch(sm, segs, p.B, 6u);
    }
    grid_barrier(bar_all, target_all, G);
    prof.mark<9>();
    cur ^= 1;
    if (sm.n_done >= p.B) break;     // every utterance has fired its stop gate (or hit max_steps)
  }
  cp_async_wait<0>();                // the last prefetch
  prof.flush(p.prof);
}

// --------------------------------------------------------------------------------------- attention CTAs
// Location-sensitive attention of utterance b (model.py:100-121, 56-60, 78-98).
//   prepare(t): everything of step t that only needs the attention weights of step t-1
//   critical(t): the part that needs h_att(t)
__device__ void attention_role(const DecParams& p, AttSmem& sm, int b) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned int G = gridDim.x;
  const int len = p.lengths[b];
  float* wprev = p.s.w_prev + (long long)b * p.T_in;
  float* wcum = p.s.w_cum + (long long)b * p.T_in;
  const float* pmem = p.pmem + (long long)b * p.T_in * A;
  const float* memory = p.memory + (long long)b * p.T_in * E;
  const int c4 = tid % (E / 4), cpart = tid / (E / 4);     // context sum: float4 column, q part (< CTXP active)

  for (int i = tid; i < A * R / 4; i += DEC_THREADS)
    reinterpret_cast<float4*>(&sm.wq[0][0])[i] = __ldg(reinterpret_cast<const float4*>(p.w.wq) + i);
  for (int i = tid; i < A; i += DEC_THREADS) sm.v2[i] = -2.0f * __ldg(p.w.v + i);
  __syncthreads();

  // Energies (model.py:94-97) e[q] = sum_a v[a] tanh(pq[a] + s[q][a]).  With tanh(x) = 1 - 2 / (exp(2x) + 1) and
  // exp(2 (pq + s)) = exp(2 pq) exp(2 s), the critical path keeps one multiply, one reciprocal and one FMA per
  // (q, a): exp(2 s) is prepared ahead, exp(2 pq) costs 150 exponentials, and the constant sum_a v[a] drops
  // out of the softmax.  Each exponent is clamped to +-43: exp(+-86) = 2^+-124 is still a normal fp32 (no 0 * inf),
  // a product that overflows saturates tanh correctly through the reciprocal (1 / inf = 0), and terms that
  // partly cancel (pq = 25, s = -24 -> tanh(1)) keep their sum, unlike a clamp at the saturation point of tanh.
  auto exp2x = [](float x) {
    x = fminf(fmaxf(x, -43.f), 43.f);
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 2.8853900817779268f));
    return r;
  };

  float4 mreg[QPP];      // encoder rows of this thread's share of the window
  int start = 0, nw = 0;
  Prof prof;
  prof.init(p.prof != nullptr, sm.prof);

  auto prepare = [&](int t) {
    int end;
    window_bounds(t, p.window, len, start, end);
    nw = end - start + 1;
    // previous / cumulative weights around the window (zero outside the sequence: conv padding)
    const int c0 = start - (KF - 1) / 2, ncat = nw + KF - 1;
    for (int i = tid; i < 2 * ncat; i += DEC_THREADS) {
      const int c = i / ncat, q = i - c * ncat, pos = c0 + q;
      float v = 0.f;
      if (pos >= 0 && pos < p.T_in) v = c == 0 ? wprev[pos] : wcum[pos];
      sm.cat[c][q] = v;
    }
    __syncthreads();
    // location_conv (model.py:57): loc[q][f] = sum_{c,k} w[c][k][f] * cat[c][q + k]
    for (int i = tid; i < nw * NF; i += DEC_THREADS) {
      const int q = i / NF, f = i - q * NF;
      float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
      for (int k = 0; k < KF; ++k) {
        acc0 = fmaf(__ldg(p.w.w_loc + k * NF + f), sm.cat[0][q + k], acc0);
        acc1 = fmaf(__ldg(p.w.w_loc + (KF + k) * NF + f), sm.cat[1][q + k], acc1);
      }
      sm.x.loc[q][f] = acc0 + acc1;
    }
    __syncthreads();
    // S[q][a] = location_dense(loc[q])[a] + processed_memory[q][a] (model.py:94-96): warp = 3 positions,
    // lane = 5 channels
    {
      float acc[3][5];
#pragma unroll
      for (int qi = 0; qi < 3; ++qi)
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const int q = warp + qi * DEC_WARPS, a = lane + 32 * j;
          acc[qi][j] = (q < nw && a < A) ? __ldg(pmem + (long long)(start + q) * A + a) : 0.f;
        }
#pragma unroll 4
      for (int f = 0; f < NF; ++f) {
        float wl[5], lq[3];
#pragma unroll
        for (int j = 0; j < 5; ++j) wl[j] = lane + 32 * j < A ? __ldg(p.w.w_ld_t + f * A + lane + 32 * j) : 0.f;
#pragma unroll
        for (int qi = 0; qi < 3; ++qi) lq[qi] = sm.x.loc[min(warp + qi * DEC_WARPS, MAXW - 1)][f];
#pragma unroll
        for (int qi = 0; qi < 3; ++qi)
#pragma unroll
          for (int j = 0; j < 5; ++j) acc[qi][j] = fmaf(wl[j], lq[qi], acc[qi][j]);
      }
#pragma unroll
      for (int qi = 0; qi < 3; ++qi)
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const int q = warp + qi * DEC_WARPS, a = lane + 32 * j;
          if (q < nw && a < A) sm.S[q][a] = exp2x(acc[qi][j]);
        }
    }
    // encoder outputs of the window -> registers (consumed by the context sum of step t)
    {
      const float4* mrow = reinterpret_cast<const float4*>(memory + (long long)start * E) + c4;
#pragma unroll
      for (int i = 0; i < QPP; ++i) {
        const int q = cpart * QPP + i;
        mreg[i] = (cpart < CTXP && q < nw) ? __ldg(mrow + (long long)q * (E / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    __syncthreads();
  };

  auto critical = [&](int t) {
    long long tp = prof.now();
    // query projection (model.py:92): pq = W_q h_att; warp per row, 75 float4 per row over the lanes
    {
      // h_att(t): spin on the (value, tag) pairs the matrix CTAs publish -- one-way latency, no barrier
      if (tid < R) {
        const volatile unsigned long long* src = reinterpret_cast<const volatile unsigned long long*>(p.s.h_tag) + b * R + tid;
        unsigned long long v;
        do {
          v = *src;
        } while ((unsigned int)(v >> 32) != (unsigned int)(t + 1));
        sm.h[tid] = __uint_as_float((unsigned int)v);
      }
      __syncthreads();
      const float4* h4 = reinterpret_cast<const float4*>(sm.h);
      const float4 hz = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 hv0 = h4[lane], hv1 = h4[lane + 32], hv2 = lane < 11 ? h4[lane + 64] : hz;
#pragma unroll 2
      for (int r = warp; r < A; r += DEC_WARPS) {
        const float4* w4 = reinterpret_cast<const float4*>(&sm.wq[r][0]);
        const float4 a0 = w4[lane], a1 = w4[lane + 32], a2 = lane < 11 ? w4[lane + 64] : hz;
        float acc = a0.x * hv0.x + a0.y * hv0.y + a0.z * hv0.z + a0.w * hv0.w;
        acc += a1.x * hv1.x + a1.y * hv1.y + a1.z * hv1.z + a1.w * hv1.w;
        acc += a2.x * hv2.x + a2.y * hv2.y + a2.z * hv2.z + a2.w * hv2.w;
        acc = warp_sum(acc);
        if (lane == 0) sm.upq[r] = exp2x(acc);
      }
    }
    __syncthreads();
    prof.sub<11>(tp);
    // energies up to a constant: e[q] = sum_a -2 v[a] / (exp(2 pq[a]) exp(2 s[q][a]) + 1)
    {
      float u[5], v2[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int a = min(lane + 32 * j, A - 1);
        u[j] = sm.upq[a];
        v2[j] = lane + 32 * j < A ? sm.v2[a] : 0.f;
      }
      float part[3];
#pragma unroll
      for (int qi = 0; qi < 3; ++qi) {
        const int q = min(warp + qi * DEC_WARPS, MAXW - 1);
        part[qi] = 0.f;
#pragma unroll
        for (int j = 0; j < 5; ++j)
          part[qi] = fmaf(v2[j], __fdividef(1.0f, fmaf(u[j], sm.S[q][min(lane + 32 * j, A - 1)], 1.0f)), part[qi]);
      }
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1)
#pragma unroll
        for (int qi = 0; qi < 3; ++qi) part[qi] += __shfl_xor_sync(0xffffffffu, part[qi], sft);
      if (lane < 3 && warp + lane * DEC_WARPS < nw) sm.e[warp + lane * DEC_WARPS] = lane == 0 ? part[0] : lane == 1 ? part[1] : part[2];
    }
    __syncthreads();
    prof.sub<12>(tp);
    // softmax over the window (everything else is -inf -> weight 0, model.py:114-117), evaluated by every
    // warp for itself
    {
      const float e0 = lane < nw ? sm.e[lane] : -INFINITY, e1 = lane + 32 < nw ? sm.e[lane + 32] : -INFINITY;
      float mx = fmaxf(e0, e1);
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, sft));
      const float x0 = lane < nw ? expf(e0 - mx) : 0.f, x1 = lane + 32 < nw ? expf(e1 - mx) : 0.f;
      const float inv_sum = 1.0f / warp_sum(x0 + x1);
      sm.wts[warp][lane] = x0 * inv_sum;
      if (lane + 32 < MAXW) sm.wts[warp][lane + 32] = x1 * inv_sum;
      __syncwarp();
    }
    // context (model.py:118): ctx = sum_q w[q] * memory[start + q], encoder rows already in registers
    if (cpart < CTXP) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < QPP; ++i) {
        const float w = sm.wts[warp][cpart * QPP + i];      // 0 past the end of the window
        acc.x = fmaf(w, mreg[i].x, acc.x);
        acc.y = fmaf(w, mreg[i].y, acc.y);
        acc.z = fmaf(w, mreg[i].z, acc.z);
        acc.w = fmaf(w, mreg[i].w, acc.w);
      }
      *reinterpret_cast<float4*>(&sm.x.ctxp[cpart][4 * c4]) = acc;
    }
    __syncthreads();
    prof.sub<13>(tp);
    for (int c = tid; c < E; c += DEC_THREADS) p.s.ctx[b * E + c] = sm.x.ctxp[0][c] + sm.x.ctxp[1][c] + sm.x.ctxp[2][c];
    if (tid == 0 && p.s.align_start) p.s.align_start[(long long)b * p.max_steps + t] = start;
    // new attention_weights (zero outside the window), cumulative weights (model.py:424), alignments
    int ostart = 0, oend = -1;
    if (t > 0) window_bounds(t - 1, p.window, len, ostart, oend);
    for (int pos = ostart + tid; pos <= oend; pos += DEC_THREADS)
      if (pos < start || pos >= start + nw) wprev[pos] = 0.f;
    if (tid < nw) {
      const float wv = sm.wts[0][tid];
      wprev[start + tid] = wv;
      wcum[start + tid] += wv;
      if (p.align) p.align[((long long)b * p.max_steps + t) * p.T_in + start + tid] = wv;
      if (p.s.align_win) p.s.align_win[((long long)b * p.max_steps + t) * (2 * p.window + 1) + tid] = wv;
    }
  };

  unsigned int* bar_all = reinterpret_cast<unsigned int*>(p.s.done + 3);
  unsigned int target_all = 0;
  int cur = 0;
  prepare(0);
  for (int t = 0; t < p.max_steps; ++t) {
    prof.mark<0>();
    critical(t);
    prof.mark<2>();
    grid_barrier(bar_all, target_all, G);      // context(t) complete -> matrix CTAs
    prof.mark<3>();
    if (t + 1 < p.max_steps) prepare(t + 1);
    prof.mark<4>();
    grid_barrier(bar_all, target_all, G);      // end of step
    prof.mark<9>();
    cur ^= 1;
    if (tid == 0) {
      const volatile int* done = p.s.done;
      sm.n_done = done[0] + done[1];
    }
    __syncthreads();
    if (sm.n_done >= p.B) {
      if (blockIdx.x == 0 && tid == 0) p.s.done[2] = t + 1;
      break;
    }
  }
  prof.flush(p.prof);
}

__global__ void __launch_bounds__(DEC_THREADS, 1) taco_decoder_kernel(const DecParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if ((int)blockIdx.x < p.B)
    attention_role(p, *reinterpret_cast<AttSmem*>(smem_raw), blockIdx.x);
  else
    matrix_role(p, *reinterpret_cast<MatSmem*>(smem_raw), blockIdx.x - p.B, gridDim.x - p.B);
}

__global__ void __launch_bounds__(DEC_THREADS, 1) grid_barrier_selftest_kernel(unsigned int* counter, int iters) {
  unsigned int target = 0;
  for (int i = 0; i < iters; ++i) grid_barrier(counter, target, gridDim.x);
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
  FAC_REQUIRE(sms > MIN_MATRIX_CTAS, "taco_decoder: needs more than %d SMs, device has %d", MIN_MATRIX_CTAS, sms);
  FAC_REQUIRE(B <= sms - MIN_MATRIX_CTAS,
              "taco_decoder: at most %d utterances per launch on this device (one attention CTA each next to >= %d "
              "matrix CTAs); split the batch", sms - MIN_MATRIX_CTAS, MIN_MATRIX_CTAS);
  const size_t smem = sizeof(MatSmem) > sizeof(AttSmem) ? sizeof(MatSmem) : sizeof(AttSmem);
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
