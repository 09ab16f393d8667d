// Autoregressive decoder of the PPG->Mel model as ONE persistent cooperative kernel
// (reference src/common/model.py:489-535 Decoder.inference, :387-442 decode, :100-121 Attention,
// :56-60 LocationLayer, :124-135 Prenet, src/common/utils.py:46-78 window mask).
//
// The reference spends ~40 kernel launches and >= 3 device->host syncs per output frame; here the
// whole loop (up to max_steps frames, all B utterances in lock step) is a single launch of one CTA per
// SM, and the CTAs are specialised:
//   * MATRIX CTAs (all but B of them) keep the weight matrices of the step in shared memory for the whole
//     sequence, ONE matrix per CTA, split by output row: the two LSTMCells (11.5 MB as half hi/lo pairs, the
//     bulk of the CTAs), [mel projection | stop gate | prenet layer 0 composed with the projection] (13 CTAs)
//     and prenet layer 1 (11 CTAs).  prenet0 has no bias or nonlinearity between it and the projection
//     (model.py:132-135, 436-438), so W_pre0 (W_proj hc + b) is evaluated as one composed matrix.  A step is
//     a chain of four batched mat-vecs: attention LSTM | decoder LSTM | projection+gate+prenet0 | prenet1.
//   * ATTENTION CTAs (one per utterance) own the location-sensitive attention of their utterance.
//     The query matrix W_q (180 KB) is resident in their shared memory, and everything that does not
//     depend on this step's attention-LSTM output -- location conv of the previous/cumulative weights,
//     location_dense, processed_memory, the encoder rows of the window (held in registers) -- is
//     prepared while the matrix CTAs run the other three mat-vecs (location conv as a sliding window,
//     location dense on mma.sync).  On the critical path remain
//     W_q h, 41 x 150 tanh, a 41-way softmax and the 41 x 600 context sum.  Only the <= 2w+1 window
//     positions are evaluated: everything outside [t-w, t+w] is masked to -inf by utils.py:46-78,
//     i.e. has softmax weight exactly 0.
//   * DATAFLOW, no grid barrier (profiles/r2_grid_handover_bench.txt: a barrier costs 2.5 k cycles before
//     a byte moves, 3.8 k with the payload of a phase; a tagged word is seen after ~490): every vector that
//     crosses CTAs (prenet output, attention hidden, context, decoder hidden, prenet layer-0 output) lives in
//     global memory as (value, version) 8-byte words, double-buffered by version parity.  A consumer polls
//     the words it needs until they carry the version it expects -- data and flag arrive in ONE store, no
//     fence, no atomic.  Every mat-vec needs the FULL input vector, so the data dependencies themselves
//     keep the CTAs within one step of each other (which is what makes two buffers enough).
//   * SPLIT mat-vecs: the K range of a matrix is cut at the vector boundaries (segments padded to whole
//     k16 steps).  The products with vectors that are complete EARLY (the hidden states and the previous
//     context) are accumulated while the CTA would otherwise wait for a hand-over; after the awaited vector
//     arrives only its own segment is left: 304 of 1216 columns for the attention LSTM, 608 for the decoder
//     LSTM, 304 of 912 for the projection.  A CTA that runs one mat-vec per step is idle three quarters of
//     it, so this early work is free -- with every CTA running all four mat-vecs (round 1's layout) it was not.
//   * utterances go through a matrix CTA in passes of 8 or 16 (one or two mma n-tiles); a producer publishes
//     pass by pass and a consumer starts on a pass as soon as its words are in, so a large batch pipelines
//     through the chain.  The arithmetic of an utterance never depends on the batch it is in.
//   * the stop decision (sigmoid(gate) > threshold) is taken on the device and travels as a tagged word too.
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
  long long* prof;            // optional [grid][32] cycle counters (tools/decoder_cycle_breakdown.py)
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
// matrix CTAs are specialised by matrix: N_PP_CTAS hold [projection | gate | prenet 0], N_P2_CTAS prenet layer 1,
// the rest is split evenly between the two LSTMCells (>= 38 CTAs each: <= 8 hidden units = 32 gate rows per CTA)
constexpr int N_PP_CTAS = 13, N_P2_CTAS = 11;
// Two builds of the kernel: utterances go through a matrix CTA in passes of 8 (one mma n-tile; up to 32 resident
// rows) or, for 8 < B <= 36, in passes of 16 (two n-tiles; the staging needs the room of 4 rows, and B <= 36 keeps
// an LSTM CTA at <= 7 units = 28 rows).  B <= 48 in all.
template <bool P16>
struct Cfg {
  static constexpr int PASS = P16 ? 16 : 8;
  static constexpr int MAXROWS = P16 ? 28 : 32;
  static constexpr int PSHIFT = P16 ? 4 : 3;
};
constexpr int MAXPASS = 6;       // passes of 8; passes of 16: 3
constexpr int MAX_B_P16 = 36;
constexpr int MAXW = 48;        // max window positions (2*window+1 <= 48)
// K segments: every vector padded to whole k16 steps, so that a mat-vec can be cut at the vector boundaries
constexpr int SEG = 304, SEGC = 608;              // a 300-vector / the 600-float context
constexpr int ST_SEG = SEG / 16, ST_SEGC = SEGC / 16;
constexpr int KP_LSTM = SEG + SEGC + SEG, KS_LSTM = KP_LSTM + 8;  // halfs per resident LSTM weight row (+8: ldmatrix rows hit distinct banks)
constexpr int KP_PP = SEG + SEGC, KS_PP = KP_PP + 8;             // projection: [h_dec | context]
constexpr int KP_P2 = SEG, KS_P2 = KP_P2 + 8;                    // prenet layer 1
// staging row of one utterance: the context plus one 300-vector (what is staged early and what is awaited never
// live at the same time, so they share columns; each role lays its slots out itself)
constexpr int XSTRIDE = SEGC + SEG + 8;
constexpr float W_SCALE = 256.f;        // resident weights are stored times 2^8 so that their fp16 lo parts stay normal
constexpr int CTXP = 3;         // q-range split of the context sum
constexpr int QPP = MAXW / CTXP;  // window positions per part

// Exchange area (fac_taco_decoder_state::xchg): (value, version) words, two copies by version parity.  Version v
// of a vector is what step v consumes as the state BEFORE the step: pre_v, hatt_v, ctx_v, hdec_v feed step v;
// step v produces hatt_{v+1}, ctx_{v+1}, hdec_{v+1}, p1_{v+1}, pre_{v+1}.  Version 0 is the zero fill of the host
// (model.py:304-335 initialises every state to zero; prenet(0) = 0 because the prenet has no bias).
enum Vec { V_PRE = 0, V_HATT = 1, V_HDEC = 2, V_P1 = 3, V_CTX = 4 };
constexpr int XCHG_WORDS = 4 * R + E;   // per utterance and parity
__device__ __forceinline__ unsigned long long* xchg_vec(unsigned long long* base, int B, int kind, unsigned int version) {
  return base + ((size_t)(version & 1u) * XCHG_WORDS + (size_t)kind * R) * B;
}
__device__ __forceinline__ unsigned long long tagged(float v, unsigned int version) {
  return ((unsigned long long)version << 32) | __float_as_uint(v);
}
// Behind the two copies: one arrival counter per vector kind, 32 bytes apart.  A producer CTA adds 1 after its
// stores (relaxed, NO fence: the words carry their own version).  A CTA that fetches a vector EARLY, long before
// it is complete, waits on that one counter with one thread instead of sweeping the payload with 512 --
// measured: ~140 CTAs polling a payload for thousands of cycles congest the L2 slices of those lines and delay
// the very stores they wait for (5 k cycles per hand-over instead of 1 k).  The one role that is next on the
// critical path polls the words themselves: one hop instead of two.  The counter is only a hint; what a consumer
// accepts is decided by the version in each word.
constexpr int XCHG_HINTS = 64;          // 8-byte words reserved for the counters (FAC_TACO_XCHG_HINTS)
static_assert(XCHG_HINTS * 8 >= 5 * 32 && XCHG_HINTS == FAC_TACO_XCHG_HINTS && XCHG_WORDS == FAC_TACO_XCHG_WORDS, "exchange area");
__device__ __forceinline__ unsigned int* xchg_hint(unsigned long long* base, int B, int kind) {
  return reinterpret_cast<unsigned int*>(base + (size_t)2 * XCHG_WORDS * B) + kind * 8;
}
__device__ __forceinline__ void hint_arrive(unsigned int* counter) {
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}
constexpr int SPIN_LIMIT = 1 << 17;     // polls before a hand-over is declared dead (>= 30 ms; a step takes ~10 us)
constexpr int DONE_ABORT = 7;           // done[7] != 0: a hand-over timed out, every CTA leaves the loop

constexpr int PROF_SLOTS = 32;
constexpr int MAT_WARPS = 8;      // warps that multiply (the others wait at the barrier)
template <bool P16>
struct MatSmem {                  // matrix CTAs
  static constexpr int PASS = Cfg<P16>::PASS, MAXROWS = Cfg<P16>::MAXROWS;
  // resident rows of this CTA's matrix as IEEE-half hi/lo pairs (w * 2^8 = hi + lo to ~2^-22): the operands of
  // mma.sync; the row stride is the matrix's own (KS_LSTM / KS_PP / KS_P2)
  __half w[2][MAXROWS * KS_LSTM];
  float bias[32];
  alignas(16) float part[MAT_WARPS][32][PASS];        // per-warp partial tiles (K split over the warps)
  float sums[32][PASS];
  int n_done, done_count, ok;
  unsigned int prof[PROF_SLOTS];
  alignas(16) float xs[PASS][XSTRIDE];  // staged input vectors of a pass of utterances
};
constexpr int LOC_LD = NF + 8;    // halfs per row of the location features (+8: ldmatrix rows hit distinct banks)
struct AttSmem {                  // attention CTAs
  float wq[A][R];                 // query_layer weight, resident
  float S[MAXW][A];               // exp(2 (location_dense(location_conv(.)) + processed_memory)) of the coming window
  float v2[A + 2];                // -2 v
  float upq[A + 2];               // exp(2 W_q h_att)
  alignas(16) float h[R];         // attention_hidden of this step
  float e[MAXW];
  float wts[DEC_WARPS][MAXW];     // softmax weights, one copy per warp
  float cat[2][MAXW + KF - 1 + 2];
  int n_done, ok;
  unsigned int prof[PROF_SLOTS];
  alignas(16) union {
    __half loc[2][MAXW][LOC_LD];  // preparation: location_conv output as half hi/lo pairs (mma operand)
    float ctxp[CTXP][E];          // critical path
  } x;
};
static_assert(sizeof(MatSmem<false>) <= 227 * 1024 && sizeof(MatSmem<true>) <= 227 * 1024 && sizeof(AttSmem) <= 227 * 1024,
              "decoder shared memory");

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  // (no memory clobber: it only ever reads the resident weights, which never change after the prologue)
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
// the same for a tile that other threads have just written (the attention's location features)
__device__ __forceinline__ void ldmatrix_x4_fresh(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr)
               : "memory");
}
// D (16 x 8, fp32) += A (16 x 16, row-major halfs) * B (16 x 8, halfs; lane holds two k-pairs of one column)
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
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
  unsigned int* acc;          // [PROF_SLOTS]: see tools/decoder_cycle_breakdown.py; [10] = total of all marks
  long long prev;
  bool on;
  __device__ __forceinline__ void init(bool enabled, unsigned int* smem_acc) {
    on = enabled && threadIdx.x == 0;
    acc = smem_acc;
    if (on) {
      for (int i = 0; i < PROF_SLOTS; ++i) acc[i] = 0;
      prev = clock64();
    }
  }
  template <int SLOT>
  __device__ __forceinline__ void mark() {
    if (on) {
      const long long now = clock64();
      acc[SLOT] += (unsigned int)(now - prev);
      acc[10] += (unsigned int)(now - prev);
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
      for (int i = 0; i < PROF_SLOTS; ++i) out[blockIdx.x * PROF_SLOTS + i] = acc[i];
  }
};

// ---- hand-over primitives
__device__ __forceinline__ void ld_tagged2(const unsigned long long* src, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(src) : "memory");
}
__device__ __forceinline__ unsigned long long ld_tagged(const unsigned long long* src) {
  unsigned long long a;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(a) : "l"(src) : "memory");
  return a;
}
__device__ __forceinline__ void st_tagged(unsigned long long* dst, float v, unsigned int version) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst), "l"(tagged(v, version)) : "memory");
}
// a spinning thread gives up when somebody else already did (checked now and then) or after SPIN_LIMIT polls
__device__ __forceinline__ bool spin_dead(int& spins, int* done) {
  ++spins;
  if ((spins & 255) == 0 && *reinterpret_cast<volatile int*>(done + DONE_ABORT) != 0) return true;
  if (spins > SPIN_LIMIT) {
    *reinterpret_cast<volatile int*>(done + DONE_ABORT) = 1;
    return true;
  }
  return false;
}

// Poll-fetch of version `ver` of a 300- or 600-float vector for the utterances [n0, n0 + nb) into staging slot
// `slot`.  With `hint` (the arrival counter of the vector) thread 0 first waits until all producers have announced
// their stores, so that the sweep is normally a single round.  The sweep: 2 .. 16 warps per utterance, a lane owns
// the 16-byte pieces (two tagged words) lane, lane + 32, ... of its warp's share -- all of them in flight at
// once, no index arithmetic beyond immediates --, keeps re-reading the ones whose versions do not match yet and
// drops the values into shared memory.
// Ends with a CTA barrier; false = the hand-over was declared dead (every thread of the CTA agrees).
template <bool P16>
__device__ __noinline__ bool fetch(MatSmem<P16>& sm, const unsigned long long* vec, int len, unsigned int ver, int n0, int nb,
                                   int slot, const unsigned int* hint, unsigned int hint_target, int* done) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool dead = false;
  int spins = 0;
  if (hint != nullptr) {              // wait quietly: one thread, one word
    if (threadIdx.x == 0) {
      unsigned int seen;
      do {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(hint) : "memory");
      } while (seen < hint_target && !(dead = spin_dead(spins, done)));
    }
    __syncthreads();
  }
  constexpr int FETCH_ROUNDS = P16 ? 10 : 5;   // 300 pieces (the context) of one utterance on one / two warps
  // warps per utterance: 1, 2, 4, 8 or 16
  const int sh = nb > 8 ? 0 : nb > 4 ? 1 : nb > 2 ? 2 : nb > 1 ? 3 : 4;
  const int n = warp >> sh, part = warp & ((1 << sh) - 1);
  const int half = len >> 1, share = (half + (1 << sh) - 1) >> sh;     // pieces per warp
  const int first = part * share, count = n < nb ? min(share, half - first) : 0;
  const unsigned long long* src = vec + (size_t)(n0 + n) * len + 2 * (first + lane);
  float* dst = &sm.xs[n < nb ? n : 0][slot + 2 * (first + lane)];
  unsigned int need = 0;
#pragma unroll
  for (int j = 0; j < FETCH_ROUNDS; ++j)
    if (lane + 32 * j < count) need |= 1u << j;
  while (need != 0 && !dead) {
    unsigned long long a[FETCH_ROUNDS], b[FETCH_ROUNDS];
#pragma unroll
    for (int j = 0; j < FETCH_ROUNDS; ++j)
      if (need >> j & 1) ld_tagged2(src + 64 * j, a[j], b[j]);
#pragma unroll
    for (int j = 0; j < FETCH_ROUNDS; ++j)
      if ((need >> j & 1) && (unsigned int)(a[j] >> 32) == ver && (unsigned int)(b[j] >> 32) == ver) {
        *reinterpret_cast<float2*>(dst + 64 * j) =
            make_float2(__uint_as_float((unsigned int)a[j]), __uint_as_float((unsigned int)b[j]));
        need &= ~(1u << j);
      }
    if (need != 0) dead = spin_dead(spins, done);
  }
  return !__syncthreads_or(dead);
}

// One piece of a mat-vec: `steps` k16 steps of the resident rows, weight columns from `koff`, inputs from staging
// slot column `slot`.
struct SegDesc {
  int koff, slot, steps;
};

// Partial products of a matrix CTA's resident rows (<= 32 = two m-tiles) with the staged vectors of one pass of
// <= 8 / 16 utterances (one / two n-tiles), on the tensor cores.  These pieces are small (19 .. 57 k16 steps) and the CTA's
// instruction issue and latencies, not the tensor pipe, bound them.  The steps of the one or two listed segments
// are dealt round-robin to at most MAT_WARPS slots (a fixed deal: the summation order of an output never depends
// on B), <= 3 steps per slot where possible; a slot is a pair of warps that share its steps' tiles; per step a warp loads the hi
// and lo weight fragments with ldmatrix, splits its slice of the inputs into half hi/lo pairs in registers and
// issues the three products hi*hi + lo*hi + hi*lo per m-tile (fp32-grade: ~2^-21 relative).  The partial tiles
// meet in shared memory in fp32; returns, in thread (row * PASS + utterance), the sum
// scaled back by 1 / W_SCALE (with `to_sums` also sm.sums[row][utterance] = sum + add).  Two CTA barriers inside.
template <bool P16>
__device__ __noinline__ float mat_part(MatSmem<P16>& sm, int ks, int n_rows, const SegDesc s0, const SegDesc s1, int nb,
                                       float add = 0.f, bool to_sums = false) {
  constexpr int PSHIFT = Cfg<P16>::PSHIFT, PASS = Cfg<P16>::PASS;
  constexpr int MT = P16 ? 2 : 1;                    // m-tiles per warp
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int total = s0.steps + s1.steps;
  const int n_slots = min(MAT_WARPS, (total + 2) / 3);
  // warp = (deal slot, half): passes of 8 split the two m-tiles over the halves, passes of 16 the two n-tiles
  const int slot = warp & (MAT_WARPS - 1), sel = warp >> 3;
  const bool m2 = n_rows > 16, n2 = P16 && nb > 8;
  if (slot < n_slots && (sel == 0 || (P16 ? n2 : m2))) {
    // ldmatrix: lane l addresses row (l & 7) + 8 * ((l >> 3) & 1) of the 8 x 8 block at k offset 8 * (l >> 4)
    const int arow = (lane & 7) + ((lane >> 3) & 1) * 8;
    const int mt0 = P16 ? 0 : sel;                    // first (P16: of two) m-tile of this warp
    const int arow0 = mt0 * 16 + arow < n_rows ? mt0 * 16 + arow : 0, arow1 = 16 + arow < n_rows ? 16 + arow : 0;   // past the end: any valid row
    const uint32_t w_hi = (uint32_t)__cvta_generic_to_shared(&sm.w[0][0]), w_lo = (uint32_t)__cvta_generic_to_shared(&sm.w[1][0]);
    const uint32_t off0 = (uint32_t)(arow0 * ks + (lane >> 4) * 8) * 2, off1 = (uint32_t)(arow1 * ks + (lane >> 4) * 8) * 2;
    // B fragments: column (utterance) lane >> 2 of the n-tile, k pairs 2 * (lane & 3) and + 8
    const int nt = P16 ? sel : 0;
    const float* xrow = &sm.xs[nt * 8 + min(lane >> 2, max(nb - nt * 8 - 1, 0))][2 * (lane & 3)];
    // independent accumulation chains (hi*hi, lo*hi, hi*lo per m-tile), added in a fixed order at the end
    float acc[MT][3][4];
#pragma unroll
    for (int i = 0; i < 12 * MT; ++i) (&acc[0][0][0])[i] = 0.f;
#pragma unroll 1
    for (int s = slot; s < total; s += n_slots) {
      int ls = s, koff = s0.koff, xo = s0.slot;
      if (ls >= s0.steps) {
        ls -= s0.steps;
        koff = s1.koff;
        xo = s1.slot;
      }
      koff = (koff + 16 * ls) * 2;
      xo += 16 * ls;
      uint32_t ah[MT][4], al[MT][4], bh[2], bl[2];
      ldmatrix_x4(ah[0], w_hi + off0 + koff);
      ldmatrix_x4(al[0], w_lo + off0 + koff);
      if (P16 && m2) {
        ldmatrix_x4(ah[MT - 1], w_hi + off1 + koff);
        ldmatrix_x4(al[MT - 1], w_lo + off1 + koff);
      }
      split_half2(*reinterpret_cast<const float2*>(xrow + xo), bh[0], bl[0]);
      split_half2(*reinterpret_cast<const float2*>(xrow + xo + 8), bh[1], bl[1]);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
        if (mt == 0 || m2) {
          mma_f16(acc[mt][0], ah[mt], bh);
          mma_f16(acc[mt][1], al[mt], bh);
          mma_f16(acc[mt][2], ah[mt], bl);
        }
    }
    // accumulator layout: rows lane >> 2 and + 8, columns 2 * (lane & 3) + {0, 1} of the n-tile
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
      if (mt == 0 || m2) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = acc[mt][0][j] + (acc[mt][1][j] + acc[mt][2][j]);
        const int row = (P16 ? mt : mt0) * 16 + (lane >> 2);
        *reinterpret_cast<float2*>(&sm.part[slot][row][nt * 8 + 2 * (lane & 3)]) = make_float2(v[0], v[1]);
        *reinterpret_cast<float2*>(&sm.part[slot][row + 8][nt * 8 + 2 * (lane & 3)]) = make_float2(v[2], v[3]);
      }
  }
  __syncthreads();
  float a = 0.f;
  if ((tid & (PASS - 1)) < nb && (tid >> PSHIFT) < n_rows) {
    const int r = tid >> PSHIFT, c = tid & (PASS - 1);
    float v[MAT_WARPS];
#pragma unroll
    for (int w = 0; w < MAT_WARPS; ++w) v[w] = w < n_slots ? sm.part[w][r][c] : 0.f;   // all loads in flight, fixed order of
#pragma unroll
    for (int w = 0; w < MAT_WARPS; ++w) a += v[w];                                     // the additions
    a *= 1.0f / W_SCALE;
    if (to_sums) sm.sums[r][c] = a + add;             // the epilogue's input: early columns + this piece
  }
  __syncthreads();                                    // sm.part is free again, sm.sums complete
  return a;
}

// per-pass register of the threads that hold a reduced (row, utterance) element
struct PassReg {
  float v[MAXPASS];
  __device__ __forceinline__ float get(int ps) const {
    return ps == 0 ? v[0] : ps == 1 ? v[1] : ps == 2 ? v[2] : ps == 3 ? v[3] : ps == 4 ? v[4] : v[5];
  }
  __device__ __forceinline__ void set(int ps, float x) {
    if (ps == 0) v[0] = x;
    else if (ps == 1) v[1] = x;
    else if (ps == 2) v[2] = x;
    else if (ps == 3) v[3] = x;
    else if (ps == 4) v[4] = x;
    else v[5] = x;
  }
};

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
// Every matrix CTA owns a slice of the rows of ONE matrix for the whole sequence and runs one mat-vec per step:
//   role A  attention_rnn   (model.py:400-402)  waits for pre_t      publishes hatt_{t+1}
//   role D  decoder_rnn     (model.py:425-428)  waits for ctx_{t+1}  publishes hdec_{t+1}
//   role P  [linear_projection | gate_layer | prenet layer 0 o projection] (model.py:436-441, 507, 132-135)
//                                               waits for hdec_{t+1} publishes mel, gate, stop count, p1_{t+1}
//   role Q  prenet layer 1  (model.py:132-135)  waits for p1_{t+1}   publishes pre_{t+1}
// A CTA is idle three quarters of a step, so the columns of its matrix that meet vectors known EARLY (hidden
// states, the previous context) are multiplied while it waits; on the critical path of a step stay four times
// [sweep of the awaited vector (polled directly) | its 19 or 38 k16 steps | epilogue] plus the attention.

// the stop counter of step t (published by the CTA that owns the gate row): false = leave the loop
__device__ __forceinline__ bool step_continues(const DecParams& p, int* n_done_s, int* ok_s, unsigned int v1, Prof& prof) {
  if (threadIdx.x == 0) {
    const unsigned long long* done_word = reinterpret_cast<const unsigned long long*>(p.s.done + 4);
    int spins = 0;
    bool dead = false;
    unsigned long long w;
    do {
      w = ld_tagged(done_word);
    } while ((unsigned int)(w >> 32) != v1 && !(dead = spin_dead(spins, p.s.done)));
    *n_done_s = (int)(unsigned int)w;
    *ok_s = !dead;
  }
  __syncthreads();
  prof.mark<12>();
  const bool go = *ok_s && *n_done_s < p.B;
  __syncthreads();                   // the two words are rewritten next step
  return go;
}

// resident rows: `n_rows` rows of a [rows][k_src] fp32 matrix -> padded-segment half hi/lo pairs
//   layout 0: [300 -> 304 | 600 -> 608 | 300 -> 304] (LSTMCells, projection: the first two)   1: [300 -> 304]
template <bool P16, typename RowFn>
__device__ __forceinline__ void load_rows(MatSmem<P16>& sm, const float* w, int k_src, int ks, int kp, int n_rows, RowFn src_row) {
  for (int i = threadIdx.x; i < n_rows * ks; i += DEC_THREADS) {
    const int q = i / ks, k = i - q * ks;
    int ksrc = -1;
    if (k < kp) {
      if (k < SEG) ksrc = k < R ? k : -1;
      else if (k < SEG + SEGC) ksrc = k - SEG < E ? R + k - SEG : -1;
      else ksrc = k - SEG - SEGC < R ? R + E + k - SEG - SEGC : -1;
    }
    put_weight(&sm.w[0][q * ks + k], &sm.w[1][q * ks + k], ksrc >= 0 ? __ldg(w + (long long)src_row(q) * k_src + ksrc) : 0.f);
  }
}

template <bool P16>
__device__ __forceinline__ void mat_common_init(const DecParams& p, MatSmem<P16>& sm) {
  // the K padding of the staged inputs must hold finite values (it meets zero weights): zeros now, later at most
  // elements of another (finite) vector that shared the columns
  for (int i = threadIdx.x; i < Cfg<P16>::PASS * XSTRIDE; i += DEC_THREADS) (&sm.xs[0][0])[i] = 0.f;
  if (threadIdx.x == 0) sm.done_count = 0;
}

// roles A and D: one LSTMCell for the units [u0, u0 + nu) of this CTA and all B utterances
//   A: input [pre_t | ctx_t], hidden hatt_t -> hatt_{t+1};  early columns: ctx_t, hatt_t;   awaited: pre_t
//   D: input [hatt_{t+1} | ctx_{t+1}], hidden hdec_t -> hdec_{t+1};  early: hatt_{t+1}, hdec_t;  awaited: ctx_{t+1}
template <bool P16, bool IS_ATT>
__device__ void lstm_role(const DecParams& p, MatSmem<P16>& sm, int idx, int n_ctas, int n_att_ctas) {
  constexpr int PASS = Cfg<P16>::PASS;
  // staging columns: A stages [ctx | hatt] early and pre over hatt; D stages [hatt | hdec] early and ctx over both
  constexpr int XC = 0, X0 = IS_ATT ? SEGC : 0, X1 = IS_ATT ? SEGC : SEG;
  const int tid = threadIdx.x, B = p.B;
  int u0, nu;
  row_range(idx, n_ctas, R, u0, nu);
  const int n_rows = 4 * nu;
  load_rows(sm, IS_ATT ? p.w.w_att : p.w.w_dec, KIN, KS_LSTM, KP_LSTM, n_rows,
            [&](int q) { return (q / nu) * R + u0 + q % nu; });
  for (int i = tid; i < n_rows; i += DEC_THREADS) sm.bias[i] = __ldg((IS_ATT ? p.w.b_att : p.w.b_dec) + (i / nu) * R + u0 + i % nu);
  mat_common_init(p, sm);
  __syncthreads();

  unsigned long long* const xb = p.s.xchg;
  float* const cell = IS_ATT ? p.s.c_att : p.s.c_dec;
  const unsigned int n_lstm = (unsigned int)n_ctas, n_att = (unsigned int)n_att_ctas;
  // K layout of both cells: [first input 304 | context 608 | hidden 304]
  const SegDesc none = {0, 0, 0}, k_in = {0, X0, ST_SEG}, k_ctx = {SEG, XC, ST_SEGC}, k_hid = {SEG + SEGC, X1, ST_SEG};
  PassReg early;
#pragma unroll
  for (int i = 0; i < MAXPASS; ++i) early.v[i] = 0.f;
  Prof prof;
  prof.init(p.prof != nullptr, sm.prof);
  // early columns of step `t` (role A: called right after the critical part of step t - 1, i.e. speculatively
  // before the stop count of that step is known -- both vectors exist by then whether or not there is a step t)
  const unsigned int n_pass = (unsigned int)((B + PASS - 1) / PASS);   // role A announces hatt once per pass
  auto early_columns = [&](int t) -> bool {
    const unsigned int v0 = (unsigned int)t, v1 = v0 + 1;
    bool ok = true;
    if (IS_ATT) {
      // ctx_t and hatt_t
      for (int ps = 0, n0 = 0; n0 < B && ok; ++ps, n0 += PASS) {
        const int nb = min(PASS, B - n0);
        ok = fetch(sm, xchg_vec(xb, B, V_HATT, v0), R, v0, n0, nb, X1, ps == 0 ? xchg_hint(xb, B, V_HATT) : nullptr, n_lstm * n_pass * v0, p.s.done) &&
             fetch(sm, xchg_vec(xb, B, V_CTX, v0), E, v0, n0, nb, XC, ps == 0 ? xchg_hint(xb, B, V_CTX) : nullptr, n_att * v0, p.s.done);
        if (ok) early.set(ps, mat_part(sm, KS_LSTM, n_rows, k_ctx, k_hid, nb));
      }
    } else {
      // hdec_t first (complete since the end of the previous step), then hatt_{t+1}, pass by pass as the role-A
      // CTAs announce it: what is left when the attention has the context ready is as little as possible
      for (int ps = 0, n0 = 0; n0 < B && ok; ++ps, n0 += PASS) {
        const int nb = min(PASS, B - n0);
        ok = fetch(sm, xchg_vec(xb, B, V_HDEC, v0), R, v0, n0, nb, X1, ps == 0 ? xchg_hint(xb, B, V_HDEC) : nullptr, n_lstm * v0, p.s.done);
        if (ok) early.set(ps, mat_part(sm, KS_LSTM, n_rows, k_hid, none, nb));
      }
      for (int ps = 0, n0 = 0; n0 < B && ok; ++ps, n0 += PASS) {
        const int nb = min(PASS, B - n0);
        ok = fetch(sm, xchg_vec(xb, B, V_HATT, v1), R, v1, n0, nb, X0, xchg_hint(xb, B, V_HATT), n_lstm * (n_pass * v0 + ps + 1), p.s.done);
        if (ok) early.set(ps, early.get(ps) + mat_part(sm, KS_LSTM, n_rows, k_in, none, nb));
      }
    }
    prof.mark<2>();
    return ok;
  };
  for (int t = 0; t < p.max_steps; ++t) {
    const unsigned int v0 = (unsigned int)t, v1 = v0 + 1;
    bool ok = true;
    if (!IS_ATT && !early_columns(t)) break;      // (role A, step 0: the zero state, nothing to add)
    // ---- the awaited vector, cell update, publication
    for (int ps = 0, n0 = 0; n0 < B && ok; ++ps, n0 += PASS) {
      const int nb = min(PASS, B - n0);
      float c_old = 0.f;
      if (tid < nu * nb) c_old = cell[(n0 + tid / nu) * R + u0 + tid % nu];     // only this thread ever touches it
      if (IS_ATT)
        ok = fetch(sm, xchg_vec(xb, B, V_PRE, v0), R, v0, n0, nb, X0, nullptr, 0, p.s.done);
      else
        ok = fetch(sm, xchg_vec(xb, B, V_CTX, v1), E, v1, n0, nb, XC, nullptr, 0, p.s.done);
      prof.mark<0>();
      if (!ok) break;
      mat_part(sm, KS_LSTM, n_rows, IS_ATT ? k_in : k_ctx, none, nb, early.get(ps), true);
      if (tid < nu * nb) {                        // cell update: one thread per (unit, utterance)
        const int u = tid % nu, n = tid / nu, j = u0 + u, b = n0 + n;
        float gv[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) gv[g] = sm.bias[g * nu + u] + sm.sums[g * nu + u][n];
        const float cn = sigmoidf_fast(gv[1]) * c_old + sigmoidf_fast(gv[0]) * tanhf_fast(gv[2]);
        cell[b * R + j] = cn;
        st_tagged(xchg_vec(xb, B, IS_ATT ? V_HATT : V_HDEC, v1) + b * R + j, sigmoidf_fast(gv[3]) * tanhf_fast(cn), v1);
      }
      if (IS_ATT) hint_arrive(xchg_hint(xb, B, V_HATT));      // per pass: the decoder LSTM's CTAs start on it
      prof.mark<1>();
    }
    if (!ok) break;
    if (!IS_ATT) hint_arrive(xchg_hint(xb, B, V_HDEC));
    if (IS_ATT && t + 1 < p.max_steps && !early_columns(t + 1)) break;
    if (!step_continues(p, &sm.n_done, &sm.ok, v1, prof)) break;
  }
  prof.flush(p.prof);
}

// role P: rows [pp0, pp0 + npp) of [linear_projection | gate_layer | prenet layer 0 o projection] on hc = [h_dec | context]
template <bool P16>
__device__ void proj_role(const DecParams& p, MatSmem<P16>& sm, int idx) {
  constexpr int PASS = Cfg<P16>::PASS, XC = 0, X0 = 0;     // the context early, h_dec over it
  const int tid = threadIdx.x, B = p.B;
  int pp0, npp;
  row_range(idx, N_PP_CTAS, NPP, pp0, npp);
  load_rows(sm, p.w.w_pp, KHC, KS_PP, KP_PP, npp, [&](int q) { return pp0 + q; });
  for (int i = tid; i < npp; i += DEC_THREADS) sm.bias[i] = __ldg(p.w.b_pp + pp0 + i);
  mat_common_init(p, sm);
  __syncthreads();

  unsigned long long* const xb = p.s.xchg;
  unsigned long long* const done_word = reinterpret_cast<unsigned long long*>(p.s.done + 4);
  const bool owns_gate = pp0 <= M && M < pp0 + npp;
  const SegDesc none = {0, 0, 0}, k_hdec = {0, X0, ST_SEG}, k_ctx = {SEG, XC, ST_SEGC};
  PassReg early;
#pragma unroll
  for (int i = 0; i < MAXPASS; ++i) early.v[i] = 0.f;
  Prof prof;
  prof.init(p.prof != nullptr, sm.prof);
  for (int t = 0; t < p.max_steps; ++t) {
    const unsigned int v1 = (unsigned int)t + 1;
    const bool more = t + 1 < p.max_steps;
    bool ok = true;
    // ---- early: the context columns
    for (int ps = 0, n0 = 0; n0 < B && ok; ++ps, n0 += PASS) {
      const int nb = min(PASS, B - n0);
      // (13 CTAs: they poll the words themselves, pass by pass, instead of waiting for all B attention CTAs)
      ok = fetch(sm, xchg_vec(xb, B, V_CTX, v1), E, v1, n0, nb, XC, nullptr, 0, p.s.done);
      if (ok) early.set(ps, mat_part(sm, KS_PP, npp, k_ctx, none, nb));
    }
    prof.mark<2>();
    if (!ok) break;
    for (int ps = 0, n0 = 0; n0 < B && ok; ++ps, n0 += PASS) {
      const int nb = min(PASS, B - n0);
      unsigned char drop0 = 0;                    // the dropout mask byte comes from DRAM: ask for it early
      if (tid < npp * nb) {
        const int row = pp0 + tid % npp;
        if (row > M && more) drop0 = p.drop[(((long long)(t + 1) * 2 + 0) * B + n0 + tid / npp) * R + row - M - 1];
      }
      ok = fetch(sm, xchg_vec(xb, B, V_HDEC, v1), R, v1, n0, nb, X0, nullptr, 0, p.s.done);
      prof.mark<0>();
      if (!ok) break;
      mat_part(sm, KS_PP, npp, k_hdec, none, nb, early.get(ps), true);
      if (tid < npp * nb) {
        const int r = tid % npp, n = tid / npp, row = pp0 + r, b = n0 + n;
        const float v = sm.bias[r] + sm.sums[r][n];
        if (row < M) {
          p.mel[((long long)b * p.max_steps + t) * M + row] = v;
        } else if (row == M) {
          p.gate[(long long)b * p.max_steps + t] = v;
          if (p.s.out_len[b] == 0) {              // stop test (model.py:524), per utterance
            if (sigmoidf_exact(v) > p.gate_threshold) {
              p.s.out_len[b] = t + 1;
              atomicAdd(p.s.done, 1);
              atomicAdd(&sm.done_count, 1);
            } else if (!more) {
              p.s.out_len[b] = p.max_steps;       // model.py:526-528 "Reached max decoder steps"
              atomicAdd(p.s.done + 1, 1);
              atomicAdd(&sm.done_count, 1);
            }
          }
        } else if (more) {
          // prenet layer 0 of the NEXT step; dropout p = 0.5 is always on -> mask * 2
          st_tagged(xchg_vec(xb, B, V_P1, v1) + b * R + row - M - 1, fmaxf(v, 0.f) * (2.0f * (float)drop0), v1);
        }
      }
      prof.mark<1>();
    }
    if (!ok) break;
    hint_arrive(xchg_hint(xb, B, V_P1));
    if (owns_gate) {                              // how many utterances have stopped, as a tagged word of this step
      if (tid == 0)                               // (hint_arrive's barrier ordered the epilogue's counting)
        asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(done_word),
                     "l"(((unsigned long long)v1 << 32) | (unsigned int)sm.done_count)
                     : "memory");
    }
    if (!step_continues(p, &sm.n_done, &sm.ok, v1, prof)) break;
  }
  prof.flush(p.prof);
}

// role Q: rows [r0, r0 + nr) of prenet layer 1
template <bool P16>
__device__ void prenet_role(const DecParams& p, MatSmem<P16>& sm, int idx) {
  constexpr int PASS = Cfg<P16>::PASS, X0 = 0;
  const int tid = threadIdx.x, B = p.B;
  int r0, nr;
  row_range(idx, N_P2_CTAS, R, r0, nr);
  for (int i = tid; i < nr * KS_P2; i += DEC_THREADS) {
    const int q = i / KS_P2, k = i - q * KS_P2;
    put_weight(&sm.w[0][q * KS_P2 + k], &sm.w[1][q * KS_P2 + k], k < R ? __ldg(p.w.w_pre2 + (long long)(r0 + q) * R + k) : 0.f);
  }
  mat_common_init(p, sm);
  __syncthreads();

  unsigned long long* const xb = p.s.xchg;
  const SegDesc none = {0, 0, 0}, k_all = {0, X0, ST_SEG};
  Prof prof;
  prof.init(p.prof != nullptr, sm.prof);
  for (int t = 0; t < p.max_steps; ++t) {
    const unsigned int v1 = (unsigned int)t + 1;
    bool ok = true;
    if (t + 1 < p.max_steps) {
      for (int ps = 0, n0 = 0; n0 < B && ok; ++ps, n0 += PASS) {
        const int nb = min(PASS, B - n0);
        unsigned char drop1 = 0;
        if (tid < nr * nb) drop1 = p.drop[(((long long)(t + 1) * 2 + 1) * B + n0 + tid / nr) * R + r0 + tid % nr];
        ok = fetch(sm, xchg_vec(xb, B, V_P1, v1), R, v1, n0, nb, X0, nullptr, 0, p.s.done);
        prof.mark<0>();
        if (!ok) break;
        mat_part(sm, KS_P2, nr, k_all, none, nb, 0.f, true);
        if (tid < nr * nb) {
          const int r = tid % nr, n = tid / nr;
          st_tagged(xchg_vec(xb, B, V_PRE, v1) + (n0 + n) * R + r0 + r, fmaxf(sm.sums[r][n], 0.f) * (2.0f * (float)drop1), v1);
        }
        prof.mark<1>();
      }
      if (!ok) break;
      hint_arrive(xchg_hint(xb, B, V_PRE));
    }
    if (!step_continues(p, &sm.n_done, &sm.ok, v1, prof)) break;
  }
  prof.flush(p.prof);
}

// --------------------------------------------------------------------------------------- attention CTAs
// Location-sensitive attention of utterance b (model.py:100-121, 56-60, 78-98).
//   prepare(t): everything of step t that only needs the attention weights of step t-1
//   critical(t): the part that needs h_att(t)
__device__ void attention_role(const DecParams& p, AttSmem& sm, int b) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
    long long tp = prof.now();
    int end;
    window_bounds(t, p.window, len, start, end);
    nw = end - start + 1;
    // previous / cumulative weights around the window (zero outside the sequence: conv padding)
    constexpr int NCAT = MAXW + KF - 1 + 2;
    const int c0 = start - (KF - 1) / 2, ncat = nw + KF - 1;
    if (tid < 2 * NCAT) {
      const int c = tid >= NCAT, q = tid - c * NCAT, pos = c0 + q;
      float v = 0.f;
      if (q < ncat && pos >= 0 && pos < p.T_in) v = c == 0 ? wprev[pos] : wcum[pos];
      sm.cat[c][q] = v;                // zeros beyond the window: positions >= nw below are computed but never used
    }
    __syncthreads();
    prof.sub<16>(tp);
    // location_conv (model.py:57): loc[q][f] = sum_{c,k} w[c][k][f] * cat[c][q + k].  Lane = filter, warp = three
    // consecutive positions: one weight and one new input per tap feed three FMAs (sliding window).
    {
      const int f = lane, q0 = 3 * warp;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float x0 = sm.cat[c][q0], x1 = sm.cat[c][q0 + 1];
#pragma unroll
        for (int k = 0; k < KF; ++k) {
          const float x2 = sm.cat[c][q0 + k + 2];
          const float w = __ldg(p.w.w_loc + (c * KF + k) * NF + f);
          a0 = fmaf(w, x0, a0);
          a1 = fmaf(w, x1, a1);
          a2 = fmaf(w, x2, a2);
          x0 = x1;
          x1 = x2;
        }
      }
      const float v[3] = {a0, a1, a2};
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const __half h = __float2half_rn(v[i]);
        sm.x.loc[0][q0 + i][f] = h;
        sm.x.loc[1][q0 + i][f] = __float2half_rn(v[i] - __half2float(h));
      }
    }
    __syncthreads();
    prof.sub<17>(tp);
    // S[q][a] = exp(2 (location_dense(loc[q])[a] + processed_memory[q][a])) (model.py:94-96) on the tensor cores:
    // [48 positions x 32 filters] x [32 x 150 channels], a warp per 8-channel n-tile (19 of them), all three
    // m-tiles; the weights come straight from global memory (L1-resident: 19 KB), split into half hi/lo pairs
    // like the features (three products: fp32-grade).
    for (int nt = warp; nt < (A + 7) / 8; nt += DEC_WARPS) {
      const int g = lane >> 2, t2 = 2 * (lane & 3), a_col = 8 * nt + t2;       // this lane's output columns a_col, + 1
      // processed_memory of this lane's outputs (rows g, g + 8 of each m-tile): ask for it first
      float2 pm[3][2];
#pragma unroll
      for (int mt = 0; mt < 3; ++mt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int q = 16 * mt + g + 8 * hh;
          pm[mt][hh] = (q < nw && a_col < A) ? __ldg(reinterpret_cast<const float2*>(pmem + (long long)(start + q) * A + a_col))
                                            : make_float2(0.f, 0.f);
        }
      float acc[3][4];
#pragma unroll
      for (int i = 0; i < 12; ++i) (&acc[0][0])[i] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        // B fragment: filters (k) 16 ks + t2, + 1 and + 8, + 9 of channel 8 nt + g
        const int an = 8 * nt + g;
        const float* wp = p.w.w_ld_t + (16 * ks + t2) * A + an;
        float2 w0 = make_float2(0.f, 0.f), w1 = w0;
        if (an < A) {
          w0 = make_float2(__ldg(wp), __ldg(wp + A));
          w1 = make_float2(__ldg(wp + 8 * A), __ldg(wp + 9 * A));
        }
        uint32_t bh[2], bl[2];
        split_half2(w0, bh[0], bl[0]);
        split_half2(w1, bh[1], bl[1]);
        const int arow = (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) {
          uint32_t ah[4], al[4];
          ldmatrix_x4_fresh(ah, (uint32_t)__cvta_generic_to_shared(&sm.x.loc[0][16 * mt + arow][16 * ks + (lane >> 4) * 8]));
          ldmatrix_x4_fresh(al, (uint32_t)__cvta_generic_to_shared(&sm.x.loc[1][16 * mt + arow][16 * ks + (lane >> 4) * 8]));
          mma_f16(acc[mt], ah, bh);
          mma_f16(acc[mt], al, bh);
          mma_f16(acc[mt], ah, bl);
        }
      }
#pragma unroll
      for (int mt = 0; mt < 3; ++mt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int q = 16 * mt + g + 8 * hh;
          if (q < nw && a_col < A) {
            sm.S[q][a_col] = exp2x(acc[mt][2 * hh] + pm[mt][hh].x);
            sm.S[q][a_col + 1] = exp2x(acc[mt][2 * hh + 1] + pm[mt][hh].y);
          }
        }
    }
    prof.sub<18>(tp);
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
    prof.sub<19>(tp);
  };

  auto critical = [&](int t) -> bool {
    long long tp = prof.now();
    // query projection (model.py:92): pq = W_q h_att; warp per row, 75 float4 per row over the lanes
    {
      // h_att_{t+1}: spin on the (value, version) words the matrix CTAs publish -- one-way latency
      bool dead = false;
      if (tid < R) {
        const unsigned long long* src = xchg_vec(p.s.xchg, p.B, V_HATT, (unsigned int)t + 1) + b * R + tid;
        unsigned long long v;
        int spins = 0;
        do {
          v = ld_tagged(src);
        } while ((unsigned int)(v >> 32) != (unsigned int)(t + 1) && !(dead = spin_dead(spins, p.s.done)));
        sm.h[tid] = __uint_as_float((unsigned int)v);
      }
      if (__syncthreads_or(dead)) return false;
      prof.sub<14>(tp);
      // W_q h: lane = row (32 consecutive rows of W_q: their float4 reads are conflict-free, and the matching
      // float4 of h is ONE broadcast address), warp = (row block, third of K); the three partial sums of a row
      // meet in shared memory (the context scratch is free at this point)
      if (warp < 15) {
        const int part = warp % 3, r = 32 * (warp / 3) + lane;
        if (r < A) {
          const float4* w4 = reinterpret_cast<const float4*>(&sm.wq[r][100 * part]);
          const float4* h4 = reinterpret_cast<const float4*>(sm.h + 100 * part);
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 5
          for (int i = 0; i < 25; ++i) {
            const float4 a = w4[i], hv = h4[i];
            s0 = fmaf(a.x, hv.x, s0);
            s1 = fmaf(a.y, hv.y, s1);
            s2 = fmaf(a.z, hv.z, s2);
            s3 = fmaf(a.w, hv.w, s3);
          }
          sm.x.ctxp[part][r] = (s0 + s1) + (s2 + s3);
        }
      }
      __syncthreads();
      if (tid < A) sm.upq[tid] = exp2x(sm.x.ctxp[0][tid] + (sm.x.ctxp[1][tid] + sm.x.ctxp[2][tid]));
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
    {
      unsigned long long* dst = xchg_vec(p.s.xchg, p.B, V_CTX, (unsigned int)t + 1) + b * E;
      for (int c = tid; c < E; c += DEC_THREADS)
        st_tagged(dst + c, sm.x.ctxp[0][c] + sm.x.ctxp[1][c] + sm.x.ctxp[2][c], (unsigned int)t + 1);
      hint_arrive(xchg_hint(p.s.xchg, p.B, V_CTX));
    }
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
    __syncthreads();                 // prepare(t + 1) reads the weights just written (other threads)
    return true;
  };

  const unsigned long long* done_word = reinterpret_cast<const unsigned long long*>(p.s.done + 4);
  prepare(0);
  for (int t = 0; t < p.max_steps; ++t) {
    prof.mark<0>();
    if (!critical(t)) break;
    prof.mark<2>();
    if (t + 1 < p.max_steps) prepare(t + 1);
    prof.mark<4>();
    // the stop counter of this step (a tagged word from the matrix CTA that owns the gate row)
    if (tid == 0) {
      int spins = 0;
      bool dead = false;
      unsigned long long w;
      do {
        w = ld_tagged(done_word);
      } while ((unsigned int)(w >> 32) != (unsigned int)(t + 1) && !(dead = spin_dead(spins, p.s.done)));
      sm.n_done = (int)(unsigned int)w;
      sm.ok = !dead;
    }
    __syncthreads();
    prof.mark<9>();
    const bool stop = sm.n_done >= p.B || !sm.ok;
    if (stop) {
      if (blockIdx.x == 0 && tid == 0) p.s.done[2] = t + 1;
      break;
    }
    __syncthreads();                 // sm.n_done / sm.ok are rewritten next step
  }
  prof.flush(p.prof);
}

template <bool P16>
__global__ void __launch_bounds__(DEC_THREADS, 1) taco_decoder_kernel(const DecParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int mi = (int)blockIdx.x - p.B, GM = (int)gridDim.x - p.B;
  const int n_lstm = (GM - N_PP_CTAS - N_P2_CTAS) / 2;          // CTAs per LSTMCell
  if (mi < 0) {
    attention_role(p, *reinterpret_cast<AttSmem*>(smem_raw), blockIdx.x);
  } else {
    MatSmem<P16>& sm = *reinterpret_cast<MatSmem<P16>*>(smem_raw);
    if (mi < n_lstm) lstm_role<P16, true>(p, sm, mi, n_lstm, p.B);
    else if (mi < 2 * n_lstm) lstm_role<P16, false>(p, sm, mi - n_lstm, n_lstm, p.B);
    else if (mi < 2 * n_lstm + N_PP_CTAS) proj_role<P16>(p, sm, mi - 2 * n_lstm);
    else if (mi < 2 * n_lstm + N_PP_CTAS + N_P2_CTAS) prenet_role<P16>(p, sm, mi - 2 * n_lstm - N_PP_CTAS);
    // (an odd CTA out, if any, has nothing to do)
  }
}

// Barrier over `n` CTAs as round 1's decoder used it between phases (kept as a diagnostic: its cost is what the
// dataflow design avoids): monotonically increasing arrival counter, release on the add, acquire spin.
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
  FAC_REQUIRE(s->c_att && s->c_dec && s->xchg && s->w_prev && s->w_cum && s->done && s->out_len,
              "taco_decoder: NULL state buffer");
  FAC_REQUIRE(B <= 8 * MAXPASS, "taco_decoder: at most %d utterances per launch", 8 * MAXPASS);
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
  const bool p16 = B > 8 && B <= MAX_B_P16;
  const size_t mat_smem = p16 ? sizeof(MatSmem<true>) : sizeof(MatSmem<false>);
  const size_t smem = mat_smem > sizeof(AttSmem) ? mat_smem : sizeof(AttSmem);
  const void* kernel = p16 ? (const void*)taco_decoder_kernel<true> : (const void*)taco_decoder_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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
  // cooperative launch = co-residency guarantee for the CTAs that poll each other's words
  e = cudaLaunchCooperativeKernel(kernel, dim3(sms), dim3(DEC_THREADS), args, smem, st);
  count_launch();
  if (e != cudaSuccess) {
    set_error("taco_decoder: cooperative launch failed: %s", cudaGetErrorString(e));
    return 2;
  }
  return check_launch("taco_decoder_kernel");
}

}  // namespace fac
