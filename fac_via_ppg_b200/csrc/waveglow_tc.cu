// WaveGlow WN layer on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), reference
// src/waveglow/glow.py:158-175.
//
// Each layer is two implicit GEMMs over 128-column time tiles:
//   G1: pre[128 x 2C] = [x(t-d) | x(t) | x(t+d) | spect(t)] (K = 3C + n_cond) * W1^T
//       epilogue: gate tanh(.)*sigmoid(.) (glow.py:33-40) -> acts (bf16 hi/lo), and the skip path
//       collapsed algebraically: `end` is linear in the skip sum (glow.py:171-175), so
//       end(sum_i skip_i) = sum_i (W_end W_skip_i) acts_i + const; each layer adds its 8-channel
//       contribution  out8 += Wc_i acts  (fp32 FFMA on the exact gate outputs).  The (B,T,C) skip
//       tensor and half of the res_skip GEMM disappear.
//   G2: x_new[128 x C] = [acts | x] (K = 2C) * [W_res | I]^T  -- the residual add (glow.py:166) is
//       done BY the tensor core through an identity block, so the epilogue has no global loads: it only
//       re-splits the fp32 accumulator into the bf16 hi/lo operand copies of the next layer.
// Operands are bf16 in channels-last HBM layout and reach shared memory through TMA (dilated taps
// are shifted box coordinates; rows outside [0, T) are zero-filled by the TMA unit = Conv1d padding).
// Accumulation is fp32 in tensor memory.  Precision modes:
//   nsplit = 1  plain bf16 operands                                  (1 UMMA per product)
//   nsplit = 2  split-bf16: v = hi + lo, a*w ~= ah*wh + al*wh + ah*wl (3 UMMAs per product),
//               ~2^-16 relative operand error, fp32 accumulate: fp32-grade results.
// Warp roles (320 threads, persistent over tiles, 1 CTA / SM): warp 0 = TMA producer, warp 1 = UMMA
// issuer (one elected lane), warps 2-9 = epilogue (two per TMEM lane quarter, half of the columns each).
// The same kernel serves the acoustic model's Conv1d / Linear layers (TC_LINEAR epilogue, IEEE-half operand
// pairs): there the (tile, column block) units are dealt round-robin to the CTA pairs and every unit walks its
// contraction in chains of <= k_chunk elements that alternate the two TMEM accumulators (see conv_gemm_tc).
#include <stdlib.h>
#include "fac_common.cuh"
#include "tc_common.cuh"
#include "tc_host.cuh"

namespace fac {
namespace {

using namespace tc;

// K per pipeline stage (BK): 64 (128-byte rows, SWIZZLE_128B) by default, 32 (64-byte rows, SWIZZLE_64B) optional:
// the TMA issue rate is per row request, so rows should be as wide as the stage budget allows.
constexpr int TC_BK_MAX = 64;
constexpr int UMMA_K = 16;
constexpr int TC_NMAX = 512;     // accumulator columns per tile (all of TMEM)
constexpr int TC_EPI_WARPS = 8;   // two warps per TMEM lane quarter, each owning half of the columns
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_NOUT = 8;       // channels of the collapsed skip path (= max 2*n_half)
constexpr int TC_CMAX = 256;     // max WN channels (Wc staging)
constexpr int TC_BAR_BYTES = 256;
constexpr int TC_WC_BYTES = TC_NOUT * TC_CMAX * 4;       // 8 KB
constexpr int TC_X8_BYTES = TC_BM * TC_NOUT * 4;         // 4 KB: out8 partials handed between paired epilogue warps
constexpr int TC_MAX_STAGES = 12;
constexpr int TC_RING_BYTES = 192 * 1024;   // operand ring; the stage count follows from the stage size

// CG = 1: one CTA per 128-row tile.  CG = 2: a CTA pair (thread-block cluster of 2, tcgen05 cta_group::2)
// works on two adjacent tiles as ONE 256-row UMMA; each CTA stages only half of the weight rows, so
// the per-SM operand traffic and shared-memory reads drop by a third and the ring gets 6 stages.
template <int CG, int BK>
struct TcCfg {
  static constexpr int ROWB = BK * 2;                                // bytes of one operand row
  static constexpr int A_BYTES = TC_BM * ROWB;                       // 8 / 16 KB
  static constexpr int W_BYTES = (TC_NHALF / CG) * ROWB;             // this CTA's share of a 256-row weight block
  // a stage holds nsplit x (A tile + weight share): split-bf16 48 KB / 32 KB (4 / 6 stages), bf16 half of that
  // (8 / 12 stages)
  static constexpr int SMEM = TC_RING_BYTES + TC_BAR_BYTES + TC_WC_BYTES + TC_X8_BYTES + 1024 /*alignment*/;
};

enum { TC_GATE = 0, TC_RESIDUAL = 1, TC_LINEAR = 2 };

struct TcSrc {
  int channels, taps, dilation, center;
};

struct TcParams {
  int n_src;
  TcSrc src[2];
  int n_total, k_steps, wlo_k_steps;   // K steps [0, wlo_k_steps) also use the W_lo term (split mode)
  int T, B, tiles_per_batch, n_tiles, nsplit, mode, C;
  const float* bias;
  __nv_bfloat16* out_hi;   // gate: acts; residual: x   (B, T, C)
  __nv_bfloat16* out_lo;
  const float* wc;         // gate: [TC_NOUT][C] collapsed skip weights (fp32)
  float* out8;             // gate: (B, T, TC_NOUT) running end() pre-activation
  int accumulate_out8;
  int out_row_mul, out_row_off;   // residual epilogue: output row = column * mul + off (phase-strided upsampler)
  // grouped launch (the upsampler's phases): n_phases GEMMs over the SAME activations, phase ph with the weight rows
  // [ph * w_phase_rows, ...) of the weight map and output row offset out_row_off + ph; the tile index space is
  // n_phases x tiles_pp (tiles_pp = the tiles of one phase rounded up to whole CTA groups).  0 = one GEMM.
  int n_phases, tiles_pp, w_phase_rows;
  // TC_LINEAR (generic Conv1d / Linear): v = act(acc + bias) [* mask] [+ residual] on the first n_valid columns,
  // written as fp32 (out_f32, row stride out_ld) and/or as the bf16 hi/lo operand copies of the next layer
  // (out_hi/out_lo, row stride C; columns >= n_valid are exact zeros because their weights and bias are)
  int act, n_valid;
  int fp16;                       // operands (and the hi/lo outputs of TC_LINEAR) are IEEE half instead of bf16
  int flat_units;                 // 1: (tile, column block) units are dealt round-robin to the CTA groups (TC_LINEAR);
                                  // 0: a CTA group owns whole tiles (the gate epilogue's out8 update needs that)
  int chunk_steps;                // K steps per accumulation chain (K-chunked accumulation, see conv_gemm_tc); 0 = all
  float* scratch;                 // fp32 partial sums of the earlier K chunks of a unit, (rows, n_valid)
  long long scratch_ld;
  const float* mask;
  const float* residual;
  float* out_f32;
  long long mask_ld, res_ld, out_ld;
  const int* row_lengths;  // TC_LINEAR, optional [B]: rows t >= row_lengths[b] are written as zeros (ragged batches)
  long long* prof;         // optional [grid][8] cycle counters
};

template <int CG, int BK>
__global__ void __launch_bounds__(TC_THREADS, 1)
wn_gemm_tc_kernel(const __grid_constant__ CUtensorMap a0_hi, const __grid_constant__ CUtensorMap a0_lo,
                  const __grid_constant__ CUtensorMap a1_hi, const __grid_constant__ CUtensorMap a1_lo,
                  const __grid_constant__ CUtensorMap w_hi, const __grid_constant__ CUtensorMap w_lo,
                  const TcParams p) {
  constexpr int W_BYTES = TcCfg<CG, BK>::W_BYTES, A_BYTES = TcCfg<CG, BK>::A_BYTES, TC_ROWB = TcCfg<CG, BK>::ROWB, TC_BK = BK;
  const int STAGE_BYTES = p.nsplit * (A_BYTES + W_BYTES);
  const int TC_STAGES = min(TC_MAX_STAGES, TC_RING_BYTES / STAGE_BYTES);
  const int W_OFF = p.nsplit * A_BYTES;      // stage layout: A_hi [A_lo] W_hi [W_lo]
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + TC_RING_BYTES);
  uint64_t* empty = full + TC_MAX_STAGES;
  uint64_t* tmem_full = empty + TC_MAX_STAGES;   // [2] one per accumulator region
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* wc_s = reinterpret_cast<float*>(smem + TC_RING_BYTES + TC_BAR_BYTES);
  float* x8_s = wc_s + TC_NOUT * TC_CMAX;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = CG == 1 ? 0 : (int)cluster_ctarank();       // position inside the CTA pair
  const int tile_first = (blockIdx.x / CG) * CG;               // pairs take adjacent tiles in lock step
  // A work unit is (time tile, block of <= 256 output columns); its accumulator is one of the two
  // 256-column TMEM regions, so the epilogue of unit u overlaps the UMMAs of unit u+1.
  const int n_blocks = (p.n_total + TC_NHALF - 1) / TC_NHALF;
  const int n_cols = p.n_total < TC_NHALF ? p.n_total : TC_NHALF;
  // unit schedule, identical in the three roles: the it-th unit of this CTA group is (first tile tb, block nb)
  const int n_groups = (int)gridDim.x / CG, group_id = (int)blockIdx.x / CG;
  // grouped launch: tile index tb -> (phase, tile of the phase); a single GEMM is one phase of all tiles
  const int n_phases = p.n_phases > 0 ? p.n_phases : 1;
  const int tiles_pp = p.n_phases > 0 ? p.tiles_pp : ((p.n_tiles + CG - 1) / CG) * CG;
  const int total_units = ((p.n_tiles + CG - 1) / CG) * n_blocks;
  // K chunks (flat schedule only): a unit's contraction is walked in chains of chunk_steps K steps, each with its
  // own accumulator; the chains of a unit are consecutive entries of the schedule of the same CTA group
  const int chunk_steps = (p.flat_units && p.chunk_steps > 0) ? p.chunk_steps : p.k_steps;
  const int n_chunks = (p.k_steps + chunk_steps - 1) / chunk_steps;
  auto unit_at = [&](int it, int& tb, int& nb, int& ck) -> bool {
    if (p.flat_units) {
      const int uidx = group_id + (it / n_chunks) * n_groups;
      if (uidx >= total_units) return false;
      tb = (uidx / n_blocks) * CG;
      nb = uidx % n_blocks;
      ck = it % n_chunks;
      return true;
    }
    tb = tile_first + (it / n_blocks) * (int)gridDim.x;
    nb = it % n_blocks;
    ck = 0;
    return tb < n_phases * tiles_pp;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int r = 0; r < 2; ++r) {
      mbar_init(&tmem_full[r], 1);
      mbar_init(&tmem_empty[r], TC_EPI_WARPS * CG);           // CG = 2: both CTAs' epilogues release the leader
    }
    fence_barrier_init();
    tma_prefetch_desc(&a0_hi);
    tma_prefetch_desc(&w_hi);
  }
  if (p.mode == TC_GATE)
    for (int i = threadIdx.x; i < TC_NOUT * p.C; i += TC_THREADS) wc_s[i] = __ldg(p.wc + i);
  if (warp == 1) {
    if (CG == 1) tmem_alloc(tmem_slot, TC_NMAX);
    else tmem_alloc_cg2(tmem_slot, TC_NMAX);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();     // the peer's barriers exist before anything is signalled remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      const int w_rows = n_cols / CG;                                   // weight rows staged by this CTA
      const uint32_t w_bytes = (uint32_t)(w_rows * TC_ROWB);
      const int steps0 = p.src[0].taps * (p.src[0].channels / TC_BK);
      int stage = 0;
      uint32_t phase = 0;
      long long prod_wait = 0;
      for (int it = 0;; ++it) {
        int tb, nb, ck;
        if (!unit_at(it, tb, nb, ck)) break;
        const int ph = tb / tiles_pp;
        const int tile = tb - ph * tiles_pp + rank;    // may be one past the end for the pair's second CTA:
        const int b = tile / p.tiles_per_batch;        // its loads are then fully out of bounds (zero fill)
        const int t0 = (tile % p.tiles_per_batch) * TC_BM;
        {
          const int ks_end = min(p.k_steps, (ck + 1) * chunk_steps);
          for (int ks = ck * chunk_steps; ks < ks_end; ++ks) {
            // decode the K step into (source, tap, channel block)
            int s = 0, rem = ks;
            if (rem >= steps0) { s = 1; rem -= steps0; }
            const int cps = p.src[s].channels / TC_BK;
            const int tap = rem / cps, c0 = (rem - tap * cps) * TC_BK;
            const int row0 = t0 + tap * p.src[s].dilation - p.src[s].center;
            const CUtensorMap* mh = s == 0 ? &a0_hi : &a1_hi;
            const CUtensorMap* ml = s == 0 ? &a0_lo : &a1_lo;
            const bool use_wlo = p.nsplit == 2 && ks < p.wlo_k_steps;
            const long long w0 = clock64();
            mbar_wait(&empty[stage], phase ^ 1);
            prod_wait += clock64() - w0;
            uint8_t* st = smem + stage * STAGE_BYTES;
            const uint32_t bytes = (uint32_t)(p.nsplit * A_BYTES) + w_bytes * (use_wlo ? 2u : 1u);
            const int w_row = ph * p.w_phase_rows + nb * TC_NHALF + rank * w_rows;
            if (CG == 1) {
              mbar_arrive_expect_tx(&full[stage], bytes);
              tma_load_3d(st, mh, &full[stage], c0, row0, b);
              if (p.nsplit == 2) tma_load_3d(st + A_BYTES, ml, &full[stage], c0, row0, b);
              tma_load_2d(st + W_OFF, &w_hi, &full[stage], ks * TC_BK, w_row);
              if (use_wlo) tma_load_2d(st + W_OFF + W_BYTES, &w_lo, &full[stage], ks * TC_BK, w_row);
            } else {
              // both CTAs' loads complete on the LEADER's barrier, which expects the pair's bytes
              if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * bytes);
              const uint32_t lead_full = mapa_u32(smem_u32(&full[stage]), 0);
              tma_load_3d_cg2(st, mh, lead_full, c0, row0, b);
              if (p.nsplit == 2) tma_load_3d_cg2(st + A_BYTES, ml, lead_full, c0, row0, b);
              tma_load_2d_cg2(st + W_OFF, &w_hi, lead_full, ks * TC_BK, w_row);
              if (use_wlo) tma_load_2d_cg2(st + W_OFF + W_BYTES, &w_lo, lead_full, ks * TC_BK, w_row);
            }
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
      if (p.prof) p.prof[blockIdx.x * 8 + 0] = prod_wait;
    }
  } else if (warp == 1) {
    // ===================================================== UMMA issuer
    if (lane == 0 && rank == 0) {                     // CG = 2: the leader issues for the pair
      const uint32_t idesc = p.fp16 ? make_idesc_f16(TC_BM * CG, n_cols) : make_idesc_bf16(TC_BM * CG, n_cols);
      auto mma = [](uint32_t d, uint64_t a, uint64_t w, uint32_t id, uint32_t acc) {
        if (CG == 1) umma_bf16(d, a, w, id, acc);
        else umma_bf16_cg2(d, a, w, id, acc);
      };
      auto commit = [](uint64_t* bar) {
        if (CG == 1) umma_commit(bar);
        else umma_commit_cg2(bar, (uint16_t)0x3);
      };
      int stage = 0;
      uint32_t phase = 0;
      uint32_t u = 0;   // unit counter
      long long wait_tmem = 0, wait_full = 0;
      const long long k_start = clock64();
      for (int it = 0;; ++it) {
        int tb, nb, ck;
        if (!unit_at(it, tb, nb, ck)) break;
        const int ks_begin = ck * chunk_steps, ks_end = min(p.k_steps, ks_begin + chunk_steps);
        for (int once = 0; once < 1; ++once, ++u) {
          const uint32_t r = u & 1;
          long long w0 = clock64();
          mbar_wait(&tmem_empty[r], ((u >> 1) & 1) ^ 1);   // epilogue has drained this region
          wait_tmem += clock64() - w0;
          tc_fence_after();
          const uint32_t d = tmem_base + r * TC_NHALF;
          for (int ks = ks_begin; ks < ks_end; ++ks) {
            w0 = clock64();
            mbar_wait(&full[stage], phase);
            wait_full += clock64() - w0;
            tc_fence_after();
            const uint32_t st = smem_u32(smem + stage * STAGE_BYTES);
#pragma unroll
            for (int kk = 0; kk < TC_BK / UMMA_K; ++kk) {
              const uint32_t koff = kk * UMMA_K * 2;   // bytes along K inside the swizzled row
              const uint64_t a_h = make_smem_desc(st + koff, TC_ROWB);
              const uint64_t w_h = make_smem_desc(st + W_OFF + koff, TC_ROWB);
              mma(d, a_h, w_h, idesc, (ks > ks_begin || kk > 0) ? 1u : 0u);
              if (p.nsplit == 2) {
                const uint64_t a_l = make_smem_desc(st + A_BYTES + koff, TC_ROWB);
                mma(d, a_l, w_h, idesc, 1u);
                if (ks < p.wlo_k_steps) {
                  const uint64_t w_l = make_smem_desc(st + W_OFF + W_BYTES + koff, TC_ROWB);
                  mma(d, a_h, w_l, idesc, 1u);
                }
              }
            }
            commit(&empty[stage]);        // frees the smem slot (in both CTAs) when these UMMAs have read it
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
          }
          commit(&tmem_full[r]);          // accumulator complete -> epilogue (of both CTAs)
        }
      }
      if (p.prof) {
        p.prof[blockIdx.x * 8 + 1] = wait_tmem;
        p.prof[blockIdx.x * 8 + 2] = wait_full;
        p.prof[blockIdx.x * 8 + 3] = clock64() - k_start;
      }
    }
  } else {
    // ===================================================== epilogue (8 warps: lane quarter x column half)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int c_begin = half * (n_cols / 2), c_end = c_begin + n_cols / 2;
    const int row = q * 32 + lane;
    uint32_t u = 0;
    long long epi_wait = 0, epi_busy = 0;
    for (int it = 0;; ++it) {
      int tb, nb, ck;
      if (!unit_at(it, tb, nb, ck)) break;
      const bool first_chunk = ck == 0, last_chunk = ck == n_chunks - 1;
      const int ph = tb / tiles_pp;
      const int tile = tb - ph * tiles_pp + rank;
      const int b = tile / p.tiles_per_batch;
      const int t = (tile % p.tiles_per_batch) * TC_BM + row;
      const bool valid = tile < p.n_tiles && t < p.T;
      const bool dead_row = valid && p.row_lengths != nullptr && t >= __ldg(p.row_lengths + b);
      const long long col = (long long)b * p.T + t;
      for (int once = 0; once < 1; ++once, ++u) {
        const uint32_t r = u & 1;
        const long long e0 = clock64();
        mbar_wait(&tmem_full[r], (u >> 1) & 1);
        const long long e1 = clock64();
        epi_wait += e1 - e0;
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + r * TC_NHALF;
        float acc8[TC_NOUT];
#pragma unroll
        for (int o = 0; o < TC_NOUT; ++o) acc8[o] = 0.f;
        for (int c = c_begin; c < c_end; c += 32) {
          uint32_t rr[32];
          tmem_ld32(trow + c, rr);
          tmem_ld_wait();
          const int n0 = nb * TC_NHALF + c;     // global output column of this chunk
          if (!valid || n0 >= p.n_total) continue;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bv = (p.bias != nullptr && last_chunk) ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j))
                                                                : make_float4(0.f, 0.f, 0.f, 0.f);
            v[j + 0] = __uint_as_float(rr[j + 0]) + bv.x;
            v[j + 1] = __uint_as_float(rr[j + 1]) + bv.y;
            v[j + 2] = __uint_as_float(rr[j + 2]) + bv.z;
            v[j + 3] = __uint_as_float(rr[j + 3]) + bv.w;
          }
          if (p.mode == TC_GATE) {
            // columns (2c, 2c+1) hold the tanh / sigmoid pre-activations of channel c (glow.py:33-40)
            float g[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) g[j] = gate_act(v[2 * j], v[2 * j + 1]);
            const int ch0 = n0 >> 1;
            // collapsed skip path: out8 += Wc[:, ch0:ch0+16] g   (fp32, exact gate outputs)
#pragma unroll
            for (int o = 0; o < TC_NOUT; ++o) {
              const float4* wrow = reinterpret_cast<const float4*>(wc_s + o * p.C + ch0);
              float a = acc8[o];
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const float4 w4 = wrow[j4];
                a = fmaf(w4.x, g[4 * j4 + 0], a);
                a = fmaf(w4.y, g[4 * j4 + 1], a);
                a = fmaf(w4.z, g[4 * j4 + 2], a);
                a = fmaf(w4.w, g[4 * j4 + 3], a);
              }
              acc8[o] = a;
            }
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) split2(g[2 * j], g[2 * j + 1], hi[j], lo[j]);
            const long long off = col * p.C + ch0;
            uint4* dh = reinterpret_cast<uint4*>(p.out_hi + off);
            dh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            dh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            if (p.nsplit == 2) {
              uint4* dl = reinterpret_cast<uint4*>(p.out_lo + off);
              dl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              dl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            }
          } else if (p.mode == TC_LINEAR) {
            // generic Conv1d / Linear epilogue (reference src/common/layers.py:40-71 users)
            if (n_chunks > 1) {
              // K-chunked accumulation: the partial sums of a unit's chains meet in fp32 (round to nearest) in
              // `scratch`; this thread wrote what it reads (same row, same columns, previous chain)
              float* srow = p.scratch + col * p.scratch_ld + n0;
              if (!first_chunk) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  if (n0 + j + 3 < p.n_valid) {
                    const float4 m = *reinterpret_cast<const float4*>(srow + j);
                    v[j] += m.x; v[j + 1] += m.y; v[j + 2] += m.z; v[j + 3] += m.w;
                  }
              }
              if (!last_chunk) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  if (n0 + j + 3 < p.n_valid)
                    *reinterpret_cast<float4*>(srow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                continue;
              }
            }
            if (p.act == FAC_ACT_RELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            } else if (p.act == FAC_ACT_TANH) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
            }
            if (p.mask != nullptr) {
              const float* mrow = p.mask + col * p.mask_ld + n0;
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                if (n0 + j + 3 < p.n_valid) {
                  const float4 m = __ldg(reinterpret_cast<const float4*>(mrow + j));
                  v[j] *= m.x; v[j + 1] *= m.y; v[j + 2] *= m.z; v[j + 3] *= m.w;
                }
            }
            if (p.residual != nullptr) {
              const float* rrow = p.residual + col * p.res_ld + n0;
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                if (n0 + j + 3 < p.n_valid) {
                  const float4 m = __ldg(reinterpret_cast<const float4*>(rrow + j));
                  v[j] += m.x; v[j + 1] += m.y; v[j + 2] += m.z; v[j + 3] += m.w;
                }
            }
            if (dead_row) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
            if (p.out_f32 != nullptr) {
              float* orow = p.out_f32 + col * p.out_ld + n0;
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                if (n0 + j + 3 < p.n_valid)
                  *reinterpret_cast<float4*>(orow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            if (p.out_hi != nullptr) {
              uint32_t hi[16], lo[16];
              if (p.fp16) {
#pragma unroll
                for (int j = 0; j < 16; ++j) split2_f16(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
              }
              const long long off = col * p.C + n0;
              uint4* dh = reinterpret_cast<uint4*>(p.out_hi + off);
#pragma unroll
              for (int j = 0; j < 4; ++j) dh[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
              if (p.nsplit == 2) {
                uint4* dl = reinterpret_cast<uint4*>(p.out_lo + off);
#pragma unroll
                for (int j = 0; j < 4; ++j) dl[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
              }
            }
          } else {
            // residual stream x_new = x + res (glow.py:166), already summed by the identity block:
            // refresh the bf16 operand copies the next layer's TMA loads read
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
            const long long off = (col * p.out_row_mul + p.out_row_off + ph) * p.C + n0;
            uint4* dh = reinterpret_cast<uint4*>(p.out_hi + off);
#pragma unroll
            for (int j = 0; j < 4; ++j) dh[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
            if (p.nsplit == 2) {
              uint4* dl = reinterpret_cast<uint4*>(p.out_lo + off);
#pragma unroll
              for (int j = 0; j < 4; ++j) dl[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
            }
          }
        }
        // this warp is done reading the TMEM region: hand it back to the UMMA issuer early
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 1) mbar_arrive(&tmem_empty[r]);
          else mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[r]), 0));
        }
        if (p.mode == TC_GATE) {
          // the warp owning the upper column half hands its partial sums to its partner (fixed
          // order: deterministic rounding), which updates out8 (unit 0 of a tile starts or continues
          // the layer sum, unit 1 always continues)
          if (half == 1) {
#pragma unroll
            for (int o = 0; o < TC_NOUT; ++o) x8_s[row * TC_NOUT + o] = acc8[o];
          }
          asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
          if (half == 0 && valid) {
#pragma unroll
            for (int o = 0; o < TC_NOUT; ++o) acc8[o] += x8_s[row * TC_NOUT + o];
            float4* o8 = reinterpret_cast<float4*>(p.out8 + col * TC_NOUT);
            float4 o0 = make_float4(acc8[0], acc8[1], acc8[2], acc8[3]);
            float4 o1 = make_float4(acc8[4], acc8[5], acc8[6], acc8[7]);
            if (p.accumulate_out8 || nb > 0) {
              const float4 p0 = o8[0], p1 = o8[1];
              o0.x += p0.x; o0.y += p0.y; o0.z += p0.z; o0.w += p0.w;
              o1.x += p1.x; o1.y += p1.y; o1.z += p1.z; o1.w += p1.w;
            }
            o8[0] = o0;
            o8[1] = o1;
          }
          asm volatile("bar.sync 2, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
        }
        epi_busy += clock64() - e1;
      }
    }
    if (p.prof && warp == 2 && lane == 0) {
      p.prof[blockIdx.x * 8 + 4] = epi_wait;
      p.prof[blockIdx.x * 8 + 5] = epi_busy;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();     // nobody leaves while the peer may still signal / read this CTA
  tc_fence_after();
  if (warp == 1) {
    if (CG == 1) tmem_dealloc(tmem_base, TC_NMAX);
    else tmem_dealloc_cg2(tmem_base, TC_NMAX);
  }
}

// mel (rows, n_mel) fp32 -> zero-padded (rows, pad) bf16 hi/lo operand copies of the upsampler GEMM.
__global__ void mel_pad_split_kernel(const float* __restrict__ mel, __nv_bfloat16* __restrict__ hi,
                                     __nv_bfloat16* __restrict__ lo, long long n_rows, int n_mel, int pad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * pad) return;
  const long long r = i / pad;
  const int c = (int)(i - r * pad);
  const float v = c < n_mel ? __ldg(mel + r * n_mel + c) : 0.f;
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[i] = h;
  if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// x = start(audio_0) (glow.py:156), written as the bf16 operand copies of the first layer.  A pure streaming
// kernel (reads 16-32 B, writes 1 KB per column): thread = 8 channels x WN_START_COLS consecutive columns (the
// weights stay in registers), 16-byte stores.
constexpr int WN_START_COLS = 4;
__global__ void __launch_bounds__(256) wn_start_tc_kernel(const float* __restrict__ audio, const float* __restrict__ w,
                                                          const float* __restrict__ bias, __nv_bfloat16* __restrict__ x_hi,
                                                          __nv_bfloat16* __restrict__ x_lo, long long n_cols, int C,
                                                          int n_group, int off, int n_half) {
  // programmatic dependent launch (see csrc/waveglow_fused.cu): this kernel may have started before the previous
  // flow's `end` (or the upsampler) completed; once that is certain, the first layer's kernel may start its prologue
  // and its conditioning K steps.  (Wait BEFORE the trigger: the layer kernels read `spect` before their own wait, and
  // every one of them starts after this point -- the upsampler's output is complete and visible by then.)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int c8 = C >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long col0 = (idx / c8) * WN_START_COLS;
  if (col0 >= n_cols) return;
  const int c = (int)(idx % c8) * 8;
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c));
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c + 4));
  float4 w0[4], w1[4];                      // n_half <= 4 (n_group <= 8)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    w0[j] = j < n_half ? __ldg(reinterpret_cast<const float4*>(w + (long long)j * C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    w1[j] = j < n_half ? __ldg(reinterpret_cast<const float4*>(w + (long long)j * C + c + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < WN_START_COLS; ++i) {
    const long long col = col0 + i;
    if (col >= n_cols) break;
    float4 o0 = b0, o1 = b1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < n_half) {
        const float a = __ldg(audio + col * n_group + off + j);
        o0.x = fmaf(a, w0[j].x, o0.x);
        o0.y = fmaf(a, w0[j].y, o0.y);
        o0.z = fmaf(a, w0[j].z, o0.z);
        o0.w = fmaf(a, w0[j].w, o0.w);
        o1.x = fmaf(a, w1[j].x, o1.x);
        o1.y = fmaf(a, w1[j].y, o1.y);
        o1.z = fmaf(a, w1[j].z, o1.z);
        o1.w = fmaf(a, w1[j].w, o1.w);
      }
    }
    uint32_t h[4], l[4];
    split2(o0.x, o0.y, h[0], l[0]);
    split2(o0.z, o0.w, h[1], l[1]);
    split2(o1.x, o1.y, h[2], l[2]);
    split2(o1.z, o1.w, h[3], l[3]);
    *reinterpret_cast<uint4*>(x_hi + col * C + c) = make_uint4(h[0], h[1], h[2], h[3]);
    if (x_lo) *reinterpret_cast<uint4*>(x_lo + col * C + c) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// out = out8 + bias8 is end(skip sum) (glow.py:175); b, s = halves (glow.py:278-279);
// a1 <- (a1 - b) / exp(s) (glow.py:280); z <- W^-1 [a0; a1] (glow.py:283, 96).  One thread per column.
__global__ void wn_end_tc_kernel(const float* __restrict__ out8, const float* __restrict__ bias8,
                                 const float* __restrict__ w_inv, float* __restrict__ audio, long long n_cols,
                                 int n_group, int n_rem, int n_half) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // the next flow's `start`
  asm volatile("griddepcontrol.wait;" ::: "memory");                  // out8 is the last layer's
  const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= n_cols) return;
  const float4 a = __ldg(reinterpret_cast<const float4*>(out8 + col * TC_NOUT));
  const float4 bq = __ldg(reinterpret_cast<const float4*>(out8 + col * TC_NOUT) + 1);
  const float o[TC_NOUT] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
  const int off = n_group - n_rem;
  float* acol = audio + col * n_group + off;
  float y[TC_NOUT];
#pragma unroll
  for (int j = 0; j < TC_NOUT; ++j) {
    y[j] = 0.f;
    if (j < n_rem) {
      const float av = acol[j];
      if (j < n_half) {
        y[j] = av;
      } else {
        float bshift = 0.f, s = 0.f;
#pragma unroll
        for (int k = 0; k < TC_NOUT; ++k) {     // static indexing keeps o[] in registers
          if (k == j - n_half) bshift = o[k] + __ldg(bias8 + k);
          if (k == j) s = o[k] + __ldg(bias8 + k);
        }
        y[j] = (av - bshift) / expf(s);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < TC_NOUT; ++i) {
    if (i < n_rem) {
      float z = 0.f;
#pragma unroll
      for (int j = 0; j < TC_NOUT; ++j)
        if (j < n_rem) z = fmaf(__ldg(w_inv + i * n_rem + j), y[j], z);
      acol[i] = z;
    }
  }
}

// ---------------------------------------------------------------- host side (tensor maps: tc_host.cuh)
int g_tc_bk = 0;   // 0 = automatic
inline int tc_pick_bk(int) { return g_tc_bk ? g_tc_bk : 64; }   // measured: 64 wins in both modes (profiles/README.md)
int g_tc_batch_group = 0; // utterances per pass of fac_waveglow_infer_tc over a flow; 0 = the whole batch
int g_tc_cta_group = 0;   // 0 = automatic: CTA pairs for the split-bf16 mode, single CTAs for plain bf16

// Measured on B200 (profiles/README.md): pairs win 9 % in split-bf16 (operand traffic and shared-memory reads
// per UMMA drop by a third, 6-stage ring); plain bf16 is TMA-feed-bound either way and is 3 % faster unpaired.
inline int tc_pick_cg(int nsplit) { return g_tc_cta_group ? g_tc_cta_group : (nsplit == 2 ? 2 : 1); }

template <int CG, int BK>
int launch_tc_cg(const CUtensorMap maps[6], const TcParams& p, cudaStream_t st) {
  static bool attr_set_on[FAC_MAX_DEVICES] = {};      // function attributes are per device
  bool& attr_set = attr_set_on[current_device_slot()];
  constexpr int smem = TcCfg<CG, BK>::SMEM;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wn_gemm_tc_kernel<CG, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("wn_gemm_tc: cannot reserve %d bytes of shared memory: %s", smem, cudaGetErrorString(e));
      return 2;
    }
    attr_set = true;
  }
  const int n_blocks = ceil_div(p.n_total, TC_NHALF);
  const int groups = ceil_div(p.n_tiles, CG) * (p.flat_units ? n_blocks : 1) * (p.n_phases > 0 ? p.n_phases : 1);
  const int max_groups = sm_count() / CG;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(CG * (groups < max_groups ? groups : max_groups)));
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, wn_gemm_tc_kernel<CG, BK>, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], p);
  count_launch();
  if (e != cudaSuccess) {
    set_error("wn_gemm_tc_kernel<%d>: launch failed: %s", CG, cudaGetErrorString(e));
    return 2;
  }
  return check_launch("wn_gemm_tc_kernel");
}

int launch_tc(const CUtensorMap maps[6], const TcParams& p, cudaStream_t st, int cg) {
  const int n_cols = p.n_total < TC_NHALF ? p.n_total : TC_NHALF;
  FAC_REQUIRE(cg == 1 || (n_cols / 2) % 8 == 0, "wn_gemm_tc: %d columns cannot be split over a CTA pair", n_cols);
  const int bk = tc_pick_bk(p.nsplit);
  if (bk == 64) return cg == 2 ? launch_tc_cg<2, 64>(maps, p, st) : launch_tc_cg<1, 64>(maps, p, st);
  return cg == 2 ? launch_tc_cg<2, 32>(maps, p, st) : launch_tc_cg<1, 32>(maps, p, st);
}

long long* g_tc_prof = nullptr;

// ---------------------------------------------------------------- generic Conv1d / Linear on the tensor cores
// (B, C, T) channel-major fp32 -> (B, T, pad) channels-last bf16 hi/lo (the PPG input of the encoder prenet,
// reference src/common/model.py:237-241): 32 x 32 tiles through shared memory.
__device__ __forceinline__ void store_split(unsigned short* hi, unsigned short* lo, long long o, float v, bool fp16) {
  if (fp16) {
    v = fminf(fmaxf(v, -65504.f), 65504.f);          // saturate instead of inf / NaN (see split2_f16)
    const __half h = __float2half_rn(v);
    hi[o] = __half_as_ushort(h);
    if (lo) lo[o] = __half_as_ushort(__float2half_rn(v - __half2float(h)));
  } else {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[o] = __bfloat16_as_ushort(h);
    if (lo) lo[o] = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(h)));
  }
}

// (n_rows, C) fp32 -> zero-padded (n_rows, pad) 16-bit hi/lo operand copies.
__global__ void pad_split_kernel(const float* __restrict__ in, unsigned short* __restrict__ hi,
                                 unsigned short* __restrict__ lo, long long n_rows, int C, int pad, bool fp16) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * pad) return;
  const long long r = i / pad;
  const int c = (int)(i - r * pad);
  store_split(hi, lo, i, c < C ? __ldg(in + r * C + c) : 0.f, fp16);
}

__global__ void transpose_split_kernel(const float* __restrict__ in, unsigned short* __restrict__ hi,
                                       unsigned short* __restrict__ lo, int C, int T, int pad, bool fp16) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;      // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + tx;
    tile[i][tx] = (c < C && t < T) ? __ldg(in + ((long long)b * C + c) * T + t) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + tx;
    if (t < T && c < pad) store_split(hi, lo, ((long long)b * T + t) * pad + c, tile[tx][i], fp16);
  }
}

}  // namespace

int tc_transpose_split(const float* in, void* hi, void* lo, int B, int C, int T, int pad, int fp16, cudaStream_t st) {
  FAC_REQUIRE(in && hi && B > 0 && C > 0 && T > 0 && pad >= C, "transpose_split: bad arguments");
  dim3 grid((unsigned)ceil_div(T, 32), (unsigned)ceil_div(pad, 32), (unsigned)B);
  transpose_split_kernel<<<grid, dim3(32, 8), 0, st>>>(in, reinterpret_cast<unsigned short*>(hi),
                                                       reinterpret_cast<unsigned short*>(lo), C, T, pad, fp16 != 0);
  count_launch();
  return check_launch("transpose_split_kernel");
}

int tc_pad_split(const float* in, void* hi, void* lo, long long n_rows, int C, int pad, int fp16, cudaStream_t st) {
  FAC_REQUIRE(in && hi && n_rows > 0 && C > 0 && pad >= C, "pad_split: bad arguments");
  pad_split_kernel<<<(unsigned)((n_rows * pad + 255) / 256), 256, 0, st>>>(
      in, reinterpret_cast<unsigned short*>(hi), reinterpret_cast<unsigned short*>(lo), n_rows, C, pad, fp16 != 0);
  count_launch();
  return check_launch("pad_split_kernel");
}

int conv_gemm_tc(const fac_tc_conv* c, cudaStream_t st) {
  FAC_REQUIRE(c != nullptr, "conv_gemm_tc: NULL descriptor");
  FAC_REQUIRE(c->nsplit == 1 || c->nsplit == 2, "conv_gemm_tc: nsplit must be 1 or 2 (got %d)", c->nsplit);
  FAC_REQUIRE(c->a_hi && c->w_hi && (c->nsplit == 1 || (c->a_lo && c->w_lo)), "conv_gemm_tc: NULL operand");
  FAC_REQUIRE(c->B > 0 && c->T > 0 && c->taps > 0, "conv_gemm_tc: empty problem");
  FAC_REQUIRE(c->c_pad % TC_BK_MAX == 0, "conv_gemm_tc: input channels %d must be padded to a multiple of %d", c->c_pad,
              TC_BK_MAX);
  FAC_REQUIRE(c->n_pad % 64 == 0 && c->n_valid > 0 && c->n_valid <= c->n_pad && c->n_valid % 4 == 0,
              "conv_gemm_tc: n_pad %d must be a multiple of 64 and n_valid %d a multiple of 4", c->n_pad, c->n_valid);
  FAC_REQUIRE(c->out || c->out_hi, "conv_gemm_tc: no output");
  FAC_REQUIRE(c->out_hi == nullptr || c->nsplit == 1 || c->out_lo, "conv_gemm_tc: out_lo missing");
  const int cg = tc_pick_cg(c->nsplit), bk = tc_pick_bk(c->nsplit);
  const int K = c->taps * c->c_pad;
  CUtensorMap maps[6];
  if (int rc = make_act_map(&maps[0], c->a_hi, c->B, c->T, c->c_pad, bk)) return rc;
  if (int rc = make_act_map(&maps[1], c->nsplit == 2 ? c->a_lo : c->a_hi, c->B, c->T, c->c_pad, bk)) return rc;
  maps[2] = maps[0];
  maps[3] = maps[1];
  if (int rc = make_weight_map(&maps[4], c->w_hi, c->n_pad, K, cg, bk)) return rc;
  if (int rc = make_weight_map(&maps[5], c->nsplit == 2 ? c->w_lo : c->w_hi, c->n_pad, K, cg, bk)) return rc;
  TcParams p{};
  p.T = c->T;
  p.B = c->B;
  p.tiles_per_batch = ceil_div(c->T, TC_BM);
  p.n_tiles = c->B * p.tiles_per_batch;
  p.nsplit = c->nsplit;
  p.C = c->n_pad;                 // row stride of the 16-bit outputs
  p.n_src = 1;
  p.src[0] = TcSrc{c->c_pad, c->taps, 1, c->center};
  p.n_total = c->n_pad;
  p.mode = TC_LINEAR;
  p.flat_units = 1;
  p.out_row_mul = 1;
  p.fp16 = c->fp16;
  p.n_valid = c->n_valid;
  // The tensor core's fp32 accumulator truncates on every accumulation, an error that grows with the number
  // of UMMAs chained into one accumulator (measured: ~3e-5 relative at K = 5824).  With k_chunk > 0 every unit
  // walks its K range in chains of at most k_chunk elements, each in its own TMEM accumulator, whose partial
  // sums are added in fp32 (round-to-nearest) by the epilogue through `scratch`; the last chain applies the
  // real epilogue.
  const int total_steps = K / bk;
  p.k_steps = total_steps;
  p.wlo_k_steps = total_steps;
  p.chunk_steps = 0;
  if (c->k_chunk > 0) {
    FAC_REQUIRE(c->k_chunk % bk == 0, "conv_gemm_tc: k_chunk %d must be a multiple of %d", c->k_chunk, bk);
    if (c->k_chunk / bk < total_steps) {
      FAC_REQUIRE(c->scratch != nullptr, "conv_gemm_tc: K-chunking needs scratch");
      p.chunk_steps = c->k_chunk / bk;
      p.scratch = c->scratch;
      p.scratch_ld = c->n_valid;
    }
  }
  p.bias = c->bias;
  p.act = c->act;
  p.mask = c->mask;
  p.mask_ld = c->mask_ld;
  p.residual = c->residual;
  p.res_ld = c->res_ld;
  p.out_f32 = c->out;
  p.out_ld = c->out_ld;
  p.out_hi = reinterpret_cast<__nv_bfloat16*>(c->out_hi);
  p.out_lo = reinterpret_cast<__nv_bfloat16*>(c->out_lo);
  p.row_lengths = c->row_lengths;
  return launch_tc(maps, p, st, cg);
}

namespace {

}  // namespace

void tc_set_prof(long long* p) { g_tc_prof = p; }
int tc_set_k_block(int bk) {
  FAC_REQUIRE(bk == 0 || bk == 32 || bk == 64, "K block must be 0 (auto), 32 or 64 (got %d)", bk);
  g_tc_bk = bk;
  return 0;
}
int tc_set_batch_group(int n) {
  FAC_REQUIRE(n >= 0, "batch group must be >= 0 (got %d)", n);
  g_tc_batch_group = n;
  return 0;
}
// L2 eviction hints of the fused kernel's TMA loads: 1 = keep the boxes that are read again, 2 = single-use boxes go
// first, 4 = keep the weights.  Measured (8 x 10 s): 3 cuts the layer's DRAM traffic from 1.67 to 1.39 GB and the step
// by 1.3 %; adding 4 gives both back.
int g_tc_fused = 2, g_tc_prefetch = 0, g_tc_l2_hints = 3;
// programmatic dependent launch between the kernels of a flow step (FAC_TC_PDL=0 turns it off for A/B runs)
bool tc_pdl_enabled() {
  static const bool on = getenv("FAC_TC_PDL") == nullptr || atoi(getenv("FAC_TC_PDL")) != 0;
  return on;
}
// mode & 15: 0 = two launches per layer; 1 = one fused launch per layer; 2 = one launch per flow step where possible;
// 3 = like 2, but start and end stay separate kernels (three launches per flow step).
// mode >> 4: L2 prefetch distance of the fused kernel's producer in K steps (experiments)
int tc_set_fused(int mode) {
  FAC_REQUIRE((mode & 15) <= 3 && mode >= 0, "fused mode must be 0 .. 3 (+ 16 x prefetch distance), got %d", mode);
  g_tc_fused = mode & 15;
  g_tc_prefetch = (mode >> 4) & 63;
  if ((mode >> 10) & 15) g_tc_l2_hints = ((mode >> 10) & 15) - 1;     // + 1024 x (h + 1): L2 hint mask h (experiments)
  return 0;
}
int tc_set_cta_group(int cg) {
  FAC_REQUIRE(cg >= 0 && cg <= 2, "cta group must be 0 (auto), 1 or 2 (got %d)", cg);
  g_tc_cta_group = cg;
  return 0;
}

int wg_check_model(const fac_wg_model* m);
bool wn_fused_supported(int C, int n_cond, int bk);
bool wn_fused_weights_ok(const fac_wg_model* m, const fac_wg_tc_flow& wf);
int wn_flow_fused(const fac_wg_model* m, const fac_wg_tc_weights* w, int flow, const fac_wg_tc_workspace* ws,
                  float* audio, int B, int T, int nsplit, int layer_first, int layer_count, int do_start, int do_end,
                  int bk, int prefetch_steps, int l2_hints, long long* prof, cudaStream_t st);

static int tc_check(const fac_wg_model* m, const fac_wg_tc_weights* w, const fac_wg_tc_workspace* ws, int nsplit) {
  if (int rc = wg_check_model(m)) return rc;
  FAC_REQUIRE(w && ws, "tensor-core path: NULL weights/workspace");
  FAC_REQUIRE(nsplit == 1 || nsplit == 2, "tensor-core path: nsplit must be 1 (bf16) or 2 (split-bf16), got %d", nsplit);
  const int C = m->n_channels, n_cond = m->n_mel * m->n_group;
  FAC_REQUIRE(C % TC_BK_MAX == 0 && n_cond % TC_BK_MAX == 0 && 2 * C <= TC_NMAX && C <= TC_CMAX,
              "tensor-core path: needs n_channels %% 16 == 0 (<= %d) and n_cond %% 16 == 0", TC_CMAX);
  FAC_REQUIRE(m->n_group <= TC_NOUT, "tensor-core path: n_group %d > %d", m->n_group, TC_NOUT);
  const bool fused_ws = ws->x2_hi && (nsplit == 1 || ws->x2_lo);
  FAC_REQUIRE(ws->spect_hi && ws->x_hi && (ws->acts_hi || fused_ws) && ws->out8, "tensor-core path: workspace incomplete");
  if (nsplit == 2) FAC_REQUIRE(ws->spect_lo && ws->x_lo && (ws->acts_lo || fused_ws), "tensor-core path: lo buffers missing");
  return 0;
}

// K block of the fused layer kernel: 32 unless overridden (next to its 64 KB acts tile and the x staging the
// operand ring is 128 KB: 4 stages of K = 32, but only 2 of K = 64, which starves the issuer -- measured)
static int tc_fused_bk() { return g_tc_bk ? g_tc_bk : 32; }

// One fused launch per layer (waveglow_fused.cu) when the workspace carries the second residual-stream pair.
static bool tc_use_fused(const fac_wg_model* m, const fac_wg_tc_flow& wf, const fac_wg_tc_workspace* ws, int nsplit,
                         int layer) {
  // (the fused kernel always works in CTA pairs; a forced cta_group of 1 selects the two-launch form)
  const bool pairs = g_tc_cta_group == 0 || g_tc_cta_group == 2;
  const int bk = tc_fused_bk();
  return g_tc_fused && pairs && ws->x2_hi && (nsplit == 1 ? bk == 32 : ws->x2_lo != nullptr) && wn_fused_weights_ok(m, wf) &&
         wn_fused_supported(m->n_channels, m->n_mel * m->n_group, bk);
}

int wg_tc_prepare_spect(const fac_wg_model* m, const fac_wg_tc_weights* w, const fac_wg_tc_workspace* ws,
                        const float* mel_cl, int B, int F, int nsplit, cudaStream_t st) {
  // glow.py:253-259 on the tensor cores: the transposed conv is `phases` GEMMs over the mel frames
  // (tap k of phase p reads frame f - k), written straight into the squeezed bf16 hi/lo layout.
  if (int rc = wg_check_model(m)) return rc;
  FAC_REQUIRE(w && w->up_hi && ws && ws->mel_hi && ws->spect_hi && mel_cl, "prepare_spect: NULL argument");
  FAC_REQUIRE(nsplit == 1 || (w->up_lo && ws->mel_lo && ws->spect_lo), "prepare_spect: lo buffers missing");
  const int n_cond = m->n_mel * m->n_group, phases = m->hop / m->n_group, pad = w->mel_pad;
  FAC_REQUIRE(pad % TC_BK_MAX == 0 && pad >= m->n_mel, "prepare_spect: mel_pad %d must be a multiple of %d", pad,
              TC_BK_MAX);
  const int bk = tc_pick_bk(nsplit);
  const long long n_rows = (long long)B * F;
  mel_pad_split_kernel<<<(unsigned)((n_rows * pad + 255) / 256), 256, 0, st>>>(
      mel_cl, reinterpret_cast<__nv_bfloat16*>(ws->mel_hi),
      nsplit == 2 ? reinterpret_cast<__nv_bfloat16*>(ws->mel_lo) : nullptr, n_rows, m->n_mel, pad);
  count_launch();
  if (int rc = check_launch("mel_pad_split_kernel")) return rc;
  CUtensorMap maps[6];
  if (int rc = make_act_map(&maps[0], ws->mel_hi, B, F, pad, bk)) return rc;
  if (int rc = make_act_map(&maps[1], nsplit == 2 ? ws->mel_lo : ws->mel_hi, B, F, pad, bk)) return rc;
  maps[2] = maps[0];
  maps[3] = maps[1];
  TcParams p{};
  p.T = F;
  p.B = B;
  p.tiles_per_batch = ceil_div(F, TC_BM);
  p.n_tiles = B * p.tiles_per_batch;
  p.nsplit = nsplit;
  p.C = n_cond;
  p.n_src = 1;
  p.src[0] = TcSrc{pad, m->upsample_taps, -1, 0};
  p.n_total = n_cond;
  p.k_steps = m->upsample_taps * pad / bk;
  p.wlo_k_steps = p.k_steps;
  p.mode = TC_RESIDUAL;
  p.bias = m->upsample_b;
  p.out_hi = reinterpret_cast<__nv_bfloat16*>(ws->spect_hi);
  p.out_lo = reinterpret_cast<__nv_bfloat16*>(ws->spect_lo);
  p.out_row_mul = phases;
  // all phases in ONE launch: the phase matrices are consecutive, so one weight map over phases * n_cond rows serves
  // them all, and the tile index space is phases x (tiles rounded up to whole CTA pairs) -- 20 launches of 1.2 waves
  // each become one of ~12 full waves
  const int cg = tc_pick_cg(nsplit);
  if (int rc = make_weight_map(&maps[4], w->up_hi, phases * n_cond, m->upsample_taps * pad, cg, bk)) return rc;
  if (int rc = make_weight_map(&maps[5], nsplit == 2 ? w->up_lo : w->up_hi, phases * n_cond, m->upsample_taps * pad, cg, bk)) return rc;
  p.out_row_off = 0;
  p.n_phases = phases;
  p.tiles_pp = ceil_div(p.n_tiles, cg) * cg;
  p.w_phase_rows = n_cond;
  return launch_tc(maps, p, st, cg);
}

int wg_tc_start(const fac_wg_model* m, int flow, const float* audio, const fac_wg_tc_workspace* ws, int B, int Tg,
                int nsplit, cudaStream_t st) {
  if (int rc = wg_check_model(m)) return rc;
  FAC_REQUIRE(flow >= 0 && flow < m->n_flows && ws && ws->x_hi, "wn_start_tc: bad arguments");
  const fac_wg_flow& f = m->flows[flow];
  const long long n_cols = (long long)B * Tg;
  FAC_REQUIRE(f.n_half <= 4, "wn_start_tc: n_half %d > 4", f.n_half);
  const long long total = ((n_cols + WN_START_COLS - 1) / WN_START_COLS) * (m->n_channels / 8);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((total + 255) / 256));
  cfg.blockDim = dim3(256);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tc_pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, wn_start_tc_kernel, audio, (const float*)f.start_w, (const float*)f.start_b,
                                     reinterpret_cast<__nv_bfloat16*>(ws->x_hi),
                                     nsplit == 2 ? reinterpret_cast<__nv_bfloat16*>(ws->x_lo) : (__nv_bfloat16*)nullptr, n_cols,
                                     (int)m->n_channels, (int)m->n_group, (int)(m->n_group - f.n_rem), (int)f.n_half);
  count_launch();
  if (e != cudaSuccess) {
    set_error("wn_start_tc_kernel: launch failed: %s", cudaGetErrorString(e));
    return 2;
  }
  return check_launch("wn_start_tc_kernel");
}

int wg_tc_end(const fac_wg_model* m, const fac_wg_tc_weights* w, int flow, const float* out8, float* audio, int B,
              int Tg, cudaStream_t st) {
  if (int rc = wg_check_model(m)) return rc;
  FAC_REQUIRE(flow >= 0 && flow < m->n_flows && w && out8 && audio, "wn_end_tc: bad arguments");
  const fac_wg_flow& f = m->flows[flow];
  const long long n_cols = (long long)B * Tg;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((n_cols + 255) / 256));
  cfg.blockDim = dim3(256);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tc_pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, wn_end_tc_kernel, out8, (const float*)w->flows[flow].out_bias, (const float*)f.w_inv,
                                     audio, n_cols, (int)m->n_group, (int)f.n_rem, (int)f.n_half);
  count_launch();
  if (e != cudaSuccess) {
    set_error("wn_end_tc_kernel: launch failed: %s", cudaGetErrorString(e));
    return 2;
  }
  return check_launch("wn_end_tc_kernel");
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
  if (tc_use_fused(m, wf, ws, nsplit, layer)) {
    // the residual stream ping-pongs: layer i reads x (i even) / x2 (i odd) and writes the other pair
    return wn_flow_fused(m, w, flow, ws, nullptr, B, Tg, nsplit, layer, 1, 0, 0, tc_fused_bk(), g_tc_prefetch & 63, g_tc_l2_hints, g_tc_prof, st);
  }
  FAC_REQUIRE(ws->acts_hi && (nsplit == 1 || ws->acts_lo), "wn_layer_tc: the two-launch form needs the acts buffers");
  CUtensorMap maps[6];
  TcParams p{};
  p.T = Tg;
  p.B = B;
  p.tiles_per_batch = ceil_div(Tg, TC_BM);
  p.n_tiles = B * p.tiles_per_batch;
  p.nsplit = nsplit;
  p.C = C;
  p.prof = g_tc_prof;
  p.out_row_mul = 1;
  // ---- G1: [x taps | spect] -> gate -> acts, out8 += Wc acts
  const int K1 = ks * C + n_cond;
  const int cg = tc_pick_cg(nsplit), bk = tc_pick_bk(nsplit);
  if (int rc = make_act_map(&maps[0], ws->x_hi, B, Tg, C, bk)) return rc;
  if (int rc = make_act_map(&maps[1], nsplit == 2 ? ws->x_lo : ws->x_hi, B, Tg, C, bk)) return rc;
  if (int rc = make_act_map(&maps[2], ws->spect_hi, B, Tg, n_cond, bk)) return rc;
  if (int rc = make_act_map(&maps[3], nsplit == 2 ? ws->spect_lo : ws->spect_hi, B, Tg, n_cond, bk)) return rc;
  if (int rc = make_weight_map(&maps[4], wf.w1_hi[layer], 2 * C, K1, cg, bk)) return rc;
  if (int rc = make_weight_map(&maps[5], nsplit == 2 ? wf.w1_lo[layer] : wf.w1_hi[layer], 2 * C, K1, cg, bk)) return rc;
  p.n_src = 2;
  p.src[0] = TcSrc{C, ks, dil, dil * (ks - 1) / 2};
  p.src[1] = TcSrc{n_cond, 1, 0, 0};
  p.n_total = 2 * C;
  p.k_steps = K1 / bk;
  p.wlo_k_steps = p.k_steps;
  p.mode = TC_GATE;
  p.bias = f.in_cond_b[layer];
  p.out_hi = reinterpret_cast<__nv_bfloat16*>(ws->acts_hi);
  p.out_lo = reinterpret_cast<__nv_bfloat16*>(ws->acts_lo);
  p.wc = wf.wc[layer];
  p.out8 = ws->out8;
  p.accumulate_out8 = layer > 0;
  if (int rc = launch_tc(maps, p, st, cg)) return rc;
  if (last) return 0;   // the last layer feeds the skip path only (glow.py:168-169)
  // ---- G2: x <- [acts | x] [W_res | I]^T + b_res
  if (int rc = make_act_map(&maps[0], ws->acts_hi, B, Tg, C, bk)) return rc;
  if (int rc = make_act_map(&maps[1], nsplit == 2 ? ws->acts_lo : ws->acts_hi, B, Tg, C, bk)) return rc;
  if (int rc = make_act_map(&maps[2], ws->x_hi, B, Tg, C, bk)) return rc;
  if (int rc = make_act_map(&maps[3], nsplit == 2 ? ws->x_lo : ws->x_hi, B, Tg, C, bk)) return rc;
  if (int rc = make_weight_map(&maps[4], wf.w2_hi[layer], C, 2 * C, cg, bk)) return rc;
  if (int rc = make_weight_map(&maps[5], nsplit == 2 ? wf.w2_lo[layer] : wf.w2_hi[layer], C, 2 * C, cg, bk)) return rc;
  p.n_src = 2;
  p.src[0] = TcSrc{C, 1, 0, 0};
  p.src[1] = TcSrc{C, 1, 0, 0};
  p.n_total = C;
  p.k_steps = 2 * C / bk;
  p.wlo_k_steps = C / bk;           // the identity block has no lo part
  p.mode = TC_RESIDUAL;
  p.bias = wf.res_b[layer];
  p.out_hi = reinterpret_cast<__nv_bfloat16*>(ws->x_hi);
  p.out_lo = reinterpret_cast<__nv_bfloat16*>(ws->x_lo);
  p.wc = nullptr;
  p.out8 = nullptr;
  if (p.prof) p.prof += 8 * 256;   // G2 counters follow G1's
  return launch_tc(maps, p, st, cg);
}

int wg_tc_flow_step(const fac_wg_model* m, const fac_wg_tc_weights* w, int flow, float* audio,
                    const fac_wg_tc_workspace* ws, int B, int Tg, int nsplit, cudaStream_t st) {
  if (int rc = tc_check(m, w, ws, nsplit)) return rc;
  FAC_REQUIRE(flow >= 0 && flow < m->n_flows && audio, "waveglow_flow_step_tc: bad arguments");
  if (g_tc_fused == 2 && ws->flow_sync && tc_use_fused(m, w->flows[flow], ws, nsplit, 0))
    return wn_flow_fused(m, w, flow, ws, audio, B, Tg, nsplit, 0, m->n_layers, 1, 1, tc_fused_bk(), g_tc_prefetch & 63, g_tc_l2_hints, g_tc_prof, st);
  if (g_tc_fused == 3 && ws->flow_sync && m->n_layers > 1 && tc_use_fused(m, w->flows[flow], ws, nsplit, 0)) {
    // every WN layer of the step in one launch; start and end as their own (whole-chip) kernels
    if (int rc = wg_tc_start(m, flow, audio, ws, B, Tg, nsplit, st)) return rc;
    if (int rc = wn_flow_fused(m, w, flow, ws, audio, B, Tg, nsplit, 0, m->n_layers, 0, 0, tc_fused_bk(), g_tc_prefetch & 63, g_tc_l2_hints, g_tc_prof, st)) return rc;
    return wg_tc_end(m, w, flow, ws->out8, audio, B, Tg, st);
  }
  if (int rc = wg_tc_start(m, flow, audio, ws, B, Tg, nsplit, st)) return rc;
  for (int i = 0; i < m->n_layers; ++i)
    if (int rc = wg_tc_layer(m, w, flow, i, ws, B, Tg, nsplit, st)) return rc;
  return wg_tc_end(m, w, flow, ws->out8, audio, B, Tg, st);
}

int wg_infer_tc(const fac_wg_model* m, const fac_wg_tc_weights* w, const float* mel_cl, float* audio,
                const fac_wg_tc_workspace* ws, int B, int F, int nsplit, cudaStream_t st) {
  if (int rc = tc_check(m, w, ws, nsplit)) return rc;
  FAC_REQUIRE(mel_cl && audio, "waveglow_infer_tc: NULL argument");
  const int Tg = F * (m->hop / m->n_group);
  if (int rc = wg_tc_prepare_spect(m, w, ws, mel_cl, B, F, nsplit, st)) return rc;
  // Utterances are independent, so a flow can run group by group: with a group whose residual stream and
  // gated activations (2 KB per column in split mode) fit the L2, the second GEMM of a layer and the next
  // layer's first GEMM find them there instead of in HBM.
  const int group = g_tc_batch_group > 0 ? (g_tc_batch_group < B ? g_tc_batch_group : B) : B;
  const int C = m->n_channels, n_cond = m->n_mel * m->n_group;
  for (int k = m->n_flows - 1; k >= 0; --k) {
    for (int b0 = 0; b0 < B; b0 += group) {
      const int nb = b0 + group <= B ? group : B - b0;
      const long long col0 = (long long)b0 * Tg;
      fac_wg_tc_workspace g = *ws;
      auto shift = [&](void* base, long long elems) -> void* {
        return base ? static_cast<void*>(static_cast<__nv_bfloat16*>(base) + elems) : nullptr;
      };
      g.spect_hi = shift(ws->spect_hi, col0 * n_cond);
      g.spect_lo = shift(ws->spect_lo, col0 * n_cond);
      g.x_hi = shift(ws->x_hi, col0 * C);
      g.x_lo = shift(ws->x_lo, col0 * C);
      g.acts_hi = shift(ws->acts_hi, col0 * C);
      g.acts_lo = shift(ws->acts_lo, col0 * C);
      g.x2_hi = shift(ws->x2_hi, col0 * C);
      g.x2_lo = shift(ws->x2_lo, col0 * C);
      g.out8 = ws->out8 + col0 * TC_NOUT;
      float* audio_g = audio + col0 * m->n_group;
      if (int rc = wg_tc_flow_step(m, w, k, audio_g, &g, nb, Tg, nsplit, st)) return rc;
    }
  }
  return 0;
}

}  // namespace fac
