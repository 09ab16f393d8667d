// One WaveGlow WN layer (reference src/waveglow/glow.py:158-174) as ONE launch on the 5th-generation tensor
// cores: the in+cond GEMM, the gate, the collapsed skip update AND the residual GEMM + residual add, with the
// gated activations handed from the first GEMM's epilogue to the second GEMM through shared memory.
//
// The two-launch form (waveglow_tc.cu) writes `acts` (hi+lo) to HBM and reads it straight back, re-reads and
// re-writes the residual stream x in a second pass and spends ~5 % of its UMMAs on an identity block that
// performs x + res on the tensor core.  Here, per CTA pair (cta_group::2, UMMA M = 256 = two adjacent 128-row
// time tiles) and per tile:
//   U(u)  u = 0,1: pre[256 x 256] = [x(t-d) | x(t) | x(t+d) | spect(t)] W1[u]^T      (K = 3C + n_cond, split-bf16:
//                  3 UMMAs per product) into TMEM region 0
//   E(u)          the 8 epilogue warps DRAIN region 0 into registers (128 fp32 per thread) and release it at
//                  once, then gate tanh*sigmoid -> acts (128 channels of this unit), out8 += Wc acts (fp32), and
//                  write acts as bf16 hi/lo straight into shared memory in the UMMA operand layout (the TMA
//                  swizzle applied by hand)
//   P(u)          res[256 x C] (+)= acts_u W_res[:, u]^T into TMEM region 1 (K = the unit's 128 channels)
//   EG            x_new = res + b_res + x (glow.py:166) -> bf16 hi/lo of the NEXT layer's input buffer
// K order of U: the conditioning (spect) steps first, then the taps of the residual stream (measured: 1 % faster
// than taps first in the one-launch-per-layer form; in the flow form it is what lets a tile start before its
// inputs are stored).
// The UMMA issue order is U(q), P(q-1), U(q+1), P(q) ... over the flattened unit sequence q: every unit is
// followed by a short residual part in the OTHER TMEM region, which hides the register drain of the unit just
// finished, so a single 256-column accumulator region serves the first GEMM without stalls and the other 256
// columns hold the residual accumulator -- all 512 TMEM columns, no double buffering needed.  acts never
// touches HBM; x is read by TMA (taps) + once more by EG from L2 and written once, into a second buffer (the
// neighbouring tiles still read the old x for their dilated taps: layers ping-pong between two x buffers).
// EG moves x through shared memory with TMA in both directions: a third producer thread streams the old x of the
// tile (32-channel boxes, one staging entry per column half) a whole tile ahead, the epilogue threads add in
// place (their own row) and one thread per half stores the box to the other x buffer with a TMA store -- the
// row-per-thread global accesses this replaces cost 32 LSU wavefronts per instruction and made EG the longest
// stage of the epilogue (measured: 33 k cycles per tile; profiles/README.md).
// Shared memory: operand ring (128 KB = 4 stages of K = 32; with K = 64 only 2 stages fit and the issuer starves:
// measured) + the acts operand tile (64 KB) + the x staging (32 KB).  The producer also prefetches the first
// unit's activation boxes of the next tile into L2 (they come from DRAM; the second unit re-reads them from L2).
// Registers: the epilogue warpgroups take 232 registers per thread from the producer / issuer warpgroup
// (setmaxnreg), which is what makes a 128-register drain possible.
// Layer without a residual output (the last one, glow.py:168-169): U / E only, alternating both TMEM regions.
#include "fac_common.cuh"
#include "tc_common.cuh"
#include "tc_host.cuh"

namespace fac {
bool tc_pdl_enabled();      // csrc/waveglow_tc.cu
namespace {

using namespace tc;

constexpr int FU_THREADS = 384;        // warpgroup 0 = {TMA producer, UMMA issuer, 2 idle warps}, warpgroups 1-2 = epilogue
constexpr int FU_EPI_WARPS = 8;
constexpr int FU_EPI_THREADS = 32 * FU_EPI_WARPS;
constexpr int FU_NOUT = 8;             // channels of the collapsed skip path
constexpr int FU_CMAX = 256;
constexpr int FU_MAX_STAGES = 8;
constexpr int FU_RING_BYTES = 128 * 1024;
constexpr int FU_ACTS_BYTES = 64 * 1024;          // 128 rows x 128 channels x (hi + lo) x 2 B
constexpr int FU_XS_COLS = 32;                    // channels of one x staging box (64-byte rows, SWIZZLE_64B)
constexpr int FU_XS_ARRAY = TC_BM * FU_XS_COLS * 2;       // 8 KB: one box (hi or lo)
constexpr int FU_XS_BYTES = 2 * 2 * FU_XS_ARRAY;          // one entry (hi + lo) per column half
constexpr int FU_BAR_BYTES = 512;
constexpr int FU_SMEM = FU_RING_BYTES + FU_ACTS_BYTES + FU_XS_BYTES + FU_BAR_BYTES + 1024;
constexpr int FU_REGS_LOW = 40, FU_REGS_HIGH = 232;   // 128 x 40 + 256 x 232 = 384 x 168

struct FusedParams {
  int T, B, tiles_per_batch, n_tiles;
  int C, n_cond, taps;
  int k1_steps;              // K steps of the first GEMM: (taps * C + n_cond) / BK
  // The launch runs layers [layer_first, layer_first + layer_count) of a WN with n_layers layers (dilation 2^l,
  // the last one has no residual output), optionally preceded by `start` (glow.py:156) and followed by `end` +
  // coupling + invertible 1x1 (glow.py:175, 278-283): the whole flow step in one launch.  Consecutive phases
  // are separated by a grid-wide barrier (cooperative launch); a single layer needs none.
  int layer_first, layer_count, n_layers;
  int do_start, do_end;
  const float* bias1[FAC_MAX_LAYERS];   // [2C] in.bias + cond.bias, gate-interleaved
  const float* res_b[FAC_MAX_LAYERS];   // [C]
  const float* wc[FAC_MAX_LAYERS];      // [8][C] collapsed skip weights
  float* out8;               // (B, T, 8)
  int accumulate_out8;       // single-layer launches: layer_first > 0
  int pdl;                   // launched with programmatic stream serialization: see grid_dep_wait()
  const float* start_w;      // [n_half][C]
  const float* start_b;      // [C]
  const float* out_bias;     // [8] end.bias + end.weight @ (sum of the skip biases)
  const float* w_inv;        // [n_rem][n_rem]
  float* audio;              // (B, T, n_group): the flow's live channels are the LAST n_rem slots of a column
  int n_group, n_rem, n_half;
  __nv_bfloat16* acts_hi;    // optional (tests, single layer): the gated activations, (B, T, C)
  __nv_bfloat16* acts_lo;
  unsigned int* grid_bar;    // [n_tiles] zeroed by the host: finished stores of each tile's residual stream (flow_signal)
  int prefetch_steps;        // L2 prefetch distance of the producer in K steps (0 = off)
  int l2_hints;              // eviction-priority hints on the TMA loads: 1 keep re-read boxes, 2 single-use boxes first, 4 keep weights
  long long* prof;           // optional [grid][16] clock64 counters (tools/tc_cycle_breakdown.py)
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
               ::"l"(m), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// clock64 accumulators of one role (tools/tc_cycle_breakdown.py).  Only the PROF instantiation of the kernel carries
// them: in the single-thread producer / issuer roles they cost a third of the 40 registers those warps keep.
template <bool ON>
struct TickT {
  long long t;
  __device__ __forceinline__ void start() {
    if (ON) t = clock64();
  }
  __device__ __forceinline__ void lap(long long& acc) {
    if (ON) {
      const long long now = clock64();
      acc += now - t;
      t = now;
    }
  }
};

__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t addr, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(addr), "r"(parity)
      : "memory");
  return done != 0;
}
// wait with cluster-scope acquire: the arrivals come from the peer CTA's threads after their shared-memory writes
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  while (!mbar_try_wait_cluster(addr, parity)) {
  }
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void ld_shared_v4(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}

// byte offset inside a K-major operand tile of `ROWB`-byte rows (dense, 1024-byte aligned) -> the offset the TMA
// swizzle mode of that row width (128B / 64B) would have stored it at: address bits [4,7) ^= bits [7,10) (128B),
// bits [4,6) ^= bits [7,9) (64B)
template <int ROWB>
__device__ __forceinline__ uint32_t swizzle_off(uint32_t off) {
  return ROWB == 128 ? off ^ (((off >> 7) & 7u) << 4) : off ^ (((off >> 7) & 3u) << 4);
}

struct FusedMaps {
  CUtensorMap xa_hi, xa_lo, xb_hi, xb_lo;   // residual stream buffers A (ws->x) and B (ws->x2): boxes of BK channels
  CUtensorMap s_hi, s_lo;                   // spect
  CUtensorMap w1_hi, w1_lo;                 // [layer][2C][taps*C + n_cond]
  CUtensorMap w2_hi, w2_lo;                 // [layer][C][C] residual half of res_skip
  CUtensorMap ya_hi, ya_lo, yb_hi, yb_lo;   // the same two buffers in boxes of FU_XS_COLS channels (EG staging, TMA store)
};

// Programmatic dependent launch (single-layer form): a layer's kernel is allowed to start while the previous
// layer's is still running -- on the SMs that one has already left or never used (a short utterance occupies a
// third of the chip) -- and does everything that does not depend on it: prologue, weights, and the conditioning
// K steps of its first unit, which come first in the K order for this reason.  Whoever touches the residual
// stream or out8 waits for the previous grid to have completed (griddepcontrol.wait) first.
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Dependencies between the phases of a flow step (cooperative launch: every CTA is resident).  No grid barrier:
// a time tile of layer l reads the residual stream of its own rows and of the two neighbouring tiles (dilated
// taps reach <= 128 rows), so what it has to wait for is THREE tiles of layer l - 1, not the grid.  flags[t]
// counts the finished TMA stores of tile t's stream (two per written version: one per column half; the storing
// thread waits for its bulk stores, fences the async proxy and adds 1 with release); a neighbour that does not
// exist is replaced by the nearest tile that does.  The CTAs own fixed tiles and walk (layer, tile) in the
// same order, so the waits cannot form a cycle; the write-after-read hazard of the two ping-pong buffers is
// covered by the same waits (a tile writes version v + 1 only after its neighbours have finished the layer
// that read version v - 1 from the buffer it overwrites).
constexpr int FLOW_SPIN_LIMIT = 1 << 24;      // x >= 200 ns: seconds
__device__ __forceinline__ void flow_signal(unsigned int* flags, int tile) {
  // The caller has waited for its bulk stores (cp.async.bulk.wait_group 0: the writes are performed, they sit in
  // L2), so the counter cannot overtake them; a release here would be a full membar that also waits for the
  // CTA's unrelated generic stores (the out8 read-modify-writes) -- measured: ~1 k cycles per tile, 3 % of a step.
  asm volatile("fence.proxy.async;" ::: "memory");
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(flags + tile) : "memory");
}
__device__ __forceinline__ void flow_wait_one(const unsigned int* flags, int idx, unsigned int need) {
  // relaxed polls with a pause (an acquire per poll would invalidate the SM's L1 over and over under the
  // epilogue's feet); the caller fences once when all its counters are there
  unsigned int seen;
  int spins = 0;
  for (;;) {
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(flags + idx) : "memory");
    if (seen >= need) break;
    if (++spins > FLOW_SPIN_LIMIT) __trap();      // a lost signal must not hang the GPU
    __nanosleep(200);
  }
}
// The loading threads of a CTA do not poll global memory themselves (three dependent L2 round trips per tile would
// drain the operand ring): an otherwise idle thread, the watcher, walks the (layer, tile) items ahead of them and
// publishes how many are cleared in a shared-memory word.
__device__ __forceinline__ void flow_wait_cleared(const uint32_t* cleared, uint32_t item) {
  uint32_t seen;
  int spins = 0;
  do {
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(seen) : "r"((uint32_t)__cvta_generic_to_shared(cleared)) : "memory");
    if (++spins > FLOW_SPIN_LIMIT) __trap();
  } while (seen <= item);
  asm volatile("fence.proxy.async;" ::: "memory");
}
// tile `tile` and its neighbours carry at least `writes` versions (0: nothing to wait for)
__device__ __forceinline__ void flow_wait_tiles(const unsigned int* flags, int tile, int n_tiles, int writes, bool neighbours) {
  if (writes <= 0) return;
  const unsigned int need = 2u * (unsigned int)writes;
  const int last = n_tiles - 1, idx = min(tile, last);  // a tile past the end waits for the last one
  flow_wait_one(flags, idx, need);
  if (neighbours) {
    if (idx > 0) flow_wait_one(flags, idx - 1, need);
    if (idx < last) flow_wait_one(flags, idx + 1, need);
  }
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
  asm volatile("fence.proxy.async;" ::: "memory");      // the acquired writes came from, and go to, the async proxy
}

// FLOW = false: the single-layer form (layer_count == 1, no start / end / tile counters): the same code with the
// phase logic compiled out, which keeps the register allocation of the hot loops as tight as it can be.
// NS = 2: split-bf16 operands (hi + lo pairs, 3 UMMAs per product, fp32-grade); NS = 1: plain bf16 (hi only).
template <int BK, int NS, bool FLOW, bool PROF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FU_THREADS, 1)
wn_flow_fused_kernel(const __grid_constant__ FusedMaps maps, const FusedParams p) {
  using Tick = TickT<PROF>;
  constexpr bool SPLIT = NS == 2;
  constexpr int ROWB = BK * 2;                       // bytes of one operand row
  constexpr int A_BYTES = TC_BM * ROWB;              // one 128-row activation tile (hi or lo)
  constexpr int W_BYTES = (TC_NHALF / 2) * ROWB;     // this CTA's half of a 256-row weight block (hi or lo)
  constexpr int STAGE_BYTES = NS * (A_BYTES + W_BYTES);
  constexpr int STAGES = FU_RING_BYTES / STAGE_BYTES < FU_MAX_STAGES ? FU_RING_BYTES / STAGE_BYTES : FU_MAX_STAGES;
  constexpr int W_OFF = NS * A_BYTES;                // stage layout: A_hi [A_lo] W_hi [W_lo]
  static_assert(STAGES >= 2 && STAGES <= FU_MAX_STAGES, "ring");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* acts_s = smem + FU_RING_BYTES;            // [k step][hi | lo][128 rows][ROWB], 1024-byte aligned
  uint8_t* xs_s = acts_s + FU_ACTS_BYTES;            // [column half][hi | lo][128 rows][64 B]
  uint64_t* full = reinterpret_cast<uint64_t*>(xs_s + FU_XS_BYTES);
  uint64_t* empty = full + FU_MAX_STAGES;
  uint64_t* tmem_full = empty + FU_MAX_STAGES;       // [2] accumulator regions (see the issuer)
  uint64_t* tmem_empty = tmem_full + 2;              // [2]
  uint64_t* acts_ready = tmem_empty + 2;             // epilogue -> issuer (leader's copy is used)
  uint64_t* acts_free = acts_ready + 1;              // issuer -> epilogue (both CTAs)
  uint64_t* xs_full = acts_free + 1;                 // [2] x producer -> epilogue half
  uint64_t* xs_empty = xs_full + 2;                  // [2] epilogue half -> x producer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xs_empty + 2);
  uint32_t* flow_cleared = tmem_slot + 2;            // FLOW: (layer, tile) items whose input tiles are complete (watcher)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int tile_first = ((int)blockIdx.x / 2) * 2;
  const int C = p.C;
  const int n_total = 2 * C;
  const int n_units = (n_total + TC_NHALF - 1) / TC_NHALF;          // accumulator blocks of the first GEMM per tile
  const int n_cols = n_total < TC_NHALF ? n_total : TC_NHALF;       // columns of one block
  const int chpu = n_cols / 2;                                      // channels one unit contributes
  const int kp_steps = chpu / BK;                                   // K steps of one residual part
  const int xs_chunks = (C / 2) / FU_XS_COLS;                       // x staging boxes per column half and tile
  int my_tiles = 0;
  for (int tb = tile_first; tb < p.n_tiles; tb += (int)gridDim.x) ++my_tiles;
  const int Q = my_tiles * n_units;                                 // units of this CTA pair per layer
  // phases: [start] layer ... layer [end]; what a tile of a layer needs from the phase before travels through the
  // per-tile counters (flow_signal / flow_wait_tiles); `end` only touches this CTA's own columns
  const int layer_count = FLOW ? p.layer_count : 1;
  const bool do_start = FLOW && p.do_start, do_end = FLOW && p.do_end;
  auto has_res = [&](int l) { return l < p.n_layers - 1; };
  // versions of the residual stream written in THIS launch before layer li reads it (start + the layers with a
  // residual output): what flags[] of a tile must have reached, halved
  auto writes_before = [&](int li) {
    int n = do_start ? 1 : 0;
    for (int i = 0; i < li; ++i) n += has_res(p.layer_first + i) ? 1 : 0;
    return FLOW ? n : 0;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int r = 0; r < 2; ++r) {
      mbar_init(&tmem_full[r], 1);
      mbar_init(&tmem_empty[r], FU_EPI_WARPS * 2);
      mbar_init(&xs_full[r], 1);
      mbar_init(&xs_empty[r], 1);
    }
    mbar_init(acts_ready, FU_EPI_WARPS * 2);
    mbar_init(acts_free, 1);
    *flow_cleared = 0;
    fence_barrier_init();
    tma_prefetch_desc(&maps.xa_hi);
    tma_prefetch_desc(&maps.w1_hi);
  }
  if (warp == 1) tmem_alloc_cg2(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (!FLOW && p.pdl && threadIdx.x == 0) grid_dep_launch();      // the next layer's kernel may start its prologue

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FU_REGS_LOW));
    // ring position and the use counters behind every mbarrier parity run on across the layers of the launch
    int stage = 0;
    uint32_t phase = 0;
    auto advance = [&]() {
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    };
    long long prod_wait = 0, w_tmem0 = 0, w_full = 0, w_acts = 0, w_tmem1 = 0, issue = 0;
    const long long k_start = PROF ? clock64() : 0;
    uint32_t use[2] = {0, 0}, acts_n = 0, xph[2] = {0, 0};
    for (int li = 0; li < layer_count; ++li) {
      const int layer = p.layer_first + li;
      const bool res = has_res(layer);
      const int dil = 1 << layer, center = dil * (p.taps - 1) / 2;
      const bool in_a = (layer & 1) == 0;              // layer l reads buffer A when l is even and writes the other
      if (warp == 0 && lane == 0) {
        // ===================================================== TMA producer (operand ring)
        const CUtensorMap* xm_hi = in_a ? &maps.xa_hi : &maps.xb_hi;
        const CUtensorMap* xm_lo = in_a ? &maps.xa_lo : &maps.xb_lo;
        const int w1_rows = n_cols / 2, w2_rows = C / 2;            // weight rows staged by this CTA
        const int steps_x = p.taps * (C / BK);
        // L2 residency: the activation boxes of a tile's first unit are read again by its later units (and, for x,
        // by the neighbouring tiles' taps) -> keep; the last unit's are not needed again in this layer -> evict first;
        // the weights are shared by every CTA of the grid -> keep
        const uint64_t keep = p.l2_hints ? l2_policy_evict_last() : 0, once = p.l2_hints ? l2_policy_evict_first() : 0;
        auto acquire = [&](uint32_t bytes) -> uint8_t* {
          const long long w0 = PROF ? clock64() : 0;
          mbar_wait(&empty[stage], phase ^ 1);
          if (PROF) prod_wait += clock64() - w0;
          if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * bytes);   // both CTAs' loads land on the leader's barrier
          return smem + stage * STAGE_BYTES;
        };
        // activation box of K step ks of unit q: (map pair, channel, first row, utterance)
        auto a_box = [&](int q, int ks, const CUtensorMap*& mh, const CUtensorMap*& ml, int& c0, int& row0, int& b) {
          const int tile = tile_first + (q / n_units) * (int)gridDim.x + rank;   // may be one past the end: OOB -> zeros
          b = tile / p.tiles_per_batch;
          const int t0 = (tile % p.tiles_per_batch) * TC_BM;
          if (ks < steps_x) {
            const int cps = C / BK, tap = ks / cps;
            c0 = (ks - tap * cps) * BK;
            row0 = t0 + tap * dil - center;
            mh = xm_hi;
            ml = xm_lo;
          } else {
            c0 = (ks - steps_x) * BK;
            row0 = t0;
            mh = &maps.s_hi;
            ml = &maps.s_lo;
          }
        };
        const int need_writes = writes_before(li);
        for (int q = 0; q <= Q; ++q) {
          if (q < Q) {
            const int u = q % n_units;
            const int w_row = u * TC_NHALF + rank * w1_rows;
            // K order: the conditioning (spect) steps first, then the taps of the residual stream.  The spect
            // boxes and all the weights do not depend on the previous layer, so the tensor pipe already works on a
            // tile while the stream of that tile and of its neighbours is still being stored (FLOW); the wait
            // sits right before the first tap.  (ki: position in this order, ks: K step of the weight matrix)
            for (int ki = 0; ki < p.k1_steps; ++ki) {
              const int ks = ki + steps_x < p.k1_steps ? ki + steps_x : ki + steps_x - p.k1_steps;
              if (FLOW && u == 0 && ks == 0 && need_writes > 0)
                flow_wait_cleared(flow_cleared, (uint32_t)(li * my_tiles + q / n_units));
              if (!FLOW && p.pdl && q == 0 && ks == 0) grid_dep_wait();     // the first read of the residual stream
              const CUtensorMap *mh, *ml;
              int c0, row0, b;
              if (p.prefetch_steps > 0) {
                // L2 prefetch, prefetch_steps ahead in the flattened (unit, K step) order, of boxes that will come
                // from DRAM: those of a tile's FIRST unit (its later units find them in L2)
                int q2 = q, ki2 = ki + p.prefetch_steps;
                if (ki2 >= p.k1_steps) {
                  ki2 -= p.k1_steps;
                  ++q2;
                }
                const int ks2 = ki2 + steps_x < p.k1_steps ? ki2 + steps_x : ki2 + steps_x - p.k1_steps;
                if (q2 < Q && q2 % n_units == 0 && (!FLOW || ks2 >= steps_x)) {
                  a_box(q2, ks2, mh, ml, c0, row0, b);
                  tma_prefetch_l2_3d(mh, c0, row0, b);
                  tma_prefetch_l2_3d(ml, c0, row0, b);
                }
              }
              uint8_t* st = acquire((uint32_t)(NS * (A_BYTES + w1_rows * ROWB)));
              const uint32_t lead_full = mapa_u32(smem_u32(&full[stage]), 0);
              a_box(q, ks, mh, ml, c0, row0, b);
              const bool reread = u < n_units - 1 || ks < steps_x;
              if ((p.l2_hints & 1) && reread) {
                tma_load_3d_cg2_hint(st, mh, lead_full, c0, row0, b, keep);
                if (SPLIT) tma_load_3d_cg2_hint(st + A_BYTES, ml, lead_full, c0, row0, b, keep);
              } else if ((p.l2_hints & 2) && !reread) {
                tma_load_3d_cg2_hint(st, mh, lead_full, c0, row0, b, once);
                if (SPLIT) tma_load_3d_cg2_hint(st + A_BYTES, ml, lead_full, c0, row0, b, once);
              } else {
                tma_load_3d_cg2(st, mh, lead_full, c0, row0, b);
                if (SPLIT) tma_load_3d_cg2(st + A_BYTES, ml, lead_full, c0, row0, b);
              }
              if (p.l2_hints & 4) {
                tma_load_3d_cg2_hint(st + W_OFF, &maps.w1_hi, lead_full, ks * BK, w_row, layer, keep);
                if (SPLIT) tma_load_3d_cg2_hint(st + W_OFF + W_BYTES, &maps.w1_lo, lead_full, ks * BK, w_row, layer, keep);
              } else {
                tma_load_3d_cg2(st + W_OFF, &maps.w1_hi, lead_full, ks * BK, w_row, layer);
                if (SPLIT) tma_load_3d_cg2(st + W_OFF + W_BYTES, &maps.w1_lo, lead_full, ks * BK, w_row, layer);
              }
              advance();
            }
          }
          if (q > 0 && res) {
            const int u = (q - 1) % n_units;
            for (int s = 0; s < kp_steps; ++s) {
              uint8_t* st = acquire((uint32_t)(NS * w2_rows * ROWB));
              const uint32_t lead_full = mapa_u32(smem_u32(&full[stage]), 0);
              const int k0 = u * chpu + s * BK;
              tma_load_3d_cg2(st + W_OFF, &maps.w2_hi, lead_full, k0, rank * w2_rows, layer);
              if (SPLIT) tma_load_3d_cg2(st + W_OFF + W_BYTES, &maps.w2_lo, lead_full, k0, rank * w2_rows, layer);
              advance();
            }
          }
        }
      } else if (warp == 1 && lane == 0) {
        // ===================================================== UMMA issuer (the leader issues for the pair)
        if (rank == 0) {
          const uint32_t idesc1 = make_idesc_bf16(2 * TC_BM, n_cols), idesc2 = make_idesc_bf16(2 * TC_BM, C);
          const uint32_t acts_a = smem_u32(acts_s);
          Tick tk;
          tk.start();
          for (int q = 0; q <= Q; ++q) {
            if (q < Q) {
              // first GEMM of unit q: with a residual part always region 0 (drained -- into registers -- by the
              // epilogue of unit q-1 while the residual part of unit q-2 ran); without, the two regions alternate
              const int r = res ? 0 : (q & 1);
              tk.lap(issue);
              mbar_wait(&tmem_empty[r], (use[r]++ & 1) ^ 1);
              tk.lap(w_tmem0);
              tc_fence_after();
              const uint32_t d1 = tmem_base + r * TC_NHALF;
              for (int ks = 0; ks < p.k1_steps; ++ks) {
                tk.lap(issue);
                mbar_wait(&full[stage], phase);
                tk.lap(w_full);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + stage * STAGE_BYTES);
#pragma unroll
                for (int kk = 0; kk < BK / 16; ++kk) {
                  const uint32_t koff = kk * 32;
                  const uint64_t a_h = make_smem_desc(st + koff, ROWB), w_h = make_smem_desc(st + W_OFF + koff, ROWB);
                  umma_bf16_cg2(d1, a_h, w_h, idesc1, (ks > 0 || kk > 0) ? 1u : 0u);
                  if (SPLIT) {
                    const uint64_t a_l = make_smem_desc(st + A_BYTES + koff, ROWB);
                    const uint64_t w_l = make_smem_desc(st + W_OFF + W_BYTES + koff, ROWB);
                    umma_bf16_cg2(d1, a_l, w_h, idesc1, 1u);
                    umma_bf16_cg2(d1, a_h, w_l, idesc1, 1u);
                  }
                }
                umma_commit_cg2(&empty[stage], (uint16_t)0x3);
                advance();
              }
              umma_commit_cg2(&tmem_full[r], (uint16_t)0x3);
            }
            if (q > 0 && res) {
              // residual part of unit q-1 into region 1: its acts were written while unit q was being multiplied
              const int u = (q - 1) % n_units;
              const uint32_t d2 = tmem_base + TC_NHALF;
              tk.lap(issue);
              mbar_wait_cluster(acts_ready, acts_n++ & 1);
              tk.lap(w_acts);
              if (u == 0) mbar_wait(&tmem_empty[1], (use[1]++ & 1) ^ 1);     // EG of the previous tile
              tk.lap(w_tmem1);
              tc_fence_after();
              for (int s = 0; s < kp_steps; ++s) {
                tk.lap(issue);
                mbar_wait(&full[stage], phase);
                tk.lap(w_full);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + stage * STAGE_BYTES);
                const uint32_t as = acts_a + s * NS * A_BYTES;
#pragma unroll
                for (int kk = 0; kk < BK / 16; ++kk) {
                  const uint32_t koff = kk * 32;
                  const uint64_t a_h = make_smem_desc(as + koff, ROWB), w_h = make_smem_desc(st + W_OFF + koff, ROWB);
                  umma_bf16_cg2(d2, a_h, w_h, idesc2, (u > 0 || s > 0 || kk > 0) ? 1u : 0u);
                  if (SPLIT) {
                    const uint64_t a_l = make_smem_desc(as + A_BYTES + koff, ROWB);
                    const uint64_t w_l = make_smem_desc(st + W_OFF + W_BYTES + koff, ROWB);
                    umma_bf16_cg2(d2, a_l, w_h, idesc2, 1u);
                    umma_bf16_cg2(d2, a_h, w_l, idesc2, 1u);
                  }
                }
                umma_commit_cg2(&empty[stage], (uint16_t)0x3);
                advance();
              }
              umma_commit_cg2(acts_free, (uint16_t)0x3);                 // the acts tile may be overwritten
              if (u == n_units - 1) umma_commit_cg2(&tmem_full[1], (uint16_t)0x3);
            }
          }
        }
      } else if (warp == 2 && lane == 0 && res) {
        // ===================================================== x producer: the tile's old residual stream for EG
        const CUtensorMap* ym_hi = in_a ? &maps.ya_hi : &maps.yb_hi;
        const CUtensorMap* ym_lo = in_a ? &maps.ya_lo : &maps.yb_lo;
        const int need_writes = writes_before(li);
        if (!FLOW && p.pdl) grid_dep_wait();
        for (int j = 0; j < my_tiles; ++j) {
          const int tile = tile_first + j * (int)gridDim.x + rank;
          const int b = tile / p.tiles_per_batch, t0 = (tile % p.tiles_per_batch) * TC_BM;
          if (FLOW && need_writes > 0) flow_wait_cleared(flow_cleared, (uint32_t)(li * my_tiles + j));
          for (int cc = 0; cc < xs_chunks; ++cc)
            for (int h = 0; h < 2; ++h) {
              uint8_t* dst = xs_s + h * 2 * FU_XS_ARRAY;
              const int c0 = h * (C / 2) + cc * FU_XS_COLS;
              mbar_wait(&xs_empty[h], xph[h] ^ 1);
              mbar_arrive_expect_tx(&xs_full[h], NS * FU_XS_ARRAY);
              if (p.l2_hints & 2) {    // the last read of the old x in this layer
                const uint64_t once = l2_policy_evict_first();
                tma_load_3d_hint(dst, ym_hi, &xs_full[h], c0, t0, b, once);
                if (SPLIT) tma_load_3d_hint(dst + FU_XS_ARRAY, ym_lo, &xs_full[h], c0, t0, b, once);
              } else {
                tma_load_3d(dst, ym_hi, &xs_full[h], c0, t0, b);
                if (SPLIT) tma_load_3d(dst + FU_XS_ARRAY, ym_lo, &xs_full[h], c0, t0, b);
              }
              xph[h] ^= 1;
            }
        }
      } else if (FLOW && warp == 3 && lane == 0) {
        // ===================================================== watcher: clears the (layer, tile) items in order
        const int need_writes = writes_before(li);
        for (int j = 0; j < my_tiles; ++j) {
          flow_wait_tiles(p.grid_bar, tile_first + j * (int)gridDim.x + rank, p.n_tiles, need_writes, true);
          asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(flow_cleared)),
                       "r"((uint32_t)(li * my_tiles + j + 1))
                       : "memory");
        }
      }
      __syncwarp();
    }
    if (PROF && p.prof) {
      long long* pr = p.prof + blockIdx.x * 16;
      if (warp == 0 && lane == 0) pr[0] = prod_wait;
      if (warp == 1 && lane == 0 && rank == 0) {
        pr[1] = w_tmem0; pr[2] = w_full; pr[3] = w_acts; pr[4] = w_tmem1; pr[5] = clock64() - k_start;
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FU_REGS_HIGH));
    if (!FLOW && p.pdl) grid_dep_wait();               // out8 and the stream are the previous kernel's
    // ===================================================== epilogue: 8 warps = TMEM lane quarter x column half
    const int qd = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = qd * 32 + lane;
    const int ncol_half = n_cols / 2;                  // accumulator columns of this thread in the first GEMM
    const int c_begin = half * ncol_half;
    const uint32_t lane_base = tmem_base + ((uint32_t)(qd * 32) << 16);
    const uint32_t lead_empty0 = mapa_u32(smem_u32(&tmem_empty[0]), 0);
    const uint32_t lead_empty1 = mapa_u32(smem_u32(&tmem_empty[1]), 0);
    const uint32_t lead_ready = mapa_u32(smem_u32(acts_ready), 0);
    const uint32_t acts_a = smem_u32(acts_s);
    const uint32_t xs_a = smem_u32(xs_s) + half * 2 * FU_XS_ARRAY;
    uint8_t* xs_mine = xs_s + half * 2 * FU_XS_ARRAY;
    const bool storer = qd == 0 && lane == 0;          // the thread of this half that issues the TMA stores
    auto half_sync = [&]() {
      if (half == 0) asm volatile("bar.sync 3, 128;" ::: "memory");
      else asm volatile("bar.sync 4, 128;" ::: "memory");
    };
    uint32_t xs_phase = 0, use[2] = {0, 0}, acts_n = 0;
    int pending_tile = -1;             // storer: a tile whose stores are committed but not yet announced (FLOW)
    long long e_wfull0 = 0, e_drain = 0, e_wfree = 0, e_busy = 0, e_wfull1 = 0, eg_busy = 0;
    const long long e_start = PROF ? clock64() : 0;
    Tick tk;
    tk.start();

    if (do_start) {
      // ------------------------------------------------- x = start(audio_0) (glow.py:156) for this CTA's tiles,
      // written to buffer A through the x staging and a TMA store
      const int off = p.n_group - p.n_rem;
      for (int j = 0; j < my_tiles; ++j) {
        const int tile = tile_first + j * (int)gridDim.x + rank;
        const int b = tile / p.tiles_per_batch, t0 = (tile % p.tiles_per_batch) * TC_BM;
        const int t = t0 + row;
        const bool valid = tile < p.n_tiles && t < p.T;
        float a_in[FU_NOUT / 2];
#pragma unroll
        for (int i = 0; i < FU_NOUT / 2; ++i)
          a_in[i] = (valid && i < p.n_half) ? __ldcg(p.audio + ((long long)b * p.T + t) * p.n_group + off + i) : 0.f;
        for (int cc = 0; cc < xs_chunks; ++cc) {
          const int n0 = half * (C / 2) + cc * FU_XS_COLS;
          uint32_t hw[16], lw[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            float2 v = __ldg(reinterpret_cast<const float2*>(p.start_b + n0 + 2 * k));
#pragma unroll
            for (int i = 0; i < FU_NOUT / 2; ++i) {
              if (i < p.n_half) {
                const float2 w = __ldg(reinterpret_cast<const float2*>(p.start_w + (long long)i * C + n0 + 2 * k));
                v.x = fmaf(a_in[i], w.x, v.x);
                v.y = fmaf(a_in[i], w.y, v.y);
              }
            }
            split2(v.x, v.y, hw[k], lw[k]);
          }
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const uint32_t o = swizzle_off<64>((uint32_t)row * 64 + j4 * 16);
            st_shared_v4(xs_a + o, hw[4 * j4], hw[4 * j4 + 1], hw[4 * j4 + 2], hw[4 * j4 + 3]);
            if (SPLIT) st_shared_v4(xs_a + FU_XS_ARRAY + o, lw[4 * j4], lw[4 * j4 + 1], lw[4 * j4 + 2], lw[4 * j4 + 3]);
          }
          fence_proxy_async_smem();
          half_sync();
          if (storer) {
            tma_store_3d(&maps.ya_hi, xs_mine, n0, t0, b);
            if (SPLIT) tma_store_3d(&maps.ya_lo, xs_mine + FU_XS_ARRAY, n0, t0, b);
            tma_store_commit();
            tma_store_wait_read();
          }
          half_sync();                         // the staging entry may be rewritten
        }
        if (storer && tile < p.n_tiles) {      // this half of the tile's first version is out
          tma_store_wait_all();
          flow_signal(p.grid_bar, tile);
        }
      }
    }

    for (int li = 0; li < layer_count; ++li) {
      const int layer = p.layer_first + li;
      const bool res = has_res(layer);
      const bool in_a = (layer & 1) == 0;
      const CUtensorMap* yo_hi = in_a ? &maps.yb_hi : &maps.ya_hi;       // the buffer this layer writes
      const CUtensorMap* yo_lo = in_a ? &maps.yb_lo : &maps.ya_lo;
      const float* bias1 = p.bias1[layer];
      const float* wcl = p.wc[layer];
      const float* res_b = p.res_b[layer];
      const bool acc_out8 = layer_count > 1 ? layer > 0 : p.accumulate_out8 != 0;
      for (int q = 0; q <= Q; ++q) {
        if (q < Q) {
          // ------------------------------------------------- E(q): drain, gate, out8, acts -> shared memory
          const int tile_j = q / n_units, u = q % n_units;
          const int tile = tile_first + tile_j * (int)gridDim.x + rank;
          const int b = tile / p.tiles_per_batch;
          const int t = (tile % p.tiles_per_batch) * TC_BM + row;
          const bool valid = tile < p.n_tiles && t < p.T;
          const long long col = (long long)b * p.T + t;
          const int r = res ? 0 : (q & 1);
          tk.lap(e_busy);
          mbar_wait(&tmem_full[r], use[r]++ & 1);
          tk.lap(e_wfull0);
          tc_fence_after();
          uint32_t acc[4][32];
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4)
            if (c4 * 32 < ncol_half) tmem_ld32(lane_base + r * TC_NHALF + c_begin + c4 * 32, acc[c4]);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(r == 0 ? lead_empty0 : lead_empty1);   // the region is free again
          tk.lap(e_drain);
          float acc8[FU_NOUT];
#pragma unroll
          for (int o = 0; o < FU_NOUT; ++o) acc8[o] = 0.f;
          bool waited = !res;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            if (c4 * 32 < ncol_half) {
              const int n0 = u * TC_NHALF + c_begin + c4 * 32;     // first output column of this chunk
              const int ch_u = (c_begin + c4 * 32) >> 1;           // first channel inside the unit (16 per chunk)
              const int ch0 = n0 >> 1;                             // global channel
              float g[16];
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const float4 bv = __ldg(reinterpret_cast<const float4*>(bias1 + n0 + 2 * j));
                g[j] = gate_act(__uint_as_float(acc[c4][2 * j]) + bv.x, __uint_as_float(acc[c4][2 * j + 1]) + bv.y);
                g[j + 1] = gate_act(__uint_as_float(acc[c4][2 * j + 2]) + bv.z, __uint_as_float(acc[c4][2 * j + 3]) + bv.w);
              }
              // collapsed skip path: out8 += Wc[:, ch0:ch0+16] g   (fp32, exact gate outputs; Wc: warp-uniform
              // addresses, L1-resident)
#pragma unroll
              for (int o = 0; o < FU_NOUT; ++o) {
                const float4* wrow = reinterpret_cast<const float4*>(wcl + o * C + ch0);
                float a = acc8[o];
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                  const float4 w4 = __ldg(wrow + j4);
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
              if (res) {
                if (!waited) {      // the previous residual part has read the acts tile (completes right after unit q)
                  tk.lap(e_busy);
                  if (acts_n > 0) mbar_wait(acts_free, (acts_n - 1) & 1);
                  ++acts_n;
                  tk.lap(e_wfree);
                  waited = true;
                }
                // rows outside the utterance hold zeros (their x_new rows are clipped by the TMA store)
                const uint32_t s = (uint32_t)(ch_u / BK);
                const uint32_t off = (uint32_t)row * ROWB + (uint32_t)(ch_u % BK) * 2;
                const uint32_t base = acts_a + s * NS * A_BYTES;
                const uint32_t o0 = swizzle_off<ROWB>(off), o1 = swizzle_off<ROWB>(off + 16);
                st_shared_v4(base + o0, hi[0], hi[1], hi[2], hi[3]);
                st_shared_v4(base + o1, hi[4], hi[5], hi[6], hi[7]);
                if (SPLIT) {
                  st_shared_v4(base + A_BYTES + o0, lo[0], lo[1], lo[2], lo[3]);
                  st_shared_v4(base + A_BYTES + o1, lo[4], lo[5], lo[6], lo[7]);
                }
              }
              if (p.acts_hi != nullptr && valid) {
                const long long goff = col * C + ch0;
                uint4* dh = reinterpret_cast<uint4*>(p.acts_hi + goff);
                dh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                dh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                if (SPLIT) {
                  uint4* dl = reinterpret_cast<uint4*>(p.acts_lo + goff);
                  dl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                  dl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                }
              }
            }
          }
          if (res) {
            fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lead_ready);
          }
          // out8 (+)= partial sums of the two column halves, in a fixed order (deterministic rounding): the lower
          // half first (it starts or continues the layer sum), then the upper half on top.  L2 is the meeting point
          // (ld.global.cg), the named barriers order the two read-modify-writes and the next unit's.
          {
            float4* o8 = reinterpret_cast<float4*>(p.out8 + col * FU_NOUT);
            float4 o0 = make_float4(acc8[0], acc8[1], acc8[2], acc8[3]);
            float4 o1 = make_float4(acc8[4], acc8[5], acc8[6], acc8[7]);
            if (half == 0 && valid) {
              if (acc_out8 || u > 0) {
                const float4 p0 = __ldcg(o8), p1 = __ldcg(o8 + 1);
                o0.x += p0.x; o0.y += p0.y; o0.z += p0.z; o0.w += p0.w;
                o1.x += p1.x; o1.y += p1.y; o1.z += p1.z; o1.w += p1.w;
              }
              o8[0] = o0;
              o8[1] = o1;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(FU_EPI_THREADS) : "memory");
            if (half == 1 && valid) {
              const float4 p0 = __ldcg(o8), p1 = __ldcg(o8 + 1);
              o0.x += p0.x; o0.y += p0.y; o0.z += p0.z; o0.w += p0.w;
              o1.x += p1.x; o1.y += p1.y; o1.z += p1.z; o1.w += p1.w;
              o8[0] = o0;
              o8[1] = o1;
            }
            asm volatile("bar.sync 2, %0;" ::"n"(FU_EPI_THREADS) : "memory");
          }
        }
        if (q > 0 && res && (q - 1) % n_units == n_units - 1) {
          // ------------------------------------------------- EG: x_new = res + b_res + x (glow.py:166), through the
          // x staging: the old x arrives by TMA, is updated in place (own row) and leaves by TMA store
          const int tile_j = (q - 1) / n_units;
          const int tile = tile_first + tile_j * (int)gridDim.x + rank;
          const int b = tile / p.tiles_per_batch, t0 = (tile % p.tiles_per_batch) * TC_BM;
          tk.lap(e_busy);
          if (FLOW && storer && pending_tile >= 0) {
            // announce the previous tile now: its stores were committed a whole tile of UMMAs ago, the wait is free
            tma_store_wait_all();
            flow_signal(p.grid_bar, pending_tile);
            pending_tile = -1;
          }
          mbar_wait(&tmem_full[1], use[1]++ & 1);
          tk.lap(e_wfull1);
          tc_fence_after();
          for (int cc = 0; cc < xs_chunks; ++cc) {
            const int n0 = half * (C / 2) + cc * FU_XS_COLS;
            uint32_t rr[32];
            tmem_ld32(lane_base + TC_NHALF + n0, rr);
            mbar_wait(&xs_full[half], xs_phase);
            xs_phase ^= 1;
            tmem_ld_wait();
            uint32_t hw[16], lw[16];
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const uint32_t o = swizzle_off<64>((uint32_t)row * 64 + j4 * 16);
              ld_shared_v4(xs_a + o, hw[4 * j4], hw[4 * j4 + 1], hw[4 * j4 + 2], hw[4 * j4 + 3]);
              if (SPLIT) ld_shared_v4(xs_a + FU_XS_ARRAY + o, lw[4 * j4], lw[4 * j4 + 1], lw[4 * j4 + 2], lw[4 * j4 + 3]);
              else lw[4 * j4] = lw[4 * j4 + 1] = lw[4 * j4 + 2] = lw[4 * j4 + 3] = 0u;
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const float2 bb = __ldg(reinterpret_cast<const float2*>(res_b + n0 + 2 * k));
              const float2 xhf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[k]));
              const float2 xlf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&lw[k]));
              const float v0 = (__uint_as_float(rr[2 * k]) + bb.x) + (xhf.x + xlf.x);
              const float v1 = (__uint_as_float(rr[2 * k + 1]) + bb.y) + (xhf.y + xlf.y);
              split2(v0, v1, hw[k], lw[k]);
            }
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const uint32_t o = swizzle_off<64>((uint32_t)row * 64 + j4 * 16);
              st_shared_v4(xs_a + o, hw[4 * j4], hw[4 * j4 + 1], hw[4 * j4 + 2], hw[4 * j4 + 3]);
              if (SPLIT) st_shared_v4(xs_a + FU_XS_ARRAY + o, lw[4 * j4], lw[4 * j4 + 1], lw[4 * j4 + 2], lw[4 * j4 + 3]);
            }
            fence_proxy_async_smem();
            half_sync();
            if (storer) {
              // rows outside the utterance (and a tile past the end) are clipped by the tensor map
              tma_store_3d(yo_hi, xs_mine, n0, t0, b);
              if (SPLIT) tma_store_3d(yo_lo, xs_mine + FU_XS_ARRAY, n0, t0, b);
              tma_store_commit();
              tma_store_wait_read();               // the staging entry may be refilled
              mbar_arrive(&xs_empty[half]);
            }
          }
          if (FLOW && storer && tile < p.n_tiles) pending_tile = tile;
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(lead_empty1);
          tk.lap(eg_busy);
        }
      }
      if (storer) {
        tma_store_wait_all();
        if (FLOW && pending_tile >= 0) {       // the last tile of the layer
          flow_signal(p.grid_bar, pending_tile);
          pending_tile = -1;
        }
      }
    }

    if (do_end && half == 0) {
      // ------------------------------------------------- out = end(skip sum) = out8 + bias8 (glow.py:175); b, s =
      // halves (278-279); a1 <- (a1 - b) / exp(s) (280); z <- W^-1 [a0; a1] (283, 96): one thread per column of this
      // CTA's tiles, in place on audio (out8 rows were completed by this CTA's own epilogue)
      const int off = p.n_group - p.n_rem;
      for (int j = 0; j < my_tiles; ++j) {
        const int tile = tile_first + j * (int)gridDim.x + rank;
        const int b = tile / p.tiles_per_batch;
        const int t = (tile % p.tiles_per_batch) * TC_BM + row;
        if (tile >= p.n_tiles || t >= p.T) continue;
        const long long col = (long long)b * p.T + t;
        const float4 qa = __ldcg(reinterpret_cast<const float4*>(p.out8 + col * FU_NOUT));
        const float4 qb = __ldcg(reinterpret_cast<const float4*>(p.out8 + col * FU_NOUT) + 1);
        const float o[FU_NOUT] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
        float* acol = p.audio + col * p.n_group + off;
        float y[FU_NOUT];
#pragma unroll
        for (int i = 0; i < FU_NOUT; ++i) {
          y[i] = 0.f;
          if (i < p.n_rem) {
            const float av = __ldcg(acol + i);
            if (i < p.n_half) {
              y[i] = av;
            } else {
              float bshift = 0.f, sc = 0.f;
#pragma unroll
              for (int k = 0; k < FU_NOUT; ++k) {     // static indexing keeps o[] in registers
                if (k == i - p.n_half) bshift = o[k] + __ldg(p.out_bias + k);
                if (k == i) sc = o[k] + __ldg(p.out_bias + k);
              }
              y[i] = (av - bshift) / expf(sc);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < FU_NOUT; ++i) {
          if (i < p.n_rem) {
            float z = 0.f;
#pragma unroll
            for (int k = 0; k < FU_NOUT; ++k)
              if (k < p.n_rem) z = fmaf(__ldg(p.w_inv + i * p.n_rem + k), y[k], z);
            acol[i] = z;
          }
        }
      }
    }
    if (PROF && p.prof && warp == 4 && lane == 0) {
      long long* pr = p.prof + blockIdx.x * 16;
      pr[6] = e_wfull0; pr[7] = e_drain; pr[8] = e_wfree; pr[9] = e_busy; pr[10] = e_wfull1; pr[11] = eg_busy;
      pr[12] = clock64() - e_start;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();     // nobody leaves while the peer may still signal / read this CTA
  tc_fence_after();
  if (warp == 1) tmem_dealloc_cg2(tmem_base, 512);
}

template <int BK, int NS, bool FLOW, bool PROF>
int launch_fused(const FusedMaps& maps, const FusedParams& p, cudaStream_t st) {
  static bool attr_set_on[FAC_MAX_DEVICES] = {};
  bool& attr_set = attr_set_on[current_device_slot()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wn_flow_fused_kernel<BK, NS, FLOW, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, FU_SMEM);
    if (e != cudaSuccess) {
      set_error("wn_flow_fused: cannot reserve %d bytes of shared memory: %s", FU_SMEM, cudaGetErrorString(e));
      return 2;
    }
    attr_set = true;
  }
  const int pairs = ceil_div(p.n_tiles, 2), max_pairs = sm_count() / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * (pairs < max_pairs ? pairs : max_pairs)));     // <= 1 CTA per SM: co-resident
  cfg.blockDim = dim3(FU_THREADS);
  cfg.dynamicSmemBytes = FU_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (FLOW) {
    attr[0].id = cudaLaunchAttributeCooperative;    // CTAs wait for each other's tiles: every CTA must be resident
    attr[0].val.cooperative = 1;
  } else {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
  }
  cfg.attrs = attr;
  cfg.numAttrs = (FLOW || p.pdl) ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, wn_flow_fused_kernel<BK, NS, FLOW, PROF>, maps, p);
  count_launch();
  if (e != cudaSuccess) {
    set_error("wn_flow_fused_kernel: launch failed: %s", cudaGetErrorString(e));
    return 2;
  }
  return check_launch("wn_flow_fused_kernel");
}

// [layers][N][K] 16-bit weights of one flow with a uniform layer stride: box = bk k x (min(256, N) / 2) rows.
int make_weight_map3(CUtensorMap* m, const void* ptr, int N, int K, int layers, long long layer_stride_bytes, int bk) {
  EncodeTiledFn fn = encode_fn();
  FAC_REQUIRE(fn != nullptr, "tensor-core path: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)N, (cuuint64_t)layers};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)layer_stride_bytes};
  cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)((N < TC_NHALF ? N : TC_NHALF) / 2), 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, tc_swizzle(bk), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FAC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights %dx%dx%d) failed: %d", layers, N, K, (int)r);
  return 0;
}

}  // namespace

bool wn_fused_supported(int C, int n_cond, int bk) {
  const int n_cols = 2 * C < TC_NHALF ? 2 * C : TC_NHALF;
  return (bk == 32 || bk == 64) && C % bk == 0 && n_cond % bk == 0 && C <= FU_CMAX && (2 * C) % n_cols == 0 &&
         (n_cols / 2) % bk == 0 && (n_cols / 2) % 32 == 0 && (C / 2) % FU_XS_COLS == 0 && C % 16 == 0;
}

// The tensor-core weights of a flow are usable by the fused kernel when every layer's matrices sit at a uniform
// stride (packing.py lays them out that way): one 3-D tensor map then serves all layers.
bool wn_fused_weights_ok(const fac_wg_model* m, const fac_wg_tc_flow& wf) {
  const int L = m->n_layers;
  if (!wf.w1_hi[0] || !wf.w1_lo[0]) return false;
  if (L == 1) return true;
  if (!wf.w2r_hi[0] || !wf.w2r_lo[0]) return false;
  const long long stride = (const char*)wf.w1_hi[1] - (const char*)wf.w1_hi[0];
  if (stride <= 0 || stride % 16) return false;
  for (int i = 0; i < L; ++i) {
    if ((const char*)wf.w1_hi[i] != (const char*)wf.w1_hi[0] + i * stride) return false;
    if ((const char*)wf.w1_lo[i] != (const char*)wf.w1_lo[0] + i * stride) return false;
    if (i < L - 1) {
      if ((const char*)wf.w2r_hi[i] != (const char*)wf.w2r_hi[0] + i * stride) return false;
      if ((const char*)wf.w2r_lo[i] != (const char*)wf.w2r_lo[0] + i * stride) return false;
      if (!wf.res_b[i]) return false;
    }
    if (!wf.wc[i]) return false;
  }
  return true;
}

// Layers [layer_first, layer_first + layer_count) of flow `flow`, optionally with start (before) and end +
// coupling + invertible 1x1 (after), as ONE launch.  The residual stream of layer l is read from ws->x when l is
// even and from ws->x2 when l is odd, and written to the other pair; start writes ws->x.
int wn_flow_fused(const fac_wg_model* m, const fac_wg_tc_weights* w, int flow, const fac_wg_tc_workspace* ws,
                  float* audio, int B, int T, int nsplit, int layer_first, int layer_count, int do_start, int do_end,
                  int bk, int prefetch_steps, int l2_hints, long long* prof, cudaStream_t st) {
  const fac_wg_flow& f = m->flows[flow];
  const fac_wg_tc_flow& wf = w->flows[flow];
  const int C = m->n_channels, n_cond = m->n_mel * m->n_group, taps = m->kernel_size, L = m->n_layers;
  FAC_REQUIRE(wn_fused_supported(C, n_cond, bk), "wn_flow_fused: unsupported geometry C=%d n_cond=%d bk=%d", C, n_cond, bk);
  FAC_REQUIRE(nsplit == 1 || nsplit == 2, "wn_flow_fused: nsplit must be 1 or 2");
  FAC_REQUIRE(wn_fused_weights_ok(m, wf), "wn_flow_fused: the flow's tensor-core weights are not uniformly strided");
  FAC_REQUIRE(layer_first >= 0 && layer_count >= 1 && layer_first + layer_count <= L, "wn_flow_fused: layer range");
  FAC_REQUIRE(ws->x_hi && ws->x2_hi && ws->spect_hi && ws->out8, "wn_flow_fused: workspace incomplete");
  FAC_REQUIRE(nsplit == 1 || (ws->x_lo && ws->x2_lo && ws->spect_lo), "wn_flow_fused: lo buffers missing");
  // plain bf16: the lo maps are never dereferenced; they alias the hi ones
  const void* x_lo = nsplit == 2 ? ws->x_lo : ws->x_hi;
  const void* x2_lo = nsplit == 2 ? ws->x2_lo : ws->x2_hi;
  const void* spect_lo = nsplit == 2 ? ws->spect_lo : ws->spect_hi;
  const bool phases = do_start || (layer_count > 1);
  FAC_REQUIRE(!phases || ws->flow_sync, "wn_flow_fused: a multi-phase launch needs ws->flow_sync");
  FAC_REQUIRE(!(do_start || do_end) || audio, "wn_flow_fused: audio missing");
  FusedMaps maps;
  const int K1 = taps * C + n_cond;
  const long long stride = L > 1 ? (const char*)wf.w1_hi[1] - (const char*)wf.w1_hi[0] : (long long)2 * C * K1 * 2;
  if (int rc = make_act_map(&maps.xa_hi, ws->x_hi, B, T, C, bk)) return rc;
  if (int rc = make_act_map(&maps.xa_lo, x_lo, B, T, C, bk)) return rc;
  if (int rc = make_act_map(&maps.xb_hi, ws->x2_hi, B, T, C, bk)) return rc;
  if (int rc = make_act_map(&maps.xb_lo, x2_lo, B, T, C, bk)) return rc;
  if (int rc = make_act_map(&maps.s_hi, ws->spect_hi, B, T, n_cond, bk)) return rc;
  if (int rc = make_act_map(&maps.s_lo, spect_lo, B, T, n_cond, bk)) return rc;
  if (int rc = make_weight_map3(&maps.w1_hi, wf.w1_hi[0], 2 * C, K1, L, stride, bk)) return rc;
  if (int rc = make_weight_map3(&maps.w1_lo, wf.w1_lo[0], 2 * C, K1, L, stride, bk)) return rc;
  if (L > 1) {
    if (int rc = make_weight_map3(&maps.w2_hi, wf.w2r_hi[0], C, C, L - 1, stride, bk)) return rc;
    if (int rc = make_weight_map3(&maps.w2_lo, wf.w2r_lo[0], C, C, L - 1, stride, bk)) return rc;
  } else {
    maps.w2_hi = maps.w1_hi;
    maps.w2_lo = maps.w1_lo;
  }
  if (int rc = make_act_map(&maps.ya_hi, ws->x_hi, B, T, C, FU_XS_COLS)) return rc;
  if (int rc = make_act_map(&maps.ya_lo, x_lo, B, T, C, FU_XS_COLS)) return rc;
  if (int rc = make_act_map(&maps.yb_hi, ws->x2_hi, B, T, C, FU_XS_COLS)) return rc;
  if (int rc = make_act_map(&maps.yb_lo, x2_lo, B, T, C, FU_XS_COLS)) return rc;
  FusedParams p{};
  p.T = T;
  p.B = B;
  p.tiles_per_batch = ceil_div(T, TC_BM);
  p.n_tiles = B * p.tiles_per_batch;
  p.C = C;
  p.n_cond = n_cond;
  p.taps = taps;
  p.k1_steps = K1 / bk;
  p.layer_first = layer_first;
  p.layer_count = layer_count;
  p.n_layers = L;
  p.do_start = do_start;
  p.do_end = do_end;
  for (int i = 0; i < L; ++i) {
    p.bias1[i] = f.in_cond_b[i];
    p.res_b[i] = i < L - 1 ? wf.res_b[i] : nullptr;
    p.wc[i] = wf.wc[i];
  }
  p.out8 = ws->out8;
  p.accumulate_out8 = layer_first > 0;
  p.pdl = tc_pdl_enabled() ? 1 : 0;
  p.start_w = f.start_w;
  p.start_b = f.start_b;
  p.out_bias = wf.out_bias;
  p.w_inv = f.w_inv;
  p.audio = audio;
  p.n_group = m->n_group;
  p.n_rem = f.n_rem;
  p.n_half = f.n_half;
  p.acts_hi = layer_count == 1 ? reinterpret_cast<__nv_bfloat16*>(ws->acts_hi) : nullptr;
  p.acts_lo = layer_count == 1 ? reinterpret_cast<__nv_bfloat16*>(ws->acts_lo) : nullptr;
  if (p.acts_hi == nullptr || (nsplit == 2 && p.acts_lo == nullptr)) p.acts_hi = p.acts_lo = nullptr;
  p.grid_bar = reinterpret_cast<unsigned int*>(ws->flow_sync);
  p.prefetch_steps = prefetch_steps < p.k1_steps ? prefetch_steps : p.k1_steps - 1;
  p.l2_hints = l2_hints;
  p.prof = prof;
  if (phases) {
    // flags[t]: finished stores of tile t's residual stream in this launch
    cudaError_t e = cudaMemsetAsync(ws->flow_sync, 0, (size_t)p.n_tiles * sizeof(unsigned int), st);
    if (e != cudaSuccess) {
      set_error("wn_flow_fused: cannot reset the tile flags: %s", cudaGetErrorString(e));
      return 2;
    }
  }
  if (nsplit == 1) {         // plain bf16 (K = 32 only: 8 ring stages)
    FAC_REQUIRE(bk == 32 && prof == nullptr, "wn_flow_fused: the bf16 form is compiled for K = 32 without cycle counters");
    return phases ? launch_fused<32, 1, true, false>(maps, p, st) : launch_fused<32, 1, false, false>(maps, p, st);
  }
  if (prof != nullptr) {     // the instrumented instantiations (K = 32 only)
    FAC_REQUIRE(bk == 32, "wn_flow_fused: cycle counters are compiled for the K = 32 kernel only");
    return phases ? launch_fused<32, 2, true, true>(maps, p, st) : launch_fused<32, 2, false, true>(maps, p, st);
  }
  if (phases) return bk == 64 ? launch_fused<64, 2, true, false>(maps, p, st) : launch_fused<32, 2, true, false>(maps, p, st);
  return bk == 64 ? launch_fused<64, 2, false, false>(maps, p, st) : launch_fused<32, 2, false, false>(maps, p, st);
}

}  // namespace fac
