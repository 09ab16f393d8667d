// WaveGlow reverse flow, exact-fp32 path (reference src/waveglow/glow.py:252-293).
//
// Data layout in HBM (all fp32, channels-last):
//   spect (B, T_g, n_mel*n_group)   squeezed upsampled conditioning, channel m*n_group+j
//   x, acts, skip (B, T_g, C)       WN residual stream, gated activations, skip sum
//   audio (B, T_g, n_group)         the flow state; a flow with n_rem live channels uses the
//                                   LAST n_rem slots of each column, so "prepending" the early
//                                   noise (glow.py:285-290) is just widening the slot window and
//                                   the final buffer is the waveform (sample = n_group*t + slot).
#include "fac_common.cuh"

namespace fac {

int launch_conv_gemm_f32(const fac_conv_src* srcs, int n_srcs, const float* w_packed, const float* bias, int B,
                         int T_out, int N, const fac_conv_epilogue* epi, int n_phases, long long w_phase_stride,
                         long long out_phase_stride, cudaStream_t stream);

namespace {

// x[col, c] = start_b[c] + sum_j start_w[j][c] * audio_0[col, j]       (glow.py:156)
__global__ void wn_start_kernel(const float* __restrict__ audio, const float* __restrict__ w,
                                const float* __restrict__ bias, float* __restrict__ x, long long n_cols, int C,
                                int n_group, int off, int n_half) {
  const int c4 = C >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_cols * c4) return;
  const long long col = idx / c4;
  const int c = (int)(idx % c4) * 4;
  float4 o = __ldg(reinterpret_cast<const float4*>(bias + c));
  for (int j = 0; j < n_half; ++j) {
    const float a = __ldg(audio + col * n_group + off + j);
    const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (long long)j * C + c));
    o.x = fmaf(a, wv.x, o.x);
    o.y = fmaf(a, wv.y, o.y);
    o.z = fmaf(a, wv.z, o.z);
    o.w = fmaf(a, wv.w, o.w);
  }
  *reinterpret_cast<float4*>(x + col * C + c) = o;
}

// One warp per column: out = end(skip) (glow.py:175); b, s = halves (glow.py:278-279);
// a1 <- (a1 - b) / exp(s) (glow.py:280); z <- W^-1 [a0; a1] (glow.py:283, 96).
constexpr int END_MAX_OUT = 8;
__global__ void __launch_bounds__(256) wn_end_coupling_kernel(const float* __restrict__ skip,
                                                              const float* __restrict__ end_w,
                                                              const float* __restrict__ end_b,
                                                              const float* __restrict__ w_inv,
                                                              float* __restrict__ audio, long long n_cols, int C,
                                                              int n_group, int n_rem, int n_half) {
  extern __shared__ float sw[];  // end_w [2*n_half][C]
  const int n_out = 2 * n_half;
  for (int i = threadIdx.x; i < n_out * C; i += blockDim.x) sw[i] = __ldg(end_w + i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const long long warp_global = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const long long n_warps = (long long)gridDim.x * warps_per_block;
  const int off = n_group - n_rem;
  for (long long col = warp_global; col < n_cols; col += n_warps) {
    float part[END_MAX_OUT];
#pragma unroll
    for (int o = 0; o < END_MAX_OUT; ++o) part[o] = 0.f;
    const float* srow = skip + col * C;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(srow + c));
#pragma unroll
      for (int o = 0; o < END_MAX_OUT; ++o) {
        if (o < n_out) {
          const float4 wv = *reinterpret_cast<const float4*>(&sw[o * C + c]);
          part[o] = fmaf(v.x, wv.x, fmaf(v.y, wv.y, fmaf(v.z, wv.z, fmaf(v.w, wv.w, part[o]))));
        }
      }
    }
#pragma unroll
    for (int o = 0; o < END_MAX_OUT; ++o) {
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) part[o] += __shfl_xor_sync(0xffffffffu, part[o], sft);
    }
    // every lane now holds the 2*n_half outputs; lane i produces output channel i
    float y[END_MAX_OUT];
    float* acol = audio + col * n_group + off;
#pragma unroll
    for (int j = 0; j < END_MAX_OUT; ++j) {
      if (j < n_rem) {
        const float a = acol[j];
        if (j < n_half) {
          y[j] = a;
        } else {
          const float bshift = part[j - n_half] + __ldg(end_b + (j - n_half));
          const float s = part[j] + __ldg(end_b + j);
          y[j] = (a - bshift) / expf(s);
        }
      } else {
        y[j] = 0.f;
      }
    }
    __syncwarp();
    if (lane < n_rem) {
      float z = 0.f;
#pragma unroll
      for (int j = 0; j < END_MAX_OUT; ++j)
        if (j < n_rem) z = fmaf(__ldg(w_inv + lane * n_rem + j), y[j], z);
      acol[lane] = z;
    }
  }
}

}  // namespace

int wg_check_model(const fac_wg_model* m) {
  FAC_REQUIRE(m != nullptr, "waveglow: model is NULL");
  FAC_REQUIRE(m->n_flows >= 1 && m->n_flows <= FAC_MAX_FLOWS, "waveglow: n_flows %d out of range", m->n_flows);
  FAC_REQUIRE(m->n_layers >= 1 && m->n_layers <= FAC_MAX_LAYERS, "waveglow: n_layers %d out of range", m->n_layers);
  FAC_REQUIRE(m->n_group >= 2 && m->n_group <= END_MAX_OUT, "waveglow: n_group %d unsupported (max %d)", m->n_group,
              END_MAX_OUT);
  FAC_REQUIRE(m->n_channels % 8 == 0, "waveglow: n_channels %d must be a multiple of 8", m->n_channels);
  FAC_REQUIRE((m->n_mel * m->n_group) % 8 == 0 && m->n_mel % 8 == 0, "waveglow: n_mel %d must be a multiple of 8",
              m->n_mel);
  FAC_REQUIRE(m->hop % m->n_group == 0, "waveglow: hop %d must be a multiple of n_group %d", m->hop, m->n_group);
  return 0;
}

int wg_upsample_squeeze(const fac_wg_model* m, const float* mel_cl, float* spect, int B, int F, cudaStream_t st) {
  if (int rc = wg_check_model(m)) return rc;
  FAC_REQUIRE(mel_cl && spect && B > 0 && F > 0, "upsample: bad arguments");
  const int n_cond = m->n_mel * m->n_group;
  const int phases = m->hop / m->n_group;
  const int n_pad = round_up(n_cond, 128);
  fac_conv_src src{mel_cl, (long long)F * m->n_mel, m->n_mel, 1, m->n_mel, m->upsample_taps, -1, 0, F, 0};
  fac_conv_epilogue epi{};
  epi.kind = FAC_EPI_LINEAR;
  epi.act = FAC_ACT_NONE;
  epi.out = spect;
  epi.out_batch_stride = (long long)F * phases * n_cond;
  epi.out_row_stride = (long long)phases * n_cond;
  return launch_conv_gemm_f32(&src, 1, m->upsample_w, m->upsample_b, B, F, n_cond, &epi, phases,
                              (long long)m->upsample_taps * m->n_mel * n_pad, n_cond, st);
}

int wg_start(const fac_wg_model* m, int flow, const float* audio, float* x, int B, int Tg, cudaStream_t st) {
  if (int rc = wg_check_model(m)) return rc;
  FAC_REQUIRE(flow >= 0 && flow < m->n_flows, "wn_start: flow %d out of range", flow);
  const fac_wg_flow& f = m->flows[flow];
  const long long n_cols = (long long)B * Tg;
  const long long total = n_cols * (m->n_channels / 4);
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  wn_start_kernel<<<(unsigned)blocks, threads, 0, st>>>(audio, f.start_w, f.start_b, x, n_cols, m->n_channels,
                                                        m->n_group, m->n_group - f.n_rem, f.n_half);
  count_launch();
  return check_launch("wn_start_kernel");
}

int wg_layer(const fac_wg_model* m, int flow, int layer, const fac_wg_workspace* ws, int B, int Tg, cudaStream_t st) {
  if (int rc = wg_check_model(m)) return rc;
  FAC_REQUIRE(flow >= 0 && flow < m->n_flows && layer >= 0 && layer < m->n_layers, "wn_layer: index out of range");
  FAC_REQUIRE(ws && ws->spect && ws->x && ws->acts && ws->skip, "wn_layer: workspace incomplete");
  const fac_wg_flow& f = m->flows[flow];
  const int C = m->n_channels, n_cond = m->n_mel * m->n_group;
  const int dil = 1 << layer;
  fac_conv_src srcs[2] = {
      {ws->x, (long long)Tg * C, C, 1, C, m->kernel_size, dil, dil * (m->kernel_size - 1) / 2, Tg, 0},
      {ws->spect, (long long)Tg * n_cond, n_cond, 1, n_cond, 1, 0, 0, Tg, 0},
  };
  fac_conv_epilogue gate{};
  gate.kind = FAC_EPI_GATE;
  gate.out = ws->acts;
  gate.out_batch_stride = (long long)Tg * C;
  gate.out_row_stride = C;
  if (int rc = launch_conv_gemm_f32(srcs, 2, f.in_cond_w[layer], f.in_cond_b[layer], B, Tg, 2 * C, &gate, 1, 0, 0, st))
    return rc;
  const bool last = layer == m->n_layers - 1;
  fac_conv_src a{ws->acts, (long long)Tg * C, C, 1, C, 1, 0, 0, Tg, 0};
  fac_conv_epilogue rs{};
  rs.kind = FAC_EPI_RES_SKIP;
  rs.out = ws->x;
  rs.out2 = ws->skip;
  rs.out_batch_stride = (long long)Tg * C;
  rs.out_row_stride = C;
  rs.n_split = last ? 0 : C;
  rs.accumulate_out2 = layer > 0;
  return launch_conv_gemm_f32(&a, 1, f.res_skip_w[layer], f.res_skip_b[layer], B, Tg, last ? C : 2 * C, &rs, 1, 0, 0,
                              st);
}

int wg_end(const fac_wg_model* m, int flow, const float* skip, float* audio, int B, int Tg, cudaStream_t st) {
  if (int rc = wg_check_model(m)) return rc;
  FAC_REQUIRE(flow >= 0 && flow < m->n_flows, "wn_end: flow %d out of range", flow);
  const fac_wg_flow& f = m->flows[flow];
  const long long n_cols = (long long)B * Tg;
  const int threads = 256;
  const int cols_per_block = 64;
  const long long blocks = (n_cols + cols_per_block - 1) / cols_per_block;
  const size_t smem = (size_t)2 * f.n_half * m->n_channels * sizeof(float);
  wn_end_coupling_kernel<<<(unsigned)blocks, threads, smem, st>>>(skip, f.end_w, f.end_b, f.w_inv, audio, n_cols,
                                                                 m->n_channels, m->n_group, f.n_rem, f.n_half);
  count_launch();
  return check_launch("wn_end_coupling_kernel");
}

int wg_infer(const fac_wg_model* m, const float* mel_cl, float* audio, const fac_wg_workspace* ws, int B, int F,
             cudaStream_t st) {
  if (int rc = wg_check_model(m)) return rc;
  FAC_REQUIRE(mel_cl && audio && ws, "waveglow_infer: NULL argument");
  const int Tg = F * (m->hop / m->n_group);
  if (int rc = wg_upsample_squeeze(m, mel_cl, ws->spect, B, F, st)) return rc;
  for (int k = m->n_flows - 1; k >= 0; --k) {
    if (int rc = wg_start(m, k, audio, ws->x, B, Tg, st)) return rc;
    for (int i = 0; i < m->n_layers; ++i)
      if (int rc = wg_layer(m, k, i, ws, B, Tg, st)) return rc;
    if (int rc = wg_end(m, k, ws->skip, audio, B, Tg, st)) return rc;
  }
  return 0;
}

}  // namespace fac
