// extern "C" surface of libfacb200.so (declared in include/fac_b200.h).
#include "fac_common.cuh"
#include <atomic>
#include <cstring>

namespace fac {

static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int launch_conv_gemm_f32(const fac_conv_src*, int, const float*, const float*, int, int, int, const fac_conv_epilogue*,
                         int, long long, long long, cudaStream_t);
int wg_upsample_squeeze(const fac_wg_model*, const float*, float*, int, int, cudaStream_t);
int wg_start(const fac_wg_model*, int, const float*, float*, int, int, cudaStream_t);
int wg_layer(const fac_wg_model*, int, int, const fac_wg_workspace*, int, int, cudaStream_t);
int wg_end(const fac_wg_model*, int, const float*, float*, int, int, cudaStream_t);
int wg_infer(const fac_wg_model*, const float*, float*, const fac_wg_workspace*, int, int, cudaStream_t);

int wg_tc_prepare_spect(const fac_wg_model*, const fac_wg_tc_weights*, const fac_wg_tc_workspace*, const float*, int, int,
                        int, cudaStream_t);
int wg_tc_start(const fac_wg_model*, int, const float*, const fac_wg_tc_workspace*, int, int, int, cudaStream_t);
int wg_tc_layer(const fac_wg_model*, const fac_wg_tc_weights*, int, int, const fac_wg_tc_workspace*, int, int, int,
                cudaStream_t);
int wg_infer_tc(const fac_wg_model*, const fac_wg_tc_weights*, const float*, float*, const fac_wg_tc_workspace*, int,
                int, int, cudaStream_t);
int wg_tc_flow_step(const fac_wg_model*, const fac_wg_tc_weights*, int, float*, const fac_wg_tc_workspace*, int, int, int,
                    cudaStream_t);
int wg_tc_end(const fac_wg_model*, const fac_wg_tc_weights*, int, const float*, float*, int, int, cudaStream_t);
int tc_set_batch_group(int);
void tc_set_prof(long long*);
void taco_set_prof(long long*);
void lstm_set_prof(long long*);
int conv_gemm_tc(const fac_tc_conv*, cudaStream_t);
int tc_transpose_split(const float*, void*, void*, int, int, int, int, int, cudaStream_t);
int tc_pad_split(const float*, void*, void*, long long, int, int, int, cudaStream_t);
int ppg_sparsify(const float*, int*, float*, int*, int, int, int, int, float, cudaStream_t);
int prenet0_sparse(const int*, const float*, const float*, int, const float*, const int*, float*, int, void*, void*, int, int,
                   int, int, int, int, cudaStream_t);
int denoise_spectrum(float*, const float*, float, long long, int, int, cudaStream_t);
int tc_set_cta_group(int);
int tc_set_k_block(int);
int tc_set_fused(int);
int selftest_grid_barrier(unsigned int*, int, cudaStream_t);
int lstm_bidir(const float*, const float*, float*, const int*, int, int, int, cudaStream_t);
int taco_decoder_run(const fac_taco_decoder_weights*, const float*, const float*, const int*, const unsigned char*,
                     const fac_taco_decoder_state*, float*, float*, float*, int, int, int, int, float, cudaStream_t);

}  // namespace fac

extern "C" {

int fac_version(void) { return 200; }
const char* fac_last_error(void) { return fac::g_error; }
long long fac_launch_count(void) { return fac::g_launches.load(); }
void fac_reset_launch_count(void) { fac::g_launches.store(0); }
void fac_add_launch_count(long long n) { fac::g_launches.fetch_add(n, std::memory_order_relaxed); }

int fac_conv_gemm_f32(const fac_conv_src* srcs, int n_srcs, const float* w_packed, const float* bias, int B, int T_out,
                      int N, const fac_conv_epilogue* epi, int n_phases, long long w_phase_stride,
                      long long out_phase_stride, void* stream) {
  return fac::launch_conv_gemm_f32(srcs, n_srcs, w_packed, bias, B, T_out, N, epi, n_phases, w_phase_stride,
                                   out_phase_stride, (cudaStream_t)stream);
}
int fac_waveglow_upsample_squeeze_f32(const fac_wg_model* m, const float* mel_cl, float* spect, int B, int F,
                                      void* stream) {
  return fac::wg_upsample_squeeze(m, mel_cl, spect, B, F, (cudaStream_t)stream);
}
int fac_wn_start_f32(const fac_wg_model* m, int flow, const float* audio, float* x, int B, int Tg, void* stream) {
  return fac::wg_start(m, flow, audio, x, B, Tg, (cudaStream_t)stream);
}
int fac_wn_layer_f32(const fac_wg_model* m, int flow, int layer, const fac_wg_workspace* ws, int B, int Tg,
                     void* stream) {
  return fac::wg_layer(m, flow, layer, ws, B, Tg, (cudaStream_t)stream);
}
int fac_wn_end_coupling_f32(const fac_wg_model* m, int flow, const float* skip, float* audio, int B, int Tg,
                            void* stream) {
  return fac::wg_end(m, flow, skip, audio, B, Tg, (cudaStream_t)stream);
}
int fac_waveglow_infer_f32(const fac_wg_model* m, const float* mel_cl, float* audio, const fac_wg_workspace* ws, int B,
                           int F, void* stream) {
  return fac::wg_infer(m, mel_cl, audio, ws, B, F, (cudaStream_t)stream);
}

int fac_waveglow_tc_prepare_spect(const fac_wg_model* m, const fac_wg_tc_weights* w, const fac_wg_tc_workspace* ws,
                                  const float* mel_cl, int B, int F, int nsplit, void* stream) {
  return fac::wg_tc_prepare_spect(m, w, ws, mel_cl, B, F, nsplit, (cudaStream_t)stream);
}
int fac_wn_start_tc(const fac_wg_model* m, int flow, const float* audio, const fac_wg_tc_workspace* ws, int B, int Tg,
                    int nsplit, void* stream) {
  return fac::wg_tc_start(m, flow, audio, ws, B, Tg, nsplit, (cudaStream_t)stream);
}
int fac_wn_layer_tc(const fac_wg_model* m, const fac_wg_tc_weights* w, int flow, int layer,
                    const fac_wg_tc_workspace* ws, int B, int Tg, int nsplit, void* stream) {
  return fac::wg_tc_layer(m, w, flow, layer, ws, B, Tg, nsplit, (cudaStream_t)stream);
}
int fac_wn_end_tc(const fac_wg_model* m, const fac_wg_tc_weights* w, int flow, const float* out8, float* audio, int B,
                  int Tg, void* stream) {
  return fac::wg_tc_end(m, w, flow, out8, audio, B, Tg, (cudaStream_t)stream);
}
int fac_waveglow_infer_tc(const fac_wg_model* m, const fac_wg_tc_weights* w, const float* mel_cl, float* audio,
                          const fac_wg_tc_workspace* ws, int B, int F, int nsplit, void* stream) {
  return fac::wg_infer_tc(m, w, mel_cl, audio, ws, B, F, nsplit, (cudaStream_t)stream);
}
int fac_waveglow_flow_step_tc(const fac_wg_model* m, const fac_wg_tc_weights* w, int flow, float* audio,
                              const fac_wg_tc_workspace* ws, int B, int Tg, int nsplit, void* stream) {
  return fac::wg_tc_flow_step(m, w, flow, audio, ws, B, Tg, nsplit, (cudaStream_t)stream);
}
void fac_tc_set_profile_buffer(long long* device_buf) { fac::tc_set_prof(device_buf); }
int fac_conv_gemm_tc(const fac_tc_conv* conv, void* stream) { return fac::conv_gemm_tc(conv, (cudaStream_t)stream); }
int fac_transpose_split_16(const float* in, void* hi, void* lo, int B, int C, int T, int pad, int fp16, void* stream) {
  return fac::tc_transpose_split(in, hi, lo, B, C, T, pad, fp16, (cudaStream_t)stream);
}
int fac_pad_split_16(const float* in, void* hi, void* lo, long long n_rows, int C, int pad, int fp16, void* stream) {
  return fac::tc_pad_split(in, hi, lo, n_rows, C, pad, fp16, (cudaStream_t)stream);
}
void fac_lstm_set_profile_buffer(long long* device_buf) { fac::lstm_set_prof(device_buf); }
void fac_taco_set_profile_buffer(long long* device_buf) { fac::taco_set_prof(device_buf); }
int fac_tc_set_batch_group(int utterances) { return fac::tc_set_batch_group(utterances); }
int fac_tc_set_cta_group(int cta_group) { return fac::tc_set_cta_group(cta_group); }
int fac_tc_set_k_block(int k_block) { return fac::tc_set_k_block(k_block); }
int fac_tc_set_fused(int enabled) { return fac::tc_set_fused(enabled); }
int fac_selftest_grid_barrier(unsigned int* zeroed_counter, int iters, void* stream) {
  return fac::selftest_grid_barrier(zeroed_counter, iters, (cudaStream_t)stream);
}
int fac_denoise_spectrum_f32(float* spec, const float* bias_mag, float strength, long long n_rows, int n_bins, int ld,
                               void* stream) {
  return fac::denoise_spectrum(spec, bias_mag, strength, n_rows, n_bins, ld, (cudaStream_t)stream);
}
int fac_ppg_sparsify(const float* ppg, int* idx, float* val, int* overflow, int B, int D, int T, int k,
                     float threshold, void* stream) {
  return fac::ppg_sparsify(ppg, idx, val, overflow, B, D, T, k, threshold, (cudaStream_t)stream);
}
int fac_prenet0_sparse_f32(const int* idx, const float* val, const float* w_t, int w_ld, const float* mask,
                           const int* row_lengths, float* out, int out_ld, void* out_hi, void* out_lo, int B, int T,
                           int k, int D, int E, int pad, void* stream) {
  return fac::prenet0_sparse(idx, val, w_t, w_ld, mask, row_lengths, out, out_ld, out_hi, out_lo, B, T, k, D, E, pad,
                             (cudaStream_t)stream);
}
int fac_lstm_bidir_f32(const float* xp, const float* w_hh, float* out, int B, int T, int H, void* stream) {
  return fac::lstm_bidir(xp, w_hh, out, nullptr, B, T, H, (cudaStream_t)stream);
}
int fac_lstm_bidir_var_f32(const float* xp, const float* w_hh, float* out, const int* lengths, int B, int T, int H,
                           void* stream) {
  return fac::lstm_bidir(xp, w_hh, out, lengths, B, T, H, (cudaStream_t)stream);
}
int fac_taco_decoder_run(const fac_taco_decoder_weights* w, const float* memory, const float* pmem, const int* lengths,
                         const unsigned char* drop, const fac_taco_decoder_state* state, float* mel, float* gate,
                         float* align, int B, int T_in, int max_steps, int window, float gate_threshold, void* stream) {
  return fac::taco_decoder_run(w, memory, pmem, lengths, drop, state, mel, gate, align, B, T_in, max_steps, window,
                               gate_threshold, (cudaStream_t)stream);
}

}  // extern "C"
