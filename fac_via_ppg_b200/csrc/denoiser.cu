// Spectral bias subtraction of the Denoiser (reference src/waveglow/denoiser.py:63-68 with
// src/common/stft.py:99-111): magnitude = sqrt(re^2 + im^2), magnitude' = max(magnitude - bias*strength, 0),
// recombined with the original phase.  cos/sin(atan2(im, re)) are re/|z| and im/|z|, so the kernel scales
// (re, im) by magnitude'/magnitude in place; the STFT / inverse STFT around it are fac_conv_gemm_f32 calls
// (hop-reshaped signal, 7 taps).
#include "fac_common.cuh"

namespace fac {
namespace {

__global__ void denoise_spectrum_kernel(float* __restrict__ spec, const float* __restrict__ bias_mag, float strength,
                                        long long n_rows, int n_bins, int ld) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rows * n_bins) return;
  const long long row = idx / n_bins;
  const int bin = (int)(idx - row * n_bins);
  float* re = spec + row * ld + bin;
  float* im = re + n_bins;
  const float r = *re, i = *im;
  const float mag = sqrtf(r * r + i * i);
  const float kept = fmaxf(mag - __ldg(bias_mag + bin) * strength, 0.f);
  const float scale = mag > 0.f ? kept / mag : 0.f;
  *re = r * scale;
  *im = i * scale;
}

}  // namespace

int denoise_spectrum(float* spec, const float* bias_mag, float strength, long long n_rows, int n_bins, int ld,
                     cudaStream_t st) {
  FAC_REQUIRE(spec && bias_mag && n_rows > 0 && n_bins > 0 && ld >= 2 * n_bins, "denoise_spectrum: bad arguments");
  const long long total = n_rows * n_bins;
  denoise_spectrum_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(spec, bias_mag, strength, n_rows, n_bins, ld);
  count_launch();
  return check_launch("denoise_spectrum_kernel");
}

}  // namespace fac
