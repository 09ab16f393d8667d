"""Drop-in for the reference ``src/waveglow/denoiser.py`` (the step right after ``infer`` on the CLI
path, generate_synthesis.py:58-62, 94-95), CUDA-native.

The reference's STFT is a dense windowed-DFT Conv1d (stride = hop) and its inverse a
ConvTranspose1d with the pseudo-inverse basis (src/common/stft.py:54-138).  With the signal reshaped
into rows of ``hop`` samples both become 7-tap implicit GEMMs (forward: taps +1 over the sample rows;
inverse: taps -1 over the frames, exactly like WaveGlow's upsampler), so they run on the same kernels as
the rest of the path: by default ``fac_conv_gemm_tc`` (tcgen05 tensor cores, split-fp16 operands, K-chunked
fp32 accumulation: the 1026 x 1120 DFT is a contraction), with ``precision = 'fp32'`` the exact FFMA
``fac_conv_gemm_f32``.  The window-sum-square normalisation (src/common/audio_processing.py:39-88) and the
hop scaling are folded into the inverse GEMM's epilogue as a multiplicative mask, and the spectral
subtraction (denoiser.py:63-68) is ``fac_denoise_spectrum_f32``.  Only reflect-padding / cropping stay in torch.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from fac_via_ppg_b200 import _ext, ops


def _hann_periodic(n):
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)       # scipy get_window('hann', n, fftbins=True)


class STFT(torch.nn.Module):
    """Windowed DFT bases of reference stft.py:54-77, packed for the hop-reshaped GEMM form."""

    def __init__(self, filter_length=800, hop_length=200, win_length=800):
        super().__init__()
        if filter_length % 2 or hop_length % 8:
            raise ValueError("filter_length must be even and hop_length a multiple of 8")
        self.filter_length, self.hop_length, self.win_length = filter_length, hop_length, win_length
        self.taps = -(-filter_length // hop_length)
        self.cutoff = filter_length // 2 + 1
        self.n_out = 2 * self.cutoff
        self.ld = (self.n_out + 7) // 8 * 8                          # leading dimension of a spectrum row
        basis = np.fft.fft(np.eye(filter_length))
        basis = np.vstack([np.real(basis[:self.cutoff]), np.imag(basis[:self.cutoff])])
        window = np.zeros(filter_length)
        lpad = (filter_length - win_length) // 2
        window[lpad:lpad + win_length] = _hann_periodic(win_length)
        scale = filter_length / hop_length
        fwd = torch.tensor(basis * window, dtype=torch.float32)                               # (n_out, n_fft)
        inv = torch.tensor(np.linalg.pinv(scale * basis).T * window, dtype=torch.float32)    # (n_out, n_fft)
        k_pad = self.taps * hop_length
        w_f = torch.zeros(k_pad, self.n_out)
        w_f[:filter_length] = fwd.t()                                # rows = sample offset tap*hop + c
        inv_pad = torch.zeros(self.n_out, k_pad)
        inv_pad[:, :filter_length] = inv
        w_i = torch.zeros(self.taps, self.ld, hop_length)            # rows = (tap, bin), cols = sample in hop
        w_i[:, :self.n_out] = inv_pad.view(self.n_out, self.taps, hop_length).permute(1, 0, 2)
        wf_p, _ = ops.pack_gemm_weight(w_f)
        wi_p, _ = ops.pack_gemm_weight(w_i.reshape(self.taps * self.ld, hop_length))
        self.register_buffer("w_forward", wf_p)
        self.register_buffer("w_inverse", wi_p)
        self.register_buffer("window_sq", torch.tensor(window ** 2, dtype=torch.float32))
        self._norm_cache = {}
        # tensor-core form (built on first use, on the buffers' device): [n_pad][taps * c_pad] IEEE-half hi/lo,
        # K contiguous and tap-major, input channels zero-padded to a multiple of 64 (fac_conv_gemm_tc)
        self._fwd_kn = w_f                                   # (taps*hop, n_out)
        self._inv_tbh = w_i                                  # (taps, ld, hop)
        self._tc = None

    precision = "fp16x3"        # 'fp16x3': tcgen05 tensor cores (fp32-grade) | 'fp32': exact FFMA implicit GEMM

    def _tc_weights(self, device):
        if self._tc is not None and self._tc[0] == str(device):
            return self._tc[1]
        hop, taps, ld = self.hop_length, self.taps, self.ld
        rnd = lambda n: (n + 63) // 64 * 64                                       # noqa: E731

        def pack(w_ntc, n_valid):                # (n, taps, c) fp32 -> dict for ops.conv_gemm_tc
            n, _, c = w_ntc.shape
            n_pad, c_pad = rnd(max(n, n_valid)), rnd(c)
            wp = torch.zeros(n_pad, taps, c_pad)
            wp[:n, :, :c] = w_ntc
            wp = wp.reshape(n_pad, taps * c_pad).to(device)
            hi = wp.to(torch.float16)
            return dict(hi=hi.contiguous(), lo=(wp - hi.float()).to(torch.float16).contiguous(), c_pad=c_pad, taps=taps,
                        n_pad=n_pad, n_valid=n_valid, bias=None)

        n_valid_f = (self.n_out + 3) // 4 * 4                                     # 1026 -> 1028 (two zero columns)
        fwd = pack(self._fwd_kn.t().reshape(self.n_out, taps, hop), n_valid_f)
        # output row r reads spectrum frames r - k: conv_gemm_tc reads rows t + tap' - center with center = taps - 1,
        # tap' = taps - 1 - k
        inv = pack(self._inv_tbh.flip(0).permute(2, 0, 1).contiguous(), hop)     # (hop, taps, ld)
        self._tc = (str(device), (fwd, inv))
        return self._tc[1]

    def transform_raw(self, x):
        """x (B, N) -> spectrum rows (B, frames, ld) = [real | imag | pad] (reference stft.py:79-97)."""
        spec, frames = self._transform_rows(x)
        return spec[:, :frames]

    def _transform_rows(self, x):
        """-> (spectrum buffer (B, >= frames, ld) whose rows past `frames` are zero, frames)."""
        _ext.require_cuda(x, "audio")
        B, N = x.shape
        hop, pad = self.hop_length, self.filter_length // 2
        xp = F.pad(x.float()[:, None, None, :], (pad, pad, 0, 0), mode="reflect").view(B, -1)
        frames = (xp.shape[1] - self.filter_length) // hop + 1
        rows = frames + self.taps - 1
        if xp.shape[1] < rows * hop:
            xp = F.pad(xp, (0, rows * hop - xp.shape[1]))
        src = xp[:, : rows * hop].contiguous().view(B, rows, hop)
        if self.precision == "fp16x3":
            fwd, _ = self._tc_weights(x.device)
            # all `rows` output rows are computed; the taps - 1 rows past the last frame are written as zeros
            spec = torch.zeros(B, rows, self.ld, device=x.device, dtype=torch.float32)
            lens = torch.full((B,), frames, dtype=torch.int32, device=x.device)
            a = ops.pad_split(src, fwd["c_pad"])
            ops.conv_gemm_tc(a, dict(fwd, center_override=0), out=spec, want_split=False, row_lengths=lens)
            return spec, frames
        spec = torch.zeros(B, frames, self.ld, device=x.device, dtype=torch.float32)
        ops.conv_gemm([ops.conv_src(src, self.taps, 1, 0)], self.w_forward, None, self.n_out, spec, batch=B,
                      rows=frames, out_batch_stride=frames * self.ld, out_row_stride=self.ld)
        return spec, frames

    def transform(self, x):
        """reference STFT.transform: magnitude and phase (B, cutoff, frames)."""
        spec = self.transform_raw(x)
        re, im = spec[..., : self.cutoff].transpose(1, 2), spec[..., self.cutoff: self.n_out].transpose(1, 2)
        return torch.sqrt(re * re + im * im), torch.atan2(im, re)

    def _normaliser(self, frames, batch, device):
        """hop scaling / window-sum-square envelope of reference stft.py:119-132, as the epilogue mask."""
        key = (frames, batch, str(device))
        if key not in self._norm_cache:
            hop, n_fft = self.hop_length, self.filter_length
            rows = frames + self.taps - 1
            ones = torch.ones(1, 1, frames, device=device)
            wss = F.conv_transpose1d(ones, self.window_sq[None, None, :].to(device), stride=hop)[0, 0]
            norm = torch.full_like(wss, float(n_fft) / hop)
            nz = wss > torch.finfo(torch.float32).tiny
            norm[nz] = norm[nz] / wss[nz]
            full = torch.zeros(rows * hop, device=device)
            full[: wss.numel()] = norm
            self._norm_cache[key] = full.view(1, rows, hop).expand(batch, rows, hop).contiguous()
        return self._norm_cache[key]

    def inverse_raw(self, spec, n_samples, frames=None):
        """spectrum rows (B, frames, ld) -> audio (B, 1, n_samples) (reference stft.py:106-138).  ``frames``: the
        buffer holds that many frames followed by taps - 1 zero rows (what _transform_rows returns)."""
        B = spec.shape[0]
        hop = self.hop_length
        padded = frames is not None
        frames = spec.shape[1] if frames is None else frames
        rows = frames + self.taps - 1
        out = torch.empty(B, rows, hop, device=spec.device, dtype=torch.float32)
        if self.precision == "fp16x3":
            _, inv = self._tc_weights(spec.device)
            if padded and spec.shape[1] == rows and spec.is_contiguous():
                full = spec
            else:
                full = torch.zeros(B, rows, self.ld, device=spec.device, dtype=torch.float32)  # frames past the end: zero
                full[:, :frames] = spec[:, :frames]
            a = ops.pad_split(full, inv["c_pad"])
            ops.conv_gemm_tc(a, dict(inv, center_override=self.taps - 1), out=out, want_split=False,
                             mask=self._normaliser(frames, B, spec.device))
        else:
            ops.conv_gemm([ops.conv_src(spec[:, :frames].contiguous(), self.taps, -1, 0)], self.w_inverse, None, hop, out,
                          batch=B, rows=rows, mask=self._normaliser(frames, B, spec.device))
        half = self.filter_length // 2
        return out.view(B, 1, rows * hop)[:, :, half: half + n_samples]


class Denoiser(torch.nn.Module):
    """Removes the model bias from WaveGlow audio (reference denoiser.py:35-68)."""

    def __init__(self, waveglow, filter_length=1024, hop_length=160, win_length=1024, mode="zeros"):
        super().__init__()
        w = waveglow.upsample.weight
        self.stft = STFT(filter_length, hop_length, win_length).to(w.device)
        if mode == "zeros":
            mel_input = torch.zeros((1, 80, 88), dtype=w.dtype, device=w.device)
        elif mode == "normal":
            mel_input = torch.randn((1, 80, 88), dtype=w.dtype, device=w.device)
        else:
            raise Exception("Denoiser mode %r is unknown: use 'zeros' or 'normal'" % (mode,))
        with torch.no_grad():
            bias_audio = waveglow.infer(mel_input, sigma=0.0).float()
            bias_spec, _ = self.stft.transform(bias_audio)
        self.register_buffer("bias_spec", bias_spec[:, :, 0][:, :, None].contiguous())

    @torch.no_grad()
    def forward(self, audio, strength=0.1):
        audio = audio.cuda().float()
        spec, frames = self.stft._transform_rows(audio)          # (B, frames [+ zero rows], ld), contiguous
        B, n_rows, ld = spec.shape
        rc = _ext.load().fac_denoise_spectrum_f32(spec.data_ptr(), self.bias_spec.data_ptr(), float(strength),
                                                  B * n_rows, self.stft.cutoff, ld, _ext.current_stream())
        _ext.check(rc, "fac_denoise_spectrum_f32")
        return self.stft.inverse_raw(spec, audio.shape[1], frames=frames)
