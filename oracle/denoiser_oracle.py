"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference Denoiser / STFT.

Restates reference src/common/stft.py:49-138 (STFT as Conv1d / ConvTranspose1d with a windowed
DFT basis and its pseudo-inverse), src/common/audio_processing.py:39-88 (window_sumsquare) and
src/waveglow/denoiser.py:35-68 (bias-spectrum subtraction).  Pinned against the unmodified
reference modules by tests/test_oracle_pinning.py when /root/reference is present.
Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def hann_periodic(n):
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)      # scipy.signal.get_window('hann', n, fftbins=True)


def stft_bases(filter_length=1024, hop_length=160, win_length=1024):
    """reference stft.py:54-77."""
    scale = filter_length / hop_length
    basis = np.fft.fft(np.eye(filter_length))
    cutoff = filter_length // 2 + 1
    basis = np.vstack([np.real(basis[:cutoff]), np.imag(basis[:cutoff])])
    window = np.zeros(filter_length)
    lpad = (filter_length - win_length) // 2
    window[lpad:lpad + win_length] = hann_periodic(win_length)
    fwd = torch.FloatTensor(basis[:, None, :]) * torch.from_numpy(window).float()
    inv = torch.FloatTensor(np.linalg.pinv(scale * basis).T[:, None, :]) * torch.from_numpy(window).float()
    return fwd.float(), inv.float(), window


def window_sumsquare(window, n_frames, hop_length, n_fft):
    """reference audio_processing.py:39-88 (norm=None)."""
    n = n_fft + hop_length * (n_frames - 1)
    x = np.zeros(n, dtype=np.float32)
    win_sq = (window ** 2).astype(np.float32)
    for i in range(n_frames):
        s = i * hop_length
        x[s:min(n, s + n_fft)] += win_sq[:max(0, min(n_fft, n - s))]
    return x


def transform(x, fwd, filter_length=1024, hop_length=160):
    """reference stft.py:79-104: reflect pad, strided conv, magnitude / phase."""
    pad = filter_length // 2
    xp = F.pad(x[:, None, None, :], (pad, pad, 0, 0), mode="reflect").squeeze(1)
    ft = F.conv1d(xp, fwd, stride=hop_length)
    cutoff = filter_length // 2 + 1
    re, im = ft[:, :cutoff], ft[:, cutoff:]
    return torch.sqrt(re ** 2 + im ** 2), torch.atan2(im, re)


def inverse(magnitude, phase, inv, window, filter_length=1024, hop_length=160):
    """reference stft.py:106-138."""
    spec = torch.cat([magnitude * torch.cos(phase), magnitude * torch.sin(phase)], dim=1)
    out = F.conv_transpose1d(spec, inv, stride=hop_length)
    wss = window_sumsquare(window, magnitude.size(-1), hop_length, filter_length)
    nz = torch.from_numpy(np.where(wss > np.finfo(np.float32).tiny)[0])
    out[:, :, nz] /= torch.from_numpy(wss)[nz]
    out *= float(filter_length) / hop_length
    return out[:, :, filter_length // 2: -(filter_length // 2)]


def denoise(audio, bias_audio, strength, filter_length=1024, hop_length=160, win_length=1024):
    """reference denoiser.py:56-68: audio (B, N), bias_audio (1, M) = infer(zeros(1,80,88), sigma=0)."""
    fwd, inv, window = stft_bases(filter_length, hop_length, win_length)
    bias_spec, _ = transform(bias_audio.float(), fwd, filter_length, hop_length)
    bias_spec = bias_spec[:, :, 0][:, :, None]
    mag, phase = transform(audio.float(), fwd, filter_length, hop_length)
    mag = torch.clamp(mag - bias_spec * strength, 0.0)
    return inverse(mag, phase, inv, window, filter_length, hop_length)
