"""Drop-in for the reference ``src/waveglow/denoiser.py`` (the step right after ``infer`` on the CLI
path, generate_synthesis.py:58-62, 94-95).  SURVEY.md section 8f ranks it "next": the bias audio comes
from the CUDA-native ``WaveGlow.infer(zeros, sigma=0)``; the STFT / inverse STFT (reference
src/common/stft.py:79-138, a dense-DFT conv) are still expressed with torch ops here -- plumbing
around the hot path, not a hand-written kernel yet.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _hann(win_length):
    n = np.arange(win_length)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)       # scipy get_window('hann', fftbins=True)


class STFT(torch.nn.Module):
    """Windowed DFT as Conv1d / ConvTranspose1d, same bases as reference stft.py:54-77."""

    def __init__(self, filter_length=800, hop_length=200, win_length=800):
        super().__init__()
        self.filter_length, self.hop_length, self.win_length = filter_length, hop_length, win_length
        basis = np.fft.fft(np.eye(filter_length))
        cutoff = filter_length // 2 + 1
        basis = np.vstack([np.real(basis[:cutoff]), np.imag(basis[:cutoff])])
        window = np.zeros(filter_length)
        lpad = (filter_length - win_length) // 2
        window[lpad:lpad + win_length] = _hann(win_length)
        scale = filter_length / hop_length
        fwd = torch.tensor(basis[:, None, :] * window, dtype=torch.float32)
        inv = torch.tensor(np.linalg.pinv(scale * basis).T[:, None, :] * window, dtype=torch.float32)
        self.register_buffer("forward_basis", fwd)
        self.register_buffer("inverse_basis", inv)
        self.register_buffer("window_sq", torch.tensor(window ** 2, dtype=torch.float32))

    def transform(self, x):
        pad = self.filter_length // 2
        x = F.pad(x[:, None, None, :], (pad, pad, 0, 0), mode="reflect").squeeze(1)
        ft = F.conv1d(x, self.forward_basis, stride=self.hop_length)
        cutoff = self.filter_length // 2 + 1
        re, im = ft[:, :cutoff], ft[:, cutoff:]
        return torch.sqrt(re * re + im * im), torch.atan2(im, re)

    def inverse(self, magnitude, phase):
        spec = torch.cat([magnitude * torch.cos(phase), magnitude * torch.sin(phase)], dim=1)
        out = F.conv_transpose1d(spec, self.inverse_basis, stride=self.hop_length)
        n_frames = magnitude.size(-1)
        # window sum-square envelope (reference audio_processing.py:39-88) as one transposed conv
        ones = torch.ones(1, 1, n_frames, device=out.device)
        wss = F.conv_transpose1d(ones, self.window_sq[None, None, :], stride=self.hop_length)[0, 0]
        nz = wss > torch.finfo(torch.float32).tiny
        out[:, :, nz] = out[:, :, nz] / wss[nz]
        out = out * (float(self.filter_length) / self.hop_length)
        half = self.filter_length // 2
        return out[:, :, half:-half]


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
            raise Exception("Mode {} if not supported".format(mode))
        with torch.no_grad():
            bias_audio = waveglow.infer(mel_input, sigma=0.0).float()
            bias_spec, _ = self.stft.transform(bias_audio)
        self.register_buffer("bias_spec", bias_spec[:, :, 0][:, :, None])

    def forward(self, audio, strength=0.1):
        spec, angles = self.stft.transform(audio.cuda().float())
        spec = torch.clamp(spec - self.bias_spec * strength, 0.0)
        return self.stft.inverse(spec, angles)
