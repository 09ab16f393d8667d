"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference WaveGlow reverse flow.

A functional, weight-dict driven restatement of ``WaveGlow.infer`` and the
modules it calls, written against the reference source (each function cites the
lines it follows).  It exists so that the GPU box -- which has no /root/reference
-- still has a checker.  It is pinned against the real reference by
tests/test_oracle_pinning.py (live, when /root/reference is present) and by the
golden vectors in tests/golden/ that oracle/make_golden.py produced by running
the *unmodified* reference modules.  The reference's own test-suite holds no
vectors for this path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  It must never be on the product path.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def gated_activation(in_act: torch.Tensor, n_channels: int) -> torch.Tensor:
    """reference src/waveglow/glow.py:33-40 (fused_add_tanh_sigmoid_multiply, after the add)."""
    return torch.tanh(in_act[:, :n_channels]) * torch.sigmoid(in_act[:, n_channels:])


def wn_forward(sd, prefix: str, audio_0: torch.Tensor, spect: torch.Tensor, n_layers: int, n_channels: int):
    """reference src/waveglow/glow.py:154-175 (WN.forward)."""
    x = F.conv1d(audio_0, sd[prefix + "start.weight"], sd[prefix + "start.bias"])
    skip_total = None
    for i in range(n_layers):
        w_in = sd[prefix + f"in_layers.{i}.weight"]
        dilation = 2 ** i
        padding = (w_in.shape[2] * dilation - dilation) // 2
        pre = F.conv1d(x, w_in, sd[prefix + f"in_layers.{i}.bias"], dilation=dilation, padding=padding)
        pre = pre + F.conv1d(spect, sd[prefix + f"cond_layers.{i}.weight"], sd[prefix + f"cond_layers.{i}.bias"])
        acts = gated_activation(pre, n_channels)
        rs = F.conv1d(acts, sd[prefix + f"res_skip_layers.{i}.weight"], sd[prefix + f"res_skip_layers.{i}.bias"])
        if i < n_layers - 1:
            x = rs[:, :n_channels] + x
            skip = rs[:, n_channels:]
        else:
            skip = rs
        skip_total = skip if skip_total is None else skip + skip_total
    return F.conv1d(skip_total, sd[prefix + "end.weight"], sd[prefix + "end.bias"])


def upsample_and_squeeze(sd, cfg, mel: torch.Tensor) -> torch.Tensor:
    """reference src/waveglow/glow.py:253-259: transposed conv, trim, group-squeeze.

    Returns (B, n_mel*n_group, T_g); channel m*n_group+j of column t is the
    upsampled mel channel m at sample n_group*t+j."""
    hop, n_group = cfg["hop_length"], cfg["n_group"]
    w = sd["upsample.weight"]
    up = F.conv_transpose1d(mel, w, sd["upsample.bias"], stride=hop)
    up = up[:, :, : -(w.shape[2] - hop)]
    B, C, T = up.shape
    up = up.unfold(2, n_group, n_group).permute(0, 2, 1, 3)
    return up.contiguous().view(B, up.size(1), -1).permute(0, 2, 1)


def invertible_1x1_reverse(weight: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """reference src/waveglow/glow.py:82-97 with reverse=True: z <- W^-1 z."""
    w_inv = weight.squeeze(-1).inverse().to(z.dtype)   # glow.py:89-95 (cast follows the input)
    return F.conv1d(z, w_inv[..., None])


def noise_shapes(cfg, batch: int, n_cols: int):
    """Shapes of the N(0,1) draws of one infer() call, in draw order
    (reference src/waveglow/glow.py:261-270, 285-290)."""
    n_rem = cfg["n_group"]
    n_early = 0
    for k in range(cfg["n_flows"]):
        if k % cfg["n_early_every"] == 0 and k > 0:
            n_rem -= cfg["n_early_size"]
            n_early += 1
    return [(batch, n_rem, n_cols)] + [(batch, cfg["n_early_size"], n_cols)] * n_early


def draw_noise(cfg, batch: int, n_cols: int, device="cpu", dtype=torch.float32):
    """Consumes torch's global generator exactly like the reference infer() does."""
    return [torch.empty(s, device=device, dtype=dtype).normal_() for s in noise_shapes(cfg, batch, n_cols)]


def waveglow_infer(sd, cfg, mel: torch.Tensor, sigma: float = 1.0, noise=None) -> torch.Tensor:
    """reference src/waveglow/glow.py:252-293 (WaveGlow.infer).

    ``noise`` is the list of unit-normal draws (see noise_shapes); when None it
    is drawn from torch's global generator in the reference order."""
    wn = cfg["WN_config"]
    spect = upsample_and_squeeze(sd, cfg, mel)
    if noise is None:
        noise = draw_noise(cfg, mel.shape[0], spect.shape[2], mel.device, mel.dtype)
    noise = list(noise)
    audio = sigma * noise.pop(0)
    for k in reversed(range(cfg["n_flows"])):
        n_half = audio.size(1) // 2
        audio_0, audio_1 = audio[:, :n_half], audio[:, n_half:]
        out = wn_forward(sd, f"WN.{k}.", audio_0, spect, wn["n_layers"], wn["n_channels"])
        b, s = out[:, :n_half], out[:, n_half:]
        audio_1 = (audio_1 - b) / torch.exp(s)
        audio = torch.cat([audio_0, audio_1], 1)
        audio = invertible_1x1_reverse(sd[f"convinv.{k}.conv.weight"], audio)
        if k % cfg["n_early_every"] == 0 and k > 0:
            audio = torch.cat((sigma * noise.pop(0), audio), 1)
    return audio.permute(0, 2, 1).contiguous().view(audio.size(0), -1)
