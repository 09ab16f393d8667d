"""CPU emulation (torch) of what the CUDA kernels compute FROM THE PACKED BUFFERS.

Test infrastructure: lets the CPU-only suite check the weight packing, the phase
decomposition of the upsampler, the gate interleave and the slot layout of the
audio buffer against the oracle without a GPU.  Mirrors csrc/gemm_conv_f32.cu and
csrc/waveglow_f32.cu one-to-one in semantics (not in performance).
"""
import torch

from fac_via_ppg_b200.packing import upsample_taps
from fac_via_ppg_b200.synth import flow_channels


def gather_rows(src, taps, dil, center):
    """src (B, T, C) channels-last -> (B, T, taps*C): row t holds rows t + k*dil - center (zero outside)."""
    B, T, C = src.shape
    cols = []
    for k in range(taps):
        shift = k * dil - center
        out = src.new_zeros(B, T, C)
        lo, hi = max(0, -shift), min(T, T - shift)
        if hi > lo:
            out[:, lo:hi] = src[:, lo + shift:hi + shift]
        cols.append(out)
    return torch.cat(cols, dim=-1)


def conv_gemm(srcs, w, bias, n):
    """srcs: list of (tensor (B,T,C), taps, dil, center); w (K, N_pad); returns (B, T, n)."""
    a = torch.cat([gather_rows(*s) for s in srcs], dim=-1)
    out = a @ w[:, :n]
    return out if bias is None else out + bias[:n]


def waveglow_infer(packed, mel, audio):
    """packed: PackedWaveGlow on CPU; mel (B, n_mel, F); audio (B, Tg, G) pre-filled with sigma*z."""
    cfg, lay, flat = packed.cfg, packed.layout, packed.flat
    wn = cfg["WN_config"]
    C, L, ks = wn["n_channels"], wn["n_layers"], wn["kernel_size"]
    G, hop, n_mel = cfg["n_group"], cfg["hop_length"], cfg["n_mel_channels"]
    n_cond, phases, taps = n_mel * G, hop // G, upsample_taps(cfg)
    B, _, F = mel.shape
    Tg = F * phases
    mel_cl = mel.transpose(1, 2).contiguous()
    spect = mel.new_zeros(B, F, phases, n_cond)
    w_up, b_up = lay.view(flat, "upsample_w"), lay.view(flat, "upsample_b")
    for p in range(phases):
        spect[:, :, p] = conv_gemm([(mel_cl, taps, -1, 0)], w_up[p], b_up, n_cond)
    spect = spect.view(B, Tg, n_cond)
    audio = audio.clone()
    chans = flow_channels(cfg)
    for k in reversed(range(cfg["n_flows"])):
        n_rem, n_half = chans[k]
        off = G - n_rem
        x = audio[:, :, off:off + n_half] @ lay.view(flat, f"{k}.start_w") + lay.view(flat, f"{k}.start_b")
        skip = None
        for i in range(L):
            d = 2 ** i
            pre = conv_gemm([(x, ks, d, d * (ks - 1) // 2), (spect, 1, 0, 0)],
                            lay.view(flat, f"{k}.{i}.in_cond_w"), lay.view(flat, f"{k}.{i}.in_cond_b"), 2 * C)
            acts = torch.tanh(pre[..., 0::2]) * torch.sigmoid(pre[..., 1::2])
            n_rs = 2 * C if i < L - 1 else C
            rs = conv_gemm([(acts, 1, 0, 0)], lay.view(flat, f"{k}.{i}.res_skip_w"),
                           lay.view(flat, f"{k}.{i}.res_skip_b"), n_rs)
            if i < L - 1:
                x = x + rs[..., :C]
                s = rs[..., C:]
            else:
                s = rs
            skip = s if skip is None else skip + s
        out = skip @ lay.view(flat, f"{k}.end_w").t() + lay.view(flat, f"{k}.end_b")
        a0 = audio[:, :, off:off + n_half]
        a1 = (audio[:, :, off + n_half:] - out[..., :n_half]) / torch.exp(out[..., n_half:])
        y = torch.cat([a0, a1], dim=-1)
        audio[:, :, off:] = y @ lay.view(flat, f"{k}.w_inv").t()
    return audio.reshape(B, Tg * G)


def fill_audio_slots(noise, sigma, G):
    """Slot layout used by WaveGlow.infer: first draw owns the last slots."""
    B, _, Tg = noise[0].shape
    audio = torch.empty(B, Tg, G)
    hi = G
    for z in noise:
        lo = hi - z.shape[1]
        audio[:, :, lo:hi] = (sigma * z).transpose(1, 2)
        hi = lo
    assert hi == 0
    return audio
