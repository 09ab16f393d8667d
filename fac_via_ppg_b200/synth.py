"""Seeded synthetic weights and inputs for the PPG -> Mel -> WaveGlow path.

There are no checkpoints or datasets on the build or GPU boxes, so every test,
the smoke run and the benchmark use weights drawn here.  The draw is a pure
function of (seed, config) on the CPU generator, so the same state can be
rebuilt on any box and loaded both into the reference modules (in the build
container) and into the drop-in modules of this package.

Key names and shapes follow the reference state dicts *after*
``WaveGlow.remove_weightnorm`` (reference src/waveglow/glow.py:295-311) and of
``Tacotron2`` (reference src/common/model.py:538-560), see SURVEY.md section 8a.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

SEED = 16807  # the reference's own seed (src/common/hparams.py:223, src/waveglow/config.json:9)

# src/waveglow/config.json:29-41
WAVEGLOW_CONFIG = {
    "n_mel_channels": 80,
    "hop_length": 160,
    "n_flows": 12,
    "n_group": 8,
    "n_early_every": 4,
    "n_early_size": 2,
    "WN_config": {"n_layers": 8, "n_channels": 256, "kernel_size": 3},
}

# A reduced geometry for fast CPU tests (same structure, fewer/lighter flows).
WAVEGLOW_CONFIG_SMALL = {
    "n_mel_channels": 80,
    "hop_length": 160,
    "n_flows": 4,
    "n_group": 8,
    "n_early_every": 2,
    "n_early_size": 2,
    "WN_config": {"n_layers": 3, "n_channels": 64, "kernel_size": 3},
}

UPSAMPLE_KERNEL = 1024  # src/waveglow/glow.py:184-186


def _uniform(gen, shape, fan_in):
    bound = 1.0 / math.sqrt(fan_in)
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2.0 - 1.0) * bound


def _normal(gen, shape, std, mean=0.0):
    return torch.randn(shape, generator=gen, dtype=torch.float32) * std + mean


def flow_channels(cfg):
    """Per-flow (n_remaining_channels, n_half), reference src/waveglow/glow.py:195-207."""
    n_half = cfg["n_group"] // 2
    n_rem = cfg["n_group"]
    out = []
    for k in range(cfg["n_flows"]):
        if k % cfg["n_early_every"] == 0 and k > 0:
            n_half -= cfg["n_early_size"] // 2
            n_rem -= cfg["n_early_size"]
        out.append((n_rem, n_half))
    return out


def waveglow_state(seed: int = SEED, cfg=None, end_std: float = 0.02, convinv: str = "orthonormal"):
    """State dict of a weight-norm-free WaveGlow with non-degenerate weights.

    ``convinv='orthonormal'`` keeps the reference's initialisation of the invertible 1x1
    convolutions (a rotation, glow.py:74-80), for which W^-1 == W^T; ``convinv='general'``
    draws W = Q1 diag(s) Q2^T with singular values s in [0.5, 2], what a trained checkpoint
    looks like: only there is the inverse of glow.py:89-95 distinguishable from a transpose.

    ``WN.end`` is zero-initialised by the reference constructor
    (src/waveglow/glow.py:126-130), which would turn every coupling into the
    identity and hide errors in the WN stack; it is drawn from N(0, end_std^2)
    here instead.
    """
    cfg = cfg or WAVEGLOW_CONFIG
    gen = torch.Generator().manual_seed(seed)
    n_mel = cfg["n_mel_channels"]
    n_group = cfg["n_group"]
    wn = cfg["WN_config"]
    C, L, ks = wn["n_channels"], wn["n_layers"], wn["kernel_size"]
    n_cond = n_mel * n_group
    sd = OrderedDict()
    sd["upsample.weight"] = _uniform(gen, (n_mel, n_mel, UPSAMPLE_KERNEL), n_mel * UPSAMPLE_KERNEL / cfg["hop_length"])
    sd["upsample.bias"] = _uniform(gen, (n_mel,), n_mel)
    for k, (n_rem, n_half) in enumerate(flow_channels(cfg)):
        p = f"WN.{k}."
        sd[p + "start.weight"] = _uniform(gen, (C, n_half, 1), n_half)
        sd[p + "start.bias"] = _uniform(gen, (C,), n_half)
        sd[p + "end.weight"] = _normal(gen, (2 * n_half, C, 1), end_std)
        sd[p + "end.bias"] = _normal(gen, (2 * n_half,), end_std)
        for i in range(L):
            sd[p + f"in_layers.{i}.weight"] = _uniform(gen, (2 * C, C, ks), C * ks)
            sd[p + f"in_layers.{i}.bias"] = _uniform(gen, (2 * C,), C * ks)
            sd[p + f"cond_layers.{i}.weight"] = _uniform(gen, (2 * C, n_cond, 1), n_cond)
            sd[p + f"cond_layers.{i}.bias"] = _uniform(gen, (2 * C,), n_cond)
            n_rs = 2 * C if i < L - 1 else C
            sd[p + f"res_skip_layers.{i}.weight"] = _uniform(gen, (n_rs, C, 1), C)
            sd[p + f"res_skip_layers.{i}.bias"] = _uniform(gen, (n_rs,), C)
        # orthonormal with det +1, as the reference initialises it (glow.py:74-80)
        q, _ = torch.linalg.qr(_normal(gen, (n_rem, n_rem), 1.0))
        if torch.det(q) < 0:
            q[:, 0] = -q[:, 0]
        if convinv == "general":
            q2, _ = torch.linalg.qr(_normal(gen, (n_rem, n_rem), 1.0))
            s = torch.exp2(torch.rand((n_rem,), generator=gen) * 2.0 - 1.0)      # log-uniform in [0.5, 2]
            q = (q * s[None, :]) @ q2.t()
        elif convinv != "orthonormal":
            raise ValueError("convinv must be 'orthonormal' or 'general'")
        sd[f"convinv.{k}.conv.weight"] = q.contiguous().view(n_rem, n_rem, 1)
    return sd


# Tacotron2 (PPG->Mel) dimensions, reference src/common/hparams.py:167-231.
TACOTRON_HPARAMS = {
    "n_symbols": 5816,
    "symbols_embedding_dim": 600,
    "encoder_embedding_dim": 600,
    "encoder_kernel_size": 5,
    "encoder_n_convolutions": 3,
    "n_acoustic_feat_dims": 80,
    "prenet_dim": 300,
    "attention_rnn_dim": 300,
    "decoder_rnn_dim": 300,
    "attention_dim": 150,
    "attention_location_n_filters": 32,
    "attention_location_kernel_size": 31,
    "attention_window_size": 20,
    "postnet_embedding_dim": 512,
    "postnet_kernel_size": 5,
    "postnet_n_convolutions": 5,
    "gate_threshold": 0.5,
    "max_decoder_steps": 1000,
    "p_attention_dropout": 0.1,
    "p_decoder_dropout": 0.1,
}


def _bn(gen, sd, prefix, n):
    sd[prefix + "weight"] = torch.rand((n,), generator=gen) + 0.5
    sd[prefix + "bias"] = _normal(gen, (n,), 0.1)
    sd[prefix + "running_mean"] = _normal(gen, (n,), 0.1)
    sd[prefix + "running_var"] = torch.rand((n,), generator=gen) + 0.5
    sd[prefix + "num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def _lstm_like(gen, sd, prefix, n_in, n_hid, suffix=""):
    sd[prefix + "weight_ih" + suffix] = _uniform(gen, (4 * n_hid, n_in), n_hid) * 2.0
    sd[prefix + "weight_hh" + suffix] = _uniform(gen, (4 * n_hid, n_hid), n_hid) * 2.0
    sd[prefix + "bias_ih" + suffix] = _uniform(gen, (4 * n_hid,), n_hid)
    sd[prefix + "bias_hh" + suffix] = _uniform(gen, (4 * n_hid,), n_hid)


def tacotron_state(seed: int = SEED, hp=None, gain: float = 1.0):
    """State dict for the PPG->Mel model with BatchNorm statistics randomised so
    that BN folding is exercised."""
    hp = dict(TACOTRON_HPARAMS, **(hp or {}))
    gen = torch.Generator().manual_seed(seed + 1)
    E = hp["encoder_embedding_dim"]
    D = hp["n_symbols"]
    M = hp["n_acoustic_feat_dims"]
    P = hp["prenet_dim"]
    A = hp["attention_dim"]
    R = hp["attention_rnn_dim"]
    Rd = hp["decoder_rnn_dim"]
    sd = OrderedDict()
    # The PPG is a posterior (rows sum to 1), so the first prenet layer sees a
    # convex combination of its columns: scale by sqrt(D) to keep activations O(1).
    sd["encoder.prenet.layers.0.linear_layer.weight"] = _uniform(gen, (E, D), 1.0) * 3.0 * gain
    sd["encoder.prenet.layers.1.linear_layer.weight"] = _uniform(gen, (E, E), E) * 2.0
    ke = hp["encoder_kernel_size"]
    for i in range(hp["encoder_n_convolutions"]):
        p = f"encoder.convolutions.{i}."
        sd[p + "0.conv.weight"] = _uniform(gen, (E, E, ke), E * ke) * 2.0
        sd[p + "0.conv.bias"] = _uniform(gen, (E,), E * ke)
        _bn(gen, sd, p + "1.", E)
    H = E // 2
    _lstm_like(gen, sd, "encoder.lstm.", E, H, "_l0")
    _lstm_like(gen, sd, "encoder.lstm.", E, H, "_l0_reverse")
    sd["decoder.prenet.layers.0.linear_layer.weight"] = _uniform(gen, (P, M), M) * 2.0
    sd["decoder.prenet.layers.1.linear_layer.weight"] = _uniform(gen, (P, P), P) * 2.0
    _lstm_like(gen, sd, "decoder.attention_rnn.", P + E, R)
    al = "decoder.attention_layer."
    sd[al + "query_layer.linear_layer.weight"] = _uniform(gen, (A, R), R) * 2.0
    sd[al + "memory_layer.linear_layer.weight"] = _uniform(gen, (A, E), E) * 6.0
    sd[al + "v.linear_layer.weight"] = _uniform(gen, (1, A), A) * 16.0
    nf, kf = hp["attention_location_n_filters"], hp["attention_location_kernel_size"]
    sd[al + "location_layer.location_conv.conv.weight"] = _uniform(gen, (nf, 2, kf), 2 * kf) * 2.0
    sd[al + "location_layer.location_dense.linear_layer.weight"] = _uniform(gen, (A, nf), nf) * 2.0
    _lstm_like(gen, sd, "decoder.decoder_rnn.", R + E, Rd)
    sd["decoder.linear_projection.linear_layer.weight"] = _uniform(gen, (M, Rd + E), Rd + E) * 12.0
    sd["decoder.linear_projection.linear_layer.bias"] = _uniform(gen, (M,), Rd + E)
    sd["decoder.gate_layer.linear_layer.weight"] = _uniform(gen, (1, Rd + E), Rd + E)
    sd["decoder.gate_layer.linear_layer.bias"] = _uniform(gen, (1,), Rd + E)
    kp = hp["postnet_kernel_size"]
    Pe = hp["postnet_embedding_dim"]
    n_post = hp["postnet_n_convolutions"]
    dims = [M] + [Pe] * (n_post - 1) + [M]
    for i in range(n_post):
        p = f"postnet.convolutions.{i}."
        sd[p + "0.conv.weight"] = _uniform(gen, (dims[i + 1], dims[i], kp), dims[i] * kp) * 2.0
        sd[p + "0.conv.bias"] = _uniform(gen, (dims[i + 1],), dims[i] * kp)
        _bn(gen, sd, p + "1.", dims[i + 1])
    return sd


def synthetic_mel(batch: int, frames: int, seed: int = SEED, n_mel: int = 80):
    """Log-mel-like input: N(-5, 2^2) clipped to the range of the reference's
    dynamic-range compression (src/common/audio_processing.py:110-116)."""
    gen = torch.Generator().manual_seed(seed + 2)
    return (torch.randn((batch, n_mel, frames), generator=gen) * 2.0 - 5.0).clamp_(-11.5, 2.0)


def synthetic_ppg(batch: int, frames: int, seed: int = SEED, n_symbols: int = 5816):
    """Posteriorgram-like input (B, n_symbols, T): every frame is a softmax so it
    sums to one, as the reference's test/test_ppg.py:54 asserts for real PPGs."""
    gen = torch.Generator().manual_seed(seed + 3)
    logits = torch.randn((batch, frames, n_symbols), generator=gen) * 3.0
    return torch.softmax(logits, dim=-1).transpose(1, 2).contiguous()


def frames_for_seconds(seconds: float, hop: int = 160, rate: int = 22050) -> int:
    """Mel frames covering ``seconds`` of audio at the benchmark's 22.05 kHz
    convention (10 s -> 1379 frames)."""
    return int(math.ceil(seconds * rate / hop))
