"""GPU parity tests for the WaveGlow reverse flow: CUDA path (through the C ABI) vs the
oracle / the golden vectors of the unmodified reference.  Tolerances: the north star
asks for <= 1e-4 RMS on the waveform; the exact-fp32 path is held to 1e-5."""
import ctypes as C
import os

import pytest
import torch

import emulate
from fac_via_ppg_b200 import _ext, ops, synth
from fac_via_ppg_b200.waveglow.glow import WaveGlow
from oracle import waveglow_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rms(a, b):
    return (a.double().cpu() - b.double().cpu()).pow(2).mean().sqrt().item()


def build_model(cfg, seed=synth.SEED, **state_kwargs):
    model = WaveGlow(**cfg)
    model = WaveGlow.remove_weightnorm(model)
    model.load_state_dict(synth.waveglow_state(seed=seed, cfg=cfg, **state_kwargs), strict=True)
    return model.to(DEV).eval().set_precision("fp32")      # this file tests the exact-fp32 FFMA path


# ---------------------------------------------------------------- kernel granularity
@pytest.mark.parametrize("rows,batch", [(1, 1), (77, 3), (128, 2), (300, 2)])
def test_conv_gemm_two_sources_dilated(rows, batch):
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(batch, rows, 64, generator=g)
    s = torch.randn(batch, rows, 24, generator=g)
    w = torch.randn(3 * 64 + 24, 200, generator=g) * 0.1
    bias = torch.randn(200, generator=g)
    mask = (torch.rand(batch, rows, 200, generator=g) > 0.5).float() * 2
    res = torch.randn(batch, rows, 200, generator=g)
    ref = torch.relu(emulate.conv_gemm([(x, 3, 4, 4), (s, 1, 0, 0)], w, bias, 200)) * mask + res
    wp, bp = ops.pack_gemm_weight(w.to(DEV), bias.to(DEV))
    out = torch.empty(batch, rows, 200, device=DEV)
    xd, sdv, md, rd = x.to(DEV), s.to(DEV), mask.to(DEV), res.to(DEV)
    ops.conv_gemm([ops.conv_src(xd, 3, 4, 4), ops.conv_src(sdv)], wp, bp, 200, out, batch=batch,
                  rows=rows, act=_ext.ACT_RELU, mask=md, residual=rd)
    assert (out.cpu() - ref).abs().max().item() <= 1e-4


def test_conv_gemm_channel_major_source_and_odd_width():
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 40, 37, generator=g)            # (B, C, T) with T = 37: unaligned rows
    w = torch.randn(5 * 40, 150, generator=g) * 0.1
    ref = torch.tanh(emulate.conv_gemm([(x.transpose(1, 2).contiguous(), 5, 1, 2)], w, None, 150))
    wp, _ = ops.pack_gemm_weight(w.to(DEV))
    out = torch.empty(2, 37, 150, device=DEV)
    xd = x.to(DEV)
    ops.conv_gemm([ops.conv_src(xd, 5, 1, 2, channel_major=True)], wp, None, 150, out, batch=2, rows=37,
                  act=_ext.ACT_TANH)
    assert (out.cpu() - ref).abs().max().item() <= 1e-4


def test_gate_and_res_skip_epilogues():
    g = torch.Generator().manual_seed(6)
    B, T, Cn = 2, 150, 64
    x = torch.randn(B, T, Cn, generator=g)
    w = torch.randn(Cn, 2 * Cn, generator=g) * 0.2
    b = torch.randn(2 * Cn, generator=g)
    pre = emulate.conv_gemm([(x, 1, 0, 0)], w, b, 2 * Cn)
    ref_gate = torch.tanh(pre[..., 0::2]) * torch.sigmoid(pre[..., 1::2])
    wp, bp = ops.pack_gemm_weight(w.to(DEV), b.to(DEV))
    acts = torch.empty(B, T, Cn, device=DEV)
    xd = x.to(DEV)
    ops.conv_gemm([ops.conv_src(xd)], wp, bp, 2 * Cn, acts, batch=B, rows=T, kind=_ext.EPI_GATE)
    assert (acts.cpu() - ref_gate).abs().max().item() <= 1e-5
    res0, skip0 = torch.randn(B, T, Cn, generator=g), torch.randn(B, T, Cn, generator=g)
    res, skip = res0.to(DEV), skip0.to(DEV)
    ops.conv_gemm([ops.conv_src(xd)], wp, bp, 2 * Cn, res, batch=B, rows=T, kind=_ext.EPI_RES_SKIP,
                  out2=skip, n_split=Cn, accumulate_out2=True)
    assert (res.cpu() - (res0 + pre[..., :Cn])).abs().max().item() <= 1e-4
    assert (skip.cpu() - (skip0 + pre[..., Cn:])).abs().max().item() <= 1e-4


@pytest.mark.parametrize("frames", [1, 3, 50])
def test_upsample_squeeze_matches_oracle(frames):
    cfg = synth.WAVEGLOW_CONFIG
    sd = synth.waveglow_state(cfg=cfg)
    model = build_model(cfg)
    mel = synth.synthetic_mel(2, frames, seed=frames)
    ref = waveglow_oracle.upsample_and_squeeze(sd, cfg, mel).transpose(1, 2)
    spect = torch.empty(2, frames * 20, 640, device=DEV)
    mel_cl = mel.transpose(1, 2).contiguous().to(DEV)
    rc = _ext.load().fac_waveglow_upsample_squeeze_f32(C.byref(model.packed().cmodel), mel_cl.data_ptr(),
                                                       spect.data_ptr(), 2, frames, _ext.current_stream())
    _ext.check(rc, "upsample")
    assert (spect.cpu() - ref).abs().max().item() <= 2e-5


# ---------------------------------------------------------------- whole infer()
@pytest.mark.parametrize("name", ["waveglow_small_b2_f6.pt", "waveglow_full_b2_f5.pt",
                                  "waveglow_full_b1_f88_sigma0.pt", "waveglow_small_b2_f6_general_convinv.pt",
                                  "waveglow_full_b2_f5_general_convinv.pt"])
def test_infer_matches_reference_golden(golden_dir, name):
    """The *_general_convinv cases carry non-orthogonal invertible 1x1 weights (singular values in [0.5, 2]), so a
    transpose in place of the inverse of glow.py:89-95, or a row/column mix-up of W^-1, fails them."""
    g = torch.load(os.path.join(golden_dir, name))
    model = build_model(g["cfg"], **g.get("state_kwargs", {}))
    mel = synth.synthetic_mel(g["batch"], g["frames"], seed=g["mel_seed"]).to(DEV)
    audio = model.infer(mel, sigma=g["sigma"], noise=[z.to(DEV) for z in g["noise"]])
    assert audio.shape == g["audio"].shape
    assert rms(audio, g["audio"]) <= 1e-5          # north star: 1e-4 RMS
    assert (audio.cpu() - g["audio"]).abs().max().item() <= 1e-4


def test_infer_default_noise_follows_reference_draw_order():
    cfg = synth.WAVEGLOW_CONFIG_SMALL
    model = build_model(cfg)
    mel = synth.synthetic_mel(2, 4).to(DEV)
    torch.manual_seed(123)
    a = model.infer(mel, sigma=0.7)
    torch.manual_seed(123)
    # small geometry: 4 flows, early output every 2 -> first draw has 6 channels, then 2 at k=2
    noise = [torch.empty(2, 6, 80, device=DEV).normal_(), torch.empty(2, 2, 80, device=DEV).normal_()]
    b = model.infer(mel, sigma=0.7, noise=noise)
    assert torch.equal(a, b)
    ref = waveglow_oracle.waveglow_infer(synth.waveglow_state(cfg=cfg), cfg, mel.cpu(), 0.7,
                                         [z.cpu() for z in noise])
    assert rms(a, ref) <= 1e-5


def test_infer_with_weightnorm_checkpoint_flavour():
    """generate_synthesis.py:58-61 runs infer() on a model that still carries weight norm."""
    cfg = synth.WAVEGLOW_CONFIG_SMALL
    torch.manual_seed(3)
    model = WaveGlow(**cfg)
    for wn in model.WN:                      # the constructor zero-inits `end`: make it matter
        torch.nn.init.normal_(wn.end.weight, std=0.05)
    model = model.to(DEV).eval().set_precision("fp32")
    mel = synth.synthetic_mel(1, 3).to(DEV)
    noise = [torch.randn(1, 6, 60, device=DEV), torch.randn(1, 2, 60, device=DEV)]
    a = model.infer(mel, 0.5, noise=noise)
    sd = {k: v.detach().cpu() for k, v in model.plain_state().items()}
    ref = waveglow_oracle.waveglow_infer(sd, cfg, mel.cpu(), 0.5, [z.cpu() for z in noise])
    assert rms(a, ref) <= 1e-5
    model = WaveGlow.remove_weightnorm(model)
    b = model.infer(mel, 0.5, noise=noise)
    assert rms(a, b) <= 1e-6


def test_infer_empty_and_dtype():
    model = build_model(synth.WAVEGLOW_CONFIG_SMALL)
    assert model.infer(torch.zeros(0, 80, 5, device=DEV)).shape == (0, 800)
    out = model.infer(synth.synthetic_mel(1, 2).to(DEV).half(), sigma=0.0)
    assert out.dtype == torch.float16 and out.shape == (1, 320)
    with pytest.raises(_ext.FacError):
        model.infer(torch.zeros(1, 80, 2))          # CPU tensor: no fallback, loud failure


def test_full_size_batch_prefix_locality():
    """BASELINE config 2 size (8 x 10 s).  Size-independent properties: (1) utterances are
    independent -- row 0 of the batch equals the same utterance run alone, bit for bit;
    (2) locality -- the first columns only see the first frames (receptive field
    12 flows x 255 columns + 7 upsample taps), so they match the oracle run on a prefix."""
    cfg = synth.WAVEGLOW_CONFIG
    model = build_model(cfg)
    B, F, Fp = 8, synth.frames_for_seconds(10.0), 240
    mel = synth.synthetic_mel(B, F, seed=31).to(DEV)
    torch.manual_seed(17)
    noise = model.noise_like_reference(B, F * 20, DEV, torch.float32)
    full = model.infer(mel, 0.6, noise=noise)
    assert full.shape == (B, F * 160) and torch.isfinite(full).all()
    alone = model.infer(mel[:1].contiguous(), 0.6, noise=[z[:1].contiguous() for z in noise])
    assert torch.equal(alone[0], full[0])
    ref = waveglow_oracle.waveglow_infer(synth.waveglow_state(cfg=cfg), cfg, mel[:1, :, :Fp].cpu(), 0.6,
                                         [z[:1, :, :Fp * 20].cpu() for z in noise])
    safe_cols = Fp * 20 - 12 * 255 - 8 * 20
    assert safe_cols > 1000
    assert rms(full[0, :safe_cols * 8], ref[0, :safe_cols * 8]) <= 1e-5
