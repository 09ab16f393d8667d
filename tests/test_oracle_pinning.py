"""Pins oracle/*.py: (a) against the golden vectors produced by the UNMODIFIED
reference modules (tests/golden, see oracle/make_golden.py) -- runs everywhere;
(b) live against the reference when /root/reference is present."""
import os

import pytest
import torch

from fac_via_ppg_b200 import synth
from oracle import ref_shim, tacotron_oracle, waveglow_oracle

WG_CASES = ["waveglow_small_b2_f6.pt", "waveglow_full_b2_f5.pt", "waveglow_full_b1_f88_sigma0.pt",
            "waveglow_small_b2_f6_general_convinv.pt", "waveglow_full_b2_f5_general_convinv.pt"]


@pytest.mark.parametrize("name", WG_CASES)
def test_waveglow_oracle_matches_golden(golden_dir, name):
    g = torch.load(os.path.join(golden_dir, name))
    sd = synth.waveglow_state(cfg=g["cfg"], **g.get("state_kwargs", {}))
    mel = synth.synthetic_mel(g["batch"], g["frames"], seed=g["mel_seed"])
    out = waveglow_oracle.waveglow_infer(sd, g["cfg"], mel, g["sigma"], g["noise"])
    assert out.shape == g["audio"].shape
    if g.get("state_kwargs", {}).get("convinv") == "general":
        # the case exists to tell W^-1 from W^T (glow.py:82-97): the transpose must NOT reproduce the golden
        wrong = dict(sd)
        for k in range(g["cfg"]["n_flows"]):
            w = sd[f"convinv.{k}.conv.weight"][:, :, 0]
            wrong[f"convinv.{k}.conv.weight"] = w.t().inverse().contiguous()[:, :, None]
        bad = waveglow_oracle.waveglow_infer(wrong, g["cfg"], mel, g["sigma"], g["noise"])
        assert (bad - g["audio"]).abs().max().item() > 1e-2
    # same arithmetic, same library: tolerance only covers thread-count dependent summation order
    assert (out - g["audio"]).abs().max().item() <= 2e-5


def test_tacotron_oracle_matches_golden(golden_dir):
    g = torch.load(os.path.join(golden_dir, "tacotron_b1_t24.pt"))
    sd = synth.tacotron_state()
    ppg = synth.synthetic_ppg(1, g["t_in"], seed=g["ppg_seed"])
    masks = [m.float() for m in g["masks"]]
    mel, mel_post, gate, align = tacotron_oracle.tacotron_inference(
        sd, synth.TACOTRON_HPARAMS, ppg, masks, gate_threshold=2.0, max_decoder_steps=g["t_in"])
    assert (mel - g["mel"]).abs().max().item() <= 1e-4
    assert (mel_post - g["mel_post"]).abs().max().item() <= 1e-4
    assert (gate - g["gate"]).abs().max().item() <= 1e-4
    assert (align - g["align"]).abs().max().item() <= 2e-4   # sharp softmax: energies are O(10)


def test_dropout_tape_reproduces_reference_draw_order():
    torch.manual_seed(5)
    tape = tacotron_oracle.record_dropout_tape(1, 7, 3)
    assert [tuple(m.shape) for m in tape] == [(1, 7, 600)] * 2 + [(1, 300)] * 6
    assert all(set(m.unique().tolist()) <= {0.0, 2.0} for m in tape)


def test_window_mask_quirk():
    # src/common/utils.py:65-69: beyond the end only the last frame stays unmasked
    m = tacotron_oracle.window_mask([10], 3, 50, 10)
    assert m[0].tolist() == [True] * 9 + [False]
    m = tacotron_oracle.window_mask([10], 3, 0, 10)
    assert m[0].tolist() == [False] * 4 + [True] * 6


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")
def test_waveglow_oracle_matches_live_reference():
    cfg = synth.WAVEGLOW_CONFIG_SMALL
    sd = synth.waveglow_state(cfg=cfg, seed=99)
    model = ref_shim.reference_waveglow(sd, cfg)
    mel = synth.synthetic_mel(3, 7, seed=4)
    torch.manual_seed(8)
    with torch.no_grad():
        ref = model.infer(mel, sigma=0.8)
    torch.manual_seed(8)
    out = waveglow_oracle.waveglow_infer(sd, cfg, mel, 0.8)
    assert (out - ref).abs().max().item() <= 2e-5


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")
def test_tacotron_oracle_matches_live_reference():
    sd = synth.tacotron_state(seed=77)
    model = ref_shim.reference_tacotron(sd)
    model.decoder.gate_threshold, model.decoder.max_decoder_steps = 2.0, 12
    ppg = synth.synthetic_ppg(1, 12, seed=6)
    torch.manual_seed(9)
    with torch.no_grad():
        ref = model.inference(ppg)
    torch.manual_seed(9)
    out = tacotron_oracle.tacotron_inference(sd, synth.TACOTRON_HPARAMS, ppg, None, 2.0, 12)
    for a, b in zip(ref, out):
        assert (a - b).abs().max().item() <= 1e-4


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")
def test_denoiser_oracle_matches_live_reference_stft():
    """oracle/denoiser_oracle.py vs the unmodified reference src/common/stft.py (transform and inverse)."""
    from oracle import denoiser_oracle as do
    ref_shim.install()
    from common.stft import STFT  # type: ignore
    torch.manual_seed(0)
    x = torch.randn(2, 3200) * 0.3
    ref = STFT(1024, 160, 1024)
    fwd, inv, window = do.stft_bases()
    mag_r, ph_r = ref.transform(x)
    mag_o, ph_o = do.transform(x, fwd)
    assert (mag_r - mag_o).abs().max().item() <= 1e-6 and (ph_r - ph_o).abs().max().item() <= 1e-6
    assert (ref.inverse(mag_r, ph_r) - do.inverse(mag_o, ph_o, inv, window)).abs().max().item() <= 1e-5
