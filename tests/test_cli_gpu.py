"""The CLI / glue surface on the GPU: generate_synthesis (reference src/script/generate_synthesis.py),
checkpoint formats (pickled WaveGlow module, Tacotron2 state dict) and the Denoiser drop-in."""
import os

import numpy as np
import pytest
import torch
from scipy.io import wavfile

import fac_via_ppg_b200
from fac_via_ppg_b200 import synth
from fac_via_ppg_b200.common.hparams import create_hparams_stage
from fac_via_ppg_b200.common.utils import (get_inference, get_mask_from_lengths_window_and_time_step,
                                           load_waveglow_model, waveglow_audio)
from fac_via_ppg_b200.script import generate_synthesis
from fac_via_ppg_b200.script.train_ppg2mel import load_model
from fac_via_ppg_b200.waveglow.denoiser import Denoiser
from fac_via_ppg_b200.waveglow.glow import WaveGlow
from oracle import tacotron_oracle

pytestmark = pytest.mark.gpu


def small_waveglow():
    cfg = synth.WAVEGLOW_CONFIG_SMALL
    m = WaveGlow(**cfg)                                     # still weight-normed, like a training checkpoint
    for wn in m.WN:
        torch.nn.init.normal_(wn.end.weight, std=0.05)
    return m


def test_cli_synthetic_end_to_end(tmp_path):
    out = tmp_path / "out"
    rc = generate_synthesis.main(["--ppg2mel_model", "none", "--waveglow_model", "none", "--teacher_utterance_path",
                                  "none", "--output_dir", str(out), "--synthetic", "0.4"])
    assert rc == 0
    fs, wav = wavfile.read(str(out / "ac.wav"))
    assert fs == 16000 and wav.dtype == np.float32 and wav.ndim == 1
    assert wav.shape[0] == 40 * 160 and np.isfinite(wav).all() and np.abs(wav).max() > 0
    assert "Done!" in (out / "debug.log").read_text()


def test_checkpoint_formats_and_glue(tmp_path):
    fac_via_ppg_b200.install_aliases()
    # WaveGlow: pickled module under key 'model' (reference train_waveglow.py:56-64)
    wg_path = str(tmp_path / "waveglow.pt")
    torch.save({"model": small_waveglow(), "iteration": 1}, wg_path)
    wg = load_waveglow_model(wg_path)
    assert not hasattr(wg.WN[0].start, "weight_g") and next(wg.parameters()).is_cuda
    # Tacotron2: {'state_dict': ...} (reference train_ppg2mel.py:143-149)
    taco_path = str(tmp_path / "taco.pt")
    torch.save({"state_dict": synth.tacotron_state(), "iteration": 1}, taco_path)
    taco = load_model(create_hparams_stage())
    taco.load_state_dict(torch.load(taco_path, weights_only=False)["state_dict"])
    taco.eval()
    taco.decoder.gate_threshold, taco.decoder.max_decoder_steps = 2.0, 12
    ppg = synth.synthetic_ppg(1, 12)[0].t().numpy()          # (T, D) like get_ppg returns
    mel = get_inference(ppg, taco)
    assert mel.shape == (1, 80, 12) and mel.is_cuda
    audio = waveglow_audio(mel, wg, 0.6, True)
    assert audio.shape == (1, 12 * 160) and torch.isfinite(audio).all()
    pcm = waveglow_audio(mel, wg, 0.6, False)
    assert pcm.dtype == np.int16 and pcm.shape == (12 * 160,)


def test_denoiser_is_identity_at_zero_strength_and_removes_bias():
    wg = WaveGlow.remove_weightnorm(small_waveglow()).cuda().eval()
    den = Denoiser(wg, mode="zeros")
    assert den.bias_spec.shape == (1, 513, 1)
    t = torch.arange(16000, device="cuda") / 16000.0
    audio = (0.3 * torch.sin(2 * np.pi * 220 * t) + 0.1 * torch.sin(2 * np.pi * 1330 * t))[None]
    same = den(audio, strength=0.0)[:, 0]
    assert same.shape == audio.shape
    assert (same - audio)[:, 1024:-1024].abs().max().item() <= 1e-3     # STFT -> iSTFT reconstructs
    less = den(audio, strength=1.0)[:, 0]
    assert less.pow(2).mean() <= audio.pow(2).mean() + 1e-6            # magnitudes only shrink


def test_window_mask_helper_matches_oracle():
    lengths = torch.tensor([10, 7, 3], device="cuda")
    for step in (0, 2, 5, 9, 30):
        got = get_mask_from_lengths_window_and_time_step(lengths, 3, step).cpu()
        ref = tacotron_oracle.window_mask([10, 7, 3], 3, step, 10)
        assert torch.equal(got, ref), step


def test_denoiser_matches_reference_restatement():
    """CUDA Denoiser (hop-reshaped STFT GEMMs + spectral kernel) vs the oracle restatement of
    reference denoiser.py / stft.py on the same bias audio."""
    from oracle import denoiser_oracle
    wg = WaveGlow.remove_weightnorm(small_waveglow()).cuda().eval()
    den = Denoiser(wg, mode="zeros")
    bias_audio = wg.infer(torch.zeros(1, 80, 88, device="cuda"), sigma=0.0).float().cpu()
    g = torch.Generator().manual_seed(1)
    audio = torch.randn(2, 4800, generator=g) * 0.2
    for precision in ("fp16x3", "fp32"):          # tensor-core STFT GEMMs (default) and the exact FFMA form
        den.stft.precision = precision
        for strength in (0.005, 0.5):
            ref = denoiser_oracle.denoise(audio, bias_audio, strength)
            out = den(audio.cuda(), strength=strength)
            assert out.shape == ref.shape == (2, 1, 4800)
            err = (out.cpu() - ref).abs().max().item()
            print("denoiser %s strength %g: max-abs vs oracle %.2e" % (precision, strength, err))
            assert err <= 2e-4, (precision, strength)
    den.stft.precision = "fp16x3"
    mag_tc, ph_tc = den.stft.transform(audio.cuda())
    den.stft.precision = "fp32"
    mag_f, ph_f = den.stft.transform(audio.cuda())
    den.stft.precision = "fp16x3"
    assert mag_tc.shape == mag_f.shape == (2, 513, 31) and (mag_tc - mag_f).abs().max().item() <= 1e-4


def test_cli_with_checkpoints_and_npy_ppg(tmp_path, monkeypatch):
    """f4: the non-synthetic branch of the CLI -- Tacotron2 state-dict checkpoint, pickled WaveGlow module, a
    precomputed (T, 5816) .npy PPG as --teacher_utterance_path (what get_ppg returns,
    reference src/common/data_utils.py:55-59) -- with the on-disk pack cache enabled (f3)."""
    fac_via_ppg_b200.install_aliases()
    monkeypatch.setenv("FAC_PACK_CACHE", str(tmp_path / "packs"))
    taco_path, wg_path, ppg_path = (str(tmp_path / n) for n in ("taco.pt", "waveglow.pt", "teacher.npy"))
    state = synth.tacotron_state()
    state["decoder.gate_layer.linear_layer.bias"] = torch.tensor([-10.0])        # never fires: 1000 frames
    torch.save({"state_dict": state, "iteration": 1}, taco_path)
    torch.save({"model": small_waveglow(), "iteration": 1}, wg_path)
    np.save(ppg_path, synth.synthetic_ppg(1, 30, seed=3)[0].t().numpy())          # (T, 5816)
    out = tmp_path / "out"
    torch.manual_seed(0)
    rc = generate_synthesis.main(["--ppg2mel_model", taco_path, "--waveglow_model", wg_path,
                                  "--teacher_utterance_path", ppg_path, "--output_dir", str(out)])
    assert rc == 0
    fs, wav = wavfile.read(str(out / "ac.wav"))
    # the gate never fires: the decoder runs to max_decoder_steps = 1000 (hparams.py:216) like the reference
    assert fs == 16000 and wav.dtype == np.float32 and wav.shape == (1000 * 160,) and np.isfinite(wav).all()
    assert os.listdir(tmp_path / "packs")                                        # the WaveGlow pack was cached
    # a missing utterance is reported like the reference does (generate_synthesis.py:99-100), not raised
    rc = generate_synthesis.main(["--ppg2mel_model", taco_path, "--waveglow_model", wg_path,
                                  "--teacher_utterance_path", str(tmp_path / "nope.npy"), "--output_dir", str(out)])
    assert rc == 1


def test_old_format_waveglow_checkpoint_loads_and_matches(tmp_path):
    """f3: a pickled WaveGlow in the res_layers / skip_layers format (reference convert_model.py:43-70) goes
    through load_waveglow_model and produces the same audio as the same weights in the current format."""
    fac_via_ppg_b200.install_aliases()
    new = small_waveglow()
    old = small_waveglow()
    old.load_state_dict(new.state_dict())
    C = synth.WAVEGLOW_CONFIG_SMALL["WN_config"]["n_channels"]
    wnorm = torch.nn.utils.weight_norm
    for wn in old.WN:
        wn.res_layers, wn.skip_layers = torch.nn.ModuleList(), torch.nn.ModuleList()
        for i, src in enumerate(wn.res_skip_layers):
            last = i == wn.n_layers - 1
            v, g = src.weight_v.detach(), src.weight_g.detach()
            w = v * (g / v.flatten(1).norm(dim=1).view(-1, 1, 1))
            b = src.bias.detach()
            if not last:
                res = torch.nn.Conv1d(C, C, 1)
                res.weight.data, res.bias.data = w[:C].clone(), b[:C].clone()
                wn.res_layers.append(wnorm(res, name="weight"))
            skip = torch.nn.Conv1d(C, C, 1)
            skip.weight.data, skip.bias.data = (w if last else w[C:]).clone(), (b if last else b[C:]).clone()
            wn.skip_layers.append(wnorm(skip, name="weight"))
        del wn.res_skip_layers
    old_path, new_path = str(tmp_path / "old.pt"), str(tmp_path / "new.pt")
    torch.save({"model": old}, old_path)
    torch.save({"model": new}, new_path)
    a, b = load_waveglow_model(old_path), load_waveglow_model(new_path)
    assert hasattr(a.WN[0], "res_skip_layers") and not hasattr(a.WN[0], "res_layers")
    mel = synth.synthetic_mel(1, 6, seed=2).cuda()
    noise = [torch.randn(1, 6, 120, device="cuda"), torch.randn(1, 2, 120, device="cuda")]
    for precision in ("fp32", "bf16x3"):
        xa = a.set_precision(precision).infer(mel, 0.6, noise=noise)
        xb = b.set_precision(precision).infer(mel, 0.6, noise=noise)
        assert (xa - xb).abs().max().item() <= 1e-5, precision
