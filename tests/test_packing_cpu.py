"""Host logic without a GPU: the packed-weight formulation (phase-decomposed upsampler,
gate-interleaved K-concatenated GEMM, slot layout of the audio buffer) reproduces the
oracle when evaluated with torch on the CPU (tests/emulate.py mirrors the kernels)."""
import os

import pytest
import torch

import emulate
from fac_via_ppg_b200 import synth
from fac_via_ppg_b200.packing import PackedWaveGlow, waveglow_layout
from oracle import waveglow_oracle


@pytest.mark.parametrize("name", ["waveglow_small_b2_f6.pt", "waveglow_full_b2_f5.pt"])
def test_packed_formulation_matches_golden(golden_dir, name):
    g = torch.load(os.path.join(golden_dir, name))
    cfg = g["cfg"]
    packed = PackedWaveGlow.from_state(synth.waveglow_state(cfg=cfg), cfg, "cpu")
    mel = synth.synthetic_mel(g["batch"], g["frames"], seed=g["mel_seed"])
    audio0 = emulate.fill_audio_slots(g["noise"], g["sigma"], cfg["n_group"])
    out = emulate.waveglow_infer(packed, mel, audio0)
    assert (out - g["audio"]).abs().max().item() <= 5e-5


def test_upsample_phase_decomposition_ragged_frames():
    cfg = synth.WAVEGLOW_CONFIG_SMALL
    sd = synth.waveglow_state(cfg=cfg, seed=3)
    packed = PackedWaveGlow.from_state(sd, cfg, "cpu")
    for frames in (1, 2, 7):        # fewer frames than taps exercises the zero fill
        mel = synth.synthetic_mel(1, frames, seed=frames)
        ref = waveglow_oracle.upsample_and_squeeze(sd, cfg, mel).transpose(1, 2)
        mel_cl = mel.transpose(1, 2).contiguous()
        w, b = packed.layout.view(packed.flat, "upsample_w"), packed.layout.view(packed.flat, "upsample_b")
        out = torch.stack([emulate.conv_gemm([(mel_cl, 7, -1, 0)], w[p], b, 640) for p in range(20)], dim=2)
        assert (out.reshape(1, frames * 20, 640) - ref).abs().max().item() <= 1e-5


def test_layout_is_deterministic_and_aligned():
    a, b = waveglow_layout(synth.WAVEGLOW_CONFIG), waveglow_layout(synth.WAVEGLOW_CONFIG)
    assert list(a.entries.items()) == list(b.entries.items())
    assert all(off % 64 == 0 for off, _ in a.entries.values())
    # 87.7 M parameters + phase padding of the upsampler
    assert 87e6 < a.size < 90e6


def test_packed_tacotron_formulation_matches_golden(golden_dir):
    """BN folding, hoisted LSTM projection, window-only attention and the packed decoder
    tables reproduce the unmodified reference (golden) when evaluated with torch on CPU."""
    from fac_via_ppg_b200.packing import PackedTacotron
    g = torch.load(os.path.join(golden_dir, "tacotron_b1_t24.pt"))
    packed = PackedTacotron.from_state(synth.tacotron_state(), synth.TACOTRON_HPARAMS, "cpu")
    ppg = synth.synthetic_ppg(1, g["t_in"], seed=g["ppg_seed"])
    masks = [m.float() for m in g["masks"]]
    mel, mel_post, gate, align = emulate.tacotron_inference(packed, ppg, masks, g["t_in"], window=20)
    # north-star tolerance on mel is 1e-3 max-abs; a different (but fp32) summation order stays well inside
    assert (mel - g["mel"]).abs().max().item() <= 3e-4
    assert (mel_post - g["mel_post"]).abs().max().item() <= 3e-4
    assert (gate - g["gate"]).abs().max().item() <= 3e-4
    assert (align - g["align"]).abs().max().item() <= 3e-4


def test_window_only_attention_equals_dense_masked_attention():
    """Longer than the window: positions outside [t-w, t+w] get exactly zero weight in the
    reference (utils.py:46-78 + masked_fill(-inf)), so evaluating only the window is exact."""
    from fac_via_ppg_b200.packing import PackedTacotron
    from oracle import tacotron_oracle
    sd = synth.tacotron_state(seed=5)
    packed = PackedTacotron.from_state(sd, synth.TACOTRON_HPARAMS, "cpu")
    T = 70
    ppg = synth.synthetic_ppg(1, T, seed=8)
    torch.manual_seed(4)
    masks = tacotron_oracle.record_dropout_tape(1, T, T)
    ref = tacotron_oracle.tacotron_inference(sd, synth.TACOTRON_HPARAMS, ppg, masks, 2.0, T)
    out = emulate.tacotron_inference(packed, ppg, masks, T, window=20)
    for a, b in zip(out, ref):
        assert (a - b).abs().max().item() <= 3e-4


@pytest.mark.parametrize("name", ["waveglow_small_b2_f6.pt", "waveglow_full_b2_f5.pt"])
def test_tensor_core_algebra_matches_golden(golden_dir, name):
    """Skip-path collapse (end() folded into per-layer 8-channel updates), residual add through an
    identity block and the bf16 hi+lo weight split reproduce the reference (torch fp32 on the CPU)."""
    g = torch.load(os.path.join(golden_dir, name))
    cfg = g["cfg"]
    packed = PackedWaveGlow.from_state(synth.waveglow_state(cfg=cfg), cfg, "cpu")
    mel = synth.synthetic_mel(g["batch"], g["frames"], seed=g["mel_seed"])
    audio0 = emulate.fill_audio_slots(g["noise"], g["sigma"], cfg["n_group"])
    out = emulate.waveglow_infer_tc(packed, mel, audio0)
    assert (out - g["audio"]).pow(2).mean().sqrt().item() <= 2e-5


def test_hop_reshaped_stft_formulation_matches_oracle():
    """The Denoiser's STFT / inverse STFT as 7-tap GEMMs over the hop-reshaped signal (forward taps +1,
    inverse taps -1, window-sum-square folded into a mask) equal the reference's strided Conv1d /
    ConvTranspose1d (oracle/denoiser_oracle.py, pinned to reference src/common/stft.py)."""
    import torch.nn.functional as F
    from fac_via_ppg_b200.waveglow.denoiser import STFT
    from oracle import denoiser_oracle as do
    stft = STFT(1024, 160, 1024)
    torch.manual_seed(2)
    x = torch.randn(2, 1600) * 0.3
    fwd, inv, window = do.stft_bases()
    mag, phase = do.transform(x, fwd)
    # forward
    xp = F.pad(x[:, None, None, :], (512, 512, 0, 0), mode="reflect").view(2, -1)
    frames = (xp.shape[1] - 1024) // 160 + 1
    rows = frames + stft.taps - 1
    xp = F.pad(xp, (0, rows * 160 - xp.shape[1]))
    spec = emulate.conv_gemm([(xp.view(2, rows, 160), stft.taps, 1, 0)], stft.w_forward, None, stft.n_out)[:, :frames]
    re, im = spec[..., :513].transpose(1, 2), spec[..., 513:].transpose(1, 2)
    assert (torch.sqrt(re * re + im * im) - mag).abs().max().item() <= 1e-4
    # inverse (zero rows appended so that every output row exists in the emulation)
    spec_ld = torch.zeros(2, rows, stft.ld)
    spec_ld[:, :frames, : stft.n_out] = spec
    out = emulate.conv_gemm([(spec_ld, stft.taps, -1, 0)], stft.w_inverse, None, 160)
    out = out * stft._normaliser(frames, 2, "cpu")
    rec = out.reshape(2, 1, rows * 160)[:, :, 512: 512 + 1600]
    ref = do.inverse(mag, phase, inv, window)
    assert (rec - ref).abs().max().item() <= 1e-4
    assert (rec[:, 0] - x).abs().max().item() <= 1e-3          # and both reconstruct the signal


def test_tacotron_tensor_core_weights_layout_cpu():
    """PackedTacotron.tc_weights(): [n_pad][taps * c_pad] half hi/lo pairs reproduce the packed fp32 weights
    (22 significand bits), padding rows / channels are exact zeros, biases are padded with zeros."""
    from fac_via_ppg_b200.packing import PackedTacotron
    packed = PackedTacotron.from_state(synth.tacotron_state(), synth.TACOTRON_HPARAMS, "cpu")
    tw = packed.tc_weights()
    assert set(tw) == {"enc.pre0", "enc.pre1", "enc.conv0", "enc.conv1", "enc.conv2", "enc.lstm_ih",
                       "post.conv0", "post.conv1", "post.conv2", "post.conv3", "post.conv4"}
    for name, (c_in, taps, n_out) in {"enc.pre0": (5816, 1, 600), "enc.conv1": (600, 5, 600),
                                      "enc.lstm_ih": (600, 1, 2400), "post.conv0": (80, 5, 512),
                                      "post.conv4": (512, 5, 80)}.items():
        w = tw[name]
        assert w["c_pad"] % 64 == 0 and w["n_pad"] % 64 == 0 and w["c_pad"] >= c_in and w["n_pad"] >= n_out
        assert w["hi"].shape == (w["n_pad"], taps * w["c_pad"]) and w["hi"].dtype == torch.float16
        both = (w["hi"].float() + w["lo"].float()).view(w["n_pad"], taps, w["c_pad"])
        ref = packed.view(name + "_w")[:, :n_out].reshape(taps, c_in, n_out).permute(2, 0, 1)
        assert (both[:n_out, :, :c_in] - ref).abs().max().item() <= 2e-6 * ref.abs().max().item()
        assert both[n_out:].abs().max().item() == 0.0 if w["n_pad"] > n_out else True
        assert both[:, :, c_in:].abs().max().item() == 0.0 if w["c_pad"] > c_in else True
        if w["bias"] is not None:
            assert w["bias"].shape == (w["n_pad"],) and w["bias"][n_out:].abs().sum().item() == 0.0
    assert tw["enc.pre0"]["bias"] is None and packed.tc_weights() is tw        # cached


def test_out_of_range_weights_fail_loudly():
    """The half-precision operand pairs of the recurrent kernels hold w * 2^8: a checkpoint whose weights do not
    fit is refused at packing time instead of producing infinities on the GPU."""
    from fac_via_ppg_b200 import _ext
    from fac_via_ppg_b200.packing import PackedTacotron
    sd = dict(synth.tacotron_state())
    sd["decoder.attention_rnn.weight_hh"] = sd["decoder.attention_rnn.weight_hh"] * 1e4
    with pytest.raises(_ext.FacError):
        PackedTacotron.from_state(sd, synth.TACOTRON_HPARAMS, "cpu")


def test_pack_cache_round_trip_and_invalidation(tmp_path, monkeypatch):
    """f3: the packed buffer is cached on disk under a hash of the weights; a hit skips the repacking, a changed
    weight misses."""
    cfg = synth.WAVEGLOW_CONFIG_SMALL
    sd = synth.waveglow_state(cfg=cfg, seed=5)
    first = PackedWaveGlow.from_state(sd, cfg, "cpu", cache_dir=str(tmp_path))
    files = os.listdir(tmp_path)
    assert len(files) == 1 and files[0].startswith("waveglow-") and not hasattr(first, "from_cache")
    monkeypatch.setattr(PackedWaveGlow, "load_state", lambda self, sd: (_ for _ in ()).throw(AssertionError("repacked")))
    again = PackedWaveGlow.from_state(sd, cfg, "cpu", cache_dir=str(tmp_path))
    assert again.from_cache.endswith(files[0]) and torch.equal(again.flat, first.flat)
    monkeypatch.setenv("FAC_PACK_CACHE", str(tmp_path))                 # the environment default
    assert torch.equal(PackedWaveGlow.from_state(sd, cfg, "cpu").flat, first.flat)
    monkeypatch.undo()
    changed = dict(sd)
    changed["WN.1.end.bias"] = sd["WN.1.end.bias"] + 1e-3
    other = PackedWaveGlow.from_state(changed, cfg, "cpu", cache_dir=str(tmp_path))
    assert len(os.listdir(tmp_path)) == 2 and not torch.equal(other.flat, first.flat)
    # a truncated entry is repacked, not trusted
    path = os.path.join(tmp_path, files[0])
    open(path, "wb").write(b"garbage")
    assert torch.equal(PackedWaveGlow.from_state(sd, cfg, "cpu", cache_dir=str(tmp_path)).flat, first.flat)


def test_old_format_checkpoint_is_converted_like_the_reference():
    """f3: res_layers / skip_layers checkpoints (reference src/waveglow/convert_model.py:43-70) are merged into
    res_skip_layers; the converted model's plain weights equal a directly built current-format model."""
    from fac_via_ppg_b200.waveglow.convert_model import _check_model_old_version, update_model
    from fac_via_ppg_b200.waveglow.glow import WaveGlow
    cfg = synth.WAVEGLOW_CONFIG_SMALL
    sd = synth.waveglow_state(cfg=cfg, seed=6)
    new = WaveGlow.remove_weightnorm(WaveGlow(**cfg))
    new.load_state_dict(sd)
    # synthesize the old format from the same weights: split every res_skip conv into weight-normed res / skip convs
    old = WaveGlow(**cfg)
    old.load_state_dict(WaveGlow(**cfg).state_dict())
    old.upsample.load_state_dict(new.upsample.state_dict())
    C = cfg["WN_config"]["n_channels"]
    wnorm = torch.nn.utils.weight_norm
    for k, wn in enumerate(old.WN):
        ref = new.WN[k]
        old.convinv[k].load_state_dict(new.convinv[k].state_dict())
        wn.end.load_state_dict(ref.end.state_dict())
        for name in ("start",):
            conv = torch.nn.Conv1d(ref.start.in_channels, C, 1)
            conv.load_state_dict(ref.start.state_dict())
            wn.start = wnorm(conv, name="weight")
        for lst in ("in_layers", "cond_layers"):
            for i, src in enumerate(getattr(ref, lst)):
                conv = torch.nn.Conv1d(src.in_channels, src.out_channels, src.kernel_size[0], dilation=src.dilation[0],
                                       padding=src.padding[0])
                conv.load_state_dict(src.state_dict())
                getattr(wn, lst)[i] = wnorm(conv, name="weight")
        wn.res_layers, wn.skip_layers = torch.nn.ModuleList(), torch.nn.ModuleList()
        for i, src in enumerate(ref.res_skip_layers):
            last = i == wn.n_layers - 1
            w, b = src.weight.detach(), src.bias.detach()
            if not last:
                res = torch.nn.Conv1d(C, C, 1)
                res.weight.data, res.bias.data = w[:C].clone(), b[:C].clone()
                wn.res_layers.append(wnorm(res, name="weight"))
            skip = torch.nn.Conv1d(C, C, 1)
            skip.weight.data, skip.bias.data = (w if last else w[C:]).clone(), (b if last else b[C:]).clone()
            wn.skip_layers.append(wnorm(skip, name="weight"))
        del wn.res_skip_layers
    assert _check_model_old_version(old) and not _check_model_old_version(new)
    assert update_model(new) is new
    conv = update_model(old)
    assert _check_model_old_version(old) and not _check_model_old_version(conv)       # the input is left alone
    conv = WaveGlow.remove_weightnorm(conv)
    got, want = conv.plain_state(), new.plain_state()
    assert set(got) == set(want)
    for name in want:
        assert (got[name] - want[name]).abs().max().item() <= 1e-6, name
