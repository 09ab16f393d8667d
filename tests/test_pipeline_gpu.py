"""End-to-end PPG -> Mel -> WaveGlow at the BASELINE.json config shapes (GPU, through the drop-in API).

configs[0] (1 x 2 s) is small enough for the CPU oracle, so it is a full parity case; the larger
configs are checked through size-independent properties: utterance independence (bit-exact alone vs
batched), causality of the decoder (a longer run reproduces a shorter one as its prefix) and
finiteness."""
import pytest
import torch

from fac_via_ppg_b200 import synth
from fac_via_ppg_b200.common.hparams import create_hparams_stage
from fac_via_ppg_b200.common.model import Tacotron2
from fac_via_ppg_b200.waveglow.glow import WaveGlow
from oracle import tacotron_oracle, waveglow_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def taco():
    m = Tacotron2(create_hparams_stage())
    m.load_state_dict(synth.tacotron_state(), strict=True)
    return m.to(DEV).eval()


@pytest.fixture(scope="module")
def waveglow():
    m = WaveGlow.remove_weightnorm(WaveGlow(**synth.WAVEGLOW_CONFIG))
    m.load_state_dict(synth.waveglow_state(), strict=True)
    return m.to(DEV).eval()


def rms(a, b):
    return (a.double().cpu() - b.double().cpu()).pow(2).mean().sqrt().item()


def test_config0_one_utterance_two_seconds_matches_cpu_reference_port(taco, waveglow):
    """BASELINE configs[0]: 1 x 2 s (276 frames) PPG -> Mel -> WaveGlow; the CPU side is the oracle."""
    T = synth.frames_for_seconds(2.0)
    taco.decoder.gate_threshold, taco.decoder.max_decoder_steps = 2.0, T
    ppg = synth.synthetic_ppg(1, T, seed=101)
    torch.manual_seed(3)
    masks = tacotron_oracle.record_dropout_tape(1, T, T)
    ref = tacotron_oracle.tacotron_inference(synth.tacotron_state(), synth.TACOTRON_HPARAMS, ppg, masks, 2.0, T)
    out = taco.inference(ppg.to(DEV), dropout_tape=masks)
    assert (out[0].cpu() - ref[0]).abs().max().item() <= 1e-3          # mel, north-star tolerance
    assert (out[1].cpu() - ref[1]).abs().max().item() <= 1e-3          # mel_postnet
    # vocoder on the reference's own mel so the 1e-4 RMS waveform bound is about WaveGlow alone
    mel = ref[1].clamp(-11.5, 2.0)
    torch.manual_seed(4)
    noise = waveglow_oracle.draw_noise(synth.WAVEGLOW_CONFIG, 1, T * 20)
    wav_ref = waveglow_oracle.waveglow_infer(synth.waveglow_state(), synth.WAVEGLOW_CONFIG, mel, 0.6, noise)
    for precision, tol in (("bf16x3", 1e-4), ("fp32", 1e-5)):
        wav = waveglow.set_precision(precision).infer(mel.to(DEV), 0.6, noise=[z.to(DEV) for z in noise])
        assert wav.shape == (1, T * 160)
        assert rms(wav, wav_ref) <= tol, precision
    waveglow.set_precision("bf16x3")
    # and the chained GPU pipeline end to end (mel from the GPU decoder)
    wav_chain = waveglow.infer(out[1].clamp(-11.5, 2.0).contiguous(), 0.6, noise=[z.to(DEV) for z in noise])
    assert torch.isfinite(wav_chain).all() and rms(wav_chain, wav_ref) <= 5e-2


def test_config2_shape_batch32_five_seconds_bf16_properties(taco, waveglow):
    """BASELINE configs[2] shape: 32 x 5 s, bf16 vocoder.  Every utterance of the batch equals the
    same utterance processed alone, bit for bit, in both models."""
    B, T = 32, synth.frames_for_seconds(5.0)
    taco.decoder.gate_threshold, taco.decoder.max_decoder_steps = 2.0, T
    taco.return_alignments = False
    ppg = synth.synthetic_ppg(B, T, seed=55).to(DEV)
    torch.manual_seed(6)
    masks = tacotron_oracle.record_dropout_tape(B, T, T)
    try:
        out = taco.inference(ppg, dropout_tape=masks)
        k = 17
        alone = taco.inference(ppg[k:k + 1].contiguous(), dropout_tape=[m[k:k + 1] for m in masks])
    finally:
        taco.return_alignments = True
    assert out[1].shape == (B, 80, T) and torch.isfinite(out[1]).all()
    assert torch.equal(out[1][k], alone[1][0])
    mel = out[1].clamp(-11.5, 2.0).contiguous()
    waveglow.set_precision("bf16")
    try:
        torch.manual_seed(8)
        noise = waveglow.noise_like_reference(B, T * 20, DEV, torch.float32)
        wav = waveglow.infer(mel, 0.6, noise=noise)
        wav_k = waveglow.infer(mel[k:k + 1].contiguous(), 0.6, noise=[z[k:k + 1].contiguous() for z in noise])
    finally:
        waveglow.set_precision("bf16x3")
    assert wav.shape == (B, T * 160) and torch.isfinite(wav).all()
    assert torch.equal(wav[k], wav_k[0])


def test_config4_shape_sixty_seconds_long_form(taco, waveglow):
    """BASELINE configs[4] shape (60 s utterances, 8269 decoder steps -- beyond the reference's 1000-step
    limit): decoding is causal, so the long run must reproduce a 300-step run as its prefix."""
    B, T = 2, synth.frames_for_seconds(60.0)
    ppg = synth.synthetic_ppg(B, T, seed=77).to(DEV)
    taco.rng_mode = "fast"
    taco.return_alignments = False
    try:
        torch.manual_seed(9)
        e0, e1, dec = taco._dropout_masks(B, T, T, DEV, None)
        tape = [e0, e1] + [dec[t, j].float() * 2 for t in range(300) for j in range(2)]
        taco.decoder.gate_threshold, taco.decoder.max_decoder_steps = 2.0, 300
        short = taco.inference(ppg, dropout_tape=tape)
        taco.decoder.max_decoder_steps = T
        full_tape = [e0, e1] + list((dec.float() * 2).view(-1, B, 300))
        long = taco.inference(ppg, dropout_tape=full_tape)
    finally:
        taco.return_alignments = True
    assert long[0].shape == (B, 80, T) and torch.isfinite(long[1]).all()
    assert torch.equal(long[0][:, :, :300], short[0])
    wav = waveglow.infer(long[1][:1].clamp(-11.5, 2.0).contiguous(), 0.6)
    assert wav.shape == (1, T * 160) and torch.isfinite(wav).all()


@pytest.mark.parametrize("t_in,steps", [(1, 4), (3, 9), (25, 40)])
def test_tiny_and_ragged_tacotron_inputs(taco, t_in, steps):
    """Inputs shorter than the attention window / the location kernel, and decodes that run past the end
    of the input (the window then collapses onto the last frame, src/common/utils.py:65-69)."""
    taco.decoder.gate_threshold, taco.decoder.max_decoder_steps = 2.0, steps
    ppg = synth.synthetic_ppg(2, t_in, seed=t_in)
    torch.manual_seed(t_in)
    masks = tacotron_oracle.record_dropout_tape(2, t_in, steps)
    ref = tacotron_oracle.tacotron_inference(synth.tacotron_state(), synth.TACOTRON_HPARAMS, ppg, masks, 2.0, steps)
    out = taco.inference(ppg.to(DEV), dropout_tape=masks)
    for name, a, b in zip(("mel", "mel_post", "gate", "align"), out, ref):
        assert a.shape == b.shape, name
        assert (a.cpu() - b).abs().max().item() <= 1e-3, name


@pytest.mark.parametrize("batch,frames", [(1, 1), (3, 2), (5, 7)])
@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_tiny_and_odd_waveglow_inputs(waveglow, batch, frames, precision):
    """A single frame, odd batches, column counts far below one 128-row tile (and odd tile counts for
    the CTA-pair kernel)."""
    mel = synth.synthetic_mel(batch, frames, seed=frames)
    torch.manual_seed(frames)
    noise = waveglow_oracle.draw_noise(synth.WAVEGLOW_CONFIG, batch, frames * 20)
    ref = waveglow_oracle.waveglow_infer(synth.waveglow_state(), synth.WAVEGLOW_CONFIG, mel, 0.6, noise)
    try:
        out = waveglow.set_precision(precision).infer(mel.to(DEV), 0.6, noise=[z.to(DEV) for z in noise])
    finally:
        waveglow.set_precision("bf16x3")
    assert out.shape == ref.shape
    assert rms(out, ref) <= (1e-5 if precision == "fp32" else 1e-4)
