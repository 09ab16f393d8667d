"""TEST INFRASTRUCTURE ONLY -- regenerates tests/golden/*.pt from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python -m oracle.make_golden
The reference modules (src/waveglow/glow.py, src/common/model.py) are imported
through oracle/ref_shim.py, loaded with the seeded synthetic weights of
fac_via_ppg_b200/synth.py, and executed on CPU fp32.  Each fixture stores the
inputs that cannot be regenerated bit-exactly elsewhere (noise draws, dropout
masks) and the reference outputs; weights and mel/PPG inputs are regenerated
from their seeds by the tests.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fac_via_ppg_b200 import synth  # noqa: E402
from oracle import ref_shim, tacotron_oracle, waveglow_oracle  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _replay_normal(noise):
    """Patch Tensor.normal_ so the reference's in-place draws return the recorded tape."""
    tape = list(noise)
    orig = torch.Tensor.normal_

    def fake(self, *a, **k):
        z = tape.pop(0)
        assert tuple(z.shape) == tuple(self.shape), (z.shape, self.shape)
        return self.copy_(z)
    return orig, fake


def waveglow_case(name, cfg, batch, frames, sigma, seed, **state_kwargs):
    sd = synth.waveglow_state(cfg=cfg, **state_kwargs)
    model = ref_shim.reference_waveglow(sd, cfg)
    mel = synth.synthetic_mel(batch, frames, seed=seed)
    torch.manual_seed(seed)
    noise = waveglow_oracle.draw_noise(cfg, batch, frames * cfg["hop_length"] // cfg["n_group"])
    orig, fake = _replay_normal(noise)
    torch.Tensor.normal_ = fake
    try:
        with torch.no_grad():
            audio = model.infer(mel, sigma=sigma)
    finally:
        torch.Tensor.normal_ = orig
    torch.save({"cfg": cfg, "batch": batch, "frames": frames, "sigma": sigma, "mel_seed": seed,
                "state_kwargs": state_kwargs, "noise": noise, "audio": audio.clone(), "generator": "reference src/waveglow/glow.py WaveGlow.infer"},
               os.path.join(GOLDEN, name))
    print(name, tuple(audio.shape), float(audio.std()))


def tacotron_case(name, t_in, seed):
    sd = synth.tacotron_state()
    model = ref_shim.reference_tacotron(sd)
    model.decoder.gate_threshold = 2.0         # never fires: forced length
    model.decoder.max_decoder_steps = t_in
    ppg = synth.synthetic_ppg(1, t_in, seed=seed)
    torch.manual_seed(seed)
    masks = tacotron_oracle.record_dropout_tape(1, t_in, t_in)
    torch.manual_seed(seed)                     # the reference now draws the same masks itself
    with torch.no_grad():
        mel, mel_post, gate, align = model.inference(ppg)
    torch.save({"t_in": t_in, "ppg_seed": seed, "masks": [m.to(torch.uint8) for m in masks],
                "mel": mel.clone(), "mel_post": mel_post.clone(), "gate": gate.clone(), "align": align.clone(),
                "generator": "reference src/common/model.py Tacotron2.inference"},
               os.path.join(GOLDEN, name))
    print(name, tuple(mel.shape), float(mel.std()))


def main():
    if not ref_shim.available():
        raise SystemExit("needs the reference tree at %s" % ref_shim.REFERENCE_ROOT)
    os.makedirs(GOLDEN, exist_ok=True)
    waveglow_case("waveglow_small_b2_f6.pt", synth.WAVEGLOW_CONFIG_SMALL, 2, 6, 0.6, 11)
    waveglow_case("waveglow_full_b2_f5.pt", synth.WAVEGLOW_CONFIG, 2, 5, 0.6, 12)
    waveglow_case("waveglow_full_b1_f88_sigma0.pt", synth.WAVEGLOW_CONFIG, 1, 88, 0.0, 13)  # the Denoiser call
    # trained-checkpoint-like invertible 1x1 convs: W^-1 != W^T (glow.py:82-97)
    waveglow_case("waveglow_small_b2_f6_general_convinv.pt", synth.WAVEGLOW_CONFIG_SMALL, 2, 6, 0.6, 14,
                  convinv="general")
    waveglow_case("waveglow_full_b2_f5_general_convinv.pt", synth.WAVEGLOW_CONFIG, 2, 5, 0.6, 15, convinv="general")
    tacotron_case("tacotron_b1_t24.pt", 24, 21)


if __name__ == "__main__":
    main()
