"""fac_via_ppg_b200.pipeline.BatchStream: the synthesis path of reference src/script/generate_synthesis.py:86-98 over a
sequence of host batches, with the upload of batch k + 1 behind the vocoder of batch k."""
import pytest
import torch

from fac_via_ppg_b200 import synth
from fac_via_ppg_b200.common.hparams import create_hparams_stage
from fac_via_ppg_b200.common.model import Tacotron2
from fac_via_ppg_b200.pipeline import BatchStream
from fac_via_ppg_b200.waveglow.denoiser import Denoiser
from fac_via_ppg_b200.waveglow.glow import WaveGlow
from oracle import tacotron_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_batch_stream_equals_the_three_calls_made_one_after_the_other():
    batch, t_in, n_batches = 2, 20, 3
    taco = Tacotron2(create_hparams_stage())
    taco.load_state_dict(synth.tacotron_state(), strict=True)
    taco = taco.to(DEV).eval()
    taco.decoder.gate_threshold, taco.decoder.max_decoder_steps = 2.0, t_in
    cfg = synth.WAVEGLOW_CONFIG
    wg = WaveGlow.remove_weightnorm(WaveGlow(**cfg))
    wg.load_state_dict(synth.waveglow_state(cfg=cfg))
    wg = wg.to(DEV).eval()
    den = Denoiser(wg, mode="zeros")
    host = [synth.synthetic_ppg(batch, t_in, seed=70 + i).pin_memory() for i in range(n_batches)]
    torch.manual_seed(1)
    tapes = [tacotron_oracle.record_dropout_tape(batch, t_in, t_in) for _ in range(n_batches)]

    # reference order of calls, batch by batch (the dropout masks and the vocoder noise are drawn from torch's
    # generator in the same order in both runs)
    def direct():
        torch.manual_seed(5)
        outs = []
        for x in host:
            mel = taco.inference(x.to(DEV))[1]
            wav = wg.infer(mel.clamp(-11.5, 2.0).contiguous(), sigma=0.6)
            outs.append(den(wav, strength=0.005)[:, 0].clone())
        return outs

    want = direct()
    torch.manual_seed(5)
    stream = BatchStream(taco, wg, den, sigma=0.6, denoiser_strength=0.005)
    got = [w.clone() for w in stream.run(iter(host), to_host=True, record_phases=True)]
    assert len(got) == n_batches and stream.phase_events is not None
    for a, b in zip(got, want):
        assert a.device.type == "cpu" and a.shape == b.shape == (batch, t_in * 160)
        assert torch.equal(a, b.cpu())
    on_gpu = list(BatchStream(taco, wg, None).run(iter(host[:1])))
    assert on_gpu[0].is_cuda and on_gpu[0].shape == (batch, t_in * 160)
    with pytest.raises(ValueError):
        list(stream.run(iter([host[0], host[1][:1]])))
    assert list(stream.run(iter([]))) == []
    del tapes
