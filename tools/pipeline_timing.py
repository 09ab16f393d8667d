"""End-to-end PPG -> Mel -> WaveGlow timing at BASELINE.json configs[2] (32 x 5 s, bf16 vocoder) and
configs[0]-like single-utterance latency.  Usage (GPU box): python tools/pipeline_timing.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fac_via_ppg_b200 import synth  # noqa: E402
from fac_via_ppg_b200.common.hparams import create_hparams_stage  # noqa: E402
from fac_via_ppg_b200.common.model import Tacotron2  # noqa: E402
from fac_via_ppg_b200.waveglow.glow import WaveGlow  # noqa: E402

taco = Tacotron2(create_hparams_stage())
taco.load_state_dict(synth.tacotron_state())
taco = taco.cuda().eval()
taco.return_alignments = False
wg = WaveGlow.remove_weightnorm(WaveGlow(**synth.WAVEGLOW_CONFIG))
wg.load_state_dict(synth.waveglow_state())
wg = wg.cuda().eval()

for name, B, seconds, precision in (("configs[2] 32 x 5 s bf16", 32, 5.0, "bf16"), ("32 x 5 s bf16x3", 32, 5.0, "bf16x3"),
                                    ("1 x 2 s bf16x3 (CLI-like latency)", 1, 2.0, "bf16x3"),
                                    ("configs[4] shard 8 x 60 s bf16x3", 8, 60.0, "bf16x3")):
    T = synth.frames_for_seconds(seconds)
    taco.decoder.gate_threshold, taco.decoder.max_decoder_steps = 2.0, T
    wg.set_precision(precision)
    ppg = synth.synthetic_ppg(B, T).cuda()
    best = None
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mel = taco.inference(ppg)[1]
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        wav = wg.infer(mel.clamp(-11.5, 2.0).contiguous(), 0.6)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        if best is None or t2 - t0 < best[0]:
            best = (t2 - t0, t1 - t0, t2 - t1)
    n = wav.numel()
    print("%-36s total %.1f ms (ppg->mel %.1f, mel->wav %.1f) | %.2f M samples/s | RTF %.0fx" %
          (name, best[0] * 1e3, best[1] * 1e3, best[2] * 1e3, n / best[0] / 1e6, n / best[0] / 22050))
    del ppg, mel, wav
    torch.cuda.empty_cache()
