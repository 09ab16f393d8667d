"""Phase timing of Tacotron2.inference (CUDA events) at BASELINE-config shapes.
Usage (GPU box): python tools/tacotron_timing.py [B T ...]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fac_via_ppg_b200 import synth  # noqa: E402
from fac_via_ppg_b200.common.hparams import create_hparams_stage  # noqa: E402
from fac_via_ppg_b200.common.model import Tacotron2  # noqa: E402

m = Tacotron2(create_hparams_stage())
m.load_state_dict(synth.tacotron_state())
m = m.cuda().eval()
m.collect_timing = True
m.return_alignments = False
args = [int(a) for a in sys.argv[1:]] or [1, 276, 32, 690, 8, 2000]
for B, T in zip(args[0::2], args[1::2]):
    m.decoder.gate_threshold, m.decoder.max_decoder_steps = 2.0, T
    ppg = synth.synthetic_ppg(B, T).cuda()
    for _ in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.inference(ppg)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
    tm = m.last_timing
    print("B=%d T=%d wall %.1f ms | encoder %.2f decoder %.2f (%.2f us/step) postnet %.2f | %.0f mel frames/s" %
          (B, T, wall, tm["encoder_ms"], tm["decoder_ms"], tm["decoder_ms"] * 1e3 / T, tm["postnet_ms"],
           B * T / (wall / 1e3)))
