"""Latency of one short utterance through WaveGlow.infer (launch-bound sizes): device time per call (CUDA events,
mean of 20 after 3 warm-ups) and library launches per call.  Usage (GPU box): python tools/latency_small.py [seconds ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fac_via_ppg_b200 import _ext, synth  # noqa: E402
from fac_via_ppg_b200.waveglow.glow import WaveGlow  # noqa: E402

m = WaveGlow.remove_weightnorm(WaveGlow(**synth.WAVEGLOW_CONFIG))
m.load_state_dict(synth.waveglow_state())
m = m.cuda().eval()
lib = _ext.load()
for seconds in [float(a) for a in sys.argv[1:]] or [2.0, 0.5, 5.0]:
    F = synth.frames_for_seconds(seconds)
    mel = synth.synthetic_mel(1, F).cuda()
    for mode in ("graph", "eager"):
        m.graph_max_frames = 4096 if mode == "graph" else 0
        for _ in range(3):
            m.infer(mel, 0.6)
        lib.fac_reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            m.infer(mel, 0.6)
        e1.record()
        torch.cuda.synchronize()
        print("1 x %.1f s (%d frames) %-5s: %.3f ms per infer, %d library launches, RTF %.0fx" %
              (seconds, F, mode, e0.elapsed_time(e1) / 20, lib.fac_launch_count() // 20,
               F * 160 / 22050 / (e0.elapsed_time(e1) / 20e3)))
