"""compute-sanitizer target: one small WaveGlow.infer in each launch mode (per layer with programmatic dependent
launch, one launch per flow step with tile counters, three launches per flow step)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fac_via_ppg_b200 import _ext, synth  # noqa: E402
from fac_via_ppg_b200.waveglow.glow import WaveGlow  # noqa: E402

cfg = synth.WAVEGLOW_CONFIG
m = WaveGlow.remove_weightnorm(WaveGlow(**cfg))
m.load_state_dict(synth.waveglow_state(cfg=cfg))
m = m.cuda().eval()
m.graph_max_frames = 0
lib = _ext.load()
mel = synth.synthetic_mel(2, 40).cuda()
outs = []
for mode, flow in ((1, False), (2, True), (3, True)):
    lib.fac_tc_set_fused(mode)
    m.flow_step_launch = flow
    torch.manual_seed(0)
    outs.append(m.infer(mel, sigma=0.6).clone())
    torch.cuda.synchronize()
    print("mode", mode, "ok", float(outs[-1].abs().max()))
print("identical:", torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]))
