import torch, sys
sys.path.insert(0, '.')
from fac_via_ppg_b200 import synth
from fac_via_ppg_b200.common.hparams import create_hparams_stage
from fac_via_ppg_b200.common.model import Tacotron2
m = Tacotron2(create_hparams_stage()); m.load_state_dict(synth.tacotron_state()); m = m.cuda().eval()
for B, T in ((3, 24), (12, 16)):
    m.decoder.gate_threshold, m.decoder.max_decoder_steps = 2.0, T
    out = m.inference(synth.synthetic_ppg(B, T).cuda(), input_lengths=[T] + [T - 3] * (B - 1))
    torch.cuda.synchronize()
    print("ok", B, T, float(out[1].abs().max()))
