"""Per-phase cycle breakdown of the persistent decoder kernel (fac_taco_set_profile_buffer).
Usage (GPU box): python tools/decoder_cycle_breakdown.py [B T ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fac_via_ppg_b200 import _ext, synth  # noqa: E402
from fac_via_ppg_b200.common.hparams import create_hparams_stage  # noqa: E402
from fac_via_ppg_b200.common.model import Tacotron2  # noqa: E402

lib = _ext.load()
m = Tacotron2(create_hparams_stage())
m.load_state_dict(synth.tacotron_state())
m = m.cuda().eval()
m.collect_timing, m.return_alignments = True, False
SLOTS = [("early columns (incl. waiting for them)", 2), ("awaited vector (wait + sweep)", 0), ("products + epilogue", 1),
         ("stop count", 12)]
ATT = [("wait h_att", 14), ("query projection", 11), ("energies", 12), ("softmax+context", 13), ("prepare", 4),
       ("wait stop", 9)]
args = [int(a) for a in sys.argv[1:]] or [1, 690, 8, 690, 32, 690]
for B, T in zip(args[0::2], args[1::2]):
    m.decoder.gate_threshold, m.decoder.max_decoder_steps = 2.0, T
    ppg = synth.synthetic_ppg(B, T).cuda()
    m.inference(ppg)
    prof = torch.zeros(256 * 32, dtype=torch.int64, device="cuda")
    lib.fac_taco_set_profile_buffer(prof.data_ptr())
    m.inference(ppg)
    torch.cuda.synchronize()
    lib.fac_taco_set_profile_buffer(None)
    p = prof.view(256, 32)[:148].double().cpu() / T
    us = m.last_timing["decoder_ms"] * 1e3 / T
    n_lstm = (148 - B - 24) // 2
    roles = [("A attention LSTM", B, B + n_lstm), ("D decoder LSTM", B + n_lstm, B + 2 * n_lstm),
             ("P projection", B + 2 * n_lstm, B + 2 * n_lstm + 13), ("Q prenet 1", B + 2 * n_lstm + 13, B + 2 * n_lstm + 24)]
    print("B=%d T=%d: %.2f us/step, %.0f cycles/step (attention CTA 0)" % (B, T, us, p[0, 10]))
    for name, lo, hi in roles:
        print("  %-18s (%2d CTAs, mean): " % (name, hi - lo) + "  ".join("%s %d" % (n, p[lo:hi, i].mean()) for n, i in SLOTS))
    print("  attention CTA 0: " + "  ".join("%s %d" % (n, p[0, i]) for n, i in ATT))
    print("  attention CTA 0, prepare: weights around the window %d, location conv %d, location dense + exp %d, encoder rows %d"
          % tuple(p[0, 16:20]))
