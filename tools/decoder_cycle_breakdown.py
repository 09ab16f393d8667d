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
names = ["lstm_att", "attention", "lstm_dec", "projection", "prenet1"]
args = [int(a) for a in sys.argv[1:]] or [1, 690, 8, 690, 32, 690]
for B, T in zip(args[0::2], args[1::2]):
    m.decoder.gate_threshold, m.decoder.max_decoder_steps = 2.0, T
    ppg = synth.synthetic_ppg(B, T).cuda()
    m.inference(ppg)
    prof = torch.zeros(256 * 16, dtype=torch.int64, device="cuda")
    lib.fac_taco_set_profile_buffer(prof.data_ptr())
    m.inference(ppg)
    torch.cuda.synchronize()
    lib.fac_taco_set_profile_buffer(None)
    p = prof.view(256, 16)[:148].double().cpu() / T
    us = m.last_timing["decoder_ms"] * 1e3 / T
    print("B=%d T=%d: %.2f us/step, %.0f cycles/step (CTA 0)" % (B, T, us, p[0, 10]))
    for cta in (0, 147):
        print("  CTA %3d: " % cta + "  ".join("%s %d+%d" % (n, p[cta, 2 * i], p[cta, 2 * i + 1]) for i, n in enumerate(names)))
    print("  CTA 147 mat-vec phases (all four): fetch wait %d, arithmetic %d, epilogue %d" % tuple(p[147, 11:14]))
    print("  CTA 147 arithmetic detail: mma section (warp 0) %d, wait for the other warps %d" % tuple(p[147, 14:16]))
    print("  CTA 0 attention critical path: h load + query projection %d, energies %d, softmax + context %d" % tuple(p[0, 11:14]))
    print("  mean   : " + "  ".join("%s %d+%d" % (n, p[:, 2 * i].mean(), p[:, 2 * i + 1].mean()) for i, n in enumerate(names)))
