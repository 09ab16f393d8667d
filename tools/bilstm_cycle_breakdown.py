"""Cycle breakdown of the encoder BiLSTM cluster kernel (fac_lstm_set_profile_buffer).
Usage (GPU box): python tools/bilstm_cycle_breakdown.py [B T ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fac_via_ppg_b200 import _ext  # noqa: E402

lib = _ext.load()
args = [int(a) for a in sys.argv[1:]] or [1, 690, 7, 690, 8, 690, 16, 690, 28, 690, 32, 690]
for B, T in zip(args[0::2], args[1::2]):
    xp = torch.randn(B, T, 2400, device="cuda") * 0.1
    w = torch.randn(2, 1200, 300, device="cuda") * 0.05
    out = torch.empty(B, T, 600, device="cuda")
    prof = torch.zeros(512 * 4, dtype=torch.int64, device="cuda")
    for it in range(2):
        lib.fac_lstm_set_profile_buffer(prof.data_ptr() if it else None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _ext.check(lib.fac_lstm_bidir_f32(xp.data_ptr(), w.data_ptr(), out.data_ptr(), B, T, 300, _ext.current_stream()), "lstm")
        e1.record()
        torch.cuda.synchronize()
    lib.fac_lstm_set_profile_buffer(None)
    p = prof.view(512, 4)[:16].double().cpu() / T
    print("B=%d T=%d: %.2f us/step | cycles/step CTA0: mma %d, cell+DSMEM %d, after hand-over %d | CTA7: %d %d %d" %
          (B, T, e0.elapsed_time(e1) * 1e3 / T, *p[0, :3], *p[7, :3]))
