"""Per-role cycle breakdown of the tensor-core WN kernel (wn_gemm_tc_kernel) on one middle layer of
BASELINE configs[1], read from the kernel's optional clock64 counters (fac_tc_set_profile_buffer).
Usage (GPU box): python tools/tc_cycle_breakdown.py [precision ...]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fac_via_ppg_b200 import _ext, synth  # noqa: E402
from fac_via_ppg_b200.waveglow.glow import WaveGlow  # noqa: E402

cfg = synth.WAVEGLOW_CONFIG
m = WaveGlow.remove_weightnorm(WaveGlow(**cfg))
m.load_state_dict(synth.waveglow_state(cfg=cfg))
m = m.cuda().eval()
lib = _ext.load()
mel = synth.synthetic_mel(8, 1379).cuda()
names = ["prod_wait_empty", "mma_wait_tmem", "mma_wait_full", "mma_total", "epi_wait_full", "epi_busy"]
for prec in (sys.argv[1:] or ["bf16x3", "bf16"]):
    m.set_precision(prec)
    packed = m.packed()
    tcw = packed.tc_weights()
    bufs, B, F, Tg = m._alloc_io(mel, 0.6, None)
    st, mm, ws, ns = _ext.current_stream(), C.byref(packed.cmodel), C.byref(bufs["ws"]), m._nsplit()
    lib.fac_waveglow_tc_prepare_spect(mm, C.byref(tcw), ws, bufs["mel_cl"].data_ptr(), B, F, ns, st)
    lib.fac_wn_start_tc(mm, 5, bufs["audio"].data_ptr(), ws, B, Tg, ns, st)
    for i in range(3):
        lib.fac_wn_layer_tc(mm, C.byref(tcw), 5, i, ws, B, Tg, ns, st)
    prof = torch.zeros(2 * 256 * 8, dtype=torch.int64, device="cuda")
    lib.fac_tc_set_profile_buffer(prof.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.fac_wn_layer_tc(mm, C.byref(tcw), 5, 3, ws, B, Tg, ns, st)
    e1.record()
    torch.cuda.synchronize()
    lib.fac_tc_set_profile_buffer(None)
    p = prof.view(2, 256, 8)[:, :148].double().cpu()
    print(prec, "layer ms %.3f" % e0.elapsed_time(e1))
    for g in range(2):
        print("  G%d" % (g + 1), " ".join("%s=%.0fk" % (n, p[g, :, i].mean().item() / 1e3) for i, n in enumerate(names)))
