"""Per-role cycle breakdown of the tensor-core WN layer kernels on one middle layer of BASELINE configs[1], read
from the kernels' optional clock64 counters (fac_tc_set_profile_buffer): the fused one-launch form
(wn_layer_fused_kernel) and the two-launch form (wn_gemm_tc_kernel).
Usage (GPU box): python tools/tc_cycle_breakdown.py [fused|split|bf16 ...]"""
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
shape = [int(v) for v in os.environ.get("FAC_BREAKDOWN_SHAPE", "8,1379").split(",")]      # utterances, frames
mel = synth.synthetic_mel(shape[0], shape[1]).cuda()
names2 = ["prod_wait_empty", "mma_wait_tmem", "mma_wait_full", "mma_total", "epi_wait_full", "epi_busy"]
namesf = ["prod_wait_empty", "mma_wait_tmem0", "mma_wait_full", "mma_wait_acts", "mma_wait_tmem1", "mma_total",
          "epi_wait_full0", "epi_drain", "epi_wait_acts_free", "epi_busy", "epi_wait_full1", "eg_busy", "epi_total"]
for mode in (sys.argv[1:] or ["fused", "split", "bf16"]):
    m.set_precision("bf16" if mode == "bf16" else "bf16x3")
    m.fused_layers = mode == "fused"              # (bf16: the two-launch form; its fused form carries no counters)
    if os.environ.get("FAC_TC_FUSED"):
        lib.fac_tc_set_fused(int(os.environ["FAC_TC_FUSED"]))
    packed = m.packed()
    tcw = packed.tc_weights()
    bufs, B, F, Tg = m._alloc_io(mel, 0.6, None)
    st, mm, ws, ns = _ext.current_stream(), C.byref(packed.cmodel), C.byref(bufs["ws"]), m._nsplit()
    lib.fac_waveglow_tc_prepare_spect(mm, C.byref(tcw), ws, bufs["mel_cl"].data_ptr(), B, F, ns, st)
    lib.fac_wn_start_tc(mm, 5, bufs["audio"].data_ptr(), ws, B, Tg, ns, st)
    for i in range(3):
        lib.fac_wn_layer_tc(mm, C.byref(tcw), 5, i, ws, B, Tg, ns, st)
    for layer in (3, 7) if mode == "fused" else (3,):
        prof = torch.zeros(2 * 256 * 8, dtype=torch.int64, device="cuda")
        lib.fac_tc_set_profile_buffer(prof.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.fac_wn_layer_tc(mm, C.byref(tcw), 5, layer, ws, B, Tg, ns, st)
        e1.record()
        torch.cuda.synchronize()
        lib.fac_tc_set_profile_buffer(None)
        print(mode, "layer %d: %.3f ms" % (layer, e0.elapsed_time(e1)))
        if mode == "fused":
            p = prof.view(256, 16)[:148].double().cpu()
            lead, every = p[0::2], p
            print("   producer/issuer (pair leaders):", " ".join("%s=%.0fk" % (n, lead[:, i].mean().item() / 1e3)
                                                                 for i, n in enumerate(namesf[:6])))
            print("   epilogue (warp 4 of every CTA):", " ".join("%s=%.0fk" % (n, every[:, 6 + i].mean().item() / 1e3)
                                                                 for i, n in enumerate(namesf[6:])))
        else:
            p = prof.view(2, 256, 8)[:, :148].double().cpu()
            for g in range(2):
                print("  G%d" % (g + 1), " ".join("%s=%.0fk" % (n, p[g, :, i].mean().item() / 1e3) for i, n in enumerate(names2)))
