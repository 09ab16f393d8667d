"""Where does the one-launch-per-flow form of the fused WN kernel lose time against one launch per layer?
Cycle counters (PROF instantiations) of one whole flow step in both forms, BASELINE configs[1] shape.
Usage (GPU box): python tools/micro/flow_vs_layers.py"""
import ctypes as C
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
lib = _ext.load()
shape = [int(v) for v in os.environ.get("FAC_BREAKDOWN_SHAPE", "8,1379").split(",")]
mel = synth.synthetic_mel(shape[0], shape[1]).cuda()
names = ["prod_wait_empty", "mma_wait_tmem0", "mma_wait_full", "mma_wait_acts", "mma_wait_tmem1", "mma_total",
         "epi_wait_full0", "epi_drain", "epi_wait_acts_free", "epi_busy", "epi_wait_full1", "eg_busy", "epi_total"]
m.set_precision("bf16x3")
m.flow_step_launch = True
packed = m.packed()
tcw = packed.tc_weights()
bufs, B, F, Tg = m._alloc_io(mel, 0.6, None)
st, mm, ws, ns = _ext.current_stream(), C.byref(packed.cmodel), C.byref(bufs["ws"]), m._nsplit()
lib.fac_waveglow_tc_prepare_spect(mm, C.byref(tcw), ws, bufs["mel_cl"].data_ptr(), B, F, ns, st)
flow, L = 5, cfg["WN_config"]["n_layers"]


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def show(tag, ms, p):
    lead = p[0::2]
    print("%s: %.3f ms" % (tag, ms))
    print("   producer/issuer (pair leaders):", " ".join("%s=%.0fk" % (n, lead[:, i].mean().item() / 1e3) for i, n in enumerate(names[:6])))
    print("   epilogue (warp 4 of every CTA):", " ".join("%s=%.0fk" % (n, p[:, 6 + i].mean().item() / 1e3) for i, n in enumerate(names[6:])))


for rep in range(2):
    # ---- one launch per layer (+ start, end)
    lib.fac_tc_set_fused(1)
    total = torch.zeros(148, 16, dtype=torch.float64)
    ms = timed(lambda: lib.fac_wn_start_tc(mm, flow, bufs["audio"].data_ptr(), ws, B, Tg, ns, st))
    for i in range(L):
        prof = torch.zeros(2 * 256 * 8, dtype=torch.int64, device="cuda")
        lib.fac_tc_set_profile_buffer(prof.data_ptr())
        ms += timed(lambda: lib.fac_wn_layer_tc(mm, C.byref(tcw), flow, i, ws, B, Tg, ns, st))
        lib.fac_tc_set_profile_buffer(None)
        total += prof.view(256, 16)[:148].double().cpu()
    ms += timed(lambda: lib.fac_wn_end_tc(mm, C.byref(tcw), flow, bufs["out8"].data_ptr(), bufs["audio"].data_ptr(), B, Tg, st))
    show("per-layer launches, one flow step (instrumented kernels, sum of start + %d layers + end)" % L, ms, total)
    # ---- one launch per flow step
    lib.fac_tc_set_fused(2)
    prof = torch.zeros(2 * 256 * 8, dtype=torch.int64, device="cuda")
    lib.fac_tc_set_profile_buffer(prof.data_ptr())
    ms = timed(lambda: lib.fac_waveglow_flow_step_tc(mm, C.byref(tcw), flow, bufs["audio"].data_ptr(), ws, B, Tg, ns, st))
    lib.fac_tc_set_profile_buffer(None)
    show("one launch per flow step (instrumented kernel)", ms, prof.view(256, 16)[:148].double().cpu())
