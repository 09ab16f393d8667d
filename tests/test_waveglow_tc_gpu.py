"""GPU parity tests for the tensor-core (tcgen05 + TMA) WaveGlow path, through the C ABI.

bf16x3 (split-bf16, fp32 accumulate) is held to the north-star tolerance for the fp32 config
(<= 1e-4 RMS on the waveform vs the unmodified fp32 reference); plain bf16 is the precision
BASELINE configs[2] names and is held to a stated looser bound."""
import ctypes as C
import os

import pytest
import torch

from fac_via_ppg_b200 import _ext, synth
from fac_via_ppg_b200.waveglow.glow import WaveGlow
from oracle import waveglow_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"
RMS_TOL = {"bf16x3": 1e-4, "bf16": 3e-2}


def rms(a, b):
    return (a.double().cpu() - b.double().cpu()).pow(2).mean().sqrt().item()


_models = {}


def build_model(cfg, precision, **state_kwargs):
    key = (str(cfg), precision, tuple(sorted(state_kwargs.items())))
    if key not in _models:
        model = WaveGlow.remove_weightnorm(WaveGlow(**cfg))
        model.load_state_dict(synth.waveglow_state(cfg=cfg, **state_kwargs), strict=True)
        _models[key] = model.to(DEV).eval().set_precision(precision)
    return _models[key]


@pytest.mark.parametrize("cfg_name,cols", [("small", 70), ("small", 300), ("full", 200), ("full", 700)])
@pytest.mark.parametrize("precision,fused", [("bf16x3", True), ("bf16x3", False), ("bf16", True), ("bf16", False)])
def test_tc_layer_matches_fp32_layer(cfg_name, cols, precision, fused):
    """One WN layer (flow 0; dilation 1 and 2/8, and the last layer, which has no residual output) on the tensor
    cores vs the exact-fp32 kernels on identical inputs: gated activations, residual stream (kept as a bf16
    hi+lo pair) and the collapsed skip path (out8 += W_end W_skip acts must equal W_end applied to the fp32 skip).
    fused = the one-launch form of csrc/waveglow_fused.cu (the residual stream moves from x to x2 or back,
    depending on the layer's parity); otherwise the two-launch form that updates x in place."""
    cfg = synth.WAVEGLOW_CONFIG_SMALL if cfg_name == "small" else synth.WAVEGLOW_CONFIG
    model = build_model(cfg, precision)
    lib, packed = _ext.load(), model.packed()
    nsplit = model._nsplit()
    B, Cn, n_cond = 2, cfg["WN_config"]["n_channels"], 640
    n_half = synth.flow_channels(cfg)[0][1]
    end_w = packed.layout.view(packed.flat, "0.end_w")                         # (2*n_half, C)
    g = torch.Generator().manual_seed(cols)
    x0 = torch.randn(B, cols, Cn, generator=g).to(DEV)
    spect = (torch.randn(B, cols, n_cond, generator=g) * 2).to(DEV)
    skip0 = torch.randn(B, cols, Cn, generator=g).to(DEV)
    out8_0 = torch.randn(B, cols, 8, generator=g).to(DEV)
    st = _ext.current_stream()
    tol = 2e-4 if precision == "bf16x3" else 8e-2
    hi = lambda t: t.to(torch.bfloat16)                                   # noqa: E731
    lo = lambda t: (t - t.to(torch.bfloat16).float()).to(torch.bfloat16)  # noqa: E731
    n_layers = cfg["WN_config"]["n_layers"]
    for layer in (0, 3 if n_layers > 3 else 1, n_layers - 1):
        last = layer == n_layers - 1
        # exact fp32 reference kernels (skip accumulates on top of skip0 for layer > 0)
        xr, sr, ar = x0.clone(), skip0.clone(), torch.empty_like(x0)
        ws = _ext.WgWorkspace(spect.data_ptr(), xr.data_ptr(), ar.data_ptr(), sr.data_ptr())
        _ext.check(lib.fac_wn_layer_f32(C.byref(packed.cmodel), 0, layer, C.byref(ws), B, cols, st), "f32 layer")
        skip_delta = sr - skip0 if layer > 0 else sr
        # the tensor-core path folds the skip biases into out_bias, so out8 carries none
        b_rs = packed.layout.view(packed.flat, f"0.{layer}.res_skip_b")
        skip_delta = skip_delta - (b_rs[:Cn] if last else b_rs[Cn:2 * Cn])
        out8_ref = skip_delta @ end_w.t() + (out8_0[..., : 2 * n_half] if layer > 0 else 0)
        # tensor-core kernels
        s_hi, s_lo = hi(spect), lo(spect)
        a_hi, a_lo = torch.zeros_like(s_hi[..., :Cn]), torch.zeros_like(s_hi[..., :Cn])
        out8 = out8_0.clone()
        if fused:        # layer i reads x (i even) / x2 (i odd) and writes the other pair
            src, dst = (hi(x0), lo(x0)), (torch.full_like(a_hi, 7.0), torch.full_like(a_hi, 7.0))
            xa, xb = (src, dst) if layer % 2 == 0 else (dst, src)
            wst = _ext.WgTcWorkspace(None, None, s_hi.data_ptr(), s_lo.data_ptr(), xa[0].data_ptr(), xa[1].data_ptr(),
                                     a_hi.data_ptr(), a_lo.data_ptr(), out8.data_ptr(), xb[0].data_ptr(), xb[1].data_ptr())
            x_hi, x_lo = src if last else dst            # the last layer leaves the stream alone
        else:
            x_hi, x_lo = hi(x0), lo(x0)
            wst = _ext.WgTcWorkspace(None, None, s_hi.data_ptr(), s_lo.data_ptr(), x_hi.data_ptr(), x_lo.data_ptr(),
                                     a_hi.data_ptr(), a_lo.data_ptr(), out8.data_ptr(), None, None)
        before = lib.fac_launch_count()
        rc = lib.fac_wn_layer_tc(C.byref(packed.cmodel), C.byref(packed.tc_weights()), 0, layer, C.byref(wst), B, cols,
                                 nsplit, st)
        _ext.check(rc, "tc layer")
        torch.cuda.synchronize()
        assert lib.fac_launch_count() - before == (1 if fused or last else 2)
        acts = a_hi.float() + (a_lo.float() if nsplit == 2 else 0)
        xt = x_hi.float() + (x_lo.float() if nsplit == 2 else 0)
        if last:
            xr = x0                                      # glow.py:168-169: no residual output
        assert (acts - ar).abs().max().item() <= tol, ("acts", layer)
        assert (xt - xr).abs().max().item() <= tol, ("x", layer)
        assert (out8[..., : 2 * n_half] - out8_ref).abs().max().item() <= tol, ("out8", layer)


@pytest.mark.parametrize("cfg_name,B,cols", [("small", 2, 300), ("full", 1, 130), ("full", 3, 1000)])
def test_flow_step_is_one_launch_and_equals_the_layer_by_layer_sequence(cfg_name, B, cols):
    """fac_waveglow_flow_step_tc (reference glow.py:272-283 for one flow: WN.forward, coupling inverse, invertible
    1x1) as ONE cooperative launch -- start, 8 fused layers chained by per-tile dependency counters, end -- against the same
    step run as start / layer / ... / end launches: identical bits (same kernel code per tile), for the first flow
    (4 + 4 channels) and the last one (2 + 2)."""
    cfg = synth.WAVEGLOW_CONFIG_SMALL if cfg_name == "small" else synth.WAVEGLOW_CONFIG
    model = build_model(cfg, "bf16x3")
    lib, packed = _ext.load(), model.packed()
    Cn, n_cond, L = cfg["WN_config"]["n_channels"], 640, cfg["WN_config"]["n_layers"]
    g = torch.Generator().manual_seed(cols)
    spect = (torch.randn(B, cols, n_cond, generator=g) * 2).to(DEV)
    audio0 = torch.randn(B, cols, 8, generator=g).to(DEV)
    s_hi = spect.to(torch.bfloat16)
    s_lo = (spect - s_hi.float()).to(torch.bfloat16)
    st, m, tcw = _ext.current_stream(), C.byref(packed.cmodel), C.byref(packed.tc_weights())
    b16 = lambda: torch.zeros(B, cols, Cn, device=DEV, dtype=torch.bfloat16)      # noqa: E731

    def workspace(sync):
        bufs = [b16() for _ in range(4)] + [torch.zeros(B, cols, 8, device=DEV), sync]
        ws = _ext.WgTcWorkspace(None, None, s_hi.data_ptr(), s_lo.data_ptr(), bufs[0].data_ptr(), bufs[1].data_ptr(),
                                None, None, bufs[4].data_ptr(), bufs[2].data_ptr(), bufs[3].data_ptr(), _ext.ptr(sync))
        return ws, bufs

    for flow in (cfg["n_flows"] - 1, 0):
        a_one, a_seq = audio0.clone(), audio0.clone()
        ws1, keep1 = workspace(torch.zeros(B * ((cols + 127) // 128), dtype=torch.int32, device=DEV))   # a counter per tile
        before = lib.fac_launch_count()
        _ext.check(lib.fac_waveglow_flow_step_tc(m, tcw, flow, a_one.data_ptr(), C.byref(ws1), B, cols, 2, st), "flow step")
        torch.cuda.synchronize()
        assert lib.fac_launch_count() - before == 1
        ws2, keep2 = workspace(None)                      # no flow_sync: start + one fused launch per layer + end
        before = lib.fac_launch_count()
        _ext.check(lib.fac_waveglow_flow_step_tc(m, tcw, flow, a_seq.data_ptr(), C.byref(ws2), B, cols, 2, st), "flow seq")
        torch.cuda.synchronize()
        assert lib.fac_launch_count() - before == L + 2
        assert torch.isfinite(a_one).all() and not torch.equal(a_one, audio0)
        assert torch.equal(a_one, a_seq), flow
        assert torch.equal(keep1[4], keep2[4])            # out8


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("name", ["waveglow_small_b2_f6.pt", "waveglow_full_b2_f5.pt",
                                  "waveglow_full_b1_f88_sigma0.pt", "waveglow_small_b2_f6_general_convinv.pt",
                                  "waveglow_full_b2_f5_general_convinv.pt"])
def test_tc_infer_matches_reference_golden(golden_dir, name, precision):
    g = torch.load(os.path.join(golden_dir, name))
    model = build_model(g["cfg"], precision, **g.get("state_kwargs", {}))
    mel = synth.synthetic_mel(g["batch"], g["frames"], seed=g["mel_seed"]).to(DEV)
    audio = model.infer(mel, sigma=g["sigma"], noise=[z.to(DEV) for z in g["noise"]])
    assert audio.shape == g["audio"].shape
    err = rms(audio, g["audio"])
    print("%s %s rms %.3e" % (name, precision, err))
    assert err <= RMS_TOL[precision]


def _oracle_window(cfg, sd, mel, noise, row, f0, f1):
    """Oracle waveform of utterance `row` restricted to mel frames [f0, f1): correct wherever the receptive
    field (12 flows x 255 columns + the 7 upsampler taps) stays inside the crop or hits a TRUE utterance edge."""
    ref = waveglow_oracle.waveglow_infer(sd, cfg, mel[row:row + 1, :, f0:f1].cpu(), 0.6,
                                         [z[row:row + 1, :, f0 * 20:f1 * 20].cpu() for z in noise])
    return ref[0]


def test_tc_full_size_whole_tensor_and_oracle_windows():
    """BASELINE configs[1] size (8 x 10 s = 8 x 27 580 columns: 215 full 128-column tiles + a ragged one per
    utterance, dealt to 74 CTA pairs) on the tensor cores.  Size-independent checks:
      (1) the WHOLE (8, 220 640) bf16x3 output vs the exact-fp32 FFMA path on the GPU, per utterance <= 1e-4 RMS;
      (2) oracle windows: the prefix of row 0, a mid-utterance window of row 3 and the ragged tails of rows 5 and 7
          (any window is checkable on the CPU: receptive field 12 x 255 + 160 columns per side);
      (3) batch rows are independent: rows 0 and 6 alone == in the batch, bit for bit."""
    cfg = synth.WAVEGLOW_CONFIG
    sd = synth.waveglow_state(cfg=cfg)
    model = build_model(cfg, "bf16x3")
    B, F = 8, synth.frames_for_seconds(10.0)
    mel = synth.synthetic_mel(B, F, seed=31).to(DEV)
    torch.manual_seed(17)
    noise = model.noise_like_reference(B, F * 20, DEV, torch.float32)
    full = model.infer(mel, 0.6, noise=noise)
    assert full.shape == (B, F * 160) and torch.isfinite(full).all()
    # (1) whole tensor vs the exact-fp32 kernels
    model.set_precision("fp32")
    try:
        exact = model.infer(mel, 0.6, noise=noise)
    finally:
        model.set_precision("bf16x3")
    per_utt = (full.double() - exact.double()).pow(2).mean(dim=1).sqrt()
    worst = (full - exact).abs().max().item()
    print("bf16x3 vs fp32 kernels, 8 x 10 s: rms per utterance max %.3e, max-abs %.3e" % (per_utt.max().item(), worst))
    assert per_utt.max().item() <= 1e-4
    # (2) oracle windows (R = columns a crop edge can influence)
    R = 12 * 255 + 8 * 20
    Tg = F * 20
    checks = []
    ref = _oracle_window(cfg, sd, mel, noise, 0, 0, 240)                   # prefix of row 0
    checks.append(("prefix row 0", full[0, :(240 * 20 - R) * 8], ref[:(240 * 20 - R) * 8]))
    f0, f1 = 560, 960                                                      # middle of row 3 (tiles 87..150)
    ref = _oracle_window(cfg, sd, mel, noise, 3, f0, f1)
    lo, hi = f0 * 20 + R, f1 * 20 - R
    checks.append(("middle row 3", full[3, lo * 8:hi * 8], ref[(lo - f0 * 20) * 8:(hi - f0 * 20) * 8]))
    for row in (5, 7):                                                     # ragged tail (last, partial tile)
        f0 = F - 240
        ref = _oracle_window(cfg, sd, mel, noise, row, f0, F)
        lo = f0 * 20 + R
        checks.append(("tail row %d" % row, full[row, lo * 8:], ref[(lo - f0 * 20) * 8:]))
    for name, got, want in checks:
        assert got.numel() == want.numel() and got.numel() > 8000, name
        err = rms(got, want)
        print("%s: %d samples, rms vs oracle %.3e" % (name, got.numel(), err))
        assert err <= 1e-4, name
    # (3) independence
    for row in (0, 6):
        alone = model.infer(mel[row:row + 1].contiguous(), 0.6, noise=[z[row:row + 1].contiguous() for z in noise])
        assert torch.equal(alone[0], full[row]), row


@pytest.mark.parametrize("mode,per_flow", [(2, 1), (3, 3)])
def test_flow_step_launch_mode_of_infer_matches_the_default(golden_dir, mode, per_flow):
    """WaveGlow.flow_step_launch: infer() with one cooperative launch per flow step (14 launches per call instead of
    122; fused mode 3 keeps start and end as separate kernels: 38) produces the same bits as the default
    one-launch-per-layer mode, and the reference golden within 1e-4."""
    g = torch.load(os.path.join(golden_dir, "waveglow_full_b2_f5.pt"))
    model = build_model(g["cfg"], "bf16x3")
    mel = synth.synthetic_mel(g["batch"], g["frames"], seed=g["mel_seed"]).to(DEV)
    noise = [z.to(DEV) for z in g["noise"]]
    lib = _ext.load()
    base = model.infer(mel, sigma=g["sigma"], noise=noise)
    model.flow_step_launch = True
    try:
        _ext.check(lib.fac_tc_set_fused(mode), "fac_tc_set_fused")
        lib.fac_reset_launch_count()
        flow = model.infer(mel, sigma=g["sigma"], noise=noise)
        assert lib.fac_launch_count() == 1 + 1 + per_flow * g["cfg"]["n_flows"]   # mel split, upsampler (all phases), flows
    finally:
        model.flow_step_launch = False
        lib.fac_tc_set_fused(2)                  # the library default (the module decides through its workspace)
    assert torch.equal(flow, base)
    assert rms(flow, g["audio"]) <= 1e-4


def test_small_inputs_replay_a_cuda_graph():
    """Launch-bound sizes are captured once per shape in a CUDA graph: same bits as the eager path (sigma = 0
    removes the noise), fresh noise on every replay, launches still accounted for."""
    cfg = synth.WAVEGLOW_CONFIG
    model = WaveGlow.remove_weightnorm(WaveGlow(**cfg))
    model.load_state_dict(synth.waveglow_state(cfg=cfg))
    model = model.to(DEV).eval()
    mel = synth.synthetic_mel(2, 9, seed=4).to(DEV)
    model.graph_max_frames = 0
    eager = model.infer(mel, sigma=0.0)
    model.graph_max_frames = 4096
    lib = _ext.load()
    first = model.infer(mel, sigma=0.0)            # captures
    lib.fac_reset_launch_count()
    again = model.infer(mel, sigma=0.0)            # replays
    assert lib.fac_launch_count() >= 122          # mel split + upsampler + 12 x (start + 8 fused layers + end)
    assert torch.equal(first, eager) and torch.equal(again, eager)
    mel2 = synth.synthetic_mel(2, 9, seed=5).to(DEV)
    assert torch.equal(model.infer(mel2, sigma=0.0), model._infer_eager(mel2, 0.0, None))   # new input, same graph
    a, b = model.infer(mel, sigma=0.6), model.infer(mel, sigma=0.6)
    assert torch.isfinite(a).all() and not torch.equal(a, b)                                 # new draws per replay
    assert len(model.packed()._graphs) == 2
    # new weights -> new pack -> the graphs captured against the old buffers are gone with it
    old_pack = model.packed()
    with torch.no_grad():
        model.WN[0].end.bias.add_(0.25)
    changed = model.infer(mel, sigma=0.0)
    assert model.packed() is not old_pack and len(model.packed()._graphs) == 1
    assert not torch.equal(changed, eager) and torch.equal(changed, model._infer_eager(mel, 0.0, None))
