"""Drop-in for the reference ``src/waveglow/glow.py`` (inference side).

Same class names, constructor arguments, attribute / parameter names and
``infer`` signature as the reference, so pickled ``{'model': WaveGlow}``
checkpoints (reference src/script/train_waveglow.py:56-64) and state dicts load
unchanged and the callers in src/common/utils.py:142-152,177-181,
src/waveglow/denoiser.py:44-61 and src/waveglow/inference.py:33-56 keep working.
The modules below are parameter containers only: every FLOP of ``infer`` runs in
the hand-written sm_100a kernels of libfacb200.so through the C ABI of
include/fac_b200.h.  There is no PyTorch or CPU fallback; the training direction
(``forward`` / ``WaveGlowLoss``) is out of scope and raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from fac_via_ppg_b200 import _ext
from fac_via_ppg_b200.packing import PackedWaveGlow


class Invertible1x1Conv(torch.nn.Module):
    """Parameter holder for the invertible 1x1 convolution (reference glow.py:62-102).

    Initialised to a random rotation (orthonormal, det +1) like the reference."""

    def __init__(self, c):
        super().__init__()
        self.conv = torch.nn.Conv1d(c, c, kernel_size=1, bias=False)
        q, _ = torch.linalg.qr(torch.randn(c, c))
        if torch.det(q) < 0:
            q[:, 0].neg_()
        self.conv.weight.data = q.contiguous().view(c, c, 1)

    def forward(self, z, reverse=False):
        raise NotImplementedError("fac_via_ppg_b200: Invertible1x1Conv runs fused inside WaveGlow.infer "
                                  "(fac_wn_end_coupling_f32); the standalone/training call is out of scope")


class WN(torch.nn.Module):
    """Parameter holder for the WaveNet-like coupling network (reference glow.py:105-152)."""

    def __init__(self, n_in_channels, n_mel_channels, n_layers, n_channels, kernel_size):
        super().__init__()
        if kernel_size % 2 != 1 or n_channels % 2 != 0:
            raise ValueError("WN needs an odd kernel size and an even channel count")
        wnorm = torch.nn.utils.weight_norm
        self.n_layers, self.n_channels = n_layers, n_channels
        self.start = wnorm(torch.nn.Conv1d(n_in_channels, n_channels, 1), name="weight")
        self.end = torch.nn.Conv1d(n_channels, 2 * n_in_channels, 1)
        torch.nn.init.zeros_(self.end.weight)
        torch.nn.init.zeros_(self.end.bias)
        self.in_layers = torch.nn.ModuleList()
        self.res_skip_layers = torch.nn.ModuleList()
        self.cond_layers = torch.nn.ModuleList()
        for i in range(n_layers):
            d = 2 ** i
            self.in_layers.append(wnorm(torch.nn.Conv1d(n_channels, 2 * n_channels, kernel_size, dilation=d,
                                                        padding=d * (kernel_size - 1) // 2), name="weight"))
            self.cond_layers.append(wnorm(torch.nn.Conv1d(n_mel_channels, 2 * n_channels, 1), name="weight"))
            width = 2 * n_channels if i < n_layers - 1 else n_channels
            self.res_skip_layers.append(wnorm(torch.nn.Conv1d(n_channels, width, 1), name="weight"))

    def forward(self, forward_input):
        raise NotImplementedError("fac_via_ppg_b200: WN runs inside WaveGlow.infer (fac_wn_layer_f32 ...)")


def _plain_weight(conv):
    """Effective weight of a conv with or without weight-norm (both checkpoint
    flavours reach infer(): generate_synthesis.py:58 vs utils.py:177-181)."""
    if hasattr(conv, "weight_g") and hasattr(conv, "weight_v"):
        v, g = conv.weight_v, conv.weight_g
        return v * (g / v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1))))
    return conv.weight


class WaveGlow(torch.nn.Module):
    """Reference-compatible WaveGlow (reference glow.py:178-303) whose ``infer`` is CUDA-native."""

    def __init__(self, n_mel_channels, hop_length, n_flows, n_group, n_early_every, n_early_size, WN_config):
        super().__init__()
        if n_group % 2 != 0:
            raise ValueError("n_group must be even")
        self.upsample = torch.nn.ConvTranspose1d(n_mel_channels, n_mel_channels, 1024, stride=hop_length)
        self.n_flows, self.n_group = n_flows, n_group
        self.n_early_every, self.n_early_size = n_early_every, n_early_size
        self.WN = torch.nn.ModuleList()
        self.convinv = torch.nn.ModuleList()
        n_half, n_rem = n_group // 2, n_group
        for k in range(n_flows):
            if k > 0 and k % n_early_every == 0:
                n_half -= n_early_size // 2
                n_rem -= n_early_size
            self.convinv.append(Invertible1x1Conv(n_rem))
            self.WN.append(WN(n_half, n_mel_channels * n_group, **WN_config))
        self.n_remaining_channels = n_rem

    # ------------------------------------------------------------------ config / packing
    def config(self):
        wn0 = self.WN[0]
        return {
            "n_mel_channels": self.upsample.in_channels,
            "hop_length": self.upsample.stride[0],
            "n_flows": self.n_flows,
            "n_group": self.n_group,
            "n_early_every": self.n_early_every,
            "n_early_size": self.n_early_size,
            "WN_config": {"n_layers": wn0.n_layers, "n_channels": wn0.n_channels,
                          "kernel_size": wn0.in_layers[0].kernel_size[0]},
        }

    def plain_state(self):
        """Weight-norm-free state dict (what remove_weightnorm would leave)."""
        sd = {"upsample.weight": self.upsample.weight, "upsample.bias": self.upsample.bias}
        for k, wn in enumerate(self.WN):
            p = f"WN.{k}."
            sd[p + "start.weight"], sd[p + "start.bias"] = _plain_weight(wn.start), wn.start.bias
            sd[p + "end.weight"], sd[p + "end.bias"] = wn.end.weight, wn.end.bias
            for i in range(wn.n_layers):
                for name, layers in (("in_layers", wn.in_layers), ("cond_layers", wn.cond_layers),
                                     ("res_skip_layers", wn.res_skip_layers)):
                    sd[p + f"{name}.{i}.weight"] = _plain_weight(layers[i])
                    sd[p + f"{name}.{i}.bias"] = layers[i].bias
            sd[f"convinv.{k}.conv.weight"] = self.convinv[k].conv.weight
        return sd

    def _weights_signature(self):
        return tuple((p.data_ptr(), p._version, p.dtype) for p in self.parameters())

    def packed(self) -> PackedWaveGlow:
        """Packed fp32 weights on the module's device, rebuilt when a parameter changes."""
        sig = self._weights_signature()
        cache = getattr(self, "_fac_packed", None)
        if cache is None or cache[0] != sig:
            dev = self.upsample.weight.device
            _ext.require_cuda(self.upsample.weight, "WaveGlow parameters")
            cache = (sig, PackedWaveGlow.from_state(self.plain_state(), self.config(), dev))
            object.__setattr__(self, "_fac_packed", cache)
        return cache[1]

    def empty_packed(self) -> PackedWaveGlow:
        """Unfilled packed buffer of the right layout on the module's device (broadcast target)."""
        return PackedWaveGlow(self.config(), self.upsample.weight.device)

    def use_packed(self, packed: PackedWaveGlow):
        """Adopt externally provided packed weights (e.g. received by NCCL broadcast)."""
        object.__setattr__(self, "_fac_packed", (self._weights_signature(), packed))

    # ------------------------------------------------------------------ inference
    def noise_like_reference(self, batch, n_cols, device, dtype):
        """The N(0,1) draws of reference infer() in its order (glow.py:261-270, 285-290)."""
        draws = [torch.empty(batch, self.n_remaining_channels, n_cols, device=device, dtype=dtype).normal_()]
        for k in reversed(range(self.n_flows)):
            if k % self.n_early_every == 0 and k > 0:
                draws.append(torch.empty(batch, self.n_early_size, n_cols, device=device, dtype=dtype).normal_())
        return draws

    # Arithmetic of the WN layer GEMMs (>99 % of the FLOPs):
    #   "fp32"   exact fp32 on the FFMA pipe (parity anchor)
    #   "bf16x3" tcgen05 tensor cores, split-bf16 operands (hi+lo, 3 UMMAs per product), fp32
    #            accumulate: fp32-grade result (<= 1e-5 RMS on the waveform vs the fp32 reference)
    #   "bf16"   tcgen05 tensor cores, plain bf16 operands (BASELINE configs[2] precision)
    PRECISIONS = ("fp32", "bf16x3", "bf16")
    precision = "bf16x3"      # default: tensor cores with fp32-grade results
    # bf16x3: one fused launch per WN layer (csrc/waveglow_fused.cu); False: two launches per layer
    fused_layers = os.environ.get("FAC_TC_FUSED", "1") != "0"
    # True: a whole flow step (start, 8 layers, end / coupling / 1x1) is ONE cooperative launch in which a time tile
    # of a layer waits for its own and its neighbour tiles of the previous layer (fac_waveglow_flow_step_tc).
    # Measured on B200 it equals one launch per layer for a single short utterance and is 8-9 % slower at 8 x 10 s
    # (profiles/README.md), so it is opt-in.  FAC_TC_FUSED=3: the same with start and end as separate kernels.
    flow_step_launch = os.environ.get("FAC_TC_FUSED", "1") in ("2", "3")
    fused_bf16 = False        # plain bf16 through the fused kernel too (slower, see _alloc_io)

    def set_precision(self, precision):
        if precision in (None, "auto"):
            precision = "bf16x3"
        if precision not in self.PRECISIONS:
            raise ValueError("precision must be one of %s" % (self.PRECISIONS,))
        self.precision = precision
        return self

    def _nsplit(self):
        return {"fp32": 0, "bf16": 1, "bf16x3": 2}[self.precision]

    def _alloc_io(self, spect, sigma, noise):
        """Allocates the audio slot buffer (pre-filled with sigma*z), mel in channels-last
        form and the workspace.  Everything here is PyTorch plumbing around raw pointers."""
        B, n_mel, F = spect.shape
        hop, G = self.upsample.stride[0], self.n_group
        Tg = F * hop // G
        dev = spect.device
        if noise is None:
            noise = self.noise_like_reference(B, Tg, dev, spect.dtype)
        # the flow with n_rem live channels owns the LAST n_rem slots of every column, so
        # the first draw fills the last slots and each early draw the slots before them
        audio = torch.empty(B, Tg, G, device=dev, dtype=torch.float32)
        hi = G
        for z in noise:
            lo = hi - z.shape[1]
            audio[:, :, lo:hi] = (sigma * z).float().transpose(1, 2)
            hi = lo
        if hi != 0:
            raise ValueError("noise draws do not cover n_group channels")
        mel_cl = spect.float().transpose(1, 2).contiguous()
        Cn, n_cond = self.WN[0].n_channels, n_mel * G
        f32 = lambda c: torch.empty(B, Tg, c, device=dev, dtype=torch.float32)      # noqa: E731
        b16 = lambda c: torch.empty(B, Tg, c, device=dev, dtype=torch.bfloat16)     # noqa: E731
        bufs = {"audio": audio, "mel_cl": mel_cl}
        nsplit = self._nsplit()
        if nsplit == 0:
            bufs["spect"], bufs["x"], bufs["skip"], bufs["acts"] = f32(n_cond), f32(Cn), f32(Cn), f32(Cn)
            bufs["ws"] = _ext.WgWorkspace(bufs["spect"].data_ptr(), bufs["x"].data_ptr(), bufs["acts"].data_ptr(),
                                          bufs["skip"].data_ptr())
        else:
            # a layer is ONE fused launch that ping-pongs the residual stream between x and x2 and keeps the gated
            # activations on the SM (otherwise: two launches per layer with acts through HBM).  Plain bf16 stays on
            # the two-launch form by default: with a third of the UMMA time per unit the fused kernel's serialized
            # register-drain epilogue becomes the bottleneck (measured 52 vs 39 ms per 8 x 10 s step)
            fused = self.fused_layers and (nsplit == 2 or self.fused_bf16)
            names = (("spect_hi", n_cond), ("x_hi", Cn)) + ((("x2_hi", Cn),) if fused else (("acts_hi", Cn),))
            for name, c in names:
                bufs[name] = b16(c)
                bufs[name[:-2] + "lo"] = b16(c) if nsplit == 2 else None
            bufs["out8"] = f32(8)
            # one counter per 128-column time tile (include/fac_b200.h: flow_sync)
            bufs["flow_sync"] = (torch.zeros(B * ((Tg + 127) // 128) + 8, device=dev, dtype=torch.int32)
                                 if fused and self.flow_step_launch else None)
            pad = self.packed().mel_pad
            bufs["mel_hi"] = torch.empty(B, F, pad, device=dev, dtype=torch.bfloat16)
            bufs["mel_lo"] = torch.empty_like(bufs["mel_hi"]) if nsplit == 2 else None
            bufs["ws"] = _ext.WgTcWorkspace(
                bufs["mel_hi"].data_ptr(), _ext.ptr(bufs["mel_lo"]),
                bufs["spect_hi"].data_ptr(), _ext.ptr(bufs["spect_lo"]),
                bufs["x_hi"].data_ptr(), _ext.ptr(bufs["x_lo"]),
                _ext.ptr(bufs.get("acts_hi")), _ext.ptr(bufs.get("acts_lo")), bufs["out8"].data_ptr(),
                _ext.ptr(bufs.get("x2_hi")), _ext.ptr(bufs.get("x2_lo")), _ext.ptr(bufs.get("flow_sync")))
        return bufs, B, F, Tg

    # Small inputs are launch-bound (a 2 s utterance is ~240 launches of a few microseconds of work each, plus
    # six tensor-map encodes per GEMM on the host): up to this many frames in the batch, infer() is captured
    # once per (batch, frames, precision, sigma) in a CUDA graph and replayed.  0 disables.
    graph_max_frames = 4096
    _GRAPH_CACHE_SIZE = 4

    @torch.no_grad()
    def infer(self, spect, sigma=1.0, noise=None):
        """mel (B, n_mel, F) -> audio (B, F*hop); reference glow.py:252-293.

        ``noise`` optionally supplies the unit-normal draws (list in reference draw
        order); by default they come from torch's generator on spect's device with
        the reference's shapes and order, so seeding reproduces the reference."""
        _ext.require_cuda(spect, "spect")
        B, n_mel, F = spect.shape
        if n_mel != self.upsample.in_channels:
            raise ValueError("spect has %d mel channels, model expects %d" % (n_mel, self.upsample.in_channels))
        if B == 0 or F == 0:
            return spect.new_zeros(B, F * self.upsample.stride[0])
        if noise is None and 0 < B * F <= self.graph_max_frames and not torch.cuda.is_current_stream_capturing():
            return self._infer_graphed(spect, sigma)
        return self._infer_eager(spect, sigma, noise)

    def _infer_graphed(self, spect, sigma):
        lib = _ext.load()
        packed = self.packed()
        B, _, F = spect.shape
        # the N(0,1) draws stay OUTSIDE the graph: same generator calls, shapes and order as the eager path (and as
        # the reference), so seeding behaves identically; the graph reads them from static buffers
        noise = self.noise_like_reference(B, F * self.upsample.stride[0] // self.n_group, spect.device, spect.dtype)
        # the graphs live ON the packed weights they captured pointers of: a rebuilt (or adopted) pack starts with
        # an empty cache and the old graphs die with the old buffers -- never keyed by id(), which CPython reuses
        key = (tuple(spect.shape), spect.dtype, spect.device, self.precision, float(sigma))
        cache = packed.__dict__.setdefault("_graphs", {})
        entry = cache.pop(key, None)
        if entry is None:
            static_in, static_noise = spect.clone(), [z.clone() for z in noise]
            side = torch.cuda.Stream(device=spect.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                       # warm-up outside the capture (function attributes,
                self._infer_eager(static_in, sigma, static_noise)   # tensor-core weight copies, allocator)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            before = lib.fac_launch_count()
            try:
                with torch.cuda.graph(graph):
                    static_out = self._infer_eager(static_in, sigma, static_noise)
            except RuntimeError as exc:       # a launch kind this driver cannot capture: run eagerly from now on
                import warnings
                warnings.warn("fac_via_ppg_b200: CUDA graph capture of WaveGlow.infer failed (%s); graphs disabled" % exc)
                self.graph_max_frames = 0
                torch.cuda.synchronize()
                return self._infer_eager(spect, sigma, noise)
            entry = (graph, static_in, static_noise, static_out, lib.fac_launch_count() - before)
            while len(cache) >= self._GRAPH_CACHE_SIZE:
                cache.pop(next(iter(cache)))
        cache[key] = entry                                      # most recently used last
        graph, static_in, static_noise, static_out, n_launches = entry
        static_in.copy_(spect)
        for dst, z in zip(static_noise, noise):
            dst.copy_(z)
        graph.replay()
        lib.fac_add_launch_count(n_launches)                    # replayed launches never pass through the library
        return static_out.clone()

    def _infer_eager(self, spect, sigma, noise):
        lib = _ext.load()
        packed = self.packed()
        bufs, B, F, Tg = self._alloc_io(spect, sigma, noise)
        nsplit = self._nsplit()
        if nsplit == 0:
            rc = lib.fac_waveglow_infer_f32(C.byref(packed.cmodel), bufs["mel_cl"].data_ptr(),
                                            bufs["audio"].data_ptr(), C.byref(bufs["ws"]), B, F,
                                            _ext.current_stream())
            _ext.check(rc, "fac_waveglow_infer_f32")
        else:
            rc = lib.fac_waveglow_infer_tc(C.byref(packed.cmodel), C.byref(packed.tc_weights()),
                                           bufs["mel_cl"].data_ptr(), bufs["audio"].data_ptr(), C.byref(bufs["ws"]),
                                           B, F, nsplit, _ext.current_stream())
            _ext.check(rc, "fac_waveglow_infer_tc")
        return bufs["audio"].view(B, Tg * self.n_group).to(spect.dtype)

    @torch.no_grad()
    def profile_dominant_kernel(self, spect, peaks, passes=1):
        """Times the dominant kernel -- the WN layer launches (glow.py:158-174; one fused launch per layer in
        bf16x3, two otherwise) -- live with CUDA events on the launching stream, one event pair per (flow, layer),
        over ``passes`` complete back-to-back infer passes after one untimed pass (so the GPU is in the same
        power / clock state as in a long timed region), and returns the ``roofline`` object bench.py prints.
        Algorithmic work per layer: columns x (2C x (ks*C + n_cond) + n_rs x C) MACs."""
        lib = _ext.load()
        packed = self.packed()
        bufs, B, F, Tg = self._alloc_io(spect, 0.6, None)
        st = _ext.current_stream()
        m = C.byref(packed.cmodel)
        ws = C.byref(bufs["ws"])
        nsplit = self._nsplit()
        tcw = C.byref(packed.tc_weights()) if nsplit else None
        cfg = self.config()
        Cn, L, ks = cfg["WN_config"]["n_channels"], cfg["WN_config"]["n_layers"], cfg["WN_config"]["kernel_size"]
        n_cond = cfg["n_mel_channels"] * cfg["n_group"]
        pairs, macs, n_launch = [], 0, 0
        for it in range(passes + 1):
            timed = it > 0
            if nsplit == 0:
                _ext.check(lib.fac_waveglow_upsample_squeeze_f32(m, bufs["mel_cl"].data_ptr(), bufs["spect"].data_ptr(),
                                                                 B, F, st), "upsample")
            else:
                _ext.check(lib.fac_waveglow_tc_prepare_spect(m, tcw, ws, bufs["mel_cl"].data_ptr(), B, F, nsplit, st),
                           "prepare_spect")
            for k in reversed(range(self.n_flows)):
                if nsplit == 0:
                    _ext.check(lib.fac_wn_start_f32(m, k, bufs["audio"].data_ptr(), bufs["x"].data_ptr(), B, Tg, st),
                               "start")
                else:
                    _ext.check(lib.fac_wn_start_tc(m, k, bufs["audio"].data_ptr(), ws, B, Tg, nsplit, st), "start")
                for i in range(L):
                    if timed:
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        before = lib.fac_launch_count()
                        e0.record()
                    if nsplit == 0:
                        _ext.check(lib.fac_wn_layer_f32(m, k, i, ws, B, Tg, st), "layer")
                    else:
                        _ext.check(lib.fac_wn_layer_tc(m, tcw, k, i, ws, B, Tg, nsplit, st), "layer")
                    if timed:
                        e1.record()
                        n_launch += lib.fac_launch_count() - before
                        pairs.append((e0, e1))
                        macs += B * Tg * (2 * Cn * (ks * Cn + n_cond) + (2 * Cn if i < L - 1 else Cn) * Cn)
                if nsplit == 0:
                    _ext.check(lib.fac_wn_end_coupling_f32(m, k, bufs["skip"].data_ptr(), bufs["audio"].data_ptr(), B,
                                                           Tg, st), "end")
                else:
                    _ext.check(lib.fac_wn_end_tc(m, tcw, k, bufs["out8"].data_ptr(), bufs["audio"].data_ptr(), B, Tg,
                                                 st), "end")
        torch.cuda.synchronize()
        total_ms = sum(a.elapsed_time(b) for a, b in pairs)
        achieved = 2.0 * macs / (total_ms / 1e3) / 1e12
        peak = peaks["tflops_sustained"]
        fused = nsplit == 2 and n_launch == len(pairs)
        # tensor-pipe work actually executed per algorithmic product: 3 UMMAs in the split scheme; the two-launch
        # form adds the identity block that performs the residual add on the tensor core
        executed = {0: 1.0, 1: 1.0, 2: 3.0}[nsplit]
        kernel = {0: "conv_gemm_f32_kernel (FFMA), two launches per WN layer",
                  1: "wn_gemm_tc_kernel (tcgen05 bf16), two launches per WN layer",
                  2: ("wn_layer_fused_kernel (tcgen05 split-bf16, 3 UMMAs per product): ONE launch per WN layer = in+cond "
                      "GEMM -> gate -> residual GEMM -> residual add" if fused else
                      "wn_gemm_tc_kernel (tcgen05 split-bf16, 3 UMMAs per product), two launches per WN layer")}[nsplit]
        roof = {
            "bound": "tensor", "kernel": kernel,
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": peaks["source"] + ", bf16 sustained (kernel timed inside a long back-to-back region)",
            "launches_timed": n_launch, "passes": passes, "avg_launch_ms": total_ms / n_launch,
            "layer_ms_per_pass": total_ms / passes, "flop_per_launch": 2.0 * macs / n_launch, "traffic": None,
            "executed_tensor_tflops": achieved * executed if nsplit else 0.0,
            "executed_frac_of_peak": achieved * executed / peak if nsplit else 0.0,
        }
        if nsplit == 0:
            roof["note"] = ("exact-fp32 FFMA path: its own pipe peak is 74.4 TFLOP/s (148 SM x 128 lanes x 2 x "
                            "1.965 GHz), frac_of_ffma_peak=%.3f" % (achieved / 74.4))
        elif nsplit == 2:
            roof["note"] = ("`achieved` counts ALGORITHMIC flops; the split-bf16 scheme executes 3 bf16 UMMAs per "
                            "algorithmic product to reach fp32-grade results, so frac <= 1/3 by construction "
                            "(profiles/r2_precision_sweep.json: no cheaper operand scheme meets the 1e-4 RMS tolerance)")
        return roof

    def forward(self, forward_input):
        raise NotImplementedError("fac_via_ppg_b200 covers the inference path only (WaveGlow.infer); "
                                  "the training direction of reference glow.py:209-250 is out of scope")

    @staticmethod
    def remove_weightnorm(model):
        """reference glow.py:295-311: fold weight_g * v / |v| into plain weights."""
        rm = torch.nn.utils.remove_weight_norm
        for wn in model.WN:
            wn.start = rm(wn.start)
            for name in ("in_layers", "cond_layers", "res_skip_layers"):
                setattr(wn, name, torch.nn.ModuleList([rm(c) for c in getattr(wn, name)]))
        return model
