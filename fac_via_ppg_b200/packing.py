"""Weight repacking: reference-format state dicts -> the flat HBM buffer the kernels read.

All packed tensors of a model live in ONE flat fp32 device buffer (256-byte
aligned sub-allocations), so that the only collective of the multi-GPU path is a
single ``torch.distributed.broadcast`` of that buffer (see dist.py).  The layout is
a pure function of the model configuration, so a rank that receives the buffer
can rebuild the pointer table without seeing the state dict.

WaveGlow input format: the state dict of a ``WaveGlow`` after
``remove_weightnorm`` (reference src/waveglow/glow.py:295-311).
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import torch

from . import _ext
from .synth import UPSAMPLE_KERNEL, flow_channels

ALIGN_ELEMS = 64  # 256 bytes


def _round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


class FlatLayout:
    """name -> (offset, shape) inside one flat fp32 buffer."""

    def __init__(self):
        self.entries = OrderedDict()
        self.size = 0

    def add(self, name, shape):
        n = 1
        for s in shape:
            n *= int(s)
        self.entries[name] = (self.size, tuple(int(s) for s in shape))
        self.size += _round_up(n, ALIGN_ELEMS)

    def view(self, flat, name):
        off, shape = self.entries[name]
        n = 1
        for s in shape:
            n *= s
        return flat[off:off + n].view(shape)

    def ptr(self, flat, name):
        return flat.data_ptr() + self.entries[name][0] * flat.element_size()


def upsample_taps(cfg) -> int:
    return -(-UPSAMPLE_KERNEL // cfg["hop_length"])


def waveglow_layout(cfg) -> FlatLayout:
    wn = cfg["WN_config"]
    Cn, L, ks = wn["n_channels"], wn["n_layers"], wn["kernel_size"]
    n_mel, n_group, hop = cfg["n_mel_channels"], cfg["n_group"], cfg["hop_length"]
    n_cond = n_mel * n_group
    phases = hop // n_group
    lay = FlatLayout()
    lay.add("upsample_w", (phases, upsample_taps(cfg) * n_mel, _round_up(n_cond, 128)))
    lay.add("upsample_b", (_round_up(n_cond, 128),))
    for k, (n_rem, n_half) in enumerate(flow_channels(cfg)):
        lay.add(f"{k}.start_w", (n_half, Cn))
        lay.add(f"{k}.start_b", (Cn,))
        lay.add(f"{k}.end_w", (2 * n_half, Cn))
        lay.add(f"{k}.end_b", (2 * n_half,))
        lay.add(f"{k}.w_inv", (n_rem, n_rem))
        for i in range(L):
            n_rs = 2 * Cn if i < L - 1 else Cn
            lay.add(f"{k}.{i}.in_cond_w", (ks * Cn + n_cond, _round_up(2 * Cn, 128)))
            lay.add(f"{k}.{i}.in_cond_b", (_round_up(2 * Cn, 128),))
            lay.add(f"{k}.{i}.res_skip_w", (Cn, _round_up(n_rs, 128)))
            lay.add(f"{k}.{i}.res_skip_b", (_round_up(n_rs, 128),))
    return lay


def _pad_cols(t, n_pad):
    if t.shape[-1] == n_pad:
        return t
    out = t.new_zeros(t.shape[:-1] + (n_pad,))
    out[..., : t.shape[-1]] = t
    return out


def validate_waveglow_cfg(cfg):
    wn = cfg["WN_config"]
    if cfg["n_flows"] > _ext.FAC_MAX_FLOWS or wn["n_layers"] > _ext.FAC_MAX_LAYERS:
        raise _ext.FacError("WaveGlow geometry exceeds FAC_MAX_FLOWS/FAC_MAX_LAYERS")
    if wn["n_channels"] % 8 or cfg["n_mel_channels"] % 8:
        raise _ext.FacError("n_channels and n_mel_channels must be multiples of 8")
    if cfg["hop_length"] % cfg["n_group"]:
        raise _ext.FacError("hop_length must be a multiple of n_group")
    if cfg["n_group"] > 8 or cfg["n_group"] % 2:
        raise _ext.FacError("n_group must be even and <= 8")


class PackedWaveGlow:
    """Flat packed weights + the ``fac_wg_model`` pointer table."""

    def __init__(self, cfg, device):
        validate_waveglow_cfg(cfg)
        self.cfg = cfg
        self.layout = waveglow_layout(cfg)
        self.flat = torch.zeros(self.layout.size, dtype=torch.float32, device=device)
        self.cmodel = self._pointer_table()

    def _pointer_table(self):
        cfg, lay, flat = self.cfg, self.layout, self.flat
        wn = cfg["WN_config"]
        m = _ext.WgModel()
        m.n_flows, m.n_layers, m.n_channels = cfg["n_flows"], wn["n_layers"], wn["n_channels"]
        m.n_group, m.n_mel, m.hop = cfg["n_group"], cfg["n_mel_channels"], cfg["hop_length"]
        m.n_early_every, m.n_early_size = cfg["n_early_every"], cfg["n_early_size"]
        m.upsample_taps, m.kernel_size = upsample_taps(cfg), wn["kernel_size"]
        m.upsample_w = lay.ptr(flat, "upsample_w")
        m.upsample_b = lay.ptr(flat, "upsample_b")
        for k, (n_rem, n_half) in enumerate(flow_channels(cfg)):
            f = m.flows[k]
            f.n_half, f.n_rem = n_half, n_rem
            for name in ("start_w", "start_b", "end_w", "end_b", "w_inv"):
                setattr(f, name, lay.ptr(flat, f"{k}.{name}"))
            for i in range(wn["n_layers"]):
                f.in_cond_w[i] = lay.ptr(flat, f"{k}.{i}.in_cond_w")
                f.in_cond_b[i] = lay.ptr(flat, f"{k}.{i}.in_cond_b")
                f.res_skip_w[i] = lay.ptr(flat, f"{k}.{i}.res_skip_w")
                f.res_skip_b[i] = lay.ptr(flat, f"{k}.{i}.res_skip_b")
        return m

    @property
    def nbytes(self):
        return self.flat.numel() * 4

    @torch.no_grad()
    def load_state(self, sd):
        """Fill the flat buffer from a weight-norm-free WaveGlow state dict."""
        cfg, lay, flat = self.cfg, self.layout, self.flat
        dev = flat.device
        wn = cfg["WN_config"]
        Cn, L, ks = wn["n_channels"], wn["n_layers"], wn["kernel_size"]
        n_mel, n_group, hop = cfg["n_mel_channels"], cfg["n_group"], cfg["hop_length"]
        phases, taps = hop // n_group, upsample_taps(cfg)

        def get(name):
            return sd[name].detach().to(device=dev, dtype=torch.float32)

        # Transposed conv as `phases` small GEMMs (glow.py:253): output sample
        # n = n_group*(phases*f + p) + j takes tap index n - hop*(f - k) = n_group*p + j + hop*k
        # of the kernel from input frame f - k.
        w_up = get("upsample.weight")                                   # (in i, out m, 1024)
        w_ext = w_up.new_zeros(n_mel, n_mel, taps * hop)
        w_ext[:, :, : w_up.shape[2]] = w_up
        w_ext = w_ext.view(n_mel, n_mel, taps, phases, n_group)         # i, m, k, p, j
        w_ph = w_ext.permute(3, 2, 0, 1, 4).reshape(phases, taps * n_mel, n_mel * n_group)
        lay.view(flat, "upsample_w").copy_(_pad_cols(w_ph, lay.entries["upsample_w"][1][2]))
        lay.view(flat, "upsample_b")[: n_mel * n_group].copy_(get("upsample.bias").repeat_interleave(n_group))

        perm = torch.stack([torch.arange(Cn), torch.arange(Cn) + Cn], dim=1).flatten().to(dev)
        for k, (n_rem, n_half) in enumerate(flow_channels(cfg)):
            p = f"WN.{k}."
            lay.view(flat, f"{k}.start_w").copy_(get(p + "start.weight")[:, :, 0].t())
            lay.view(flat, f"{k}.start_b").copy_(get(p + "start.bias"))
            lay.view(flat, f"{k}.end_w").copy_(get(p + "end.weight")[:, :, 0])
            lay.view(flat, f"{k}.end_b").copy_(get(p + "end.bias"))
            # glow.py:89-95: W^-1 is computed once in fp32 and cached
            lay.view(flat, f"{k}.w_inv").copy_(get(f"convinv.{k}.conv.weight")[:, :, 0].inverse())
            for i in range(L):
                w_in = get(p + f"in_layers.{i}.weight")                 # (2C, C, ks)
                w_cond = get(p + f"cond_layers.{i}.weight")[:, :, 0]    # (2C, n_cond)
                w1 = torch.cat([w_in.permute(2, 1, 0).reshape(ks * Cn, 2 * Cn), w_cond.t()], dim=0)[:, perm]
                b1 = (get(p + f"in_layers.{i}.bias") + get(p + f"cond_layers.{i}.bias"))[perm]
                v = lay.view(flat, f"{k}.{i}.in_cond_w")
                v.copy_(_pad_cols(w1, v.shape[1]))
                lay.view(flat, f"{k}.{i}.in_cond_b")[: 2 * Cn].copy_(b1)
                w2 = get(p + f"res_skip_layers.{i}.weight")[:, :, 0].t()  # (C, n_rs)
                v = lay.view(flat, f"{k}.{i}.res_skip_w")
                v.copy_(_pad_cols(w2, v.shape[1]))
                b2 = get(p + f"res_skip_layers.{i}.bias")
                lay.view(flat, f"{k}.{i}.res_skip_b")[: b2.numel()].copy_(b2)
        return self

    @classmethod
    def from_state(cls, sd, cfg, device):
        return cls(cfg, device).load_state(sd)
