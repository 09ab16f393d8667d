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
import hashlib
import json
import os
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
        self._tc = None
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

    # ------------------------------------------------------------------ tensor-core copies
    @property
    def mel_pad(self):
        """mel channels rounded up to the largest tensor-core K block (64)."""
        return _round_up(self.cfg["n_mel_channels"], 64)

    def tc_layouts(self):
        """Layouts of the derived tensor-core weights: (bf16 operand matrices, fp32 side tables)."""
        cfg = self.cfg
        wn = cfg["WN_config"]
        Cn, L, ks = wn["n_channels"], wn["n_layers"], wn["kernel_size"]
        n_cond = cfg["n_mel_channels"] * cfg["n_group"]
        l16, l32 = FlatLayout(), FlatLayout()
        phases, taps = cfg["hop_length"] // cfg["n_group"], upsample_taps(cfg)
        for part in ("hi", "lo"):
            l16.add(f"up_{part}", (phases, n_cond, taps * self.mel_pad))
        for k in range(cfg["n_flows"]):
            l32.add(f"{k}.out_bias", (8,))
            for i in range(L):
                # every layer occupies the same block (the last one leaves its residual matrices zero): a uniform
                # layer stride lets ONE 3-D tensor map per matrix kind serve all layers of a flow (waveglow_fused.cu)
                for part in ("hi", "lo"):
                    l16.add(f"{k}.{i}.w1_{part}", (2 * Cn, ks * Cn + n_cond))
                    l16.add(f"{k}.{i}.w2_{part}", (Cn, 2 * Cn))
                    l16.add(f"{k}.{i}.w2r_{part}", (Cn, Cn))
                l32.add(f"{k}.{i}.wc", (8, Cn))
                if i < L - 1:
                    l32.add(f"{k}.{i}.res_b", (Cn,))
        return l16, l32

    @torch.no_grad()
    def tc_weights(self):
        """Builds (once) the tensor-core weight buffers from the packed fp32 buffer and returns the
        ``fac_wg_tc_weights`` pointer table (see include/fac_b200.h for the algebra).  Derived data:
        after a broadcast of ``flat`` every rank derives its own copy locally."""
        cached = getattr(self, "_tc", None)
        if cached is not None:
            return cached[-1]
        cfg, lay = self.cfg, self.layout
        wn = cfg["WN_config"]
        Cn, L = wn["n_channels"], wn["n_layers"]
        dev = self.flat.device
        l16, l32 = self.tc_layouts()
        flat16 = torch.zeros(l16.size, dtype=torch.bfloat16, device=dev)
        flat32 = torch.zeros(l32.size, dtype=torch.float32, device=dev)
        table = _ext.WgTcWeights()
        eye = torch.eye(Cn, device=dev)
        # upsampler phase matrices: fp32 [phases][taps*n_mel][n_cond] -> [phases][n_cond][taps*mel_pad]
        n_mel, taps, pad = cfg["n_mel_channels"], upsample_taps(cfg), self.mel_pad
        n_cond = n_mel * cfg["n_group"]
        w_up = lay.view(self.flat, "upsample_w")[:, :, :n_cond]
        w_up = w_up.reshape(w_up.shape[0], taps, n_mel, n_cond).permute(0, 3, 1, 2)          # p, n, tap, c
        w_pad = w_up.new_zeros(w_up.shape[0], n_cond, taps, pad)
        w_pad[..., :n_mel] = w_up
        w_pad = w_pad.reshape(w_up.shape[0], n_cond, taps * pad)
        up_hi = w_pad.to(torch.bfloat16)
        l16.view(flat16, "up_hi").copy_(up_hi)
        l16.view(flat16, "up_lo").copy_((w_pad - up_hi.float()).to(torch.bfloat16))
        table.up_hi, table.up_lo, table.mel_pad = l16.ptr(flat16, "up_hi"), l16.ptr(flat16, "up_lo"), pad

        def put16(name, w, flow, i):
            hi = w.to(torch.bfloat16)
            lo = (w - hi.float()).to(torch.bfloat16)
            stem = name.split(".")[-1]
            l16.view(flat16, name + "_hi").copy_(hi)
            l16.view(flat16, name + "_lo").copy_(lo)
            getattr(flow, stem + "_hi")[i] = l16.ptr(flat16, name + "_hi")
            getattr(flow, stem + "_lo")[i] = l16.ptr(flat16, name + "_lo")

        for k, (n_rem, n_half) in enumerate(flow_channels(cfg)):
            flow = table.flows[k]
            end_w = lay.view(self.flat, f"{k}.end_w").double()                    # (2*n_half, C)
            bias8 = lay.view(self.flat, f"{k}.end_b").double().clone()
            for i in range(L):
                last = i == L - 1
                n_rs = Cn if last else 2 * Cn
                put16(f"{k}.{i}.w1", lay.view(self.flat, f"{k}.{i}.in_cond_w")[:, : 2 * Cn].t().contiguous(), flow, i)
                w_rs = lay.view(self.flat, f"{k}.{i}.res_skip_w")[:, :n_rs].t()       # (n_rs, C)
                b_rs = lay.view(self.flat, f"{k}.{i}.res_skip_b")[:n_rs]
                w_skip, b_skip = (w_rs, b_rs) if last else (w_rs[Cn:], b_rs[Cn:])
                wc = l32.view(flat32, f"{k}.{i}.wc")
                wc[: 2 * n_half].copy_((end_w @ w_skip.double()).float())            # composed in fp64
                flow.wc[i] = l32.ptr(flat32, f"{k}.{i}.wc")
                bias8 += end_w @ b_skip.double()
                if not last:
                    put16(f"{k}.{i}.w2", torch.cat([w_rs[:Cn], eye], dim=1).contiguous(), flow, i)
                    put16(f"{k}.{i}.w2r", w_rs[:Cn].contiguous(), flow, i)
                    l32.view(flat32, f"{k}.{i}.res_b").copy_(b_rs[:Cn])
                    flow.res_b[i] = l32.ptr(flat32, f"{k}.{i}.res_b")
            l32.view(flat32, f"{k}.out_bias")[: 2 * n_half].copy_(bias8.float())
        table = self._tc_table(flat16, flat32)
        self._tc = (flat16, flat32, table)
        return table

    def _tc_table(self, flat16, flat32):
        """``fac_wg_tc_weights`` pointer table over the two derived buffers (a pure function of the layouts)."""
        l16, l32 = self.tc_layouts()
        L = self.cfg["WN_config"]["n_layers"]
        table = _ext.WgTcWeights()
        table.up_hi, table.up_lo, table.mel_pad = l16.ptr(flat16, "up_hi"), l16.ptr(flat16, "up_lo"), self.mel_pad
        for k in range(self.cfg["n_flows"]):
            flow = table.flows[k]
            flow.out_bias = l32.ptr(flat32, f"{k}.out_bias")
            for i in range(L):
                flow.w1_hi[i], flow.w1_lo[i] = l16.ptr(flat16, f"{k}.{i}.w1_hi"), l16.ptr(flat16, f"{k}.{i}.w1_lo")
                flow.wc[i] = l32.ptr(flat32, f"{k}.{i}.wc")
                if i < L - 1:
                    flow.w2_hi[i], flow.w2_lo[i] = l16.ptr(flat16, f"{k}.{i}.w2_hi"), l16.ptr(flat16, f"{k}.{i}.w2_lo")
                    flow.w2r_hi[i], flow.w2r_lo[i] = l16.ptr(flat16, f"{k}.{i}.w2r_hi"), l16.ptr(flat16, f"{k}.{i}.w2r_lo")
                    flow.res_b[i] = l32.ptr(flat32, f"{k}.{i}.res_b")
        return table

    def to(self, device):
        """A copy of the packed weights (and of the derived tensor-core buffers, when built) on ``device``: pack
        once on the host -- or load from the pack cache -- and ship three flat buffers."""
        new = PackedWaveGlow(self.cfg, device)
        new.flat.copy_(self.flat)
        if getattr(self, "_tc", None) is not None:
            f16, f32 = self._tc[0].to(device), self._tc[1].to(device)
            new._tc = (f16, f32, new._tc_table(f16, f32))
        return new

    @classmethod
    def from_state(cls, sd, cfg, device, cache_dir=None):
        """Packs ``sd``.  With ``cache_dir`` (default: the FAC_PACK_CACHE environment variable; unset = no cache)
        the flat buffer is kept on disk under a key derived from the configuration, the pack format and a hash of
        every weight, so a checkpoint is repacked (transposes, gate interleaving, W^-1 through cuSOLVER ...) once
        per machine instead of once per process (reference load path: src/common/utils.py:177-181)."""
        cache_dir = os.environ.get("FAC_PACK_CACHE") if cache_dir is None else cache_dir
        if not cache_dir:
            return cls(cfg, device).load_state(sd)
        path = os.path.join(cache_dir, "waveglow-%s.pack" % state_fingerprint(sd, cfg))
        packed = cls(cfg, device)
        if os.path.isfile(path):
            try:
                blob = torch.load(path, map_location="cpu", weights_only=True)
                if blob["format"] == PACK_FORMAT and blob["flat"].numel() == packed.flat.numel():
                    packed.flat.copy_(blob["flat"])
                    packed.from_cache = path
                    return packed
            except Exception:          # unreadable / truncated cache entry: repack and overwrite
                pass
        packed.load_state(sd)
        os.makedirs(cache_dir, exist_ok=True)
        tmp = "%s.tmp.%d" % (path, os.getpid())
        torch.save({"format": PACK_FORMAT, "flat": packed.flat.detach().cpu()}, tmp)
        os.replace(tmp, path)          # atomic: concurrent ranks see the old or the new file
        return packed


PACK_FORMAT = 2        # bump when waveglow_layout() or load_state() change what the flat buffer holds


def state_fingerprint(sd, cfg) -> str:
    """sha256 over the pack format, the configuration and every tensor (name, shape, bytes) of a state dict."""
    h = hashlib.sha256()
    h.update(("fac-pack-%d|" % PACK_FORMAT).encode())
    h.update(json.dumps(cfg, sort_keys=True).encode())
    for name in sorted(sd):
        t = sd[name].detach().to("cpu", torch.float32).contiguous()
        h.update(("|%s|%s|" % (name, tuple(t.shape))).encode())
        h.update(t.numpy().tobytes())
    return h.hexdigest()[:32]


# ===================================================================== Tacotron2 (PPG -> Mel)
def _fold_bn(sd, prefix, dev, eps=1e-5):
    """Conv1d + BatchNorm1d(eval) -> one conv: w' = w * g / sqrt(var + eps), b' = (b - mean) * g / sqrt(...) + beta
    (reference src/common/model.py:143-176, 199-209; BatchNorm1d default eps)."""
    f32 = dict(device=dev, dtype=torch.float32)
    w = sd[prefix + "0.conv.weight"].detach().to(**f32)
    b = sd[prefix + "0.conv.bias"].detach().to(**f32)
    scale = sd[prefix + "1.weight"].detach().to(**f32) / torch.sqrt(sd[prefix + "1.running_var"].detach().to(**f32) + eps)
    shift = sd[prefix + "1.bias"].detach().to(**f32) - sd[prefix + "1.running_mean"].detach().to(**f32) * scale
    return w * scale[:, None, None], b * scale + shift


def tacotron_layout(hp) -> FlatLayout:
    E, D, M = hp["encoder_embedding_dim"], hp["n_symbols"], hp["n_acoustic_feat_dims"]
    P, A, R = hp["prenet_dim"], hp["attention_dim"], hp["attention_rnn_dim"]
    H = E // 2
    ke, kp, Pe = hp["encoder_kernel_size"], hp["postnet_kernel_size"], hp["postnet_embedding_dim"]
    pad = lambda n: _round_up(n, 128)  # noqa: E731
    lay = FlatLayout()
    lay.add("enc.pre0_w", (D, pad(E)))
    lay.add("enc.pre1_w", (E, pad(E)))
    for i in range(hp["encoder_n_convolutions"]):
        lay.add(f"enc.conv{i}_w", (ke * E, pad(E)))
        lay.add(f"enc.conv{i}_b", (pad(E),))
    lay.add("enc.lstm_ih_w", (E, pad(8 * H)))
    lay.add("enc.lstm_ih_b", (pad(8 * H),))
    lay.add("enc.lstm_hh", (2, 4 * H, H))
    lay.add("dec.mem_w", (E, pad(A)))
    kin = P + E + R
    lay.add("dec.w_att", (4 * R, kin))
    lay.add("dec.b_att", (4 * R,))
    lay.add("dec.w_dec", (4 * R, kin))
    lay.add("dec.b_dec", (4 * R,))
    lay.add("dec.wq", (A, R))
    lay.add("dec.w_loc", (2, hp["attention_location_kernel_size"], hp["attention_location_n_filters"]))
    lay.add("dec.w_ld_t", (hp["attention_location_n_filters"], A))
    lay.add("dec.v", (A,))
    lay.add("dec.w_pp", (M + 1 + P, R + E))
    lay.add("dec.b_pp", (M + 1 + P,))
    lay.add("dec.w_pre2", (P, P))
    n_post = hp["postnet_n_convolutions"]
    dims = [M] + [Pe] * (n_post - 1) + [M]
    for i in range(n_post):
        lay.add(f"post.conv{i}_w", (kp * dims[i], pad(dims[i + 1])))
        lay.add(f"post.conv{i}_b", (pad(dims[i + 1]),))
    return lay


def validate_tacotron_hparams(hp):
    """The persistent decoder kernel is compiled for the geometry of create_hparams_stage()
    (reference src/common/hparams.py:167-231); other sizes fail loudly instead of silently."""
    want = {"encoder_embedding_dim": 600, "prenet_dim": 300, "attention_rnn_dim": 300, "decoder_rnn_dim": 300,
            "attention_dim": 150, "attention_location_n_filters": 32, "attention_location_kernel_size": 31,
            "n_acoustic_feat_dims": 80}
    bad = {k: hp[k] for k, v in want.items() if hp[k] != v}
    if bad:
        raise _ext.FacError("decoder kernel is built for %s; got %s" % (want, bad))
    if hp["n_symbols"] % 8 or hp["postnet_embedding_dim"] % 8:
        raise _ext.FacError("n_symbols and postnet_embedding_dim must be multiples of 8")
    w = hp["attention_window_size"]
    if w is None or 2 * w + 1 > 48:
        raise _ext.FacError("attention_window_size must be set and <= 23 (the attention CTA keeps the window's "
                            "encoder rows in registers and its location terms in shared memory)")


class PackedTacotron:
    """Flat packed weights of the PPG->Mel model + the decoder pointer table."""

    def __init__(self, hp, device):
        validate_tacotron_hparams(hp)
        self.hp = dict(hp)
        self.layout = tacotron_layout(hp)
        self.flat = torch.zeros(self.layout.size, dtype=torch.float32, device=device)
        w = _ext.TacoDecoderWeights()
        for name in ("w_att", "b_att", "w_dec", "b_dec", "wq", "w_loc", "w_ld_t", "v", "w_pp", "b_pp", "w_pre2"):
            setattr(w, name, self.layout.ptr(self.flat, "dec." + name))
        self.cdecoder = w

    def view(self, name):
        return self.layout.view(self.flat, name)

    @property
    def nbytes(self):
        return self.flat.numel() * 4

    @torch.no_grad()
    def load_state(self, sd):
        hp, dev = self.hp, self.flat.device
        E, M = hp["encoder_embedding_dim"], hp["n_acoustic_feat_dims"]
        H = E // 2

        def get(name):
            return sd[name].detach().to(device=dev, dtype=torch.float32)

        def put(name, t):
            v = self.view(name)
            if v.dim() == 2 and t.shape[1] != v.shape[1]:
                t = _pad_cols(t, v.shape[1])
            elif v.dim() == 1 and t.numel() != v.numel():
                t = torch.cat([t, t.new_zeros(v.numel() - t.numel())])
            v.copy_(t)

        put("enc.pre0_w", get("encoder.prenet.layers.0.linear_layer.weight").t())
        put("enc.pre1_w", get("encoder.prenet.layers.1.linear_layer.weight").t())
        for i in range(hp["encoder_n_convolutions"]):
            w, b = _fold_bn(sd, f"encoder.convolutions.{i}.", dev)
            put(f"enc.conv{i}_w", w.permute(2, 1, 0).reshape(-1, w.shape[0]))      # rows (tap, c_in)
            put(f"enc.conv{i}_b", b)
        w_ih = torch.cat([get("encoder.lstm.weight_ih_l0"), get("encoder.lstm.weight_ih_l0_reverse")], dim=0)
        b_ih = torch.cat([get("encoder.lstm.bias_ih_l0") + get("encoder.lstm.bias_hh_l0"),
                          get("encoder.lstm.bias_ih_l0_reverse") + get("encoder.lstm.bias_hh_l0_reverse")])
        put("enc.lstm_ih_w", w_ih.t())
        put("enc.lstm_ih_b", b_ih)
        self.view("enc.lstm_hh").copy_(torch.stack([get("encoder.lstm.weight_hh_l0"),
                                                    get("encoder.lstm.weight_hh_l0_reverse")]))
        al = "decoder.attention_layer."
        put("dec.mem_w", get(al + "memory_layer.linear_layer.weight").t())
        for cell, name in (("attention_rnn", "att"), ("decoder_rnn", "dec")):
            p = f"decoder.{cell}."
            put(f"dec.w_{name}", torch.cat([get(p + "weight_ih"), get(p + "weight_hh")], dim=1))
            put(f"dec.b_{name}", get(p + "bias_ih") + get(p + "bias_hh"))
        put("dec.wq", get(al + "query_layer.linear_layer.weight"))
        self.view("dec.w_loc").copy_(get(al + "location_layer.location_conv.conv.weight").permute(1, 2, 0))
        put("dec.w_ld_t", get(al + "location_layer.location_dense.linear_layer.weight").t())
        put("dec.v", get(al + "v.linear_layer.weight")[0])
        # The next step's prenet layer 0 is applied straight to the projected frame (no bias, no
        # nonlinearity in between): compose it with the projection in fp64 once.
        w_proj, b_proj = get("decoder.linear_projection.linear_layer.weight"), get("decoder.linear_projection.linear_layer.bias")
        w_pre0 = get("decoder.prenet.layers.0.linear_layer.weight").double()
        put("dec.w_pp", torch.cat([w_proj, get("decoder.gate_layer.linear_layer.weight"),
                                   (w_pre0 @ w_proj.double()).float()], dim=0))
        put("dec.b_pp", torch.cat([b_proj, get("decoder.gate_layer.linear_layer.bias"),
                                   (w_pre0 @ b_proj.double()).float()]))
        put("dec.w_pre2", get("decoder.prenet.layers.1.linear_layer.weight"))
        for i in range(hp["postnet_n_convolutions"]):
            w, b = _fold_bn(sd, f"postnet.convolutions.{i}.", dev)
            put(f"post.conv{i}_w", w.permute(2, 1, 0).reshape(-1, w.shape[0]))
            put(f"post.conv{i}_b", b)
        # the recurrent kernels keep these matrices in shared memory as half hi/lo pairs of w * 2^8
        for name in ("enc.lstm_hh", "dec.w_att", "dec.w_dec", "dec.w_pp", "dec.w_pre2"):
            peak = float(self.view(name).abs().max())
            if not peak < 200.0:
                raise _ext.FacError("%s: |weight| up to %.3g is outside the range (< 200) the resident half-precision "
                                    "operand pairs of the recurrent kernels support" % (name, peak))
        return self

    # ------------------------------------------------------------------ tensor-core copies
    @torch.no_grad()
    def tc_weights(self, dtype=torch.float16):
        """16-bit hi/lo operand matrices (IEEE half: hi + lo carries 22 significand bits; the weights and
        activations of this model are far inside its range) of the encoder / postnet GEMMs for fac_conv_gemm_tc, derived (once) from the
        packed fp32 buffer: name -> dict(hi, lo, c_pad, taps, n_pad, n_valid, bias).  Layout [n_pad][taps*c_pad]
        (K contiguous, tap-major), input channels zero-padded to a multiple of 64, rows to a multiple of 64."""
        cached = getattr(self, "_tc", None)
        if cached is not None:
            return cached
        hp = self.hp
        E, D, M = hp["encoder_embedding_dim"], hp["n_symbols"], hp["n_acoustic_feat_dims"]
        ke, kp, Pe = hp["encoder_kernel_size"], hp["postnet_kernel_size"], hp["postnet_embedding_dim"]
        n_post = hp["postnet_n_convolutions"]
        dims = [M] + [Pe] * (n_post - 1) + [M]
        specs = [("enc.pre0", D, 1, E, False), ("enc.pre1", E, 1, E, False)]
        specs += [(f"enc.conv{i}", E, ke, E, True) for i in range(hp["encoder_n_convolutions"])]
        specs += [("enc.lstm_ih", E, 1, 4 * E, True)]
        specs += [(f"post.conv{i}", dims[i], kp, dims[i + 1], True) for i in range(n_post)]
        out = {}
        for name, c_in, taps, n_out, has_bias in specs:
            w = self.view(name + "_w")[:, :n_out]                              # (taps*c_in, n_out) fp32
            c_pad, n_pad = _round_up(c_in, 64), _round_up(n_out, 64)
            wp = w.new_zeros(n_pad, taps, c_pad)
            wp[:n_out, :, :c_in] = w.reshape(taps, c_in, n_out).permute(2, 0, 1)
            wp = wp.reshape(n_pad, taps * c_pad)
            if dtype == torch.float16 and float(wp.abs().max()) > 3.0e4:
                raise _ext.FacError("%s: |weight| up to %.3g does not fit the half-precision operand range of the "
                                    "tensor-core path; use Tacotron2.set_precision('fp32')" % (name, float(wp.abs().max())))
            hi = wp.to(dtype)
            lo = (wp - hi.float()).to(dtype)
            bias = None
            if has_bias:
                bias = w.new_zeros(n_pad)
                bias[:n_out] = self.view(name + "_b")[:n_out]
            out[name] = dict(hi=hi.contiguous(), lo=lo.contiguous(), c_pad=c_pad, taps=taps, n_pad=n_pad,
                             n_valid=n_out, bias=bias)
        self._tc = out
        return out

    @classmethod
    def from_state(cls, sd, hp, device):
        return cls(hp, device).load_state(sd)

    def to(self, device):
        """A copy on ``device`` (flat buffer + the derived tensor-core operand matrices, when built)."""
        new = PackedTacotron(self.hp, device)
        new.flat.copy_(self.flat)
        if getattr(self, "_tc", None) is not None:
            new._tc = {name: {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in w.items()}
                       for name, w in self._tc.items()}
        return new
