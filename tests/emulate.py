"""CPU emulation (torch) of what the CUDA kernels compute FROM THE PACKED BUFFERS.

Test infrastructure: lets the CPU-only suite check the weight packing, the phase
decomposition of the upsampler, the gate interleave and the slot layout of the
audio buffer against the oracle without a GPU.  Mirrors csrc/gemm_conv_f32.cu and
csrc/waveglow_f32.cu one-to-one in semantics (not in performance).
"""
import torch

from fac_via_ppg_b200.packing import upsample_taps
from fac_via_ppg_b200.synth import flow_channels


def gather_rows(src, taps, dil, center):
    """src (B, T, C) channels-last -> (B, T, taps*C): row t holds rows t + k*dil - center (zero outside)."""
    B, T, C = src.shape
    cols = []
    for k in range(taps):
        shift = k * dil - center
        out = src.new_zeros(B, T, C)
        lo, hi = max(0, -shift), min(T, T - shift)
        if hi > lo:
            out[:, lo:hi] = src[:, lo + shift:hi + shift]
        cols.append(out)
    return torch.cat(cols, dim=-1)


def conv_gemm(srcs, w, bias, n):
    """srcs: list of (tensor (B,T,C), taps, dil, center); w (K, N_pad); returns (B, T, n)."""
    a = torch.cat([gather_rows(*s) for s in srcs], dim=-1)
    out = a @ w[:, :n]
    return out if bias is None else out + bias[:n]


def waveglow_infer(packed, mel, audio):
    """packed: PackedWaveGlow on CPU; mel (B, n_mel, F); audio (B, Tg, G) pre-filled with sigma*z."""
    cfg, lay, flat = packed.cfg, packed.layout, packed.flat
    wn = cfg["WN_config"]
    C, L, ks = wn["n_channels"], wn["n_layers"], wn["kernel_size"]
    G, hop, n_mel = cfg["n_group"], cfg["hop_length"], cfg["n_mel_channels"]
    n_cond, phases, taps = n_mel * G, hop // G, upsample_taps(cfg)
    B, _, F = mel.shape
    Tg = F * phases
    mel_cl = mel.transpose(1, 2).contiguous()
    spect = mel.new_zeros(B, F, phases, n_cond)
    w_up, b_up = lay.view(flat, "upsample_w"), lay.view(flat, "upsample_b")
    for p in range(phases):
        spect[:, :, p] = conv_gemm([(mel_cl, taps, -1, 0)], w_up[p], b_up, n_cond)
    spect = spect.view(B, Tg, n_cond)
    audio = audio.clone()
    chans = flow_channels(cfg)
    for k in reversed(range(cfg["n_flows"])):
        n_rem, n_half = chans[k]
        off = G - n_rem
        x = audio[:, :, off:off + n_half] @ lay.view(flat, f"{k}.start_w") + lay.view(flat, f"{k}.start_b")
        skip = None
        for i in range(L):
            d = 2 ** i
            pre = conv_gemm([(x, ks, d, d * (ks - 1) // 2), (spect, 1, 0, 0)],
                            lay.view(flat, f"{k}.{i}.in_cond_w"), lay.view(flat, f"{k}.{i}.in_cond_b"), 2 * C)
            acts = torch.tanh(pre[..., 0::2]) * torch.sigmoid(pre[..., 1::2])
            n_rs = 2 * C if i < L - 1 else C
            rs = conv_gemm([(acts, 1, 0, 0)], lay.view(flat, f"{k}.{i}.res_skip_w"),
                           lay.view(flat, f"{k}.{i}.res_skip_b"), n_rs)
            if i < L - 1:
                x = x + rs[..., :C]
                s = rs[..., C:]
            else:
                s = rs
            skip = s if skip is None else skip + s
        out = skip @ lay.view(flat, f"{k}.end_w").t() + lay.view(flat, f"{k}.end_b")
        a0 = audio[:, :, off:off + n_half]
        a1 = (audio[:, :, off + n_half:] - out[..., :n_half]) / torch.exp(out[..., n_half:])
        y = torch.cat([a0, a1], dim=-1)
        audio[:, :, off:] = y @ lay.view(flat, f"{k}.w_inv").t()
    return audio.reshape(B, Tg * G)


def fill_audio_slots(noise, sigma, G):
    """Slot layout used by WaveGlow.infer: first draw owns the last slots."""
    B, _, Tg = noise[0].shape
    audio = torch.empty(B, Tg, G)
    hi = G
    for z in noise:
        lo = hi - z.shape[1]
        audio[:, :, lo:hi] = (sigma * z).transpose(1, 2)
        hi = lo
    assert hi == 0
    return audio


# ===================================================================== Tacotron2 from the packed buffer
def tacotron_inference(packed, inputs, masks, n_steps, window, gate_threshold=2.0):
    """Mirror of Tacotron2.inference in fac_via_ppg_b200/common/model.py + csrc/tacotron_*.cu:
    BN-folded convs, hoisted LSTM input projection, windowed attention, packed decoder weights."""
    hp = packed.hp
    v = packed.view
    B, D, T = inputs.shape
    E, H, A, R, M = 600, 300, 150, 300, 80
    x = inputs.transpose(1, 2).contiguous()
    h = torch.relu(conv_gemm([(x, 1, 0, 0)], v("enc.pre0_w"), None, E)) * masks[0]
    h = torch.relu(conv_gemm([(h, 1, 0, 0)], v("enc.pre1_w"), None, E)) * masks[1]
    for i in range(hp["encoder_n_convolutions"]):
        h = torch.relu(conv_gemm([(h, 5, 1, 2)], v(f"enc.conv{i}_w"), v(f"enc.conv{i}_b"), E))
    xp = conv_gemm([(h, 1, 0, 0)], v("enc.lstm_ih_w"), v("enc.lstm_ih_b"), 8 * H)
    w_hh = v("enc.lstm_hh")
    memory = torch.zeros(B, T, E)
    for d in range(2):
        hh, cc = torch.zeros(B, H), torch.zeros(B, H)
        for step in range(T):
            t = step if d == 0 else T - 1 - step
            g = xp[:, t, d * 4 * H:(d + 1) * 4 * H] + hh @ w_hh[d].t()
            i_, f_, g_, o_ = g.chunk(4, dim=-1)
            cc = torch.sigmoid(f_) * cc + torch.sigmoid(i_) * torch.tanh(g_)
            hh = torch.sigmoid(o_) * torch.tanh(cc)
            memory[:, t, d * H:(d + 1) * H] = hh
    pmem = conv_gemm([(memory, 1, 0, 0)], v("dec.mem_w"), None, A)

    def cell(w, b, xin, c):
        g = xin @ w.t() + b
        i_, f_, g_, o_ = g.chunk(4, dim=-1)
        c = torch.sigmoid(f_) * c + torch.sigmoid(i_) * torch.tanh(g_)
        return torch.sigmoid(o_) * torch.tanh(c), c

    h_att, c_att, h_dec, c_dec = (torch.zeros(B, R) for _ in range(4))
    ctx, pre = torch.zeros(B, E), torch.zeros(B, R)
    w_prev, w_cum = torch.zeros(B, T), torch.zeros(B, T)
    mel, gate, align = torch.zeros(B, n_steps, M), torch.zeros(B, n_steps), torch.zeros(B, n_steps, T)
    w_loc = v("dec.w_loc").permute(2, 0, 1)      # packed (2, 31, 32) -> (32, 2, 31)
    for t in range(n_steps):
        h_att, c_att = cell(v("dec.w_att"), v("dec.b_att"), torch.cat([pre, ctx, h_att], -1), c_att)
        start, end = min(max(0, t - window), T - 1), min(t + window, T - 1)
        nw = end - start + 1
        pq = h_att @ v("dec.wq").t()
        cat = torch.zeros(B, 2, nw + 30)
        for q in range(nw + 30):
            pos = start - 15 + q
            if 0 <= pos < T:
                cat[:, 0, q], cat[:, 1, q] = w_prev[:, pos], w_cum[:, pos]
        loc = torch.stack([torch.einsum("fck,bck->bf", w_loc, cat[:, :, q:q + 31]) for q in range(nw)], dim=1)
        e = torch.tanh(pq[:, None, :] + loc @ v("dec.w_ld_t") + pmem[:, start:end + 1]) @ v("dec.v")
        w = torch.softmax(e, dim=1)
        ctx = torch.einsum("bq,bqc->bc", w, memory[:, start:end + 1])
        w_prev = torch.zeros(B, T)
        w_prev[:, start:end + 1] = w
        w_cum[:, start:end + 1] += w
        align[:, t, start:end + 1] = w
        h_dec, c_dec = cell(v("dec.w_dec"), v("dec.b_dec"), torch.cat([h_att, ctx, h_dec], -1), c_dec)
        out = torch.cat([h_dec, ctx], -1) @ v("dec.w_pp").t() + v("dec.b_pp")
        mel[:, t], gate[:, t] = out[:, :M], out[:, M]
        if t + 1 < n_steps:          # rows M+1.. = prenet layer 0 composed with the projection
            p1 = torch.relu(out[:, M + 1:]) * masks[2 + 2 * (t + 1)]
            pre = torch.relu(p1 @ v("dec.w_pre2").t()) * masks[3 + 2 * (t + 1)]
    hpost = mel
    n = hp["postnet_n_convolutions"]
    for i in range(n):
        width = M if i == n - 1 else hp["postnet_embedding_dim"]
        hpost = conv_gemm([(hpost, 5, 1, 2)], v(f"post.conv{i}_w"), v(f"post.conv{i}_b"), width)
        hpost = torch.tanh(hpost) if i < n - 1 else hpost + mel
    return [mel.transpose(1, 2), hpost.transpose(1, 2), gate.unsqueeze(-1), align]


# ===================================================================== tensor-core formulation (csrc/waveglow_tc.cu)
def waveglow_infer_tc(packed, mel, audio):
    """Mirror of the tensor-core path's ALGEBRA with torch fp32 on the CPU: bf16 hi+lo operand
    matrices (summed), skip path collapsed into out8 += Wc acts, residual add through the
    identity block of w2.  Checks packing.tc_weights() without a GPU."""
    cfg, lay, flat = packed.cfg, packed.layout, packed.flat
    packed.tc_weights()
    flat16, flat32, _ = packed._tc
    l16, l32 = packed.tc_layouts()
    wn = cfg["WN_config"]
    C, L, ks = wn["n_channels"], wn["n_layers"], wn["kernel_size"]
    G, hop, n_mel = cfg["n_group"], cfg["hop_length"], cfg["n_mel_channels"]
    n_cond, phases, taps = n_mel * G, hop // G, upsample_taps(cfg)
    B, _, F = mel.shape
    Tg = F * phases
    mel_cl = mel.transpose(1, 2).contiguous()
    spect = mel.new_zeros(B, F, phases, n_cond)
    # upsampler from the tensor-core phase matrices [phases][n_cond][taps*mel_pad] (mel zero-padded)
    pad = packed.mel_pad
    mel_pad = mel_cl.new_zeros(B, F, pad)
    mel_pad[..., :n_mel] = mel_cl
    w_up = l16.view(flat16, "up_hi").float() + l16.view(flat16, "up_lo").float()
    b_up = lay.view(flat, "upsample_b")[:n_cond]
    a_up = gather_rows(mel_pad, taps, -1, 0)
    for p in range(phases):
        spect[:, :, p] = a_up @ w_up[p].t() + b_up
    spect = spect.view(B, Tg, n_cond)
    audio = audio.clone()

    def w16(name):
        return l16.view(flat16, name + "_hi").float() + l16.view(flat16, name + "_lo").float()

    for k in reversed(range(cfg["n_flows"])):
        n_rem, n_half = flow_channels(cfg)[k]
        off = G - n_rem
        x = audio[:, :, off:off + n_half] @ lay.view(flat, f"{k}.start_w") + lay.view(flat, f"{k}.start_b")
        out8 = audio.new_zeros(B, Tg, 8)
        for i in range(L):
            d = 2 ** i
            a = torch.cat([gather_rows(x, ks, d, d * (ks - 1) // 2), spect], dim=-1)
            pre = a @ w16(f"{k}.{i}.w1").t() + lay.view(flat, f"{k}.{i}.in_cond_b")[: 2 * C]
            acts = torch.tanh(pre[..., 0::2]) * torch.sigmoid(pre[..., 1::2])
            out8 = out8 + acts @ l32.view(flat32, f"{k}.{i}.wc").t()
            if i < L - 1:
                x = torch.cat([acts, x], dim=-1) @ w16(f"{k}.{i}.w2").t() + l32.view(flat32, f"{k}.{i}.res_b")
        out = out8 + l32.view(flat32, f"{k}.out_bias")
        a0 = audio[:, :, off:off + n_half]
        a1 = (audio[:, :, off + n_half:] - out[..., :n_half]) / torch.exp(out[..., n_half:2 * n_half])
        audio[:, :, off:] = torch.cat([a0, a1], dim=-1) @ lay.view(flat, f"{k}.w_inv").t()
    return audio.reshape(B, Tg * G)
