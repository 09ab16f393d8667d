"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference PPG->Mel inference.

Functional restatement of ``Tacotron2.inference`` (reference
src/common/model.py:597-610) and everything below it, driven by a plain state
dict.  Pinned against the unmodified reference modules by
tests/test_oracle_pinning.py and tests/golden/ (see oracle/make_golden.py); the
reference's own tests hold no vectors for this path.

The reference draws two kinds of random numbers at inference: Prenet dropout is
*always on* (src/common/model.py:132-135).  ``DropoutTape`` reproduces those
draws: in ``record`` mode it calls ``F.dropout`` with the same shapes in the
same order as the reference (so that, under the same seed, masks are
bit-identical) and stores the masks; in ``replay`` mode it feeds stored masks.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


class DropoutTape:
    """Masks take values {0, 2} (= keep / (1-p) with p = 0.5)."""

    def __init__(self, masks=None):
        self.masks = [] if masks is None else list(masks)
        self.replay = masks is not None
        self.pos = 0

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if self.replay:
            m = self.masks[self.pos].to(x.device)
            self.pos += 1
            return x * m.view_as(x)
        m = F.dropout(torch.ones_like(x), p=0.5, training=True)
        self.masks.append(m)
        return x * m


def record_dropout_tape(batch: int, t_in: int, n_steps: int, enc_dim: int = 600, prenet_dim: int = 300):
    """Draws, from torch's global generator, exactly the masks one reference
    inference() call of ``n_steps`` decoder steps consumes, in the same order
    (encoder prenet x2 then decoder prenet x2 per step)."""
    tape = DropoutTape()
    ones_e = torch.ones(batch, t_in, enc_dim)
    tape(ones_e), tape(ones_e)
    ones_d = torch.ones(batch, prenet_dim)
    for _ in range(n_steps):
        tape(ones_d), tape(ones_d)
    return tape.masks


def prenet(sd, prefix: str, x: torch.Tensor, drop) -> torch.Tensor:
    """reference src/common/model.py:124-135 (Prenet.forward; dropout always on)."""
    for i in range(2):
        x = drop(F.relu(F.linear(x, sd[prefix + f"layers.{i}.linear_layer.weight"])))
    return x


def conv_bn(sd, prefix: str, x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """ConvNorm + BatchNorm1d in eval mode (reference src/common/layers.py:53-71,
    src/common/model.py:143-176 / 199-209)."""
    w = sd[prefix + "0.conv.weight"]
    y = F.conv1d(x, w, sd[prefix + "0.conv.bias"], padding=(w.shape[2] - 1) // 2)
    return F.batch_norm(y, sd[prefix + "1.running_mean"], sd[prefix + "1.running_var"],
                        sd[prefix + "1.weight"], sd[prefix + "1.bias"], False, 0.0, eps)


def lstm_cell(sd, prefix: str, x, h, c, suffix: str = ""):
    """torch.nn.LSTMCell semantics (gate order i, f, g, o), used at
    reference src/common/model.py:400-402, 425-428."""
    gates = F.linear(x, sd[prefix + "weight_ih" + suffix], sd[prefix + "bias_ih" + suffix]) + \
        F.linear(h, sd[prefix + "weight_hh" + suffix], sd[prefix + "bias_hh" + suffix])
    i, f, g, o = gates.chunk(4, dim=-1)
    c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h = torch.sigmoid(o) * torch.tanh(c)
    return h, c


def bilstm(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """torch.nn.LSTM(bidirectional, batch_first, 1 layer), reference
    src/common/model.py:211-213, 246-247.  x: (B, T, E) -> (B, T, 2H)."""
    B, T, _ = x.shape
    H = sd[prefix + "weight_hh_l0"].shape[1]
    outs = []
    for suffix, order in (("_l0", range(T)), ("_l0_reverse", reversed(range(T)))):
        h = x.new_zeros(B, H)
        c = x.new_zeros(B, H)
        ys = [None] * T
        # input projection hoisted out of the loop (same arithmetic)
        xp = F.linear(x, sd[prefix + "weight_ih" + suffix], sd[prefix + "bias_ih" + suffix])
        for t in order:
            gates = xp[:, t] + F.linear(h, sd[prefix + "weight_hh" + suffix], sd[prefix + "bias_hh" + suffix])
            i, f, g, o = gates.chunk(4, dim=-1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
            ys[t] = h
        outs.append(torch.stack(ys, dim=1))
    return torch.cat(outs, dim=-1)


def encoder_inference(sd, hp, x: torch.Tensor, drop) -> torch.Tensor:
    """reference src/common/model.py:237-249 (Encoder.inference).  x: (B, D, T) -> (B, T, E)."""
    x = prenet(sd, "encoder.prenet.", x.transpose(1, 2), drop).transpose(1, 2)
    for i in range(hp["encoder_n_convolutions"]):
        x = F.relu(conv_bn(sd, f"encoder.convolutions.{i}.", x))
    return bilstm(sd, "encoder.lstm.", x.transpose(1, 2))


def window_mask(lengths, window: int, time_step: int, max_len: int) -> torch.Tensor:
    """reference src/common/utils.py:46-78.  True = masked.  Keeps the documented
    quirk: past the end of an utterance its last frame stays unmasked."""
    mask = torch.ones(len(lengths), max_len, dtype=torch.bool)
    for ii, n in enumerate(lengths):
        max_idx = int(n) - 1
        start = min(max(0, time_step - window), max_idx)
        end = min(time_step + window, max_idx)
        if start > end:
            continue
        mask[ii, start:end + 1] = False
    return mask


def attention(sd, query, memory, processed_memory, weights_cat, mask):
    """reference src/common/model.py:78-121 (get_alignment_energies + Attention.forward)
    and :56-60 (LocationLayer.forward)."""
    al = "decoder.attention_layer."
    pq = F.linear(query.unsqueeze(1), sd[al + "query_layer.linear_layer.weight"])
    w_loc = sd[al + "location_layer.location_conv.conv.weight"]
    loc = F.conv1d(weights_cat, w_loc, padding=(w_loc.shape[2] - 1) // 2).transpose(1, 2)
    loc = F.linear(loc, sd[al + "location_layer.location_dense.linear_layer.weight"])
    energies = F.linear(torch.tanh(pq + loc + processed_memory), sd[al + "v.linear_layer.weight"]).squeeze(-1)
    if mask is not None:
        energies = energies.masked_fill(mask, -float("inf"))
    weights = F.softmax(energies, dim=1)
    context = torch.bmm(weights.unsqueeze(1), memory).squeeze(1)
    return context, weights


def decoder_inference(sd, hp, memory: torch.Tensor, lengths, drop, gate_threshold=None, max_decoder_steps=None):
    """reference src/common/model.py:489-535 (Decoder.inference) with :387-442 (decode),
    :304-335 (initialize_decoder_states), :289-302 (go frame), :358-385 (output parsing).

    The stop test (``sigmoid(gate) > threshold``) is only defined for B == 1 in the
    reference (:524); for B > 1 this restatement stops when *all* rows fire, which
    coincides with the reference for B == 1 and for forced-length runs."""
    gate_threshold = hp["gate_threshold"] if gate_threshold is None else gate_threshold
    max_decoder_steps = hp["max_decoder_steps"] if max_decoder_steps is None else max_decoder_steps
    B, T_in, E = memory.shape
    M, R, Rd = hp["n_acoustic_feat_dims"], hp["attention_rnn_dim"], hp["decoder_rnn_dim"]
    dec_in = memory.new_zeros(B, M)
    h_att, c_att = memory.new_zeros(B, R), memory.new_zeros(B, R)
    h_dec, c_dec = memory.new_zeros(B, Rd), memory.new_zeros(B, Rd)
    w_att, w_cum = memory.new_zeros(B, T_in), memory.new_zeros(B, T_in)
    context = memory.new_zeros(B, E)
    processed_memory = F.linear(memory, sd["decoder.attention_layer.memory_layer.linear_layer.weight"])
    mels, gates, aligns = [], [], []
    while True:
        x = prenet(sd, "decoder.prenet.", dec_in, drop)
        mask = None
        if hp["attention_window_size"] is not None:
            mask = window_mask(lengths, hp["attention_window_size"], len(mels), T_in).to(memory.device)
        h_att, c_att = lstm_cell(sd, "decoder.attention_rnn.", torch.cat((x, context), -1), h_att, c_att)
        cat = torch.stack((w_att, w_cum), dim=1)
        context, w_att = attention(sd, h_att, memory, processed_memory, cat, mask)
        w_cum = w_cum + w_att
        h_dec, c_dec = lstm_cell(sd, "decoder.decoder_rnn.", torch.cat((h_att, context), -1), h_dec, c_dec)
        hc = torch.cat((h_dec, context), dim=1)
        mel = F.linear(hc, sd["decoder.linear_projection.linear_layer.weight"],
                       sd["decoder.linear_projection.linear_layer.bias"])
        gate = F.linear(hc, sd["decoder.gate_layer.linear_layer.weight"], sd["decoder.gate_layer.linear_layer.bias"])
        mels.append(mel), gates.append(gate), aligns.append(w_att)
        if bool((torch.sigmoid(gate) > gate_threshold).all()):
            break
        if len(mels) == max_decoder_steps:
            break
        dec_in = mel
    mel_out = torch.stack(mels).transpose(0, 1).contiguous().transpose(1, 2)   # (B, M, T_out)
    gate_out = torch.stack(gates).transpose(0, 1).contiguous()                 # (B, T_out, 1)
    align_out = torch.stack(aligns).transpose(0, 1)                            # (B, T_out, T_in)
    return mel_out, gate_out, align_out


def postnet(sd, hp, x: torch.Tensor) -> torch.Tensor:
    """reference src/common/model.py:178-184 (Postnet.forward, eval mode)."""
    n = hp["postnet_n_convolutions"]
    for i in range(n - 1):
        x = torch.tanh(conv_bn(sd, f"postnet.convolutions.{i}.", x))
    return conv_bn(sd, f"postnet.convolutions.{n - 1}.", x)


def tacotron_inference(sd, hp, inputs: torch.Tensor, dropout_masks=None, gate_threshold=None, max_decoder_steps=None):
    """reference src/common/model.py:597-610 (Tacotron2.inference).

    inputs (B, n_symbols, T_in) -> [mel, mel_postnet, gate, alignments]."""
    drop = DropoutTape(dropout_masks)
    lengths = [inputs.shape[2]] * inputs.shape[0]      # model.py:599
    memory = encoder_inference(sd, hp, inputs, drop)
    mel, gate, align = decoder_inference(sd, hp, memory, lengths, drop, gate_threshold, max_decoder_steps)
    mel_post = mel + postnet(sd, hp, mel)
    return [mel, mel_post, gate, align]
