"""Drop-in for the reference ``src/common/model.py`` (inference side of the PPG->Mel model).

Module tree, constructor arguments and parameter names mirror the reference so
``load_state_dict(torch.load(ckpt)['state_dict'])`` (generate_synthesis.py:81-83) works
unchanged; ``Tacotron2.inference(inputs)`` keeps its signature and return list
``[mel, mel_postnet, gate, alignments]`` (reference model.py:597-610).  The modules hold
parameters only.  All arithmetic runs in libfacb200.so:

  encoder prenet / convs+BN / LSTM input projection / postnet
      -> fac_conv_gemm_tc (tcgen05 tensor cores, split-fp16 operands, K-chunked fp32 accumulation; BN folded,
         fused ReLU/tanh/dropout mask), or fac_conv_gemm_f32 (exact FFMA) with set_precision('fp32')
  memory layer              -> fac_conv_gemm_f32
  encoder BiLSTM recurrence -> fac_lstm_bidir_f32 (8-CTA clusters, W_hh resident in DSMEM, mma.sync)
  decoder loop              -> fac_taco_decoder_run (one persistent cooperative kernel: matrix CTAs + one
                               attention CTA per utterance)

Differences from the reference that a caller can observe:
  * B > 1 works (the reference's stop test only supports B == 1, model.py:524): every
    utterance stops on its own gate; frames after an utterance's stop are zeroed and the
    per-utterance lengths are left in ``model.last_output_lengths``.  A batch is defined as B
    independent B == 1 runs: ``inference(inputs, input_lengths=...)`` accepts a ragged batch
    (zero-padded to the longest input; the reference builds ``input_lengths`` at model.py:599 and
    masks per utterance in utils.py:46-78): convolutions see each utterance's own zero padding, the
    reverse LSTM direction starts at each utterance's own last frame, the attention window is
    clamped to each utterance's length, and the postnet stops at each utterance's own last frame.
  * the always-on prenet dropout (model.py:132-135) draws its masks from torch's generator
    on the input's device.  ``rng_mode='reference'`` replays the reference's draw order
    call by call (seed-exact); the default ``'fast'`` makes one draw per tensor kind.
    A recorded tape can be injected with ``inference(inputs, dropout_tape=...)``.
  * ``inference`` also accepts a pruned posteriorgram (``ops.SparsePPG``: per-frame (index, value) lists, or a
    dense input pruned on the GPU when ``model.ppg_prune = (k, threshold)`` is set): the first prenet layer then
    gathers k weight rows per frame instead of running the K = 5816 GEMM, and a host caller ships k * 8 bytes per
    frame instead of 23 KB.  Exact for the pruned posteriorgram; pruning itself perturbs the first layer's
    pre-activations by at most (dropped probability mass) x max |W|.
The training direction (``forward`` / ``parse_batch``) is out of scope and raises.
"""
from __future__ import annotations

import ctypes as C
import sys

import torch
from torch import nn
from torch.nn import functional as F

from fac_via_ppg_b200 import _ext, ops
from fac_via_ppg_b200.packing import PackedTacotron
from fac_via_ppg_b200.common.layers import ConvNorm, LinearNorm

_HP_KEYS = ("n_symbols", "symbols_embedding_dim", "encoder_embedding_dim", "encoder_kernel_size",
            "encoder_n_convolutions", "n_acoustic_feat_dims", "prenet_dim", "attention_rnn_dim", "decoder_rnn_dim",
            "attention_dim", "attention_location_n_filters", "attention_location_kernel_size",
            "attention_window_size", "postnet_embedding_dim", "postnet_kernel_size", "postnet_n_convolutions",
            "gate_threshold", "max_decoder_steps", "p_attention_dropout", "p_decoder_dropout")


def _no_forward(name):
    def forward(self, *a, **k):
        raise NotImplementedError("fac_via_ppg_b200: %s is a parameter holder; use Tacotron2.inference" % name)
    return forward


class LocationLayer(nn.Module):
    def __init__(self, attention_n_filters, attention_kernel_size, attention_dim):
        super().__init__()
        self.location_conv = ConvNorm(2, attention_n_filters, kernel_size=attention_kernel_size,
                                      padding=(attention_kernel_size - 1) // 2, bias=False)
        self.location_dense = LinearNorm(attention_n_filters, attention_dim, bias=False, w_init_gain="tanh")

    forward = _no_forward("LocationLayer")


class Attention(nn.Module):
    def __init__(self, attention_rnn_dim, embedding_dim, attention_dim, attention_location_n_filters,
                 attention_location_kernel_size):
        super().__init__()
        self.query_layer = LinearNorm(attention_rnn_dim, attention_dim, bias=False, w_init_gain="tanh")
        self.memory_layer = LinearNorm(embedding_dim, attention_dim, bias=False, w_init_gain="tanh")
        self.v = LinearNorm(attention_dim, 1, bias=False)
        self.location_layer = LocationLayer(attention_location_n_filters, attention_location_kernel_size,
                                            attention_dim)
        self.score_mask_value = -float("inf")

    forward = _no_forward("Attention")


class Prenet(nn.Module):
    def __init__(self, in_dim, sizes):
        super().__init__()
        dims = [in_dim] + list(sizes)
        self.layers = nn.ModuleList(LinearNorm(a, b, bias=False) for a, b in zip(dims[:-1], dims[1:]))

    forward = _no_forward("Prenet")


def _conv_bn(c_in, c_out, k, gain):
    return nn.Sequential(ConvNorm(c_in, c_out, kernel_size=k, padding=(k - 1) // 2, w_init_gain=gain),
                         nn.BatchNorm1d(c_out))


class Postnet(nn.Module):
    def __init__(self, hparams):
        super().__init__()
        n, k = hparams.postnet_n_convolutions, hparams.postnet_kernel_size
        dims = [hparams.n_acoustic_feat_dims] + [hparams.postnet_embedding_dim] * (n - 1) + \
               [hparams.n_acoustic_feat_dims]
        self.convolutions = nn.ModuleList(
            _conv_bn(dims[i], dims[i + 1], k, "tanh" if i < n - 1 else "linear") for i in range(n))

    forward = _no_forward("Postnet")


class Encoder(nn.Module):
    def __init__(self, hparams):
        super().__init__()
        E = hparams.encoder_embedding_dim
        self.prenet = Prenet(hparams.n_symbols, [hparams.symbols_embedding_dim, hparams.symbols_embedding_dim])
        self.convolutions = nn.ModuleList(
            _conv_bn(E, E, hparams.encoder_kernel_size, "relu") for _ in range(hparams.encoder_n_convolutions))
        self.lstm = nn.LSTM(E, E // 2, 1, batch_first=True, bidirectional=True)

    forward = _no_forward("Encoder")


class Decoder(nn.Module):
    def __init__(self, hparams):
        super().__init__()
        for key in ("n_acoustic_feat_dims", "encoder_embedding_dim", "attention_rnn_dim", "decoder_rnn_dim",
                    "prenet_dim", "max_decoder_steps", "gate_threshold", "p_attention_dropout", "p_decoder_dropout",
                    "attention_window_size"):
            setattr(self, key, getattr(hparams, key))
        E = hparams.encoder_embedding_dim
        self.prenet = Prenet(hparams.n_acoustic_feat_dims, [hparams.prenet_dim, hparams.prenet_dim])
        self.attention_rnn = nn.LSTMCell(hparams.prenet_dim + E, hparams.attention_rnn_dim)
        self.attention_layer = Attention(hparams.attention_rnn_dim, E, hparams.attention_dim,
                                         hparams.attention_location_n_filters,
                                         hparams.attention_location_kernel_size)
        self.decoder_rnn = nn.LSTMCell(hparams.attention_rnn_dim + E, hparams.decoder_rnn_dim, 1)
        self.linear_projection = LinearNorm(hparams.decoder_rnn_dim + E, hparams.n_acoustic_feat_dims)
        self.gate_layer = LinearNorm(hparams.decoder_rnn_dim + E, 1, bias=True, w_init_gain="sigmoid")

    forward = _no_forward("Decoder")


class Tacotron2(nn.Module):
    """Reference-compatible PPG->Mel model (reference model.py:538-610), CUDA-native inference."""

    rng_mode = "fast"            # 'fast' | 'reference'  (see module docstring)
    precision = "fp16x3"         # encoder / postnet GEMMs: 'fp16x3' = split-fp16 (IEEE half hi + lo) on the tcgen05 tensor cores
                                 # (fp32-grade, 3 UMMAs per product) | 'fp32' = exact FFMA implicit GEMM
    collect_timing = False       # True: CUDA-event times of encoder / decoder / postnet in .last_timing (ms)
    # True: dense (B, T_out, T_in) like the reference; False: None (saves 273 MB per 60 s utterance); 'window': the
    # sparse form -- a pair (weights (B, T_out, 2w+1), start (B, T_out)): step t attends input positions
    # start[b, t] + [0, 2w], the only ones the attention window leaves unmasked (reference utils.py:46-78)
    return_alignments = True
    max_utterances_per_launch = 36      # decoder launches beyond this are split (see _decode)
    ppg_prune = None             # (k, threshold): prune dense inputs on the GPU and use the gather prenet (see above)

    def __init__(self, hparams):
        super().__init__()
        self.mask_padding = hparams.mask_padding
        self.fp16_run = hparams.fp16_run
        self.n_acoustic_feat_dims = hparams.n_acoustic_feat_dims
        self.hp = {k: getattr(hparams, k) for k in _HP_KEYS}
        self.encoder = Encoder(hparams)
        self.decoder = Decoder(hparams)
        self.postnet = Postnet(hparams)
        self.last_output_lengths = None

    # ------------------------------------------------------------------ packing
    def _weights_signature(self):
        return tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))

    def packed(self) -> PackedTacotron:
        sig = self._weights_signature()
        cache = getattr(self, "_fac_packed", None)
        if cache is None or cache[0] != sig:
            w = self.decoder.linear_projection.linear_layer.weight
            _ext.require_cuda(w, "Tacotron2 parameters")
            cache = (sig, PackedTacotron.from_state(self.state_dict(), self.hp, w.device))
            object.__setattr__(self, "_fac_packed", cache)
        return cache[1]

    def empty_packed(self) -> PackedTacotron:
        return PackedTacotron(self.hp, self.decoder.linear_projection.linear_layer.weight.device)

    def use_packed(self, packed: PackedTacotron):
        object.__setattr__(self, "_fac_packed", (self._weights_signature(), packed))

    # ------------------------------------------------------------------ dropout masks
    def _dropout_masks(self, B, T_in, n_steps, dev, tape):
        """Returns (enc0, enc1) float masks in {0,2} of shape (B,T_in,E) and the decoder mask
        tensor (n_steps, 2, B, P) uint8 in {0,1}; draw order = reference call order."""
        E, P = self.hp["symbols_embedding_dim"], self.hp["prenet_dim"]
        if tape is not None:
            tape = [m.to(dev) for m in tape]
            if len(tape) < 2 + 2 * n_steps:
                raise ValueError("dropout tape holds %d masks, need %d" % (len(tape), 2 + 2 * n_steps))
            enc = [(m.float() > 0).float().mul_(2.0).view(B, T_in, E).contiguous() for m in tape[:2]]
            dec = torch.stack([(m.float() > 0).view(B, P) for m in tape[2:2 + 2 * n_steps]])
            return enc[0], enc[1], dec.view(n_steps, 2, B, P).to(torch.uint8).contiguous()
        if self.rng_mode == "reference":
            ones_e = torch.ones(B, T_in, E, device=dev)
            enc0, enc1 = F.dropout(ones_e, 0.5, True), F.dropout(ones_e, 0.5, True)
            ones_d = torch.ones(B, P, device=dev)
            dec = torch.stack([F.dropout(ones_d, 0.5, True) for _ in range(2 * n_steps)])
            return enc0, enc1, (dec > 0).view(n_steps, 2, B, P).to(torch.uint8).contiguous()
        enc = (torch.rand(2, B, T_in, E, device=dev) >= 0.5).float().mul_(2.0)
        dec = (torch.rand(n_steps, 2, B, P, device=dev) >= 0.5).to(torch.uint8)
        return enc[0], enc[1], dec

    # ------------------------------------------------------------------ stages
    def set_precision(self, precision):
        if precision not in ("fp16x3", "fp32"):
            raise ValueError("precision must be 'fp16x3' or 'fp32', got %r" % (precision,))
        self.precision = precision
        return self

    def _encode_tc(self, packed, inputs, enc0, enc1, lens):
        """Encoder.inference with every GEMM on the tensor cores (split-fp16): the PPG is transposed and split
        once, each layer writes the fp16 hi/lo operand copies of the next one directly.  ``lens`` (int32 [B] or
        None): rows beyond an utterance's length are written as zeros by every layer, which is the zero padding
        the next Conv1d would see if the utterance were processed alone."""
        hp = self.hp
        B, T = (inputs.shape[0], inputs.shape[1]) if isinstance(inputs, ops.SparsePPG) else (inputs.shape[0], inputs.shape[2])
        E = hp["encoder_embedding_dim"]
        tw = packed.tc_weights()
        if isinstance(inputs, ops.SparsePPG):
            _, a = ops.prenet0_sparse(inputs, packed.view("enc.pre0_w"), E, mask=enc0, row_lengths=lens,
                                      pad=tw["enc.pre1"]["c_pad"])
        else:
            a = ops.transpose_split(inputs, tw["enc.pre0"]["c_pad"])
            _, a = ops.conv_gemm_tc(a, tw["enc.pre0"], act=_ext.ACT_RELU, mask=enc0, row_lengths=lens)
        _, a = ops.conv_gemm_tc(a, tw["enc.pre1"], act=_ext.ACT_RELU, mask=enc1, row_lengths=lens)
        for i in range(hp["encoder_n_convolutions"]):
            _, a = ops.conv_gemm_tc(a, tw[f"enc.conv{i}"], act=_ext.ACT_RELU, row_lengths=lens)
        xp = torch.empty(B, T, 4 * E, device=a[0].device, dtype=torch.float32)
        ops.conv_gemm_tc(a, tw["enc.lstm_ih"], out=xp, want_split=False)
        return xp

    def _encode(self, packed, inputs, enc0, enc1, lens=None):
        """reference model.py:237-249 (Encoder.inference): (B, D, T) -> memory (B, T, E); rows beyond an
        utterance's length (ragged batch) are zero."""
        hp = self.hp
        sparse = isinstance(inputs, ops.SparsePPG)
        B, T = (inputs.shape[0], inputs.shape[1]) if sparse else (inputs.shape[0], inputs.shape[2])
        E, H = hp["encoder_embedding_dim"], hp["encoder_embedding_dim"] // 2
        dev = inputs.values.device if sparse else inputs.device
        new = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)  # noqa: E731
        if self.precision == "fp16x3":
            xp = self._encode_tc(packed, inputs, enc0, enc1, lens)
        else:
            if sparse:
                h, _ = ops.prenet0_sparse(inputs, packed.view("enc.pre0_w"), E, mask=enc0, row_lengths=lens,
                                          out=new(B, T, E), want_split=False)
            else:
                h = ops.conv_gemm([ops.conv_src(inputs, channel_major=True)], packed.view("enc.pre0_w"), None, E,
                                  new(B, T, E), batch=B, rows=T, act=_ext.ACT_RELU, mask=enc0, row_lengths=lens)
            h = ops.conv_gemm([ops.conv_src(h)], packed.view("enc.pre1_w"), None, E, new(B, T, E), batch=B, rows=T,
                              act=_ext.ACT_RELU, mask=enc1, row_lengths=lens)
            k = hp["encoder_kernel_size"]
            for i in range(hp["encoder_n_convolutions"]):
                h = ops.conv_gemm([ops.conv_src(h, k, 1, (k - 1) // 2)], packed.view(f"enc.conv{i}_w"),
                                  packed.view(f"enc.conv{i}_b"), E, new(B, T, E), batch=B, rows=T, act=_ext.ACT_RELU,
                                  row_lengths=lens)
            xp = ops.conv_gemm([ops.conv_src(h)], packed.view("enc.lstm_ih_w"), packed.view("enc.lstm_ih_b"), 8 * H,
                               new(B, T, 8 * H), batch=B, rows=T)
        memory = new(B, T, E) if lens is None else torch.zeros(B, T, E, device=dev, dtype=torch.float32)
        rc = _ext.load().fac_lstm_bidir_var_f32(xp.data_ptr(), packed.view("enc.lstm_hh").data_ptr(),
                                                memory.data_ptr(), _ext.ptr(lens), B, T, H, _ext.current_stream())
        _ext.check(rc, "fac_lstm_bidir_var_f32")
        return memory

    def _decode(self, packed, memory, dec_masks, n_steps, lens=None):
        """reference model.py:489-535 (Decoder.inference) -> mel_cl (B, n_steps, M), gate, align, lengths.
        One launch can decode (SMs - 100) utterances (one attention CTA each next to >= 100 matrix CTAs), but
        beyond 36 a matrix CTA takes them in passes of 8 instead of 16 (csrc/tacotron_decoder.cu), which is slower
        per utterance than two launches: larger batches run as consecutive groups of 32.  A ragged batch is length-sorted into the groups
        (src/common/data_utils.py sorts training batches the same way), so that a group of short utterances
        retires as soon as its own longest member has fired its stop gate; an utterance's result does not
        depend on the group it runs in (the decoder's arithmetic is batch-invariant)."""
        B = memory.shape[0]
        limit = max(1, min(torch.cuda.get_device_properties(memory.device).multi_processor_count - 100,
                           self.max_utterances_per_launch))
        group = min(limit, 32)
        if B <= limit:
            return self._decode_group(packed, memory, dec_masks, n_steps, lens)
        order = torch.argsort(lens, descending=True, stable=True) if lens is not None else None
        if order is not None:
            memory, dec_masks, lens = memory[order], dec_masks[:, :, order], lens[order]
        parts = [self._decode_group(packed, memory[i:i + group].contiguous(),
                                    dec_masks[:, :, i:i + group].contiguous(), n_steps,
                                    None if lens is None else lens[i:i + group].contiguous())
                 for i in range(0, B, group)]
        mel, gate, align, out_len, done = zip(*parts)
        if align[0] is None:
            align_all = None
        elif isinstance(align[0], tuple):
            align_all = (torch.cat([a[0] for a in align]), torch.cat([a[1] for a in align]))
        else:
            align_all = torch.cat(align)
        out = [torch.cat(mel), torch.cat(gate), align_all, torch.cat(out_len)]
        if order is not None:
            inv = torch.empty_like(order)
            inv[order] = torch.arange(B, device=order.device)
            pick = lambda t: None if t is None else (tuple(x[inv] for x in t) if isinstance(t, tuple) else t[inv])  # noqa: E731
            out = [pick(t) for t in out]
        return (*out, torch.stack(done).sum(0))

    def _decode_group(self, packed, memory, dec_masks, n_steps, lens=None):
        hp = self.hp
        B, T, E = memory.shape
        dev = memory.device
        A, R, M = hp["attention_dim"], hp["attention_rnn_dim"], hp["n_acoustic_feat_dims"]
        pmem = ops.conv_gemm([ops.conv_src(memory)], packed.view("dec.mem_w"), None, A,
                             torch.empty(B, T, A, device=dev), batch=B, rows=T)
        z = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)  # noqa: E731
        st = {"c_att": z(B, R), "c_dec": z(B, R),
              # every vector that crosses CTAs, as (value, version) 8-byte words, two copies (include/fac_b200.h)
              "xchg": torch.zeros(2 * _ext.TACO_XCHG_WORDS * B + _ext.TACO_XCHG_HINTS, dtype=torch.int64, device=dev),
              "w_prev": z(B, T), "w_cum": z(B, T),
              "done": torch.zeros(8, dtype=torch.int32, device=dev),
              "out_len": torch.zeros(B, dtype=torch.int32, device=dev)}
        lengths = lens if lens is not None else torch.full((B,), T, dtype=torch.int32, device=dev)   # model.py:599
        mel = z(B, n_steps, M)
        gate = z(B, n_steps)
        align = z(B, n_steps, T) if self.return_alignments is True else None
        if self.return_alignments == "window":
            st["align_win"] = z(B, n_steps, 2 * hp["attention_window_size"] + 1)
            st["align_start"] = torch.zeros(B, n_steps, dtype=torch.int32, device=dev)
        cstate = _ext.TacoDecoderState(*[_ext.ptr(st.get(n)) for n, _ in _ext.TacoDecoderState._fields_])
        rc = _ext.load().fac_taco_decoder_run(
            C.byref(packed.cdecoder), memory.data_ptr(), pmem.data_ptr(), lengths.data_ptr(), dec_masks.data_ptr(),
            C.byref(cstate), mel.data_ptr(), gate.data_ptr(), _ext.ptr(align), B, T, n_steps,
            hp["attention_window_size"], float(self.decoder.gate_threshold), _ext.current_stream())
        _ext.check(rc, "fac_taco_decoder_run")
        if self.return_alignments == "window":
            align = (st["align_win"], st["align_start"])
        return mel, gate, align, st["out_len"], st["done"]

    def _postnet(self, packed, mel_cl, out_len=None):
        """reference model.py:178-184 + 604-605: mel_post = mel + postnet(mel), channels-last in/out.  ``out_len``
        (int32 [B] or None): every layer writes zeros beyond an utterance's own last frame, so an utterance that
        stopped early sees the Conv1d zero padding it would see alone."""
        hp = self.hp
        B, T, M = mel_cl.shape
        n, k, Pe = hp["postnet_n_convolutions"], hp["postnet_kernel_size"], hp["postnet_embedding_dim"]
        if self.precision == "fp16x3":
            tw = packed.tc_weights()
            a = ops.pad_split(mel_cl, tw["post.conv0"]["c_pad"])
            for i in range(n - 1):
                _, a = ops.conv_gemm_tc(a, tw[f"post.conv{i}"], act=_ext.ACT_TANH, row_lengths=out_len)
            out = torch.empty(B, T, M, device=mel_cl.device, dtype=torch.float32)
            ops.conv_gemm_tc(a, tw[f"post.conv{n - 1}"], residual=mel_cl, out=out, want_split=False,
                             row_lengths=out_len)
            return out
        h = mel_cl
        for i in range(n):
            last = i == n - 1
            width = M if last else Pe
            h = ops.conv_gemm([ops.conv_src(h, k, 1, (k - 1) // 2)], packed.view(f"post.conv{i}_w"),
                              packed.view(f"post.conv{i}_b"), width,
                              torch.empty(B, T, width, device=mel_cl.device), batch=B, rows=T,
                              act=_ext.ACT_NONE if last else _ext.ACT_TANH, residual=mel_cl if last else None,
                              row_lengths=out_len)
        return h

    # ------------------------------------------------------------------ public API
    def parse_input(self, inputs):
        return inputs.half() if self.fp16_run else inputs            # fp16_optimizer.py:53-63

    def parse_output(self, outputs, output_lengths=None):
        if not self.fp16_run:
            return outputs                                          # model.py:566-578 (no masking at inference)
        return [o.float() if torch.is_tensor(o) else o for o in outputs]

    @torch.no_grad()
    def inference(self, inputs, dropout_tape=None, input_lengths=None):
        """inputs (B, n_symbols, T_in) -> [mel (B,M,T_out), mel_postnet, gate (B,T_out,1), alignments (B,T_out,T_in)].

        ``input_lengths`` (optional, B ints <= T_in) makes the batch ragged: utterance b only has its first
        input_lengths[b] frames (the rest is padding and is ignored), and its outputs equal those of a B == 1
        call on inputs[b:b+1, :, :input_lengths[b]] with the same dropout masks."""
        if isinstance(inputs, ops.SparsePPG):
            _ext.require_cuda(inputs.values, "inputs")
            _ext.require_cuda(inputs.indices, "inputs")
            x = inputs
            (B, T, _), D = x.shape, x.n_symbols
        else:
            _ext.require_cuda(inputs, "inputs")
            inputs = self.parse_input(inputs)
            x = inputs.float().contiguous()
            B, D, T = x.shape
        if D != self.hp["n_symbols"]:
            raise ValueError("inputs have %d symbols, model expects %d" % (D, self.hp["n_symbols"]))
        if B == 0 or T == 0:
            raise ValueError("empty input batch")
        if self.ppg_prune is not None and not isinstance(x, ops.SparsePPG):
            x = ops.sparsify_ppg(x, *self.ppg_prune)
        dev = x.values.device if isinstance(x, ops.SparsePPG) else x.device
        lens = None
        if input_lengths is not None:
            lens_host = torch.as_tensor(input_lengths).to("cpu", torch.int64).flatten()
            if lens_host.numel() != B or int(lens_host.min()) < 1 or int(lens_host.max()) > T:
                raise ValueError("input_lengths must hold %d values in [1, %d]" % (B, T))
            if int(lens_host.min()) < T:                     # all full length: the plain path
                lens = lens_host.to(device=dev, dtype=torch.int32)
        packed = self.packed()
        n_steps = int(self.decoder.max_decoder_steps)
        enc0, enc1, dec_masks = self._dropout_masks(B, T, n_steps, dev, dropout_tape)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if self.collect_timing else None
        if ev:
            ev[0].record()
        memory = self._encode(packed, x, enc0, enc1, lens)
        if ev:
            ev[1].record()
        mel_cl, gate, align, out_len, done = self._decode(packed, memory, dec_masks, n_steps, lens)
        if ev:
            ev[2].record()
        out_lens = out_len.cpu()                                      # the one device->host sync of the call
        t_out = int(out_lens.max())
        done_host = done.cpu()
        if int(done_host[7]) > 0:
            raise _ext.FacError("fac_taco_decoder_run: a hand-over between the decoder's CTAs timed out (the GPU is "
                                "shared with another kernel, or the state buffers were not zero-filled)")
        if int(done_host[1]) > 0:
            print("Warning! Reached max decoder steps", file=sys.stderr)   # model.py:526-528 (stderr: keeps stdout machine-readable)
        mel_cl = mel_cl[:, :t_out].contiguous()
        ragged_out = B > 1 and int(out_lens.min()) < t_out
        if ragged_out:                                               # per-utterance stop: zero the tail
            keep = (torch.arange(t_out, device=dev)[None, :] < out_len[:, None]).unsqueeze(-1)
            mel_cl = mel_cl * keep
            gate = gate[:, :t_out] * keep[..., 0]
            if isinstance(align, tuple):
                align = (align[0][:, :t_out] * keep, align[1][:, :t_out] * keep[..., 0].to(torch.int32))
            elif align is not None:
                align = align[:, :t_out] * keep
        self.last_output_lengths = out_lens
        post_cl = self._postnet(packed, mel_cl, out_len.contiguous() if ragged_out else None)
        if ev:
            ev[3].record()
            torch.cuda.synchronize()
            self.last_timing = {"encoder_ms": ev[0].elapsed_time(ev[1]), "decoder_ms": ev[1].elapsed_time(ev[2]),
                                "postnet_ms": ev[2].elapsed_time(ev[3]), "decoder_steps": t_out}
        if isinstance(align, tuple):
            align_out = (align[0][:, :t_out], align[1][:, :t_out])
        else:
            align_out = align[:, :t_out] if align is not None else None
        outputs = [mel_cl.transpose(1, 2), post_cl.transpose(1, 2), gate[:, :t_out].unsqueeze(-1), align_out]
        return self.parse_output(outputs)

    def forward(self, inputs):
        raise NotImplementedError("fac_via_ppg_b200 covers Tacotron2.inference only; teacher-forced training "
                                  "(reference model.py:580-595) is out of scope")

    def parse_batch(self, batch):
        raise NotImplementedError("training batches are out of scope (reference model.py:547-560)")
