"""Thin Python wrappers over the C-ABI entry points (raw device pointers in, nothing
allocated behind the caller's back).  Used by the drop-in modules and by the tests."""
from __future__ import annotations

import ctypes as C

import torch

from . import _ext


def conv_src(t: torch.Tensor, taps: int = 1, dilation: int = 0, center: int = 0, channel_major: bool = False):
    """Describe one GEMM input.  ``t`` is (B, T, C) channels-last, or (B, C, T) when
    ``channel_major`` (the layout the reference hands to Tacotron2.inference)."""
    _ext.require_cuda(t, "conv source")
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise _ext.FacError("conv sources must be contiguous fp32 tensors")
    if channel_major and t.shape[2] == 1:       # (B, C, 1) is the same memory as channels-last (B, 1, C)
        t, channel_major = t.view(t.shape[0], 1, t.shape[1]), False
    if channel_major:
        B, Cc, T = t.shape
        src = _ext.ConvSrc(t.data_ptr(), Cc * T, 1, T, Cc, taps, dilation, center, T, 0)
    else:
        B, T, Cc = t.shape
        src = _ext.ConvSrc(t.data_ptr(), T * Cc, Cc, 1, Cc, taps, dilation, center, T, 0)
    src.keepalive = t   # the descriptor only holds a raw pointer: pin the tensor to it
    return src


def conv_gemm(srcs, w_packed, bias, n_out, out, *, batch, rows, kind=_ext.EPI_LINEAR, act=_ext.ACT_NONE,
              mask=None, residual=None, out2=None, n_split=0, accumulate_out2=False, out_batch_stride=None,
              out_row_stride=None, phases=1, w_phase_stride=0, out_phase_stride=0, row_lengths=None):
    """out[b, t, :n_out] = epilogue(sum over sources/taps/channels ... ) -- fac_conv_gemm_f32."""
    lib = _ext.load()
    arr = (_ext.ConvSrc * len(srcs))(*srcs)
    width = out.shape[-1]
    epi = _ext.ConvEpilogue(kind, act, out.data_ptr(),
                            out_batch_stride if out_batch_stride is not None else rows * width,
                            out_row_stride if out_row_stride is not None else width,
                            _ext.ptr(mask), _ext.ptr(residual), _ext.ptr(out2), n_split, int(accumulate_out2),
                            _ext.ptr(row_lengths))
    rc = lib.fac_conv_gemm_f32(arr, len(srcs), w_packed.data_ptr(), _ext.ptr(bias), batch, rows, n_out,
                               C.byref(epi), phases, w_phase_stride, out_phase_stride, _ext.current_stream())
    _ext.check(rc, "fac_conv_gemm_f32")
    return out


def pack_gemm_weight(w_kn: torch.Tensor, bias=None):
    """(K, N) weight [+ (N,) bias] -> zero-padded (K, N_pad) / (N_pad,) with N_pad % 128 == 0."""
    K, N = w_kn.shape
    n_pad = (N + 127) // 128 * 128
    w = w_kn.new_zeros(K, n_pad)
    w[:, :N] = w_kn
    b = None
    if bias is not None:
        b = bias.new_zeros(n_pad)
        b[:N] = bias
    return w.contiguous(), b


# ---------------------------------------------------------------------- tensor-core Conv1d / Linear
def split_pair(shape, device, nsplit=2, dtype=torch.float16):
    """Uninitialised 16-bit (hi, lo) operand buffers; lo is None in the single-operand mode."""
    hi = torch.empty(*shape, dtype=dtype, device=device)
    return hi, (torch.empty_like(hi) if nsplit == 2 else None)


def transpose_split(x_cm: torch.Tensor, pad: int, nsplit=2, dtype=torch.float16):
    """(B, C, T) channel-major fp32 -> (B, T, pad) channels-last 16-bit hi/lo (fac_transpose_split_16)."""
    _ext.require_cuda(x_cm, "transpose_split input")
    B, Cc, T = x_cm.shape
    hi, lo = split_pair((B, T, pad), x_cm.device, nsplit, dtype)
    rc = _ext.load().fac_transpose_split_16(x_cm.data_ptr(), hi.data_ptr(), _ext.ptr(lo), B, Cc, T, pad,
                                            int(dtype == torch.float16), _ext.current_stream())
    _ext.check(rc, "fac_transpose_split_16")
    return hi, lo


def pad_split(x_cl: torch.Tensor, pad: int, nsplit=2, dtype=torch.float16):
    """(B, T, C) channels-last fp32 -> (B, T, pad) 16-bit hi/lo, padding channels zero (fac_pad_split_16)."""
    _ext.require_cuda(x_cl, "pad_split input")
    B, T, Cc = x_cl.shape
    hi, lo = split_pair((B, T, pad), x_cl.device, nsplit, dtype)
    rc = _ext.load().fac_pad_split_16(x_cl.data_ptr(), hi.data_ptr(), _ext.ptr(lo), B * T, Cc, pad,
                                      int(dtype == torch.float16), _ext.current_stream())
    _ext.check(rc, "fac_pad_split_16")
    return hi, lo


TC_K_CHUNK = 512   # contraction elements per tensor-core accumulation chain (see fac_tc_conv.k_chunk)


def conv_gemm_tc(a, w, *, act=_ext.ACT_NONE, mask=None, residual=None, out=None, want_split=True, nsplit=2,
                 k_chunk=None, row_lengths=None):
    """Conv1d / Linear on the tensor cores (fac_conv_gemm_tc).  ``a`` = (hi, lo) 16-bit (B, T, c_pad) from
    transpose_split/pad_split or a previous call; ``w`` = one entry of PackedTacotron.tc_weights() (same
    16-bit type).  ``row_lengths`` (int32 [B], optional): rows beyond an utterance's length are written as
    zeros.  Returns (out_f32 or None, (hi, lo) or None)."""
    a_hi, a_lo = a
    B, T, c_pad = a_hi.shape
    if c_pad != w["c_pad"]:
        raise _ext.FacError("conv_gemm_tc: input has %d channels, weight expects %d" % (c_pad, w["c_pad"]))
    if a_hi.dtype != w["hi"].dtype:
        raise _ext.FacError("conv_gemm_tc: operand types differ (%s vs %s)" % (a_hi.dtype, w["hi"].dtype))
    nxt = split_pair((B, T, w["n_pad"]), a_hi.device, nsplit, a_hi.dtype) if want_split else (None, None)
    ld = lambda t: 0 if t is None else t.shape[-1]  # noqa: E731
    k_chunk = TC_K_CHUNK if k_chunk is None else k_chunk
    scratch = None
    if k_chunk and w["taps"] * c_pad > k_chunk:
        scratch = torch.empty(B * T, w["n_valid"], dtype=torch.float32, device=a_hi.device)
    d = _ext.TcConv(a_hi.data_ptr(), _ext.ptr(a_lo), w["hi"].data_ptr(), w["lo"].data_ptr() if nsplit == 2 else None,
                    _ext.ptr(w["bias"]), _ext.ptr(mask), _ext.ptr(residual), _ext.ptr(out), _ext.ptr(nxt[0]),
                    _ext.ptr(nxt[1]), ld(mask), ld(residual), ld(out), B, T, c_pad, w["taps"],
                    w.get("center_override", (w["taps"] - 1) // 2), w["n_pad"], w["n_valid"], act, nsplit, int(a_hi.dtype == torch.float16),
                    k_chunk or 0, 0, _ext.ptr(scratch), _ext.ptr(row_lengths))
    rc = _ext.load().fac_conv_gemm_tc(C.byref(d), _ext.current_stream())
    _ext.check(rc, "fac_conv_gemm_tc")
    return out, (nxt if want_split else None)


# ---------------------------------------------------------------------- pruned posteriorgrams
class SparsePPG:
    """A posteriorgram as per-frame (index, value) lists: ``indices`` (B, T, k) int32, ``values`` (B, T, k) fp32,
    padded with (0, 0.0); ``n_symbols`` = the dense width (5816).  What ``Tacotron2.inference`` accepts in place
    of the dense (B, n_symbols, T) tensor: k * 8 bytes per frame instead of 23 KB."""

    def __init__(self, indices, values, n_symbols):
        if indices.shape != values.shape or indices.dim() != 3:
            raise ValueError("indices and values must both be (B, T, k)")
        if indices.shape[2] > 64:
            raise ValueError("at most 64 entries per frame")
        self.indices = indices.to(torch.int32).contiguous()
        self.values = values.to(torch.float32).contiguous()
        self.n_symbols = int(n_symbols)

    @property
    def shape(self):
        return self.indices.shape

    @property
    def device(self):
        return self.values.device

    def to(self, device, non_blocking=False):
        return SparsePPG(self.indices.to(device, non_blocking=non_blocking),
                         self.values.to(device, non_blocking=non_blocking), self.n_symbols)

    def pin_memory(self):
        return SparsePPG(self.indices.pin_memory(), self.values.pin_memory(), self.n_symbols)

    def dense(self):
        """(B, n_symbols, T) dense tensor holding exactly the listed entries (tests, oracle input)."""
        B, T, k = self.indices.shape
        out = torch.zeros(B, T, self.n_symbols, dtype=torch.float32, device=self.values.device)
        out.scatter_add_(2, self.indices.long(), self.values)
        return out.transpose(1, 2).contiguous()

    @classmethod
    def from_dense_host(cls, ppg, k=64, threshold=0.0):
        """Host-side pruning of a dense (B, n_symbols, T) posteriorgram (numpy / CPU tensor): per frame the k
        largest entries, of those only the ones > threshold, in ascending channel order."""
        x = torch.as_tensor(ppg, dtype=torch.float32).cpu()
        B, D, T = x.shape
        vals, idx = x.transpose(1, 2).topk(min(k, D), dim=2)
        vals = torch.where(vals > threshold, vals, torch.zeros_like(vals))
        idx = torch.where(vals > 0, idx, torch.zeros_like(idx))
        order = torch.argsort(torch.where(vals > 0, idx, torch.full_like(idx, D)), dim=2)
        return cls(idx.gather(2, order), vals.gather(2, order), D)


def sparsify_ppg(ppg: torch.Tensor, k: int = 64, threshold: float = 1e-4) -> SparsePPG:
    """Dense (B, n_symbols, T) posteriorgram on the GPU -> SparsePPG holding the entries > threshold
    (fac_ppg_sparsify).  Raises FacError when a frame has more than k such entries: nothing is truncated
    silently -- raise the threshold, or keep the dense path."""
    _ext.require_cuda(ppg, "ppg")
    x = ppg.float().contiguous()
    B, D, T = x.shape
    idx = torch.empty(B, T, k, dtype=torch.int32, device=x.device)
    val = torch.empty(B, T, k, dtype=torch.float32, device=x.device)
    overflow = torch.zeros(1, dtype=torch.int32, device=x.device)
    rc = _ext.load().fac_ppg_sparsify(x.data_ptr(), idx.data_ptr(), val.data_ptr(), overflow.data_ptr(), B, D, T, k,
                                      float(threshold), _ext.current_stream())
    _ext.check(rc, "fac_ppg_sparsify")
    n_over = int(overflow.item())
    if n_over:
        raise _ext.FacError("%d of %d frames have more than %d entries above %g: raise the threshold or use the "
                            "dense path" % (n_over, B * T, k, threshold))
    return SparsePPG(idx, val, D)


def prenet0_sparse(sp: SparsePPG, w_t: torch.Tensor, E: int, *, mask=None, row_lengths=None, out=None, pad=None,
                   want_split=True):
    """First encoder prenet layer on a SparsePPG (fac_prenet0_sparse_f32).  ``w_t`` = packed enc.pre0_w
    (n_symbols, ld).  Returns (out_f32 or None, (hi, lo) fp16 operand copies or None)."""
    _ext.require_cuda(sp.values, "sparse PPG")
    B, T, k = sp.indices.shape
    pad = pad or (E + 63) // 64 * 64
    nxt = split_pair((B, T, pad), sp.values.device, 2, torch.float16) if want_split else (None, None)
    rc = _ext.load().fac_prenet0_sparse_f32(sp.indices.data_ptr(), sp.values.data_ptr(), w_t.data_ptr(), w_t.shape[1],
                                            _ext.ptr(mask), _ext.ptr(row_lengths), _ext.ptr(out),
                                            0 if out is None else out.shape[-1], _ext.ptr(nxt[0]), _ext.ptr(nxt[1]),
                                            B, T, k, sp.n_symbols, E, pad, _ext.current_stream())
    _ext.check(rc, "fac_prenet0_sparse_f32")
    return out, (nxt if want_split else None)
