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
              out_row_stride=None, phases=1, w_phase_stride=0, out_phase_stride=0):
    """out[b, t, :n_out] = epilogue(sum over sources/taps/channels ... ) -- fac_conv_gemm_f32."""
    lib = _ext.load()
    arr = (_ext.ConvSrc * len(srcs))(*srcs)
    width = out.shape[-1]
    epi = _ext.ConvEpilogue(kind, act, out.data_ptr(),
                            out_batch_stride if out_batch_stride is not None else rows * width,
                            out_row_stride if out_row_stride is not None else width,
                            _ext.ptr(mask), _ext.ptr(residual), _ext.ptr(out2), n_split, int(accumulate_out2))
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
