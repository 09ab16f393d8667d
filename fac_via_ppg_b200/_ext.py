"""ctypes binding of libfacb200.so (C ABI declared in include/fac_b200.h).

The product path fails loudly when the library is missing or a call fails --
there is no eager-PyTorch or CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

FAC_MAX_FLOWS = 16
FAC_MAX_LAYERS = 16

ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2
EPI_LINEAR, EPI_GATE, EPI_RES_SKIP = 0, 1, 2

_fp = C.c_void_p  # device pointers cross the boundary as integers


class ConvSrc(C.Structure):
    _fields_ = [("ptr", _fp), ("batch_stride", C.c_longlong), ("row_stride", C.c_longlong),
                ("ch_stride", C.c_longlong), ("channels", C.c_int), ("taps", C.c_int),
                ("dilation", C.c_int), ("center", C.c_int), ("rows", C.c_int), ("_pad", C.c_int)]


class ConvEpilogue(C.Structure):
    _fields_ = [("kind", C.c_int), ("act", C.c_int), ("out", _fp),
                ("out_batch_stride", C.c_longlong), ("out_row_stride", C.c_longlong),
                ("mask", _fp), ("residual", _fp), ("out2", _fp),
                ("n_split", C.c_int), ("accumulate_out2", C.c_int), ("row_lengths", _fp)]


class TcConv(C.Structure):
    _fields_ = [("a_hi", _fp), ("a_lo", _fp), ("w_hi", _fp), ("w_lo", _fp), ("bias", _fp), ("mask", _fp),
                ("residual", _fp), ("out", _fp), ("out_hi", _fp), ("out_lo", _fp),
                ("mask_ld", C.c_longlong), ("res_ld", C.c_longlong), ("out_ld", C.c_longlong),
                ("B", C.c_int), ("T", C.c_int), ("c_pad", C.c_int), ("taps", C.c_int), ("center", C.c_int),
                ("n_pad", C.c_int), ("n_valid", C.c_int), ("act", C.c_int), ("nsplit", C.c_int), ("fp16", C.c_int),
                ("k_chunk", C.c_int), ("_pad", C.c_int), ("scratch", _fp), ("row_lengths", _fp)]


class WgFlow(C.Structure):
    _fields_ = [("n_half", C.c_int), ("n_rem", C.c_int),
                ("start_w", _fp), ("start_b", _fp), ("end_w", _fp), ("end_b", _fp), ("w_inv", _fp),
                ("in_cond_w", _fp * FAC_MAX_LAYERS), ("in_cond_b", _fp * FAC_MAX_LAYERS),
                ("res_skip_w", _fp * FAC_MAX_LAYERS), ("res_skip_b", _fp * FAC_MAX_LAYERS)]


class WgModel(C.Structure):
    _fields_ = [("n_flows", C.c_int), ("n_layers", C.c_int), ("n_channels", C.c_int), ("n_group", C.c_int),
                ("n_mel", C.c_int), ("hop", C.c_int), ("n_early_every", C.c_int), ("n_early_size", C.c_int),
                ("upsample_taps", C.c_int), ("kernel_size", C.c_int),
                ("upsample_w", _fp), ("upsample_b", _fp),
                ("flows", WgFlow * FAC_MAX_FLOWS)]


class WgWorkspace(C.Structure):
    _fields_ = [("spect", _fp), ("x", _fp), ("acts", _fp), ("skip", _fp)]


class WgTcFlow(C.Structure):
    _fields_ = [("w1_hi", _fp * FAC_MAX_LAYERS), ("w1_lo", _fp * FAC_MAX_LAYERS),
                ("w2_hi", _fp * FAC_MAX_LAYERS), ("w2_lo", _fp * FAC_MAX_LAYERS),
                ("w2r_hi", _fp * FAC_MAX_LAYERS), ("w2r_lo", _fp * FAC_MAX_LAYERS),
                ("wc", _fp * FAC_MAX_LAYERS), ("res_b", _fp * FAC_MAX_LAYERS), ("out_bias", _fp)]


class WgTcWeights(C.Structure):
    _fields_ = [("up_hi", _fp), ("up_lo", _fp), ("mel_pad", C.c_int), ("_pad", C.c_int),
                ("flows", WgTcFlow * FAC_MAX_FLOWS)]


class WgTcWorkspace(C.Structure):
    _fields_ = [(n, _fp) for n in ("mel_hi", "mel_lo", "spect_hi", "spect_lo", "x_hi", "x_lo", "acts_hi", "acts_lo",
                                   "out8", "x2_hi", "x2_lo", "flow_sync")]


class TacoDecoderWeights(C.Structure):
    _fields_ = [(n, _fp) for n in ("w_att", "b_att", "w_dec", "b_dec", "wq", "w_loc", "w_ld_t", "v", "w_pp", "b_pp",
                                   "w_pre2")]


class TacoDecoderState(C.Structure):
    _fields_ = [(n, _fp) for n in ("c_att", "c_dec", "xchg", "w_prev", "w_cum", "done", "out_len", "align_win",
                                   "align_start")]


TACO_XCHG_WORDS, TACO_XCHG_HINTS = 1800, 64        # include/fac_b200.h FAC_TACO_XCHG_*


# name -> (restype, argtypes); every symbol include/fac_b200.h declares.
_P = C.POINTER
SIGNATURES = {
    "fac_version": (C.c_int, []),
    "fac_last_error": (C.c_char_p, []),
    "fac_launch_count": (C.c_longlong, []),
    "fac_reset_launch_count": (None, []),
    "fac_add_launch_count": (None, [C.c_longlong]),
    "fac_conv_gemm_f32": (C.c_int, [_P(ConvSrc), C.c_int, _fp, _fp, C.c_int, C.c_int, C.c_int, _P(ConvEpilogue),
                                    C.c_int, C.c_longlong, C.c_longlong, _fp]),
    "fac_waveglow_upsample_squeeze_f32": (C.c_int, [_P(WgModel), _fp, _fp, C.c_int, C.c_int, _fp]),
    "fac_wn_start_f32": (C.c_int, [_P(WgModel), C.c_int, _fp, _fp, C.c_int, C.c_int, _fp]),
    "fac_wn_layer_f32": (C.c_int, [_P(WgModel), C.c_int, C.c_int, _P(WgWorkspace), C.c_int, C.c_int, _fp]),
    "fac_wn_end_coupling_f32": (C.c_int, [_P(WgModel), C.c_int, _fp, _fp, C.c_int, C.c_int, _fp]),
    "fac_waveglow_infer_f32": (C.c_int, [_P(WgModel), _fp, _fp, _P(WgWorkspace), C.c_int, C.c_int, _fp]),
    "fac_waveglow_tc_prepare_spect": (C.c_int, [_P(WgModel), _P(WgTcWeights), _P(WgTcWorkspace), _fp, C.c_int, C.c_int,
                                                C.c_int, _fp]),
    "fac_wn_start_tc": (C.c_int, [_P(WgModel), C.c_int, _fp, _P(WgTcWorkspace), C.c_int, C.c_int, C.c_int, _fp]),
    "fac_wn_layer_tc": (C.c_int, [_P(WgModel), _P(WgTcWeights), C.c_int, C.c_int, _P(WgTcWorkspace), C.c_int,
                                  C.c_int, C.c_int, _fp]),
    "fac_waveglow_flow_step_tc": (C.c_int, [_P(WgModel), _P(WgTcWeights), C.c_int, _fp, _P(WgTcWorkspace), C.c_int,
                                            C.c_int, C.c_int, _fp]),
    "fac_wn_end_tc": (C.c_int, [_P(WgModel), _P(WgTcWeights), C.c_int, _fp, _fp, C.c_int, C.c_int, _fp]),
    "fac_waveglow_infer_tc": (C.c_int, [_P(WgModel), _P(WgTcWeights), _fp, _fp, _P(WgTcWorkspace), C.c_int, C.c_int,
                                        C.c_int, _fp]),
    "fac_conv_gemm_tc": (C.c_int, [_P(TcConv), _fp]),
    "fac_transpose_split_16": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    "fac_pad_split_16": (C.c_int, [_fp, _fp, _fp, C.c_longlong, C.c_int, C.c_int, C.c_int, _fp]),
    "fac_tc_set_profile_buffer": (None, [_fp]),
    "fac_taco_set_profile_buffer": (None, [_fp]),
    "fac_lstm_set_profile_buffer": (None, [_fp]),
    "fac_tc_set_cta_group": (C.c_int, [C.c_int]),
    "fac_tc_set_batch_group": (C.c_int, [C.c_int]),
    "fac_tc_set_k_block": (C.c_int, [C.c_int]),
    "fac_tc_set_fused": (C.c_int, [C.c_int]),
    "fac_selftest_grid_barrier": (C.c_int, [_fp, C.c_int, _fp]),
    "fac_denoise_spectrum_f32": (C.c_int, [_fp, _fp, C.c_float, C.c_longlong, C.c_int, C.c_int, _fp]),
    "fac_ppg_sparsify": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _fp]),
    "fac_prenet0_sparse_f32": (C.c_int, [_fp, _fp, _fp, C.c_int, _fp, _fp, _fp, C.c_int, _fp, _fp, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    "fac_lstm_bidir_f32": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, C.c_int, _fp]),
    "fac_lstm_bidir_var_f32": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, _fp]),
    "fac_taco_decoder_run": (C.c_int, [_P(TacoDecoderWeights), _fp, _fp, _fp, _fp, _P(TacoDecoderState), _fp, _fp,
                                       _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _fp]),
}

_lib = None


class FacError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True):
    """Load (building first if necessary) the shared library and bind every symbol."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if build_if_missing:
        path = _build.build()
    if not os.path.isfile(path):
        raise FacError("libfacb200.so not found at %s -- run `python -m fac_via_ppg_b200.build`" % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    cta_group, k_block = os.environ.get("FAC_TC_CTA_GROUP"), os.environ.get("FAC_TC_K_BLOCK")
    if os.environ.get("FAC_TC_BATCH_GROUP"):
        check(lib.fac_tc_set_batch_group(int(os.environ["FAC_TC_BATCH_GROUP"])), "fac_tc_set_batch_group")
    if cta_group:
        check(lib.fac_tc_set_cta_group(int(cta_group)), "fac_tc_set_cta_group")
    if k_block:
        check(lib.fac_tc_set_k_block(int(k_block)), "fac_tc_set_k_block")
    if os.environ.get("FAC_TC_FUSED"):
        check(lib.fac_tc_set_fused(int(os.environ["FAC_TC_FUSED"])), "fac_tc_set_fused")
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().fac_last_error()
        raise FacError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else "?"))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t, name: str):
    """Every entry point launches on the CURRENT device's current stream (one process per GPU): a CPU tensor or a
    tensor of another device is refused instead of launching with foreign pointers."""
    if not t.is_cuda:
        raise FacError("%s must live on a CUDA device: this package has no CPU path (got %s)" % (name, t.device))
    import torch
    if t.device.index != torch.cuda.current_device():
        raise FacError("%s lives on %s but the current device is cuda:%d; wrap the call in "
                       "torch.cuda.device(...) (one process per GPU)" % (name, t.device, torch.cuda.current_device()))
