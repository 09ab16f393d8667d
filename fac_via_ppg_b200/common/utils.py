"""Drop-in for the inference glue of the reference ``src/common/utils.py``.

Same function names / arguments as the reference (utils.py:39-78, 107-112, 142-181).  The two
mask helpers are kept for API compatibility only: the decoder kernel evaluates the attention
window in place (csrc/tacotron_decoder.cu), so nothing on the hot path builds a mask or syncs
with the host per step.  Data-preparation helpers (wav loading, notch filter, get_mel) are out
of scope.
"""
from __future__ import annotations

import numpy as np
import torch


def get_mask_from_lengths(lengths):
    """(B,) lengths -> (B, max_len) bool, True inside the sequence (reference utils.py:39-43)."""
    ids = torch.arange(int(lengths.max()), device=lengths.device)
    return ids[None, :] < lengths[:, None]


def get_mask_from_lengths_window_and_time_step(lengths, attention_window_size, time_step):
    """True = masked (reference utils.py:46-78), vectorised and free of host syncs.  Keeps the
    documented quirk: past an utterance's end its last frame stays unmasked."""
    max_idx = lengths.to(torch.long) - 1
    start = torch.clamp(torch.full_like(max_idx, time_step - attention_window_size), min=0)
    start = torch.minimum(start, max_idx)
    end = torch.minimum(torch.full_like(max_idx, time_step + attention_window_size), max_idx)
    ids = torch.arange(int(lengths.max()), device=lengths.device)[None, :]
    return ~((ids >= start[:, None]) & (ids <= end[:, None]))


def to_gpu(x):
    """reference utils.py:107-112."""
    x = x.contiguous()
    if torch.cuda.is_available():
        x = x.cuda(non_blocking=True)
    return x


def waveglow_audio(mel, waveglow, sigma, is_cuda_output=False):
    """reference utils.py:142-152: mel (B, 80, F) -> int16 numpy of utterance 0, or CUDA float (B, T)."""
    mel = mel.cuda()
    with torch.no_grad():
        audio = waveglow.infer(mel, sigma=sigma)
    if is_cuda_output:
        return audio
    return (32768 * audio[0]).cpu().numpy().astype("int16")


def get_inference(seq, model, is_clip=False, ppg_topk=0, ppg_threshold=0.0):
    """reference utils.py:155-174: (T, D) numpy PPG -> mel_outputs_postnet (1, 80, T_out) on the GPU.
    ``ppg_topk`` > 0 (an extension): prune every frame on the host to its ppg_topk largest posteriors (> threshold)
    and ship the (index, value) lists instead of the dense matrix (ops.SparsePPG)."""
    seq = torch.from_numpy(np.asarray(seq)).float().transpose(0, 1).unsqueeze(0)
    if ppg_topk > 0:
        from fac_via_ppg_b200.ops import SparsePPG
        lists = SparsePPG.from_dense_host(seq, k=ppg_topk, threshold=ppg_threshold).to("cuda", non_blocking=True)
        _, mel_outputs_postnet, _, _ = model.inference(lists)
        if is_clip:
            return mel_outputs_postnet[:, :, 10:(seq.size(2) - 10)]
        return mel_outputs_postnet
    seq = to_gpu(seq)
    _, mel_outputs_postnet, _, _ = model.inference(seq)
    if is_clip:
        return mel_outputs_postnet[:, :, 10:(seq.size(2) - 10)]
    return mel_outputs_postnet


def load_waveglow_model(path):
    """reference utils.py:177-181: a pickled ``{'model': WaveGlow}`` checkpoint -> eval model on the GPU.
    The pickle refers to ``waveglow.glow.WaveGlow``; ``fac_via_ppg_b200.install_aliases()`` (called here)
    makes that name resolve to the drop-in class.  Checkpoints in the old res_layers / skip_layers format are
    converted on the fly; with FAC_PACK_CACHE set the packed weight buffer is cached on disk (packing.py)."""
    import fac_via_ppg_b200
    from fac_via_ppg_b200.waveglow.convert_model import update_model
    fac_via_ppg_b200.install_aliases()
    model = torch.load(path, weights_only=False)["model"]
    model = update_model(model)              # old res_layers/skip_layers format (reference convert_model.py:43-70)
    model = model.remove_weightnorm(model)
    model.cuda().eval()
    return model
