"""Drop-in for the reference ``src/waveglow/convert_model.py``: old-format WaveGlow checkpoints.

Early WaveGlow checkpoints keep the residual and skip 1x1 convolutions of every WN layer as two module
lists, ``res_layers`` (n_layers - 1 entries; the last layer has no residual output) and ``skip_layers``
(n_layers entries), instead of the merged ``res_skip_layers`` of reference src/waveglow/glow.py:132-152.
``update_model`` (reference convert_model.py:43-70) stacks them -- residual rows first, then skip rows,
the order glow.py:164-169 slices them in -- so that ``WaveGlow.infer`` and the weight packer see the
current format.  ``load_waveglow_model`` (common/utils.py) applies it to every checkpoint it loads.

    python -m fac_via_ppg_b200.waveglow.convert_model old.pt new.pt
"""
from __future__ import annotations

import io
import sys

import torch


def _check_model_old_version(model) -> bool:
    """reference convert_model.py:37-41."""
    return len(model.WN) > 0 and hasattr(model.WN[0], "res_layers")


def _plain(conv):
    """(weight, bias) of a 1x1 conv with or without weight norm."""
    if hasattr(conv, "weight_g") and hasattr(conv, "weight_v"):
        v, g = conv.weight_v, conv.weight_g
        w = v * (g / v.flatten(1).norm(dim=1).view(-1, 1, 1))
    else:
        w = conv.weight
    return w.detach(), conv.bias.detach()


def update_model(old_model):
    """Old-format module -> a copy in the current format (the input is returned untouched when it already is)."""
    if not _check_model_old_version(old_model):
        return old_model
    # a pickle round trip instead of copy.deepcopy: weight-normed convs carry a derived, non-leaf `weight`
    # tensor that current PyTorch refuses to deep-copy (it pickles fine -- it is how the checkpoint got here)
    buf = io.BytesIO()
    torch.save(old_model, buf)
    buf.seek(0)
    model = torch.load(buf, weights_only=False)
    for wn in model.WN:
        merged = torch.nn.ModuleList()
        for i in range(wn.n_layers):
            w_skip, b_skip = _plain(wn.skip_layers[i])
            if i < wn.n_layers - 1:
                w_res, b_res = _plain(wn.res_layers[i])
                w, b = torch.cat([w_res, w_skip]), torch.cat([b_res, b_skip])
            else:
                w, b = w_skip, b_skip
            conv = torch.nn.Conv1d(wn.n_channels, w.shape[0], 1).to(device=w.device, dtype=w.dtype)
            with torch.no_grad():
                conv.weight.copy_(w)
                conv.bias.copy_(b)
            merged.append(torch.nn.utils.weight_norm(conv, name="weight"))   # what remove_weightnorm expects
        wn.res_skip_layers = merged
        del wn.res_layers
        del wn.skip_layers
    return model


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 2:
        raise SystemExit("usage: python -m fac_via_ppg_b200.waveglow.convert_model OLD_CHECKPOINT NEW_CHECKPOINT")
    import fac_via_ppg_b200
    fac_via_ppg_b200.install_aliases()
    ckpt = torch.load(argv[0], map_location="cpu", weights_only=False)
    ckpt["model"] = update_model(ckpt["model"])
    torch.save(ckpt, argv[1])


if __name__ == "__main__":
    main()
