"""Drop-in for the parameter wrappers of the reference ``src/common/layers.py:40-71``.

``LinearNorm`` / ``ConvNorm`` only exist so that state-dict keys keep their
``.linear_layer.`` / ``.conv.`` infixes; the arithmetic runs in fac_conv_gemm_f32.
``TacotronSTFT`` (training-target extraction, constructed but unused by the CLI) is out of scope.
"""
from __future__ import annotations

import torch


def _gain(name):
    return torch.nn.init.calculate_gain(name)


class LinearNorm(torch.nn.Module):
    def __init__(self, in_dim, out_dim, bias=True, w_init_gain="linear"):
        super().__init__()
        self.linear_layer = torch.nn.Linear(in_dim, out_dim, bias=bias)
        torch.nn.init.xavier_uniform_(self.linear_layer.weight, gain=_gain(w_init_gain))

    def forward(self, x):
        raise NotImplementedError("fac_via_ppg_b200: LinearNorm is a parameter holder; see Tacotron2.inference")


class ConvNorm(torch.nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=None, dilation=1, bias=True,
                 w_init_gain="linear"):
        super().__init__()
        if padding is None:
            if kernel_size % 2 != 1:
                raise ValueError("ConvNorm needs an odd kernel when padding is implicit")
            padding = dilation * (kernel_size - 1) // 2
        self.conv = torch.nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                                    padding=padding, dilation=dilation, bias=bias)
        torch.nn.init.xavier_uniform_(self.conv.weight, gain=_gain(w_init_gain))

    def forward(self, signal):
        raise NotImplementedError("fac_via_ppg_b200: ConvNorm is a parameter holder; see Tacotron2.inference")


class TacotronSTFT(torch.nn.Module):
    """Placeholder: the CLI constructs it (generate_synthesis.py:75-78) but never calls it."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.args, self.kwargs = args, kwargs

    def mel_spectrogram(self, y):
        raise NotImplementedError("mel extraction for training targets is out of scope (SURVEY.md section 2, row 4)")
