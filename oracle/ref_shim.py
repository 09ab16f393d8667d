"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference modules.

Only usable where /root/reference exists (the build container); it is what
oracle/make_golden.py uses to produce tests/golden/*.pt and what the CPU tests
use to pin oracle/*.py against the real thing.  Nothing in the product path,
the GPU tests, smoke() or bench.py may import this file: the reference tree is
absent on the GPU box.

Three shims, none of which touches /root/reference:
 1. ``common`` is registered as a bare namespace so ``common/__init__.py`` (which
    star-imports pykaldi/textgrid/protobuf code) never runs; ``librosa`` is
    stubbed with the two helpers stft.py needs.
 2. The legacy CUDA-only tensor constructors (``torch.cuda.FloatTensor`` ...)
    are pointed at CPU factories when no GPU is present; ByteTensor -> bool
    because ``masked_fill_`` rejects uint8 masks on torch >= 2.
 3. Determinism is obtained by seeding torch's generator before each call.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("FAC_REFERENCE_ROOT", "/root/reference")
REFERENCE_SRC = os.path.join(REFERENCE_ROOT, "src")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "waveglow", "glow.py"))


def _pad_center(data, size, axis=-1):
    n = data.shape[axis]
    lpad = int((size - n) // 2)
    lengths = [(0, 0)] * data.ndim
    lengths[axis] = (lpad, int(size - n - lpad))
    return np.pad(data, lengths, mode="constant")


def _tiny(x):
    x = np.asarray(x)
    dtype = x.dtype if np.issubdtype(x.dtype, np.floating) else np.float32
    return np.finfo(dtype).tiny


def _normalize(S, norm=np.inf, axis=0):
    if norm is None:
        return S
    mag = np.abs(S).astype(float)
    length = np.max(mag, axis=axis, keepdims=True)
    length[length < _tiny(S)] = 1.0
    return S / length


_installed = False


def install():
    """Make ``waveglow.glow`` / ``common.model`` ... importable from the reference."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    # (1) bare 'common' package + librosa stub
    # (drop-in aliases of fac_via_ppg_b200.install_aliases() under the same names would shadow the reference)
    for name in [n for n in sys.modules if n.split(".")[0] in ("common", "waveglow")
                 and "fac_via_ppg_b200" in (getattr(sys.modules[n], "__name__", "") or "")]:
        del sys.modules[name]
    common = types.ModuleType("common")
    common.__path__ = [os.path.join(REFERENCE_SRC, "common")]
    sys.modules.setdefault("common", common)
    if "librosa" not in sys.modules:
        librosa = types.ModuleType("librosa")
        util = types.ModuleType("librosa.util")
        util.pad_center, util.tiny, util.normalize = _pad_center, _tiny, _normalize
        filters = types.ModuleType("librosa.filters")
        filters.mel = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("librosa.filters.mel stub"))
        librosa.util, librosa.filters = util, filters
        sys.modules["librosa"] = librosa
        sys.modules["librosa.util"] = util
        sys.modules["librosa.filters"] = filters
    if REFERENCE_SRC not in sys.path:
        sys.path.append(REFERENCE_SRC)
    # (2) legacy constructors on a GPU-less host
    if not torch.cuda.is_available():
        class _Bool:
            def __new__(cls, *shape):
                return torch.empty(*shape, dtype=torch.bool)
        torch.cuda.FloatTensor = torch.FloatTensor
        torch.cuda.HalfTensor = torch.HalfTensor
        torch.cuda.LongTensor = torch.LongTensor
        torch.cuda.ByteTensor = _Bool
        # Module.cuda()/Tensor.cuda() become no-ops so utils.py / denoiser.py run on CPU
        torch.nn.Module.cuda = lambda self, device=None: self
        torch.Tensor.cuda = lambda self, *a, **k: self
    else:
        class _BoolCuda:
            def __new__(cls, *shape):
                return torch.empty(*shape, dtype=torch.bool, device="cuda")
        torch.cuda.ByteTensor = _BoolCuda
    _installed = True


def reference_waveglow(state, cfg):
    """A reference ``WaveGlow`` (src/waveglow/glow.py:178) carrying ``state``."""
    install()
    from waveglow.glow import WaveGlow  # type: ignore
    model = WaveGlow(**cfg)
    model = WaveGlow.remove_weightnorm(model)
    model.load_state_dict(state, strict=True)
    return model.eval()


def reference_tacotron(state, **hparam_overrides):
    """A reference ``Tacotron2`` (src/common/model.py:538) carrying ``state``."""
    install()
    from common.hparams import create_hparams_stage  # type: ignore
    from common.model import Tacotron2  # type: ignore
    hp = create_hparams_stage(**hparam_overrides)
    model = Tacotron2(hp)
    model.load_state_dict(state, strict=True)
    return model.eval()
