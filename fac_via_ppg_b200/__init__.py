"""fac_via_ppg_b200 -- B200-native (sm_100a) inference path for guanlongzhao/fac-via-ppg:
PPG -> Mel (Tacotron2 variant) -> WaveGlow, behind the reference's own Python call signatures.

Layout: ``csrc/`` hand-written CUDA behind the C ABI of ``include/fac_b200.h``; ``_ext.py`` the
ctypes binding; ``waveglow/glow.py`` and ``common/{model,layers,hparams,utils}.py`` the drop-ins
for the reference files of the same names; ``script/generate_synthesis.py`` the CLI.
"""
import importlib
import sys

_ALIASES = {
    "waveglow": "fac_via_ppg_b200.waveglow",
    "waveglow.glow": "fac_via_ppg_b200.waveglow.glow",
    "waveglow.denoiser": "fac_via_ppg_b200.waveglow.denoiser",
    "waveglow.convert_model": "fac_via_ppg_b200.waveglow.convert_model",
    "common": "fac_via_ppg_b200.common",
    "common.model": "fac_via_ppg_b200.common.model",
    "common.layers": "fac_via_ppg_b200.common.layers",
    "common.hparams": "fac_via_ppg_b200.common.hparams",
    "common.utils": "fac_via_ppg_b200.common.utils",
}


def install_aliases(force: bool = False):
    """Register the drop-in modules under the reference's import names (``waveglow.glow``,
    ``common.model`` ...) so that pickled reference checkpoints and unmodified caller code
    (``from common.utils import get_inference``) resolve to this package."""
    for alias, target in _ALIASES.items():
        if force or alias not in sys.modules:
            sys.modules[alias] = importlib.import_module(target)
