"""Drop-in for the reference CLI ``src/script/generate_synthesis.py``: same four required flags, same
constants (sigma 0.6, denoiser strength 0.005, 16 kHz float32 WAV) and outputs
(``output_dir/ac.wav`` + ``debug.log``).

PPG extraction (Kaldi nnet3 via pykaldi, reference src/ppg/) is out of scope, so
``--teacher_utterance_path`` accepts a precomputed PPG: a ``.npy`` file holding a (T, 5816) float
array (what the reference's ``get_ppg`` returns, src/common/data_utils.py:55-59).  A recording (``.wav``)
is accepted only when ``FAC_REFERENCE_SRC`` points at a reference ``src/`` tree whose Kaldi front-end
imports: the original ``get_ppg`` then runs unchanged, with the reference's own ``common`` / ``ppg``
packages imported in isolation from this package's drop-in aliases.  Two extras help on boxes without
checkpoints: ``--synthetic SECONDS`` ignores the model/utterance paths and runs seeded random-init
models on a synthetic PPG, and ``--no_denoiser`` skips the post-filter.  ``--ppg_topk K`` prunes every PPG
frame to its K largest posteriors on the host (k * 8 bytes per frame cross PCIe instead of 23 KB).
"""
import argparse
import logging
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from scipy.io import wavfile  # noqa: E402

import fac_via_ppg_b200  # noqa: E402
from fac_via_ppg_b200 import synth  # noqa: E402
from fac_via_ppg_b200.common.hparams import create_hparams_stage  # noqa: E402
from fac_via_ppg_b200.common.layers import TacotronSTFT  # noqa: E402
from fac_via_ppg_b200.common.utils import get_inference, load_waveglow_model, waveglow_audio  # noqa: E402
from fac_via_ppg_b200.script.train_ppg2mel import load_model  # noqa: E402
from fac_via_ppg_b200.waveglow.denoiser import Denoiser  # noqa: E402
from fac_via_ppg_b200.waveglow.glow import WaveGlow  # noqa: E402


def load_teacher_ppg(path):
    if path.endswith(".npy"):
        ppg = np.load(path)
        if ppg.ndim != 2:
            raise ValueError("expected a (T, n_symbols) PPG array in %s" % path)
        return ppg.astype(np.float32)
    return _reference_get_ppg(path)


def _reference_get_ppg(path):
    """The reference's own front-end (src/common/data_utils.py:55-59 get_ppg + src/ppg DependenciesPPG) from the
    tree FAC_REFERENCE_SRC names.  install_aliases() binds ``common`` to this package's drop-ins, which have no
    ``data_utils``: the reference packages are therefore imported with those names temporarily un-aliased and
    are removed again afterwards, so both worlds keep resolving to their own modules."""
    src = os.environ.get("FAC_REFERENCE_SRC")
    if not src or not os.path.isdir(os.path.join(src, "common")):
        raise SystemExit("PPG extraction from a recording needs the reference's Kaldi front-end: set "
                         "FAC_REFERENCE_SRC to the reference's src/ directory (with pykaldi installed), or pass a "
                         "precomputed (T, 5816) .npy PPG as --teacher_utterance_path")
    mine = lambda name: name.split(".")[0] in ("common", "ppg")                       # noqa: E731
    saved = {name: sys.modules.pop(name) for name in list(sys.modules) if mine(name)}
    sys.path.insert(0, src)
    try:
        import ppg as ppg_module
        from common.data_utils import get_ppg
        return np.asarray(get_ppg(path, ppg_module.DependenciesPPG()), dtype=np.float32)
    except ImportError as exc:
        raise SystemExit("the reference front-end under %s does not import (%s); pass a .npy PPG instead" % (src, exc))
    finally:
        sys.path.remove(src)
        for name in [n for n in sys.modules if mine(n)]:
            del sys.modules[name]
        sys.modules.update(saved)


def main(argv=None):
    parser = argparse.ArgumentParser(description="Generate accent conversion speech using pre-trained models.")
    parser.add_argument("--ppg2mel_model", type=str, required=True, help="Path to the PPG-to-Mel model.")
    parser.add_argument("--waveglow_model", type=str, required=True, help="Path to the WaveGlow model.")
    parser.add_argument("--teacher_utterance_path", type=str, required=True,
                        help="Path to a native speaker recording (or a .npy PPG).")
    parser.add_argument("--output_dir", type=str, required=True, help="Output dir, will save the audio and log info.")
    parser.add_argument("--synthetic", type=float, default=0.0, metavar="SECONDS",
                        help="run seeded random-init models on a synthetic PPG of this length")
    parser.add_argument("--no_denoiser", action="store_true")
    parser.add_argument("--ppg_topk", type=int, default=0, metavar="K",
                        help="prune every PPG frame to its K (<= 64) largest posteriors on the host and run the "
                             "gather prenet (0 = dense, the reference's input)")
    args = parser.parse_args(argv)

    os.makedirs(args.output_dir, exist_ok=True)
    logging.basicConfig(filename=os.path.join(args.output_dir, "debug.log"), level=logging.DEBUG, force=True)
    logging.info("Output dir: %s", args.output_dir)
    fs, waveglow_sigma, denoiser_strength, denoiser_mode, is_clip = 16000, 0.6, 0.005, "zeros", False
    for key, val in (("Tacotron", args.ppg2mel_model), ("Waveglow", args.waveglow_model), ("is_clip", is_clip),
                     ("Fs", fs), ("Sigma", waveglow_sigma), ("Denoiser strength", denoiser_strength),
                     ("Denoiser mode", denoiser_mode)):
        logging.debug("%s: %s", key, val)

    fac_via_ppg_b200.install_aliases()
    hparams = create_hparams_stage()
    TacotronSTFT(hparams.filter_length, hparams.hop_length, hparams.win_length, hparams.n_acoustic_feat_dims,
                 hparams.sampling_rate, hparams.mel_fmin, hparams.mel_fmax)      # constructed, unused (as upstream)
    tacotron_model = load_model(hparams)
    if args.synthetic > 0:
        tacotron_model.load_state_dict(synth.tacotron_state())
        waveglow_model = WaveGlow.remove_weightnorm(WaveGlow(**synth.WAVEGLOW_CONFIG))
        waveglow_model.load_state_dict(synth.waveglow_state())
        waveglow_model.cuda().eval()
        n_frames = int(round(args.synthetic * fs / hparams.hop_length))
        teacher_ppg = synth.synthetic_ppg(1, n_frames)[0].t().numpy()
        tacotron_model.decoder.gate_threshold = 2.0          # random weights never learn to stop
        tacotron_model.decoder.max_decoder_steps = n_frames
    else:
        tacotron_model.load_state_dict(torch.load(args.ppg2mel_model, weights_only=False)["state_dict"])
        waveglow_model = load_waveglow_model(args.waveglow_model)
        if not os.path.isfile(args.teacher_utterance_path):
            logging.warning("Missing %s", args.teacher_utterance_path)
            return 1
        teacher_ppg = load_teacher_ppg(args.teacher_utterance_path)
    tacotron_model.eval()
    denoiser = None if args.no_denoiser else Denoiser(waveglow_model, mode=denoiser_mode)

    logging.info("Perform AC on %s", args.teacher_utterance_path)
    ac_mel = get_inference(teacher_ppg, tacotron_model, is_clip, ppg_topk=args.ppg_topk)
    ac_wav = waveglow_audio(ac_mel, waveglow_model, waveglow_sigma, True)
    if denoiser is not None:
        ac_wav = denoiser(ac_wav, strength=denoiser_strength)[:, 0]
    ac_wav = ac_wav.float().cpu().numpy().T
    wavfile.write(os.path.join(args.output_dir, "ac.wav"), fs, ac_wav)
    logging.info("Done!")
    return 0


if __name__ == "__main__":
    sys.exit(main())
