"""Drop-in for the one symbol of the reference ``src/script/train_ppg2mel.py`` that the inference
CLI imports: ``load_model`` (reference train_ppg2mel.py:113-119).  Training itself is out of scope."""
from fac_via_ppg_b200.common.model import Tacotron2


def load_model(hparams):
    model = Tacotron2(hparams).cuda()
    if hparams.fp16_run:
        raise NotImplementedError("fp16_run is not supported by the B200 path (the reference README marks it broken)")
    return model
