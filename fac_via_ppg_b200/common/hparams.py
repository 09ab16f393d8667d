"""Drop-in for the reference ``src/common/hparams.py`` (read-only configuration surface).

``create_hparams_stage()`` returns the Interspeech'19 values the CLI uses (reference
hparams.py:161-241, generate_synthesis.py:74); ``create_hparams()`` the training defaults
(hparams.py:40-158).  Unknown keyword overrides are rejected like in the reference.  The
tables are grouped by what consumes them; only MODEL_DIMS reaches the CUDA path.
"""
from __future__ import annotations

from types import SimpleNamespace

# geometry of the PPG->Mel network: the only keys the inference path reads
MODEL_DIMS = dict(
    n_symbols=5816, symbols_embedding_dim=600, n_acoustic_feat_dims=80,
    encoder_kernel_size=5, encoder_n_convolutions=3, encoder_embedding_dim=600,
    prenet_dim=300, attention_rnn_dim=300, decoder_rnn_dim=300,
    attention_dim=150, attention_location_n_filters=32, attention_location_kernel_size=31,
    attention_window_size=20,
    postnet_embedding_dim=512, postnet_kernel_size=5, postnet_n_convolutions=5,
    max_decoder_steps=1000, gate_threshold=0.5, p_attention_dropout=0.1, p_decoder_dropout=0.1,
    fp16_run=False, mask_padding=True,
)
AUDIO = dict(max_wav_value=32768.0, sampling_rate=16000, filter_length=1024, hop_length=160, win_length=1024,
             mel_fmin=0.0, mel_fmax=8000.0)
RUNTIME = dict(epochs=1000, seed=16807, dynamic_loss_scaling=True, distributed_run=False, dist_backend="nccl",
               dist_url="tcp://localhost:54321", cudnn_enabled=True, cudnn_benchmark=False, log_directory="log",
               warm_start=False, n_gpus=1, rank=0, group_name="group_name", training_files="", validation_files="",
               is_full_ppg=True, is_append_f0=False, ppg_subsampling_factor=1, is_cache_feats=False,
               feats_cache_path="", use_saved_learning_rate=False, weight_decay=1e-6, grad_clip_thresh=1.0,
               batch_size=6, mel_weight=1, gate_weight=0.005)
# values that differ between the two factory functions
DEFAULT_ONLY = dict(iters_per_checkpoint=200, output_directory=None, checkpoint_path="", load_feats_from_disk=False,
                    learning_rate=1e-5)
STAGE_ONLY = dict(iters_per_checkpoint=100, output_directory="", checkpoint_path=None, load_feats_from_disk=True,
                  learning_rate=1e-4, is_large_set=False, is_skip_sil=False, mvn_stats_file="",
                  sequence_level="sentence")


class HParamsView(SimpleNamespace):
    """Attribute view over the hyper-parameter table (reference hparams.py:35-37)."""


def _make(tables, overrides):
    merged = {}
    for t in tables:
        merged.update(t)
    unknown = [k for k in overrides if k not in merged]
    if unknown:
        raise ValueError("The hyper-parameter %s is not supported." % unknown[0])
    merged.update(overrides)
    return HParamsView(**merged)


def create_hparams(**kwargs):
    return _make((RUNTIME, AUDIO, MODEL_DIMS, DEFAULT_ONLY), kwargs)


def create_hparams_stage(**kwargs):
    return _make((RUNTIME, AUDIO, MODEL_DIMS, STAGE_ONLY), kwargs)
