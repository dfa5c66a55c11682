"""Hyper-parameters of the path, as plain dicts shaped like the reference's YAMLs.

`default_model_config()` mirrors config/ESD-Chinese-Singing-MFA/model.yaml:1-31 and
`default_preprocess_config()` the keys of preprocess.yaml:1-36 that the model
constructor reads (model/fastspeech2.py:16-71, model/modules.py:20-78).  The
kernels are compiled for exactly these dimensions; `check_supported` raises on
anything else (there is no generic fallback by design).
"""
import copy

# PostNet dimensions are hard-coded defaults in the reference (transformer/Layers.py:72-78),
# and the encoder vocabulary is len(text.symbols_ipa.symbols)+1 = 139 (transformer/Models.py:40).
N_SRC_VOCAB = 139
POSTNET_DIM = 512
POSTNET_KERNEL = 5
POSTNET_LAYERS = 5

_MODEL = {
    "transformer": {
        "encoder_layer": 4,
        "encoder_head": 2,
        "encoder_hidden": 256,
        "decoder_layer": 6,
        "decoder_head": 2,
        "decoder_hidden": 256,
        "conv_filter_size": 1024,
        "conv_kernel_size": [9, 1],
        "encoder_dropout": 0.2,
        "decoder_dropout": 0.2,
    },
    "variance_predictor": {"filter_size": 256, "kernel_size": 3, "dropout": 0.5},
    "variance_embedding": {
        "pitch_quantization": "linear",
        "energy_quantization": "linear",
        "n_bins": 256,
    },
    "multi_speaker": True,
    "multi_emotion": True,
    "max_seq_len": 2000,
    "vocoder": {"model": "HiFi-GAN", "speaker": "universal"},
}

_PREPROCESS = {
    "dataset": "ESD-Chinese-Singing-MFA",
    "path": {"preprocessed_path": "./preprocessed_data/ESD-Chinese-Singing-MFA"},
    "preprocessing": {
        "mel": {"n_mel_channels": 80, "mel_fmin": 0, "mel_fmax": 8000},
        "pitch": {"feature": "phoneme_level", "normalization": True},
        "energy": {"feature": "phoneme_level", "normalization": True},
        "audio": {"sampling_rate": 22050, "max_wav_value": 32768.0},
        "stft": {"filter_length": 1024, "hop_length": 256, "win_length": 1024},
    },
}


def default_model_config():
    return copy.deepcopy(_MODEL)


def default_preprocess_config(preprocessed_path=None):
    cfg = copy.deepcopy(_PREPROCESS)
    if preprocessed_path is not None:
        cfg["path"]["preprocessed_path"] = str(preprocessed_path)
    return cfg


def check_supported(preprocess_config, model_config):
    """Raise ValueError unless the configs name exactly the compiled-for architecture."""
    t = model_config["transformer"]
    want = _MODEL["transformer"]
    for k in ("encoder_layer", "encoder_head", "encoder_hidden", "decoder_layer",
              "decoder_head", "decoder_hidden", "conv_filter_size"):
        if int(t[k]) != want[k]:
            raise ValueError(f"fs2_b200 kernels are specialised for transformer.{k}={want[k]}, got {t[k]}")
    if [int(v) for v in t["conv_kernel_size"]] != [9, 1]:
        raise ValueError("fs2_b200 kernels are specialised for conv_kernel_size [9, 1]")
    vp = model_config["variance_predictor"]
    if int(vp["filter_size"]) != 256 or int(vp["kernel_size"]) != 3:
        raise ValueError("fs2_b200 kernels are specialised for variance_predictor 256/k3")
    ve = model_config["variance_embedding"]
    if int(ve["n_bins"]) != 256:
        raise ValueError("fs2_b200 kernels are specialised for n_bins=256")
    for q in ("pitch_quantization", "energy_quantization"):
        if ve[q] not in ("linear", "log"):
            raise ValueError(f"{q} must be 'linear' or 'log'")
    if not model_config["multi_speaker"] or not model_config["multi_emotion"]:
        raise ValueError("fs2_b200 implements the multi_speaker + multi_emotion model only")
    pp = preprocess_config["preprocessing"]
    if int(pp["mel"]["n_mel_channels"]) != 80:
        raise ValueError("fs2_b200 kernels are specialised for 80 mel channels")
    for f in ("pitch", "energy"):
        if pp[f]["feature"] not in ("phoneme_level", "frame_level"):        # model/modules.py:34-35
            raise ValueError(f"{f}.feature must be 'phoneme_level' or 'frame_level'")
