"""fs2_b200 -- B200-native FastSpeech2 acoustic-model inference forward.

Drop-in for `FastSpeech2.forward` of Napoliee/Expressive-FastSpeech2-Mandarin
(model/fastspeech2.py:73-149): hand-written sm_100a CUDA kernels behind a C-ABI
shared library (include/fs2_b200.h), called from a thin Python/PyTorch facade.
There is no CPU path: everything that computes lives in csrc/ and fails loudly
when the library is not built.
"""
from . import config, synthetic  # noqa: F401

__all__ = ["config", "synthetic", "FastSpeech2B200", "HiFiGANGeneratorB200", "get_vocoder", "vocoder_infer", "load_library"]


def __getattr__(name):
    # The facade needs the compiled library; import lazily so that host-only
    # utilities (config, synthetic, partition) stay usable without it.
    if name in ("FastSpeech2B200", "get_model"):
        from . import model
        return getattr(model, name)
    if name in ("HiFiGANGeneratorB200", "get_vocoder", "vocoder_infer"):
        from . import vocoder
        return getattr(vocoder, name)
    if name == "load_library":
        from ._lib import load_library
        return load_library
    raise AttributeError(name)
