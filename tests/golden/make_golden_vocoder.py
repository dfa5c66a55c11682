"""Generate tests/golden/vocoder.npz by running the UNMODIFIED reference `hifigan.Generator`
(/root/reference/hifigan/models.py) in float64 on the seeded synthetic weights.  Authoring container only:
    python tests/golden/make_golden_vocoder.py
The weights go in through the reference's own path: a weight_norm module (weight_g / weight_v) is given the plain
weights, then `remove_weight_norm()` is called as utils/model.py:66 does."""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, REPO)


def build_reference_generator(sd, dtype=torch.float64):
    if REF not in sys.path:
        sys.path.insert(1, REF)
    if "hifigan" in sys.modules and not hasattr(sys.modules["hifigan"], "AttrDict"):
        del sys.modules["hifigan"]            # a stub left by another test
    import hifigan
    with open(os.path.join(REF, "hifigan", "config.json")) as f:
        cfg = hifigan.AttrDict(json.load(f))
    gen = hifigan.Generator(cfg)
    gen.eval()
    gen.remove_weight_norm()
    missing = gen.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return gen.to(dtype)


def main():
    import fs2_b200
    sd = fs2_b200.synthetic.synthetic_vocoder_state_dict(seed=0)
    gen = build_reference_generator(sd)
    keys = [(k, tuple(v.shape)) for k, v in gen.state_dict().items()]
    assert sorted(keys) == sorted((k, tuple(s)) for k, s in fs2_b200.synthetic.vocoder_schema()), "schema differs from the reference's"
    g = torch.Generator().manual_seed(5)
    mel = torch.randn(2, 80, 21, generator=g, dtype=torch.float64) * 1.5 - 2.0
    with torch.no_grad():
        wav = gen(mel)
    out_dir = os.path.dirname(os.path.abspath(__file__))
    np.savez_compressed(os.path.join(out_dir, "vocoder.npz"), mel=mel.numpy(), wav=wav.numpy(),
                        n_keys=np.array(len(keys)), rms=np.array(float(wav.pow(2).mean().sqrt())))
    print("vocoder fixture:", tuple(wav.shape), "rms", float(wav.pow(2).mean().sqrt()), "absmax", float(wav.abs().max()))


if __name__ == "__main__":
    main()
