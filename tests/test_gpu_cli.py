"""The CLI drop-in end to end on the GPU (random-init weights, mel .npy output)."""
import json
import os
import tempfile

import numpy as np
import pytest
import yaml

import fs2_b200
from fs2_b200 import cli

pytestmark = pytest.mark.gpu


def test_cli_single_sentence(sd32):
    d = tempfile.mkdtemp(prefix="fs2_cli_")
    fs2_b200.synthetic.write_fixture_jsons(d)
    paths = {}
    cfgs = {"p": fs2_b200.config.default_preprocess_config(d), "m": fs2_b200.config.default_model_config(),
            "t": {"path": {"ckpt_path": d, "result_path": os.path.join(d, "result")}}}
    for k, v in cfgs.items():
        paths[k] = os.path.join(d, k + ".yaml")
        with open(paths[k], "w") as f:
            yaml.safe_dump(v, f)
    cli.main(["--restore_step", "1", "--mode", "single", "--text", "今天天气真好", "--speaker_id", "0001", "--emotion",
              "Happy", "-p", paths["p"], "-m", paths["m"], "-t", paths["t"], "--random_init", "--duration_control", "1.2"])
    mel = np.load(os.path.join(d, "result", "synthesis_0001_Happy.npy"))
    meta = json.load(open(os.path.join(d, "result", "synthesis_0001_Happy.json")))
    assert mel.ndim == 2 and mel.shape[1] == 80 and mel.shape[0] == meta["n_frames"] > 16
    assert meta["n_phonemes"] == 16 and np.isfinite(mel).all()
