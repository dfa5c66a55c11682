"""The CLI's text front end: symbol ids, syllable splitting and batch tuple (no GPU needed)."""
import os
import sys
import tempfile
import types

import numpy as np
import pytest

from fs2_b200 import cli, synthetic


def test_symbol_ids_match_the_reference_table():
    # text/symbols_pinyin.py: pad, '-', punctuation, letters, 44 pinyin phonemes; later duplicates win
    assert len(cli.SYMBOLS) == 108 and len(cli.SYMBOL_TO_ID) == 83
    assert cli.SYMBOL_TO_ID["_"] == 0 and cli.SYMBOL_TO_ID["a"] == 64 and cli.SYMBOL_TO_ID["n"] == 86
    assert cli.SYMBOL_TO_ID["zh"] == 107 and cli.SYMBOL_TO_ID["A"] == 12


def test_config1_sentence_gives_the_fixture_ids():
    ph = cli.text_to_phonemes("今天天气真好")
    assert ph == "j i n t ia n t ia n q i zh e n h ao".split()
    assert cli.phonemes_to_ids(ph).tolist() == synthetic.C1_IDS
    assert cli.text_to_phonemes("{ni hao shi jie}") == ["ni", "hao", "shi", "jie"]
    assert cli.phonemes_to_ids(["ni", "h", "??"]).tolist() == [0, 76, 0]   # unknown -> padding id


@pytest.mark.parametrize("syllable", ["zhuang", "xiong", "lv", "nve", "er", "a", "yuan", "shi", "qiu", "wen", "o", "juan"])
def test_syllable_split_matches_reference_rules(syllable):
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference tree only exists in the authoring container")
    # run the reference's own splitter with pypinyin stubbed to return the syllable
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden
    make_golden.import_reference()
    ref_mod = sys.modules.get("synthesize_chinese_pinyin")
    stub = getattr(ref_mod, "pypinyin", None) if ref_mod is not None else None
    if stub is None or not hasattr(stub, "_fs2_stub"):
        stub = types.ModuleType("pypinyin")
        stub._fs2_stub = True
        stub.Style = types.SimpleNamespace(NORMAL=0)
        stub.lazy_pinyin = lambda text, style=None: [stub.current]
    sys.modules["pypinyin"] = stub
    stub.current = syllable
    for name in ("dataset_chinese", "dataset"):      # hifigan itself imports fine (torch only)
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["dataset_chinese"].TextDataset = object
    import importlib
    try:
        ref = importlib.import_module("synthesize_chinese_pinyin")
    except Exception as e:  # the script imports vocoder / dataset helpers that need missing packages
        pytest.skip(f"reference CLI not importable here: {e}")
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        want = ref.chinese_to_pinyin_phonemes("x")
    sys.modules.pop("pypinyin", None)       # do not leak the stub into the other tests
    ref.pypinyin = stub                      # the reference module keeps its own handle
    assert cli.syllable_to_phonemes(syllable) == want


def test_single_batch_tuple():
    d = synthetic.write_fixture_jsons(tempfile.mkdtemp(prefix="fs2_json_"))
    args = cli.build_parser().parse_args(["--restore_step", "1", "--mode", "single", "--text", "今天天气真好", "--speaker_id",
                                          "0001", "--emotion", "Happy", "-p", "p", "-m", "m", "-t", "t"])
    b = cli.single_batch(args, {"path": {"preprocessed_path": d}})
    assert b[0] == ["synthesis_0001_Happy"] and b[8] == 16
    assert b[2].tolist() == [0] and b[3].tolist() == [1] and b[4].tolist() == [1] and b[5].tolist() == [1]
    assert b[6].tolist() == [synthetic.C1_IDS]


def test_batch_source_file():
    d = synthetic.write_fixture_jsons(tempfile.mkdtemp(prefix="fs2_json_"))
    src = os.path.join(d, "val.txt")
    with open(src, "w", encoding="utf-8") as f:
        f.write("0001_000001|0001|{j i n t ia n}|今天|x|Happy|0.8|0.8\n")
        f.write("0002_000002|0002|{h ao}|好|x|Sad|0.3|0.2\n")
    (ids, raw, spk, emo, aro, val, texts, lens, mx), = list(cli.source_batches(src, {"path": {"preprocessed_path": d}}))
    assert ids == ["0001_000001", "0002_000002"] and mx == 6 and lens.tolist() == [6, 2]
    assert texts[1].tolist() == [76, 66, 0, 0, 0, 0] and spk.tolist() == [0, 1] and emo.tolist() == [1, 3]
    assert aro.tolist() == [1, 3] and val.tolist() == [1, 3]


def test_batch_source_drops_utterances_longer_than_max_seq_len():
    """TextDataset.process_meta (dataset_chinese.py:244-255): a line whose recorded mel has more than max_seq_len rows is
    skipped; a listed utterance without a mel file raises, as np.load does there."""
    d = synthetic.write_fixture_jsons(tempfile.mkdtemp(prefix="fs2_json_"))
    os.makedirs(os.path.join(d, "mel"))
    np.save(os.path.join(d, "mel", "0001-mel-short.npy"), np.zeros((30, 80), dtype=np.float32))
    np.save(os.path.join(d, "mel", "0002-mel-long.npy"), np.zeros((41, 80), dtype=np.float32))
    src = os.path.join(d, "val.txt")
    with open(src, "w", encoding="utf-8") as f:
        f.write("short|0001|{j i n}|今|x|Happy|0.8|0.8\n")
        f.write("long|0002|{h ao}|好|x|Sad|0.3|0.2\n")
    cfg = {"path": {"preprocessed_path": d}}
    (ids, *_), = list(cli.source_batches(src, cfg, max_seq_len=40))
    assert ids == ["short"]
    (ids, *_), = list(cli.source_batches(src, cfg, max_seq_len=41))
    assert ids == ["short", "long"]
    (ids, *_), = list(cli.source_batches(src, cfg, max_seq_len=40, mel_filter=False))
    assert ids == ["short", "long"]
    with open(src, "a", encoding="utf-8") as f:
        f.write("absent|0001|{h ao}|好|x|Sad|0.3|0.2\n")
    with pytest.raises(FileNotFoundError):
        list(cli.source_batches(src, cfg, max_seq_len=40))


def test_expand_matches_the_reference_rule():
    """utils/tools.py:163-167: value j repeated max(0, int(d_j)) times."""
    vals, durs = np.array([1.5, -2.0, 3.0, 4.0]), np.array([2.0, 0.0, 3.9, -1.0])
    want = []
    for v, dd in zip(vals, durs):
        want += [v] * max(0, int(dd))
    assert cli.expand(vals, durs).tolist() == want
