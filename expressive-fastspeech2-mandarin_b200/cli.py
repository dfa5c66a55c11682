"""Command-line front end with the flags of the reference's synthesize_chinese_pinyin.py
(:156-308): single-sentence or batch synthesis of mel spectrograms on the B200 engine.

    python -m fs2_b200.cli --restore_step 900000 --mode single --text "今天天气真好" \\
        --speaker_id 0001 --emotion Happy -p preprocess.yaml -m model.yaml -t train.yaml

What differs from the reference script, by design: the acoustic model is `FastSpeech2B200` and the vocoder is
`HiFiGANGeneratorB200` (both on libfs2b200.so).  Per utterance the CLI writes `<result_path>/<id>.wav` (int16,
`synth_samples`, utils/tools.py:258-271) when vocoder weights are available (`hifigan/generator_<speaker>.pth.tar` as in
utils/model.py:60-63, or `--vocoder_ckpt`), and always the postnet mel `<id>.npy` plus `<id>.json` (text, phoneme and
frame counts).  The spectrogram PNG of the reference needs matplotlib and is not produced.  `--random_init` replaces
both checkpoints by the seeded synthetic weights so that the CLI runs end to end without trained weights;
`--no_vocoder` stops after the mel.
"""
import argparse
import json
import os

import numpy as np
import torch
import yaml

# ---- symbol table of text/symbols_pinyin.py:3-26 (ids are positions; later duplicates win) ----
_PAD, _SPECIAL, _PUNCT = "_", "-", "!'(),.:;? "
_LETTERS = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz"
PHONEMES = ("a ai ao b c ch d e ei er f g h i ia iao ie iu j k l m n ng o ou p q r s sh spn t u ua uai ue ui uo "
            "w x y z zh").split()
SYMBOLS = [_PAD] + list(_SPECIAL) + list(_PUNCT) + list(_LETTERS) + PHONEMES
SYMBOL_TO_ID = {s: i for i, s in enumerate(SYMBOLS)}

# emotion -> (arousal, valence) keys (synthesize_chinese_pinyin.py:281-287)
EMOTION_AV = {"Angry": ("0.9", "0.1"), "Happy": ("0.8", "0.8"), "Neutral": ("0.5", "0.5"), "Sad": ("0.3", "0.2"),
              "Surprise": ("0.8", "0.6")}

# syllable -> phonemes (synthesize_chinese_pinyin.py:35-96): longest initial first, finals table,
# unknown finals fall back to per-character lookup
_INITIALS = ["zh", "ch", "sh", "b", "p", "m", "f", "d", "t", "n", "l", "g", "k", "h", "j", "q", "x", "r", "z", "c", "s",
             "y", "w"]
_FINALS = {"a": "a", "o": "o", "e": "e", "i": "i", "u": "u", "v": "y", "ai": "ai", "ei": "ei", "ui": "ui", "ao": "ao",
           "ou": "ou", "iu": "iu", "ie": "ie", "ue": "ue", "ve": "ue", "an": "a n", "en": "e n", "in": "i n", "un": "u n",
           "vn": "y n", "ang": "a ng", "eng": "e ng", "ing": "i ng", "ong": "o ng", "er": "er", "iao": "iao",
           "ian": "ia n", "iang": "ia ng", "iong": "io ng", "uai": "uai", "uan": "ua n", "uang": "ua ng"}
# enough hanzi for the repository's own demo sentences when pypinyin is not installed
_DEMO_PINYIN = {"今": "jin", "天": "tian", "气": "qi", "真": "zhen", "好": "hao", "你": "ni", "世": "shi", "界": "jie"}


def syllable_to_phonemes(syllable):
    initial = next((i for i in _INITIALS if syllable.startswith(i)), "")
    final = syllable[len(initial):]
    out = [initial] if initial else []
    if final in _FINALS:
        out += _FINALS[final].split()
    else:
        for ch in final:
            out += _FINALS[ch].split() if ch in _FINALS else [ch]
    return out


def text_to_phonemes(text):
    """'{a b c}' is a literal phoneme list (:110-112); anything else is Chinese text (:24-104)."""
    if text.startswith("{") and text.endswith("}"):
        return text[1:-1].split()
    try:
        from pypinyin import Style, lazy_pinyin
        syllables = lazy_pinyin(text, style=Style.NORMAL)
    except ImportError:
        missing = [ch for ch in text if ch not in _DEMO_PINYIN]
        if missing:
            raise SystemExit("pypinyin is not installed; pass phonemes as '{j i n ...}' or install pypinyin")
        syllables = [_DEMO_PINYIN[ch] for ch in text]
    phonemes = []
    for s in syllables:
        phonemes += syllable_to_phonemes(s)
    return phonemes


def phonemes_to_ids(phonemes):
    """Unknown phonemes map to the padding id 0, as in the reference (:120-124)."""
    return np.array([SYMBOL_TO_ID.get(p, SYMBOL_TO_ID[_PAD]) for p in phonemes], dtype=np.int64)


def single_batch(args, preprocess_config):
    """The 9-tuple of synthesize_chinese_pinyin.py:262-300."""
    root = preprocess_config["path"]["preprocessed_path"]
    with open(os.path.join(root, "speakers.json")) as f:
        speaker_map = json.load(f)
    with open(os.path.join(root, "emotions.json")) as f:
        emo = json.load(f)
    arousal, valence = EMOTION_AV[args.emotion]
    ids = phonemes_to_ids(text_to_phonemes(args.text))
    name = args.output_name or f"synthesis_{args.speaker_id}_{args.emotion}"
    return ([name], [args.text], np.array([speaker_map[args.speaker_id]]), np.array([emo["emotion_dict"][args.emotion]]),
            np.array([emo["arousal_dict"][arousal]]), np.array([emo["valence_dict"][valence]]), np.array([ids]),
            np.array([len(ids)]), int(len(ids)))


def expand(values, durations):
    """utils/tools.py:163-167: per-phoneme values repeated max(0, int(d)) times -> per-frame array."""
    reps = np.maximum(np.trunc(np.asarray(durations, dtype=np.float64)).astype(np.int64), 0)
    return np.repeat(np.asarray(values), reps)


def source_batches(path, preprocess_config, batch_size=8, max_seq_len=None, mel_filter=True):
    """Batch mode: lines `basename|speaker|{p1 p2 ...}|raw_text|...|emotion|arousal|valence`
    (dataset_chinese.py:236-262,221-232); texts zero-padded per batch (:264-276).
    The reference's TextDataset.process_meta loads `<preprocessed_path>/mel/<speaker>-mel-<basename>.npy` for every
    line and drops the utterances whose recorded mel is longer than model_config["max_seq_len"] (:244-255; a missing
    file raises there).  mel_filter=True reproduces that whenever the mel directory exists; synthesis from a bare
    phoneme list (no preprocessed corpus on disk) simply has nothing to filter."""
    root = preprocess_config["path"]["preprocessed_path"]
    mel_dir = os.path.join(root, "mel")
    use_filter = mel_filter and max_seq_len is not None and os.path.isdir(mel_dir)
    with open(os.path.join(root, "speakers.json")) as f:
        speaker_map = json.load(f)
    with open(os.path.join(root, "emotions.json")) as f:
        emo = json.load(f)
    rows = []
    with open(path, encoding="utf-8") as f:
        for line in f:
            parts = line.strip("\n").split("|")
            if len(parts) < 4:
                continue
            emotion, arousal, valence = parts[-3], parts[-2], parts[-1]
            if use_filter:
                mel = np.load(os.path.join(mel_dir, "{}-mel-{}.npy".format(parts[1], parts[0])), mmap_mode="r")
                if mel.shape[0] > max_seq_len:
                    continue
            rows.append((parts[0], parts[3], speaker_map[parts[1]], emo["emotion_dict"][emotion],
                         emo["arousal_dict"][arousal], emo["valence_dict"][valence],
                         phonemes_to_ids(text_to_phonemes(parts[2]))))
    for i in range(0, len(rows), batch_size):
        chunk = rows[i:i + batch_size]
        lens = np.array([len(r[6]) for r in chunk])
        texts = np.zeros((len(chunk), int(lens.max())), dtype=np.int64)
        for j, r in enumerate(chunk):
            texts[j, : len(r[6])] = r[6]
        yield ([r[0] for r in chunk], [r[1] for r in chunk], np.array([r[2] for r in chunk]),
               np.array([r[3] for r in chunk]), np.array([r[4] for r in chunk]), np.array([r[5] for r in chunk]), texts,
               lens, int(lens.max()))


def build_parser():
    p = argparse.ArgumentParser(description="FastSpeech2 synthesis on the B200 engine (mel output)")
    p.add_argument("--restore_step", type=int, required=True)
    p.add_argument("--mode", type=str, choices=["batch", "single"], required=True)
    p.add_argument("--source", type=str, default=None)
    p.add_argument("--text", type=str, default=None)
    p.add_argument("--speaker_id", type=str, default="0001")
    p.add_argument("--emotion", type=str, default="Neutral", choices=list(EMOTION_AV))
    p.add_argument("--output_name", type=str, default=None)
    p.add_argument("-p", "--preprocess_config", type=str, required=True)
    p.add_argument("-m", "--model_config", type=str, required=True)
    p.add_argument("-t", "--train_config", type=str, required=True)
    p.add_argument("--pitch_control", type=float, default=1.0)
    p.add_argument("--energy_control", type=float, default=1.0)
    p.add_argument("--duration_control", type=float, default=1.0)
    p.add_argument("--random_init", action="store_true", help="seeded synthetic weights instead of the checkpoints")
    p.add_argument("--vocoder_ckpt", type=str, default=None, help="generator checkpoint (default hifigan/generator_<speaker>.pth.tar)")
    p.add_argument("--no_vocoder", action="store_true", help="write the mel only")
    p.add_argument("--math_mode", type=str, default="tf32", choices=["tf32", "bf16", "parity"],
                   help="arithmetic of the contractions (bf16: faster, tolerance in tests/test_gpu_bf16.py; parity: split-operand "
                        "3xTF32 acoustic model, ~1e-4 of the fp64 reference)")
    p.add_argument("--ragged_vocoder", action="store_true",
                   help="skip the padded frames in the vocoder (each utterance generated as if alone; about a third of the work "
                        "at batch 64) instead of the reference's run-padded-then-trim")
    return p


def load_model(args, preprocess_config, model_config, train_config, device="cuda"):
    """utils/model.py:11-34: construct, torch.load(<ckpt_path>/<step>.pth.tar)["model"], eval."""
    from .model import FastSpeech2B200
    model = FastSpeech2B200(preprocess_config, model_config, math_mode=args.math_mode)
    if args.random_init:
        from .synthetic import synthetic_state_dict
        model.load_state_dict(synthetic_state_dict(seed=0))
    else:
        path = os.path.join(train_config["path"]["ckpt_path"], f"{args.restore_step}.pth.tar")
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        model.load_state_dict(ckpt["model"])
    return model.to(device).eval()


def _voc_mode(args):
    return "bf16" if args.math_mode == "bf16" else "tf32"    # the parity mode is the acoustic model's


def load_vocoder(args, model_config, device="cuda"):
    """utils/model.py:37-71.  Returns None when no generator weights can be found (the reference ships none)."""
    if args.no_vocoder:
        return None
    from .vocoder import get_vocoder
    if args.random_init:
        return get_vocoder(model_config, device, random_init=True, math_mode=_voc_mode(args))
    path = args.vocoder_ckpt or os.path.join("hifigan", f"generator_{model_config['vocoder']['speaker']}.pth.tar")
    if not os.path.exists(path):
        print(f"[fs2_b200] no vocoder weights at {path}: writing mel spectrograms only")
        return None
    return get_vocoder(model_config, device, ckpt_path=path, math_mode=_voc_mode(args))


def _plot(out_dir, name, mel, feats, i, n_src, n_mel, pp):
    """The spectrogram figure of synth_samples (utils/tools.py:247-256) when matplotlib is importable (it is not part of
    this image; the curves are always written as .npy next to the mel)."""
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        return
    fig, ax = plt.subplots(1, 1)
    ax.imshow(np.asarray(mel).T, origin="lower", aspect="auto")
    ax.set_title("Synthetized Spectrogram")
    fig.savefig(os.path.join(out_dir, f"{name}.png"))
    plt.close(fig)


def write_wavs(out_dir, ids, wavs, sampling_rate):
    """utils/tools.py:268-271: one int16 wav per utterance."""
    from scipy.io import wavfile
    for name, wav in zip(ids, wavs):
        wavfile.write(os.path.join(out_dir, f"{name}.wav"), sampling_rate, wav)


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.mode == "batch":
        assert args.source is not None and args.text is None
    else:
        assert args.source is None and args.text is not None
    with open(args.preprocess_config) as f:
        preprocess_config = yaml.load(f, Loader=yaml.FullLoader)
    with open(args.model_config) as f:
        model_config = yaml.load(f, Loader=yaml.FullLoader)
    with open(args.train_config) as f:
        train_config = yaml.load(f, Loader=yaml.FullLoader)
    model = load_model(args, preprocess_config, model_config, train_config)
    vocoder = load_vocoder(args, model_config)
    hop = preprocess_config["preprocessing"]["stft"]["hop_length"]
    sampling_rate = preprocess_config["preprocessing"]["audio"]["sampling_rate"]
    batches = [single_batch(args, preprocess_config)] if args.mode == "single" else \
        source_batches(args.source, preprocess_config, max_seq_len=int(model_config["max_seq_len"]))
    pp = preprocess_config["preprocessing"]
    out_dir = train_config["path"]["result_path"]
    os.makedirs(out_dir, exist_ok=True)
    for ids, raw_texts, speakers, emotions, arousals, valences, texts, text_lens, max_len in batches:
        host = dict(speakers=speakers, emotions=emotions, arousals=arousals, valences=valences, texts=texts,
                    src_lens=text_lens, max_src_len=max_len)
        mel, mel_lens, _, _ = model.synthesize_host(host, p_control=args.pitch_control, e_control=args.energy_control,
                                                    d_control=args.duration_control)
        if vocoder is not None:        # utils/tools.py:258-271
            from .vocoder import vocoder_infer
            mels_dev = model.last_postnet.transpose(1, 2)
            if mels_dev.shape[2] > 0:
                wavs = vocoder_infer(mels_dev, vocoder, model_config, preprocess_config, lengths=[int(n) * hop for n in mel_lens],
                                     skip_padding=args.ragged_vocoder)
                write_wavs(out_dir, ids, wavs, sampling_rate)
        feats = model.last_host
        for i, name in enumerate(ids):
            np.save(os.path.join(out_dir, f"{name}.npy"), mel[i])
            # the per-frame pitch / energy curves synth_samples draws under the spectrogram (utils/tools.py:229-243):
            # phoneme_level predictions are expanded by the durations, frame_level ones are cut at mel_len
            n_src, n_mel = int(text_lens[i]), int(mel_lens[i])
            dur = feats["durations"][i, :n_src]
            for key in ("pitch", "energy"):
                v = feats[key][i]
                curve = expand(v[:n_src], dur) if pp[key]["feature"] == "phoneme_level" else v[:n_mel]
                np.save(os.path.join(out_dir, f"{name}.{key}.npy"), np.asarray(curve, dtype=np.float32))
            _plot(out_dir, name, mel[i], feats, i, n_src, n_mel, pp)
            with open(os.path.join(out_dir, f"{name}.json"), "w", encoding="utf-8") as f:
                json.dump({"text": raw_texts[i], "n_phonemes": int(text_lens[i]), "n_frames": int(mel_lens[i])}, f,
                          ensure_ascii=False)
            print(f"{name}: {int(text_lens[i])} phonemes -> {int(mel_lens[i])} mel frames -> {out_dir}/{name}.npy" +
                  (f", {out_dir}/{name}.wav" if vocoder is not None else ""))


if __name__ == "__main__":
    main()
