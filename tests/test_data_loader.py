"""Caller side of the hot path (wav2letter_pytorch_b200/data_loader.py) on a machine without a GPU: audio reading, manifests,
label encoding and collation against the reference's data/data_loader.py:19-163 -- imported verbatim where it can run here
(``_collator``), restated where it needs soundfile (``load_audio``, ``SpectrogramDataset.__getitem__``)."""
import json

import numpy as np
import pytest
import torch
from scipy.io import wavfile

from oracle import ref_loader as rl
from wav2letter_pytorch_b200 import config, data_loader as DL
from wav2letter_pytorch_b200.label_sets import english_lowercase_labels as LABELS

SR = 16000


@pytest.fixture()
def corpus(tmp_path):
    rng = np.random.default_rng(0)
    rows = []
    for i, (sec, text) in enumerate([(0.31, "hello world"), (0.52, "it's a _test_ 42!"), (0.40, "")]):
        pcm = (rng.standard_normal(int(sec * SR)) * 3000).astype(np.int16)
        path = str(tmp_path / ("utt%d.wav" % i))
        wavfile.write(path, SR, pcm)
        rows.append(dict(audio_filepath=path, text=text, pcm=pcm))
    return rows


def audio_conf():
    return config.compose().model.audio_conf


def test_load_audio_scaling_offset_duration(corpus, tmp_path):
    pcm, path = corpus[0]["pcm"], corpus[0]["audio_filepath"]
    a = DL.load_audio(path)
    assert a.dtype == np.float32 and a.shape == pcm.shape
    assert np.array_equal(a, pcm.astype(np.float32) / 32768.0)            # libsndfile's float32 read of 16-bit PCM (data_loader.py:21-28)
    b = DL.load_audio(path, duration=0.1, offset=0.05)                      # seek(int(offset*sr)); read(int(duration*sr))
    assert np.array_equal(b, a[int(0.05 * SR):int(0.05 * SR) + int(0.1 * SR)])
    c = DL.load_audio(path, duration=-1, offset=0.2)
    assert np.array_equal(c, a[int(0.2 * SR):])
    stereo = np.stack([pcm, -pcm], axis=1)
    p2 = str(tmp_path / "st.wav")
    wavfile.write(p2, SR, stereo)
    s = DL.load_audio(p2)
    assert s.shape == (2, len(pcm))                                         # samples.transpose() (data_loader.py:30)


def test_dataset_jsonl_and_csv_manifests(corpus, tmp_path):
    import pandas as pd
    man = str(tmp_path / "m.json")
    with open(man, "w") as f:
        for r in corpus:
            f.write(json.dumps(dict(audio_filepath=r["audio_filepath"], text=r["text"])) + "\n")
    ds = DL.SpectrogramDataset(man, audio_conf(), LABELS, mel_spec=64)
    assert len(ds) == 3 and ds.data_channels() == 64
    assert list(ds.df.offset) == [0, 0, 0] and list(ds.df.duration) == [-1, -1, -1]
    sig, target, path, text = ds[1]
    assert path == corpus[1]["audio_filepath"] and text == corpus[1]["text"]
    assert sig.dtype == np.float32 and sig.shape == corpus[1]["pcm"].shape
    # data_loader.py:126 -- list(filter(None, [labels_map.get(x) ...])): unknown characters AND index 0 ('_', the blank) vanish
    labels_map = {c: i for i, c in enumerate(LABELS)}
    assert target == list(filter(None, [labels_map.get(ch) for ch in corpus[1]["text"]]))
    assert 0 not in target and len(target) == len("it's a test ")
    assert ds[2][1] == []
    csv = str(tmp_path / "m.csv")
    pd.DataFrame([dict(audio_filepath=r["audio_filepath"], text=r["text"], offset=0.1, duration=0.1) for r in corpus]).to_csv(csv)
    ds2 = DL.SpectrogramDataset(csv, audio_conf(), LABELS, mel_spec=64)
    assert len(ds2) == 3 and len(ds2[0][0]) == int(0.1 * SR)


def test_dataset_rejects_wrong_sample_rate_and_missing_mel(corpus, tmp_path):
    man = str(tmp_path / "m.json")
    with open(man, "w") as f:
        f.write(json.dumps(dict(audio_filepath=corpus[0]["audio_filepath"], text="a")) + "\n")
    conf = dict(audio_conf())
    conf["sample_rate"] = 8000
    with pytest.raises(AssertionError, match="Expected sample rate 8000 but found 16000"):
        DL.SpectrogramDataset(man, conf, LABELS, mel_spec=64)
    with pytest.raises(ValueError):
        DL.SpectrogramDataset(man, audio_conf(), LABELS, mel_spec=None)


def _ragged_batch(seed=0):
    rng = np.random.default_rng(seed)
    feats = [rng.standard_normal((64, t)).astype(np.float32) for t in (31, 50, 17)]
    targets = [[3, 4, 5, 28, 7], [9], [1, 2, 3, 4, 5, 6, 7, 8]]
    return [(f, t, "p%d" % i, "text%d" % i) for i, (f, t) in enumerate(zip(feats, targets))]


def test_collator_restated():
    batch = _ragged_batch()
    inputs, il, tg, tl, paths, texts = DL._collator(batch)
    assert inputs.dtype == torch.float32 and inputs.shape == (3, 64, 50)
    assert il.dtype == torch.int32 and il.tolist() == [31, 50, 17]
    assert tg.dtype == torch.int32 and tg.shape == (3, 8) and tl.tolist() == [5, 1, 8]
    for i, (f, t, _, _) in enumerate(batch):
        assert np.array_equal(inputs[i].numpy(), np.pad(f, ((0, 0), (0, 50 - f.shape[1]))))
        assert tg[i].tolist() == t + [0] * (8 - len(t))
    assert paths == ("p0", "p1", "p2") and texts == ("text0", "text1", "text2")


@pytest.mark.skipif(not rl.reference_available(), reason="reference tree not mounted")
def test_collator_equals_reference():
    ref = rl.load_reference_features()
    batch = _ragged_batch(3)
    got, want = DL._collator(batch), ref._collator(batch)
    for g, w in zip(got[:4], want[:4]):
        assert g.dtype == w.dtype and torch.equal(g, w)
    assert got[4:] == want[4:]


def test_device_collator_and_loader_wiring():
    """DeviceCollator hands the raw signals to extract_batch in batch order and pads the targets; BatchAudioDataLoader picks it
    for a raw-audio dataset (an extractor double stands in for the CUDA front-end)"""
    calls = []

    class Extractor:
        def extract_batch(self, signals, dither=True):
            calls.append((len(signals), dither))
            lens = torch.tensor([1 + len(s) // 160 for s in signals], dtype=torch.int32)
            return torch.zeros((len(signals), 64, int(lens.max()))), lens

    class DS(torch.utils.data.Dataset):
        return_audio = True
        extractor = Extractor()

        def __len__(self):
            return 5

        def __getitem__(self, i):
            return np.zeros(1600 * (i + 1), np.float32), [1] * i, "p%d" % i, "t" * i

    loader = DL.BatchAudioDataLoader(DS(), batch_size=2, shuffle=False)
    batches = list(loader)
    assert [b[0].shape[0] for b in batches] == [2, 2, 1] and calls == [(2, True), (2, True), (1, True)]
    inputs, il, tg, tl, paths, texts = batches[1]
    assert il.tolist() == [31, 41] and inputs.shape == (2, 64, 41)
    assert tg.tolist() == [[1, 1, 0], [1, 1, 1]] and tl.tolist() == [2, 3] and paths == ("p2", "p3") and texts == ("tt", "ttt")
    with pytest.raises(ValueError):
        DL.BatchAudioDataLoader(DS(), batch_size=2, num_workers=2)
    DS.return_audio = False                                              # feature matrices in: the reference-shaped collator
    assert DL.BatchAudioDataLoader(DS(), batch_size=2).collate_fn is DL._collator
