"""Audio files -> manifest -> BatchAudioDataLoader -> the batch the model's training_step takes, features made on the GPU
(wav2letter_pytorch_b200/data_loader.py vs the reference's data/data_loader.py:89-163 restated by the oracle).

Sorts last for the same reason as test_gpu_zz_strided.py: written after the last GPU session of round 1."""
import json

import numpy as np
import pytest
import torch
from scipy.io import wavfile

from oracle import w2l_oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def test_loader_batches_match_oracle_features(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from wav2letter_pytorch_b200 import config, data_loader as DL
    from wav2letter_pytorch_b200.label_sets import english_lowercase_labels as labels
    rs = np.random.RandomState(11)
    texts = ["hello world", "it's fine", "a", "the quick brown fox"]
    man = str(tmp_path / "m.json")
    pcm = []
    with open(man, "w") as f:
        for i, (n, text) in enumerate(zip((16000, 9000, 12345, 20000), texts)):
            pcm.append((rs.randn(n) * 4000).astype(np.int16))
            path = str(tmp_path / ("u%d.wav" % i))
            wavfile.write(path, 16000, pcm[-1])
            f.write(json.dumps(dict(audio_filepath=path, text=text)) + "\n")
    conf = config.compose().model.audio_conf
    ds = DL.SpectrogramDataset(man, conf, labels, mel_spec=64)
    loader = DL.BatchAudioDataLoader(ds, batch_size=2, shuffle=False)
    loader.collate_fn.dither = False                                       # deterministic: no dither on either side
    seen = 0
    for inputs, il, tg, tl, paths, got_texts in loader:
        assert inputs.is_cuda and il.is_cuda and tg.is_cuda and tl.is_cuda
        assert inputs.dtype == torch.float32 and il.dtype == tg.dtype == tl.dtype == torch.int32
        sigs = [p.astype(np.float32) / 32768.0 for p in pcm[seen:seen + 2]]
        want, want_lens = O.collate_features([O.spectrogram_extract(s) for s in sigs])
        assert il.cpu().tolist() == want_lens.tolist()
        assert inputs.shape == want.shape and (inputs.cpu() - want).abs().max() < 2e-4
        for i, text in enumerate(texts[seen:seen + 2]):
            ids = [labels.index(ch) for ch in text]
            assert tg[i, :len(ids)].cpu().tolist() == ids and int(tl[i]) == len(ids) and (tg[i, len(ids):] == 0).all()
        assert list(got_texts) == texts[seen:seen + 2]
        seen += 2
    assert seen == 4


def test_device_prefetcher_hands_out_the_batches_in_order():
    """DevicePrefetcher: pinned host batches -> device batches, next copy in flight on a side stream; values, order, pass-through of
    non-tensor entries, StopIteration at the end"""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from wav2letter_pytorch_b200.data_loader import DevicePrefetcher
    g = torch.Generator().manual_seed(0)
    host = [(torch.randn(4, 64, 300, generator=g).pin_memory(), torch.randint(1, 300, (4,), generator=g, dtype=torch.int32).pin_memory(), ["a", "b"], i)
            for i in range(5)]
    got = []
    for x, il, texts, idx in DevicePrefetcher(iter(host), "cuda:0"):
        assert x.is_cuda and il.is_cuda and texts == ["a", "b"]
        got.append(((x * 2).sum().item(), il.clone(), idx))          # some work on the compute stream between batches
    assert [i for _, _, i in got] == list(range(5))
    for (s, il, i), (hx, hil, _, _) in zip(got, host):
        assert abs(s - float((hx * 2).sum())) < 1e-2 * max(1.0, abs(s)) and torch.equal(il.cpu(), hil)
