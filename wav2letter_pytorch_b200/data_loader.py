"""The caller side of the hot path with the reference's surface (data/data_loader.py:19-163): ``load_audio``,
``SpectrogramDataset(manifest_filepath, audio_conf, labels, mel_spec)``, ``_collator`` and ``BatchAudioDataLoader``.

What differs, B200-first: the reference extracts features per utterance inside ``Dataset.__getitem__`` (single-threaded
torch.stft + matmul + numpy on the host) and its collator pads the finished feature matrices.  Here the dataset hands out the
RAW SIGNAL by default and ``DeviceCollator`` turns the whole batch into what the model takes -- ``inputs [B, F, T_max]`` fp32
zero padded, ``input_lengths``, ``targets [B, S_max]`` int32 zero padded, ``target_lengths`` -- with ONE pinned upload and the
two feature kernels (features.SpectrogramExtractor.extract_batch); all four tensors are born on the GPU.  The reference-shaped
``_collator`` over already extracted feature matrices is kept for callers that bring their own features.

Manifests are read as the reference reads them: ``.csv`` through ``pandas.read_csv(index_col=0)``, anything else as JSON
lines with ``audio_filepath`` and ``text`` (optional ``offset`` / ``duration`` in seconds)."""
import json

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

from .features import SpectrogramExtractor


def _read_wav(path):
    """(samples float32 [frames] or [frames, channels], sample_rate) -- integer PCM scaled to [-1, 1) as libsndfile does for
    dtype='float32' (data_loader.py:20-31 reads through soundfile, which is optional here)."""
    try:
        import soundfile as sf
    except Exception:  # noqa: BLE001  (not installed: plain RIFF/WAVE through scipy)
        sf = None
    if sf is not None and hasattr(sf, "read"):
        data, sr = sf.read(path, dtype="float32")
        return data, sr
    from scipy.io import wavfile
    sr, data = wavfile.read(path)
    if data.dtype == np.int16:
        data = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        data = (data.astype(np.float64) / 2147483648.0).astype(np.float32)
    elif data.dtype == np.uint8:
        data = (data.astype(np.float32) - 128.0) / 128.0
    else:
        data = data.astype(np.float32)
    return data, sr


def load_audio(path, duration=-1, offset=0):
    """data_loader.py:19-31: float32 samples, ``offset`` / ``duration`` in seconds (duration <= 0: to the end); multi-channel
    files come back as [channels, frames] like the reference's ``samples.transpose()``."""
    data, sr = _read_wav(path)
    start = int(offset * sr) if offset > 0 else 0
    stop = start + int(duration * sr) if duration > 0 else None
    return data[start:stop].transpose()


def wav_sample_rate(path):
    return _read_wav(path)[1]


class SpectrogramDataset(Dataset):
    """data_loader.py:89-147.  ``return_audio=True`` (default): items are ``(signal float32 [L], target, path, text)`` and the
    features are made per batch on the GPU by ``DeviceCollator``; ``return_audio=False``: items carry the [F, T] feature matrix
    of ``SpectrogramExtractor.extract`` (one utterance per call, still on the GPU -- there is no host feature path)."""

    def __init__(self, manifest_filepath, audio_conf, labels, mel_spec=None, use_cuda=False, return_audio=True):
        super().__init__()
        import pandas as pd
        if manifest_filepath.endswith(".csv"):
            self.df = pd.read_csv(manifest_filepath, index_col=0)
        else:
            with open(manifest_filepath) as f:
                self.df = pd.DataFrame([json.loads(line) for line in f if line.strip()])
        if "offset" not in self.df.columns:
            self.df["offset"] = 0
        if "duration" not in self.df.columns:
            self.df["duration"] = -1
        self.size = len(self.df)
        self.window_stride = audio_conf["window_stride"]
        self.window_size = audio_conf["window_size"]
        self.sample_rate = audio_conf["sample_rate"]
        self.use_cuda, self.mel_spec, self.return_audio = use_cuda, mel_spec, return_audio
        self.labels_map = {labels[i]: i for i in range(len(labels))}
        self.validate_sample_rate()
        if not mel_spec:                                  # the reference's extractor always applies a mel filterbank (data_loader.py:38-46)
            raise ValueError("SpectrogramDataset: mel_spec (number of mel bins) is required")
        self.extractor = SpectrogramExtractor(audio_conf, mel_spec, use_cuda)

    def encode(self, transcript):
        """data_loader.py:126: characters outside the label set AND label 0 (the blank) are dropped"""
        return [i for i in (self.labels_map.get(ch) for ch in transcript) if i]

    def __getitem__(self, index):
        sample = self.df.iloc[index]
        audio_path, transcript = sample.audio_filepath, sample.text
        signal = load_audio(audio_path, sample.duration, sample.offset)
        item = np.ascontiguousarray(signal, dtype=np.float32) if self.return_audio else self.extractor.extract(signal)
        return item, self.encode(transcript), audio_path, transcript

    def parse_audio(self, audio_path, duration, offset):
        return self.extractor.extract(load_audio(audio_path, duration, offset))

    def validate_sample_rate(self):
        sr = wav_sample_rate(self.df.iloc[0].audio_filepath)
        assert sr == self.sample_rate, "Expected sample rate %d but found %d in first file" % (self.sample_rate, sr)

    def __len__(self):
        return self.size

    def data_channels(self):
        """How many channels are returned in each example."""
        return self.mel_spec or int(1 + (int(self.sample_rate * self.window_size) / 2))


def pad_targets(targets):
    """list of int lists -> (targets int32 [B, S_max] zero padded, target_lengths int32 [B]) -- data_loader.py:152-157"""
    lengths = torch.tensor([len(t) for t in targets], dtype=torch.int32)
    out = torch.zeros((len(targets), int(lengths.max()) if len(targets) else 0), dtype=torch.int32)
    for i, t in enumerate(targets):
        if len(t):
            out[i, :len(t)] = torch.as_tensor(t, dtype=torch.int32)
    return out, lengths


def _collator(batch):
    """data_loader.py:149-158 over already extracted [F, T_i] feature matrices (host or device): zero pads to the longest."""
    inputs, targets, file_paths, texts = zip(*batch)
    inputs = [torch.as_tensor(x, dtype=torch.float32) for x in inputs]
    input_lengths = torch.tensor([x.shape[1] for x in inputs], dtype=torch.int32)
    longest = int(input_lengths.max())
    padded = inputs[0].new_zeros((len(inputs), inputs[0].shape[0], longest))
    for i, x in enumerate(inputs):
        padded[i, :, :x.shape[1]] = x
    tg, target_lengths = pad_targets(targets)
    return padded, input_lengths, tg, target_lengths, file_paths, texts


class DeviceCollator:
    """Raw signals in, the model's batch out, features made on the GPU for the whole batch at once.  Use with ``num_workers=0``
    (or let workers only read audio and call this in the training process): it launches CUDA work."""

    def __init__(self, extractor, dither=True):
        self.extractor, self.dither = extractor, dither

    def __call__(self, batch):
        signals, targets, file_paths, texts = zip(*batch)
        inputs, input_lengths = self.extractor.extract_batch(list(signals), dither=self.dither)
        tg, target_lengths = pad_targets(targets)
        dev = inputs.device
        return inputs, input_lengths, tg.to(dev, non_blocking=True), target_lengths.to(dev, non_blocking=True), file_paths, texts


class BatchAudioDataLoader(DataLoader):
    """data_loader.py:160-163.  Picks ``DeviceCollator`` for a dataset that hands out raw audio, ``_collator`` otherwise."""

    def __init__(self, dataset, *args, **kwargs):
        kwargs.pop("collate_fn", None)
        if getattr(dataset, "return_audio", False):
            collate = DeviceCollator(dataset.extractor)
            if kwargs.get("num_workers", 0) != 0:
                raise ValueError("BatchAudioDataLoader: the device collator launches CUDA work; use num_workers=0")
        else:
            collate = _collator
        super().__init__(dataset, *args, collate_fn=collate, **kwargs)


class DevicePrefetcher:
    """Hands out the batches of ``loader`` (tuples whose tensors live in PINNED host memory, as ``DataLoader(pin_memory=True)`` or the
    reference's collated batches after ``.pin_memory()`` give them) as device batches -- with the host->device copy of batch i+1
    enqueued on a side stream the moment batch i is handed out, so that it runs beside step i's kernels instead of in front of step
    i+1's (24.6 MB per 64 x 15 s batch: ~1 ms of PCIe time per step when it is not hidden).  The compute stream waits for the copy
    stream before a batch is returned, and the batch's memory is tied to the compute stream (``record_stream``), so the consumer
    needs no extra care.  Non-tensor entries (paths, transcripts) pass through.  Lightning does the same for the reference
    (train.py hands the DataLoader to ``pl.Trainer``, which moves batches asynchronously)."""

    def __init__(self, loader, device):
        self.it = iter(loader)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._next = None
        self._preload()

    def _preload(self):
        try:
            host = next(self.it)
        except StopIteration:
            self._next = None
            return
        with torch.cuda.stream(self.stream):
            self._next = tuple(t.to(self.device, non_blocking=True) if torch.is_tensor(t) else t for t in host)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.stream)
        batch = self._next
        for t in batch:
            if torch.is_tensor(t):
                t.record_stream(cur)
        self._preload()
        return batch
