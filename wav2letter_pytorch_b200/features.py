"""Feature front-end with the reference's interface (data/data_loader.py:33-88): ``SpectrogramExtractor(audio_conf,
mel_spec)`` with ``extract(signal)``, plus ``extract_batch(signals)`` which produces what the reference's ``_collator``
(data_loader.py:149-158) hands to the model -- ``inputs [B, F, T_max]`` fp32 zero padded and ``input_lengths [B]`` int32 --
directly on the GPU, for the whole batch in two kernel launches (csrc/features.cu) instead of one single-threaded
torch.stft + matmul + numpy pass per utterance inside the DataLoader.

    dither (1e-5 * N(0,1)) -> pre-emphasis 0.97 -> STFT(n_fft = 2^ceil(log2(win)), hop, hamming(win), center, reflect) ->
    |X|^2 -> mel filterbank (Slaney, librosa.filters.mel defaults) -> log1p(. + 2^-24) -> per-feature (v - mean) / (std + 1e-5)

There is no CPU fallback: CPU inputs are copied to the current CUDA device."""
import math

import numpy as np
import torch

from . import _lib
from . import functional as F


def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sample_rate, n_fft, n_mels, fmin=0.0, fmax=None):
    """[n_mels, n_fft/2+1] float32 triangular filters on the Slaney mel scale with Slaney area normalisation -- the published
    algorithm of ``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm='slaney')``, which the reference calls at
    data_loader.py:40-44 (librosa itself is not a dependency of this package)."""
    fmax = sample_rate / 2.0 if fmax is None else fmax
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0.0, sample_rate / 2.0, n_bins)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return (weights * enorm[:, None]).astype(np.float32)


_WINDOWS = {"hann": torch.hann_window, "hamming": torch.hamming_window, "blackman": torch.blackman_window,
            "bartlett": torch.bartlett_window}


class SpectrogramExtractor(torch.nn.Module):
    dithering = 1e-5
    preemph = 0.97
    log_zero_guard_value = 2.0 ** -24
    epsilon = 1e-5

    def __init__(self, audio_conf, mel_spec=64, use_cuda=True):
        super().__init__()
        get = audio_conf.get if hasattr(audio_conf, "get") else (lambda k, d=None: getattr(audio_conf, k, d))
        self.sample_rate = int(get("sample_rate"))
        self.win_length = int(self.sample_rate * get("window_size"))
        self.hop_length = int(self.sample_rate * get("window_stride"))
        self.n_fft = 2 ** math.ceil(math.log2(self.win_length))
        if not mel_spec:
            raise ValueError("SpectrogramExtractor: mel_spec (number of mel bins) is required")
        self.n_mels = int(mel_spec)
        self.register_buffer("fb", torch.from_numpy(mel_filterbank(self.sample_rate, self.n_fft, self.n_mels, 0.0, self.sample_rate / 2))
                             .unsqueeze(0))
        window_fn = _WINDOWS.get(get("window"), None)
        if window_fn is None:
            raise ValueError("SpectrogramExtractor: unsupported window %r" % (get("window"),))   # the reference dies in torch.stft here
        self.register_buffer("window", window_fn(self.win_length, periodic=False).float())

    def n_frames(self, n_samples):
        """torch.stft(center=True): 1 + n_samples // hop."""
        return 1 + int(n_samples) // self.hop_length

    # ---- whole batch: list of 1-D signals (or a padded [B, L] tensor + lengths) -> (inputs [B,F,T_max], input_lengths [B])
    def extract_batch(self, signals, lengths=None, noise=None, dither=True):
        dev = self.fb.device
        if dev.type != "cuda":
            if not torch.cuda.is_available():
                raise RuntimeError("SpectrogramExtractor: a CUDA device is required (no CPU path); call .cuda() first")
            self.cuda()
            dev = self.fb.device
        if torch.is_tensor(signals) and signals.dim() == 2:
            audio = signals.to(device=dev, dtype=torch.float32).contiguous()
            lens = torch.as_tensor(lengths if lengths is not None else [audio.shape[1]] * audio.shape[0], dtype=torch.int32)
        else:
            sigs = [torch.as_tensor(np.asarray(s) if not torch.is_tensor(s) else s, dtype=torch.float32).reshape(-1) for s in signals]
            lens = torch.tensor([s.numel() for s in sigs], dtype=torch.int32)
            # pinned staging buffer, kept across calls (cudaHostAlloc per batch would cost more than the kernels); the previous
            # upload must have finished before it is overwritten
            need = (len(sigs), int(lens.max()))
            stage = getattr(self, "_host_stage", None)
            if stage is None or stage[0].numel() < need[0] * need[1]:
                stage = [torch.zeros((need[0] * need[1],), dtype=torch.float32).pin_memory(), None]
                self._host_stage = stage
            if stage[1] is not None:
                stage[1].synchronize()
            host = stage[0][:need[0] * need[1]].view(need)          # contiguous [B, Lmax] view of the flat buffer
            host.zero_()
            for i, s in enumerate(sigs):
                host[i, :s.numel()] = s
            audio = host.to(dev, non_blocking=True)
            stage[1] = torch.cuda.current_stream(dev).record_event()
        if int(lens.min()) <= self.n_fft // 2:
            raise RuntimeError("SpectrogramExtractor: every signal must be longer than n_fft/2 = %d samples (reflect padding, as torch.stft)"
                               % (self.n_fft // 2))
        B, Lmax = audio.shape
        T_max = self.n_frames(int(lens.max()))
        if noise is None and dither:
            noise = torch.randn((B, Lmax), device=dev, dtype=torch.float32)       # the reference draws torch.randn(audio.shape) per utterance
        elif noise is not None:
            noise = noise.to(device=dev, dtype=torch.float32).contiguous()
        lens_d = lens.to(dev)
        lib = _lib.load()
        out = torch.empty((B, self.n_mels, T_max), dtype=torch.float32, device=dev)
        ws = torch.empty((lib.w2l_logmel_workspace_bytes(B, T_max, self.n_mels),), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.w2l_logmel_features(F._ptr(audio), audio.stride(0), F._ptr(noise), F._ptr(lens_d), B, self.n_fft, self.win_length,
                                               self.hop_length, F._ptr(self.window), F._ptr(self.fb), self.n_mels, self.dithering,
                                               self.preemph, self.log_zero_guard_value, self.epsilon, F._ptr(out), T_max, F._ptr(ws),
                                               ws.numel(), F._stream()), "logmel_features")
        return out, (lens_d // self.hop_length + 1).to(torch.int32)

    # ---- reference-shaped single-utterance call (data_loader.py:76-88): [F, T] normalised log-mel features
    def extract(self, signal, noise=None, dither=True):
        out, _ = self.extract_batch([signal], noise=None if noise is None else torch.as_tensor(noise).reshape(1, -1), dither=dither)
        return out[0]
