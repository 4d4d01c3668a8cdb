"""GPU feature front-end timing: B utterances of S seconds of 16 kHz audio -> [B, 64, T] normalised log-mel features."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wav2letter_pytorch_b200.features import SpectrogramExtractor  # noqa: E402

B, S = (int(v) for v in (sys.argv[1:3] if len(sys.argv) >= 3 else (64, 15)))
ex = SpectrogramExtractor(dict(sample_rate=16000, window_size=0.02, window_stride=0.01, window="hamming"), mel_spec=64).cuda()
audio = 0.1 * torch.randn(B, 16000 * S, device="cuda")
noise = torch.randn_like(audio)
for _ in range(3):
    out, lens = ex.extract_batch(audio, noise=noise)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out, lens = ex.extract_batch(audio, noise=noise)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("front-end B=%d x %d s: %.3f ms per batch (%.0f audio-s/s), output %s" % (B, S, ms, B * S / (ms / 1e3), tuple(out.shape)))
