"""Event timeline of one train step (no profiler): every library call is bracketed by CUDA events on its launching stream.
    python tools/timeline.py [--mid-layers 20] [--from 0 --to 1e9]  ->  start_ms  dur_ms  stream  call"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import BATCH, UTT_SEC, synthetic_batch  # noqa: E402
from wav2letter_pytorch_b200 import config  # noqa: E402
from wav2letter_pytorch_b200 import functional as F  # noqa: E402
from wav2letter_pytorch_b200.wav2letter import Wav2Letter  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mid-layers", type=int, default=20)
ap.add_argument("--lo", type=float, default=0.0)
ap.add_argument("--hi", type=float, default=1e9)
args = ap.parse_args()
dev = torch.device("cuda", 0)
cfg = config.compose(overrides=["model.mid_layers=%d" % args.mid_layers, "optimizer=novograd"]).model
torch.manual_seed(0)
model = Wav2Letter(cfg).to(dev).train()
(opt,), _ = model.configure_optimizers()
x, il, tg, tl, texts = synthetic_batch(BATCH, UTT_SEC, 0)
batch = tuple(t.to(dev) for t in (x, il, tg, tl)) + (None, texts)


def step():
    opt.zero_grad(set_to_none=True)
    loss = model.training_step(batch, 0)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
spans = []
names = [n for n in dir(F) if callable(getattr(F, n)) and not n.startswith("_") and n not in ("make_desc", "ConvDesc")]
for n in names:
    def make(fn, tag):
        def timed(*a, **k):
            st = torch.cuda.current_stream()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            r = fn(*a, **k)
            e1.record(st)
            spans.append((tag, st.cuda_stream, e0, e1))
            return r
        return timed
    setattr(F, n, make(getattr(F, n), n))
base = torch.cuda.Event(enable_timing=True)
base.record()
step()
end = torch.cuda.Event(enable_timing=True)
end.record()
torch.cuda.synchronize()
streams = {}
print("step %.3f ms" % base.elapsed_time(end))
for tag, st, e0, e1 in sorted(spans, key=lambda s: base.elapsed_time(s[2])):
    t0 = base.elapsed_time(e0)
    if args.lo <= t0 <= args.hi:
        print("%9.3f %8.3f  s%d  %s" % (t0, e0.elapsed_time(e1), streams.setdefault(st, len(streams)), tag))
