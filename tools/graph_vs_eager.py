"""A/B on one GPU: the eager training step vs the same step replayed from a CUDA graph (graph_step.GraphedTrainStep).
usage: python tools/graph_vs_eager.py [mid_layers] [batch] [seconds] [steps]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wav2letter_pytorch_b200 import config                       # noqa: E402
from wav2letter_pytorch_b200.graph_step import GraphedTrainStep   # noqa: E402
from wav2letter_pytorch_b200.wav2letter import Wav2Letter         # noqa: E402

mid, B, sec, steps = (int(a) for a in (sys.argv[1:5] + ["20", "64", "15", "20"][len(sys.argv) - 1:]))
cfg = config.compose(overrides=["model.mid_layers=%d" % mid, "optimizer=novograd"]).model
torch.manual_seed(0)
model = Wav2Letter(cfg).cuda().train()
(opt,), _ = model.configure_optimizers()
T, S = sec * 100, sec * 15
g = torch.Generator().manual_seed(1)
x = torch.randn(B, cfg.input_size, T, generator=g).cuda()
il = torch.full((B,), T, dtype=torch.int32).cuda()
tg = torch.randint(1, len(cfg.labels), (B, S), generator=g, dtype=torch.int32).cuda()
tl = torch.full((B,), S, dtype=torch.int32).cuda()
texts = ["".join(cfg.labels[c] for c in row.tolist()) for row in tg.cpu()]
batch = (x, il, tg, tl, None, texts)


def eager(it):
    opt.zero_grad(set_to_none=True)
    loss = model.training_step(batch, it)
    loss.backward()
    opt.step()
    return loss.detach()


def timed(fn, n):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        last = fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, float(last)


res = []
res.append(("eager", ) + timed(eager, steps))
step = GraphedTrainStep(model, opt, batch, warmup=1)
res.append(("graph", ) + timed(lambda i: step(batch, i), steps))
step.close()
res.append(("eager", ) + timed(eager, steps))
for name, ms, loss in res:
    print("mid_layers=%d B=%d x %d s  %s: %.3f ms/step  loss %.4f" % (mid, B, sec, name, ms, loss))
