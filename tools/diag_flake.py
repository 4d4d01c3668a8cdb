"""Reproducibility probe: run the same small train step many times and report, per library call, how far its outputs move
between runs (max-norm relative).  Localises rare races / uninitialised reads.   python tools/diag_flake.py [runs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wav2letter_pytorch_b200 import config, layers  # noqa: E402
from wav2letter_pytorch_b200 import functional as F  # noqa: E402
from wav2letter_pytorch_b200.wav2letter import Wav2Letter  # noqa: E402

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = torch.device("cuda", 0)
cfg = config.compose(overrides=["model.mid_layers=3", "optimizer=novograd"]).model
torch.manual_seed(0)
model = Wav2Letter(cfg).to(dev).train()
g = torch.Generator().manual_seed(100)
B, T, S = 4, 301, 20
x = torch.randn(B, 64, T, generator=g).to(dev)
il = torch.full((B,), T, dtype=torch.int32, device=dev)
tg = torch.randint(1, 29, (B, S), generator=g, dtype=torch.int32).to(dev)
tl = torch.full((B,), S, dtype=torch.int32, device=dev)

log = []
names = ["conv1d_fwd", "bn_finalize", "bn_act_pad", "log_softmax", "ctc_loss_raw", "log_softmax_bwd", "conv1d_dgrad_wt", "conv1d_wgrad",
         "bn_act_bwd", "colsum", "im2col_ncw"]
orig = {n: getattr(F, n) for n in names}


def flat(r):
    if torch.is_tensor(r):
        return [r]
    if isinstance(r, (tuple, list)):
        return [t for t in r if torch.is_tensor(t)]
    return []


for n in names:
    def make(fn, tag):
        def wrapped(*a, **k):
            r = fn(*a, **k)
            torch.cuda.synchronize()
            log.append((tag, [t.detach().float().clone() for t in flat(r)]))
            return r
        return wrapped
    setattr(F, n, make(orig[n], n))


def step():
    layers._seed_counter[0] = 0
    log.clear()
    model.zero_grad(set_to_none=True)
    out, ol = model(x, il)
    loss = model.criterion(out.transpose(0, 1), tg, ol, tl)
    loss.backward()
    torch.cuda.synchronize()
    return list(log)


ref = step()
worst_seen = {}
for r in range(runs):
    cur = step()
    first = None
    for i, ((tag, outs), (_, routs)) in enumerate(zip(cur, ref)):
        for a, b in zip(outs, routs):
            e = float((a - b).abs().max() / (b.abs().max() + 1e-20))
            key = "%02d:%s" % (i, tag)
            worst_seen[key] = max(worst_seen.get(key, 0.0), e)
            if e > 0.02 and first is None:
                first = (key, e)
    if first:
        print("run %d: first call deviating > 2%%: %s (%.3f)" % (r, first[0], first[1]), flush=True)
print("worst deviation per call over %d runs:" % runs)
for k in sorted(worst_seen):
    print("  %-24s %.5f" % (k, worst_seen[k]))
