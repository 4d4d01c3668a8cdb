"""Library yard-stick: the SAME Wav2Letter-20 training shape run through torch's own CUDA path on this GPU --
nn.ReflectionPad1d / nn.Conv1d (cuDNN) / nn.BatchNorm1d / Dropout / clamp, F.log_softmax, nn.CTCLoss, backward -- in fp32 (TF32
allowed) and under bf16 autocast.  No greedy decode (the reference's per-frame .item() loop would add ~10^5 device syncs per step)
and no optimizer step, i.e. a generous lower bound for the stock path.  Allowed for comparison only; nothing here is on the product
path.   python tools/yardstick_torch_cuda.py [--steps 5]"""
import argparse
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wav2letter_pytorch_b200 import config  # noqa: E402  (only the yaml composer: layer list of the default config)


class Block(nn.Module):
    def __init__(self, cin, cout, k, s, d, p, bn=True):
        super().__init__()
        rows = (cin + s - 1) // s
        pad = max(0, (rows - 1) * s + (k - 1) * d + 1 - cin)                 # wav2letter.py:24-34
        self.pad = nn.ReflectionPad1d((pad // 2, (pad + 1) // 2)) if pad else nn.Identity()
        self.conv = nn.Conv1d(cin, cout, k, s, dilation=d)
        self.bn = nn.BatchNorm1d(cout, eps=1e-3, momentum=0.9) if bn else None
        self.drop = nn.Dropout(p) if p and p > 0 else nn.Identity()

    def forward(self, x):
        x = self.conv(self.pad(x))
        if self.bn is None:
            return x
        return torch.clamp(self.drop(self.bn(x)), 0, 20)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    cfg = config.compose(overrides=["model.mid_layers=20"]).model
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    blocks, width = [], cfg.input_size
    for lp in cfg.layers[:20]:
        blocks.append(Block(width, lp.output_size, lp.kernel_size, lp.stride, lp.dilation, lp.dropout))
        width = lp.output_size
    blocks.append(Block(width, 29, 1, 1, 1, 0, bn=False))
    model = nn.Sequential(*blocks).to(dev).train()
    B, T, S = 64, 1501, 225
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, 64, T, generator=g).to(dev)
    tg = torch.randint(1, 29, (B, S), generator=g, dtype=torch.int32).to(dev)
    il = torch.full((B,), T // 2, dtype=torch.int32, device=dev)
    tl = torch.full((B,), S, dtype=torch.int32, device=dev)
    crit = nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)

    def step(amp):
        model.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            out = model(x)
        lp = torch.log_softmax(out.float().transpose(1, 2), -1)
        loss = crit(lp.transpose(0, 1), tg, il, tl)
        loss.backward()

    for amp in (False, True):
        for _ in range(3):
            step(amp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step(amp)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print("torch CUDA path (%s): %.1f ms per fwd+CTC+bwd step, %.0f audio-s/s  [B=64 x 15 s, W2L-20, no decode, no optimizer]"
              % ("bf16 autocast" if amp else "fp32/TF32", ms, B * 15 / (ms / 1e3)))


if __name__ == "__main__":
    main()
