import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wav2letter_pytorch_b200 import functional as F
N, T, S, C = [int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (64, 750, 225, 29))]
g = torch.Generator(device="cuda").manual_seed(0)
lp = torch.log_softmax(torch.randn(N, T, C, generator=g, device="cuda"), -1)
tg = torch.randint(1, C, (N, S), generator=g, device="cuda", dtype=torch.int32)
il = torch.full((N,), T, dtype=torch.int32, device="cuda")
tl = torch.full((N,), S, dtype=torch.int32, device="cuda")
for _ in range(3):
    F.ctc_loss_raw(lp, tg, il, tl)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    F.ctc_loss_raw(lp, tg, il, tl)
e1.record(); torch.cuda.synchronize()
print("ctc N=%d T=%d S=%d: %.3f ms per call" % (N, T, S, e0.elapsed_time(e1) / 5))
