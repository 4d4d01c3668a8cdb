"""Where does the HOST time of a small (launch-bound) training step go?  cProfile over N steps of Wav2Letter mid_layers=M at B=64 x 15 s,
next to the CUDA-event step time; run under gpurun:  python tools/host_profile.py [mid_layers] [steps]"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from wav2letter_pytorch_b200 import config, reserve_device_memory  # noqa: E402
from wav2letter_pytorch_b200.wav2letter import Wav2Letter  # noqa: E402

mid = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda", 0)
reserve_device_memory(dev, gib=16)
x, il, tg, tl, texts = bench.synthetic_batch(64, 15, 0)
batch = tuple(t.to(dev) for t in (x, il, tg, tl)) + (None, texts)
if "--after-big" in sys.argv:                     # the bench's sequence: the 20-layer model first, dropped, then the small one
    cfg = config.compose(overrides=["model.mid_layers=20", "optimizer=novograd"]).model
    big = Wav2Letter(cfg).to(dev).train()
    (bopt,), _ = big.configure_optimizers()
    for it in range(4):
        bopt.zero_grad(set_to_none=True)
        big.training_step(batch, it).backward()
        bopt.step()
    torch.cuda.synchronize()
    del big, bopt
    if "--no-empty-cache" not in sys.argv:
        torch.cuda.empty_cache()
cfg = config.compose(overrides=["model.mid_layers=%d" % mid, "optimizer=novograd"]).model
torch.manual_seed(0)
model = Wav2Letter(cfg).to(dev).train()
(opt,), _ = model.configure_optimizers()


def step(it):
    opt.zero_grad(set_to_none=True)
    loss = model.training_step(batch, it)
    loss.backward()
    opt.step()


for it in range(5):
    step(it)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for it in range(steps):
    step(it)
e1.record()
host_ms = (time.perf_counter() - t0) / steps * 1e3
torch.cuda.synchronize()
print("mid_layers=%d: device %.3f ms/step, host enqueue %.3f ms/step" % (mid, e0.elapsed_time(e1) / steps, host_ms))
pr = cProfile.Profile()
pr.enable()
for it in range(steps):
    step(it)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(25)
