"""Per-layer timing of the conv kernels (CUDA events, L2 flushed between reps) -- development aid, run under gpurun."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wav2letter_pytorch_b200 import functional as F

B, T = 64, 750
LAYERS = {"L3": (256, 256, 11, 1), "L6": (384, 384, 13, 1), "L9": (512, 512, 17, 1), "L12": (640, 640, 21, 1), "L15": (768, 768, 25, 1),
          "L17": (896, 896, 29, 2), "L19": (896, 1024, 1, 1)}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


names = sys.argv[1:] or list(LAYERS)
print("layer   GFLOP | fwd ms TF/s | dgrad ms TF/s | dgrad-as-fwd(K-major Wt) ms TF/s | wgrad ms TF/s (splits)")
for name in names:
    ci, co, k, d = LAYERS[name]
    pad = (k - 1) * d
    Tp = T + pad
    xp = torch.randn(B, Tp, ci, device="cuda").to(torch.bfloat16)
    w = (torch.randn(k, co, ci, device="cuda") / (ci * k) ** 0.5).to(torch.bfloat16)
    wt = w.flip(0).transpose(1, 2).contiguous()                 # [k, ci, co], taps reversed
    dz = torch.randn(B, T, co, device="cuda").to(torch.bfloat16)
    z = torch.empty(B, T, co, dtype=torch.bfloat16, device="cuda")
    dx = torch.empty(B, Tp, ci, dtype=torch.bfloat16, device="cuda")
    dx2 = torch.empty_like(dx)
    dw = torch.empty(k, co, ci, dtype=torch.float32, device="cuda")
    desc = F.make_desc(B, T, ci, co, co, k, d, Tp, 0, T, 0, co)
    # dgrad expressed as a forward conv over dz with transposed/flipped weights: out rows = Tp, input rows = T, offset -(k-1)d
    desc_t = F.make_desc(B, Tp, co, ci, ci, k, d, T, -pad, Tp, 0, ci)
    fl = 2.0 * B * T * ci * co * k
    t_f = timeit(lambda: F.conv1d_fwd(xp, w, desc, z))
    t_d = timeit(lambda: F.conv1d_dgrad(dz, w, desc, dx))
    t_e = timeit(lambda: F.conv1d_fwd(dz, wt, desc_t, dx2))
    t_w = timeit(lambda: F.conv1d_wgrad(dz, xp, desc, dw))
    import ctypes
    from wav2letter_pytorch_b200 import _lib
    sp = _lib.load().w2l_conv1d_wgrad_splits(ctypes.byref(desc))
    err = (dx.float() - dx2.float()).abs().max().item()
    tf = lambda t: fl / t / 1e9
    print("%-5s %7.1f | %6.3f %5.0f | %6.3f %5.0f | %6.3f %5.0f (maxdiff %.3g) | %6.3f %5.0f (%d)" % (name, fl / 1e9, t_f, tf(t_f), t_d, tf(t_d), t_e, tf(t_e), err, t_w, tf(t_w), sp))
