"""BASELINE config 5: standalone CTC loss+grad and greedy decode sweep (T x S x N, C=29, fp32) with achieved HBM GB/s
(algorithmic bytes of SURVEY 8d / CUDA-event time) next to torch's own CUDA ctc_loss fwd+bwd as the library yard-stick.
Run under gpurun; writes a markdown table to stdout."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as TF
from wav2letter_pytorch_b200 import functional as F

C = 29
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


quick = "--quick" in sys.argv
Ts = [200, 500, 1000, 2000, 3000]
Ss = [50, 100, 300, 600]
Ns = [1, 8, 64, 512]
if quick:
    Ts, Ss, Ns = [750, 3000], [225, 600], [64, 512]
print("| N | T | S | ctc ours ms | GB/s | torch ctc ms | speed-up | decode ours ms | GB/s |")
print("|---|---|---|---|---|---|---|---|---|")
for N in Ns:
    for T in Ts:
        g = torch.Generator(device="cuda").manual_seed(N + T)
        logits = torch.randn(N, T, C, generator=g, device="cuda")
        lp = torch.log_softmax(logits, -1)
        il = torch.full((N,), T, dtype=torch.int32, device="cuda")
        t_dec = timeit(lambda: F.greedy_decode(lp, il))
        dec_bytes = N * T * C * 4 + N * T * 4 + N * 4
        for S in Ss:
            if S > T // 2:
                continue
            tg = torch.randint(1, C, (N, S), generator=g, device="cuda", dtype=torch.int32)
            tl = torch.full((N,), S, dtype=torch.int32, device="cuda")
            t_ours = timeit(lambda: F.ctc_loss_raw(lp, tg, il, tl))
            lpr = lp.detach().clone().requires_grad_(True)

            def torch_ctc():
                lpr.grad = None
                TF.ctc_loss(lpr.transpose(0, 1), tg, il, tl, blank=0, reduction="mean", zero_infinity=True).backward()
            try:
                t_torch = timeit(torch_ctc)
            except RuntimeError:
                t_torch = float("nan")
            ctc_bytes = N * T * C * 8 + N * S * 4 + 12 * N
            print("| %d | %d | %d | %.3f | %.1f | %.3f | %.2fx | %.3f | %.1f |" % (N, T, S, t_ours, ctc_bytes / t_ours / 1e6, t_torch, t_torch / t_ours,
                                                                            t_dec, dec_bytes / t_dec / 1e6))
