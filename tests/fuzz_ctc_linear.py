"""Manual differential fuzzing of the CTC schedules on the host emulation (not collected by pytest):

    python tests/fuzz_ctc_linear.py [cases=100] [seed=0]

Random shapes (N 1-3, T 1-700, S 0-300, C 2-100), logit scales 0.1 ... 60 (from near-uniform rows to rows peaked far beyond the fp32
range), repeated labels, ragged lengths, classes with probability exactly 0.  Every case runs through ``w2l_ctc_loss`` compiled for the
host (tests/_emu_cabi.py) under W2L_CTC_LINEAR = 0 (log space, the default), 1 (linear domain + log-space redo) and 2 (linear domain
alone) and is held to torch's fp64 ``ctc_loss`` at the GPU tests' bars (loss 1e-4 relative, gradient 2e-3 of its largest element).
What must hold: mode 1 passes wherever mode 0 passes (the range checks catch every case mode 2 gets wrong).  Cases that fail in mode 0
too are the fuzzer's own extremes (-inf clamped to -1e30 is finite for torch but log(0) for the kernels; scale-60 rows at the edge of
fp32 log values)."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import pytest  # noqa: E402
import torch  # noqa: E402

import _emu_cabi  # noqa: E402
from oracle import w2l_oracle as O  # noqa: E402


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    F = _emu_cabi.install(pytest.MonkeyPatch())
    fails = {"0": set(), "1": set(), "2": set()}
    for it in range(cases):
        N, C = rng.choice([1, 2, 3]), rng.choice([2, 3, 5, 29, 29, 29, 40, 100])
        S, T = rng.choice([0, 1, 2, 5, 17, 40, 70, 130, 200, 300]), rng.choice([1, 2, 7, 33, 64, 65, 150, 400, 700])
        scale = rng.choice([0.1, 1.0, 1.5, 3.0, 6.0, 10.0, 15.0, 25.0, 60.0])
        g = torch.Generator().manual_seed(it * 7 + 1)
        x = torch.randn(N, T, C, generator=g) * scale
        if rng.random() < 0.2:
            x[:, ::3, rng.randrange(C)] = -float("inf")
        lp = torch.log_softmax(x, -1).clamp_min(-1e30)
        tg = torch.randint(1, C, (N, max(S, 1)), generator=g, dtype=torch.int32)
        if rng.random() < 0.5 and S > 2:
            tg[:, 1::2] = tg[:, 0::2][:, : tg[:, 1::2].shape[1]]
        il = torch.tensor([rng.randint(1, T) for _ in range(N)], dtype=torch.int32)
        tl = torch.tensor([rng.randint(0, S) for _ in range(N)], dtype=torch.int32)
        il[0], tl[0] = T, S
        for n in range(N):
            tg[n, tl[n]:] = 0
        loss_ref, grad_ref = O.ctc_loss_torch(lp.double(), tg, il, tl, dtype=torch.float64)
        gref = torch.nan_to_num(grad_ref, nan=0.0)
        gmax = grad_ref.abs().max().item() + 1e-12
        for mode in ("0", "1", "2"):
            os.environ["W2L_CTC_LINEAR"] = mode
            loss, _nll, grad = F.ctc_loss_raw(lp, tg, il, tl)
            ok = (abs(loss.item() - loss_ref.item()) <= 1e-4 * max(1.0, abs(loss_ref.item()))
                  and (grad.double() - gref).abs().max().item() <= 2e-3 * gmax and not torch.isnan(grad).any())
            if not ok:
                fails[mode].add(it)
                print("mode %s case %d: N=%d T=%d S=%d C=%d scale=%g  loss %.6g (ref %.6g)" % (mode, it, N, T, S, C, scale, loss.item(), loss_ref.item()))
    print("failed: log space %d, linear+redo %d, linear alone %d of %d" % (len(fails["0"]), len(fails["1"]), len(fails["2"]), cases))
    regress = fails["1"] - fails["0"]
    print("linear+redo fails where log space passes:", sorted(regress) or "never")
    return 1 if regress else 0


if __name__ == "__main__":
    sys.exit(main())
