"""GPU tests of the linear-domain CTC schedule (csrc/ctc.cu ``ctc_lattice_lin_kernel`` / ``ctc_grad_lin_kernel``, W2L_CTC_LINEAR).

EXPERIMENTAL and opt-in, like the code path itself: written after round 1's GPU budget was spent, so these have only run on the
emulated GPU (`pytest -m gpu --emulate-gpu`, and tests/test_kernel_emu_ctc_linear.py in the CPU suite).  On a real GPU they run only
with W2L_TEST_EXPERIMENTAL=1 (tools/final_run.sh sets it), so that the default `-m gpu` run exercises exactly the measured default path.
The file sorts last: a fault in an experimental kernel must not take the suite's other tests with it."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


@pytest.fixture
def F(request):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    if not (os.environ.get("W2L_TEST_EXPERIMENTAL") or request.config.getoption("--emulate-gpu")):
        pytest.skip("experimental path: set W2L_TEST_EXPERIMENTAL=1 to run it on the GPU")
    from wav2letter_pytorch_b200 import functional
    return functional


@pytest.mark.parametrize("mode", ["1", "2"])
@pytest.mark.parametrize("name", ["ragged", "infeasible", "long", "single"])
def test_ctc_linear_golden(F, golden, monkeypatch, name, mode):
    """fixtures frozen from the reference's nn.CTCLoss; mode 2 = the linear-domain kernels alone, mode 1 = with the log-space redo"""
    from test_gpu_kernels import _check_ctc
    monkeypatch.setenv("W2L_CTC_LINEAR", mode)
    g = golden("ctc")
    nll = _check_ctc(F, g[name + ":lp"], g[name + ":tg"], g[name + ":il"], g[name + ":tl"])
    torch.testing.assert_close(nll.cpu(), torch.from_numpy(g[name + ":nll"]), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("N,T,S,C", [(8, 500, 150, 29), (64, 750, 225, 29), (4, 200, 50, 29), (3, 1000, 300, 29), (2, 1300, 600, 29),
                                     (5, 64, 10, 5), (2, 40, 1, 29), (4, 3000, 50, 29)])
@pytest.mark.parametrize("from_logits", [False, True])
def test_ctc_linear_random(F, monkeypatch, N, T, S, C, from_logits):
    from test_gpu_kernels import _check_ctc
    monkeypatch.setenv("W2L_CTC_LINEAR", "2")
    g = torch.Generator().manual_seed(T + S)
    x = torch.randn(N, T, C, generator=g) * 1.5
    lp = x if from_logits else torch.log_softmax(x, -1)
    tg = torch.randint(1, C, (N, S), generator=g, dtype=torch.int32)
    tg[:, 1::3] = tg[:, 0::3][:, : tg[:, 1::3].shape[1]]
    il = torch.randint(max(1, T // 2), T + 1, (N,), generator=g, dtype=torch.int32)
    tl = torch.randint(0, S + 1, (N,), generator=g, dtype=torch.int32)
    il[0], tl[0] = T, S
    for n in range(N):
        tg[n, tl[n]:] = 0
    _check_ctc(F, lp, tg, il, tl, from_logits=from_logits)


@pytest.mark.parametrize("scale", [8.0, 30.0])
def test_ctc_linear_peaked_rows(F, monkeypatch, scale):
    """rows peaked beyond the mantissa range are noticed on the device and redone in log space (mode 1)"""
    from test_gpu_kernels import _check_ctc
    monkeypatch.setenv("W2L_CTC_LINEAR", "1")
    g = torch.Generator().manual_seed(int(scale))
    N, T, S, C = 6, 300, 40, 29
    lp = torch.log_softmax(torch.randn(N, T, C, generator=g) * scale, -1)
    tg = torch.randint(1, C, (N, S), generator=g, dtype=torch.int32)
    il = torch.tensor([300, 261, 300, 150, 300, 41], dtype=torch.int32)
    tl = torch.tensor([40, 40, 4, 33, 0, 40], dtype=torch.int32)
    for n in range(N):
        tg[n, tl[n]:] = 0
    _check_ctc(F, lp, tg, il, tl)


def test_ctc_linear_matches_log_space_at_training_shape(F, monkeypatch):
    """B=64 x 15 s (T'=750, 225 labels): same loss and gradient from both schedules of the same build"""
    g = torch.Generator().manual_seed(0)
    N, T, S, C = 64, 750, 225, 29
    lp = torch.log_softmax(torch.randn(N, T, C, generator=g), -1).cuda()
    tg = torch.randint(1, C, (N, S), generator=g, dtype=torch.int32).cuda()
    il = torch.full((N,), T, dtype=torch.int32).cuda()
    tl = torch.full((N,), S, dtype=torch.int32).cuda()
    monkeypatch.setenv("W2L_CTC_LINEAR", "2")
    loss_lin, nll_lin, grad_lin = F.ctc_loss_raw(lp, tg, il, tl)
    monkeypatch.setenv("W2L_CTC_LINEAR", "0")
    loss_log, nll_log, grad_log = F.ctc_loss_raw(lp, tg, il, tl)
    torch.testing.assert_close(nll_lin, nll_log, rtol=2e-6, atol=0)
    assert (grad_lin - grad_log).abs().max().item() <= 1e-4 * grad_log.abs().max().item()
