"""The linear-domain schedule of the CTC loss + gradient (csrc/ctc.cu: ``ctc_lattice_lin_kernel`` / ``ctc_grad_lin_kernel``, opt-in
with W2L_CTC_LINEAR=1) executed on the host from its own source through the library's C wrapper ``w2l_ctc_loss`` (tests/_emu_cabi.py)
and held to the bars of the log-space kernels: CTC loss within 1e-4 relative, gradient within 2e-3 of its largest element, exact zeros
past the input length and for infeasible utterances, fixtures frozen from the reference's nn.CTCLoss.

The mantissa + per-thread exponent representation is stressed where it differs from log space: long utterances with short targets
(the lattice's mass spans > 2^130 between its ends), sharply peaked rows, rows with exact zeros, lattices over several warps."""
import os

import numpy as np
import pytest
import torch

import _emu_cabi
import _kernel_emu as KE
from oracle import w2l_oracle as O

pytestmark = pytest.mark.skipif(not KE.available(), reason="needs g++ and the CUDA headers")


@pytest.fixture
def F(monkeypatch):
    """W2L_CTC_LINEAR=2: the linear-domain kernels ALONE (no second pass in log space for flagged utterances), so that the bars below are
    met by the new kernels themselves"""
    monkeypatch.setenv("W2L_CTC_LINEAR", "2")
    return _emu_cabi.install(monkeypatch)


def check(F, lp, tg, il, tl, from_logits=False, loss_tol=1e-4, grad_tol=2e-3):
    lp = torch.as_tensor(lp, dtype=torch.float32)
    tg, il, tl = (torch.as_tensor(v, dtype=torch.int32) for v in (tg, il, tl))
    ref_in = torch.log_softmax(lp.double(), -1) if from_logits else lp.double()
    loss_ref, grad_ref = O.ctc_loss_torch(ref_in, tg, il, tl, dtype=torch.float64)
    if from_logits:
        x = lp.double().clone().requires_grad_(True)
        l = torch.nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)(torch.log_softmax(x, -1).transpose(0, 1), tg, il, tl)
        (grad_ref,) = torch.autograd.grad(l, x)
    launches = F._lib.load().w2l_launch_count()
    loss, nll, grad = F.ctc_loss_raw(lp, tg, il, tl, from_logits=from_logits)
    # prep, lattice (alpha | beta), gradient, finish (+ the two predicated log-space launches in mode 1)
    assert F._lib.load().w2l_launch_count() - launches == (4 if os.environ["W2L_CTC_LINEAR"] == "2" else 6)
    assert abs(loss.item() - loss_ref.item()) <= loss_tol * max(1.0, abs(loss_ref.item())), (loss.item(), loss_ref.item())
    assert not torch.isnan(grad).any()                                      # every gradient element written
    gmax = grad_ref.abs().max().item() + 1e-12
    err = (grad.double() - grad_ref).abs().max().item()
    assert err <= grad_tol * gmax, (err, gmax)
    for n in range(lp.shape[0]):
        assert (grad[n, int(il[n]):] == 0).all()
    return nll, grad


def test_linear_mode_is_what_runs(F, monkeypatch):
    """the workspace grows by the two exponent planes, and only when the switch is set"""
    lib = F._lib.load()
    with_lin = lib.w2l_ctc_loss_workspace_bytes(4, 100, 20)
    monkeypatch.setenv("W2L_CTC_LINEAR", "0")
    without = lib.w2l_ctc_loss_workspace_bytes(4, 100, 20)
    assert with_lin - without == 2 * ((4 * 100 * 32 * 4 + 255) // 256 * 256)   # S=20: 41 states, R=2 -> one warp


@pytest.mark.parametrize("name", ["ragged", "infeasible", "single", "long"])
def test_reference_fixtures(F, golden, name):
    g = golden("ctc")
    nll, grad = check(F, g[name + ":lp"], g[name + ":tg"], g[name + ":il"], g[name + ":tl"])
    np.testing.assert_allclose(nll.numpy(), g[name + ":nll"], rtol=1e-4, atol=1e-5)
    gref = g[name + ":grad"]
    assert np.abs(grad.numpy() - gref).max() <= 2e-3 * np.abs(gref).max() + 1e-7


@pytest.mark.parametrize("N,T,S,C", [(3, 70, 12, 29), (2, 150, 100, 29), (2, 90, 40, 5), (2, 120, 70, 29), (2, 40, 1, 29), (1, 300, 140, 29)])
@pytest.mark.parametrize("from_logits", [False, True])
def test_random_ragged(F, N, T, S, C, from_logits):
    """the GPU suite's random case at emulation sizes: repeated labels, ragged input and target lengths, one to five warps"""
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
    check(F, lp, tg, il, tl, from_logits=from_logits)


def test_long_utterance_short_target(F):
    """T = 1500 frames over 30 labels at near-uniform probabilities: alpha piles up at the end of the lattice and beta at its start
    while the occupancy sits in the middle, 2^100 and more below either maximum -- one exponent per CTA would flush it"""
    g = torch.Generator().manual_seed(5)
    N, T, S, C = 2, 1500, 30, 29
    lp = torch.log_softmax(torch.randn(N, T, C, generator=g) * 0.3, -1)
    tg = torch.randint(1, C, (N, S), generator=g, dtype=torch.int32)
    il = torch.tensor([T, T - 123], dtype=torch.int32)
    tl = torch.tensor([S, S - 7], dtype=torch.int32)
    tg[1, S - 7:] = 0
    check(F, lp, tg, il, tl)


def _peaked(scale):
    g = torch.Generator().manual_seed(int(scale))
    N, T, S, C = 3, 80, 15, 29
    lp = torch.log_softmax(torch.randn(N, T, C, generator=g) * scale, -1)
    tg = torch.randint(1, C, (N, S), generator=g, dtype=torch.int32)
    il = torch.tensor([80, 61, 80], dtype=torch.int32)
    tl = torch.tensor([15, 15, 4], dtype=torch.int32)
    tg[2, 4:] = 0
    return lp, tg, il, tl


def test_peaked_rows_within_range(F):
    """sharply peaked frames (probabilities down to e^-40 next to ~1): single-frame factors far outside what a normalise-every-few-frames
    scheme could hold, still inside the 2^126 a lane's mantissas span"""
    check(F, *_peaked(8.0))


def test_peaked_rows_beyond_range_are_redone_in_log_space(F, monkeypatch):
    """emissions of one lane's classes more than 2^126 apart (log-probs of -100 and below beside ~0): the linear kernels alone lose live
    states, notice it (CtcMeta::redo_*), and with W2L_CTC_LINEAR=1 those utterances come from the log-space kernels"""
    lp, tg, il, tl = _peaked(30.0)
    with pytest.raises(AssertionError):
        check(F, lp, tg, il, tl)
    monkeypatch.setenv("W2L_CTC_LINEAR", "1")
    check(F, lp, tg, il, tl)


def test_exact_zero_probabilities_and_infeasible(F):
    """-inf log-probs (probability exactly 0) on some classes, a target longer than its input (infeasible: loss 0 and gradient 0 with
    zero_infinity), an empty target and a one-frame utterance"""
    g = torch.Generator().manual_seed(2)
    N, T, S, C = 5, 30, 8, 6
    x = torch.randn(N, T, C, generator=g)
    x[:, ::4, 3] = -float("inf")
    lp = torch.log_softmax(x, -1)
    tg = torch.randint(1, C, (N, S), generator=g, dtype=torch.int32)
    tg[0] = torch.tensor([1, 2, 1, 2, 4, 5, 4, 5])                       # never needs class 3
    il = torch.tensor([30, 5, 30, 1, 17], dtype=torch.int32)
    tl = torch.tensor([8, 8, 0, 1, 3], dtype=torch.int32)                # utterance 1: 8 labels in 5 frames
    for n in range(N):
        tg[n, tl[n]:] = 0
    lp = lp.clamp_min(-1e30)                                             # torch's CPU ctc_loss turns -inf inputs into NaN gradients
    nll, grad = check(F, lp, tg, il, tl)
    assert nll[1].item() == 0.0 and (grad[1] == 0).all()


def test_agrees_with_log_space_kernels(F, monkeypatch):
    """same inputs through both schedules of the same library build"""
    g = torch.Generator().manual_seed(3)
    N, T, S, C = 2, 200, 45, 29
    lp = torch.log_softmax(torch.randn(N, T, C, generator=g), -1)
    tg = torch.randint(1, C, (N, S), generator=g, dtype=torch.int32)
    il, tl = torch.tensor([200, 150], dtype=torch.int32), torch.tensor([45, 31], dtype=torch.int32)
    tg[1, 31:] = 0
    loss_lin, nll_lin, grad_lin = F.ctc_loss_raw(lp, tg, il, tl)
    monkeypatch.setenv("W2L_CTC_LINEAR", "0")
    loss_log, nll_log, grad_log = F.ctc_loss_raw(lp, tg, il, tl)
    np.testing.assert_allclose(nll_lin.numpy(), nll_log.numpy(), rtol=2e-6)
    assert (grad_lin - grad_log).abs().max().item() <= 1e-4 * grad_log.abs().max().item()
