"""The CTC loss + gradient kernels and the greedy decode kernels (kernels 2 and 3), executed on the host from
their own source text (tests/_kernel_emu.py) in the launch sequences of ``w2l_ctc_loss`` / ``w2l_greedy_decode`` and held to
the oracle: transcripts bit-exact, CTC loss within 1e-4 relative, gradient within 2e-3 of its largest element -- the bars of
the `-m gpu` tests in tests/test_gpu_kernels.py, on cases small enough for a fiber-per-thread emulation.

Both CTC schedules run: the parallel one (alpha and beta lattices in separate CTAs with the tagged-slot wavefront across
warps, then the per-frame gradient kernel) and the serial one (alpha, then beta fused with the gradient).  The inline-PTX
helpers of ctc.cu are replaced by host functions (tests/_emu_backend.py CTC_PTX); everything else is the library's code,
including ``make_plan``."""
import numpy as np
import pytest
import torch

import _emu_backend as E
import _kernel_emu as KE
from oracle import w2l_oracle as O

pytestmark = pytest.mark.skipif(not KE.available(), reason="needs g++ and the CUDA headers")


@pytest.fixture(scope="module")
def ctc():
    return E.ctc()


def emu_ctc_loss(ctc, x, targets, il, tl, from_logits=False, serial=False):
    """tests/_emu_backend.py ctc_loss_raw: csrc/ctc.cu w2l_ctc_loss + launch_ctc, launch for launch"""
    return E.ctc_loss_raw(x, targets, il, tl, from_logits=from_logits, serial=serial, return_plan=True)


def check_ctc(ctc, lp, tg, il, tl, from_logits=False, serial=False, loss_tol=1e-4, grad_tol=2e-3):
    lp = torch.as_tensor(lp, dtype=torch.float32)
    tg, il, tl = (torch.as_tensor(v, dtype=torch.int32) for v in (tg, il, tl))
    ref_in = torch.log_softmax(lp.double(), -1) if from_logits else lp.double()
    loss_ref, grad_ref = O.ctc_loss_torch(ref_in, tg, il, tl, dtype=torch.float64)
    if from_logits:
        x = lp.double().clone().requires_grad_(True)
        l = torch.nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)(torch.log_softmax(x, -1).transpose(0, 1), tg, il, tl)
        (grad_ref,) = torch.autograd.grad(l, x)
    loss, nll, grad, plan = emu_ctc_loss(ctc, lp, tg, il, tl, from_logits=from_logits, serial=serial)
    assert abs(loss.item() - loss_ref.item()) <= loss_tol * max(1.0, abs(loss_ref.item())), (loss.item(), loss_ref.item())
    assert not torch.isnan(grad).any()                                    # every gradient element written
    gmax = grad_ref.abs().max().item() + 1e-12
    err = (grad.double() - grad_ref).abs().max().item()
    assert err <= grad_tol * gmax, (err, gmax)
    for n in range(lp.shape[0]):
        assert (grad[n, int(il[n]):] == 0).all()
    return nll, plan


@pytest.mark.parametrize("name", ["ragged", "infeasible", "single"])
@pytest.mark.parametrize("serial", [False, True])
def test_ctc_source_on_reference_fixtures(ctc, golden, name, serial):
    """tests/golden/ctc.npz: losses and gradients frozen from the reference's nn.CTCLoss(blank=0, 'mean', zero_infinity=True)"""
    g = golden("ctc")
    nll, _ = check_ctc(ctc, g[name + ":lp"], g[name + ":tg"], g[name + ":il"], g[name + ":tl"], serial=serial)
    np.testing.assert_allclose(nll.numpy(), g[name + ":nll"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("N,T,S,C,want_threads", [(3, 70, 12, 29, (2, 32)), (2, 150, 100, 29, (2, 128)), (2, 90, 40, 5, (2, 64)),
                                                  (2, 40, 1, 29, (2, 32)), (1, 170, 150, 29, (4, 96)), (1, 340, 300, 29, (8, 96))])
@pytest.mark.parametrize("from_logits", [False, True])
def test_ctc_source_random(ctc, N, T, S, C, want_threads, from_logits):
    """ragged lengths, repeated labels, an empty target; lattices of one to four warps (the wavefront across warps) with 2, 4
    and 8 states per lane"""
    g = torch.Generator().manual_seed(T + S)
    x = torch.randn(N, T, C, generator=g) * 1.5
    lp = x if from_logits else torch.log_softmax(x, -1)
    tg = torch.randint(1, C, (N, S), generator=g, dtype=torch.int32)
    tg[:, 1::3] = tg[:, 0::3][:, : tg[:, 1::3].shape[1]]
    il = torch.randint(max(1, T // 2), T + 1, (N,), generator=g, dtype=torch.int32)
    tl = torch.randint(0, S + 1, (N,), generator=g, dtype=torch.int32)
    il[0], tl[0] = T, S
    if N > 1:
        tl[1] = 0
    for n in range(N):
        tg[n, tl[n]:] = 0
    _, plan = check_ctc(ctc, lp, tg, il, tl, from_logits=from_logits)
    assert (plan["R"], plan["threads"]) == want_threads and plan["parallel"]
    if not from_logits:
        check_ctc(ctc, lp, tg, il, tl, serial=True)


@pytest.mark.parametrize("C", [33, 100, 128])
def test_ctc_source_wide_alphabet_small_lattice(ctc, C):
    """more classes than the CTA has threads (one warp for a short target): the serial schedule used to emit only the first
    blockDim.x gradient columns and to stage only blockDim.x * 4 columns of each frame (found by running it here; C = 29, the
    shipped label sets, was never affected)"""
    g = torch.Generator().manual_seed(C)
    N, T, S = 2, 40, 5
    lp = torch.log_softmax(torch.randn(N, T, C, generator=g) * 1.5, -1)
    tg = torch.randint(1, C, (N, S), generator=g, dtype=torch.int32)
    il, tl = torch.tensor([40, 31], dtype=torch.int32), torch.tensor([5, 3], dtype=torch.int32)
    tg[1, 3:] = 0
    for serial in (False, True):
        _, plan = check_ctc(ctc, lp, tg, il, tl, serial=serial)
        assert plan["threads"] == 32


def test_ctc_source_zero_length_input_and_renorm(ctc):
    """an utterance with input length 0 (grad exactly 0, feasible only for an empty target) and peaked frames long enough to
    cross several re-centring blocks"""
    g = torch.Generator().manual_seed(2)
    N, T, S, C = 3, 130, 9, 7
    lp = torch.log_softmax(torch.randn(N, T, C, generator=g) * 6, -1)
    tg = torch.randint(1, C, (N, S), generator=g, dtype=torch.int32)
    il = torch.tensor([130, 0, 97], dtype=torch.int32)
    tl = torch.tensor([9, 0, 4], dtype=torch.int32)
    tg[1] = 0
    tg[2, 4:] = 0
    for serial in (False, True):
        loss, nll, grad, _ = emu_ctc_loss(ctc, lp, tg, il, tl, serial=serial)
        assert nll[1].item() == 0.0 and (grad[1] == 0).all()
        ref = torch.nn.functional.ctc_loss(lp[[0, 2]].double().transpose(0, 1), tg[[0, 2]], il[[0, 2]], tl[[0, 2]], reduction="none")
        np.testing.assert_allclose(nll[[0, 2]].numpy(), ref.numpy(), rtol=1e-4)


# ------------------------------------------------------------------------------------------------ greedy decode
@pytest.fixture(scope="module")
def dec():
    return E.decode()


def emu_greedy_decode(dec, scores, sizes=None, blank=0):
    """tests/_emu_backend.py greedy_decode: csrc/decode.cu w2l_greedy_decode, launch for launch"""
    return E.greedy_decode(scores, sizes, blank)


def check_decode(dec, lp, sizes):
    am, tok, off, cnt = emu_greedy_decode(dec, lp, sizes)
    want_am = O.greedy_argmax(lp.numpy())
    assert np.array_equal(am.numpy(), want_am)
    toks, offs = O.greedy_collapse(want_am, sizes)
    for n in range(lp.shape[0]):
        k = int(cnt[n])
        assert tok[n, :k].tolist() == toks[n] and off[n, :k].tolist() == offs[n]           # bit-exact transcripts and offsets
        assert (tok[n, k:] == -1).all() and (off[n, k:] == -1).all()                       # deterministic tail


def test_decode_source_reference_vectors(dec):
    """the reference's own decoder vectors: unit_tests/decoder_test.py:40-42 and decoder.py:305-311"""
    labels = ["_", "A", "B", " "]
    probs = torch.tensor([[[0.8, 0.2, 0, 0], [0.6, 0.4, 0, 0]]])
    _, tok, _, cnt = emu_greedy_decode(dec, probs)
    assert "".join(labels[i] for i in tok[0, :int(cnt[0])].tolist()) == ""
    probs = torch.tensor([[[0.1, 0.7, 0.1, 0.1], [0.7, 0.1, 0.1, 0.1], [0.1, 0.1, 0.7, 0.1], [0.1, 0.7, 0.1, 0.1]],
                          [[0.1, 0.1, 0.1, 0.7], [0.1, 0.1, 0.1, 0.7], [0.7, 0.1, 0.1, 0.1], [0.7, 0.1, 0.1, 0.1]]])
    am, tok, off, cnt = emu_greedy_decode(dec, probs)
    lab = ["_", "a", "b", " "]
    assert ["".join(lab[i] for i in tok[n, :int(cnt[n])].tolist()) for n in range(2)] == ["aba", " "]
    assert off[0, :3].tolist() == [0, 2, 3] and off[1, :1].tolist() == [0]
    # ties -> lowest index, NaN wins the argmax (torch.max semantics)
    probs = torch.tensor([[[0.5, 0.5, 0, 0], [0.1, float("nan"), 0.9, 0.0], [0.2, 0.3, 0.3, 0.2]]])
    am, _, _, _ = emu_greedy_decode(dec, probs)
    assert am[0].tolist() == [0, 1, 1]


@pytest.mark.parametrize("N,T,C", [(3, 50, 29), (2, 256, 29), (2, 257, 5), (2, 700, 29), (1, 1, 2), (4, 300, 31)])
def test_decode_source_random(dec, N, T, C):
    g = torch.Generator().manual_seed(N * 1000 + T)
    lp = torch.log_softmax(torch.randn(N, T, C, generator=g) * 2, -1)
    lp[:, :, 0] += 1.0
    lp = torch.round(lp * 4) / 4                                           # plenty of exact ties and repeats
    sizes = torch.randint(0, T + 1, (N,), generator=g).tolist()
    sizes[0] = T
    check_decode(dec, lp, sizes)
    check_decode(dec, lp, None)
    v = lp.transpose(0, 1).contiguous().transpose(0, 1)                    # [N,T,C] view of a [T,N,C] tensor: the strided path
    am, _, _, _ = emu_greedy_decode(dec, v)
    assert np.array_equal(am.numpy(), O.greedy_argmax(lp.numpy()))
    w = torch.zeros(N * T * C + 3)[3:].view(N, T, C).copy_(lp)             # base 12 bytes off 16: the peeled, vectorised copy
    check_decode(dec, w, sizes)
