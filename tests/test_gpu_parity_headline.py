"""Parity at the configurations the headline numbers are quoted on, with CLOSED tolerances (every bound below is a constant):

* Wav2Letter ``mid_layers=20`` (configuration/model/wav2letter.yaml: widths 256..1024, k up to 29, dilation 2) -- one real training
  step (train-mode BatchNorm, CTC, backward), checked block by block against the oracle (tests/_layerwise.py explains why teacher
  forcing is the only way to a closed gradient tolerance at this depth), plus the end-to-end quantities that ARE well conditioned
  (loss, log-probs, head gradients) against the fp32 oracle run over the whole stack;
* the Jasper 10x5 block shapes (256 k11 / 384 k13 / 512 k17 / 640 k21 / 768 k25, repeat 5, residual; 896 k29 d2; 1024 k1) the same way,
  per conv+BN group, with ragged lengths so that the masks do real work;
* backward-data and backward-weights of the largest layer at FULL size (B=64, T'=750, 896->896, k=29, d=2: stream-K, 4-D tensor
  maps, vector atomics) against torch fp32 on slices of utterances / taps;
* the CTC gradient at N=512, T=3000, S=600 against torch fp64 on 32 utterances (and every utterance's loss).

Fixed bounds.  Teacher-forced blocks, device vs the oracle with the device's bf16 storage points emulated:
output 5e-3, input gradient 2e-2, parameter gradients 2e-2; vs the plain fp32 oracle: output 1e-2, gradients 1.5e-1 (a bf16-stored
pre-activation flips the clamp/ReLU gate of the ~0.1% of elements that sit within one bf16 ulp of the gate, which moves the gated
gradient by sqrt(0.001) ~ 3-5% in relative L2 -- measured 3-6e-2 for every layer at the real widths, up to 9e-2 on the toy fixtures)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as TF

import _layerwise as L
from oracle import w2l_oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1800)]

from test_gpu_models import TOL_EMU, TOL_REF  # noqa: E402  (one set of fixed bounds for every model-level parity test)
JASPER_10X5_SHAPES = [(256, 11, 2, 1, False, 1), (256, 11, 1, 1, True, 5), (384, 13, 1, 1, True, 5), (512, 17, 1, 1, True, 5),
                      (640, 21, 1, 1, True, 5), (768, 25, 1, 1, True, 5), (896, 29, 1, 2, False, 1), (1024, 1, 1, 1, False, 1)]


@pytest.fixture(scope="module")
def F():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from wav2letter_pytorch_b200 import functional
    return functional


def _ctc_checks(out, out_lens, tg, tl, loss):
    """the CTC kernel on the scores the stack produced: loss <= 1e-4 relative, gradient <= 2e-3 * max|g| vs torch fp64"""
    lp = out.detach().double().cpu().requires_grad_(True)
    ref = torch.nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)(lp.transpose(0, 1), tg.cpu(), out_lens.cpu(), tl.cpu())
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item()), (loss.item(), ref.item())
    err = (out.grad.double().cpu() - lp.grad).abs().max().item()
    assert err <= 2e-3 * lp.grad.abs().max().item(), err


def test_w2l20_train_step_parity(F):
    from wav2letter_pytorch_b200 import config
    from wav2letter_pytorch_b200.wav2letter import Wav2Letter
    cfg = config.compose(overrides=["model.mid_layers=20"]).model
    for l in cfg.layers:
        l["dropout"] = 0.0
    torch.manual_seed(0)
    model = Wav2Letter(cfg).cuda().train()
    sd0 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    x, il, tg, tl = O.synthetic_batch(4, 3, seed=5, ragged=True)
    hs, out, out_lens, loss = L.w2l_run_blocks(model, x.cuda(), il.cuda(), tg.cuda(), tl.cuda())
    torch.cuda.synchronize()
    assert out.shape == (4, 150, 29) and torch.isfinite(out).all()
    # ---- every block on its own, at its real width / kernel / dilation
    table = L.w2l_layerwise_table(model, hs, out)
    print("\n" + L.format_table(table))
    assert len(table) == 21
    bad = L.check_table(table, TOL_EMU, TOL_REF)
    assert not bad, bad
    _ctc_checks(out, out_lens, tg, tl, loss)
    # ---- end to end against the fp32 oracle over the whole stack (wav2letter.py:84-92 + base_asr_models.py:81): the quantities
    # that are well conditioned at depth 20
    specs = O.w2l_layer_specs(20, dropout=0.0)
    params = {k: v.clone().requires_grad_(True) for k, v in sd0.items() if v.is_floating_point() and "running" not in k}
    sd = dict(sd0)
    sd.update(params)
    lp_ref, ol_ref = O.w2l_forward(x, il, sd, specs, training=True, update_running=False)
    loss_ref = torch.nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)(lp_ref.transpose(0, 1), tg, ol_ref, tl)
    loss_ref.backward()
    assert np.array_equal(out_lens.cpu().numpy(), ol_ref.numpy())
    e_lp = L.rel_l2(out, lp_ref.detach())
    print("end to end vs fp32 oracle: loss %.5f vs %.5f, log-probs rel-L2 %.3e" % (loss.item(), loss_ref.item(), e_lp))
    assert abs(loss.item() - loss_ref.item()) <= 1e-2 * abs(loss_ref.item())
    assert e_lp <= 5e-2
    named = dict(model.named_parameters())
    for name in ("conv1ds.conv1d_20.conv1.weight", "conv1ds.conv1d_20.conv1.bias"):
        e = L.rel_l2(named[name].grad, params[name].grad)
        print("  %-40s %.3e" % (name, e))
        assert e <= 1e-1, (name, e)
    # reported, not bounded (chaotic amplification ~1.2x per layer of a freshly initialised BatchNorm stack: the host emulation of
    # the same storage precision is 0.5-0.85 away from fp32 in these rows, see tests/_layerwise.py)
    for i in (0, 5, 10, 15, 19):
        name = "conv1ds.conv1d_%d.conv1.weight" % i
        print("  %-40s %.3e (end to end, informational)" % (name, L.rel_l2(named[name].grad, params[name].grad)))
    # running statistics moved the way nn.BatchNorm1d(momentum=0.9) moves them (wav2letter.py:37)
    rm = model.conv1ds.conv1d_0.batch_norm.running_mean
    assert int(model.conv1ds.conv1d_0.batch_norm.num_batches_tracked) == 1 and float(rm.abs().max()) > 0


def test_jasper10x5_block_shapes_parity(F):
    from wav2letter_pytorch_b200 import config
    from wav2letter_pytorch_b200.jasper import Jasper
    blocks = [dict(layer_size=w, kernel_size=k, stride=s, dilation=d, residual=res, repeat=rep, separable=False, dropout=0.0)
              for (w, k, s, d, res, rep) in JASPER_10X5_SHAPES]
    cfg = config.compose(overrides=["model=jasper", "model.mid_layers=%d" % len(blocks)]).model
    cfg["jasper_blocks"] = config.to_attr(blocks)
    torch.manual_seed(0)
    model = Jasper(cfg).cuda().train()
    specs = O.jasper_block_specs(blocks)
    x, il, tg, tl = O.synthetic_batch(2, 3, seed=6, ragged=True)
    hs, taps, rows, out, out_lens, loss = L.jasper_run_blocks(model, x.cuda(), il.cuda(), tg.cuda(), tl.cuda())
    torch.cuda.synchronize()
    table = L.jasper_layerwise_table(model, specs, hs, taps, rows, out)
    print("\n" + L.format_table(table))
    assert len(table) == sum(s[5] for s in JASPER_10X5_SHAPES) + 1
    bad = L.check_table(table, TOL_EMU, TOL_REF)
    assert not bad, bad
    _ctc_checks(out, out_lens, tg, tl, loss)
    # lengths: the float chain of jasper.py:107-119 through every masked conv
    want = il.clone().long()
    want = torch.div(want + 2 * 5 - 10 - 1, 2, rounding_mode="floor") + 1
    assert np.array_equal(out_lens.cpu().numpy(), want.numpy())


def test_conv_full_size_backward_vs_torch(F):
    """B=64, T'=750, 896->896, k=29, d=2 -- the largest layer of the headline step: backward-data (tap-reversed K-major shadow) on
    three whole utterances and backward-weights (stream-K over the batch, red.global.add.v4) on three whole taps against torch fp32
    on the same bf16 operands.  Bounds: dx is stored as bf16 (one rounding: <= 4e-3 rel-L2), dW is fp32 (<= 1e-4 at 48 000 terms)."""
    g = torch.Generator().manual_seed(7)
    B, T, C, k, d = 64, 750, 896, 29, 2
    halo = (k - 1) * d
    Tp = T + halo
    xp = torch.randn(B, Tp, C, generator=g).to(torch.bfloat16)              # halo-carrying input, as inside the Wav2Letter stack
    w = (torch.randn(C, C, k, generator=g) / (C * k) ** 0.5).to(torch.bfloat16)
    dy = torch.randn(B, T, C, generator=g).to(torch.bfloat16)
    dz = torch.zeros(B, Tp, C, dtype=torch.bfloat16, device="cuda")           # input row pitch, zero tails (the flat dgrad layout)
    dz[:, :T] = dy.cuda()
    wt = torch.empty(k, C, C, dtype=torch.bfloat16, device="cuda")
    F.pack_wt(w.float().permute(2, 0, 1).contiguous().cuda(), wt, C, C)
    dx = torch.empty(B, Tp, C, dtype=torch.bfloat16, device="cuda")
    F.conv1d_dgrad_wt(dz, wt, F.make_desc(1, B * Tp, C, C, C, k, d, B * Tp, 0, B * Tp, 0, C), dx)
    xc = xp.cuda()
    dw = torch.empty(k, C, C, dtype=torch.float32, device="cuda")
    F.conv1d_wgrad(dz, xc, F.make_desc(B, T, C, C, C, k, d, Tp, 0, Tp, 0, C), dw)
    torch.cuda.synchronize()
    wf = w.float()
    for b in (0, 37, 63):
        ref = TF.conv_transpose1d(dy[b:b + 1].float().transpose(1, 2), wf, dilation=d).transpose(1, 2)       # [1, Tp, C]
        e = L.rel_l2(dx[b:b + 1], ref)
        assert e < 4e-3, (b, e)
    xf, dyf = xp.float(), dy.float()
    for j in (0, 14, 28):
        ref = torch.einsum("bto,bti->oi", dyf, xf[:, j * d: j * d + T])                                      # dW[j][co, ci]
        e = L.rel_l2(dw[j], ref)
        assert e < 1e-4, (j, e)
    # the per-utterance layout (zero padding from TMA out-of-bounds fill, Jasper) at full size: forward + backward-data on a slice
    pad = halo // 2
    xj = xp[:, :T].contiguous().cuda()
    y = torch.empty(B, T, C, dtype=torch.float32, device="cuda")
    F.conv1d_fwd(xj, w.permute(2, 0, 1).contiguous().cuda(), F.make_desc(B, T, C, C, C, k, d, T, -pad, T, 0, C, F.DT_F32, F.ACT_NONE), y)
    dxj = torch.empty(B, T, C, dtype=torch.bfloat16, device="cuda")
    F.conv1d_dgrad_wt(dy.cuda(), wt, F.make_desc(B, T, C, C, C, k, d, T, -pad, T, 0, C), dxj)
    torch.cuda.synchronize()
    for b in (11, 63):
        xb = xp[b:b + 1, :T].float().transpose(1, 2).clone().requires_grad_(True)
        yb = TF.conv1d(xb, wf, padding=pad, dilation=d)
        yb.backward(dyf[b:b + 1].transpose(1, 2))
        assert L.rel_l2(y[b:b + 1], yb.detach().transpose(1, 2)) < 1e-4
        assert L.rel_l2(dxj[b:b + 1], xb.grad.transpose(1, 2)) < 4e-3


def test_ctc_full_size_gradient_vs_torch(F):
    """N=512, T=3000, S=600, C=29 (the top corner of BASELINE config 5): every utterance's loss and the gradient of 32 utterances
    spread over the batch against torch's CPU CTC in fp64; ragged lengths, repeated labels."""
    g = torch.Generator().manual_seed(2)
    N, T, S, C = 512, 3000, 600, 29
    lp = torch.log_softmax(torch.randn(N, T, C, generator=g) * 1.5, -1)
    tg = torch.randint(1, C, (N, S), generator=g, dtype=torch.int32)
    tg[:, 1::5] = tg[:, 0::5][:, : tg[:, 1::5].shape[1]]
    il = torch.randint(2 * T // 3, T + 1, (N,), generator=g, dtype=torch.int32)
    tl = torch.randint(S // 2, S + 1, (N,), generator=g, dtype=torch.int32)
    il[0], tl[0] = T, S
    for n in range(N):
        tg[n, tl[n]:] = 0
    loss, nll, grad = F.ctc_loss_raw(lp.cuda(), tg.cuda(), il.cuda(), tl.cuda())
    torch.cuda.synchronize()
    ref_nll = torch.cat([TF.ctc_loss(lp[i:i + 64].double().transpose(0, 1), tg[i:i + 64], il[i:i + 64], tl[i:i + 64], reduction="none",
                                     zero_infinity=True) for i in range(0, N, 64)])        # 64 at a time: torch keeps alpha [N,T,2S+1] in fp64
    np.testing.assert_allclose(nll.cpu().numpy(), ref_nll.numpy(), rtol=1e-4)
    want_loss = (ref_nll / tl.clamp(min=1).double()).mean()
    assert abs(loss.item() - want_loss.item()) <= 1e-4 * abs(want_loss.item())
    pick = list(range(0, N, 16))
    sub = lp[pick].double().requires_grad_(True)
    ref = TF.ctc_loss(sub.transpose(0, 1), tg[pick], il[pick], tl[pick], reduction="none", zero_infinity=True)
    (ref / (N * tl[pick].clamp(min=1).double())).sum().backward()                      # the 'mean' scaling of the full batch
    got = grad[pick].double().cpu()
    assert (got - sub.grad).abs().max().item() <= 2e-3 * sub.grad.abs().max().item()
    for i, n in enumerate(pick):
        assert (got[i, int(il[n]):] == 0).all()
    assert grad.sum(-1).abs().max().item() < 1e-6
