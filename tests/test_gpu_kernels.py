"""Parity of each CUDA kernel (called through the C ABI) against the CPU oracle / plain torch fp32, on a B200."""
import numpy as np
import pytest
import torch
import torch.nn.functional as TF

from oracle import w2l_oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.fixture(scope="module")
def F():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from wav2letter_pytorch_b200 import functional
    return functional


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


# ------------------------------------------------------------------------------------------- decode
def _check_decode(F, probs, sizes, blank=0):
    p = torch.as_tensor(probs, dtype=torch.float32)
    sz = None if sizes is None else torch.as_tensor(sizes, dtype=torch.int32)
    am, tok, off, cnt = F.greedy_decode(p.cuda(), None if sz is None else sz.cuda(), blank)
    am_ref = O.greedy_argmax(p.numpy())
    tok_ref, off_ref = O.greedy_collapse(am_ref, None if sizes is None else list(sizes), blank)
    assert np.array_equal(am.cpu().numpy(), am_ref)
    cnt = cnt.cpu().numpy()
    for n in range(p.shape[0]):
        assert cnt[n] == len(tok_ref[n])
        assert tok[n, :cnt[n]].cpu().tolist() == tok_ref[n]
        assert off[n, :cnt[n]].cpu().tolist() == off_ref[n]
        assert (tok[n, cnt[n]:] == -1).all()


def test_decode_golden(F, golden):
    g = golden("decoder")
    for name in sorted({k.split(":")[0] for k in g.files}):
        sizes = g[name + ":sizes"]
        _check_decode(F, g[name + ":probs"], None if sizes[0] == -1 else sizes)


@pytest.mark.parametrize("N,T,C", [(1, 1, 4), (3, 255, 29), (2, 256, 29), (5, 257, 29), (4, 750, 29), (2, 1031, 32), (3, 300, 5)])
def test_decode_random(F, N, T, C):
    g = torch.Generator().manual_seed(N * 1000 + T)
    lp = torch.log_softmax(torch.randn(N, T, C, generator=g) * 2, -1)
    lp[:, :, 0] += 1.0
    lp = torch.round(lp * 4) / 4                       # plenty of exact ties and repeats
    sizes = torch.randint(0, T + 1, (N,), generator=g).tolist()
    sizes[0] = T
    _check_decode(F, lp, sizes)
    _check_decode(F, lp, None)
    # non-contiguous view [N,T,C] of a [T,N,C] tensor (stride_t != C path)
    v = lp.transpose(0, 1).contiguous().cuda().transpose(0, 1)
    am, tok, off, cnt = F.greedy_decode(v, None)
    assert np.array_equal(am.cpu().numpy(), O.greedy_argmax(lp.numpy()))


def test_decode_full_size_properties(F):
    """BASELINE config 5 upper end (N=512, T=3000): size-independent properties."""
    g = torch.Generator().manual_seed(0)
    N, T, C = 512, 3000, 29
    lp = torch.log_softmax(torch.randn(N, T, C, generator=g), -1).cuda()
    am, tok, off, cnt = F.greedy_decode(lp, None)
    assert torch.equal(am.long(), lp.argmax(-1))
    keep = (am != 0)
    keep[:, 1:] &= am[:, 1:] != am[:, :-1]
    assert torch.equal(cnt.long(), keep.sum(1))
    n = 17
    assert torch.equal(tok[n, :cnt[n]], am[n][keep[n]])
    assert torch.equal(off[n, :cnt[n]].long(), torch.nonzero(keep[n]).flatten())
    # idempotence: decoding a one-hot rendering of the collapsed path gives the same tokens
    o1 = off[n, :cnt[n]].long()
    assert (o1[1:] > o1[:-1]).all()


# ------------------------------------------------------------------------------------------- CTC
def _check_ctc(F, lp, tg, il, tl, from_logits=False, loss_tol=1e-4, grad_tol=2e-3):
    lp = torch.as_tensor(lp, dtype=torch.float32)
    tg, il, tl = (torch.as_tensor(v, dtype=torch.int32) for v in (tg, il, tl))
    ref_in = torch.log_softmax(lp.double(), -1) if from_logits else lp.double()
    loss_ref, grad_ref = O.ctc_loss_torch(ref_in, tg, il, tl, dtype=torch.float64)
    if from_logits:   # chain through log_softmax in fp64
        x = lp.double().clone().requires_grad_(True)
        l = torch.nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)(torch.log_softmax(x, -1).transpose(0, 1), tg, il, tl)
        (grad_ref,) = torch.autograd.grad(l, x)
    loss, nll, grad = F.ctc_loss_raw(lp.cuda(), tg.cuda(), il.cuda(), tl.cuda(), from_logits=from_logits)
    assert abs(loss.item() - loss_ref.item()) <= loss_tol * max(1.0, abs(loss_ref.item()))      # CTC loss within 1e-4 relative
    gmax = grad_ref.abs().max().item() + 1e-12
    err = (grad.cpu().double() - grad_ref).abs().max().item()
    assert err <= grad_tol * gmax, (err, gmax)
    for n in range(lp.shape[0]):
        assert (grad[n, int(il[n]):] == 0).all()
    return nll


@pytest.mark.parametrize("name", ["ragged", "infeasible", "long", "single"])
def test_ctc_golden(F, golden, name):
    g = golden("ctc")
    nll = _check_ctc(F, g[name + ":lp"], g[name + ":tg"], g[name + ":il"], g[name + ":tl"])
    np.testing.assert_allclose(nll.cpu().numpy(), g[name + ":nll"], rtol=1e-4, atol=1e-5)
    # vs the reference's own numbers (fp32 torch through nn.CTCLoss)
    lp = torch.from_numpy(g[name + ":lp"]).cuda()
    _, _, grad = F.ctc_loss_raw(lp, torch.from_numpy(g[name + ":tg"]).cuda(), torch.from_numpy(g[name + ":il"]).cuda(),
                                torch.from_numpy(g[name + ":tl"]).cuda())
    gref = g[name + ":grad"]
    assert np.abs(grad.cpu().numpy() - gref).max() <= 2e-3 * np.abs(gref).max() + 1e-7


@pytest.mark.parametrize("N,T,S,C", [(8, 500, 150, 29), (64, 750, 225, 29), (4, 200, 50, 29), (3, 1000, 300, 29), (2, 1300, 600, 29),
                                     (5, 64, 10, 5), (2, 40, 1, 29)])
@pytest.mark.parametrize("from_logits", [False, True])
def test_ctc_random(F, N, T, S, C, from_logits):
    g = torch.Generator().manual_seed(T + S)
    x = torch.randn(N, T, C, generator=g) * 1.5
    lp = x if from_logits else torch.log_softmax(x, -1)
    tg = torch.randint(1, C, (N, S), generator=g, dtype=torch.int32)
    tg[:, 1::3] = tg[:, 0::3][:, : tg[:, 1::3].shape[1]]         # force repeated labels
    il = torch.randint(max(1, T // 2), T + 1, (N,), generator=g, dtype=torch.int32)
    tl = torch.randint(0, S + 1, (N,), generator=g, dtype=torch.int32)
    il[0], tl[0] = T, S
    for n in range(N):
        tg[n, tl[n]:] = 0
    _check_ctc(F, lp, tg, il, tl, from_logits=from_logits)


def test_ctc_serial_schedule(F, monkeypatch):
    """the large-lattice schedule (alpha, then beta fused with the gradient) must agree with the oracle too"""
    monkeypatch.setenv("W2L_CTC_DBG", "8")
    g = torch.Generator().manual_seed(11)
    N, T, S, C = 5, 90, 20, 29
    lp = torch.log_softmax(torch.randn(N, T, C, generator=g), -1)
    tg = torch.randint(1, C, (N, S), generator=g, dtype=torch.int32)
    il = torch.tensor([90, 64, 41, 90, 12], dtype=torch.int32)
    tl = torch.tensor([20, 7, 20, 0, 11], dtype=torch.int32)
    _check_ctc(F, lp, tg, il, tl)


def test_ctc_transposed_view(F):
    """the reference passes out.transpose(0,1): a [T,N,C] view of the contiguous [N,T,C] tensor."""
    g = torch.Generator().manual_seed(9)
    N, T, C, S = 6, 120, 29, 30
    lp = torch.log_softmax(torch.randn(T, N, C, generator=g), -1)          # genuinely [T,N,C]-contiguous
    tg = torch.randint(1, C, (N, S), generator=g, dtype=torch.int32)
    il = torch.full((N,), T, dtype=torch.int32)
    tl = torch.full((N,), S, dtype=torch.int32)
    loss, nll, grad = F.ctc_loss_raw(lp.cuda().transpose(0, 1), tg.cuda(), il.cuda(), tl.cuda())
    loss_ref, grad_ref = O.ctc_loss_torch(lp.transpose(0, 1).contiguous(), tg, il, tl, dtype=torch.float64)
    assert abs(loss.item() - loss_ref.item()) <= 1e-4 * abs(loss_ref.item())
    assert (grad.cpu().double() - grad_ref).abs().max() <= 2e-3 * grad_ref.abs().max()


def test_ctc_full_size_properties(F):
    """N=512, T=3000, S=600: per-frame gradient rows sum to ~0 (softmax minus occupancies), loss finite & positive,
    and equals the loss of a 16-utterance slice computed by the oracle."""
    g = torch.Generator().manual_seed(1)
    N, T, S, C = 512, 3000, 600, 29
    lp = torch.log_softmax(torch.randn(N, T, C, generator=g), -1)
    tg = torch.randint(1, C, (N, S), generator=g, dtype=torch.int32)
    il = torch.full((N,), T, dtype=torch.int32)
    tl = torch.full((N,), S, dtype=torch.int32)
    loss, nll, grad = F.ctc_loss_raw(lp.cuda(), tg.cuda(), il.cuda(), tl.cuda())
    assert torch.isfinite(nll).all() and (nll > 0).all()
    assert grad.sum(-1).abs().max().item() < 1e-6
    ref = TF.ctc_loss(lp[:4].double().transpose(0, 1), tg[:4], il[:4], tl[:4], reduction="none")
    np.testing.assert_allclose(nll[:4].cpu().numpy(), ref.numpy(), rtol=1e-4)


# ------------------------------------------------------------------------------------------- conv (tcgen05)
def _bf(x):
    return x.to(torch.bfloat16).float()


def _pack_w(w, cout_pad):
    """[Co, Ci, k] fp32 -> packed [k, Co_pad, Ci] bf16 (zero rows in the pad)."""
    co, ci, k = w.shape
    p = torch.zeros(k, cout_pad, ci, dtype=torch.bfloat16)
    p[:, :co] = w.permute(2, 0, 1).to(torch.bfloat16)
    return p


CONV_CASES = [
    # B, T, Cin, Cout, k, d, pad(left,right: rows of zero padding)
    (2, 300, 64, 256, 1, 1, (0, 0)),
    (3, 200, 128, 224, 5, 1, (2, 2)),
    (2, 131, 64, 160, 11, 1, (5, 5)),
    (2, 260, 192, 29, 1, 1, (0, 0)),
    (2, 150, 64, 96, 7, 2, (6, 6)),
    (1, 128, 704, 256, 1, 1, (0, 0)),
    (2, 97, 256, 384, 3, 1, (1, 1)),
    (1, 200, 640, 640, 3, 1, (1, 1)),          # N = 640: 256-wide tiles with a partly empty last tile
    (2, 140, 896, 896, 2, 2, (1, 1)),          # N = 896: 224-wide tiles (fwd/dgrad), 256-wide 4-D chunked tiles (wgrad)
    (1, 150, 704, 320, 1, 1, (0, 0)),
]


# the FWD-kind GEMMs (forward, backward-data over the transposed weight shadow) run as CTA pairs (tcgen05 cta_group::2) by default and
# as the single-CTA kernel under W2L_CG2=0 (read per call): every case under both
CG2_CASES = CONV_CASES + [
    (3, 300, 128, 512, 3, 1, (1, 1)),          # 3 M tiles x 3 utterances: pairs along M, the last pair half empty
    (4, 300, 128, 160, 3, 1, (1, 1)),          # 3 M tiles x 4 utterances: pairs along the batch
    (1, 100, 64, 29, 1, 1, (0, 0)),            # a single M tile: the peer CTA of the only pair has no rows
    (2, 520, 256, 1024, 1, 1, (0, 0)),         # 4 N tiles in groups of 4 (small weights)
]


@pytest.fixture(params=["pair", "single"])
def gemm_kernel(request, monkeypatch):
    if request.param == "pair" and request.config.getoption("--emulate-gpu"):
        pytest.skip("the CTA-pair kernel (cluster of 2, cta_group::2) is outside the emulated surface")
    monkeypatch.setenv("W2L_CG2", "1" if request.param == "pair" else "0")
    return request.param


@pytest.mark.parametrize("B,T,Cin,Cout,k,d,pad", CG2_CASES)
def test_conv_fwd_dgrad_wgrad(F, gemm_kernel, B, T, Cin, Cout, k, d, pad):
    g = torch.Generator().manual_seed(B * T + Cin + k)
    pl, pr = pad
    x = _bf(torch.randn(B, T, Cin, generator=g))                 # time-major, UNpadded: zero padding via TMA OOB fill
    w = _bf(torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5)
    bias = torch.randn(Cout, generator=g)
    T_out = T + pl + pr - d * (k - 1)
    cout_pad = max(64, (Cout + 15) // 16 * 16)
    ldy = (Cout + 7) // 8 * 8
    # ---- reference: plain torch fp32 on the same bf16-rounded operands
    xr = x.transpose(1, 2).clone().requires_grad_(True)          # NCW
    wr = w.clone().requires_grad_(True)
    y_ref = TF.conv1d(TF.pad(xr, (pl, pr)), wr, bias, dilation=d)
    dy = _bf(torch.randn(B, Cout, T_out, generator=g))
    y_ref.backward(dy)
    # ---- forward
    xc, wc = x.to(torch.bfloat16).cuda(), _pack_w(w, cout_pad).cuda()
    desc = F.make_desc(B, T_out, Cin, Cout, cout_pad, k, d, T, -pl, T_out, 0, ldy, F.DT_F32, F.ACT_NONE)
    y = torch.zeros(B, T_out, ldy, dtype=torch.float32, device="cuda")
    F.conv1d_fwd(xc, wc, desc, y, bias=bias.cuda())
    got = y[:, :, :Cout].cpu()
    want = y_ref.detach().transpose(1, 2)
    assert rel_l2(got, want) < 2e-5, rel_l2(got, want)           # fp32 accumulate of identical bf16 operands
    # bf16 output + fused scale/shift + clamp epilogue
    desc2 = F.make_desc(B, T_out, Cin, Cout, cout_pad, k, d, T, -pl, T_out + 3, 2, ldy, F.DT_BF16, F.ACT_CLAMP20)
    y2 = torch.full((B, T_out + 3, ldy), 7.0, dtype=torch.bfloat16, device="cuda")
    sc, sh = torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g)
    F.conv1d_fwd(xc, wc, desc2, y2, bias=bias.cuda(), scale=sc.cuda(), shift=sh.cuda())
    want2 = torch.clamp(want * sc + sh, 0, 20)
    assert rel_l2(y2[:, 2:2 + T_out, :Cout].float().cpu(), want2) < 6e-3
    assert (y2[:, :2] == 7.0).all() and (y2[:, 2 + T_out:] == 7.0).all()      # rows outside [off, off+T_out) untouched
    # BatchNorm statistics from the epilogue: sum / sum of squares of exactly the bf16 values it stored
    desc_s = F.make_desc(B, T_out, Cin, Cout, cout_pad, k, d, T, -pl, T_out, 0, ldy, F.DT_BF16, F.ACT_NONE)
    ys = torch.zeros(B, T_out, ldy, dtype=torch.bfloat16, device="cuda")
    st = torch.zeros(2 * Cout, dtype=torch.float32, device="cuda")
    F.conv1d_fwd(xc, wc, desc_s, ys, bn_stats=st)
    yd = ys[:, :, :Cout].double().reshape(-1, Cout)
    assert torch.allclose(st[:Cout].double(), yd.sum(0), rtol=1e-4, atol=1e-3 * float(yd.abs().sum(0).max()) / 100)
    assert torch.allclose(st[Cout:].double(), (yd * yd).sum(0), rtol=1e-4)
    assert torch.equal(F.bn_stats(ys[:, :, :Cout].contiguous(), Cout)[Cout:] > 0, st[Cout:] > 0) if Cout % 8 == 0 else True
    # ---- dgrad
    dyc = torch.zeros(B, T_out, cout_pad, dtype=torch.bfloat16, device="cuda")
    dyc[:, :, :Cout] = dy.transpose(1, 2).to(torch.bfloat16).cuda()
    desc3 = F.make_desc(B, T_out, Cin, Cout, cout_pad, k, d, T, -pl, T_out, 0, cout_pad)
    dx = torch.empty(B, T, Cin, dtype=torch.bfloat16, device="cuda")
    F.conv1d_dgrad(dyc, wc, desc3, dx)
    assert rel_l2(dx.float().cpu(), xr.grad.transpose(1, 2)) < 6e-3
    # ---- dgrad through the transposed (K-major) weight shadow
    cin_pad = (Cin + 15) // 16 * 16
    wt = torch.full((k, cin_pad, cout_pad), 9.0, dtype=torch.bfloat16, device="cuda")
    F.pack_wt(w.permute(2, 0, 1).contiguous().cuda(), wt, Cout, Cin)
    assert torch.equal(wt[:, :Cin, :Cout].float().cpu(), w.permute(2, 1, 0).flip(0)) and (wt[:, :, Cout:] == 0).all()
    dx2 = torch.empty(B, T, Cin, dtype=torch.bfloat16, device="cuda")
    F.conv1d_dgrad_wt(dyc, wt, desc3, dx2)
    assert rel_l2(dx2.float().cpu(), xr.grad.transpose(1, 2)) < 6e-3
    # ---- wgrad
    dw = torch.full((k, Cout, Cin), 3.0, dtype=torch.float32, device="cuda")
    F.conv1d_wgrad(dyc, xc, desc3, dw)
    assert rel_l2(dw.cpu(), wr.grad.permute(2, 0, 1)) < 2e-5


def _tf32(x):
    """fp32 with the mantissa cut to 10 bits, as kind::tf32 reads its operands"""
    return (x.contiguous().view(torch.int32) & -8192).view(torch.float32)


@pytest.mark.parametrize("B,T,Cin,Cout,k,d,pad", [(2, 150, 64, 256, 1, 1, (0, 0)), (2, 140, 128, 224, 5, 1, (2, 2)), (2, 150, 64, 96, 7, 2, (6, 6)),
                                                  (3, 203, 192, 384, 3, 1, (1, 1)), (1, 130, 72, 40, 3, 1, (1, 1)), (2, 260, 160, 29, 1, 1, (0, 0))])
def test_conv_fp32_operands_tf32(F, B, T, Cin, Cout, k, d, pad):
    """the fp32-faithful mode (w2l_conv_desc::x_dtype = F32): fp32 activations / weights in memory, tcgen05 kind::tf32, fp32
    accumulation -- forward, backward-data (transposed fp32 weight shadow) and the weight gradient over transposed operands
    (w2l_conv1d_wgrad_t), against torch fp32 on operands cut to tf32 (rel-L2 <= 2e-5: only the summation order differs) and against
    plain fp32 (<= 1.5e-3: the tf32 rounding of the operands, the stated tolerance of this mode)"""
    g = torch.Generator().manual_seed(B * T + Cin + k + 1)
    pl, pr = pad
    x = torch.randn(B, T, Cin, generator=g)
    w = torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5
    bias = torch.randn(Cout, generator=g)
    T_out = T + pl + pr - d * (k - 1)
    cout_pad = max(64, (Cout + 15) // 16 * 16)
    ldy = (Cout + 7) // 8 * 8
    dy = torch.randn(B, Cout, T_out, generator=g)

    def reference(xv, wv, dyv):
        xr, wr = xv.transpose(1, 2).clone().requires_grad_(True), wv.clone().requires_grad_(True)
        y = TF.conv1d(TF.pad(xr, (pl, pr)), wr, None, dilation=d)
        y.backward(dyv)
        return y.detach().transpose(1, 2), xr.grad.transpose(1, 2), wr.grad.permute(2, 0, 1)

    y_t, dx_t, dw_t = reference(_tf32(x), _tf32(w), _tf32(dy))          # what the tensor core computes, up to summation order
    y_f, dx_f, dw_f = reference(x, w, dy)
    xc = x.cuda()
    wc = torch.zeros(k, cout_pad, Cin, device="cuda")
    wc[:, :Cout] = w.permute(2, 0, 1).cuda()
    # ---- forward, fp32 output + bias + BatchNorm statistics of the (unrounded) fp32 output
    desc = F.make_desc(B, T_out, Cin, Cout, cout_pad, k, d, T, -pl, T_out, 0, ldy, F.DT_F32, F.ACT_NONE, x_dtype=F.DT_F32)
    y = torch.zeros(B, T_out, ldy, device="cuda")
    st = torch.zeros(2 * Cout, device="cuda")
    F.conv1d_fwd(xc, wc, desc, y, bn_stats=st)
    got = y[:, :, :Cout].cpu()
    assert rel_l2(got, y_t) < 2e-5, rel_l2(got, y_t)
    assert rel_l2(got, y_f) < 1.5e-3, rel_l2(got, y_f)
    yd = got.double().reshape(-1, Cout)
    assert torch.allclose(st[:Cout].cpu().double(), yd.sum(0), rtol=1e-4, atol=1e-3) and torch.allclose(st[Cout:].cpu().double(), (yd * yd).sum(0), rtol=1e-4)
    y2 = torch.zeros(B, T_out, ldy, device="cuda")
    F.conv1d_fwd(xc, wc, desc, y2, bias=bias.cuda())
    assert rel_l2(y2[:, :, :Cout].cpu(), y_t + bias) < 2e-5
    # ---- backward-data over the transposed, tap-reversed fp32 weights: dx fp32
    dyc = torch.zeros(B, T_out, cout_pad, device="cuda")
    dyc[:, :, :Cout] = dy.transpose(1, 2).cuda()
    desc3 = F.make_desc(B, T_out, Cin, Cout, cout_pad, k, d, T, -pl, T_out, 0, cout_pad, x_dtype=F.DT_F32)
    cin_pad = (Cin + 15) // 16 * 16
    wt = torch.zeros(k, cin_pad, cout_pad, device="cuda")
    wt[:, :Cin, :Cout] = w.permute(2, 1, 0).flip(0).cuda()
    dx = torch.full((B, T, Cin), 7.0, device="cuda")
    F.conv1d_dgrad_wt(dyc, wt, desc3, dx)
    assert rel_l2(dx.cpu(), dx_t) < 2e-5 and rel_l2(dx.cpu(), dx_f) < 1.5e-3
    # ---- weight gradient: operands transposed to [B, C, rows] inside conv1d_wgrad_t (time contiguous: K-major like the forward GEMM)
    xT = F.tm_to_ct_f32(xc, T, lead=3)
    assert torch.equal(xT[:, :, 3:T + 3].cpu(), x.transpose(1, 2)) and (xT[:, :, :3] == 0).all()
    dw = torch.full((k, Cout, Cin), 3.0, device="cuda")
    F.conv1d_wgrad_t(dyc, xc, desc3, dw)
    assert rel_l2(dw.cpu(), dw_t) < 2e-5 and rel_l2(dw.cpu(), dw_f) < 1.5e-3


@pytest.mark.parametrize("B,T,C,Co,k,d,pl,pr,act,drop,flat,masked", [
    (3, 150, 256, 128, 5, 1, 2, 2, 2, 0.0, True, False),       # Wav2Letter: reflect halo (k-1)d/2 each side, clamp, one flat row space
    (2, 140, 896, 64, 29, 2, 28, 28, 2, 0.25, True, False),    # halo 28, dropout keep-bits, 4 N tiles of 224
    (4, 97, 264, 96, 3, 1, 0, 0, 1, 0.0, False, True),         # Jasper: no halo, ReLU, length mask, C not a multiple of 32
    (2, 260, 128, 256, 7, 1, 3, 3, 1, 0.1, False, False),      # per-utterance launch with a halo
])
def test_dgrad_fused_bn_reduce(F, gemm_kernel, B, T, C, Co, k, d, pl, pr, act, drop, flat, masked):
    """w2l_conv1d_dgrad_wt_bnred: the BatchNorm-backward reduction of the block that produced a layer's input, formed in the epilogue
    of that layer's backward-data GEMM -- against the separate reduce pass over the gradient the plain GEMM stored (sums to fp32
    summation order), and the apply pass fed with the raw sums (W2L_RED_RAW) against the two-pass result"""
    g = torch.Generator().manual_seed(B * T + C + k)
    Tp = pl + T + pr                                 # rows of the block's padded output = this conv's input
    halo = (k - 1) * d
    T_out = Tp - halo                                # this conv's output rows (valid convolution over the padded input)
    z = _bf(torch.randn(B, T, C, generator=g) * 1.5 + 0.3).to(torch.bfloat16).cuda()          # the block's conv output
    gamma, beta = (torch.rand(C, generator=g) + 0.5).cuda(), torch.randn(C, generator=g).cuda()
    stats = F.bn_stats(z, C)
    fin = F.bn_finalize(stats, B * T, C, gamma, beta, None, 1e-3, 0.1, torch.zeros(C, device="cuda"), torch.ones(C, device="cuda"))
    lens = torch.randint(T // 2, T + 1, (B,), generator=g, dtype=torch.int32).cuda() if masked else None
    seed, mask = 77, (torch.zeros(B * T * C // 8, dtype=torch.uint8, device="cuda") if drop > 0 else None)
    F.bn_act_pad(z, fin[0], fin[1], B, T, C, pl, pr, act, drop, seed, lens, drop_mask=mask)   # (writes the keep-bits)
    w = _bf(torch.randn(Co, C, k, generator=g) / (C * k) ** 0.5)
    dy = _bf(torch.randn(B, T_out, Co, generator=g))
    wt = torch.empty(k, (C + 15) // 16 * 16, max(64, (Co + 15) // 16 * 16), dtype=torch.bfloat16, device="cuda")
    F.pack_wt(w.permute(2, 0, 1).contiguous().cuda(), wt, Co, C)
    cout_pad = wt.shape[2]
    dz_up = torch.zeros(B, Tp if flat else T_out, cout_pad, dtype=torch.bfloat16, device="cuda")      # the layer's own dz, zero tails when flat
    dz_up[:, :T_out, :Co] = dy.to(torch.bfloat16).cuda()
    desc = (F.make_desc(1, B * Tp, C, Co, cout_pad, k, d, B * Tp, 0, B * Tp, 0, cout_pad) if flat
            else F.make_desc(B, T_out, C, Co, cout_pad, k, d, Tp, 0, T_out, 0, cout_pad))
    # ---- two-pass reference: plain GEMM, then reduce + apply over what it stored
    dx_a = torch.empty(B, Tp, C, dtype=torch.bfloat16, device="cuda")
    F.conv1d_dgrad_wt(dz_up, wt, desc, dx_a)
    dz_a, red_a, _ = F.bn_act_bwd(dx_a, z, fin[0], fin[1], fin[2], fin[3], gamma, B, T, C, pl, pr, act, drop, seed, lens, drop_mask=mask)
    # ---- fused: the GEMM's epilogue accumulates the raw sums, the apply pass takes them
    dx_b = torch.empty_like(dx_a)
    red_raw = torch.zeros(2 * C, device="cuda")
    F.conv1d_dgrad_wt(dz_up, wt, desc, dx_b, bnred=dict(z=z, mask=mask, scale=fin[0], shift=fin[1], mean=fin[2], lens=lens, red=red_raw, B=B, T=T,
                                                       pad_left=pl, pad_right=pr, act=act, drop_p=drop))
    assert torch.equal(dx_a, dx_b)
    dz_b, red_b, _ = F.bn_act_bwd(dx_b, z, fin[0], fin[1], fin[2], fin[3], gamma, B, T, C, pl, pr, act, drop, seed, lens, drop_mask=mask,
                                  red_raw=red_raw)
    scale_ref = float(red_a.abs().max())
    assert float((red_b - red_a).abs().max()) < 2e-4 * max(scale_ref, 1.0), float((red_b - red_a).abs().max())
    assert rel_l2(red_b, red_a) < 1e-4 and rel_l2(dz_b.float(), dz_a.float()) < 2e-3


@pytest.mark.parametrize("B,T,C,Co,k,d", [(3, 200, 64, 128, 5, 1), (5, 131, 128, 64, 7, 2), (2, 750, 256, 256, 11, 1)])
def test_conv_dgrad_flat_prepadded(F, gemm_kernel, B, T, C, Co, k, d):
    """Wav2Letter layout: the input carries its own halo (x_rows = T + (k-1)d), dz is stored with the input's row pitch and zero
    tails, and backward-data runs over ONE flat [B*x_rows] row space; wgrad reads the same pitched dz."""
    g = torch.Generator().manual_seed(B + T + k)
    halo = (k - 1) * d
    Tp = T + halo
    xp = _bf(torch.randn(B, Tp, C, generator=g))
    w = _bf(torch.randn(Co, C, k, generator=g) / (C * k) ** 0.5)
    dy = _bf(torch.randn(B, T, Co, generator=g))
    xr = xp.transpose(1, 2).clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    TF.conv1d(xr, wr, dilation=d).backward(dy.transpose(1, 2))
    dz = torch.zeros(B, Tp, Co, dtype=torch.bfloat16, device="cuda")
    dz[:, :T] = dy.to(torch.bfloat16).cuda()
    wt = torch.empty(k, C, Co, dtype=torch.bfloat16, device="cuda")
    F.pack_wt(w.permute(2, 0, 1).contiguous().cuda(), wt, Co, C)
    dx = torch.empty(B, Tp, C, dtype=torch.bfloat16, device="cuda")
    F.conv1d_dgrad_wt(dz, wt, F.make_desc(1, B * Tp, C, Co, Co, k, d, B * Tp, 0, B * Tp, 0, Co), dx)
    assert rel_l2(dx.float().cpu(), xr.grad.transpose(1, 2)) < 6e-3
    dw = torch.empty(k, Co, C, dtype=torch.float32, device="cuda")
    F.conv1d_wgrad(dz, xp.to(torch.bfloat16).cuda(), F.make_desc(B, T, C, Co, Co, k, d, Tp, 0, Tp, 0, Co), dw)
    assert rel_l2(dw.cpu(), wr.grad.permute(2, 0, 1)) < 2e-5


def test_conv_full_size_layer(F, gemm_kernel):
    """one config-2-sized layer (B=64, T'=750, 896->896, k=29, d=2): linearity + spot-check vs torch fp32 on a slice."""
    g = torch.Generator().manual_seed(3)
    B, T, C, k, d = 64, 750, 896, 29, 2
    pad = (k - 1) * d // 2
    x = torch.randn(B, T, C, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(C, C, k, generator=g) / (C * k) ** 0.5).to(torch.bfloat16)
    wc = w.permute(2, 0, 1).contiguous().cuda()
    desc = F.make_desc(B, T, C, C, C, k, d, T, -pad, T, 0, C, F.DT_F32, F.ACT_NONE)
    y = torch.empty(B, T, C, dtype=torch.float32, device="cuda")
    F.conv1d_fwd(x, wc, desc, y)
    ref = TF.conv1d(x[5:6].float().cpu().transpose(1, 2), w.float(), padding=pad, dilation=d).transpose(1, 2)
    assert rel_l2(y[5:6].cpu(), ref) < 1e-4      # K = 25 984 terms, fp32 accumulation order differs
    y2 = torch.empty_like(y)
    F.conv1d_fwd((x.float() * 2).to(torch.bfloat16), wc, desc, y2)           # exact in bf16: linearity
    assert rel_l2(y2, 2 * y) < 1e-6


# ------------------------------------------------------------------------------------------- elementwise
def test_im2col_and_transposes(F):
    g = torch.Generator().manual_seed(4)
    B, Fdim, T = 3, 64, 201
    x = torch.randn(B, Fdim, T, generator=g)
    pl, pr = O.reflect_pad_amounts(64, 11, 2, 1)
    rows = (T + pl + pr - 11) // 2 + 1
    got = F.im2col_ncw(x.cuda(), rows, 11, 2, 1, pl, F.PAD_REFLECT).float().cpu()
    xp = TF.pad(x, (pl, pr), mode="reflect")
    want = xp.unfold(2, 11, 2).permute(0, 2, 3, 1).reshape(B, rows, 11 * Fdim)      # [B, rows, j*F + f]
    assert torch.equal(got, _bf(want))
    # padded time-major copy with zero pad + length mask
    lens = torch.tensor([201, 100, 7], dtype=torch.int32)
    got = F.im2col_ncw(x.cuda(), T + 4, 1, 1, 1, 2, F.PAD_ZERO, lens.cuda()).float().cpu()
    want = TF.pad(x * (torch.arange(T)[None, None] < lens[:, None, None]), (2, 2)).transpose(1, 2)
    assert torch.equal(got, _bf(want))
    back = F.tm_to_ncw(got.to(torch.bfloat16).cuda(), T, Fdim, x_row_offset=2).cpu()
    assert torch.equal(back, _bf(want[:, 2:2 + T]).transpose(1, 2))


@pytest.mark.parametrize("act,drop", [(2, 0.0), (1, 0.0), (2, 0.25)])
def test_bn_act_forward_backward(F, act, drop):
    g = torch.Generator().manual_seed(5)
    B, T, C, pl, pr = 3, 90, 264, 4, 5
    z = _bf(torch.randn(B, T, C, generator=g) * 2 + 0.5)
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    rm, rv = torch.zeros(C), torch.ones(C)
    zc = z.to(torch.bfloat16).cuda()
    stats = F.bn_stats(zc, C)
    np.testing.assert_allclose(stats[:C].cpu().numpy(), z.sum((0, 1)).numpy(), rtol=1e-4, atol=1e-2)
    rmc, rvc = rm.cuda(), rv.cuda()
    fin = F.bn_finalize(stats, B * T, C, gamma.cuda(), beta.cuda(), None, 1e-3, 0.9, rmc, rvc)
    scale, shift, mean, invstd = fin[0], fin[1], fin[2], fin[3]
    # torch reference (channel-first)
    zr = z.transpose(1, 2).clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    bn = TF.batch_norm(zr, rm, rv, gr, br, training=True, momentum=0.9, eps=1e-3)
    np.testing.assert_allclose(rmc.cpu().numpy(), rm.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(rvc.cpu().numpy(), rv.numpy(), rtol=1e-4, atol=1e-5)
    seed = 1234
    if drop > 0:   # export the kernel's own keep-mask: z=0, scale=0, shift=1, no act
        ones = F.bn_act_pad(torch.zeros_like(zc), torch.zeros(C, device="cuda"), torch.ones(C, device="cuda"), B, T, C, 0, 0, 0, drop, seed)
        mult = ones.float().cpu().transpose(1, 2)
        keep = (mult > 0).float().mean().item()
        assert abs(keep - (1 - drop)) < 0.02
        np.testing.assert_allclose(mult[mult > 0].numpy(), 1 / (1 - drop), rtol=1e-2)
        bn = bn * mult
    y_ref = torch.clamp(bn, 0, 20) if act == 2 else torch.relu(bn)
    yp_ref = TF.pad(y_ref, (pl, pr), mode="reflect")
    yp = F.bn_act_pad(zc, scale, shift, B, T, C, pl, pr, act, drop, seed)
    assert rel_l2(yp.float().cpu(), yp_ref.detach().transpose(1, 2)) < 4e-3
    dyp = _bf(torch.randn(B, C, T + pl + pr, generator=g))
    yp_ref.backward(dyp)
    dz, red, _ = F.bn_act_bwd(dyp.transpose(1, 2).to(torch.bfloat16).contiguous().cuda(), zc, scale, shift, mean, invstd, gamma.cuda(),
                              B, T, C, pl, pr, act, drop, seed)
    assert rel_l2(dz.float().cpu(), zr.grad.transpose(1, 2)) < 1e-2
    assert rel_l2(red[:C].cpu(), br.grad) < 5e-3 and rel_l2(red[C:].cpu(), gr.grad) < 5e-3
    if drop > 0:   # keep-bits stored by the forward pass and read back by the backward passes == re-derived Philox bits
        bits = torch.zeros(B * T * C // 8, dtype=torch.uint8, device="cuda")
        yp2 = F.bn_act_pad(zc, scale, shift, B, T, C, pl, pr, act, drop, seed, drop_mask=bits)
        assert torch.equal(yp2, yp)
        dz2, red2, _ = F.bn_act_bwd(dyp.transpose(1, 2).to(torch.bfloat16).contiguous().cuda(), zc, scale, shift, mean, invstd,
                                    gamma.cuda(), B, T, C, pl, pr, act, drop, seed, drop_mask=bits)
        assert rel_l2(dz2.float(), dz.float()) < 1e-3 and rel_l2(red2, red) < 1e-4
    dzp, _, _ = F.bn_act_bwd(dyp.transpose(1, 2).to(torch.bfloat16).contiguous().cuda(), zc, scale, shift, mean, invstd, gamma.cuda(),
                             B, T, C, pl, pr, act, drop, seed, dz_rows=T + 7)
    assert rel_l2(dzp[:, :T].float(), dz.float()) < 1e-3 and (dzp[:, T:] == 0).all()      # reductions use fp32 atomics: last-bit noise


@pytest.mark.parametrize("B,T,C,pl,pr,act,res", [(2, 37, 64, 0, 0, 1, False), (3, 50, 264, 4, 5, 2, False), (2, 41, 896, 28, 28, 2, False),
                                                 (2, 33, 768, 0, 0, 1, True), (1, 19, 2056, 2, 3, 1, True), (5, 7, 8, 1, 1, 2, False)])
def test_bn_passes_geometry_and_fused_finalize(F, B, T, C, pl, pr, act, res):
    """the row-looping BatchNorm / activation passes over every thread geometry (C/8 = 8 ... 257 channel vectors: bx x ny = 8x32 ... 256x1,
    two column blocks at 2056), the finalize fold (w2l_bn_finalize_act_pad == w2l_bn_finalize + w2l_bn_act_pad to 1 ulp), the
    persistent reduction buffer (red_ws) and the buffers the passes clear for each other -- against torch autograd"""
    g = torch.Generator().manual_seed(C + T)
    z = _bf(torch.randn(B, T, C, generator=g) * 1.5 + 0.3)
    zr = _bf(torch.randn(B, T, C, generator=g)) if res else None
    gamma, beta, cbias = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g), torch.randn(C, generator=g)
    lens = torch.randint(T // 2, T + 1, (B,), generator=g, dtype=torch.int32) if res else None
    zc = z.to(torch.bfloat16).cuda()
    zrc = zr.to(torch.bfloat16).cuda() if res else None
    lens_c = lens.cuda() if res else None
    rs, rsh = (torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)) if res else (None, None)
    stats = F.bn_stats(zc, C)
    rm0, rv0 = torch.randn(C, generator=g), torch.rand(C, generator=g) + 0.5
    # ---- separate finalize + apply
    rm_a, rv_a, nbt_a = rm0.clone().cuda(), rv0.clone().cuda(), torch.tensor(3, dtype=torch.long).cuda()
    fin_a = F.bn_finalize(stats, B * T, C, gamma.cuda(), beta.cuda(), cbias.cuda(), 1e-3, 0.1, rm_a, rv_a, nbt_a)
    kw = dict(res=zrc, res_scale=None if not res else rs.cuda(), res_shift=None if not res else rsh.cuda())
    y_a = F.bn_act_pad(zc, fin_a[0], fin_a[1], B, T, C, pl, pr, act, 0.0, 0, lens_c, **kw)
    # ---- folded: same bits, and the launch clears the buffer handed to it
    rm_b, rv_b, nbt_b = rm0.clone().cuda(), rv0.clone().cuda(), torch.tensor(3, dtype=torch.long).cuda()
    dirty = torch.full((2 * C,), 7.0, device="cuda")
    y_b, fin_b = F.bn_finalize_act_pad(zc, stats, gamma.cuda(), beta.cuda(), cbias.cuda(), 1e-3, 0.1, rm_b, rv_b, nbt_b, B, T, C, pl, pr, act,
                                       0.0, 0, lens_c, zero_after=dirty, **kw)
    # the fold derives 1/sqrt(var + eps) with rsqrt + one Newton step where the standalone kernel divides in fp64: <= 1 ulp apart
    torch.testing.assert_close(fin_a, fin_b, rtol=3e-7, atol=1e-7)
    torch.testing.assert_close(rm_a, rm_b, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(rv_a, rv_b, rtol=1e-6, atol=1e-7)
    assert rel_l2(y_b.float(), y_a.float()) < 1e-3
    assert int(nbt_b) == 4 and int(nbt_a) == 4 and float(dirty.abs().max()) == 0.0
    # ---- torch reference
    zt = z.transpose(1, 2).clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_t, rv_t = rm0.clone(), rv0.clone()
    bn = TF.batch_norm(zt + cbias[None, :, None], rm_t, rv_t, gr, br, training=True, momentum=0.1, eps=1e-3)
    np.testing.assert_allclose(rm_b.cpu().numpy(), rm_t.numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(rv_b.cpu().numpy(), rv_t.numpy(), rtol=1e-3, atol=1e-4)
    if res:
        bn = bn + (zr.transpose(1, 2) * rs[None, :, None] + rsh[None, :, None])
    y_ref = torch.clamp(bn, 0, 20) if act == 2 else torch.relu(bn)
    if res:
        keep = (torch.arange(T)[None, :] < lens[:, None]).float()[:, None, :]
        y_ref = y_ref * keep
    yp_ref = TF.pad(y_ref, (pl, pr), mode="reflect") if pl + pr else y_ref
    assert rel_l2(y_b.float().cpu(), yp_ref.detach().transpose(1, 2)) < 5e-3
    dyp = _bf(torch.randn(B, C, T + pl + pr, generator=g))
    yp_ref.backward(dyp)
    dyc = dyp.transpose(1, 2).to(torch.bfloat16).contiguous().cuda()
    # ---- backward with a persistent reduction buffer: zero on entry, dirty afterwards, sums returned in a fresh tensor
    red_ws = torch.zeros(2 * C, device="cuda")
    dirty.fill_(3.0)
    dz, red, gout = F.bn_act_bwd(dyc, zc, fin_b[0], fin_b[1], fin_b[2], fin_b[3], gamma.cuda(), B, T, C, pl, pr, act, 0.0, 0, lens_c,
                                 want_g=res, dz_rows=T + 3, red_ws=red_ws, zero_after=dirty, **kw)
    assert red.data_ptr() != red_ws.data_ptr() and torch.equal(red, red_ws) and float(dirty.abs().max()) == 0.0
    assert (dz[:, T:] == 0).all()
    scale_err = 2e-2 if T * B < 100 else 1e-2
    assert rel_l2(dz[:, :T].float().cpu(), zt.grad.transpose(1, 2)) < scale_err
    assert rel_l2(red[:C].cpu(), br.grad) < 5e-3 and rel_l2(red[C:].cpu(), gr.grad) < 5e-3
    # ---- and without one (the buffer is allocated zero-filled per call): same numbers up to the order of the atomics
    dz2, red2, _ = F.bn_act_bwd(dyc, zc, fin_b[0], fin_b[1], fin_b[2], fin_b[3], gamma.cuda(), B, T, C, pl, pr, act, 0.0, 0, lens_c,
                                want_g=res, dz_rows=T + 3, **kw)
    assert rel_l2(red2, red) < 1e-5 and rel_l2(dz2.float(), dz.float()) < 1e-3


def test_log_softmax_and_colsum(F):
    g = torch.Generator().manual_seed(6)
    logits = torch.randn(4, 75, 32, generator=g) * 3
    lp = F.log_softmax(logits.cuda(), 29)
    assert rel_l2(lp.cpu(), torch.log_softmax(logits[..., :29], -1)) < 1e-6
    sm = F.log_softmax(logits.cuda(), 29, mode=1)
    assert rel_l2(sm.cpu(), torch.softmax(logits[..., :29], -1)) < 1e-6
    gr = torch.randn(4, 75, 29, generator=g)
    x = logits[..., :29].clone().requires_grad_(True)
    torch.log_softmax(x, -1).backward(gr)
    dl = F.log_softmax_bwd(gr.cuda(), lp, 64)
    assert rel_l2(dl[..., :29].float().cpu(), x.grad) < 5e-3 and (dl[..., 29:] == 0).all()
    m = _bf(torch.randn(1000, 64, generator=g))
    cs = F.colsum(m.to(torch.bfloat16).cuda(), 29)
    np.testing.assert_allclose(cs.cpu().numpy(), m[:, :29].sum(0).numpy(), rtol=1e-4, atol=1e-3)
    w = torch.randn(1001, generator=g)
    assert torch.equal(F.cast_bf16(w.cuda()).cpu(), w.to(torch.bfloat16))


@pytest.mark.parametrize("T", [120, 700])
def test_string_metrics_device(F, T):
    """device-side CER/WER/len-ratio vs the reference's host arithmetic (base_asr_models.py:58-69); T=700: transcripts longer than
    the 128-symbol chunks of the split kernel (words that straddle a chunk boundary, several chunks of running offsets)."""
    import random
    from wav2letter_pytorch_b200.decoder import GreedyDecoder
    labels = O.ENGLISH_LOWERCASE
    dec = GreedyDecoder(labels)
    rnd = random.Random(5)
    N, C = 9, 29
    g = torch.Generator().manual_seed(2)
    probs = torch.softmax(torch.randn(N, T, C, generator=g) * 3, -1)
    probs[:, :, 28] *= 6                                        # plenty of spaces -> many words
    probs[3] = 0
    probs[3, :, 0] = 1                                          # an all-blank (empty) hypothesis
    sizes = torch.tensor([T, T, 60, T, 1, T, 100, T, 7], dtype=torch.int32)
    texts = ["".join(rnd.choice(labels[1:]) for _ in range(rnd.randint(1, 90 if T == 120 else 600))) for _ in range(N)]
    texts[1] = "  " + texts[1] + "  a "                        # leading / trailing / double spaces
    texts[5] = "word"
    hyps = dec.decode(probs.cuda(), sizes.cuda())
    ratios = dec.error_ratios_device(probs.cuda(), sizes.cuda(), texts)
    assert ratios is not None and ratios.is_cuda
    cer = sum(dec.cer(t, h) for t, h in zip(texts, hyps)) / sum(len(t.replace(" ", "")) for t in texts)
    wer = sum(dec.wer(t, h) for t, h in zip(texts, hyps)) / sum(len(t.split()) for t in texts)
    lr = sum(map(len, hyps)) / sum(map(len, texts))
    np.testing.assert_allclose(ratios.cpu().numpy(), [cer, wer, lr], rtol=1e-6)
    assert dec.error_ratios_device(probs.cuda(), sizes.cuda(), ["tab\there"] * N) is None     # exotic whitespace -> host path
    with pytest.raises(ZeroDivisionError):
        dec.error_ratios_device(probs.cuda(), sizes.cuda(), [" "] * N)


@pytest.mark.parametrize("ref_lens", [(1, 7, 8, 9, 33, 255, 256), (257, 600, 1023, 1, 512, 40, 300)])
def test_string_metrics_strip_boundaries(F, ref_lens):
    """the warp-per-pair edit distance at the edges of its column strips (8 columns per lane up to 256 reference symbols, 32 above):
    reference lengths 1 / 8 / 9 / 255 / 256 / 257 / 1023, hypotheses shorter, equal and longer, identical pairs, an empty hypothesis --
    bit-exact sums against the reference's host arithmetic (decoder.py:49-68)"""
    import random
    from wav2letter_pytorch_b200.decoder import GreedyDecoder
    labels = O.ENGLISH_LOWERCASE
    dec = GreedyDecoder(labels)
    rnd = random.Random(sum(ref_lens))
    alpha = [l for l in labels[1:] if len(l) == 1]                # every single-character label but the blank (space included)

    def text(n):
        t = "".join(rnd.choice(alpha) for _ in range(n))
        return t if t.strip() else "a" * n                       # at least one word: the reference divides by the word count
    refs = [text(n) for n in ref_lens]
    hyps = []
    for i, r in enumerate(refs):
        kind = i % 4
        if kind == 0:
            hyps.append(r)                                       # identical
        elif kind == 1:
            hyps.append("".join(ch for ch in r if rnd.random() > 0.3) + text(rnd.randint(0, 5)))
        elif kind == 2:
            hyps.append(text(max(1, len(r) // 2)))
        else:
            hyps.append(text(min(1000, len(r) + rnd.randint(1, 40))))
    hyps[-1] = ""                                               # nothing decoded
    # hypotheses as one-hot scores: every character followed by a blank frame (so repeats survive the collapse)
    T = 2 * max(1, max(len(h) for h in hyps))
    probs = torch.zeros(len(hyps), T, len(labels))
    probs[:, :, 0] = 1.0
    for i, h in enumerate(hyps):
        for j, ch in enumerate(h):
            probs[i, 2 * j, 0] = 0.0
            probs[i, 2 * j, labels.index(ch)] = 1.0
    sizes = torch.full((len(hyps),), T, dtype=torch.int32)
    got_h = dec.decode(probs.cuda(), sizes.cuda())
    assert got_h == hyps
    ratios = dec.error_ratios_device(probs.cuda(), sizes.cuda(), refs)
    assert ratios is not None
    cer = sum(dec.cer(t, h) for t, h in zip(refs, hyps)) / sum(len(t.replace(" ", "")) for t in refs)
    wer = sum(dec.wer(t, h) for t, h in zip(refs, hyps)) / sum(len(t.split()) for t in refs)
    lr = sum(map(len, hyps)) / sum(map(len, refs))
    np.testing.assert_allclose(ratios.cpu().numpy(), [cer, wer, lr], rtol=1e-6)


@pytest.mark.parametrize("B,T,C,k,s,d", [(3, 97, 64, 33, 1, 1), (2, 201, 256, 13, 2, 1), (2, 60, 128, 7, 1, 2), (1, 40, 72, 1, 1, 1)])
def test_depthwise_conv(F, B, T, C, k, s, d):
    """depthwise (groups=C) conv fwd / dgrad / wgrad vs torch fp32 on the same bf16-rounded operands, with the length masks"""
    g = torch.Generator().manual_seed(B * T + k)
    p = (d * k) // 2 - 1 if d > 1 else k // 2                      # jasper.py:61-66
    x = _bf(torch.randn(B, T, C, generator=g))
    w = torch.randn(C, 1, k, generator=g) / k ** 0.5
    T_out = (T + 2 * p - d * (k - 1) - 1) // s + 1
    lens = torch.randint(T_out // 2, T_out + 1, (B,), generator=g, dtype=torch.int32)
    lens[0] = T_out
    xr = x.transpose(1, 2).clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    y_ref = TF.conv1d(xr, wr, stride=s, padding=p, dilation=d, groups=C)
    mask = (torch.arange(T_out)[None] < lens[:, None]).float()[:, None]
    y_ref = y_ref * mask
    dy = _bf(torch.randn(B, C, T_out, generator=g))
    y_ref.backward(dy)
    ws = w.permute(2, 1, 0).reshape(k, C).contiguous().cuda()
    xc = x.to(torch.bfloat16).cuda()
    y = F.depthwise_fwd(xc, ws, T_out, k, s, d, p, lens.cuda())
    assert rel_l2(y.float().cpu(), y_ref.detach().transpose(1, 2)) < 6e-3
    dyc = dy.transpose(1, 2).to(torch.bfloat16).contiguous().cuda()
    dw = F.depthwise_wgrad(dyc, xc, k, s, d, p, lens.cuda())
    assert rel_l2(dw.cpu(), wr.grad.permute(2, 1, 0).reshape(k, C)) < 1e-4
    if s == 1:
        dx = F.depthwise_dgrad(dyc, ws, T, k, d, p, lens.cuda())
        assert rel_l2(dx.float().cpu(), xr.grad.transpose(1, 2)) < 6e-3


# ------------------------------------------------------------------------------------------- feature front-end
def test_features_golden_and_ragged_batch(F, golden):
    """GPU front-end vs the reference's SpectrogramExtractor (golden, same dither noise) and vs the oracle on a ragged batch"""
    from wav2letter_pytorch_b200.features import SpectrogramExtractor
    conf = dict(sample_rate=16000, window_size=0.02, window_stride=0.01, window="hamming")
    ex = SpectrogramExtractor(conf, mel_spec=64).cuda()
    g = golden("features")
    assert np.allclose(ex.fb[0].cpu().numpy(), g["fb"], atol=1e-7)
    for name in ("a", "b"):
        got = ex.extract(g[name + ":signal"], noise=g[name + ":noise"]).cpu().numpy()
        want = g[name + ":feats"]
        assert got.shape == want.shape
        assert np.abs(got - want).max() < 2e-4, (name, np.abs(got - want).max())     # fp32 FFT / matmul summation order
    # ragged batch: per-utterance reflect padding at each utterance's own end, zero padding past its last frame
    rs = np.random.RandomState(3)
    sigs = [(0.2 * rs.randn(n)).astype(np.float32) for n in (24000, 8011, 16000, 400)]
    noise = torch.randn(len(sigs), 24000, generator=torch.Generator().manual_seed(5))
    out, lens = ex.extract_batch(sigs, noise=noise)
    want, want_lens = O.collate_features([O.spectrogram_extract(s, noise=noise[i, :len(s)].numpy()) for i, s in enumerate(sigs)])
    assert lens.cpu().tolist() == want_lens.tolist() == [151, 51, 101, 3]
    assert out.shape == want.shape
    assert (out.cpu() - want).abs().max() < 2e-4
    for i, n in enumerate(lens.cpu().tolist()):
        assert (out[i, :, n:] == 0).all()
    # without explicit noise the dither is drawn on the device: same features to within the dither's effect
    out2, _ = ex.extract_batch(sigs)
    assert (out2 - out).abs().max() < 0.2
    with pytest.raises(RuntimeError):
        ex.extract(np.zeros(200, dtype=np.float32))


def test_lens_chain_matches_reference_arithmetic(F):
    """one launch vs the reference's per-conv tensor arithmetic (jasper.py:91-95,107-119): lens.to(long), then
    (lens + 2p - d(k-1) - 1) / stride + 1 as a float tensor handed to the next conv"""
    g = torch.Generator().manual_seed(9)
    lens0 = torch.randint(1, 3000, (37,), generator=g, dtype=torch.int64)
    chain = [(11, 2, 1, 5), (11, 1, 1, 5), (13, 1, 1, 6), (29, 1, 2, 28), (1, 1, 1, 0), (33, 2, 1, 16), (7, 3, 1, 3), (5, 0, 1, 2), (9, 1, 1, 4)]
    want_rows, lens = [lens0.clone()], lens0
    for k, s, d, p in chain:
        if s != 0:
            lens = lens.to(dtype=torch.long)
            lens = (lens + 2 * p - d * (k - 1) - 1) / s + 1
        want_rows.append(lens.to(dtype=torch.long))
    for dt in (torch.int64, torch.int32):
        rows, final = F.lens_chain(lens0.to(dt).cuda(), chain)
        assert rows.dtype == torch.int32 and final.dtype == torch.int64
        assert torch.equal(rows.cpu().long(), torch.stack(want_rows))
        assert torch.equal(final.cpu(), want_rows[-1])


def test_conv_fwd_tail_split(F):
    """the forward tail split (last wave of tiles cut along K, pieces summed in an fp32 scratch by the last arriver), forced on
    a small problem by capping the persistent grid at 8 SMs: 18 tiles = 2 full waves + 2 tiles -> 4 K-pieces each.  Checked for
    the plain store, the fused affine+clamp bf16 epilogue and the BatchNorm statistics, twice (the arrival counters self-clean)."""
    import ctypes
    from wav2letter_pytorch_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(12)
    B, T, Cin, Cout, k = 9, 200, 256, 256, 11
    pad = k // 2
    x = _bf(torch.randn(B, T, Cin, generator=g))
    w = _bf(torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5)
    want = TF.conv1d(x.transpose(1, 2), w, padding=pad).transpose(1, 2)
    xc, wc = x.to(torch.bfloat16).cuda(), _pack_w(w, Cout).cuda()
    F.ensure_gemm_scratch(xc.device)
    try:
        lib.w2l_set_sm_budget(8)
        desc = F.make_desc(B, T, Cin, Cout, Cout, k, 1, T, -pad, T, 0, Cout, F.DT_F32, F.ACT_NONE)
        assert lib.w2l_conv1d_fwd_tail_parts(ctypes.byref(desc)) == 4
        for rep in range(2):
            y = torch.zeros(B, T, Cout, dtype=torch.float32, device="cuda")
            F.conv1d_fwd(xc, wc, desc, y)
            assert rel_l2(y.cpu(), want) < 2e-5
        sc, sh = torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g)
        desc2 = F.make_desc(B, T, Cin, Cout, Cout, k, 1, T, -pad, T, 0, Cout, F.DT_BF16, F.ACT_CLAMP20)
        y2 = torch.zeros(B, T, Cout, dtype=torch.bfloat16, device="cuda")
        st = torch.zeros(2 * Cout, dtype=torch.float32, device="cuda")
        F.conv1d_fwd(xc, wc, desc2, y2, scale=sc.cuda(), shift=sh.cuda(), bn_stats=st)
        assert rel_l2(y2.float().cpu(), torch.clamp(want * sc + sh, 0, 20)) < 6e-3
        yd = y2.double().reshape(-1, Cout)
        assert torch.allclose(st[:Cout].double(), yd.sum(0), rtol=1e-4, atol=1e-2) and torch.allclose(st[Cout:].double(), (yd * yd).sum(0), rtol=1e-4)
    finally:
        lib.w2l_set_sm_budget(0)
    assert lib.w2l_conv1d_fwd_tail_parts(ctypes.byref(desc)) <= 1          # 18 tiles fit one wave of the full machine


def test_conv_slab_mode_opt_in():
    """the opt-in resident-slab forward path (W2L_SLAB=1, read once per process): the conv parity cases in a fresh interpreter"""
    import os
    import subprocess
    import sys
    env = dict(os.environ, W2L_SLAB="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-m", "gpu", "-k",
                        "conv_fwd_dgrad_wgrad or conv_dgrad_flat or conv_fwd_tail_split", "-p", "no:cacheprovider"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
