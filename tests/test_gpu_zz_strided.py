"""GPU tests written after round 1's last GPU session (they sort last on purpose, so that a surprise here cannot mask the rest of the
suite): strided layers beyond the first one (legal in the reference's Conv1dBlock / JasperBlock, absent from its shipped yamls) --
the unfold / fold kernels against the oracle's restatement of the same index arithmetic and whole models against fixtures frozen
from the unmodified reference (tests/golden/{w2l,jasper}_strided.npz) --, the stand-alone head block, hidden widths that are
multiples of 8 (w2l_narrow.npz), the CTC schedules with alphabets wider than a CTA, and conv shapes at the edges of the tiling.

None of them had run on hardware when committed; all of them run, through the unmodified host package and the library's own C
wrappers and kernels compiled for the host, under `pytest -m gpu --emulate-gpu` (tests/_fake_cuda.py), which the CPU suite does."""
import numpy as np
import pytest
import torch

from oracle import w2l_oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


@pytest.fixture(scope="module")
def F():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from wav2letter_pytorch_b200 import functional
    return functional


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import wav2letter_pytorch_b200 as p
    return p


@pytest.mark.parametrize("B,rows,C,k,s,d,pad", [(3, 109, 64, 7, 2, 1, 0), (2, 51, 72, 5, 2, 1, 2), (2, 40, 8, 5, 3, 1, 2),
                                                (1, 33, 128, 3, 1, 2, 2), (2, 17, 16, 1, 2, 1, 0), (4, 300, 256, 11, 2, 1, 5)])
def test_im2col_tm_and_fold(F, B, rows, C, k, s, d, pad):
    gen = torch.Generator().manual_seed(B * 1000 + rows)
    x = torch.randn(B, rows, C, generator=gen).to(torch.bfloat16)
    T_out = (rows + 2 * pad - d * (k - 1) - 1) // s + 1
    col = F.im2col_tm(x.cuda(), T_out, k, s, d, pad)
    want = O.im2col_tm(x.float().numpy(), T_out, k, s, d, pad)
    assert col.shape == (B, T_out, k * C) and col.dtype == torch.bfloat16
    assert np.array_equal(col.float().cpu().numpy(), want)                      # a pure copy: bit-exact
    dcol = torch.randn(B, T_out, k * C, generator=gen).to(torch.bfloat16)
    dx = F.col2im_tm(dcol.cuda(), rows, C, k, s, d, pad)
    want = O.col2im_tm(dcol.float().numpy(), rows, C, k, s, d, pad)             # float64 sums of the bf16 inputs
    assert dx.shape == (B, rows, C)
    # fp32 accumulation, one bf16 rounding of the result: half an ulp of bf16 (2^-9 relative) plus fp32 summation noise
    np.testing.assert_allclose(dx.float().cpu().numpy(), want, rtol=2.0 ** -8, atol=1e-5)
    # adjoint identity <im2col(x), dcol> == <x, col2im(dcol)> on the exact restatement (guards the test itself)
    lhs = float((O.im2col_tm(x.double().numpy(), T_out, k, s, d, pad) * dcol.double().numpy()).sum())
    rhs = float((x.double().numpy() * want).sum())
    assert abs(lhs - rhs) <= 1e-9 * max(1.0, abs(lhs))


@pytest.mark.parametrize("B,T,C,k,s,d,pad", [(2, 51, 64, 6, 2, 1, 3), (3, 40, 72, 5, 3, 1, 2), (2, 26, 128, 33, 2, 1, 16), (2, 19, 8, 3, 1, 2, 2)])
def test_depthwise_dgrad_strided(F, B, T, C, k, s, d, pad):
    gen = torch.Generator().manual_seed(T * 10 + k)
    T_out = (T + 2 * pad - d * (k - 1) - 1) // s + 1
    dy = torch.randn(B, T_out, C, generator=gen).to(torch.bfloat16)
    w = torch.randn(k, C, generator=gen)
    lens = torch.tensor([T_out] + [max(1, T_out - 2 - i) for i in range(B - 1)], dtype=torch.int32)
    for dl in (None, lens):
        dx = F.depthwise_dgrad(dy.cuda(), w.cuda(), T, k, d, pad, None if dl is None else dl.cuda(), stride=s)
        want = O.depthwise_dgrad_strided(dy.float().numpy(), w.numpy(), T, s, d, pad, None if dl is None else dl.numpy())
        scale = np.abs(want).max() + 1e-6
        np.testing.assert_allclose(dx.float().cpu().numpy(), want, rtol=2.0 ** -7, atol=2.0 ** -9 * scale)
    if s == 1:                                                                   # the stride-1 kernel and the strided one agree
        # device tensors are bound to names: a raw pointer taken from a `.cuda()` temporary dangles as soon as the temporary dies
        dy_c, w_c, lens_c = dy.cuda(), w.cuda(), lens.cuda()
        a = F.depthwise_dgrad(dy_c, w_c, T, k, d, pad, lens_c, stride=1)
        lib = F._lib.load()
        b = torch.empty_like(a)
        F._lib.check(lib.w2l_depthwise_dgrad_strided(F._ptr(dy_c), F._ptr(w_c), F._ptr(b), B, T, C, T_out, k, 1, d, pad,
                                                     F._ptr(lens_c), F._stream()), "depthwise_dgrad_strided")
        torch.cuda.synchronize()
        # same sums, taps visited in the opposite order: equal up to one bf16 rounding of an fp32 sum
        torch.testing.assert_close(a.float(), b.float(), rtol=2.0 ** -7, atol=2.0 ** -9 * float(a.float().abs().max()))


def test_unfold_fn_gradient(F):
    """UnfoldTmFn under autograd: gradient of sum(col * g) with respect to the rows equals the fold of g"""
    from wav2letter_pytorch_b200.layers import UnfoldTmFn
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(2, 45, 64, generator=gen).to(torch.bfloat16).cuda().requires_grad_(True)
    col = UnfoldTmFn.apply(x, 21, 5, 2, 1, 0)
    g = torch.randn(col.shape, generator=gen).to(torch.bfloat16)
    col.backward(g.cuda())
    want = O.col2im_tm(g.float().numpy(), 45, 64, 5, 2, 1, 0)
    np.testing.assert_allclose(x.grad.float().cpu().numpy(), want, rtol=2.0 ** -8, atol=1e-5)


def test_w2l_strided_golden(pkg, golden):
    """Wav2Letter with TWO stride-2 blocks (reflection halo + unfold inside the stack; output lengths // 4)"""
    from test_gpu_models import check_w2l_golden
    check_w2l_golden(pkg, golden("w2l_strided"))


def test_w2l_narrow_golden(pkg, golden):
    """Wav2Letter with hidden widths 72 / 136 / 72 (multiples of 8, not of 16) against the unmodified reference"""
    from test_gpu_models import check_w2l_golden
    check_w2l_golden(pkg, golden("w2l_narrow"))


def test_jasper_strided_golden(pkg, golden):
    """Jasper with a strided dense block (repeat 2: both repeats stride) and a strided separable block after the prologue"""
    from test_gpu_models import check_jasper_golden
    check_jasper_golden(pkg, golden("jasper_strided"), seed=10)


def test_strided_block_with_residual_raises(pkg):
    """the reference's `out + res_out` fails on the time dimension for a strided block with a residual (jasper.py:412)"""
    from wav2letter_pytorch_b200 import config
    from wav2letter_pytorch_b200.jasper import Jasper
    blocks = [dict(layer_size=64, kernel_size=5, stride=1, residual=False, separable=False, repeat=1, dropout=0),
              dict(layer_size=64, kernel_size=5, stride=2, residual=True, separable=False, repeat=1, dropout=0)]
    cfg = config.compose(overrides=["model=jasper", "model.mid_layers=2"]).model
    cfg["jasper_blocks"] = config.to_attr(blocks)
    model = Jasper(cfg).cuda().train()
    x = torch.randn(2, 64, 50, device="cuda")
    with pytest.raises(RuntimeError):
        model(x, torch.tensor([50, 40], device="cuda"))


def test_head_block_standalone(pkg):
    """Conv1dBlock(..., bn=False, activation_use=False) called on its own (wav2letter.py:40-47, 69): conv + bias, NCW in / NCW out"""
    from wav2letter_pytorch_b200.wav2letter import Conv1dBlock
    torch.manual_seed(0)
    blk = Conv1dBlock(64, 29, (1,), 1, bn=False, activation_use=False).cuda()
    x = torch.randn(2, 64, 37)
    y = blk(x.cuda())
    w, b = blk.conv1.weight.detach().cpu(), blk.conv1.bias.detach().cpu()
    want = torch.nn.functional.conv1d(x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), b)
    assert y.shape == (2, 29, 37) and y.dtype == torch.float32
    assert float((y.detach().cpu() - want).norm() / want.norm()) < 1e-5
    g = torch.randn(y.shape)
    y.backward(g.cuda())
    wref, bref = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    torch.nn.functional.conv1d(x.to(torch.bfloat16).float(), wref, bref).backward(g)
    assert float((blk.conv1.weight.grad.cpu() - wref.grad).norm() / wref.grad.norm()) < 1e-2
    assert float((blk.conv1.bias.grad.cpu() - bref.grad).norm() / bref.grad.norm()) < 1e-2


def test_conv_backward_narrow_rows(F):
    """hidden widths that are multiples of 8 but not of 16 (e.g. 72): backward-data reads dy rows of Cout columns against weights
    packed with Cout_pad columns -- the tail of the last contraction chunk comes from TMA's out-of-bounds zero fill"""
    from test_kernel_emu_gemm import _narrow_rows_case
    _narrow_rows_case(F.conv1d_dgrad_wt, F.conv1d_wgrad, F.pack_wt, "cuda")


@pytest.mark.parametrize("C", [33, 100, 128])
def test_ctc_wide_alphabet_small_lattice(F, C, monkeypatch):
    """more classes than the CTC CTA has threads (a one-warp lattice for a short target), both schedules: the serial one used to emit
    only the first blockDim.x gradient columns (found on the host emulation, tests/test_kernel_emu_ctc_decode.py)"""
    from test_gpu_kernels import _check_ctc
    g = torch.Generator().manual_seed(C)
    N, T, S = 2, 40, 5
    lp = torch.log_softmax(torch.randn(N, T, C, generator=g) * 1.5, -1)
    tg = torch.randint(1, C, (N, S), generator=g, dtype=torch.int32)
    il, tl = torch.tensor([40, 31], dtype=torch.int32), torch.tensor([5, 3], dtype=torch.int32)
    tg[1, 3:] = 0
    _check_ctc(F, lp, tg, il, tl)
    monkeypatch.setenv("W2L_CTC_DBG", "8")                      # the alpha -> beta+gradient schedule of very large lattices
    _check_ctc(F, lp, tg, il, tl)


@pytest.mark.parametrize("B,T,Cin,Cout,k,d,pl,pr", [(1, 1, 64, 1, 1, 1, 0, 0), (2, 2, 72, 8, 3, 1, 1, 1), (1, 5, 64, 130, 5, 2, 8, 0),
                                                    (3, 17, 136, 29, 7, 1, 0, 6), (2, 129, 80, 272, 2, 3, 3, 0)])
def test_conv_small_and_ragged_shapes(F, B, T, Cin, Cout, k, d, pl, pr):
    """shapes at the edges of the tiling (one output row, one output channel, partial channel chunks, asymmetric zero padding) drawn
    from the host-emulation fuzzer (tests/fuzz_kernels.py), against torch conv1d autograd on the same bf16-rounded operands"""
    import torch.nn.functional as TF
    from test_gpu_kernels import _bf, _pack_w, rel_l2
    g = torch.Generator().manual_seed(B * T + Cin + k)
    T_out = T + pl + pr - d * (k - 1)
    x = _bf(torch.randn(B, T, Cin, generator=g))
    w = _bf(torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5)
    xr, wr = x.transpose(1, 2).clone().requires_grad_(True), w.clone().requires_grad_(True)
    y_ref = TF.conv1d(TF.pad(xr, (pl, pr)), wr, dilation=d)
    dy = _bf(torch.randn(B, Cout, T_out, generator=g))
    y_ref.backward(dy)
    cout_pad, ldy, cin_pad = max(64, (Cout + 15) // 16 * 16), (Cout + 7) // 8 * 8, (Cin + 15) // 16 * 16
    xc, wc = x.to(torch.bfloat16).cuda(), _pack_w(w, cout_pad).cuda()
    y = torch.zeros(B, T_out, ldy, dtype=torch.float32, device="cuda")
    F.conv1d_fwd(xc, wc, F.make_desc(B, T_out, Cin, Cout, cout_pad, k, d, T, -pl, T_out, 0, ldy, F.DT_F32, F.ACT_NONE), y)
    assert rel_l2(y[:, :, :Cout].cpu(), y_ref.detach().transpose(1, 2)) < 2e-5
    dyc = torch.zeros(B, T_out, cout_pad, dtype=torch.bfloat16, device="cuda")
    dyc[:, :, :Cout] = dy.transpose(1, 2).to(torch.bfloat16).cuda()
    desc = F.make_desc(B, T_out, Cin, Cout, cout_pad, k, d, T, -pl, T_out, 0, cout_pad)
    wt = torch.full((k, cin_pad, cout_pad), 9.0, dtype=torch.bfloat16, device="cuda")
    F.pack_wt(w.permute(2, 0, 1).contiguous().cuda(), wt, Cout, Cin)
    dx = torch.empty(B, T, Cin, dtype=torch.bfloat16, device="cuda")
    F.conv1d_dgrad_wt(dyc, wt, desc, dx)
    assert rel_l2(dx.float().cpu(), xr.grad.transpose(1, 2)) < 8e-3
    dw = torch.full((k, Cout, Cin), 3.0, dtype=torch.float32, device="cuda")
    F.conv1d_wgrad(dyc, xc, desc, dw)
    assert rel_l2(dw.cpu(), wr.grad.permute(2, 0, 1)) < 2e-5
