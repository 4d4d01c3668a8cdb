"""The register-tiled depthwise kernels (csrc/depthwise.cu ``depthwise_corr_tiled_kernel`` / ``depthwise_wgrad_tiled_kernel``, opt-in
with W2L_DW_TILED=1) executed on the host from their own source through the library's C wrappers (tests/_emu_cabi.py): forward and
backward-data must be BIT-IDENTICAL to the default kernels (same taps in the same order), the weight gradient equal up to fp32
summation order; all three against torch's grouped conv1d autograd on the same bf16-rounded operands, with length masks."""
import pytest
import torch
import torch.nn.functional as TF

import _emu_cabi
import _kernel_emu as KE

pytestmark = pytest.mark.skipif(not KE.available(), reason="needs g++ and the CUDA headers")


def _bf(x):
    return x.to(torch.bfloat16).float()


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


# (B, T, C, k): rows not a multiple of the 8-row tile, one-tile utterances, k below / at / above / far above the tile, even k
# (the shipped jasper.yaml uses k = 32 ... 74, jasper.py:61-66 pads k // 2), channel counts that leave a partly filled block
@pytest.mark.parametrize("B,T,C,k", [(2, 120, 64, 33), (3, 75, 72, 11), (2, 7, 8, 3), (1, 64, 264, 8), (2, 50, 16, 32), (2, 41, 32, 74),
                                     (3, 9, 8, 1)])
def test_tiled_depthwise_equals_default_kernels(monkeypatch, B, T, C, k):
    F = _emu_cabi.install(monkeypatch)
    g = torch.Generator().manual_seed(B * T + k)
    p = k // 2
    x = _bf(torch.randn(B, T, C, generator=g))
    w = torch.randn(C, 1, k, generator=g) / k ** 0.5
    T_out = T + 2 * p - (k - 1)
    lens = torch.randint(max(1, T_out // 2), T_out + 1, (B,), generator=g, dtype=torch.int32)
    lens[0] = T_out
    xr = x.transpose(1, 2).clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    mask = (torch.arange(T_out)[None] < lens[:, None]).float()[:, None]
    y_ref = TF.conv1d(xr, wr, padding=p, groups=C) * mask
    dy = _bf(torch.randn(B, C, T_out, generator=g))
    y_ref.backward(dy)
    ws = w.permute(2, 1, 0).reshape(k, C).contiguous()
    xc = x.to(torch.bfloat16)
    dyc = dy.transpose(1, 2).to(torch.bfloat16).contiguous()

    def run():
        n0 = F._lib.load().w2l_launch_count()
        y = F.depthwise_fwd(xc, ws, T_out, k, 1, 1, p, lens)
        dw = F.depthwise_wgrad(dyc, xc, k, 1, 1, p, lens)
        dx = F.depthwise_dgrad(dyc, ws, T, k, 1, p, lens)
        dx_nolens = F.depthwise_dgrad(dyc, ws, T, k, 1, p, None)
        y_nolens = F.depthwise_fwd(xc, ws, T_out, k, 1, 1, p, None)
        assert F._lib.load().w2l_launch_count() - n0 == 5
        return y, dw, dx, dx_nolens, y_nolens

    monkeypatch.setenv("W2L_DW_TILED", "0")
    y0, dw0, dx0, dxn0, yn0 = run()
    monkeypatch.setenv("W2L_DW_TILED", "1")
    y1, dw1, dx1, dxn1, yn1 = run()
    assert not torch.isnan(y1.float()).any() and not torch.isnan(dx1.float()).any() and not torch.isnan(dw1).any()
    assert torch.equal(y0, y1) and torch.equal(dx0, dx1) and torch.equal(dxn0, dxn1) and torch.equal(yn0, yn1)      # bit-identical
    assert rel_l2(dw1, dw0) < 1e-5
    assert rel_l2(y1.float(), y_ref.detach().transpose(1, 2)) < 6e-3
    assert rel_l2(dw1, wr.grad.permute(2, 1, 0).reshape(k, C)) < 1e-4
    assert rel_l2(dx1.float(), xr.grad.transpose(1, 2)) < 6e-3
    assert (y1[B - 1, int(lens[B - 1]):] == 0).all()


def test_tiled_kernels_leave_strided_and_dilated_layers_to_the_default_ones(monkeypatch):
    """stride 2 (the prologue block of jasper.yaml) and dilation 2 do not slide by one row per tap: same results with the switch set"""
    F = _emu_cabi.install(monkeypatch)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 60, 16, generator=g).to(torch.bfloat16)
    ws = torch.randn(5, 16, generator=g)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("W2L_DW_TILED", mode)
        out[mode] = (F.depthwise_fwd(x, ws, 30, 5, 2, 1, 2, None), F.depthwise_fwd(x, ws, 60, 5, 1, 2, 4, None))
    assert torch.equal(out["0"][0], out["1"][0]) and torch.equal(out["0"][1], out["1"][1])
