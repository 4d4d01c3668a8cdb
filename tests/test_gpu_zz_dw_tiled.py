"""GPU tests of the register-tiled depthwise kernels (csrc/depthwise.cu): the default for stride 1 / dilation 1 since round 2's A/B on
hardware (profiles/r2_dw_ab.md); W2L_DW_TILED=0 selects the plain kernels they are compared with here."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.fixture
def F():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from wav2letter_pytorch_b200 import functional
    return functional


@pytest.mark.parametrize("B,T,C,k", [(2, 120, 64, 33), (3, 75, 72, 11), (2, 7, 8, 3), (2, 50, 16, 32), (2, 41, 32, 74), (64, 751, 256, 32),
                                     (64, 751, 512, 74)])
def test_dw_tiled_equals_default(F, monkeypatch, B, T, C, k):
    """forward / backward-data bit-identical to the default kernels (same taps, same order); weight gradient up to summation order.
    The last two shapes are the first and the last separable block of the shipped jasper.yaml at B=64 x 15 s."""
    g = torch.Generator().manual_seed(B * T + k)
    p = k // 2
    T_out = T + 2 * p - (k - 1)
    x = torch.randn(B, T, C, generator=g).to(torch.bfloat16).cuda()
    dy = torch.randn(B, T_out, C, generator=g).to(torch.bfloat16).cuda()
    ws = (torch.randn(k, C, generator=g) / k ** 0.5).cuda()
    lens = torch.randint(max(1, T_out // 2), T_out + 1, (B,), generator=g, dtype=torch.int32)
    lens[0] = T_out
    lens = lens.cuda()
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("W2L_DW_TILED", mode)
        out[mode] = (F.depthwise_fwd(x, ws, T_out, k, 1, 1, p, lens), F.depthwise_dgrad(dy, ws, T, k, 1, p, lens),
                     F.depthwise_dgrad(dy, ws, T, k, 1, p, None), F.depthwise_wgrad(dy, x, k, 1, 1, p, lens))
    for a, b in zip(out["0"][:3], out["1"][:3]):
        assert torch.equal(a, b)
    dw0, dw1 = out["0"][3], out["1"][3]
    assert float((dw0 - dw1).norm() / dw0.norm()) < 1e-5


def test_plain_depthwise_jasper_separable_golden(monkeypatch, golden):
    """whole separable Jasper (fixture frozen from the unmodified reference) with the PLAIN depthwise kernels in place
    (W2L_DW_TILED=0); the default (tiled) path is what tests/test_gpu_models.py::test_jasper_dense_golden[jasper_small] runs"""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import wav2letter_pytorch_b200 as pkg
    from test_gpu_models import check_jasper_golden
    monkeypatch.setenv("W2L_DW_TILED", "0")
    check_jasper_golden(pkg, golden("jasper_small"), seed=2)
