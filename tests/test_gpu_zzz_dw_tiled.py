"""GPU tests of the register-tiled depthwise kernels (csrc/depthwise.cu, W2L_DW_TILED=1).  EXPERIMENTAL and opt-in like the code path
itself (written after round 1's GPU budget was spent; on the host they run in tests/test_kernel_emu_depthwise_tiled.py and under
`pytest -m gpu --emulate-gpu`): on a real GPU only with W2L_TEST_EXPERIMENTAL=1 (tools/final_run.sh sets it).  Sorts last."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.fixture
def F(request):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    if not (os.environ.get("W2L_TEST_EXPERIMENTAL") or request.config.getoption("--emulate-gpu")):
        pytest.skip("experimental path: set W2L_TEST_EXPERIMENTAL=1 to run it on the GPU")
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


def test_dw_tiled_jasper_separable_golden(request, monkeypatch, golden):
    """whole separable Jasper (fixture frozen from the unmodified reference) with the tiled kernels in place"""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    if not (os.environ.get("W2L_TEST_EXPERIMENTAL") or request.config.getoption("--emulate-gpu")):
        pytest.skip("experimental path: set W2L_TEST_EXPERIMENTAL=1 to run it on the GPU")
    import wav2letter_pytorch_b200 as pkg
    from test_gpu_models import check_jasper_golden
    monkeypatch.setenv("W2L_DW_TILED", "1")
    check_jasper_golden(pkg, golden("jasper_small"), seed=2)
