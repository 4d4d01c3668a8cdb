"""The tcgen05 / TMEM / TMA implicit-GEMM kernel (csrc/conv_gemm.cu, SURVEY section 8 kernel 1) executed on the host from its own
source text, THROUGH ITS OWN C ABI: tests/_emu_backend.gemm() compiles the kernel template and the library's extern "C" wrappers
(descriptor checks, tile-width choice, ring / stream-K / tail-split planning, tensor-map construction) for the host, against
functional stand-ins for the inline-PTX wrappers of common.cuh (tests/kernel_emu_runtime.h):

  mbarrier   phase bit, pending arrivals, transaction bytes (init / arrive / expect_tx / complete_tx / try_wait.parity)
  TMA        cp.async.bulk.tensor 3-D / 4-D tiled loads: dense box, zero fill outside the tensor, 128-byte swizzle on the
             absolute shared-memory address, bytes credited to the barrier
  tcgen05    mma.kind::f16 reading K-major / MN-major 128B-swizzled shared-memory descriptors (start, LBO, SBO) and the
             instruction descriptor (M, N, operand majorness), fp32 accumulators in a 128-lane x 512-column TMEM, commit as an
             arrive, ld.32x32b for the epilogue warps

The warp-specialised protocol therefore runs as written: producer / MMA issuer / epilogue warps are fibers that meet only through
the barriers.  What a passing run shows: the descriptor arithmetic, tap / chunk / tile iteration, pipeline phases, epilogues
(bias, BatchNorm fold, clamp, BN statistics, fp32 atomics of the split tiles, the K-split tail) and the host planning are
mutually consistent and compute the convolution.  What it cannot show: that the hardware agrees with this model of it -- that is
the job of the `-m gpu` tests, which these cases mirror (tests/test_gpu_kernels.py::test_conv_fwd_dgrad_wgrad)."""
import pytest
import torch
import torch.nn.functional as TF

import _emu_backend as E
import _kernel_emu as KE
from wav2letter_pytorch_b200._lib import ConvDesc

pytestmark = pytest.mark.skipif(not KE.available(), reason="needs g++ and the CUDA headers")

DT_BF16, DT_F32, ACT_NONE, ACT_CLAMP20 = 0, 1, 0, 2


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _bf(x):
    return x.to(torch.bfloat16).float()


def _pack_w(w, cout_pad):
    """[Co, Ci, k] fp32 -> packed [k, Co_pad, Ci] bf16 (zero rows in the pad)."""
    co, ci, k = w.shape
    p = torch.zeros(k, cout_pad, ci, dtype=torch.bfloat16)
    p[:, :co] = w.permute(2, 0, 1).to(torch.bfloat16)
    return p


def make_desc(B, T_out, Cin, Cout, Cout_pad, k, dilation, x_rows, x_row_offset, y_rows, y_row_offset, ldy, y_dtype=DT_BF16, act=ACT_NONE):
    return ConvDesc(B, T_out, Cin, Cout, Cout_pad, k, dilation, x_rows, x_row_offset, y_rows, y_row_offset, ldy, y_dtype, act)


CONV_CASES = [
    # B, T, Cin, Cout, k, d, pad(left,right: rows of zero padding)
    (2, 150, 64, 256, 1, 1, (0, 0)),
    (2, 140, 128, 224, 5, 1, (2, 2)),
    (2, 131, 64, 160, 11, 1, (5, 5)),
    (2, 260, 192, 29, 1, 1, (0, 0)),
    (2, 150, 64, 96, 7, 2, (6, 6)),
    (1, 97, 256, 384, 3, 1, (1, 1)),           # N = 384: two N tiles
    (1, 130, 72, 40, 3, 1, (1, 1)),            # Cin, Cout not multiples of 64 / 16: partial K chunk, 3-D (not chunked) wgrad maps
]


@pytest.mark.parametrize("B,T,Cin,Cout,k,d,pad", CONV_CASES)
def test_conv_fwd_dgrad_wgrad_source(B, T, Cin, Cout, k, d, pad):
    g = torch.Generator().manual_seed(B * T + Cin + k)
    pl, pr = pad
    x = _bf(torch.randn(B, T, Cin, generator=g))                 # time-major, UNpadded: zero padding via TMA OOB fill
    w = _bf(torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5)
    bias = torch.randn(Cout, generator=g)
    T_out = T + pl + pr - d * (k - 1)
    cout_pad = max(64, (Cout + 15) // 16 * 16)
    ldy = (Cout + 7) // 8 * 8
    xr = x.transpose(1, 2).clone().requires_grad_(True)          # NCW
    wr = w.clone().requires_grad_(True)
    y_ref = TF.conv1d(TF.pad(xr, (pl, pr)), wr, bias, dilation=d)
    dy = _bf(torch.randn(B, Cout, T_out, generator=g))
    y_ref.backward(dy)
    # ---- forward, fp32 output + bias
    xc, wc = x.to(torch.bfloat16), _pack_w(w, cout_pad)
    desc = make_desc(B, T_out, Cin, Cout, cout_pad, k, d, T, -pl, T_out, 0, ldy, DT_F32, ACT_NONE)
    y = torch.zeros(B, T_out, ldy)
    E.conv1d_fwd(xc, wc, desc, y, bias=bias)
    want = y_ref.detach().transpose(1, 2)
    assert rel_l2(y[:, :, :Cout], want) < 2e-5                   # fp32 accumulate of identical bf16 operands
    # ---- bf16 output + fused scale/shift + clamp epilogue, rows outside [off, off+T_out) untouched
    desc2 = make_desc(B, T_out, Cin, Cout, cout_pad, k, d, T, -pl, T_out + 3, 2, ldy, DT_BF16, ACT_CLAMP20)
    y2 = torch.full((B, T_out + 3, ldy), 7.0, dtype=torch.bfloat16)
    sc, sh = torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g)
    E.conv1d_fwd(xc, wc, desc2, y2, bias=bias, scale=sc, shift=sh)
    want2 = torch.clamp(want * sc + sh, 0, 20)
    assert rel_l2(y2[:, 2:2 + T_out, :Cout].float(), want2) < 6e-3
    assert (y2[:, :2] == 7.0).all() and (y2[:, 2 + T_out:] == 7.0).all()
    # ---- BatchNorm statistics from the epilogue: sum / sum of squares of exactly the bf16 values it stored
    desc_s = make_desc(B, T_out, Cin, Cout, cout_pad, k, d, T, -pl, T_out, 0, ldy, DT_BF16, ACT_NONE)
    ys = torch.zeros(B, T_out, ldy, dtype=torch.bfloat16)
    st = torch.zeros(2 * Cout)
    E.conv1d_fwd(xc, wc, desc_s, ys, bn_stats=st)
    yd = ys[:, :, :Cout].double().reshape(-1, Cout)
    assert torch.allclose(st[:Cout].double(), yd.sum(0), rtol=1e-4, atol=1e-3 * float(yd.abs().sum(0).max()) / 100)
    assert torch.allclose(st[Cout:].double(), (yd * yd).sum(0), rtol=1e-4)
    # ---- dgrad (MN-major weight operand)
    dyc = torch.zeros(B, T_out, cout_pad, dtype=torch.bfloat16)
    dyc[:, :, :Cout] = dy.transpose(1, 2).to(torch.bfloat16)
    desc3 = make_desc(B, T_out, Cin, Cout, cout_pad, k, d, T, -pl, T_out, 0, cout_pad)
    dx = torch.full((B, T, Cin), float("nan"), dtype=torch.bfloat16)
    E.conv1d_dgrad(dyc, wc, desc3, dx)
    assert rel_l2(dx.float(), xr.grad.transpose(1, 2)) < 6e-3
    # ---- dgrad through the transposed (K-major) weight shadow made by pack_wt_kernel
    cin_pad = (Cin + 15) // 16 * 16
    wt = torch.full((k, cin_pad, cout_pad), 9.0, dtype=torch.bfloat16)
    E.pack_wt(w.permute(2, 0, 1).contiguous(), wt, Cout, Cin)
    dx2 = torch.full((B, T, Cin), float("nan"), dtype=torch.bfloat16)
    E.conv1d_dgrad_wt(dyc, wt, desc3, dx2)
    assert rel_l2(dx2.float(), xr.grad.transpose(1, 2)) < 6e-3
    # ---- wgrad (both operands MN-major, stream-K with fp32 atomics on shared tiles)
    dw = torch.full((k, Cout, Cin), 3.0)
    E.conv1d_wgrad(dyc, xc, desc3, dw)
    assert rel_l2(dw, wr.grad.permute(2, 0, 1)) < 2e-5


def _narrow_rows_case(conv1d_dgrad_wt, conv1d_wgrad, pack_wt, dev):
    """hidden width 72 (a multiple of 8, not of 16): dy rows carry 72 columns while the weights are packed with Cout_pad = 80"""
    g = torch.Generator().manual_seed(72)
    B, T, Cin, Cout, k, d, pad = 2, 90, 136, 72, 5, 1, 2
    x = _bf(torch.randn(B, T, Cin, generator=g))
    w = _bf(torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5)
    xr, wr = x.transpose(1, 2).clone().requires_grad_(True), w.clone().requires_grad_(True)
    y = TF.conv1d(TF.pad(xr, (pad, pad)), wr, dilation=d)
    dy = _bf(torch.randn(B, Cout, T, generator=g))
    y.backward(dy)
    cout_pad, cin_pad = 80, 144
    dyc = dy.transpose(1, 2).to(torch.bfloat16).contiguous().to(dev)                 # [B, T, 72]: row pitch == Cout < Cout_pad
    desc = make_desc(B, T, Cin, Cout, cout_pad, k, d, T, -pad, T, 0, Cout)
    wt = torch.full((k, cin_pad, cout_pad), 9.0, dtype=torch.bfloat16, device=dev)
    pack_wt(w.permute(2, 0, 1).contiguous().to(dev), wt, Cout, Cin)
    dx = torch.full((B, T, Cin), float("nan"), dtype=torch.bfloat16, device=dev)
    conv1d_dgrad_wt(dyc, wt, desc, dx)
    assert rel_l2(dx.float().cpu(), xr.grad.transpose(1, 2)) < 6e-3
    dw = torch.full((k, Cout, Cin), 3.0, device=dev)
    conv1d_wgrad(dyc, x.to(torch.bfloat16).to(dev), desc, dw)
    assert rel_l2(dw.cpu(), wr.grad.permute(2, 0, 1)) < 2e-5


def test_conv_backward_narrow_rows_source():
    _narrow_rows_case(E.conv1d_dgrad_wt, E.conv1d_wgrad, E.pack_wt, "cpu")


def test_conv_fwd_tail_split_source():
    """the forward tail split (last wave of tiles cut along K, pieces summed in an fp32 scratch by the last arriver), forced on a
    small problem by capping the persistent grid at 8 CTAs: 18 tiles = 2 full waves + 2 tiles -> 4 K-pieces each; plain store,
    fused affine + clamp + BatchNorm statistics, twice (the arrival counters clean themselves)"""
    import ctypes
    lib = E.gemm().lib
    g = torch.Generator().manual_seed(12)
    B, T, Cin, Cout, k = 9, 200, 192, 256, 11
    pad = k // 2
    x = _bf(torch.randn(B, T, Cin, generator=g))
    w = _bf(torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5)
    want = TF.conv1d(x.transpose(1, 2), w, padding=pad).transpose(1, 2)
    xc, wc = x.to(torch.bfloat16), _pack_w(w, Cout)
    E.ensure_gemm_scratch()
    try:
        lib.emu_set_sm_budget(8)
        desc = make_desc(B, T, Cin, Cout, Cout, k, 1, T, -pad, T, 0, Cout, DT_F32, ACT_NONE)
        assert lib.w2l_conv1d_fwd_tail_parts(ctypes.byref(desc)) == 4
        for rep in range(2):
            y = torch.zeros(B, T, Cout)
            E.conv1d_fwd(xc, wc, desc, y)
            assert rel_l2(y, want) < 2e-5
        sc, sh = torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g)
        desc2 = make_desc(B, T, Cin, Cout, Cout, k, 1, T, -pad, T, 0, Cout, DT_BF16, ACT_CLAMP20)
        y2 = torch.zeros(B, T, Cout, dtype=torch.bfloat16)
        st = torch.zeros(2 * Cout)
        E.conv1d_fwd(xc, wc, desc2, y2, scale=sc, shift=sh, bn_stats=st)
        assert rel_l2(y2.float(), torch.clamp(want * sc + sh, 0, 20)) < 6e-3
        yd = y2.double().reshape(-1, Cout)
        assert torch.allclose(st[:Cout].double(), yd.sum(0), rtol=1e-4, atol=1e-2) and torch.allclose(st[Cout:].double(), (yd * yd).sum(0), rtol=1e-4)
    finally:
        lib.emu_set_sm_budget(0)
    assert lib.w2l_conv1d_fwd_tail_parts(ctypes.byref(desc)) <= 1          # 18 tiles fit one wave of the full machine


def test_conv_slab_mode_source():
    """the opt-in resident-slab forward path (W2L_SLAB=1, read once per process): tap j's A operand is a descriptor that starts j*d
    rows into a slab fetched once per channel chunk -- the conv cases above in a fresh interpreter"""
    import os
    import subprocess
    import sys
    env = dict(os.environ, W2L_SLAB="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-k", "conv_fwd_dgrad_wgrad_source or tail_split",
                        "-p", "no:cacheprovider"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
