"""The source text of the thread-independent kernels, executed on the host (tests/_kernel_emu.py) and held to the oracle on
the same cases, and the same tolerances, as their `-m gpu` tests in tests/test_gpu_zz_strided.py / test_gpu_kernels.py.

Why: the unfold / fold kernels and the strided depthwise backward-data kernel were written when no GPU session was left in
round 1.  The emulation executes their index arithmetic, masks, tap loops and bf16 roundings exactly as written in the .cu
files; depthwise_corr_kernel (verified on a B200 in both of its roles) rides along as the control that the emulation itself
reproduces a kernel known to be right on hardware."""
import numpy as np
import pytest
import torch

import _emu_backend as E
import _kernel_emu as KE
from oracle import w2l_oracle as O

pytestmark = pytest.mark.skipif(not KE.available(), reason="needs g++ and the CUDA headers")


class _Both:
    """the two builds behind one launch(): elementwise.cu and depthwise.cu are separate translation units in the library too"""

    def __init__(self):
        self.parts = [KE.build(["elementwise.cu"], ["im2col_tm_kernel", "col2im_tm_kernel"], drop=E.ELEMENTWISE_DROP, extra=E.ELEMENTWISE_PTX),
                      KE.build(["depthwise.cu"], ["depthwise_corr_kernel<__nv_bfloat16>", "depthwise_dgrad_strided_kernel<__nv_bfloat16>", "depthwise_wgrad_kernel<__nv_bfloat16>"])]

    def launch(self, kernel, *a, **k):
        next(p for p in self.parts if kernel in p._sigs).launch(kernel, *a, **k)


@pytest.fixture(scope="module")
def emu():
    return _Both()


def grid_for(items, threads, sms=148):
    """csrc/elementwise.cu grid_for: a grid-stride grid capped at 16 blocks per SM"""
    return int(max(1, min((items + threads - 1) // threads, sms * 16)))


def bf16(t):
    return t.to(torch.bfloat16).contiguous()


def emu_im2col_tm(emu, x, T_out, k, s, d, pad):
    B, rows, C = x.shape
    out = torch.full((B, T_out, k * C), float("nan"), dtype=torch.bfloat16)
    emu.launch("im2col_tm_kernel", grid_for(B * T_out * k * (C // 8), 256), 256, x.data_ptr(), out.data_ptr(), B, rows, C, T_out, k, s, d, pad)
    return out


def emu_col2im_tm(emu, dcol, rows, C, k, s, d, pad):
    B, T_out, _ = dcol.shape
    dx = torch.full((B, rows, C), float("nan"), dtype=torch.bfloat16)
    emu.launch("col2im_tm_kernel", grid_for(B * rows * (C // 8), 256), 256, dcol.data_ptr(), dx.data_ptr(), B, rows, C, T_out, k, s, d, pad)
    return dx


@pytest.mark.parametrize("B,rows,C,k,s,d,pad", [(3, 109, 64, 7, 2, 1, 0), (2, 51, 72, 5, 2, 1, 2), (2, 40, 8, 5, 3, 1, 2),
                                                (1, 33, 128, 3, 1, 2, 2), (2, 17, 16, 1, 2, 1, 0), (4, 300, 256, 11, 2, 1, 5)])
def test_im2col_tm_and_fold_source(emu, B, rows, C, k, s, d, pad):
    gen = torch.Generator().manual_seed(B * 1000 + rows)
    x = bf16(torch.randn(B, rows, C, generator=gen))
    T_out = (rows + 2 * pad - d * (k - 1) - 1) // s + 1
    col = emu_im2col_tm(emu, x, T_out, k, s, d, pad)
    assert np.array_equal(col.float().numpy(), O.im2col_tm(x.float().numpy(), T_out, k, s, d, pad))      # a pure copy: bit-exact
    dcol = bf16(torch.randn(B, T_out, k * C, generator=gen))
    dx = emu_col2im_tm(emu, dcol, rows, C, k, s, d, pad)
    want = O.col2im_tm(dcol.float().numpy(), rows, C, k, s, d, pad)
    assert not torch.isnan(dx.float()).any()                                                             # every element written
    np.testing.assert_allclose(dx.float().numpy(), want, rtol=2.0 ** -8, atol=1e-5)


def test_fold_small_grid_covers_everything(emu):
    """the grid-stride loop with a grid far smaller than the item count (what the 16-blocks-per-SM cap does at full size)"""
    gen = torch.Generator().manual_seed(1)
    B, rows, C, k, s, d, pad = 2, 77, 64, 5, 2, 1, 1
    T_out = (rows + 2 * pad - d * (k - 1) - 1) // s + 1
    x = bf16(torch.randn(B, rows, C, generator=gen))
    out = torch.full((B, T_out, k * C), float("nan"), dtype=torch.bfloat16)
    emu.launch("im2col_tm_kernel", 3, 64, x.data_ptr(), out.data_ptr(), B, rows, C, T_out, k, s, d, pad)
    assert np.array_equal(out.float().numpy(), O.im2col_tm(x.float().numpy(), T_out, k, s, d, pad))


def dw_grid(C, rows):
    return ((C // 8 + 31) // 32, (rows + 7) // 8), (32, 8)          # csrc/depthwise.cu wrappers


@pytest.mark.parametrize("B,T,C,k,s,d,pad", [(2, 51, 64, 6, 2, 1, 3), (3, 40, 72, 5, 3, 1, 2), (2, 26, 128, 33, 2, 1, 16), (2, 19, 8, 3, 1, 2, 2),
                                             (2, 30, 264, 4, 2, 1, 0)])
def test_depthwise_dgrad_strided_source(emu, B, T, C, k, s, d, pad):
    gen = torch.Generator().manual_seed(T * 10 + k)
    T_out = (T + 2 * pad - d * (k - 1) - 1) // s + 1
    dy = bf16(torch.randn(B, T_out, C, generator=gen))
    w = torch.randn(k, C, generator=gen).contiguous()
    lens = torch.tensor([T_out] + [max(1, T_out - 2 - i) for i in range(B - 1)], dtype=torch.int32)
    grid, block = dw_grid(C, B * T)
    for dl in (None, lens):
        dx = torch.full((B, T, C), float("nan"), dtype=torch.bfloat16)
        emu.launch("depthwise_dgrad_strided_kernel<__nv_bfloat16>", grid, block, dy.data_ptr(), w.data_ptr(), dx.data_ptr(), B, T, T_out, C, k, s, d, pad,
                   None if dl is None else dl.data_ptr())
        want = O.depthwise_dgrad_strided(dy.float().numpy(), w.numpy(), T, s, d, pad, None if dl is None else dl.numpy())
        assert not torch.isnan(dx.float()).any()
        scale = np.abs(want).max() + 1e-6
        np.testing.assert_allclose(dx.float().numpy(), want, rtol=2.0 ** -7, atol=2.0 ** -9 * scale)
    if s == 1:
        # control: the stride-1 backward-data path of the hardware-verified correlation kernel computes the same thing
        a = torch.full((B, T, C), float("nan"), dtype=torch.bfloat16)
        emu.launch("depthwise_corr_kernel<__nv_bfloat16>", grid, block, dy.data_ptr(), w.data_ptr(), a.data_ptr(), B, T_out, T, C, k, 1, d, pad - (k - 1) * d, 1,
                   lens.data_ptr(), None)
        torch.testing.assert_close(a.float(), dx.float(), rtol=2.0 ** -7, atol=2.0 ** -9 * float(a.float().abs().max()))


@pytest.mark.parametrize("B,T,C,k,s,d,pad", [(2, 40, 64, 5, 1, 1, 2), (2, 41, 72, 6, 2, 1, 3), (1, 50, 8, 3, 1, 2, 2)])
def test_depthwise_fwd_source_is_adjoint_of_strided_dgrad(emu, B, T, C, k, s, d, pad):
    """<fwd(x), dy> == <x, dgrad(dy)> up to bf16 rounding of the two outputs: ties the new kernel to the verified forward one"""
    gen = torch.Generator().manual_seed(7 * T + k)
    T_out = (T + 2 * pad - d * (k - 1) - 1) // s + 1
    x, dy = bf16(torch.randn(B, T, C, generator=gen)), bf16(torch.randn(B, T_out, C, generator=gen))
    w = torch.randn(k, C, generator=gen).contiguous()
    y = torch.full((B, T_out, C), float("nan"), dtype=torch.bfloat16)
    grid, block = dw_grid(C, B * T_out)
    emu.launch("depthwise_corr_kernel<__nv_bfloat16>", grid, block, x.data_ptr(), w.data_ptr(), y.data_ptr(), B, T, T_out, C, k, s, d, -pad, 0, None, None)
    want = torch.nn.functional.conv1d(x.float().transpose(1, 2), w.t().reshape(C, 1, k), stride=s, padding=pad, dilation=d, groups=C).transpose(1, 2)
    torch.testing.assert_close(y.float(), want, rtol=2.0 ** -7, atol=2.0 ** -9 * float(want.abs().max()))
    dx = torch.full((B, T, C), float("nan"), dtype=torch.bfloat16)
    grid, block = dw_grid(C, B * T)
    emu.launch("depthwise_dgrad_strided_kernel<__nv_bfloat16>", grid, block, dy.data_ptr(), w.data_ptr(), dx.data_ptr(), B, T, T_out, C, k, s, d, pad, None)
    lhs = float((y.double() * dy.double()).sum())
    rhs = float((x.double() * dx.double()).sum())
    norm = float(y.double().norm() * dy.double().norm())
    assert abs(lhs - rhs) <= 2.0 ** -7 * norm


@pytest.mark.parametrize("B,T,C,k,s,d,pad", [(2, 40, 64, 5, 1, 1, 2), (3, 41, 72, 6, 2, 1, 3), (2, 300, 264, 33, 1, 1, 16)])
def test_depthwise_wgrad_source(emu, B, T, C, k, s, d, pad):
    """the cooperative one (shared-memory reduction over the 8 row-threads, barrier, atomics into dw) under the fiber scheduler"""
    gen = torch.Generator().manual_seed(T + k)
    T_out = (T + 2 * pad - d * (k - 1) - 1) // s + 1
    x, dy = bf16(torch.randn(B, T, C, generator=gen)), bf16(torch.randn(B, T_out, C, generator=gen))
    lens = torch.tensor([T_out] + [max(1, T_out - 3 - i) for i in range(B - 1)], dtype=torch.int32)
    rows = B * T_out
    rpb = max(64, (rows + 147) // 148)                                          # csrc/depthwise.cu w2l_depthwise_wgrad
    grid = ((C // 8 + 31) // 32, (rows + rpb - 1) // rpb, (k + 3) // 4)
    for dl in (None, lens):
        dw = torch.zeros(k, C)
        emu.launch("depthwise_wgrad_kernel<__nv_bfloat16>", grid, (32, 8), dy.data_ptr(), x.data_ptr(), dw.data_ptr(), B, T, T_out, C, k, s, d, pad,
                   None if dl is None else dl.data_ptr(), rpb)
        g = dy.float().clone()
        if dl is not None:
            for b in range(B):
                g[b, int(dl[b]):] = 0
        xr = x.float().transpose(1, 2).requires_grad_(False)
        w = torch.zeros(C, 1, k, requires_grad=True)
        torch.nn.functional.conv1d(xr, w, stride=s, padding=pad, dilation=d, groups=C).backward(g.transpose(1, 2))
        want = w.grad[:, 0, :].t()
        torch.testing.assert_close(dw, want, rtol=1e-4, atol=1e-4 * float(want.abs().max()))


def _probe(src, kernel):
    import os
    import tempfile
    path = os.path.join(tempfile.mkdtemp(), "probe.cu")
    open(path, "w").write("namespace w2l {\n" + src + "\n}\n")
    return KE.build([path], [kernel])


def test_runtime_warp_collectives():
    """the emulation's own shuffles / ballots: butterfly reduction, scan neighbours, partial last warp, lanes that exit right
    after the last collective (their deposit must stay readable)"""
    e = _probe("""
__global__ void collectives_kernel(const float* x, float* red, float* up, float* down, unsigned* ballot, int n) {
  const int i = threadIdx.x;
  float m = x[i];
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  red[i] = m;
  up[i] = __shfl_up_sync(0xffffffffu, x[i], 1);
  down[i] = __shfl_down_sync(0xffffffffu, x[i], 2);
  ballot[i] = __ballot_sync(0xffffffffu, x[i] > 10.f);
}""", "collectives_kernel")
    n = 80                                                               # two full warps and a 16-lane one
    x = torch.randn(n, generator=torch.Generator().manual_seed(0)) * 20
    red, up, down = torch.zeros(n), torch.zeros(n), torch.zeros(n)
    ballot = torch.zeros(n, dtype=torch.int32)
    e.launch("collectives_kernel", 1, n, x.data_ptr(), red.data_ptr(), up.data_ptr(), down.data_ptr(), ballot.data_ptr(), n)
    for w0 in (0, 32, 64):
        w1 = min(n, w0 + 32)
        assert (red[w0:w1] == x[w0:w1].max()).all()
        assert up[w0] == x[w0] and torch.equal(up[w0 + 1:w1], x[w0:w1 - 1])
        assert torch.equal(down[w0:w1 - 2], x[w0 + 2:w1]) and torch.equal(down[w1 - 2:w1], x[w1 - 2:w1])
        want = sum(1 << j for j in range(w1 - w0) if x[w0 + j] > 10)
        assert all((int(b) & 0xFFFFFFFF) == want for b in ballot[w0:w1])


def test_runtime_flags_shared_memory_overrun():
    e = _probe("__global__ void overrun_kernel(int n) {\n  extern __shared__ float buf[];\n  buf[threadIdx.x] = 1.f;\n}", "overrun_kernel")
    e.launch("overrun_kernel", 2, 32, 0, smem=32 * 4)
    with pytest.raises(KE.EmuError, match="past the end of its dynamic shared memory"):
        e.launch("overrun_kernel", 2, 32, 0, smem=31 * 4)


def test_scheduler_reports_deadlock_instead_of_hanging():
    """a barrier only half the block reaches: on hardware a hang, here an error"""
    e = _probe("__global__ void half_barrier_kernel(int* out) {\n  if (threadIdx.x < 16) { __syncthreads(); out[0] = 1; }\n"
               "  else { __syncwarp(); out[1] = 1; }\n}", "half_barrier_kernel")
    out = torch.zeros(2, dtype=torch.int32)
    with pytest.raises(KE.EmuError, match="deadlock"):
        e.launch("half_barrier_kernel", 1, 32, out.data_ptr())
