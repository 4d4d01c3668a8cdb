"""The memory-bound companions of the conv kernels (csrc/elementwise.cu, csrc/depthwise.cu) executed on the host from their
source text through tests/_emu_backend.py: the bodies of the corresponding `-m gpu` tests in tests/test_gpu_kernels.py with
CPU tensors, same references (plain torch fp32 of the same op on the same bf16-rounded operands), same tolerances."""
import numpy as np
import pytest
import torch
import torch.nn.functional as TF

import _emu_backend as E
import _kernel_emu as KE
from oracle import w2l_oracle as O

pytestmark = pytest.mark.skipif(not KE.available(), reason="needs g++ and the CUDA headers")

PAD_ZERO, PAD_REFLECT = 0, 1


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _bf(x):
    return x.to(torch.bfloat16).float()


def test_im2col_and_transposes_source():
    g = torch.Generator().manual_seed(4)
    B, Fdim, T = 3, 64, 201
    x = torch.randn(B, Fdim, T, generator=g)
    pl, pr = O.reflect_pad_amounts(64, 11, 2, 1)
    rows = (T + pl + pr - 11) // 2 + 1
    got = E.im2col_ncw(x, rows, 11, 2, 1, pl, PAD_REFLECT).float()
    xp = TF.pad(x, (pl, pr), mode="reflect")
    want = xp.unfold(2, 11, 2).permute(0, 2, 3, 1).reshape(B, rows, 11 * Fdim)      # [B, rows, j*F + f]
    assert torch.equal(got, _bf(want))
    lens = torch.tensor([201, 100, 7], dtype=torch.int32)                           # zero pad + length mask
    got = E.im2col_ncw(x, T + 4, 1, 1, 1, 2, PAD_ZERO, lens).float()
    want = TF.pad(x * (torch.arange(T)[None, None] < lens[:, None, None]), (2, 2)).transpose(1, 2)
    assert torch.equal(got, _bf(want))
    back = E.tm_to_ncw(got.to(torch.bfloat16), T, Fdim, x_row_offset=2)
    assert torch.equal(back, _bf(want[:, 2:2 + T]).transpose(1, 2))
    back = E.tm_to_ncw(got, T, Fdim, x_row_offset=2)                                # fp32 time-major input
    assert torch.equal(back, _bf(want[:, 2:2 + T]).transpose(1, 2))
    x5 = torch.randn(2, 5, 40, generator=g)                                         # F not a multiple of 8: the scalar store path
    got = E.im2col_ncw(x5, 40, 3, 1, 1, 1, PAD_ZERO).float()
    want = TF.pad(x5, (1, 1)).unfold(2, 3, 1).permute(0, 2, 3, 1).reshape(2, 40, 15)
    assert torch.equal(got, _bf(want))


@pytest.mark.parametrize("act,drop", [(2, 0.0), (1, 0.0), (2, 0.25), (0, 0.0)])
def test_bn_act_forward_backward_source(act, drop):
    g = torch.Generator().manual_seed(5)
    B, T, C, pl, pr = 3, 90, 264, 4, 5
    z = _bf(torch.randn(B, T, C, generator=g) * 2 + 0.5)
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    rm, rv = torch.zeros(C), torch.ones(C)
    zc = z.to(torch.bfloat16)
    stats = E.bn_stats(zc, C)
    np.testing.assert_allclose(stats[:C].numpy(), z.sum((0, 1)).numpy(), rtol=1e-4, atol=1e-2)
    rmc, rvc, nbt = rm.clone(), rv.clone(), torch.tensor(0, dtype=torch.long)
    fin = E.bn_finalize(stats, B * T, C, gamma, beta, None, 1e-3, 0.9, rmc, rvc, nbt)
    assert int(nbt) == 1
    scale, shift, mean, invstd = fin[0], fin[1], fin[2], fin[3]
    zr = z.transpose(1, 2).clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    bn = TF.batch_norm(zr, rm, rv, gr, br, training=True, momentum=0.9, eps=1e-3)
    np.testing.assert_allclose(rmc.numpy(), rm.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(rvc.numpy(), rv.numpy(), rtol=1e-4, atol=1e-5)
    seed = 1234
    if drop > 0:   # export the kernel's own keep-mask: z=0, scale=0, shift=1, no act
        ones = E.bn_act_pad(torch.zeros_like(zc), torch.zeros(C), torch.ones(C), B, T, C, 0, 0, 0, drop, seed)
        mult = ones.float().transpose(1, 2)
        keep = (mult > 0).float().mean().item()
        assert abs(keep - (1 - drop)) < 0.02
        np.testing.assert_allclose(mult[mult > 0].numpy(), 1 / (1 - drop), rtol=1e-2)
        bn = bn * mult
    y_ref = torch.clamp(bn, 0, 20) if act == 2 else (torch.relu(bn) if act == 1 else bn)
    yp_ref = TF.pad(y_ref, (pl, pr), mode="reflect")
    yp = E.bn_act_pad(zc, scale, shift, B, T, C, pl, pr, act, drop, seed)
    assert not torch.isnan(yp.float()).any()
    assert rel_l2(yp.float(), yp_ref.detach().transpose(1, 2)) < 4e-3
    dyp = _bf(torch.randn(B, C, T + pl + pr, generator=g))
    yp_ref.backward(dyp)
    dypc = dyp.transpose(1, 2).to(torch.bfloat16).contiguous()
    dz, red, _ = E.bn_act_bwd(dypc, zc, scale, shift, mean, invstd, gamma, B, T, C, pl, pr, act, drop, seed)
    assert rel_l2(dz.float(), zr.grad.transpose(1, 2)) < 1e-2
    assert rel_l2(red[:C], br.grad) < 5e-3 and rel_l2(red[C:], gr.grad) < 5e-3
    if drop > 0:   # keep-bits stored by the forward pass and read back by the backward passes == re-derived Philox bits
        bits = torch.zeros(B * T * C // 8, dtype=torch.uint8)
        yp2 = E.bn_act_pad(zc, scale, shift, B, T, C, pl, pr, act, drop, seed, drop_mask=bits)
        assert torch.equal(yp2, yp)
        dz2, red2, _ = E.bn_act_bwd(dypc, zc, scale, shift, mean, invstd, gamma, B, T, C, pl, pr, act, drop, seed, drop_mask=bits)
        assert torch.equal(dz2, dz) and torch.equal(red2, red)              # the emulation runs in one fixed order: bit-equal
    dzp, _, _ = E.bn_act_bwd(dypc, zc, scale, shift, mean, invstd, gamma, B, T, C, pl, pr, act, drop, seed, dz_rows=T + 7)
    assert torch.equal(dzp[:, :T], dz) and (dzp[:, T:] == 0).all()


def test_bn_act_residual_and_length_mask_source():
    """the Jasper form: residual branch added before the activation (jasper.py:400-412), rows >= lens written / differentiated as 0"""
    g = torch.Generator().manual_seed(8)
    B, T, C = 3, 70, 72
    z, zres = _bf(torch.randn(B, T, C, generator=g)), _bf(torch.randn(B, T, C, generator=g))
    scale, shift = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    rsc, rsh = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    lens = torch.tensor([70, 33, 1], dtype=torch.int32)
    yp = E.bn_act_pad(z.to(torch.bfloat16), scale, shift, B, T, C, 0, 0, 1, 0.0, 0, lens, res=zres.to(torch.bfloat16), res_scale=rsc,
                      res_shift=rsh)
    live = (torch.arange(T)[None, :, None] < lens[:, None, None])
    want = torch.relu(z * scale + shift + zres * rsc + rsh) * live
    assert rel_l2(yp.float(), want) < 4e-3 and (yp.float()[~live.expand_as(want)] == 0).all()
    # backward with the statistics of z: dz against autograd of the same expression, g (gradient at the BN outputs) returned
    mean, var = z.mean((0, 1)), z.var((0, 1), unbiased=False)
    invstd = torch.rsqrt(var + 1e-3)
    gamma = torch.rand(C, generator=g) + 0.5
    beta = torch.randn(C, generator=g) * 0.1
    sc, sh = gamma * invstd, beta - mean * gamma * invstd
    zr = z.clone().requires_grad_(True)
    m2, v2 = zr.mean((0, 1)), zr.var((0, 1), unbiased=False)
    out = torch.relu((zr - m2) * torch.rsqrt(v2 + 1e-3) * gamma + beta + zres * rsc + rsh) * live
    dy = _bf(torch.randn(B, T, C, generator=g))
    out.backward(dy)
    dz, red, gout = E.bn_act_bwd(dy.to(torch.bfloat16), z.to(torch.bfloat16), sc, sh, mean, invstd, gamma, B, T, C, 0, 0, 1, 0.0, 0, lens,
                                 res=zres.to(torch.bfloat16), res_scale=rsc, res_shift=rsh, want_g=True)
    assert rel_l2(dz.float(), zr.grad) < 1e-2
    pre = (z - mean) * invstd * gamma + beta + zres * rsc + rsh
    assert torch.equal(gout.float(), _bf(dy * (pre > 0) * live))


def test_log_softmax_and_colsum_source():
    g = torch.Generator().manual_seed(6)
    logits = torch.randn(4, 75, 32, generator=g) * 3
    lp = E.log_softmax(logits, 29)
    assert rel_l2(lp, torch.log_softmax(logits[..., :29], -1)) < 1e-6
    flag = torch.zeros(1, dtype=torch.int32)
    sm = E.log_softmax(logits, 29, mode=1, nan_flag=flag)
    assert rel_l2(sm, torch.softmax(logits[..., :29], -1)) < 1e-6 and int(flag) == 0
    bad = logits.clone()
    bad[2, 40, 3] = float("nan")
    E.log_softmax(bad, 29, mode=1, nan_flag=flag)
    assert int(flag) == 1                                                        # jasper.py:474 from a device flag
    gr = torch.randn(4, 75, 29, generator=g)
    x = logits[..., :29].clone().requires_grad_(True)
    torch.log_softmax(x, -1).backward(gr)
    dl = E.log_softmax_bwd(gr, lp, 64)
    assert rel_l2(dl[..., :29].float(), x.grad) < 5e-3 and (dl[..., 29:] == 0).all()
    sc = torch.tensor([0.25])
    dl2 = E.log_softmax_bwd(gr, None, 64, gscale=sc, fused_identity=True)       # CTC hands over d/dlogits: identity, scaled
    assert torch.equal(dl2[..., :29].float(), _bf(gr * 0.25)) and (dl2[..., 29:] == 0).all()
    m = _bf(torch.randn(1000, 64, generator=g))
    cs = E.colsum(m.to(torch.bfloat16), 29)
    np.testing.assert_allclose(cs.numpy(), m[:, :29].sum(0).numpy(), rtol=1e-4, atol=1e-3)
    w = torch.randn(1001, generator=g)
    assert torch.equal(E.cast_bf16(w), w.to(torch.bfloat16))


def test_reflect_halo_pack_wt_lens_chain_source():
    g = torch.Generator().manual_seed(7)
    B, T, C, pl, pr = 2, 30, 24, 5, 6
    core = _bf(torch.randn(B, T, C, generator=g))
    y = torch.full((B, pl + T + pr, C), float("nan"), dtype=torch.bfloat16)
    y[:, pl:pl + T] = core.to(torch.bfloat16)
    E.reflect_halo(y, T, pl, pr)
    assert torch.equal(y.float(), TF.pad(core.transpose(1, 2), (pl, pr), mode="reflect").transpose(1, 2))
    k, co, ci = 3, 29, 40
    w = torch.randn(k, co, ci, generator=g)
    wt = torch.full((k, 48, 64), float("nan"), dtype=torch.bfloat16)
    E.pack_wt(w, wt, co, ci)
    want = torch.zeros(k, 48, 64)
    want[:, :ci, :co] = w.flip(0).transpose(1, 2)
    assert torch.equal(wt.float(), _bf(want))
    # lens_chain: the reference's per-conv tensor arithmetic (jasper.py:91-95,107-119)
    lens0 = torch.randint(1, 3000, (37,), generator=g, dtype=torch.int64)
    chain = [(11, 2, 1, 5), (11, 1, 1, 5), (13, 1, 1, 6), (29, 1, 2, 28), (1, 1, 1, 0), (33, 2, 1, 16), (7, 3, 1, 3), (5, 0, 1, 2), (9, 1, 1, 4)]
    want_rows, lens = [lens0.clone()], lens0
    for kk, s, d, p in chain:
        if s != 0:
            lens = lens.to(dtype=torch.long)
            lens = (lens + 2 * p - d * (kk - 1) - 1) / s + 1
        want_rows.append(lens.to(dtype=torch.long))
    for dt in (torch.int64, torch.int32):
        rows, final = E.lens_chain(lens0.to(dt), chain)
        assert torch.equal(rows.long(), torch.stack(want_rows)) and torch.equal(final, want_rows[-1])


@pytest.mark.parametrize("B,T,C,k,s,d", [(2, 120, 64, 33, 1, 1), (3, 75, 72, 11, 2, 1), (2, 60, 128, 5, 1, 2), (2, 64, 8, 3, 3, 1)])
def test_depthwise_conv_source(B, T, C, k, s, d):
    """depthwise (groups=C) conv fwd / dgrad (both kernels) / wgrad vs torch fp32 on the same bf16-rounded operands, length masks"""
    g = torch.Generator().manual_seed(B * T + k)
    p = (d * k) // 2 - 1 if d > 1 else k // 2                      # jasper.py:61-66
    x = _bf(torch.randn(B, T, C, generator=g))
    w = torch.randn(C, 1, k, generator=g) / k ** 0.5
    T_out = (T + 2 * p - d * (k - 1) - 1) // s + 1
    lens = torch.randint(T_out // 2, T_out + 1, (B,), generator=g, dtype=torch.int32)
    lens[0] = T_out
    xr = x.transpose(1, 2).clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    mask = (torch.arange(T_out)[None] < lens[:, None]).float()[:, None]
    y_ref = TF.conv1d(xr, wr, stride=s, padding=p, dilation=d, groups=C) * mask
    dy = _bf(torch.randn(B, C, T_out, generator=g))
    y_ref.backward(dy)
    ws = w.permute(2, 1, 0).reshape(k, C).contiguous()
    xc = x.to(torch.bfloat16)
    y = E.depthwise_fwd(xc, ws, T_out, k, s, d, p, lens)
    assert rel_l2(y.float(), y_ref.detach().transpose(1, 2)) < 6e-3
    dyc = dy.transpose(1, 2).to(torch.bfloat16).contiguous()
    dw = E.depthwise_wgrad(dyc, xc, k, s, d, p, lens)
    assert rel_l2(dw, wr.grad.permute(2, 1, 0).reshape(k, C)) < 1e-4
    dx = E.depthwise_dgrad(dyc, ws, T, k, d, p, lens, stride=s)
    assert rel_l2(dx.float(), xr.grad.transpose(1, 2)) < 6e-3
