"""TEST INFRASTRUCTURE ONLY -- a host-side stand-in for the C ABI behind ``wav2letter_pytorch_b200.functional``.

The product has no CPU path (tests/test_host_cpu.py::test_no_cpu_fallback).  To exercise the HOST logic above the C ABI on a
machine without a GPU -- autograd Functions, descriptor geometry, halo / mask / unfold wiring, length bookkeeping of
layers.py / wav2letter.py / jasper.py -- ``install(monkeypatch)`` swaps every ``functional`` entry point the models call for a
plain torch restatement of what the kernel behind it is specified to do in include/w2l_sm100.h (same layouts, same bf16
storage points, same argument meaning).  Only tests may import this module; nothing here is shipped or timed, and the `-m gpu`
suite never uses it (there the real kernels are compared with the oracle and the reference's fixtures)."""
import torch
import torch.nn.functional as TF

BF16 = torch.bfloat16


def _r(t):
    """one bf16 store"""
    return t.to(BF16)


def _act(v, act):
    if act == 1:
        return torch.where(v < 0, torch.zeros_like(v), v)
    if act == 2:
        t = torch.where(v < 0, torch.zeros_like(v), v)
        return torch.where(t > 20, torch.full_like(t, 20.0), t)
    return v


def _act_pass(pre, act):
    if act == 1:
        return pre > 0
    if act == 2:
        return (pre >= 0) & (pre <= 20)
    return torch.ones_like(pre, dtype=torch.bool)


def _row_mask(lens, B, T):
    """[B, T, 1] bool, True where t >= lens[b]"""
    if lens is None:
        return torch.zeros((B, T, 1), dtype=torch.bool)
    return (torch.arange(T).view(1, T, 1) >= lens.view(B, 1, 1).to(torch.int64))


# ------------------------------------------------------------------------------------------------ layout
def im2col_ncw(x, rows, k, stride, dilation, pad_left, pad_mode, lens=None, out_dtype=torch.bfloat16):
    assert out_dtype == torch.bfloat16, "the fp32-faithful mode is covered by the C-ABI emulation (tests/_emu_cabi.py) and the -m gpu tests"
    x = x.contiguous().float()
    B, F, T = x.shape
    t = (torch.arange(rows).view(rows, 1) * stride + torch.arange(k).view(1, k) * dilation - pad_left)      # [rows, k]
    if pad_mode == 1:
        t = torch.where(t < 0, -t, t)
        t = torch.where(t >= T, 2 * (T - 1) - t, t)
    valid = (t >= 0) & (t < T)
    g = x[:, :, t.clamp(0, T - 1)]                                             # [B, F, rows, k]
    g = g * valid.view(1, 1, rows, k)
    if lens is not None:
        ln = lens.to(torch.int64).clamp(0, T).view(B, 1, 1, 1)
        g = g * (t.view(1, 1, rows, k) < ln)
    return _r(g.permute(0, 2, 3, 1).reshape(B, rows, k * F))


def tm_to_ncw(x, T, C, x_rows=None, x_row_offset=0):
    return x[:, x_row_offset:x_row_offset + T, :C].float().transpose(1, 2).contiguous()


def _unfold_index(T_out, k, stride, dilation, pad_left):
    return torch.arange(T_out).view(T_out, 1) * stride + torch.arange(k).view(1, k) * dilation - pad_left   # [T_out, k]


def im2col_tm(x, T_out, k, stride, dilation, pad_left):
    assert x.dtype == BF16 and x.is_contiguous()
    B, rows, C = x.shape
    r = _unfold_index(T_out, k, stride, dilation, pad_left)
    valid = (r >= 0) & (r < rows)
    g = x[:, r.clamp(0, rows - 1), :] * valid.view(1, T_out, k, 1).to(BF16)     # [B, T_out, k, C]
    return g.reshape(B, T_out, k * C).contiguous()


def col2im_tm(dcol, x_rows, C, k, stride, dilation, pad_left):
    assert dcol.dtype == BF16 and dcol.is_contiguous()
    B, T_out, KC = dcol.shape
    assert KC == k * C
    r = _unfold_index(T_out, k, stride, dilation, pad_left)
    valid = ((r >= 0) & (r < x_rows)).view(1, T_out * k, 1)
    dx = torch.zeros((B, x_rows, C), dtype=torch.float32)
    src = dcol.float().view(B, T_out * k, C) * valid
    dx.index_add_(1, r.clamp(0, x_rows - 1).reshape(-1), src)
    return _r(dx)


def cast_bf16(src, dst=None):
    out = _r(src.contiguous())
    if dst is None:
        return out
    dst.copy_(out.view(dst.shape))
    return dst


def pack_wt(w_store, wt, cout, cin):
    """fp32 [k, Cout, Cin] -> wt [k, Cin_pad, Cout_pad] bf16, wt[k-1-j][ci][co] = w[j][co][ci]"""
    wt.zero_()
    wt[:, :cin, :cout] = _r(w_store.flip(0).transpose(1, 2))
    return wt


# ------------------------------------------------------------------------------------------------ BatchNorm / activation passes
def bn_stats(z, C, out=None):
    v = z.float().reshape(-1, C)
    st = torch.cat([v.sum(0), (v * v).sum(0)])
    return st if out is None else out.add_(st)


def bn_finalize(stats, rows, C, gamma, beta, conv_bias, eps, momentum, running_mean, running_var, num_batches_tracked=None):
    n = float(rows)
    mean = stats[:C].double() / n
    var = (stats[C:].double() / n - mean * mean).clamp_min(0.0)
    invstd = (1.0 / torch.sqrt(var + eps)).float()
    g = gamma.detach().float() if gamma is not None else torch.ones(C)
    bt = beta.detach().float() if beta is not None else torch.zeros(C)
    out = torch.empty((4, C), dtype=torch.float32)
    out[0] = g * invstd
    out[1] = bt - mean.float() * g * invstd
    out[2] = mean.float()
    out[3] = invstd
    if running_mean is not None:
        full = mean.float() + (conv_bias.detach().float() if conv_bias is not None else 0.0)
        running_mean.mul_(1 - momentum).add_(momentum * full)
        unbiased = var * n / (n - 1.0) if rows > 1 else var
        running_var.mul_(1 - momentum).add_(momentum * unbiased.float())
    if num_batches_tracked is not None:
        num_batches_tracked += 1
    return out


def lens_chain(lens, conv_params):
    if lens.dtype not in (torch.int32, torch.int64):
        lens = lens.to(torch.int64)
    li = lens.to(torch.int64).clone()
    rows = [li.to(torch.int32)]
    for (k, s, d, p) in conv_params:
        if s != 0:
            lf = (li + 2 * p - d * (k - 1) - 1).to(torch.float32) / float(s) + 1.0          # true division (jasper.py:91-95)
            li = lf.to(torch.int64)
        rows.append(li.to(torch.int32))
    return torch.stack(rows), li


def _pre(z, scale, shift, res, res_scale, res_shift):
    pre = z.float() * scale + shift
    if res is not None:
        pre = pre + (res.float() * res_scale + res_shift)
    return pre


def bn_act_pad(z, scale, shift, B, T, C, pad_left, pad_right, act, drop_p=0.0, seed=0, lens=None, res=None, res_scale=None,
               res_shift=None, out=None, drop_mask=None):
    assert drop_p == 0.0, "the host simulation has no dropout (Philox stream of the kernel is not restated)"
    pre = _pre(z.view(B, T, C), scale, shift, None if res is None else res.view(B, T, C), res_scale, res_shift)
    y = _r(torch.where(_row_mask(lens, B, T), torch.zeros_like(pre), _act(pre, act)))
    if out is None:
        out = torch.empty((B, pad_left + T + pad_right, C), dtype=BF16)
    out[:, pad_left:pad_left + T] = y
    for t in range(1, pad_left + 1):                     # mirror rows (reflection halo of the consumer)
        out[:, pad_left - t] = y[:, t]
    for d in range(1, pad_right + 1):
        out[:, pad_left + T - 1 + d] = y[:, T - 1 - d]
    return out


def reflect_halo(y, T, pad_left, pad_right):
    for t in range(1, pad_left + 1):
        y[:, pad_left - t] = y[:, pad_left + t]
    for d in range(1, pad_right + 1):
        y[:, pad_left + T - 1 + d] = y[:, pad_left + T - 1 - d]
    return y


def bn_finalize_act_pad(z, stats, gamma, beta, conv_bias, eps, momentum, running_mean, running_var, num_batches_tracked, B, T, C,
                        pad_left, pad_right, act, drop_p=0.0, seed=0, lens=None, res=None, res_scale=None, res_shift=None, drop_mask=None,
                        zero_after=None):
    """w2l_bn_finalize_act_pad: the two restatements above in one call; the launch also clears ``zero_after``"""
    fin = bn_finalize(stats, B * T, C, gamma, beta, conv_bias, eps, momentum, running_mean, running_var, num_batches_tracked)
    out = bn_act_pad(z, fin[0], fin[1], B, T, C, pad_left, pad_right, act, drop_p, seed, lens, res, res_scale, res_shift, None, drop_mask)
    if zero_after is not None:
        zero_after.zero_()
    return out, fin


def bn_act_bwd(dyp, z, scale, shift, mean, invstd, gamma, B, T, C, pad_left, pad_right, act, drop_p=0.0, seed=0, lens=None,
               res=None, res_scale=None, res_shift=None, want_g=False, dz_rows=None, drop_mask=None, red_ws=None, zero_after=None,
               red_raw=None):
    assert drop_p == 0.0 and red_raw is None
    dz_rows = T if dz_rows is None else dz_rows
    zf = z.view(B, T, C).float()
    pre = _pre(z.view(B, T, C), scale, shift, None if res is None else res.view(B, T, C), res_scale, res_shift)
    d = dyp.view(B, pad_left + T + pad_right, C).float()
    g = d[:, pad_left:pad_left + T].clone()
    for t in range(1, pad_left + 1):                     # fold the halo rows back onto the rows they mirror
        g[:, t] += d[:, pad_left - t]
    for k in range(1, pad_right + 1):
        g[:, T - 1 - k] += d[:, pad_left + T - 1 + k]
    g = torch.where(_act_pass(pre, act) & ~_row_mask(lens, B, T), g, torch.zeros_like(g))
    sg = g.sum((0, 1))
    sx = (g * (zf - mean)).sum((0, 1)) * invstd
    red = torch.cat([sg, sx])
    if red_ws is not None:                               # the kernels ADD into the persistent buffer (zero on entry is the caller's duty):
        red_ws += red                                    # a buffer that was not cleared shows up in the result, as on the device
        red = red_ws.clone()
        sg, sx = red[:C], red[C:]
    if zero_after is not None:
        zero_after.zero_()
    inv_m = 1.0 / float(B * T)
    coef = (gamma.detach().float() if gamma is not None else 1.0) * invstd
    kB = -coef * sx * inv_m * invstd
    kC = -coef * sg * inv_m - kB * mean
    dz = torch.zeros((B, dz_rows, C), dtype=BF16)
    dz[:, :T] = _r(coef * g + kB * zf + kC)
    return dz, red, (_r(g) if want_g else None)


def log_softmax(logits, C, mode=0, nan_flag=None):
    x = logits[..., :C].float()
    out = torch.log_softmax(x, -1) if mode == 0 else torch.softmax(x, -1)
    if nan_flag is not None and bool(torch.isnan(out).any()):
        nan_flag.fill_(1)
    return out.contiguous()


def log_softmax_bwd(g, lp, ld_out, gscale=None, fused_identity=False, out_dtype=torch.bfloat16):
    assert out_dtype == torch.bfloat16
    C = g.shape[-1]
    v = g.float()
    if not fused_identity:
        v = v - torch.exp(lp) * v.sum(-1, keepdim=True)
    if gscale is not None:
        v = v * gscale
    out = torch.zeros(g.shape[:-1] + (ld_out,), dtype=BF16)
    out[..., :C] = _r(v)
    return out


def colsum(x, C):
    return x.float().reshape(-1, x.shape[-1])[:, :C].sum(0)


# ------------------------------------------------------------------------------------------------ implicit-GEMM conv
def _gather_taps(x, T_out, k, dil, off):
    """x [B, x_rows, C] -> [B, T_out, k, C] with [b, t, j] = x[b, t + off + j*dil] (zero outside the buffer: TMA OOB fill)"""
    B, rows, C = x.shape
    r = torch.arange(T_out).view(T_out, 1) + off + torch.arange(k).view(1, k) * dil
    valid = ((r >= 0) & (r < rows)).view(1, T_out, k, 1)
    return x[:, r.clamp(0, rows - 1), :].float() * valid


def conv1d_fwd(x, w, desc, y, bias=None, scale=None, shift=None, bn_stats=None):
    """x bf16 [B, x_rows, Cin]; w bf16 [k, Cout_pad, Cin]; y [B, y_rows, ldy] (bf16 | fp32) rows [y_row_offset, +T_out), cols [0, Cout)"""
    d = desc
    assert x.dtype == BF16 and w.dtype == BF16 and tuple(x.shape) == (d.B, d.x_rows, d.Cin), (x.shape, d.B, d.x_rows, d.Cin)
    assert tuple(w.shape) == (d.k, d.Cout_pad, d.Cin) and d.Cin >= 64 and d.Cin % 8 == 0 and d.Cout_pad % 16 == 0
    assert y.shape[0] == d.B and y.shape[1] == d.y_rows and y.shape[2] == d.ldy and d.y_rows >= d.T_out + d.y_row_offset
    a = _gather_taps(x, d.T_out, d.k, d.dilation, d.x_row_offset)                        # [B, T, k, Cin]
    acc = torch.einsum("btjc,joc->bto", a, w[:, :d.Cout].float())
    if bias is not None or scale is not None:
        sc = scale.float() if scale is not None else torch.ones(d.Cout)
        sh = (bias.detach().float() * sc if bias is not None else 0.0) + (shift.float() if shift is not None else 0.0)
        acc = acc * sc + sh
    acc = _act(acc, d.act)
    stored = acc.to(y.dtype)
    if bn_stats is not None:
        v = stored.float().reshape(-1, d.Cout)
        bn_stats[:d.Cout] += v.sum(0)
        bn_stats[d.Cout:] += (v * v).sum(0)
    y[:, d.y_row_offset:d.y_row_offset + d.T_out, :d.Cout] = stored
    return y


def conv1d_dgrad_wt(dy, wt, desc, dx, bnred=None):
    assert bnred is None, "the fused BatchNorm-backward reduction is switched off under the torch restatement (install())"
    """dx[b, u, ci] = sum_j sum_co dy[b, u - off - j*dil, co] * w[j][co][ci]; wt [k, Cin_pad16, Cout_pad] = tap-reversed transpose"""
    d = desc
    assert dy.dtype == BF16 and wt.dtype == BF16 and dy.shape[-1] == d.ldy and d.ldy >= d.Cout and d.Cout_pad >= 64
    assert wt.shape[0] == d.k and wt.shape[2] == d.Cout_pad and wt.shape[1] >= d.Cin
    assert dy.numel() == d.B * d.y_rows * d.ldy and dx.numel() == d.B * d.x_rows * d.Cin
    cols = d.Cout_pad if d.ldy >= d.Cout_pad else d.Cout      # rows narrower than Cout_pad: the missing columns read as zero
    g = dy.reshape(d.B, d.y_rows, d.ldy)[:, d.y_row_offset:d.y_row_offset + d.T_out, :cols]
    a = _gather_taps(g.contiguous(), d.x_rows, d.k, d.dilation, -d.x_row_offset - (d.k - 1) * d.dilation)   # [B, x_rows, k, cols]
    out = torch.einsum("bujo,jco->buc", a, wt[:, :d.Cin, :cols].float())
    dx.view(d.B, d.x_rows, d.Cin).copy_(_r(out))
    return dx


def conv1d_dgrad(dy, w, desc, dx):
    d = desc
    wt = torch.zeros((d.k, (d.Cin + 15) // 16 * 16, d.Cout_pad), dtype=BF16)
    wt[:, :d.Cin] = w.flip(0).transpose(1, 2)
    return conv1d_dgrad_wt(dy, wt, desc, dx)


def conv1d_wgrad(dy, x, desc, dw):
    """dw[j, co, ci] = sum_{b,t} dy[b, y_row_offset + t, co] * x[b, t + x_row_offset + j*dil, ci]"""
    d = desc
    assert dy.dtype == BF16 and x.dtype == BF16 and dw.dtype == torch.float32 and tuple(dw.shape) == (d.k, d.Cout, d.Cin)
    assert dy.numel() == d.B * d.y_rows * d.ldy and tuple(x.shape) == (d.B, d.x_rows, d.Cin) and d.ldy >= 64 and d.ldy % 8 == 0
    g = dy.reshape(d.B, d.y_rows, d.ldy)[:, d.y_row_offset:d.y_row_offset + d.T_out, :d.Cout].float()
    a = _gather_taps(x, d.T_out, d.k, d.dilation, d.x_row_offset)
    dw.copy_(torch.einsum("bto,btjc->joc", g, a))
    return dw


# ------------------------------------------------------------------------------------------------ depthwise
def _ncw(x):
    return x.float().transpose(1, 2)


def depthwise_fwd(x, w, T_out, k, stride, dilation, pad, out_lens=None):
    B, T, C = x.shape
    y = TF.conv1d(_ncw(x), w.t().reshape(C, 1, k).float(), stride=stride, dilation=dilation, padding=pad, groups=C).transpose(1, 2)
    assert y.shape[1] == T_out
    return _r(torch.where(_row_mask(out_lens, B, T_out), torch.zeros_like(y), y)).contiguous()


def depthwise_dgrad(dy, w, T, k, dilation, pad, dy_lens=None, stride=1):
    B, T_out, C = dy.shape
    g = torch.where(_row_mask(dy_lens, B, T_out), torch.zeros(1), dy.float())
    with torch.enable_grad():                                  # (called from inside autograd.Function.backward)
        x = torch.zeros((B, C, T), requires_grad=True)
        y = TF.conv1d(x, w.t().reshape(C, 1, k).float(), stride=stride, dilation=dilation, padding=pad, groups=C)
        (dx,) = torch.autograd.grad(y, x, g.transpose(1, 2))
    return _r(dx.transpose(1, 2)).contiguous()


def depthwise_wgrad(dy, x, k, stride, dilation, pad, dy_lens=None):
    B, T_out, C = dy.shape
    g = torch.where(_row_mask(dy_lens, B, T_out), torch.zeros(1), dy.float())
    with torch.enable_grad():
        w = torch.zeros((C, 1, k), requires_grad=True)
        y = TF.conv1d(_ncw(x), w, stride=stride, dilation=dilation, padding=pad, groups=C)
        (dw,) = torch.autograd.grad(y, w, g.transpose(1, 2))
    return dw[:, 0, :].t().contiguous()


# ------------------------------------------------------------------------------------------------ CTC
def ctc_loss_raw(x, targets, input_lengths, target_lengths, blank=0, zero_infinity=True, reduction_mean=True, from_logits=False,
                 need_grad=True):
    """torch's CPU ctc_loss stands in for the CTC kernel (the reference's own criterion, base_asr_models.py:23): same loss, and its
    autograd returns the ATen-convention gradient with respect to the log-probs that the kernel emits"""
    assert not from_logits
    il, tl = input_lengths.to(torch.int64), target_lengths.to(torch.int64)
    with torch.enable_grad():                                  # (called from inside autograd.Function.forward)
        lp = x.detach().float().clone().requires_grad_(need_grad)
        nll = TF.ctc_loss(lp.transpose(0, 1), targets.to(torch.int64), il, tl, blank=blank, reduction="none", zero_infinity=zero_infinity)
        loss = (nll / tl.clamp_min(1).float()).mean() if reduction_mean else nll.sum()
        grad = None
        if need_grad:
            (grad,) = torch.autograd.grad(loss, lp)
    return loss.detach().reshape(1), nll.detach(), grad


# ------------------------------------------------------------------------------------------------ install
_NAMES = ["im2col_ncw", "tm_to_ncw", "im2col_tm", "col2im_tm", "cast_bf16", "pack_wt", "bn_stats", "bn_finalize", "lens_chain",
          "bn_act_pad", "bn_finalize_act_pad", "reflect_halo", "bn_act_bwd", "log_softmax", "log_softmax_bwd", "colsum", "conv1d_fwd", "conv1d_dgrad",
          "conv1d_dgrad_wt", "conv1d_wgrad", "depthwise_fwd", "depthwise_dgrad", "depthwise_wgrad", "ctc_loss_raw"]


def install(monkeypatch):
    """Swap the C-ABI wrappers of ``functional`` for the restatements above (for the duration of one test)."""
    from wav2letter_pytorch_b200 import functional as F
    from wav2letter_pytorch_b200 import layers
    for n in _NAMES:
        assert hasattr(F, n), n
        monkeypatch.setattr(F, n, globals()[n])
    monkeypatch.setattr(F, "_need_cuda", lambda *ts: None)
    monkeypatch.setattr(layers.WgradStream, "enabled", False)
    monkeypatch.setattr(layers.FusedBnReduce, "enabled", False)      # (the emu / cabi backends switch it back on: their GEMM epilogue has it)
    return F
