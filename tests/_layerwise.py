"""TEST INFRASTRUCTURE -- teacher-forced, layer-by-layer parity of a training step at the widths of the headline configurations.

Why layer by layer.  A freshly initialised BatchNorm + clamp/ReLU stack is *chaotic*: removing the per-channel mean enlarges every
perturbation relative to the signal by sqrt(E[h^2] / Var[h]) ~ 1.2 per layer (the gradient-explosion rate (pi/(pi-1))^(1/2) per layer of
mean-field BatchNorm theory), i.e. ~45x over the 20 layers of the reference's wav2letter.yaml.  Any implementation that stores bf16
anywhere (rounding 2^-9) therefore differs from an fp32 run by tens of percent in the early layers' gradients END TO END -- the CPU
oracle with emulated bf16 storage differs from the fp32 oracle by 0.5-0.85 relative L2 there (measured) -- so an
end-to-end gradient comparison cannot carry a closed tolerance at that depth, and an open-ended one would let a wiring error through.
Teacher forcing removes the amplification and keeps everything else: the model runs its real training step once (all layers, real
widths / kernel sizes / dilation, CTC loss at the end); then EVERY block is checked on its own against the oracle -- the oracle block gets
the activations the device actually fed into that block, and the gradient the device actually fed back into it, and must reproduce the
block's output, its input gradient and its parameter gradients within fixed bounds.  A wrong row offset, a dropped halo / residual /
mask term or a mis-scaled BatchNorm reduction in any layer is an O(1) error in that layer's row of the table."""
import torch
import torch.nn.functional as TF

from oracle import w2l_oracle as O


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().flatten().cpu(), torch.as_tensor(b).double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _ncw(h, C=None):
    """time-major device tensor [B, rows, Cp] -> NCW fp32 on the host; ``C``: the logical channel count when the buffer's width is
    padded (ConvParams.phys) -- the surplus channels must then be exact zeros"""
    h = h.detach().float().cpu()
    if C is not None and h.shape[2] != C:
        assert float(h[:, :, C:].abs().max()) == 0.0, "surplus channels of a padded activation / gradient buffer must be exactly zero"
        h = h[:, :, :C]
    return h.transpose(1, 2).contiguous()


def _bf(t):
    return t.to(torch.bfloat16).to(t.dtype)


# ------------------------------------------------------------------------------------------------ Wav2Letter
def w2l_run_blocks(model, x, il, tg, tl):
    """the model's own forward (Wav2Letter.forward, with its parity tap switched on so that every block's input keeps its gradient),
    then CTC and backward.  Returns (hs, out, out_lens, loss): hs[i] = input of block i."""
    model._tap = []
    try:
        out, out_lens = model(x, il)
    finally:
        hs = model.__dict__.pop("_tap")
    loss = model.criterion(out.transpose(0, 1), tg, out_lens, tl)
    loss.backward()
    return hs, out, out_lens, loss


def w2l_block_oracle(blk, hin, first, last, emu):
    """wav2letter.py:40-47 for ONE block on the host: (pad ->) conv -> BatchNorm(batch statistics) -> clamp(0, 20) (-> the consumer's
    reflection pad, which this build's producer writes).  ``emu``: bf16 weights, conv output and result stored as bf16 (DESIGN.md
    section 4), fp32 arithmetic otherwise.  Returns (result, dict of leaf tensors)."""
    k, s, d = blk.kernel_size[0], blk.stride, blk.dilation
    w = blk.conv1.weight.detach().float().cpu().contiguous().clone().requires_grad_(True)
    b = blk.conv1.bias.detach().float().cpu().clone().requires_grad_(True)
    leaves = {"conv1.weight": w, "conv1.bias": b}
    h = hin
    if first:
        h = _bf(h) if emu else h
        if sum(blk.pad_lr) > 0:
            h = TF.pad(h, blk.pad_lr, mode="reflect")
    W = O._bf16_weight(w) if emu else w
    if last:                                                 # the label head: conv + bias, log_softmax over the labels
        z = TF.conv1d(h, W, b, stride=s, dilation=d)
        return torch.log_softmax(z.transpose(1, 2), -1), leaves
    z = TF.conv1d(h, W, None if emu else b, stride=s, dilation=d)   # train-mode BN cancels the bias; the device never adds it
    if emu:
        z = O._RoundBF16.apply(z)
    gamma = blk.batch_norm.weight.detach().float().cpu().clone().requires_grad_(True)
    beta = blk.batch_norm.bias.detach().float().cpu().clone().requires_grad_(True)
    leaves.update({"batch_norm.weight": gamma, "batch_norm.bias": beta})
    y = TF.batch_norm(z, None, None, gamma, beta, training=True, momentum=0.9, eps=1e-3)
    y = torch.clamp(y, 0, 20)
    if sum(blk.next_pad) > 0:
        y = TF.pad(y, blk.next_pad, mode="reflect")
    if emu:
        y = O._RoundBF16.apply(y)
    return y, leaves


def w2l_layerwise_table(model, hs, out):
    """[(block name, {quantity: (err vs emulated-bf16 oracle, err vs fp32 oracle)})] for every block of the stack."""
    blocks = list(model.conv1ds.named_children())
    table = []
    for i, (name, blk) in enumerate(blocks):
        first, last = i == 0, i == len(blocks) - 1
        dev_out = out if last else hs[i + 1]
        cot = dev_out.grad.detach().float().cpu()
        c_in, c_out = blk.input_channels, blk.output_channels         # logical widths (the device buffers may be padded)
        row = {}
        for emu in (True, False):
            hin = (hs[i].detach().float().cpu() if first else _ncw(hs[i], c_in)).clone().requires_grad_(not first)
            y, leaves = w2l_block_oracle(blk, hin, first, last, emu)
            if last:
                want, got = y, dev_out.detach().float().cpu()
                y.backward(cot)
            else:
                want, got = y, _ncw(dev_out, c_out)
                y.backward(cot[:, :, :c_out].transpose(1, 2))
            col = 0 if emu else 1
            row.setdefault("out", [None, None])[col] = rel_l2(got, want.detach())
            if not first:
                row.setdefault("d_input", [None, None])[col] = rel_l2(_ncw(hs[i].grad, c_in), hin.grad)
            for pname, leaf in leaves.items():
                p = dict(blk.named_parameters())[pname]
                if pname == "conv1.bias" and not last:       # analytically zero under train-mode BN: the device returns exact zeros
                    assert float(p.grad.abs().max()) == 0.0, (name, "conv bias gradient must be exactly zero")
                    continue
                row.setdefault("d_" + pname, [None, None])[col] = rel_l2(p.grad, leaf.grad)
        table.append((name, row))
    return table


# ------------------------------------------------------------------------------------------------ Jasper
def jasper_run_blocks(model, x, il, tg, tl, F=None):
    """Jasper.forward with its parity tap switched on, CTC, backward.  Returns (hs, taps, rows, out, out_lens, loss): hs[i] = input of
    block i (hs[-1] = encoder output), taps[i] = [(input of sub-block r, index into ``rows`` of the lengths entering it)], rows = the
    truncated lengths entering every masked conv (``F.lens_chain``)."""
    model._tap = {"hs": [], "taps": []}
    try:
        out, out_lens = model(x, il)
    finally:
        tap = model.__dict__.pop("_tap")
    loss = model.criterion(out.transpose(0, 1), tg, out_lens, tl)
    loss.backward()
    return tap["hs"], tap["taps"], tap["rows"].detach().cpu().long(), out, out_lens, loss


def _mask_rows(y_ncw, lens):
    T = y_ncw.shape[2]
    m = torch.arange(T).expand(len(lens), T) >= lens.unsqueeze(1)
    return y_ncw.masked_fill(m.unsqueeze(1), 0)


def jasper_sub_oracle(spec, sd_i, bi, r, hin, lens, block_in, lens_block, emu):
    """ONE conv+BN group of a JasperBlock on the host (jasper.py:300-368 via the oracle's ``_jasper_conv_bn``), then -- for the last
    group -- the 1x1 residual branch of the block input and the add (jasper.py:400-412), then ReLU (dropout 0) and, with ``emu``, the
    bf16 store of the result.  Returns the NCW output."""
    Wq, Rz, Ry = (O._bf16_weight, O._RoundBF16.apply, O._RoundBF16.apply) if emu else (O._ident, O._ident, O._ident)
    step = 5 if spec["separable"] else 4
    cin = spec["cin"] if r == 0 else spec["cout"]
    out, _, _ = O._jasper_conv_bn(hin, lens, sd_i, "jasper_encoder.%d.mconv." % bi, r * step, spec, cin, spec["k"], spec["stride"],
                                  spec["pad"], spec["dilation"], True, spec["separable"], emu)
    if r == spec["repeat"] - 1 and spec["residual"]:
        rp = "jasper_encoder.%d.res.0." % bi
        res, _ = O.masked_conv1d(block_in, lens_block, Wq(sd_i[rp + "0.conv.weight"]), 1, 0, 1, 1, spec["conv_mask"])
        res = TF.batch_norm(Rz(res), None, None, sd_i[rp + "1.weight"], sd_i[rp + "1.bias"], training=True, momentum=0.1, eps=1e-3)
        if emu:
            res = O._RoundGradBF16.apply(res)
        out = out + res
    return Ry(TF.relu(out))


def jasper_layerwise_table(model, specs, hs, taps, rows, out):
    """one row per conv+BN group of every JasperBlock (named block<i>.<r>) + the head; the block's input gradient is checked as the
    sum of what the first group and the residual branch return"""
    blocks = list(model.jasper_encoder)
    params = dict(model.named_parameters())
    sd = {k: v.detach().float().cpu().clone() if v.is_floating_point() else v.detach().cpu().clone()
          for k, v in model.state_dict().items()}
    table = []
    for i, blk in enumerate(blocks):
        spec, R = specs[i], specs[i]["repeat"]
        assert len(taps[i]) == R
        prefix = "jasper_encoder.%d." % i
        # logical widths (the device buffers may be padded, see _ncw): the block's input and its BatchNorms' feature count
        cin, planes = blk.mconv[0].conv.in_channels, [m for m in blk.mconv if hasattr(m, "num_features")][-1].num_features
        rows_of = [dict() for _ in range(R)]
        for emu in (True, False):
            col = 0 if emu else 1
            leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()
                      if k.startswith(prefix) and v.is_floating_point() and "running" not in k}
            sd_i = dict(sd)
            sd_i.update(leaves)
            d_block_in = None
            for r in reversed(range(R)):
                first = i == 0 and r == 0
                h_dev, ri = taps[i][r]
                c_in = cin if r == 0 else planes
                hin = (hs[0].detach().float().cpu() if first else _ncw(h_dev, c_in)).clone().requires_grad_(not first)
                xin = _bf(hin) if (first and emu) else hin
                lens = rows[ri].clone()
                last_sub = r == R - 1
                block_in = None
                if last_sub and spec["residual"]:
                    block_in = hin if R == 1 else _ncw(hs[i], cin).clone().requires_grad_(True)
                y = jasper_sub_oracle(spec, sd_i, i, r, xin, lens, block_in, rows[taps[i][0][1]].clone(), emu)
                dev_out = hs[i + 1] if last_sub else taps[i][r + 1][0]
                nxt_ri = (taps[i + 1][0][1] if i + 1 < len(blocks) else None) if last_sub else taps[i][r + 1][1]
                if spec["conv_mask"] and nxt_ri is not None:
                    y = _mask_rows(y, rows[nxt_ri])         # the consumer's masked_fill (jasper.py:116-119), written by this build's producer
                y.backward(_ncw(dev_out.grad, planes))
                rows_of[r].setdefault("out", [None, None])[col] = rel_l2(_ncw(dev_out, planes), y.detach())
                if last_sub and spec["residual"] and R > 1:
                    d_block_in = block_in.grad
                if not first:
                    want = hin.grad
                    if r == 0 and d_block_in is not None:
                        want = want + d_block_in
                    # rows past an utterance's length carry no gradient in the reference (masked_fill); the device leaves them unspecified
                    got = _mask_rows(_ncw(h_dev.grad, c_in), lens) if spec["conv_mask"] else _ncw(h_dev.grad, c_in)
                    rows_of[r].setdefault("d_input", [None, None])[col] = rel_l2(got, want)
            step = 5 if spec["separable"] else 4
            for k, leaf in leaves.items():
                local = k[len(prefix):]
                r = int(local.split(".")[1]) // step if local.startswith("mconv.") else R - 1
                rows_of[r].setdefault("d_" + local, [None, None])[col] = rel_l2(params[k].grad, leaf.grad)
        for r in range(R):
            table.append(("block%d.%d" % (i, r), rows_of[r]))
    # head: unmasked 1x1 conv with bias -> log_softmax (jasper.py:433, 468-472)
    row = {}
    head = model.final_layer[0]
    for emu in (True, False):
        w = head.weight.detach().float().cpu().contiguous().clone().requires_grad_(True)
        b = head.bias.detach().float().cpu().clone().requires_grad_(True)
        hin = _ncw(hs[-1], head.in_channels).clone().requires_grad_(True)
        lp = torch.log_softmax(TF.conv1d(hin, O._bf16_weight(w) if emu else w, b).transpose(1, 2), -1)
        lp.backward(out.grad.detach().float().cpu())
        col = 0 if emu else 1
        row.setdefault("out", [None, None])[col] = rel_l2(out.detach().float().cpu(), lp.detach())
        row.setdefault("d_input", [None, None])[col] = rel_l2(_ncw(hs[-1].grad, head.in_channels), hin.grad)
        row.setdefault("d_weight", [None, None])[col] = rel_l2(head.weight.grad, w.grad)
        row.setdefault("d_bias", [None, None])[col] = rel_l2(head.bias.grad, b.grad)
    table.append(("head", row))
    return table


def format_table(table):
    lines = ["%-12s %-28s %10s %10s" % ("block", "quantity", "vs bf16emu", "vs fp32")]
    for name, row in table:
        for q, (e, r) in row.items():
            lines.append("%-12s %-28s %10.2e %10.2e" % (name, q, e, r))
    return "\n".join(lines)


def check_table(table, tol_emu, tol_ref):
    """every entry under its fixed bound: ``tol_*`` map a quantity kind ('out', 'd_input', 'd_param') to a number"""
    bad = []
    for name, row in table:
        for q, (e, r) in row.items():
            kind = q if q in ("out", "d_input") else "d_param"
            if not (e < tol_emu[kind]) or not (r < tol_ref[kind]):
                bad.append((name, q, e, r))
    return bad
