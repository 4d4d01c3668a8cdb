"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module, and only as the checker / the CPU yard-stick -- never as the thing
shipped.  The product path (``wav2letter_pytorch_b200``) must not import anything from ``oracle/``.

What is restated (reference file:line given at each function):
  * Conv1dBlock / Wav2Letter forward      -- wav2letter.py:12-92
  * MaskedConv1d / JasperBlock / Jasper   -- jasper.py:53-132, 154-419, 422-475
  * CTC loss (mean, blank 0, zero_inf)    -- base_asr_models.py:23,81 (+ torch.nn.CTCLoss semantics)
  * GreedyDecoder argmax + collapse       -- decoder.py:89-145
  * Novograd.step                         -- novograd.py:52-114

The arithmetic of conv / batch-norm / log_softmax in the reference is the installed torch's CPU
library (the reference calls nn.Conv1d etc. and pins no version), so the oracle calls the same
``torch.nn.functional`` primitives on CPU tensors, in fp32 or fp64.  CTC and greedy decoding are
additionally restated from scratch (numpy / python loops) so that they can be checked independently
of torch.  Parity pinning: ``tests/test_oracle_golden.py`` checks every function here against
``tests/golden/*.npz``, which ``oracle/gen_golden.py`` produced by running the *unmodified*
reference modules in the build container (where /root/reference is mounted).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# labels (data/label_sets.py:2-14): blank '_' at 0, then ', a..z, then ' ' at 28
# --------------------------------------------------------------------------------------------
ENGLISH_LOWERCASE = ["_", "'"] + [chr(ord("a") + i) for i in range(26)] + [" "]

# --------------------------------------------------------------------------------------------
# Wav2Letter layer table (configuration/model/wav2letter.yaml:5-104):
# (output_size, kernel_size, stride, dilation, dropout)
# --------------------------------------------------------------------------------------------
W2L_LAYERS = (
    [(256, 11, 2, 1, 0.2)] + [(256, 11, 1, 1, 0.2)] * 3 + [(384, 13, 1, 1, 0.2)] * 3
    + [(512, 17, 1, 1, 0.2)] * 3 + [(640, 21, 1, 1, 0.3)] * 3 + [(768, 25, 1, 1, 0.3)] * 3
    + [(896, 29, 1, 2, 0.4)] * 3 + [(1024, 1, 1, 1, 0.4)]
)


def w2l_layer_specs(mid_layers, input_size=64, n_labels=29, dropout=None):
    """wav2letter.py:59-71 -- blocks from cfg.layers[:mid_layers] plus the k=1 head (no BN/activation)."""
    specs, cin = [], input_size
    for (co, k, s, d, p) in W2L_LAYERS[:mid_layers]:
        specs.append(dict(cin=cin, cout=co, k=k, stride=s, dilation=d,
                          dropout=p if dropout is None else dropout, bn=True, act=True))
        cin = co
    specs.append(dict(cin=cin, cout=n_labels, k=1, stride=1, dilation=1, dropout=-1.0, bn=False, act=False))
    return specs


def reflect_pad_amounts(cin, k, stride, dilation):
    """wav2letter.py:24-34.  NB: the rule is computed from the *channel count*, not the time length."""
    out_rows = (cin + stride - 1) // stride
    pad = max(0, (out_rows - 1) * stride + (k - 1) * dilation + 1 - cin)
    return pad // 2, (pad + 1) // 2


def conv1d_block_forward(x, sd, prefix, spec, training, dropout_mask=None, update_running=True):
    """wav2letter.py:40-47: ReflectionPad1d -> Conv1d(bias) -> BatchNorm1d(eps 1e-3, momentum 0.9)
    -> Dropout -> clamp(0, 20).  ``x`` is [B, Cin, T] (NCW).  ``sd`` maps reference state_dict names to
    tensors; running stats are updated in place when training (as nn.BatchNorm1d does)."""
    pl, pr = reflect_pad_amounts(spec["cin"], spec["k"], spec["stride"], spec["dilation"])
    if pl + pr > 0:
        x = F.pad(x, (pl, pr), mode="reflect")
    z = F.conv1d(x, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"], stride=spec["stride"],
                 dilation=spec["dilation"])
    if spec["bn"]:
        rm, rv = sd[prefix + "batch_norm.running_mean"], sd[prefix + "batch_norm.running_var"]
        if training and not update_running:
            rm, rv = rm.clone(), rv.clone()
        z = F.batch_norm(z, rm, rv, sd[prefix + "batch_norm.weight"], sd[prefix + "batch_norm.bias"],
                         training=training, momentum=0.9, eps=1e-3)
    if training and spec["dropout"] != -1 and spec["dropout"] > 0:
        if dropout_mask is not None:           # mask shared with the implementation under test
            z = z * dropout_mask / (1.0 - spec["dropout"])
        else:
            z = F.dropout(z, spec["dropout"], training=True)
    if spec["act"]:
        z = torch.clamp(z, min=0, max=20)
    return z


def w2l_forward(x, input_lengths, sd, specs, training, update_running=True, return_logits=False):
    """wav2letter.py:84-92.  Returns (log_probs [B,T',C], output_lengths)."""
    for i, spec in enumerate(specs):
        x = conv1d_block_forward(x, sd, "conv1ds.conv1d_%d." % i, spec, training, update_running=update_running)
    logits = x.transpose(1, 2)
    lp = F.log_softmax(logits, dim=-1)
    scaling = int(np.prod([s["stride"] for s in specs]))
    out_len = None if input_lengths is None else input_lengths // scaling   # base_asr_models.py:33-39
    if return_logits:
        return lp, out_len, logits
    return lp, out_len


def w2l_init_state_dict(specs, seed=0, dtype=torch.float32):
    """Random init with nn.Conv1d's default scheme (kaiming-uniform a=sqrt(5) => U(-1/sqrt(fan_in), ..))."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for i, s in enumerate(specs):
        p = "conv1ds.conv1d_%d." % i
        bound = 1.0 / math.sqrt(s["cin"] * s["k"])
        sd[p + "conv1.weight"] = ((torch.rand(s["cout"], s["cin"], s["k"], generator=g) * 2 - 1) * bound).to(dtype)
        sd[p + "conv1.bias"] = ((torch.rand(s["cout"], generator=g) * 2 - 1) * bound).to(dtype)
        if s["bn"]:
            sd[p + "batch_norm.weight"] = torch.ones(s["cout"], dtype=dtype)
            sd[p + "batch_norm.bias"] = torch.zeros(s["cout"], dtype=dtype)
            sd[p + "batch_norm.running_mean"] = torch.zeros(s["cout"], dtype=dtype)
            sd[p + "batch_norm.running_var"] = torch.ones(s["cout"], dtype=dtype)
            sd[p + "batch_norm.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return sd


# --------------------------------------------------------------------------------------------
# Jasper (dense / separable, batch norm, 'add' residual, groups=1 -- what Jasper._build_encoder
# can reach, jasper.py:436-453)
# --------------------------------------------------------------------------------------------
def jasper_kernel_size(k, factor=1.0):
    """jasper.py:53-58 -- even kernel sizes are bumped to the next odd value."""
    k = max(int(k * factor), 1)
    return k + 1 if k % 2 == 0 else k


def jasper_same_padding(k, stride, dilation):
    """jasper.py:61-66."""
    if stride > 1 and dilation > 1:
        raise ValueError("Only stride OR dilation may be greater than 1")
    if dilation > 1:
        return (dilation * k) // 2 - 1
    return k // 2


def jasper_block_specs(blocks, input_size=64):
    """``blocks``: list of dicts with the reference's yaml keys (layer_size, kernel_size, stride,
    dilation, residual, repeat, separable, dropout, conv_mask) -- defaults as jasper.py:440-449."""
    specs, cin = [], input_size
    for b in blocks:
        k = jasper_kernel_size(b["kernel_size"])
        s, d = b.get("stride", 1), b.get("dilation", 1)
        specs.append(dict(cin=cin, cout=b["layer_size"], k=k, stride=s, dilation=d,
                          pad=jasper_same_padding(k, s, d), residual=bool(b["residual"]),
                          repeat=b.get("repeat", 1), separable=b.get("separable", True),
                          conv_mask=b.get("conv_mask", True), dropout=b.get("dropout", 0)))
        cin = b["layer_size"]
    return specs


def masked_conv1d(x, lens, w, stride, pad, dilation, groups, use_mask, bias=None):
    """jasper.py:114-132 -- zero t >= len, conv, lens' = (lens + 2p - d(k-1) - 1)/s + 1 (true division)."""
    if use_mask:
        lens = lens.to(dtype=torch.long)
        T = x.size(2)
        mask = torch.arange(T).expand(len(lens), T) >= lens.unsqueeze(1)
        x = x.masked_fill(mask.unsqueeze(1), 0)
        lens = (lens + 2 * pad - dilation * (w.shape[2] - 1) - 1) / stride + 1
    return F.conv1d(x, w, bias, stride=stride, padding=pad, dilation=dilation, groups=groups), lens


def _ident(t):
    return t


def _jasper_conv_bn(x, lens, sd, prefix, idx, spec, cin, k, stride, pad, dilation, training, separable, emu=False):
    """one _get_conv_bn_layer group (jasper.py:300-368): [depthwise+]conv, BatchNorm1d(eps 1e-3, mom 0.1).
    ``emu``: bf16 weights / bf16-stored conv outputs (the CUDA path's storage precision), fp32 accumulation."""
    Wq, Rz = (_bf16_weight, _RoundBF16.apply) if emu else (_ident, _ident)
    if separable:
        x, lens = masked_conv1d(x, lens, sd["%s%d.conv.weight" % (prefix, idx)], stride, pad, dilation,
                                cin, spec["conv_mask"])           # depthwise weights stay fp32 in the CUDA path
        x = Rz(x)                                                  # ... and its output is stored as bf16
        idx += 1
        x, lens = masked_conv1d(x, lens, Wq(sd["%s%d.conv.weight" % (prefix, idx)]), 1, 0, 1, 1, spec["conv_mask"])
        idx += 1
    else:
        x, lens = masked_conv1d(x, lens, Wq(sd["%s%d.conv.weight" % (prefix, idx)]), stride, pad, dilation, 1,
                                spec["conv_mask"])
        idx += 1
    x = Rz(x)
    p = "%s%d." % (prefix, idx)
    x = F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                     training=training, momentum=0.1, eps=1e-3)
    return x, lens, idx + 1


def jasper_block_forward(x, lens, sd, bi, spec, training, emu=False):
    """JasperBlock.forward, jasper.py:379-419 (non-dense residual, 'add')."""
    Wq, Rz, Ry = (_bf16_weight, _RoundBF16.apply, _RoundBF16.apply) if emu else (_ident, _ident, _ident)
    xs, lens_orig = x, lens
    prefix = "jasper_encoder.%d.mconv." % bi
    idx, cin, out = 0, spec["cin"], x
    for r in range(spec["repeat"]):
        out, lens, idx = _jasper_conv_bn(out, lens, sd, prefix, idx, spec, cin, spec["k"], spec["stride"],
                                         spec["pad"], spec["dilation"], training, spec["separable"], emu)
        cin = spec["cout"]
        if r != spec["repeat"] - 1:
            out = F.relu(out)
            if training and spec["dropout"] > 0:
                out = F.dropout(out, spec["dropout"], True)
            out = Ry(out)
            idx += 2                                     # activation + dropout occupy ModuleList slots
    if spec["residual"]:
        rp = "jasper_encoder.%d.res.0." % bi
        res, _ = masked_conv1d(xs, lens_orig, Wq(sd[rp + "0.conv.weight"]), 1, 0, 1, 1, spec["conv_mask"])
        res = Rz(res)
        res = F.batch_norm(res, sd[rp + "1.running_mean"], sd[rp + "1.running_var"], sd[rp + "1.weight"],
                           sd[rp + "1.bias"], training=training, momentum=0.1, eps=1e-3)
        if emu:
            res = _RoundGradBF16.apply(res)              # the masked upstream gradient is handed over as bf16
        out = out + res
    out = F.relu(out)
    if training and spec["dropout"] > 0:
        out = F.dropout(out, spec["dropout"], True)
    return Ry(out), lens


def jasper_forward(x, input_lengths, sd, specs, training, emu=False):
    """Jasper.forward, jasper.py:462-475: encoder, unmasked 1x1 head with bias, transpose,
    log_softmax when training / softmax in eval (reference quirk, replicated)."""
    lens = input_lengths
    if emu:
        x = x.to(torch.bfloat16).float()
    for bi, spec in enumerate(specs):
        x, lens = jasper_block_forward(x, lens, sd, bi, spec, training, emu)
    out_len = lens.to(dtype=int)
    w = _bf16_weight(sd["final_layer.0.weight"]) if emu else sd["final_layer.0.weight"]
    z = F.conv1d(x, w, sd["final_layer.0.bias"])
    if emu:
        z = _RoundGradBF16.apply(z)
    z = z.transpose(2, 1)
    z = F.log_softmax(z, dim=-1) if training else F.softmax(z, dim=-1)
    assert not (z != z).any()
    return z, out_len


# --------------------------------------------------------------------------------------------
# CTC
# --------------------------------------------------------------------------------------------
def ctc_loss_torch(log_probs_ntc, targets, input_lengths, target_lengths, dtype=torch.float32):
    """The reference's exact call (base_asr_models.py:23,81): nn.CTCLoss(blank=0, reduction='mean',
    zero_infinity=True) on ``out.transpose(0,1)`` ([T,N,C] view of the contiguous [N,T,C] tensor).
    Returns (loss, d loss / d log_probs as [N,T,C])."""
    lp = log_probs_ntc.detach().to("cpu", dtype).clone().requires_grad_(True)
    crit = torch.nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)
    loss = crit(lp.transpose(0, 1), targets.cpu(), input_lengths.cpu(), target_lengths.cpu())
    (g,) = torch.autograd.grad(loss, lp)
    return loss.detach(), g


def _lse(a, b):
    if a == -math.inf:
        return b
    if b == -math.inf:
        return a
    m = max(a, b)
    return m + math.log(math.exp(a - m) + math.exp(b - m))


def ctc_nll_and_grad_numpy(lp, target, T, blank=0):
    """From-scratch log-space alpha/beta recursion over the blank-extended label lattice for ONE
    utterance (float64).  ``lp`` [Tmax, C] log-probs, ``target`` 1-D ints (no blanks), ``T`` valid frames.
    Returns (nll, dnll/dlp [Tmax, C]) with the ATen convention grad = exp(lp) - occupancy/exp(lp)
    for t < T and 0 beyond (torch CTCLoss applied to a log_softmax output)."""
    lp = np.asarray(lp, dtype=np.float64)
    S = len(target)
    L = 2 * S + 1
    ext = [blank] * L
    for i, c in enumerate(target):
        ext[2 * i + 1] = int(c)
    NEG = -math.inf
    grad = np.zeros_like(lp)
    if T == 0:
        return (0.0 if S == 0 else math.inf), grad
    alpha = np.full((T, L), NEG)
    alpha[0, 0] = lp[0, blank]
    if L > 1:
        alpha[0, 1] = lp[0, ext[1]]
    for t in range(1, T):
        for s in range(L):
            a = alpha[t - 1, s]
            if s >= 1:
                a = _lse(a, alpha[t - 1, s - 1])
            if s >= 2 and ext[s] != blank and ext[s] != ext[s - 2]:
                a = _lse(a, alpha[t - 1, s - 2])
            alpha[t, s] = a + lp[t, ext[s]] if a != NEG else NEG
    ll = alpha[T - 1, L - 1]
    if L > 1:
        ll = _lse(ll, alpha[T - 1, L - 2])
    nll = -ll
    if not math.isfinite(nll):
        return math.inf, grad
    beta = np.full((T, L), NEG)
    beta[T - 1, L - 1] = lp[T - 1, blank]
    if L > 1:
        beta[T - 1, L - 2] = lp[T - 1, ext[L - 2]]
    for t in range(T - 2, -1, -1):
        for s in range(L):
            b = beta[t + 1, s]
            if s + 1 < L:
                b = _lse(b, beta[t + 1, s + 1])
            if s + 2 < L and ext[s + 2] != blank and ext[s + 2] != ext[s]:
                b = _lse(b, beta[t + 1, s + 2])
            beta[t, s] = b + lp[t, ext[s]] if b != NEG else NEG
    for t in range(T):
        occ = np.zeros(lp.shape[1])
        for s in range(L):
            ab = alpha[t, s] + beta[t, s]
            if ab != NEG:
                occ[ext[s]] += math.exp(ab + nll - lp[t, ext[s]])
        grad[t] = np.exp(lp[t]) - occ
    return nll, grad


def ctc_loss_numpy(log_probs_ntc, targets, input_lengths, target_lengths, blank=0, zero_infinity=True):
    """mean reduction of torch.nn.CTCLoss: mean_n(nll_n / max(S_n, 1)); inf -> 0 (+ zero grad)."""
    lp = np.asarray(log_probs_ntc, dtype=np.float64)
    N = lp.shape[0]
    grad = np.zeros_like(lp)
    total = 0.0
    nlls = []
    for n in range(N):
        S = int(target_lengths[n])
        nll, g = ctc_nll_and_grad_numpy(lp[n], np.asarray(targets[n][:S]), int(input_lengths[n]), blank)
        if not math.isfinite(nll):
            if zero_infinity:
                nll, g = 0.0, np.zeros_like(g)
        nlls.append(nll)
        total += nll / max(S, 1)
        grad[n] = g / (max(S, 1) * N)
    return total / N, grad, np.asarray(nlls)


# --------------------------------------------------------------------------------------------
# Greedy decoding (decoder.py:89-145)
# --------------------------------------------------------------------------------------------
def greedy_argmax(probs):
    """decoder.py:136 ``torch.max(probs, 2)``: first index wins ties, NaN counts as the maximum."""
    p = np.asarray(probs)
    N, T, C = p.shape
    out = np.zeros((N, T), dtype=np.int64)
    for n in range(N):
        for t in range(T):
            best, bi = p[n, t, 0], 0
            for c in range(1, C):
                v = p[n, t, c]
                if best != best:           # NaN already holds the maximum
                    break
                if v != v or v > best:
                    best, bi = v, c
            out[n, t] = bi
    return out


def greedy_collapse(argmax_nt, sizes=None, blank=0):
    """decoder.py:104-119 with remove_repetitions=True: drop blanks, drop a symbol equal to the *previous
    frame's raw argmax*.  Returns (tokens list per utterance, frame offsets list per utterance)."""
    toks, offs = [], []
    for n in range(len(argmax_nt)):
        seq = argmax_nt[n]
        size = int(sizes[n]) if sizes is not None else len(seq)
        tk, of = [], []
        for i in range(size):
            c = int(seq[i])
            if c != blank and not (i != 0 and c == int(seq[i - 1])):
                tk.append(c)
                of.append(i)
        toks.append(tk)
        offs.append(of)
    return toks, offs


def greedy_decode(probs, sizes=None, labels=ENGLISH_LOWERCASE, blank=0):
    """GreedyDecoder.decode (decoder.py:121-145): strings + offsets."""
    am = greedy_argmax(np.asarray(probs))
    toks, offs = greedy_collapse(am, sizes, blank)
    return ["".join(labels[c] for c in tk) for tk in toks], offs


# --------------------------------------------------------------------------------------------
# Novograd (novograd.py:52-114), betas=(0.95, 0) default, no amsgrad
# --------------------------------------------------------------------------------------------
def novograd_step(params, grads, state, lr=1e-3, betas=(0.95, 0.0), eps=1e-8, weight_decay=0.0,
                  grad_averaging=False, amsgrad=False):
    """In-place on ``params`` (list of tensors); ``state`` is a list of dicts carried across steps."""
    b1, b2 = betas
    for p, g, st in zip(params, grads, state):
        if not st:
            st["exp_avg"] = torch.zeros_like(p)
            st["exp_avg_sq"] = torch.zeros((), dtype=p.dtype)
        norm = torch.sum(g * g)
        if st["exp_avg_sq"] == 0:
            st["exp_avg_sq"] = norm.clone()
        else:
            st["exp_avg_sq"] = st["exp_avg_sq"] * b2 + (1 - b2) * norm
        second = st["exp_avg_sq"]
        if amsgrad:                                   # novograd.py:98-102: running maximum of the second moment
            st["max_exp_avg_sq"] = torch.maximum(st.get("max_exp_avg_sq", torch.zeros((), dtype=p.dtype)), second)
            second = st["max_exp_avg_sq"]
        g = g / (second.sqrt() + eps)
        if weight_decay != 0:
            g = g + weight_decay * p
        if grad_averaging:
            g = g * (1 - b1)
        st["exp_avg"].mul_(b1).add_(g)
        p.add_(st["exp_avg"], alpha=-lr)


# --------------------------------------------------------------------------------------------
# synthetic batch: the collated-batch layout of data_loader.py:149-158
# --------------------------------------------------------------------------------------------
def synthetic_batch(B, seconds, seed=0, n_labels=29, ragged=False, feat=64):
    g = torch.Generator().manual_seed(seed)
    T = 1 + 100 * seconds
    S = 15 * seconds
    x = torch.randn(B, feat, T, generator=g)
    targets = torch.randint(1, n_labels, (B, S), generator=g, dtype=torch.int32)
    if ragged:
        il = torch.randint(int(0.6 * T), T + 1, (B,), generator=g, dtype=torch.int32)
        il[0] = T
        tl = torch.randint(S // 2, S + 1, (B,), generator=g, dtype=torch.int32)
        for n in range(B):
            x[n, :, il[n]:] = 0
            targets[n, tl[n]:] = 0
    else:
        il = torch.full((B,), T, dtype=torch.int32)
        tl = torch.full((B,), S, dtype=torch.int32)
    return x, il, targets, tl


# --------------------------------------------------------------------------------------------
# The same Wav2Letter forward with the implementation's STATED storage precision emulated on the CPU:
# bf16 operands / fp32 accumulation, activations and activation-gradients stored as bf16 (DESIGN.md "precision").
# Used as the fairness yard-stick next to the fp32 oracle: a CUDA result must match this tightly, and may differ
# from the fp32 reference only by about as much as this emulation itself does.
# --------------------------------------------------------------------------------------------
class _RoundBF16(torch.autograd.Function):
    """value and gradient both pass through a bf16 store"""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


class _RoundGradBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def _bf16_weight(w):
    return w + (w.to(torch.bfloat16).to(w.dtype) - w).detach()       # bf16 shadow forward, fp32 master gradient


def w2l_forward_bf16emu(x, input_lengths, sd, specs, training=True):
    h = x.to(torch.bfloat16).float()
    for i, spec in enumerate(specs):
        p = "conv1ds.conv1d_%d." % i
        pl, pr = reflect_pad_amounts(spec["cin"], spec["k"], spec["stride"], spec["dilation"])
        if i == 0 and pl + pr > 0:
            h = F.pad(h, (pl, pr), mode="reflect")
        w = _bf16_weight(sd[p + "conv1.weight"])
        if not spec["bn"]:
            z = F.conv1d(h, w, sd[p + "conv1.bias"], stride=spec["stride"], dilation=spec["dilation"])
            z = _RoundGradBF16.apply(z)
            h = z
            continue
        z = _RoundBF16.apply(F.conv1d(h, w, None, stride=spec["stride"], dilation=spec["dilation"]))
        if training:
            mean = z.mean((0, 2), keepdim=True)
            var = z.var((0, 2), unbiased=False, keepdim=True)
        else:
            mean = (sd[p + "batch_norm.running_mean"] - sd[p + "conv1.bias"]).view(1, -1, 1)
            var = sd[p + "batch_norm.running_var"].view(1, -1, 1)
        y = (z - mean) * torch.rsqrt(var + 1e-3) * sd[p + "batch_norm.weight"].view(1, -1, 1) + sd[p + "batch_norm.bias"].view(1, -1, 1)
        if spec["act"]:
            y = torch.clamp(y, 0, 20)
        if i + 1 < len(specs):
            nxt = specs[i + 1]
            npl, npr = reflect_pad_amounts(nxt["cin"], nxt["k"], nxt["stride"], nxt["dilation"])
            if npl + npr > 0:
                y = F.pad(y, (npl, npr), mode="reflect")
        h = _RoundBF16.apply(y)
    lp = F.log_softmax(h.transpose(1, 2), dim=-1)
    scaling = int(np.prod([s["stride"] for s in specs]))
    return lp, (None if input_lengths is None else input_lengths // scaling)


# ------------------------------------------------------------------------------------------------ strided layers: unfold / fold
# Restatement (numpy loops, the same index arithmetic the CUDA kernels use) of the layout steps around a strided layer that is
# not the first one: the strided Conv1d of wav2letter.py:35-36 / jasper.py:96-105 equals a k=1 contraction of the unfolded rows
# with the weights stored [Cout, k, Cin]; backward-data is the fold (adjoint) of the column gradient.
def im2col_tm(x, T_out, k, stride, dilation, pad_left):
    """x [B, rows, C] (time-major) -> [B, T_out, k*C], out[b, t, j*C + c] = x[b, t*stride + j*dilation - pad_left, c] (0 outside)."""
    x = np.asarray(x)
    B, rows, C = x.shape
    out = np.zeros((B, T_out, k * C), dtype=x.dtype)
    for t in range(T_out):
        for j in range(k):
            r = t * stride + j * dilation - pad_left
            if 0 <= r < rows:
                out[:, t, j * C:(j + 1) * C] = x[:, r, :]
    return out


def col2im_tm(dcol, x_rows, C, k, stride, dilation, pad_left):
    """adjoint of ``im2col_tm`` in gather form: for every input row the taps j with r + pad_left - j*dilation = t*stride."""
    dcol = np.asarray(dcol)
    B, T_out, _ = dcol.shape
    dx = np.zeros((B, x_rows, C), dtype=np.float64)
    for r in range(x_rows):
        for j in range(k):
            u = r + pad_left - j * dilation
            if u < 0:
                break
            t = u // stride
            if t * stride != u or t >= T_out:
                continue
            dx[:, r, :] += dcol[:, t, j * C:(j + 1) * C]
    return dx


def depthwise_dgrad_strided(dy, w_kc, x_rows, stride, dilation, pad, dy_lens=None):
    """dy [B, T_out, C], w [k, C] -> dx [B, x_rows, C]: backward-data of a depthwise conv with stride >= 1 (gather form)."""
    dy, w_kc = np.asarray(dy, dtype=np.float64), np.asarray(w_kc, dtype=np.float64)
    B, T_out, C = dy.shape
    k = w_kc.shape[0]
    dx = np.zeros((B, x_rows, C), dtype=np.float64)
    for b in range(B):
        lim = T_out if dy_lens is None else min(T_out, max(0, int(dy_lens[b])))
        for u in range(x_rows):
            for j in range(k):
                v = u + pad - j * dilation
                if v < 0:
                    break
                t = v // stride
                if t * stride != v or t >= lim:
                    continue
                dx[b, u] += dy[b, t] * w_kc[j]
    return dx


# ------------------------------------------------------------------------------------------------ feature front-end
def mel_filterbank_slaney(sample_rate, n_fft, n_mels, fmin=0.0, fmax=None):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with its defaults (htk=False, norm='slaney'), the call at
    data/data_loader.py:40-44.  librosa is NOT installed here: this restates its published algorithm (Slaney's Auditory
    Toolbox mel scale: linear 200/3 Hz per mel below 1 kHz, log spacing ln(6.4)/27 above; triangles between neighbouring
    centre frequencies; each filter scaled by 2 / bandwidth).  PARITY UNPINNED for the filter weights themselves."""
    fmax = sample_rate / 2.0 if fmax is None else fmax
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0

    def hz_to_mel(f):
        return min_log_mel + np.log(f / min_log_hz) / logstep if f >= min_log_hz else f / f_sp

    def mel_to_hz(m):
        return min_log_hz * np.exp(logstep * (m - min_log_mel)) if m >= min_log_mel else f_sp * m

    n_bins = 1 + n_fft // 2
    freqs = [k * (sample_rate / 2.0) / (n_bins - 1) for k in range(n_bins)]
    lo, hi = hz_to_mel(fmin), hz_to_mel(fmax)
    centres = [mel_to_hz(lo + (hi - lo) * i / (n_mels + 1)) for i in range(n_mels + 2)]
    fb = np.zeros((n_mels, n_bins), dtype=np.float64)
    for i in range(n_mels):
        left, mid, right = centres[i], centres[i + 1], centres[i + 2]
        for k, f in enumerate(freqs):
            fb[i, k] = max(0.0, min((f - left) / (mid - left), (right - f) / (right - mid))) * 2.0 / (right - left)
    return fb.astype(np.float32)


def spectrogram_extract(signal, sample_rate=16000, window_size=0.02, window_stride=0.01, window="hamming", n_mels=64, noise=None,
                        fb=None):
    """SpectrogramExtractor._get_spect + extract (data/data_loader.py:64-88) on the CPU: dither with the given standard-normal
    ``noise`` (the reference draws torch.randn(audio.shape)), pre-emphasis 0.97, torch.stft(center=True), magnitude -> power,
    mel filterbank, log1p(. + 2^-24), per-feature (v - mean) / (std_unbiased + 1e-5).  Returns [n_mels, T] fp32."""
    win, hop = int(sample_rate * window_size), int(sample_rate * window_stride)
    n_fft = 2 ** int(np.ceil(np.log2(win)))
    wfn = {"hann": torch.hann_window, "hamming": torch.hamming_window, "blackman": torch.blackman_window, "bartlett": torch.bartlett_window}[window]
    x = torch.as_tensor(np.asarray(signal), dtype=torch.float32)
    if noise is not None:
        x = x + torch.as_tensor(np.asarray(noise), dtype=torch.float32) * 1e-5
    x = torch.cat((x[0].unsqueeze(0), x[1:] - 0.97 * x[:-1]), dim=0)
    X = torch.view_as_real(torch.stft(x, n_fft=n_fft, hop_length=hop, win_length=win, center=True,
                                      window=wfn(win, periodic=False).float(), return_complex=True))
    mag = torch.sqrt(X.pow(2).sum(-1))
    power = mag.pow(2)
    fbt = torch.as_tensor(mel_filterbank_slaney(sample_rate, n_fft, n_mels, 0.0, sample_rate / 2) if fb is None else fb, dtype=torch.float32)
    spect = torch.log1p(torch.matmul(fbt, power) + 2 ** -24)
    mean = spect.mean(dim=1, keepdim=True)
    std = spect.std(dim=1, keepdim=True) + 1e-5
    return (spect - mean) / std


def collate_features(feats):
    """_collator (data/data_loader.py:149-158) for the inputs: zero padding to the longest utterance + lengths."""
    T = max(f.shape[1] for f in feats)
    out = torch.zeros((len(feats), feats[0].shape[0], T), dtype=torch.float32)
    for i, f in enumerate(feats):
        out[i, :, :f.shape[1]] = f
    return out, torch.tensor([f.shape[1] for f in feats], dtype=torch.int32)


# ------------------------------------------------------------------------------------------------ prefix beam search
def prefix_beam_search(ctc, labels, blank_index=0, lm=None, k=5, alpha=0.3, beta=5, prune=0.001, end_char=">"):
    """decoder.py:147-231 restated with plain dicts (small cases only: pure Python).  Returns (best prefix, its score).
    Kept faithful where it matters for exact agreement: float64 arithmetic in the reference's association order (the reference
    vstack()s a float64 zero frame in front, which promotes any input to float64), Counter semantics (missing key = 0, `+=`
    inserts, Counter addition keeps Pb's keys first and drops non-positive sums), stable descending sort, and the word-count
    prior (len(re.findall(r'\\w+[\\s|>]', l)) + 1) ** beta."""
    import re
    ctc = np.asarray(ctc, dtype=np.float64)
    T, F = ctc.shape
    lm = (lambda l: 1) if lm is None else lm
    words = lambda l: len(re.findall(r"\w+[\s|>]", l))
    blank_char = labels[blank_index]
    pb_prev, pnb_prev, beam = {"": 1}, {"": 0}, [""]
    best_score = 0.0
    for t in range(T):
        row = ctc[t]
        alphabet = [labels[i] for i in range(F) if row[i] > prune]
        pb, pnb = {}, {}

        def add(d, key, v):
            d[key] = d.get(key, 0) + v
        for l in beam:
            if len(l) > 0 and l[-1] == end_char:
                pb[l] = pb_prev.get(l, 0)
                pnb[l] = pnb_prev.get(l, 0)
                continue
            for c in alphabet:
                ci = labels.index(c)
                if c == blank_char:
                    add(pb, l, row[blank_index] * (pb_prev.get(l, 0) + pnb_prev.get(l, 0)))
                    continue
                lp = l + c
                if len(l) > 0 and c == l[-1]:
                    add(pnb, lp, row[ci] * pb_prev.get(l, 0))
                    add(pnb, l, row[ci] * pnb_prev.get(l, 0))
                elif len(l.replace(" ", "")) > 0 and c in (" ", end_char):
                    lm_prob = lm(lp.strip(" " + end_char)) ** alpha
                    add(pnb, lp, lm_prob * row[ci] * (pb_prev.get(l, 0) + pnb_prev.get(l, 0)))
                else:
                    add(pnb, lp, row[ci] * (pb_prev.get(l, 0) + pnb_prev.get(l, 0)))
                if lp not in beam:
                    add(pb, lp, row[blank_index] * (pb_prev.get(lp, 0) + pnb_prev.get(lp, 0)))
                    add(pnb, lp, row[ci] * pnb_prev.get(lp, 0))
        nxt = {}
        for key, v in pb.items():
            s = v + pnb.get(key, 0)
            if s > 0:
                nxt[key] = s
        for key, v in pnb.items():
            if key not in pb and v > 0:
                nxt[key] = v
        score = lambda l: nxt[l] * (words(l) + 1) ** beta
        beam = sorted(nxt, key=score, reverse=True)[:k]
        best_score = score(beam[0]) if beam else 0.0
        pb_prev, pnb_prev = pb, pnb
    if not beam:
        beam = [""]
    return beam[0], float(best_score)
