"""Wav2Letter with the reference's constructor / forward / attribute / state_dict surface (wav2letter.py:12-92),
running on the sm_100a kernels.

    model = Wav2Letter(cfg.model)                       # cfg.model as composed from the Hydra yamls
    log_probs, out_lens = model(x, input_lengths)       # x [B, F, T] fp32 CUDA  ->  [B, T', n_labels] fp32

Per Conv1dBlock the reference runs ReflectionPad1d -> Conv1d(bias) -> BatchNorm1d(eps 1e-3, momentum 0.9) ->
Dropout -> clamp(0, 20) as separate library kernels.  Here a block is: tcgen05 implicit-GEMM conv over time-major
bf16 -> BN statistics -> one pass applying BN/dropout/clamp that writes directly into the NEXT block's
reflection-padded input (so no standalone pad kernel or NCW<->time-major transposes exist inside the stack)."""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from . import functional as F
from .base_asr_models import ConvCTCASR
from .layers import BatchNormParams, ConvBNActFn, ConvHeadFn, ConvParams, UnfoldTmFn, conv_bn_act_eval


def reflect_padding(input_channels, kernel, stride, dilation):
    """(left, right) of the reference's padding rule -- computed from the CHANNEL count (wav2letter.py:24-34)."""
    out_rows = (input_channels + stride - 1) // stride
    total = max(0, (out_rows - 1) * stride + (kernel - 1) * dilation + 1 - input_channels)
    return total // 2, (total + 1) // 2


class Conv1dBlock(nn.Module):
    def __init__(self, input_channels, output_channels, kernel_size, stride, drop_out_prob=-1.0, dilation=1, bn=True,
                 activation_use=True):
        super().__init__()
        self.input_channels, self.output_channels = input_channels, output_channels
        self.kernel_size, self.stride, self.dilation = kernel_size, stride, dilation
        self.drop_out_prob, self.activation_use = drop_out_prob, activation_use
        k = kernel_size[0]
        self.padding = k
        self.pad_lr = reflect_padding(input_channels, k, stride, dilation)
        self.padding_rows = sum(self.pad_lr)
        self.conv1 = ConvParams(input_channels, output_channels, k, stride=stride, dilation=dilation, bias=True, unfold=stride > 1)
        self.batch_norm = BatchNormParams(output_channels, eps=0.001, momentum=0.9) if bn else nn.Identity()
        self.has_bn = bn
        self.conv1.is_head = not bn  # the bias-only label head
        self.next_pad = (0, 0)       # reflect halo the consumer block wants; set by the owning model

    # ---- geometry
    def out_rows(self, t_in):
        k, s, d = self.kernel_size[0], self.stride, self.dilation
        return (t_in + self.padding_rows - d * (k - 1) - 1) // s + 1

    @property
    def drop_p(self):
        return float(self.drop_out_prob) if self.drop_out_prob != -1 and self.drop_out_prob > 0 else 0.0

    # ---- time-major fast path (used by Wav2Letter.forward)
    def forward_tm(self, xin, t_in, from_ncw, head_mode=0):
        """xin: NCW fp32 [B,C,T] when ``from_ncw`` (first block) else time-major bf16 already carrying this block's
        reflection halo ([B, pl+T+pr, C]).  Returns the next block's padded input (or, for the head, fp32 scores [B,T',labels]:
        log-probs when ``head_mode`` is 0, raw logits when 2)."""
        conv = self.conv1
        t_out = self.out_rows(t_in)
        pl, pr = self.pad_lr
        if from_ncw:
            Fp = conv.phys(conv.in_channels)
            if Fp != xin.shape[1]:                        # feature counts that are not a multiple of 8 (161 STFT bins) or below 64: zero rows
                xin = torch.nn.functional.pad(xin, (0, 0, 0, Fp - xin.shape[1]))
            if conv.unfold:
                xin = F.im2col_ncw(xin, t_out, self.kernel_size[0], self.stride, self.dilation, pl, F.PAD_REFLECT, out_dtype=conv.act_dtype)
            else:
                xin = F.im2col_ncw(xin, t_in + pl + pr, 1, 1, 1, pl, F.PAD_REFLECT, out_dtype=conv.act_dtype)
        elif conv.unfold:                                 # strided block inside the stack: unfold the halo-carrying input
            if conv.f32:
                raise NotImplementedError("precision='tf32': a strided block beyond the first layer is only implemented for bf16 activations")
            xin = UnfoldTmFn.apply(xin, t_out, self.kernel_size[0], self.stride, self.dilation, 0)
        if not self.has_bn:
            if self.activation_use or self.drop_p > 0:
                raise NotImplementedError("a block without BatchNorm is implemented as the bias-only head (activation_use=False, no dropout)")
            return ConvHeadFn.apply(xin, conv.weight, conv.bias, conv, head_mode), t_out
        geo = {"T_out": t_out, "x_row_offset": 0, "out_pad": self.next_pad, "drop_p": self.drop_p if self.training else 0.0,
               "act": F.ACT_CLAMP20 if self.activation_use else F.ACT_NONE}
        if self.training:
            bn = self.batch_norm
            y = ConvBNActFn.apply(xin, conv.weight, conv.bias, bn.weight, bn.bias, None, None, conv, bn, geo)
        else:
            y = conv_bn_act_eval(xin, conv, self.batch_norm, geo)
        return y, t_out

    # ---- reference-shaped standalone call: [B, C, T] fp32 in, [B, C', T'] fp32 out (wav2letter.py:40-47)
    def forward(self, xs):
        saved, self.next_pad = self.next_pad, (0, 0)
        try:
            y, t_out = self.forward_tm(xs, xs.shape[2], from_ncw=True, head_mode=2)
        finally:
            self.next_pad = saved
        if not self.has_bn:                  # the head on its own returns conv + bias (wav2letter.py:40-47 with bn=False, no clamp)
            return y.transpose(1, 2)         # [B, labels, T'] view of the fp32 logits
        return _TmToNcw.apply(y, t_out, self.output_channels)


class _TmToNcw(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, t, c):
        ctx.shape, ctx.f32 = y.shape, y.dtype == torch.float32
        return F.tm_to_ncw(y, t, c)

    @staticmethod
    def backward(ctx, g):
        B, C, T = g.shape
        if ctx.shape[2] != C:                             # padded width: the surplus channels of the buffer have no gradient
            out = torch.zeros(ctx.shape, dtype=torch.float32 if ctx.f32 else torch.bfloat16, device=g.device)
            out[:, :, :C] = g.transpose(1, 2)
            return out, None, None
        if ctx.f32:
            return g.transpose(1, 2).contiguous().float(), None, None
        out = torch.empty(ctx.shape, dtype=torch.bfloat16, device=g.device)
        F._lib.check(F._lib.load().w2l_ncw_to_tm(F._ptr(g.contiguous().float()), F._ptr(out), B, C, T, F._stream()), "ncw_to_tm")
        return out, None, None


class Wav2Letter(ConvCTCASR):
    def __init__(self, cfg):
        super().__init__(cfg)
        self.mid_layers = cfg.mid_layers
        if not cfg.input_size:
            nfft = self.audio_conf["sample_rate"] * self.audio_conf["window_size"]
            self.input_size = int(1 + nfft / 2)
        else:
            self.input_size = cfg.input_size
        width = self.input_size
        blocks = []
        for idx, lp in enumerate(cfg.layers[: self.mid_layers]):
            blocks.append(("conv1d_%d" % idx, Conv1dBlock(width, lp.output_size, (lp.kernel_size,), lp.stride, dilation=lp.dilation,
                                                          drop_out_prob=lp.dropout)))
            width = lp.output_size
        blocks.append(("conv1d_%d" % len(blocks), Conv1dBlock(width, len(self.labels), (1,), 1, bn=False, activation_use=False)))
        self.conv1ds = nn.Sequential(OrderedDict(blocks))
        mods = list(self.conv1ds.children())
        for cur, nxt in zip(mods[:-1], mods[1:]):
            cur.next_pad = nxt.pad_lr
        # precision: "bf16" (default: bf16 operands and activations, fp32 accumulation and statistics) or "tf32" -- the fp32-faithful
        # mode: activations, weights and gradients stay fp32 in memory (the reference's nn.Conv1d arithmetic, wav2letter.py:35-36),
        # the GEMMs multiply them as tf32 with fp32 accumulation.  Logits within 1.5e-3 rel-L2 of torch fp32 per layer.
        self.precision = str(getattr(cfg, "precision", "bf16") or "bf16").lower()
        if self.precision in ("fp32", "float32"):
            self.precision = "tf32"
        if self.precision not in ("bf16", "tf32"):
            raise ValueError("Wav2Letter: precision must be 'bf16' or 'tf32', got %r" % (self.precision,))
        for m in mods:
            m.conv1.f32 = self.precision == "tf32"

    @property
    def scaling_factor(self):
        if not hasattr(self, "_scaling_factor"):
            self._scaling_factor = int(np.prod([m.conv1.stride[0] for m in self.conv1ds.children()]))
        return self._scaling_factor

    def forward(self, x, input_lengths=None):
        F._need_cuda(x)                                  # RuntimeError on a CPU tensor: this build has no CPU path
        t = x.shape[2]
        h, first = x, True
        tap = getattr(self, "_tap", None)                # parity instrumentation (tests/_layerwise.py): every block's input
        for block in self.conv1ds.children():
            if tap is not None:
                tap.append(h)
            h, t = block.forward_tm(h, t, from_ncw=first)
            if tap is not None and h.requires_grad:
                h.retain_grad()
            first = False
        out_lens = self.compute_output_lengths(input_lengths) if input_lengths is not None else None
        return h, out_lens
