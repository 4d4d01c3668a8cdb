"""Jasper with the reference's constructor / forward / state_dict surface (jasper.py:22-475) on the sm_100a kernels.

What ``Jasper._build_encoder`` can reach (jasper.py:436-453) is supported: batch normalisation, ReLU, 'add' residual
through a 1x1 conv + BN, groups=1, length masking, ``repeat`` sub-blocks, dense convolutions.  Time-major execution:

  * masked_fill(t >= len, 0) before every MaskedConv1d (jasper.py:116-119) is folded into the PRODUCER's fused
    BN/ReLU pass (rows past the utterance are written as zeros), and into the unfold pass for the raw features;
  * 'same' zero padding (jasper.py:61-66) costs nothing: the conv kernel starts its taps at row -p and TMA zero-fills
    rows outside the tensor;
  * conv -> BN -> (+ BN(conv1x1(block input))) -> ReLU -> dropout is one GEMM plus one fused elementwise pass
    (two GEMMs with a residual), instead of the reference's 8-10 separate library kernels per sub-block.

Separable sub-blocks (the shipped model/jasper.yaml: depthwise k-tap + pointwise 1x1, jasper.py:318-341) run the depthwise
half on a CUDA-core kernel (csrc/depthwise.cu) and the pointwise half on the tensor-core GEMM.
Not implemented (raise NotImplementedError at construction): group/instance/layer norm, groups > 1 with shuffle, heads,
residual_mode='max', dense residuals -- none of which Jasper._build_encoder can produce."""
import numpy as np
import torch
import torch.nn as nn

from . import functional as F
from .base_asr_models import ConvCTCASR
from .layers import (BatchNormParams, ConvBNActFn, ConvHeadFn, ConvParams, DepthwiseFn, DepthwiseParams, ResidualBranchFn,
                     UnfoldTmFn, conv_bn_act_eval, conv_desc, pad_channels)

jasper_activations = {"hardtanh": nn.Hardtanh, "relu": nn.ReLU, "selu": nn.SELU}


def compute_new_kernel_size(kernel_size, kernel_width):
    """jasper.py:53-58: scale, then bump even sizes to the next odd one."""
    k = max(int(kernel_size * kernel_width), 1)
    return k + 1 if k % 2 == 0 else k


def get_same_padding(kernel_size, stride, dilation):
    """jasper.py:61-66."""
    if stride > 1 and dilation > 1:
        raise ValueError("Only stride OR dilation may be greater than 1")
    if dilation > 1:
        return (dilation * kernel_size) // 2 - 1
    return kernel_size // 2


def init_weights(m, mode="xavier_uniform"):
    """jasper.py:29-50 applied to this package's parameter containers (same RNG consumption: the values are drawn
    into a reference-shaped contiguous tensor and copied into the kernel-layout storage)."""
    if isinstance(m, MaskedConv1d):
        init_weights(m.conv, mode)
    if isinstance(m, (ConvParams, DepthwiseParams)):
        w = torch.empty(m.out_channels, m.in_channels // m.groups, m.kernel_size[0])
        if mode == "xavier_uniform":
            nn.init.xavier_uniform_(w, gain=1.0)
        elif mode == "xavier_normal":
            nn.init.xavier_normal_(w, gain=1.0)
        elif mode == "kaiming_uniform":
            nn.init.kaiming_uniform_(w, nonlinearity="relu")
        elif mode == "kaiming_normal":
            nn.init.kaiming_normal_(w, nonlinearity="relu")
        else:
            raise ValueError("Unknown Initialization mode: {0}".format(mode))
        with torch.no_grad():
            m.weight.copy_(w)
    elif isinstance(m, BatchNormParams):
        with torch.no_grad():
            m.running_mean.zero_()
            m.running_var.fill_(1)
            m.num_batches_tracked.zero_()
            m.weight.fill_(1.0)
            m.bias.zero_()


class MaskedConv1d(nn.Module):
    """Parameter holder with the reference's layout (``.conv.weight``); the masking itself is fused upstream."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, heads=-1, bias=False,
                 use_mask=True):
        super().__init__()
        if not (heads == -1 or groups == in_channels):
            raise ValueError("Only use heads for depthwise convolutions")
        if heads != -1 or groups not in (1, in_channels) or (groups != 1 and (out_channels != in_channels or bias)):
            raise NotImplementedError("only dense (groups=1) and depthwise (groups=channels) MaskedConv1d are implemented")
        self.real_out_channels = out_channels
        if groups == 1:
            self.conv = ConvParams(in_channels, out_channels, kernel_size, stride=stride, padding=padding, dilation=dilation, bias=bias,
                                   unfold=stride > 1)
        else:
            self.conv = DepthwiseParams(in_channels, kernel_size, stride=stride, padding=padding, dilation=dilation)
        self.use_mask = use_mask
        self.heads = heads
        if not use_mask:
            # conv_mask=False: the reference builds a bare nn.Conv1d here (jasper.py:289-298), so its checkpoints say
            # `mconv.N.weight` where this holder has `mconv.N.conv.weight` -- speak the reference's keys on both ways
            self._register_state_dict_hook(self._flat_keys_out)
            self._register_load_state_dict_pre_hook(self._flat_keys_in)

    @staticmethod
    def _flat_keys_out(module, state_dict, prefix, local_metadata):
        for name in ("weight", "bias"):
            if prefix + "conv." + name in state_dict:
                state_dict[prefix + name] = state_dict.pop(prefix + "conv." + name)

    @staticmethod
    def _flat_keys_in(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        for name in ("weight", "bias"):
            if prefix + name in state_dict:
                state_dict[prefix + "conv." + name] = state_dict.pop(prefix + name)

    def get_seq_len(self, lens):
        c = self.conv
        return (lens + 2 * c.padding[0] - c.dilation[0] * (c.kernel_size[0] - 1) - 1) / c.stride[0] + 1       # true division


class JasperBlock(nn.Module):
    def __init__(self, inplanes, planes, repeat=3, kernel_size=11, kernel_size_factor=1, stride=1, dilation=1, padding="same",
                 dropout=0, activation=None, residual=True, groups=1, separable=False, heads=-1, normalization="batch",
                 norm_groups=1, residual_mode="add", residual_panes=[], conv_mask=False):
        super().__init__()
        if padding != "same":
            raise ValueError("currently only 'same' padding is supported")
        if normalization != "batch":
            raise NotImplementedError("only batch normalisation is reachable from Jasper._build_encoder and implemented")
        if groups != 1 or heads != -1:
            raise NotImplementedError("grouped (shuffled) Jasper sub-blocks are not reachable from Jasper._build_encoder and not implemented")
        if residual_mode != "add" or len(residual_panes) != 0:
            raise NotImplementedError("only the plain 'add' residual is implemented")
        kernel_size = compute_new_kernel_size(kernel_size, float(kernel_size_factor))
        pad = get_same_padding(kernel_size, stride, dilation)
        self.conv_mask, self.separable, self.residual_mode = conv_mask, separable, residual_mode
        self.dropout_p = float(dropout)
        if activation is None:
            activation = nn.Hardtanh(min_val=0.0, max_val=20.0)
        if isinstance(activation, nn.ReLU):
            self.act = F.ACT_RELU
        elif isinstance(activation, nn.Hardtanh) and activation.min_val == 0.0 and activation.max_val == 20.0:
            self.act = F.ACT_CLAMP20
        else:
            raise NotImplementedError("activation %r is not implemented in the fused epilogues" % (activation,))
        if self.act == F.ACT_CLAMP20 and self.dropout_p > 0:
            # the reference applies activation THEN dropout (jasper.py:389-393); the fused pass applies dropout then activation,
            # which is the same thing for ReLU (what Jasper._build_encoder always passes) but not for a clamp at 20
            raise NotImplementedError("JasperBlock: Hardtanh(0, 20) with dropout > 0 is not implemented (use ReLU, or dropout=0)")
        # ModuleList positions mirror the reference so that checkpoints load:
        # [conv | depthwise, pointwise][bn][act, drop] * (repeat-1) + [conv | depthwise, pointwise][bn]
        mods, cin = [], inplanes
        for r in range(repeat):
            if separable:
                mods.append(MaskedConv1d(cin, cin, kernel_size, stride=stride, dilation=dilation, padding=pad, groups=cin, use_mask=conv_mask))
                mods.append(MaskedConv1d(cin, planes, 1, use_mask=conv_mask))
            else:
                mods.append(MaskedConv1d(cin, planes, kernel_size, stride=stride, dilation=dilation, padding=pad, use_mask=conv_mask))
            mods.append(BatchNormParams(planes, eps=1e-3, momentum=0.1))
            if r != repeat - 1:
                mods.extend([type(activation)() if not isinstance(activation, nn.Hardtanh) else nn.Hardtanh(0.0, 20.0), nn.Dropout(p=dropout)])
            cin = planes
        self.mconv = nn.ModuleList(mods)
        self.dense_residual = False
        if residual:
            self.res = nn.ModuleList([nn.ModuleList([MaskedConv1d(inplanes, planes, 1, use_mask=conv_mask),
                                                     BatchNormParams(planes, eps=1e-3, momentum=0.1)])])
        else:
            self.res = None
        self.mout = nn.Sequential(nn.ReLU() if self.act == F.ACT_RELU else nn.Hardtanh(0.0, 20.0), nn.Dropout(p=dropout))
        self.repeat, self.stride, self.pad = repeat, stride, pad

    def sub_blocks(self):
        """[(depthwise MaskedConv1d | None, conv MaskedConv1d, BatchNormParams)] in execution order"""
        out, step = [], 5 if self.separable else 4
        for i in range(0, len(self.mconv), step):
            if self.separable:
                out.append((self.mconv[i], self.mconv[i + 1], self.mconv[i + 2]))
            else:
                out.append((None, self.mconv[i], self.mconv[i + 1]))
        return out

    def chain_params(self):
        """(kernel, stride | 0 when unmasked, dilation, padding) of this block's MaskedConv1d modules in execution order -- the
        table ``F.lens_chain`` walks (the residual 1x1 conv masks with the block's INPUT lengths and hands nothing on)."""
        out = []
        for dwm, mc, _bn in self.sub_blocks():
            for m in ((dwm, mc) if dwm is not None else (mc,)):
                c = m.conv
                out.append((c.kernel_size[0], c.stride[0] if self.conv_mask else 0, c.dilation[0], c.padding[0]))
        return out

    def forward_tm(self, h, t, rows, ri, from_ncw, mask_output):
        """h: NCW fp32 (first block, ``from_ncw``) or time-major bf16 [B, t, C] whose rows >= len are already zero.
        ``rows`` [n+1, B] int32 = truncated lengths entering every masked conv of the encoder (``F.lens_chain``), ``ri`` = index of
        this block's first conv in it (``rows`` is None when no lengths were given).  Returns (time-major bf16 output, t_out, ri)."""
        subs = self.sub_blocks()
        first = subs[0][0] if self.separable else subs[0][1]
        if from_ncw:
            m0 = first.conv
            k, s, d, p = m0.kernel_size[0], m0.stride[0], m0.dilation[0], m0.padding[0]
            Fp = ConvParams.phys(m0.in_channels)
            if Fp != h.shape[1]:                         # feature counts that are not a multiple of 8 (161 STFT bins) or below 64: zero rows
                h = torch.nn.functional.pad(h, (0, 0, 0, Fp - h.shape[1]))
            mask = rows[ri] if (self.conv_mask and rows is not None) else None
            if not self.separable and m0.unfold:
                t_first = (t + 2 * p - d * (k - 1) - 1) // s + 1
                h = F.im2col_ncw(h, t_first, k, s, d, p, F.PAD_ZERO, mask, out_dtype=m0.act_dtype)   # masked, zero padded, unfolded
            else:
                h = F.im2col_ncw(h, t, 1, 1, 1, 0, F.PAD_ZERO, mask, out_dtype=getattr(m0, "act_dtype", torch.bfloat16))   # masked time-major copy
        block_in, res_pair = h, None
        training = self.training
        if self.res is not None:
            if self.stride != 1:                                                 # jasper.py:400-415: the 1x1 residual conv is not strided
                raise RuntimeError("JasperBlock: a strided block cannot carry a residual branch (the reference's `out + res_out` "
                                   "fails on the time dimension, jasper.py:412)")
            rconv, rbn = self.res[0][0].conv, self.res[0][1]
            if training:
                res_pair = ResidualBranchFn.apply(block_in, rconv.weight, rbn.weight, rbn.bias, rconv, rbn)
            else:
                zr = torch.empty((h.shape[0], t, rconv.cout_phys), dtype=rconv.act_dtype, device=h.device)
                F.conv1d_fwd(block_in, rconv.packed(), conv_desc(rconv, h.shape[0], t, t, 0), zr)
                rfold = rbn.eval_scale_shift(None)
                if rconv.cout_phys != rconv.out_channels:                        # padded width: the surplus channels come out as 0 * 0 + 0
                    rfold = (pad_channels(rfold[0], rconv.cout_phys), pad_channels(rfold[1], rconv.cout_phys))
                res_pair = (zr, rfold)

        tap = getattr(self, "_tap", None)           # parity instrumentation (tests/_layerwise.py): every sub-block's input
        for r, (dwm, mc, bn) in enumerate(subs):
            conv = mc.conv
            last = r == len(subs) - 1
            if tap is not None:
                if h.requires_grad:
                    h.retain_grad()
                tap.append((h, ri))
            if dwm is not None:                                                  # separable: depthwise k-tap, then pointwise 1x1
                dc = dwm.conv
                k, s, d, p = dc.kernel_size[0], dc.stride[0], dc.dilation[0], dc.padding[0]
                t_dw = (t + 2 * p - d * (k - 1) - 1) // s + 1
                ri += 1                                                          # lengths after the depthwise conv
                dmask = rows[ri] if (self.conv_mask and rows is not None) else None
                if training:
                    h = DepthwiseFn.apply(h, dc.weight, dc, t_dw, dmask)
                else:
                    h = F.depthwise_fwd(h, dc.storage_phys(), t_dw, k, s, d, p, dmask)
                t = t_dw
            if conv.unfold:
                if not (from_ncw and r == 0):                                    # (the encoder's first conv was unfolded from NCW above)
                    k, s, d, p = conv.kernel_size[0], conv.stride[0], conv.dilation[0], conv.padding[0]
                    h = UnfoldTmFn.apply(h, (t + 2 * p - d * (k - 1) - 1) // s + 1, k, s, d, p)
                t_out, x_off = h.shape[1], 0
            else:
                k, d, p = conv.kernel_size[0], conv.dilation[0], conv.padding[0]
                t_out, x_off = t + 2 * p - d * (k - 1), -p
            ri += 1                                                              # lengths after this conv (truncated, as its consumer sees them)
            out_mask = None
            if self.conv_mask and rows is not None and (not last or mask_output):
                out_mask = rows[ri]
            geo = {"T_out": t_out, "x_row_offset": x_off, "out_pad": (0, 0), "act": self.act,
                   "drop_p": self.dropout_p if training else 0.0, "lens": out_mask}
            use_res = res_pair if last else None
            if training:
                h = ConvBNActFn.apply(h, conv.weight, None, bn.weight, bn.bias, use_res[0] if use_res else None,
                                      use_res[1] if use_res else None, conv, bn, geo)
            else:
                h = conv_bn_act_eval(h, conv, bn, geo, res=use_res)
            t = t_out
        return h, t, ri


class Jasper(ConvCTCASR):
    def __init__(self, cfg):
        super().__init__(cfg)
        self.mid_layers = cfg.mid_layers
        if not cfg.input_size:
            nfft = self.audio_conf["sample_rate"] * self.audio_conf["window_size"]
            self.input_size = int(1 + nfft / 2)
        else:
            self.input_size = cfg.input_size
        self._build_encoder(cfg)
        last = self.jasper_encoder[-1].mconv[-1].num_features
        self.final_layer = nn.Sequential(ConvParams(last, len(self.labels), 1, bias=True))
        self.final_layer.apply(init_weights)
        self.final_layer[0].is_head = True
        # channel counts that are not multiples of 8 (or below 64) are padded internally like Wav2Letter's (ConvParams.phys: exact-zero
        # surplus channels in every buffer, logical shapes for parameters and gradients).  One combination is not wired: a strided dense
        # block beyond the first whose INPUT width is padded (its unfolded weights would need a per-tap padded backward-data copy).
        convs = [m for m in self.modules() if isinstance(m, ConvParams)]
        for m in convs[1:]:
            if m.unfold and m.cin_phys != m.cin_eff:
                raise NotImplementedError("Jasper: a strided dense block beyond the first with an input width that is not a multiple "
                                          "of 8 (got %d)" % m.in_channels)
        # precision: "bf16" (default) or "tf32" -- the fp32-faithful mode (fp32 activations / weights / gradients in memory, tf32
        # multiplies in the GEMMs, plain fp32 FMAs in the depthwise convs, fp32 accumulation; see Wav2Letter).  Not for strided inner
        # blocks: their time-major unfold exists for bf16 activations.
        self.precision = str(getattr(cfg, "precision", "bf16") or "bf16").lower()
        if self.precision in ("fp32", "float32"):
            self.precision = "tf32"
        if self.precision not in ("bf16", "tf32"):
            raise ValueError("Jasper: precision must be 'bf16' or 'tf32', got %r" % (self.precision,))
        if self.precision == "tf32":
            convs = [m for m in self.modules() if isinstance(m, ConvParams)]
            if any(c.unfold for c in convs[1:]) or any(m.stride[0] != 1 for m in self.modules() if isinstance(m, DepthwiseParams)
                                                       and m is not getattr(self.jasper_encoder[0].mconv[0], "conv", None)):
                raise NotImplementedError("Jasper: precision='tf32' is not implemented for strided blocks beyond the first")
            for c in convs + [m for m in self.modules() if isinstance(m, DepthwiseParams)]:
                c.f32 = True

    def _build_encoder(self, cfg):
        width, blocks = self.input_size, []
        for l in cfg.jasper_blocks[: cfg.mid_layers]:
            blocks.append(JasperBlock(inplanes=width, planes=l.layer_size, kernel_size=l.kernel_size, stride=l.get("stride", 1),
                                      dilation=l.get("dilation", 1), residual=l.residual, repeat=l.get("repeat", 1),
                                      conv_mask=l.get("conv_mask", True), separable=l.get("separable", True),
                                      activation=torch.nn.ReLU(), dropout=l.get("dropout", 0)))
            width = l.layer_size
        self.jasper_encoder = nn.Sequential(*blocks)
        self.jasper_encoder.apply(init_weights)

    @property
    def scaling_factor(self):
        if not hasattr(self, "_scaling_factor"):
            self._scaling_factor = int(np.prod([b.mconv[0].conv.stride[0] for b in self.jasper_encoder]))
        return self._scaling_factor

    def forward(self, xs, input_lengths):
        """[B, F, T] fp32 CUDA, lengths [B] -> ([B, T', n_labels] log-probs in training / probabilities in eval --
        the reference's behaviour, jasper.py:470-473 -- , output lengths int64 [B])."""
        F._need_cuda(xs)                                 # RuntimeError on a CPU tensor: this build has no CPU path
        blocks = list(self.jasper_encoder)
        rows, out_lens = None, None
        if input_lengths is not None:
            # every MaskedConv1d's length arithmetic (truncate, mask, (len + 2p - d(k-1) - 1) / stride + 1) in ONE launch
            if getattr(self, "_chain", None) is None:
                self._chain = [q for blk in blocks for q in blk.chain_params()]
            rows, out_lens = F.lens_chain(input_lengths.to(xs.device), self._chain)
        h, t, ri = xs, xs.shape[2], 0
        tap = getattr(self, "_tap", None)                # parity instrumentation (tests/_layerwise.py): block / sub-block inputs
        for i, blk in enumerate(blocks):
            if tap is not None:
                tap["hs"].append(h)
                blk._tap = []
            # the head (final_layer) is NOT masked in the reference (jasper.py:468): the last block keeps its padded rows
            h, t, ri = blk.forward_tm(h, t, rows, ri, from_ncw=(i == 0), mask_output=(i != len(blocks) - 1))
            if tap is not None:
                tap["taps"].append(blk.__dict__.pop("_tap"))
                if h.requires_grad:
                    h.retain_grad()
        head = self.final_layer[0]
        mode = getattr(self, "nan_check", "sync")
        flag = torch.zeros(1, dtype=torch.int32, device=xs.device) if mode != "off" else None
        scores = ConvHeadFn.apply(h, head.weight, head.bias, head, 0 if self.training else 1, flag)
        if tap is not None:
            tap["hs"].append(h)
            tap["rows"] = rows
            if scores.requires_grad:
                scores.retain_grad()
        self._assert_no_nan(flag, mode)
        return scores, out_lens

    # ---- jasper.py:474 `assert not (jasper_res != jasper_res).any()`: the softmax kernel raises a device flag; "sync" (default)
    # reads it here like the reference does (one device sync per forward), "deferred" copies it to pinned memory and raises at
    # the next forward / explicit check_nan() once the copy has landed (no sync on the training path), "off" skips the check
    nan_check = "sync"

    def _assert_no_nan(self, flag, mode):
        if mode == "off":
            return
        if F.capturing():
            # inside a CUDA graph capture (graph_step.py) nothing can be read back: the flag is a static operand of the graph, and
            # GraphedTrainStep copies it out after every replay and raises like "deferred" does
            self._nan_flag_graph = flag
            return
        if mode == "sync":
            assert not bool(flag.item())  # is there any NAN in result?
            return
        self.check_nan(block=False)
        host = torch.zeros(1, dtype=torch.int32).pin_memory()
        host.copy_(flag, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._nan_pending = (host, ev)

    def check_nan(self, block=True):
        pend = getattr(self, "_nan_pending", None)
        if pend is None:
            return
        host, ev = pend
        if block:
            ev.synchronize()
        elif not ev.query():
            return
        self._nan_pending = None
        assert not bool(host.item())  # is there any NAN in result?
