"""Building blocks shared by Wav2Letter and Jasper: parameter containers whose state_dict keys and shapes match the
reference's nn.Conv1d / nn.BatchNorm1d, and autograd Functions that run conv -> BatchNorm -> dropout -> activation
(-> reflection halo / length mask of the consumer) on the hand-written sm_100a kernels.

Layout contract between blocks: activations are time-major bf16 ``[B, rows, C]``; a producer writes the rows its
consumer's padding rule needs (Wav2Letter: mirrored halo rows; Jasper: none, zero padding comes from TMA OOB fill).
Conv weights keep the reference's parameter shape ``[Cout, Cin, k]`` but live in memory in the layout the GEMM
kernels read (``[k, Cout, Cin]``, or ``[Cout, k, Cin]`` for the unfolded strided first layer): the Parameter is a
permuted view, so packing a bf16 shadow is a plain cast and weight gradients need no re-layout."""
import math
import os

import torch
import torch.nn as nn

from . import functional as F

_seed_counter = [0]


class WgradStream:
    """Weight-gradient kernels run on a per-device side stream.  In backward the critical chain is
    dgrad_i -> BatchNorm-backward_{i-1} -> dgrad_{i-1}; wgrad_i only feeds the optimizer, so it is launched AFTER dgrad_i on
    the side stream and the (HBM-bound) BatchNorm-backward kernels of the next layer run beside it instead of serialising with
    the (tensor-bound) GEMMs.  The compute stream re-joins the side stream when the backward pass ends (engine callback), so
    ``.grad`` consumers need no extra care; ``GradientReducer`` enqueues its collectives behind the side stream."""
    enabled = os.environ.get("W2L_WGRAD_STREAM", "1") != "0"
    _side = {}
    _pending = {}
    _task = -1

    @classmethod
    def side(cls, device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        s = cls._side.get(idx)
        if s is None:
            s = cls._side[idx] = torch.cuda.Stream(device=idx)
        return s

    @classmethod
    def fork(cls, device, param=None):
        """Call after the producers of the wgrad operands were enqueued: the side stream waits for them.  Returns the side
        stream, or None when disabled / outside a backward pass / when ``param`` already holds a gradient -- autograd then ADDS
        the new gradient into it on the compute stream as soon as backward() returns (gradient accumulation,
        ``zero_grad(set_to_none=False)``), so that wgrad has to be ordered on the compute stream."""
        task = torch._C._current_graph_task_id()
        if not cls.enabled or task < 0 or (param is not None and param.grad is not None):
            return None
        main, side = torch.cuda.current_stream(device), cls.side(device)
        side.wait_stream(main)
        if task != cls._task:                          # first fork of this backward pass: join when the pass ends
            cls._task = task
            cls._pending.clear()
            torch.autograd.Variable._execution_engine.queue_callback(cls._join)
        cls._pending[(main.device.index, main.cuda_stream)] = main
        return side

    @classmethod
    def _join(cls):
        for main in cls._pending.values():
            main.wait_stream(cls.side(main.device))
        cls._pending.clear()


def wgrad_async(side, dy, x, desc, dw, conv=None):
    """conv1d_wgrad on the side stream returned by ``WgradStream.fork`` (plain call when it is None).  Returns the gradient in the
    Parameter's own [k_eff, Cout, cin_eff] layout: for a conv whose channel counts are padded internally (``conv.padded``) the logical
    part is cut out of the padded result on the same stream."""
    wgrad = F.conv1d_wgrad_t if desc.x_dtype == F.DT_F32 else F.conv1d_wgrad      # fp32-faithful mode: tf32 over transposed operands
    if side is None:
        wgrad(dy, x, desc, dw)
        return conv.unpad_dw(dw) if conv is not None and conv.padded else dw
    with torch.cuda.stream(side):
        wgrad(dy, x, desc, dw)
        out = conv.unpad_dw(dw) if conv is not None and conv.padded else dw
    for t in (dy, x) if getattr(dw, "_w2l_arena", False) else (dy, x, dw, out):
        t.record_stream(side)
    return out


def alloc_dw(conv, device, cout=None):
    """fp32 [k_eff, Cout, cin_eff] buffer for the weight gradient: the layer's slice of the data-parallel gradient arena when a
    ``PeerGradientReducer`` owns one (wgrad then writes where the NVLink all-reduce reads: no copy), else a fresh tensor."""
    buf = getattr(conv, "_grad_buffer", None)
    if buf is not None and conv.weight.grad is None and buf.device == device and not conv.padded:
        return buf
    rows = conv.out_channels if cout is None else cout            # (padded channel counts: the kernel writes the padded shape)
    return torch.empty((conv.k_eff, rows, conv.cin_phys), dtype=torch.float32, device=device)


def next_dropout_seed():
    _seed_counter[0] += 1
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + _seed_counter[0] * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF


class ConvParams(nn.Module):
    """Stands where the reference has an ``nn.Conv1d`` (wav2letter.py:35-36, jasper.py:96-105): same attribute
    names (weight, bias, kernel_size, stride, dilation, padding, in/out_channels) and state_dict entries."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, bias=True, unfold=False,
                 init="conv_default"):
        super().__init__()
        k = kernel_size[0] if isinstance(kernel_size, (tuple, list)) else kernel_size
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.dilation, self.padding = (k,), (stride,), (dilation,), (padding,)
        self.groups = 1
        self.unfold = bool(unfold)            # strided layer: input is unfolded (im2col) and the conv runs with k=1
        # reference-shaped init (same RNG consumption order as nn.Conv1d.reset_parameters), then re-layout
        w = torch.empty(out_channels, in_channels, k)
        if init == "xavier_uniform":          # jasper.py:29-35
            nn.init.kaiming_uniform_(w, a=math.sqrt(5))
            b0 = self._default_bias(w) if bias else None
            nn.init.xavier_uniform_(w, gain=1.0)
        else:
            nn.init.kaiming_uniform_(w, a=math.sqrt(5))
            b0 = self._default_bias(w) if bias else None
        perm = (0, 2, 1) if self.unfold else (2, 0, 1)
        store = w.permute(*perm).contiguous()
        self.weight = nn.Parameter(store.permute(*self._inverse(perm)))
        self.bias = nn.Parameter(b0) if bias else None
        self._shadow = None
        self._shadow_version = None
        # fp32-faithful mode (set by the owning model, ``precision="tf32"``): the GEMMs read fp32 activations and these fp32 weights
        # as tf32 with fp32 accumulation -- no bf16 shadow exists, activations between the layers stay fp32
        self.f32 = False
        self.is_head = False                  # the bias-only label head (set by the owning model): its output is not an activation buffer

    @property
    def act_dtype(self):
        return torch.float32 if self.f32 else torch.bfloat16

    @property
    def x_dtype(self):
        return F.DT_F32 if self.f32 else F.DT_BF16

    @staticmethod
    def _default_bias(w):
        fan_in = w.shape[1] * w.shape[2]
        bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
        return torch.empty(w.shape[0]).uniform_(-bound, bound)

    @staticmethod
    def _inverse(perm):
        inv = [0] * len(perm)
        for i, p in enumerate(perm):
            inv[p] = i
        return tuple(inv)

    # ---- kernel-side geometry
    @property
    def k_eff(self):
        return 1 if self.unfold else self.kernel_size[0]

    @property
    def cin_eff(self):
        return self.in_channels * self.kernel_size[0] if self.unfold else self.in_channels

    @property
    def cout_pad(self):
        """rows per tap of the packed weights: multiple of 16 (UMMA N granularity), at least 64 (dgrad K chunk)"""
        return max(64, (self.out_channels + 15) // 16 * 16)

    # ---- physical channel counts.  Every memory-bound pass moves 16-byte vectors of 8 channels and the GEMMs contract over at least
    # 64, while the reference accepts ANY width (wav2letter.py:59-64; 250 in the original paper; 161 STFT bins when input_size is unset):
    # activation buffers therefore carry max(64, ceil8(C)) columns, the surplus ones exact zeros -- zero weight rows produce them, zero
    # weight columns ignore them, BatchNorm maps 0 to 0 there (mean 0, shift 0) and every activation keeps 0 -- and parameters /
    # gradients keep the reference's logical shapes.  ``padded`` is False for every shipped yaml: nothing changes for those.
    @staticmethod
    def phys(c):
        return max(64, (c + 7) // 8 * 8)

    @property
    def cin_phys(self):
        return self.phys(self.in_channels) * (self.kernel_size[0] if self.unfold else 1)

    @property
    def cout_phys(self):
        return self.out_channels if self.is_head else self.phys(self.out_channels)     # (the label head writes its logical columns)

    @property
    def padded(self):
        return self.cin_phys != self.cin_eff or self.cout_phys != self.out_channels

    def unpad_dw(self, dw):
        """[k_eff, rows, cin_phys] as the weight-gradient kernel wrote it -> contiguous [k_eff, Cout, cin_eff] (the Parameter's layout)"""
        Co, Ci, k = self.out_channels, self.in_channels, self.kernel_size[0]
        if self.unfold:
            return dw.view(dw.shape[1], k, self.phys(Ci))[:Co, :, :Ci].reshape(1, Co, k * Ci)
        return dw[:, :Co, :Ci].contiguous()

    def storage(self):
        """fp32 weights in kernel layout [k_eff, Cout, cin_eff] (a view of the Parameter's memory)."""
        perm = (0, 2, 1) if self.unfold else (2, 0, 1)
        v = self.weight.detach().permute(*perm)
        if not v.is_contiguous():             # e.g. a Parameter re-created by external code: restore the layout
            v = v.contiguous()
            self.weight.data = v.permute(*self._inverse(perm))
        return v.reshape(self.k_eff, self.out_channels, self.cin_eff)

    def grad_view(self, dw_store):
        """[k_eff, Cout, cin_eff] fp32 -> tensor shaped/strided like the Parameter."""
        k = self.kernel_size[0]
        if self.unfold:
            return dw_store.view(self.out_channels, k, self.in_channels).permute(0, 2, 1)
        return dw_store.view(k, self.out_channels, self.in_channels).permute(1, 2, 0)

    def packed(self):
        """bf16 shadow [k_eff, cout_pad, cin_eff] read by the GEMM kernels; refreshed when the Parameter changed."""
        w = self.weight
        ver = (w._version, w.data_ptr())
        if self.f32 and self.cout_pad == self.out_channels and not self.padded:
            return self.storage()              # fp32 mode: the master weights ARE the operand
        dt = self.act_dtype
        if self._shadow is None or self._shadow.device != w.device or self._shadow.dtype != dt:
            self._shadow = torch.zeros((self.k_eff, self.cout_pad, self.cin_phys), dtype=dt, device=w.device)
            self._shadow_version = None
        if self._shadow_version != ver:
            st = self.storage()
            Co, Ci, k = self.out_channels, self.in_channels, self.kernel_size[0]
            if self.cin_phys != self.cin_eff:  # padded input channels: zero columns (per tap for the unfolded first layer)
                if self.unfold:
                    self._shadow[0].view(self.cout_pad, k, self.phys(Ci))[:Co, :, :Ci].copy_(st[0].view(Co, k, Ci))
                else:
                    self._shadow[:, :Co, :Ci].copy_(st)
            elif self.f32:                     # padded fp32 copy (the head: Cout is not a multiple of 16)
                self._shadow[:, :self.out_channels].copy_(st)
            elif self.cout_pad == self.out_channels:
                F.cast_bf16(st, self._shadow)
            else:                              # padded rows stay zero; k_eff == 1 for the only such layer (the head)
                for j in range(self.k_eff):
                    F.cast_bf16(st[j], self._shadow[j, :self.out_channels])
            self._shadow_version = ver
        return self._shadow

    def packed_t(self):
        """bf16 shadow for backward-data: [k_eff, cin_pad16, cout_pad], taps reversed and transposed (K-major operand)."""
        w = self.weight
        ver = (w._version, w.data_ptr())
        if self.unfold and self.cin_phys != self.cin_eff:
            raise NotImplementedError("backward-data through an unfolded layer with padded input channels")     # (never needed: the model input has no gradient)
        if getattr(self, "_shadow_t", None) is None or self._shadow_t.device != w.device or self._shadow_t.dtype != self.act_dtype:
            cin_pad = (self.cin_phys + 15) // 16 * 16
            self._shadow_t = torch.zeros((self.k_eff, cin_pad, self.cout_pad), dtype=self.act_dtype, device=w.device)
            self._shadow_t_version = None
        if self._shadow_t_version != ver:
            F.pack_wt(self.storage(), self._shadow_t, self.out_channels, self.cin_eff)
            self._shadow_t_version = ver
        return self._shadow_t

    def prefetch_packed_t(self):
        """Called from the forward pass: re-pack the backward-data shadow on the side stream, where it runs beside the forward
        GEMMs instead of sitting on the backward critical path."""
        if not WgradStream.enabled or not self.weight.is_cuda:
            return
        if getattr(self, "_shadow_t", None) is not None and self._shadow_t_version == (self.weight._version, self.weight.data_ptr()):
            return
        dev = self.weight.device
        side = WgradStream.side(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self.packed_t()
            self._shadow_t_event = side.record_event()

    def packed_t_synced(self):
        """``packed_t()`` for the compute stream: waits (on the device) for a prefetch in flight."""
        ev = getattr(self, "_shadow_t_event", None)
        if ev is not None:
            torch.cuda.current_stream(self.weight.device).wait_event(ev)
            self._shadow_t_event = None
        return self.packed_t()

    def mark_shadow_fresh(self):
        """Called by the fused optimizer, which rewrites the shadow itself."""
        self._shadow_version = (self.weight._version, self.weight.data_ptr())

    def extra_repr(self):
        return "%d, %d, kernel_size=%s, stride=%s, dilation=%s" % (self.in_channels, self.out_channels, self.kernel_size,
                                                                  self.stride, self.dilation)


class DepthwiseParams(nn.Module):
    """Stands where the reference has ``nn.Conv1d(C, C, k, groups=C, bias=False)`` (the depthwise half of a separable
    Jasper sub-block, jasper.py:318-330): weight ``[C, 1, k]``, stored in memory as fp32 ``[k, C]``."""

    def __init__(self, channels, kernel_size, stride=1, padding=0, dilation=1):
        super().__init__()
        k = kernel_size[0] if isinstance(kernel_size, (tuple, list)) else kernel_size
        self.in_channels = self.out_channels = self.groups = channels
        self.kernel_size, self.stride, self.dilation, self.padding = (k,), (stride,), (dilation,), (padding,)
        w = torch.empty(channels, 1, k)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))                     # nn.Conv1d's default, same RNG consumption
        self.weight = nn.Parameter(w.permute(2, 1, 0).contiguous().permute(2, 1, 0))
        self.bias = None
        self.f32 = False                      # fp32-faithful mode (set by the owning model): fp32 activations in and out

    @property
    def act_dtype(self):
        return torch.float32 if self.f32 else torch.bfloat16

    def storage(self):
        v = self.weight.detach().permute(2, 1, 0)
        if not v.is_contiguous():
            v = v.contiguous()
            self.weight.data = v.permute(2, 1, 0)
        return v.reshape(self.kernel_size[0], self.in_channels)

    def grad_view(self, dw_store):
        return dw_store.view(self.kernel_size[0], 1, self.in_channels).permute(2, 1, 0)

    # channel counts that are not a multiple of 8 (or below 64): the activation buffers carry ConvParams.phys(C) channels whose surplus
    # ones are exact zeros; the depthwise kernels see zero taps there, the gradient keeps the Parameter's logical shape
    @property
    def c_phys(self):
        return ConvParams.phys(self.in_channels)

    def storage_phys(self):
        st = self.storage()
        if self.c_phys == self.in_channels:
            return st
        out = torch.zeros((st.shape[0], self.c_phys), dtype=st.dtype, device=st.device)
        out[:, :self.in_channels].copy_(st)
        return out

    def unpad_dw(self, dw):
        return dw if self.c_phys == self.in_channels else dw[:, :self.in_channels].contiguous()


class DepthwiseFn(torch.autograd.Function):
    """depthwise conv over time-major bf16 (zero 'same' padding); rows >= out_lens are written as zeros because the
    consumer is a MaskedConv1d (jasper.py:116-119), and their gradient is ignored accordingly."""

    @staticmethod
    def forward(ctx, xin, weight, conv, T_out, out_lens):
        k, s, d, p = conv.kernel_size[0], conv.stride[0], conv.dilation[0], conv.padding[0]
        y = F.depthwise_fwd(xin, conv.storage_phys(), T_out, k, s, d, p, out_lens)
        ctx.conv, ctx.out_lens = conv, out_lens
        ctx.save_for_backward(xin)
        return y

    @staticmethod
    def backward(ctx, dy):
        (xin,) = ctx.saved_tensors
        conv = ctx.conv
        k, s, d, p = conv.kernel_size[0], conv.stride[0], conv.dilation[0], conv.padding[0]
        dy = dy.contiguous()
        dw = conv.unpad_dw(F.depthwise_wgrad(dy, xin, k, s, d, p, ctx.out_lens))
        dx = None
        if ctx.needs_input_grad[0]:
            dx = F.depthwise_dgrad(dy, conv.storage_phys(), xin.shape[1], k, d, p, ctx.out_lens, stride=s)
        return dx, conv.grad_view(dw), None, None, None


class UnfoldTmFn(torch.autograd.Function):
    """Unfold in front of a strided layer that is not the first one (Conv1dBlock stride > 1, wav2letter.py:24-38; a strided dense
    JasperBlock, jasper.py:289-298): time-major bf16 [B, rows, C] -> [B, T_out, k*C]; the conv then runs as a k=1 GEMM with weights
    stored [Cout, k, Cin] (``ConvParams(unfold=True)``).  backward folds the column gradient back onto the rows (fp32 sums)."""

    @staticmethod
    def forward(ctx, xin, T_out, k, stride, dilation, pad_left):
        ctx.geo = (xin.shape[1], xin.shape[2], k, stride, dilation, pad_left)
        return F.im2col_tm(xin, T_out, k, stride, dilation, pad_left)

    @staticmethod
    def backward(ctx, dcol):
        rows, C, k, stride, dilation, pad_left = ctx.geo
        return F.col2im_tm(dcol.contiguous(), rows, C, k, stride, dilation, pad_left), None, None, None, None, None


class BatchNormParams(nn.Module):
    """Stands where the reference has ``nn.BatchNorm1d`` (wav2letter.py:37, jasper.py:363)."""

    def __init__(self, num_features, eps=1e-3, momentum=0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.affine, self.track_running_stats = True, True
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))

    def eval_scale_shift(self, conv_bias=None):
        """inference fold: y = z*scale + shift; computed once per parameter / buffer version (five small launches otherwise, per
        layer and forward call: the eval-mode forward of config 1 was host-bound on them)"""
        ts = (self.weight, self.bias, self.running_mean, self.running_var) + ((conv_bias,) if conv_bias is not None else ())
        key = tuple((t._version, t.data_ptr()) for t in ts)
        hit = self.__dict__.get("_w2l_eval_fold")
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        scale = self.weight.detach() * torch.rsqrt(self.running_var + self.eps)
        shift = self.bias.detach() - self.running_mean * scale
        if conv_bias is not None:
            shift = shift + conv_bias.detach() * scale
        scale, shift = scale.contiguous(), shift.contiguous()
        self.__dict__["_w2l_eval_fold"] = (key, scale, shift)
        return scale, shift

    def extra_repr(self):
        return "%d, eps=%g, momentum=%g" % (self.num_features, self.eps, self.momentum)


def conv_desc(conv, B, T_out, x_rows, x_row_offset, y_rows=None, y_row_offset=0, ldy=None, y_dtype=F.DT_BF16, act=F.ACT_NONE, cout=None):
    """``cout``: output columns of the GEMM -- default the layer's PHYSICAL width (the activation buffer's columns; the surplus ones come
    out as exact zeros from the zero rows of the packed weights); the label head passes its logical width."""
    if conv.f32 and y_dtype == F.DT_BF16:
        y_dtype = F.DT_F32                     # fp32-faithful mode: every activation buffer is fp32
    cout = conv.cout_phys if cout is None else cout
    return F.make_desc(B, T_out, conv.cin_phys, cout, conv.cout_pad, conv.k_eff, conv.dilation[0], x_rows, x_row_offset,
                       T_out if y_rows is None else y_rows, y_row_offset, cout if ldy is None else ldy, y_dtype, act,
                       conv.x_dtype)


_EPILOGUE_STATS = os.environ.get("W2L_EPILOGUE_STATS", "1") != "0"      # 0: separate bn_stats pass over z (A/B measurements)
_STATS_PASS_K1 = os.environ.get("W2L_STATS_PASS_K1", "1") != "0"        # 0: k = 1 convs keep their statistics in the GEMM epilogue too


class FusedBnReduce:
    """The first of the two BatchNorm-backward passes of a block (sum g, sum g*xhat over all rows) folded into the epilogue of the
    backward-data GEMM of the layer ABOVE it, which holds exactly those rows of the gradient in registers when it stores them
    (``F.conv1d_dgrad_wt(bnred=...)``).  The producer block tags its output tensor with what the epilogue needs; the consumer (the next
    ``ConvBNActFn`` / ``ConvHeadFn``) claims the tag in its forward pass and, in its backward pass, hands it to its GEMM and marks the
    reduction done, so that the producer runs the apply pass only.  Not fused -- the separate reduce pass runs as before -- when the
    output has a second consumer (a Jasper block output also feeds the next block's residual conv: two gradients are summed), for
    blocks with a residual branch, in the fp32-faithful mode.
    OPT-IN (W2L_FUSE_BN_REDUCE=1): measured on B200 (profiles/r2_bn_reduce_fusion.md) it removes 1.2 ms of serialized kernel time per
    W2L-20 step but the step gets 0.3 ms SLOWER (37.4 -> 37.7; Jasper 10x5 81.4 -> 82.5): the separate pass was already hidden beside
    the weight-gradient GEMM on the side stream, while the heavier epilogue (row-strided z loads, 124 more shuffles per 32 columns)
    lengthens the backward-data GEMMs that sit on the critical chain."""
    enabled = os.environ.get("W2L_FUSE_BN_REDUCE", "0") == "1"
    fused_launches = 0           # diagnostic counter (tests)

    @staticmethod
    def tag(yp, **info):
        info.update(consumers=0, red=None)
        yp._w2l_producer = info
        return info

    @staticmethod
    def claim(xin):
        info = getattr(xin, "_w2l_producer", None)
        if info is not None:
            info["consumers"] += 1
        return info

    @classmethod
    def for_dgrad(cls, info, xin):
        """the ``bnred`` argument for the consumer's backward-data GEMM, or None"""
        if not cls.enabled or info is None or info["consumers"] != 1 or info["red"] is not None:
            return None
        B, T, (pl, pr) = info["B"], info["T"], info["pad"]
        if xin.dtype != torch.bfloat16 or tuple(xin.shape) != (B, pl + T + pr, info["C"]) or (info["drop_p"] > 0 and info["mask"] is None):
            return None
        fin = info["fin"]
        red = info["scratch"].take_red()
        info["red"] = red
        cls.fused_launches += 1
        return dict(z=info["z"], mask=info["mask"], scale=fin[0], shift=fin[1], mean=fin[2], lens=info["lens"], red=red, B=B, T=T,
                    pad_left=pl, pad_right=pr, act=info["act"], drop_p=info["drop_p"])


class BnScratch:
    """Two small persistent fp32 buffers of a BatchNorm layer that kernels accumulate into with atomics -- ``stats`` [2C] (the conv
    epilogue's batch sums, forward) and ``red`` [2C] (sum g, sum g*xhat, backward) -- each cleared by a kernel of the OTHER pass that
    runs anyway (the forward BN/activation pass clears ``red``, the backward apply pass clears ``stats``), so a steady-state training
    step needs no memset launch for them (round 1 launched one ``torch.zeros`` per buffer, layer and step: 60 of the ~125 fill
    kernels per step).  The flags track which buffer is known to be zero; a pass that finds its buffer dirty (a forward that was never
    followed by its backward, two backward passes over one graph) clears it explicitly."""

    def __init__(self, C, device):
        self.stats = torch.zeros((2 * C,), dtype=torch.float32, device=device)
        self.red = torch.zeros((2 * C,), dtype=torch.float32, device=device)
        self.stats_clean = self.red_clean = True

    def take_stats(self):
        if not self.stats_clean:
            self.stats.zero_()
        self.stats_clean = False
        return self.stats

    def take_red(self):
        if not self.red_clean:
            self.red.zero_()
        self.red_clean = False
        return self.red


def bn_scratch(bn, device, C=None):
    C = bn.num_features if C is None else C               # (the physical width when the channel count is padded)
    sc = bn.__dict__.get("_w2l_scratch")
    if sc is None or sc.stats.device != device or sc.stats.numel() != 2 * C:
        sc = bn.__dict__["_w2l_scratch"] = BnScratch(C, device)
    return sc


def pad_channels(t, Cp, value=0.0):
    """per-channel vector [C] -> [Cp] (a copy with ``value`` in the surplus channels); None and already-wide vectors pass through"""
    if t is None or t.numel() == Cp:
        return t
    out = torch.full((Cp,), value, dtype=t.dtype, device=t.device)
    out[:t.numel()].copy_(t.detach())
    return out


def conv_fwd_with_stats(xin, conv, desc, z, stats=None):
    """conv forward into ``z`` + BatchNorm batch statistics [2*Cout] (sum, sum of squares of the stored bf16 values), accumulated into
    ``stats`` (zero on entry) when given."""
    if stats is None:
        stats = torch.zeros((2 * conv.cout_phys,), dtype=torch.float32, device=xin.device)
    # k = 1 (Jasper's residual 1x1 convs, the 1024-wide tail layer, the unfolded first layer): the mainloop of a tile is too short to
    # hide the statistics block of the epilogue -- the forward GEMM ran at half the speed of the backward-data GEMM of the same shape
    # (profiles/r2_gemm_cg2.md, addendum) -- while the stored output is still L2-resident for a separate pass
    if not _EPILOGUE_STATS or (_STATS_PASS_K1 and conv.k_eff == 1 and not conv.f32):
        F.conv1d_fwd(xin, conv.packed(), desc, z)
        return F.bn_stats(z, conv.cout_phys, out=stats)
    F.conv1d_fwd(xin, conv.packed(), desc, z, bn_stats=stats)
    return stats


def zero_bias_grad(conv, device):
    """the conv bias gradient under training-mode BatchNorm is exactly zero: one cached zero tensor per layer instead of a fill
    kernel per layer and step (a fresh one whenever a gradient is already being accumulated into)"""
    if conv.bias.grad is not None:
        return torch.zeros(conv.out_channels, dtype=torch.float32, device=device)
    z = conv.__dict__.get("_w2l_zero_dbias")
    if z is None or z.device != device:
        z = conv.__dict__["_w2l_zero_dbias"] = torch.zeros(conv.out_channels, dtype=torch.float32, device=device)
    return z


class ConvBNActFn(torch.autograd.Function):
    """conv (+bias) -> BatchNorm(train statistics) [+ residual branch] -> dropout -> activation, written into the
    consumer's padded buffer.  Inputs: time-major bf16 ``xin`` ([B, x_rows, cin_eff]), the conv/BN parameters, the
    optional residual pair (z_res, fin_res) produced by ``ResidualBranchFn``, and a geometry dict:
      T_out, x_row_offset, out_pad=(pl, pr) reflect halo wanted by the consumer, act, drop_p, lens (int32 [B] | None).
    Returns yp [B, pl+T_out+pr, Cout] bf16.

    Gradient convention for the residual pair: the value returned for ``z_res`` is the (masked) gradient with
    respect to the residual branch's BatchNorm OUTPUT, which is what ``ResidualBranchFn.backward`` expects."""

    @staticmethod
    def forward(ctx, xin, weight, bias, gamma, beta, z_res, fin_res, conv, bn, geo):
        B, x_rows, _ = xin.shape
        T_out = geo["T_out"]
        Co, C_log = conv.cout_phys, conv.out_channels      # physical (buffer) / logical (parameter) width: equal unless padded
        pl, pr = geo.get("out_pad", (0, 0))
        z = torch.empty((B, T_out, Co), dtype=conv.act_dtype, device=xin.device)
        desc = conv_desc(conv, B, T_out, x_rows, geo["x_row_offset"])
        if ctx.needs_input_grad[0]:
            conv.prefetch_packed_t()                       # side stream: runs beside the GEMM launched next
        sc = bn_scratch(bn, xin.device, Co)
        rm, rv = bn.running_mean, bn.running_var
        if Co != C_log:                                    # padded width: per-channel vectors with neutral surplus entries (see ConvParams.phys)
            gamma, beta, bias = pad_channels(gamma, Co, 1.0), pad_channels(beta, Co), pad_channels(bias, Co)
            rm, rv = pad_channels(rm, Co), pad_channels(rv, Co, 1.0)
        stats = conv_fwd_with_stats(xin, conv, desc, z, sc.take_stats())      # batch statistics from the GEMM epilogue
        drop_p = geo.get("drop_p", 0.0)
        seed = next_dropout_seed() if drop_p > 0 else 0
        mask = torch.empty((B * T_out * Co // 8,), dtype=torch.uint8, device=xin.device) if drop_p > 0 else None
        has_res = z_res is not None
        # ONE launch: statistics -> scale/shift (+ running statistics, conv bias folded into the running mean), BatchNorm apply,
        # residual, dropout, activation, the consumer's halo / mask; it also clears this layer's backward reduction buffer
        yp, fin = F.bn_finalize_act_pad(z, stats, gamma, beta, bias, bn.eps, bn.momentum, rm, rv,
                                        bn.num_batches_tracked, B, T_out, Co, pl, pr, geo["act"], drop_p, seed, geo.get("lens"),
                                        res=z_res, res_scale=fin_res[0] if has_res else None, res_shift=fin_res[1] if has_res else None,
                                        drop_mask=mask, zero_after=sc.red)
        sc.red_clean = True
        if Co != C_log:
            bn.running_mean.copy_(rm[:C_log])
            bn.running_var.copy_(rv[:C_log])
        ctx.bn = bn
        ctx.conv, ctx.geo, ctx.seed, ctx.desc, ctx.has_res = conv, geo, seed, desc, has_res
        ctx.has_bias = bias is not None
        ctx.prod_in = FusedBnReduce.claim(xin) if ctx.needs_input_grad[0] else None
        ctx.prod_out = None
        if not has_res and not conv.f32:
            ctx.prod_out = FusedBnReduce.tag(yp, z=z, fin=fin, mask=mask, B=B, T=T_out, C=Co, pad=(pl, pr), act=geo["act"], drop_p=drop_p,
                                             lens=geo.get("lens"), scratch=sc)
        ctx.save_for_backward(xin, z, fin, gamma, z_res, fin_res, mask)
        return yp

    @staticmethod
    def backward(ctx, dyp):
        xin, z, fin, gamma, z_res, fin_res, mask = ctx.saved_tensors
        conv, geo, has_res = ctx.conv, ctx.geo, ctx.has_res
        B, T_out, Co = z.shape
        pl, pr = geo.get("out_pad", (0, 0))
        x_rows = xin.shape[1]
        halo = (conv.k_eff - 1) * conv.dilation[0]
        # Inputs that carry their own halo (x_rows == T_out + (k-1)d, Wav2Letter): dz is stored with the INPUT's row pitch and
        # zero tails, so that backward-data runs over one flat [B*x_rows] row space (no per-utterance tile padding).
        flat = ctx.needs_input_grad[0] and geo["x_row_offset"] == 0 and x_rows == T_out + halo
        dz_rows = x_rows if flat else T_out
        sc = bn_scratch(ctx.bn, z.device, Co)
        pre_red = ctx.prod_out["red"] if ctx.prod_out is not None else None      # the consumer's GEMM epilogue already reduced (FusedBnReduce)
        dz, red, g = F.bn_act_bwd(dyp.contiguous(), z, fin[0], fin[1], fin[2], fin[3], gamma, B, T_out, Co, pl, pr, geo["act"],
                                  geo.get("drop_p", 0.0), ctx.seed, geo.get("lens"), res=z_res,
                                  res_scale=fin_res[0] if has_res else None, res_shift=fin_res[1] if has_res else None,
                                  want_g=has_res, dz_rows=dz_rows, drop_mask=mask, red_ws=None if pre_red is not None else sc.take_red(),
                                  zero_after=sc.stats, red_raw=pre_red)
        sc.stats_clean = True                            # the apply pass cleared the forward statistics for the next step
        dw = alloc_dw(conv, z.device, cout=Co)
        side = WgradStream.fork(z.device, conv.weight)   # dz is enqueued: wgrad may start; dgrad goes first on the compute stream
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(xin)
            bnred = FusedBnReduce.for_dgrad(ctx.prod_in, xin)         # fold the reduction of the block that produced xin into this GEMM
            if flat:
                F.conv1d_dgrad_wt(dz, conv.packed_t_synced(), conv_desc(conv, 1, B * x_rows, B * x_rows, 0), dx, bnred=bnred)
            else:
                F.conv1d_dgrad_wt(dz, conv.packed_t_synced(), ctx.desc, dx, bnred=bnred)
        dw = wgrad_async(side, dz, xin, conv_desc(conv, B, T_out, x_rows, geo["x_row_offset"], y_rows=dz_rows), dw, conv)
        dbias = zero_bias_grad(conv, z.device) if ctx.has_bias else None                          # exactly 0 under train BN
        C_log = conv.out_channels                                                                 # (Co is the physical width: padded layers)
        return dx, conv.grad_view(dw), dbias, red[Co:Co + C_log], red[:C_log], g, None, None, None, None


class ResidualBranchFn(torch.autograd.Function):
    """Jasper residual branch (jasper.py:400-412): 1x1 conv of the block input, BatchNorm statistics.  Returns
    (z_res bf16 [B,T,C], fin_res fp32 [4,C] = scale, shift, mean, invstd); the BatchNorm is APPLIED inside the main
    branch's fused bn_act_pad pass.  backward() receives, for z_res, the gradient w.r.t. the branch's BN output."""

    @staticmethod
    def forward(ctx, xin, weight, gamma, beta, conv, bn):
        B, T, _ = xin.shape
        Co, C_log = conv.cout_phys, conv.out_channels      # physical (buffer) / logical (parameter) width: equal unless padded
        z = torch.empty((B, T, Co), dtype=conv.act_dtype, device=xin.device)
        desc = conv_desc(conv, B, T, T, 0)
        if ctx.needs_input_grad[0]:
            conv.prefetch_packed_t()
        rm, rv = bn.running_mean, bn.running_var
        if Co != C_log:                                    # neutral surplus entries, as in ConvBNActFn
            gamma, beta = pad_channels(gamma, Co, 1.0), pad_channels(beta, Co)
            rm, rv = pad_channels(rm, Co), pad_channels(rv, Co, 1.0)
        stats = conv_fwd_with_stats(xin, conv, desc, z)
        fin = F.bn_finalize(stats, B * T, Co, gamma, beta, None, bn.eps, bn.momentum, rm, rv, bn.num_batches_tracked)
        if Co != C_log:
            bn.running_mean.copy_(rm[:C_log])
            bn.running_var.copy_(rv[:C_log])
        FusedBnReduce.claim(xin)            # a second consumer of the block input: its producer's reduction cannot be folded into one GEMM
        ctx.conv, ctx.desc = conv, desc
        ctx.save_for_backward(xin, z, fin, gamma)
        ctx.mark_non_differentiable(fin)
        return z, fin

    @staticmethod
    def backward(ctx, g, _unused):
        xin, z, fin, gamma = ctx.saved_tensors
        conv = ctx.conv
        B, T, Co = z.shape
        dz, red, _ = F.bn_act_bwd(g.contiguous(), z, fin[0], fin[1], fin[2], fin[3], gamma, B, T, Co, 0, 0, F.ACT_NONE)
        dw = alloc_dw(conv, z.device, cout=Co)
        side = WgradStream.fork(z.device, conv.weight)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(xin)
            F.conv1d_dgrad_wt(dz, conv.packed_t_synced(), ctx.desc, dx)
        dw = wgrad_async(side, dz, xin, ctx.desc, dw, conv)
        C_log = conv.out_channels
        return dx, conv.grad_view(dw), red[Co:Co + C_log], red[:C_log], None, None


class ConvHeadFn(torch.autograd.Function):
    """k=1 conv with bias to the label logits (fp32), then log_softmax (mode 0) / softmax (mode 1) (wav2letter.py:66,86-87;
    jasper.py:433,468-473), or the raw logits (mode 2: the head block called on its own).  Returns [B, T, n_labels] fp32
    contiguous."""

    @staticmethod
    def forward(ctx, xin, weight, bias, conv, mode, nan_flag=None):
        B, T, _ = xin.shape
        Co = conv.out_channels
        ld = (Co + 7) // 8 * 8
        logits = torch.empty((B, T, ld), dtype=torch.float32, device=xin.device)
        desc = conv_desc(conv, B, T, T, 0, ldy=ld, y_dtype=F.DT_F32, cout=Co)
        if ctx.needs_input_grad[0]:
            conv.prefetch_packed_t()
        F.conv1d_fwd(xin, conv.packed(), desc, logits, bias=bias)
        out = logits[..., :Co].contiguous() if mode == 2 else F.log_softmax(logits, Co, mode, nan_flag)     # 2: raw logits
        ctx.prod_in = FusedBnReduce.claim(xin) if ctx.needs_input_grad[0] else None
        ctx.conv, ctx.mode = conv, mode
        ctx.has_bias = bias is not None
        ctx.save_for_backward(xin, out)
        return out

    @staticmethod
    def backward(ctx, dout):
        xin, out = ctx.saved_tensors
        conv = ctx.conv
        if ctx.mode == 1:
            raise NotImplementedError("backward through the eval-mode softmax head is not supported")
        B, T, _ = xin.shape
        Co, cp = conv.out_channels, conv.cout_pad
        dl = F.log_softmax_bwd(dout.contiguous(), out, cp, fused_identity=ctx.mode == 2, out_dtype=conv.act_dtype)    # [B,T,cout_pad], zero padded
        desc = conv_desc(conv, B, T, T, 0, ldy=cp, cout=Co)
        dw = alloc_dw(conv, xin.device, cout=Co)
        side = WgradStream.fork(xin.device, conv.weight)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(xin)
            F.conv1d_dgrad_wt(dl, conv.packed_t_synced(), desc, dx, bnred=FusedBnReduce.for_dgrad(ctx.prod_in, xin))
        dw = wgrad_async(side, dl, xin, desc, dw, conv)
        dbias = F.colsum(dl, Co) if ctx.has_bias else None
        return dx, conv.grad_view(dw), dbias, None, None, None


def conv_bn_act_eval(xin, conv, bn, geo, res=None):
    """Inference path: BatchNorm folded into the conv epilogue (scale/shift/activation fused), then the reflect-halo
    fill; with a residual branch or a length mask the BN/act/mask pass runs as in training (running statistics).
    ``res`` = (z_res, (scale_res, shift_res))."""
    B, x_rows, _ = xin.shape
    T_out, Co = geo["T_out"], conv.cout_phys
    pl, pr = geo.get("out_pad", (0, 0))
    lens = geo.get("lens")
    scale, shift = bn.eval_scale_shift(conv.bias)
    if Co != conv.out_channels:                    # padded width: the surplus channels come out as 0 * 0 + 0
        scale, shift = pad_channels(scale, Co), pad_channels(shift, Co)
    if res is None and lens is None:
        y = torch.empty((B, pl + T_out + pr, Co), dtype=conv.act_dtype, device=xin.device)
        desc = conv_desc(conv, B, T_out, x_rows, geo["x_row_offset"], y_rows=pl + T_out + pr, y_row_offset=pl, act=geo["act"])
        F.conv1d_fwd(xin, conv.packed(), desc, y, scale=scale, shift=shift)
        return F.reflect_halo(y, T_out, pl, pr)
    z = torch.empty((B, T_out, Co), dtype=conv.act_dtype, device=xin.device)
    desc = conv_desc(conv, B, T_out, x_rows, geo["x_row_offset"])
    F.conv1d_fwd(xin, conv.packed(), desc, z)
    return F.bn_act_pad(z, scale, shift, B, T_out, Co, pl, pr, geo["act"], 0.0, 0, lens, res=None if res is None else res[0],
                        res_scale=None if res is None else res[1][0], res_shift=None if res is None else res[1][1])
