"""Thin tensor-level wrappers over the C ABI (include/w2l_sm100.h).

PyTorch is used here for device memory, streams and autograd plumbing only; every computation is done by the
hand-written sm_100a kernels in libw2l_sm100.so.  CPU tensors are rejected -- there is no fallback.
"""
import ctypes

import torch

from . import _lib
from ._lib import BnReduce, ConvDesc

ACT_NONE, ACT_RELU, ACT_CLAMP20 = 0, 1, 2
PAD_ZERO, PAD_REFLECT = 0, 1
DT_BF16, DT_F32 = 0, 1
RED_RAW = 0x200            # W2L_RED_RAW: the apply pass takes raw sums of g*(z-mean) from the GEMM epilogue
STORE_F32 = 0x100          # W2L_STORE_F32: OR-ed into `act` when the activation buffers of a BatchNorm / activation pass are fp32


def _act_flag(act, t):
    """`act` with W2L_STORE_F32 when tensor ``t`` (the pass's conv output) is fp32: the fp32-faithful mode"""
    return act | STORE_F32 if t.dtype == torch.float32 else act


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


# The two per-call host costs of this module were torch.cuda.current_stream() and the torch.cuda.device() context (~15 us each, tens
# of calls per training step: the literal default yaml -- one hidden layer -- was host-bound on them).  Both have a direct route: the
# raw stream handle of the current device, and no device switch at all when the tensor already lives on the current device.
_raw = {"ok": hasattr(torch._C, "_cuda_getCurrentRawStream") and hasattr(torch._C, "_cuda_getDevice")}


def _stream():
    if _raw["ok"]:
        try:
            return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))
        except Exception:  # noqa: BLE001  (no CUDA runtime in this process, e.g. the host emulation: use the public API from now on)
            _raw["ok"] = False
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def capturing():
    """True while the current stream is being captured into a CUDA graph (graph_step.py).  False in a process without a CUDA runtime."""
    try:
        return torch.cuda.is_current_stream_capturing()
    except Exception:  # noqa: BLE001
        return False


class _NoSwitch:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_SWITCH = _NoSwitch()


def _on(device):
    """context that makes ``device`` the current CUDA device -- a no-op object when it already is"""
    if _raw["ok"]:
        try:
            if device.index is None or device.index == torch._C._cuda_getDevice():
                return _NO_SWITCH
        except Exception:  # noqa: BLE001
            _raw["ok"] = False
    return torch.cuda.device(device)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("wav2letter_pytorch_b200: CUDA tensors required (got a %s tensor); there is no CPU path"
                               % t.device.type)


def to_cuda(t, who="wav2letter_pytorch_b200"):
    """host tensor -> current CUDA device; there is no CPU fallback, so without a device this raises"""
    if not torch.cuda.is_available():
        raise RuntimeError("%s: a CUDA device is required (no CPU fallback)" % who)
    return t.cuda()


# --------------------------------------------------------------------------------------------- decode
def greedy_decode(scores, sizes=None, blank=0):
    """scores [N,T,C] fp32 cuda -> (argmax [N,T], tokens [N,T], offsets [N,T], counts [N]) int32 cuda.
    decoder.py:121-145 (argmax + collapse)."""
    _need_cuda(scores, sizes)
    lib = _lib.load()
    if scores.dtype != torch.float32:
        scores = scores.float()
    if scores.stride(2) != 1:
        scores = scores.contiguous()
    N, T, C = scores.shape
    dev = scores.device
    argmax = torch.empty((N, T), dtype=torch.int32, device=dev)
    tokens = torch.empty((N, T), dtype=torch.int32, device=dev)
    offsets = torch.empty((N, T), dtype=torch.int32, device=dev)
    counts = torch.empty((N,), dtype=torch.int32, device=dev)
    if sizes is not None:
        sizes = sizes.to(device=dev, dtype=torch.int32).contiguous()
    wsb = lib.w2l_greedy_decode_workspace_bytes(N, T)
    ws = torch.empty((max(wsb, 4),), dtype=torch.uint8, device=dev)
    with _on(dev):
        _lib.check(lib.w2l_greedy_decode(_ptr(scores), N, T, C, scores.stride(0), scores.stride(1), _ptr(sizes), blank, _ptr(argmax),
                                         _ptr(tokens), _ptr(offsets), _ptr(counts), _ptr(ws), ws.numel(), _stream()),
                   "greedy_decode")
    return argmax, tokens, offsets, counts


# --------------------------------------------------------------------------------------------- CTC
def ctc_loss_raw(x, targets, input_lengths, target_lengths, blank=0, zero_infinity=True, reduction_mean=True,
                 from_logits=False, need_grad=True):
    """x [N,T,C] fp32 cuda (any N/T strides).  Returns (loss[1], nll[N], grad[N,T,C] | None).
    grad already carries the 1/(N*max(S,1)) factor of reduction='mean'."""
    _need_cuda(x, targets)
    lib = _lib.load()
    if x.dtype != torch.float32:
        x = x.float()
    if x.stride(2) != 1:
        x = x.contiguous()
    N, T, C = x.shape
    dev = x.device
    targets = targets.to(device=dev, dtype=torch.int32)
    if targets.dim() != 2:
        raise ValueError("targets must be the 2-D zero padded [N, S_max] tensor the collator emits")
    targets = targets.contiguous()
    il = input_lengths.to(device=dev, dtype=torch.int32).contiguous()
    tl = target_lengths.to(device=dev, dtype=torch.int32).contiguous()
    S = targets.shape[1]
    wsb = lib.w2l_ctc_loss_workspace_bytes(N, T, S)
    if wsb == 0:
        raise RuntimeError("ctc_loss: target length %d not supported" % S)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
    nll = torch.empty((N,), dtype=torch.float32, device=dev)
    loss = torch.empty((1,), dtype=torch.float32, device=dev)
    grad = torch.empty((N, T, C), dtype=torch.float32, device=dev) if need_grad else None
    with _on(dev):
        _lib.check(lib.w2l_ctc_loss(_ptr(x), int(from_logits), N, T, C, x.stride(0), x.stride(1), _ptr(targets), S, _ptr(il), _ptr(tl),
                                    blank, int(zero_infinity), int(reduction_mean), _ptr(nll), _ptr(grad), _ptr(loss), _ptr(ws),
                                    ws.numel(), _stream()), "ctc_loss")
    return loss, nll, grad


# --------------------------------------------------------------------------------------------- conv
def make_desc(B, T_out, Cin, Cout, Cout_pad, k, dilation, x_rows, x_row_offset, y_rows, y_row_offset, ldy, y_dtype=DT_BF16,
              act=ACT_NONE, x_dtype=DT_BF16):
    return ConvDesc(B, T_out, Cin, Cout, Cout_pad, k, dilation, x_rows, x_row_offset, y_rows, y_row_offset, ldy, y_dtype, act, x_dtype)


_gemm_scratch = {}


def ensure_gemm_scratch(device):
    """Registers (once per process) the zero-filled fp32 scratch the forward GEMM uses to split its last wave of tiles along K."""
    if _gemm_scratch:
        if device not in _gemm_scratch:                  # the library keeps per-process state for ONE device (scratch, kernel attributes)
            raise RuntimeError("wav2letter_pytorch_b200 runs one process per GPU: this process already uses %s, got a tensor on %s"
                               % (next(iter(_gemm_scratch)), device))
        return
    nbytes = 4096 + 148 * 128 * 256 * 4
    buf = torch.zeros((nbytes,), dtype=torch.uint8, device=device)
    _lib.check(_lib.load().w2l_set_gemm_scratch(_ptr(buf), nbytes), "set_gemm_scratch")
    _gemm_scratch[device] = buf


def set_dropout_epoch(epoch):
    """Registers (None: unregisters) a device-resident int64 step counter that every dropout launch folds into its seed
    (w2l_set_dropout_epoch): what lets a captured step draw a fresh mask on every replay.  The caller keeps the tensor alive."""
    if epoch is not None:
        _need_cuda(epoch)
        if epoch.dtype != torch.int64 or epoch.numel() != 1:
            raise ValueError("set_dropout_epoch: one int64 CUDA element expected")
    _lib.check(_lib.load().w2l_set_dropout_epoch(_ptr(epoch)), "set_dropout_epoch")


def conv1d_fwd(x, w, desc, y, bias=None, scale=None, shift=None, bn_stats=None):
    """``bn_stats`` (fp32 [2*Cout], zero-filled): receives the per-channel sum / sum of squares of the stored output."""
    _need_cuda(x, w, y)
    ensure_gemm_scratch(x.device)
    with _on(x.device):
        _lib.check(_lib.load().w2l_conv1d_fwd(_ptr(x), _ptr(w), _ptr(bias), _ptr(scale), _ptr(shift), _ptr(bn_stats), _ptr(y),
                                              ctypes.byref(desc), _stream()), "conv1d_fwd")
    return y


def conv1d_dgrad(dy, w, desc, dx):
    _need_cuda(dy, w, dx)
    with _on(dy.device):
        _lib.check(_lib.load().w2l_conv1d_dgrad(_ptr(dy), _ptr(w), _ptr(dx), ctypes.byref(desc), _stream()), "conv1d_dgrad")
    return dx


def conv1d_dgrad_wt(dy, wt, desc, dx, bnred=None):
    """backward-data with the transposed (K-major) weight shadow; see w2l_conv1d_dgrad_wt.  ``bnred`` (dict: z, mask, scale, shift,
    mean, lens, red, B, T, pad_left, pad_right, act, drop_p) folds the BatchNorm-backward reduction of the block that PRODUCED this
    layer's input into the epilogue (w2l_conv1d_dgrad_wt_bnred): ``red`` [2C] fp32, zero on entry, receives sum g and the raw
    sum g*(z-mean)."""
    _need_cuda(dy, wt, dx)
    with _on(dy.device):
        if bnred is None:
            _lib.check(_lib.load().w2l_conv1d_dgrad_wt(_ptr(dy), _ptr(wt), _ptr(dx), ctypes.byref(desc), _stream()), "conv1d_dgrad_wt")
        else:
            p = lambda t: None if t is None else t.data_ptr()       # noqa: E731
            r = BnReduce(p(bnred["z"]), p(bnred.get("mask")), p(bnred["scale"]), p(bnred["shift"]), p(bnred["mean"]), p(bnred.get("lens")),
                         p(bnred["red"]), bnred["B"], bnred["T"], bnred["pad_left"], bnred["pad_right"], bnred["act"], float(bnred.get("drop_p", 0.0)))
            _lib.check(_lib.load().w2l_conv1d_dgrad_wt_bnred(_ptr(dy), _ptr(wt), _ptr(dx), ctypes.byref(desc), ctypes.byref(r), _stream()),
                       "conv1d_dgrad_wt_bnred")
    return dx


def pack_wt(w_store, wt, cout, cin):
    """fp32 [k, Cout, Cin] -> bf16 wt [k, Cin_pad, Cout_pad], tap-reversed + transposed"""
    k = w_store.shape[0]
    fn = _lib.load().w2l_pack_wt_f32 if wt.dtype == torch.float32 else _lib.load().w2l_pack_wt      # fp32 shadow: the fp32-faithful mode
    with _on(w_store.device):
        _lib.check(fn(_ptr(w_store), _ptr(wt), k, cout, cin, wt.shape[2], wt.shape[1], _stream()), "pack_wt")
    return wt


def conv1d_wgrad(dy, x, desc, dw):
    """dw [k, Cout, Cin] fp32; zero-filled here when the kernel will run split-K (atomic accumulation)."""
    _need_cuda(dy, x, dw)
    lib = _lib.load()
    with _on(dy.device):
        if lib.w2l_conv1d_wgrad_splits(ctypes.byref(desc)) > 1:
            dw.zero_()
        _lib.check(lib.w2l_conv1d_wgrad(_ptr(dy), _ptr(x), _ptr(dw), ctypes.byref(desc), _stream()), "conv1d_wgrad")
    return dw


def tm_to_ct_f32(x, T, x_row_offset=0, pitch=None, C=None, lead=0):
    """time-major fp32 rows [x_row_offset, x_row_offset + T), first C columns, of x [B, rows, ld] -> channel-major [B, C, pitch] fp32
    with ``lead`` zeros in front of every row (out[b, c, lead + t] = x[b, x_row_offset + t, c]; pitch defaults to lead + T rounded up
    to 4 floats; the pad behind is never read): the transposed operands of the fp32-faithful weight gradient"""
    _need_cuda(x)
    B, rows, ld = x.shape
    C = ld if C is None else C
    pitch = (lead + T + 3) // 4 * 4 if pitch is None else pitch
    out = torch.empty((B, C, pitch), dtype=torch.float32, device=x.device)
    if lead:
        out[:, :, :lead].zero_()
    with _on(x.device):
        _lib.check(_lib.load().w2l_tm_to_ct_f32(_ptr(x), ctypes.c_void_p(out.data_ptr() + 4 * lead), B, T, C, rows, x_row_offset, ld, pitch,
                                                _stream()), "tm_to_ct_f32")
    return out


def conv1d_wgrad_t(dy, x, desc, dw):
    """fp32-faithful weight gradient (tf32 multiply, fp32 accumulate) of time-major fp32 dy [B, y_rows, Cout] (rows [0, T_out) used)
    and x [B, x_rows, Cin]: both are transposed to time-contiguous copies first (K-major operands, like the forward GEMM) -- x once
    per residue mod 4 of the taps' row offsets, DELAYED by that many rows (s zeros in front), because TMA wants the time coordinate
    of a load 16-byte aligned.  dw [k, Cout, Cin] fp32 is zero-filled here (tiles shared between CTAs are accumulated with atomics)."""
    _need_cuda(dy, x, dw)
    B, x_rows, _ = x.shape
    dyT = tm_to_ct_f32(dy, desc.T_out, desc.y_row_offset, C=desc.Cout)
    pitch = (x_rows + 3 + 3) // 4 * 4
    need = sorted({(-(desc.x_row_offset + j * desc.dilation)) & 3 for j in range(desc.k)})
    xs = {s: tm_to_ct_f32(x, x_rows, 0, pitch, lead=s) for s in need}
    arr = (ctypes.c_void_p * 4)(*[xs[s].data_ptr() if s in xs else None for s in range(4)])
    with _on(dy.device):
        dw.zero_()
        _lib.check(_lib.load().w2l_conv1d_wgrad_t(_ptr(dyT), dyT.shape[2], arr, pitch, _ptr(dw), ctypes.byref(desc), _stream()),
                   "conv1d_wgrad_t")
    return dw


# --------------------------------------------------------------------------------------------- depthwise
def _dw(name, t):
    """the depthwise entry point for the storage type of tensor ``t`` (bf16, or fp32 in the fp32-faithful mode)"""
    return getattr(_lib.load(), name + ("_f32" if t.dtype == torch.float32 else ""))


def depthwise_fwd(x, w, T_out, k, stride, dilation, pad, out_lens=None):
    """x [B,T,C] bf16 (fp32), w fp32 [k,C] -> y [B,T_out,C] of x's type"""
    _need_cuda(x, w)
    B, T, C = x.shape
    y = torch.empty((B, T_out, C), dtype=x.dtype, device=x.device)
    with _on(x.device):
        _lib.check(_dw("w2l_depthwise_fwd", x)(_ptr(x), _ptr(w), _ptr(y), B, T, C, T_out, k, stride, dilation, pad, _ptr(out_lens), _stream()),
                   "depthwise_fwd")
    return y


def depthwise_dgrad(dy, w, T, k, dilation, pad, dy_lens=None, stride=1):
    B, T_out, C = dy.shape
    dx = torch.empty((B, T, C), dtype=dy.dtype, device=dy.device)
    with _on(dy.device):
        if stride == 1:
            _lib.check(_dw("w2l_depthwise_dgrad", dy)(_ptr(dy), _ptr(w), _ptr(dx), B, T, C, T_out, k, dilation, pad, _ptr(dy_lens),
                                                      _stream()), "depthwise_dgrad")
        else:
            _lib.check(_dw("w2l_depthwise_dgrad_strided", dy)(_ptr(dy), _ptr(w), _ptr(dx), B, T, C, T_out, k, stride, dilation, pad,
                                                              _ptr(dy_lens), _stream()), "depthwise_dgrad_strided")
    return dx


def depthwise_wgrad(dy, x, k, stride, dilation, pad, dy_lens=None):
    B, T_out, C = dy.shape
    T = x.shape[1]
    if dy.dtype != x.dtype:
        raise RuntimeError("depthwise_wgrad: dy %s and x %s must have the same storage type" % (dy.dtype, x.dtype))
    dw = torch.zeros((k, C), dtype=torch.float32, device=dy.device)
    with _on(dy.device):
        _lib.check(_dw("w2l_depthwise_wgrad", dy)(_ptr(dy), _ptr(x), _ptr(dw), B, T, C, T_out, k, stride, dilation, pad, _ptr(dy_lens),
                                                  _stream()), "depthwise_wgrad")
    return dw


# --------------------------------------------------------------------------------------------- elementwise
def im2col_ncw(x, rows, k, stride, dilation, pad_left, pad_mode, lens=None, out_dtype=torch.bfloat16):
    """x [B,F,T] fp32 -> [B, rows, k*F] bf16 (fp32 with ``out_dtype=torch.float32``): unfold + pad + transpose + cast."""
    _need_cuda(x, lens)
    x = x.contiguous().float()
    B, F, T = x.shape
    out = torch.empty((B, rows, k * F), dtype=out_dtype, device=x.device)
    fn = _lib.load().w2l_im2col_ncw_f32 if out_dtype == torch.float32 else _lib.load().w2l_im2col_ncw
    with _on(x.device):
        _lib.check(fn(_ptr(x), _ptr(out), B, F, T, rows, k, stride, dilation, pad_left, pad_mode, _ptr(lens), _stream()), "im2col_ncw")
    return out


def im2col_tm(x, T_out, k, stride, dilation, pad_left):
    """time-major bf16 x [B, x_rows, C] -> [B, T_out, k*C] with out[b, t, j*C + c] = x[b, t*stride + j*dilation - pad_left, c]
    (zero outside the buffer): the unfold in front of a strided layer that is not the first one."""
    _need_cuda(x)
    if x.dtype != torch.bfloat16 or not x.is_contiguous():
        raise RuntimeError("im2col_tm: contiguous bf16 [B, rows, C] required")
    B, rows, C = x.shape
    out = torch.empty((B, T_out, k * C), dtype=torch.bfloat16, device=x.device)
    with _on(x.device):
        _lib.check(_lib.load().w2l_im2col_tm(_ptr(x), _ptr(out), B, rows, C, T_out, k, stride, dilation, pad_left, _stream()), "im2col_tm")
    return out


def col2im_tm(dcol, x_rows, C, k, stride, dilation, pad_left):
    """adjoint of ``im2col_tm``: dcol [B, T_out, k*C] bf16 -> dx [B, x_rows, C] bf16."""
    _need_cuda(dcol)
    if dcol.dtype != torch.bfloat16 or not dcol.is_contiguous():
        raise RuntimeError("col2im_tm: contiguous bf16 [B, T_out, k*C] required")
    B, T_out, KC = dcol.shape
    if KC != k * C:
        raise RuntimeError("col2im_tm: dcol has %d columns, expected k*C = %d" % (KC, k * C))
    dx = torch.empty((B, x_rows, C), dtype=torch.bfloat16, device=dcol.device)
    with _on(dcol.device):
        _lib.check(_lib.load().w2l_col2im_tm(_ptr(dcol), _ptr(dx), B, x_rows, C, T_out, k, stride, dilation, pad_left, _stream()), "col2im_tm")
    return dx


def tm_to_ncw(x, T, C, x_rows=None, x_row_offset=0):
    """time-major [B, x_rows, ld] (bf16|f32) -> NCW fp32 [B, C, T]."""
    _need_cuda(x)
    B, rows, ld = x.shape
    out = torch.empty((B, C, T), dtype=torch.float32, device=x.device)
    dt = DT_BF16 if x.dtype == torch.bfloat16 else DT_F32
    with _on(x.device):
        _lib.check(_lib.load().w2l_tm_to_ncw(_ptr(x), dt, _ptr(out), B, T, C, rows, x_row_offset, ld, _stream()), "tm_to_ncw")
    return out


def bn_stats(z, C, out=None):
    """per-channel (sum, sum of squares) of bf16 ``z`` [rows, C], ADDED into ``out`` (fp32 [2C], zero-filled by the caller) when given"""
    rows = z.numel() // C
    stats = torch.zeros((2 * C,), dtype=torch.float32, device=z.device) if out is None else out
    with _on(z.device):
        _lib.check(_lib.load().w2l_bn_stats(_ptr(z), rows, C, _ptr(stats), _stream()), "bn_stats")
    return stats


def bn_finalize(stats, rows, C, gamma, beta, conv_bias, eps, momentum, running_mean, running_var, num_batches_tracked=None):
    """``num_batches_tracked`` (int64 0-dim CUDA buffer): incremented by the same kernel."""
    dev = stats.device
    out = torch.empty((4, C), dtype=torch.float32, device=dev)     # scale, shift, mean, invstd
    with _on(dev):
        _lib.check(_lib.load().w2l_bn_finalize(_ptr(stats), rows, C, _ptr(gamma), _ptr(beta), _ptr(conv_bias), eps, momentum,
                                               _ptr(running_mean), _ptr(running_var), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]),
                                               _ptr(out[3]), _ptr(num_batches_tracked), _stream()), "bn_finalize")
    return out


def lens_chain(lens, conv_params):
    """lens [B] int32/int64 CUDA, conv_params: list of (kernel, stride|0, dilation, padding) per masked conv in execution order
    -> (rows int32 [n+1, B], output lengths int64 [B]); see w2l_lens_chain."""
    _need_cuda(lens)
    if lens.dtype not in (torch.int32, torch.int64):
        lens = lens.to(torch.int64)                                  # the reference's lens.to(dtype=torch.long)
    lens = lens.contiguous()
    B, n = lens.numel(), len(conv_params)
    rows = torch.empty((n + 1, B), dtype=torch.int32, device=lens.device)
    final = torch.empty((B,), dtype=torch.int64, device=lens.device)
    table = (ctypes.c_int32 * (4 * max(n, 1)))(*[int(v) for q in conv_params for v in q])
    with _on(lens.device):
        _lib.check(_lib.load().w2l_lens_chain(_ptr(lens), int(lens.dtype == torch.int64), B, table, n, _ptr(rows), _ptr(final), _stream()),
                   "lens_chain")
    return rows, final


def bn_act_pad(z, scale, shift, B, T, C, pad_left, pad_right, act, drop_p=0.0, seed=0, lens=None, res=None, res_scale=None,
               res_shift=None, out=None, drop_mask=None):
    """``drop_mask`` (uint8 [B*T*C/8], optional) receives the dropout keep-bits for the backward pass."""
    if out is None:
        out = torch.empty((B, pad_left + T + pad_right, C), dtype=z.dtype, device=z.device)
    with _on(z.device):
        _lib.check(_lib.load().w2l_bn_act_pad(_ptr(z), _ptr(scale), _ptr(shift), _ptr(res), _ptr(res_scale), _ptr(res_shift), _ptr(out),
                                              B, T, C, pad_left, pad_right, _act_flag(act, z), float(drop_p), int(seed), _ptr(lens),
                                              _ptr(drop_mask), _stream()), "bn_act_pad")
    return out


def reflect_halo(y, T, pad_left, pad_right):
    B, rows, C = y.shape
    fn = _lib.load().w2l_reflect_halo_f32 if y.dtype == torch.float32 else _lib.load().w2l_reflect_halo
    with _on(y.device):
        _lib.check(fn(_ptr(y), B, T, C, pad_left, pad_right, _stream()), "reflect_halo")
    return y


def bn_finalize_act_pad(z, stats, gamma, beta, conv_bias, eps, momentum, running_mean, running_var, num_batches_tracked, B, T, C,
                        pad_left, pad_right, act, drop_p=0.0, seed=0, lens=None, res=None, res_scale=None, res_shift=None, drop_mask=None,
                        zero_after=None):
    """The training-mode pass: ``bn_finalize`` folded into ``bn_act_pad`` (one launch instead of two).  ``stats`` [2C] = the batch
    sums the conv epilogue left.  Returns (y bf16 [B, pl+T+pr, C], fin fp32 [4, C] = scale, shift, mean, invstd).  ``zero_after``
    (fp32 tensor, optional) is cleared by the same launch for a later kernel."""
    dev = z.device
    out = torch.empty((B, pad_left + T + pad_right, C), dtype=z.dtype, device=dev)
    fin = torch.empty((4, C), dtype=torch.float32, device=dev)
    with _on(dev):
        _lib.check(_lib.load().w2l_bn_finalize_act_pad(
            _ptr(z), _ptr(stats), B * T, _ptr(gamma), _ptr(beta), _ptr(conv_bias), float(eps), float(momentum), _ptr(running_mean),
            _ptr(running_var), _ptr(num_batches_tracked), _ptr(fin), _ptr(res), _ptr(res_scale), _ptr(res_shift), _ptr(out), B, T, C,
            pad_left, pad_right, _act_flag(act, z), float(drop_p), int(seed), _ptr(lens), _ptr(drop_mask), _ptr(zero_after),
            0 if zero_after is None else zero_after.numel(), _stream()), "bn_finalize_act_pad")
    return out, fin


def bn_act_bwd(dyp, z, scale, shift, mean, invstd, gamma, B, T, C, pad_left, pad_right, act, drop_p=0.0, seed=0, lens=None,
               res=None, res_scale=None, res_shift=None, want_g=False, dz_rows=None, drop_mask=None, red_ws=None, zero_after=None,
               red_raw=None):
    """Returns (dz bf16 [B, dz_rows, C] (rows >= T zero), red fp32 [2C] = (dbeta, dgamma), g bf16 [B,T,C] | None).
    ``red_ws`` (fp32 [2C], ZERO on entry, dirty afterwards): a persistent accumulation buffer -- the sums are then returned in a
    fresh tensor and no memset launch is needed; without it a zero-filled buffer is allocated per call.  ``zero_after`` (fp32
    tensor, optional) is cleared by the second pass for a later kernel."""
    dev = z.device
    dz_rows = T if dz_rows is None else dz_rows
    if red_raw is not None:        # the reduction came out of the backward-data GEMM's epilogue (conv1d_dgrad_wt(bnred=...)): apply only
        red, red_out = red_raw, torch.empty((2 * C,), dtype=torch.float32, device=dev)
    elif red_ws is None:
        red = torch.zeros((2 * C,), dtype=torch.float32, device=dev)
        red_out = None
    else:
        red, red_out = red_ws, torch.empty((2 * C,), dtype=torch.float32, device=dev)
    if dyp.dtype != z.dtype:
        raise RuntimeError("bn_act_bwd: upstream gradient %s and conv output %s must have the same storage type" % (dyp.dtype, z.dtype))
    dz = torch.empty((B, dz_rows, C), dtype=z.dtype, device=dev)
    g = torch.empty((B, T, C), dtype=z.dtype, device=dev) if want_g else None
    act = _act_flag(act, z)
    lib = _lib.load()
    with _on(dev):
        if red_raw is None:
            _lib.check(lib.w2l_bn_act_bwd_reduce(_ptr(dyp), _ptr(z), _ptr(res), _ptr(scale), _ptr(shift), _ptr(res_scale), _ptr(res_shift),
                                                 _ptr(mean), _ptr(invstd), _ptr(red), B, T, C, pad_left, pad_right, act, float(drop_p),
                                                 int(seed), _ptr(lens), _ptr(drop_mask), _stream()), "bn_act_bwd_reduce")
        _lib.check(lib.w2l_bn_act_bwd_apply(_ptr(dyp), _ptr(z), _ptr(res), _ptr(scale), _ptr(shift), _ptr(res_scale), _ptr(res_shift),
                                            _ptr(mean), _ptr(invstd), _ptr(gamma), _ptr(red), _ptr(dz), dz_rows, _ptr(g), B, T, C,
                                            pad_left, pad_right, act | (RED_RAW if red_raw is not None else 0), float(drop_p), int(seed),
                                            _ptr(lens), _ptr(drop_mask),
                                            _ptr(red_out), _ptr(zero_after), 0 if zero_after is None else zero_after.numel(),
                                            _stream()), "bn_act_bwd_apply")
    return dz, (red if red_out is None else red_out), g


def log_softmax(logits, C, mode=0, nan_flag=None):
    """logits [..., ld] fp32 -> [..., C] fp32 (mode 0 log_softmax, 1 softmax) over the first C columns.  ``nan_flag`` (int32 [1],
    zero) is set to 1 when an output is NaN."""
    ld = logits.shape[-1]
    rows = logits.numel() // ld
    out = torch.empty(logits.shape[:-1] + (C,), dtype=torch.float32, device=logits.device)
    with _on(logits.device):
        _lib.check(_lib.load().w2l_log_softmax(_ptr(logits), ld, _ptr(out), rows, C, mode, _ptr(nan_flag), _stream()), "log_softmax")
    return out


def log_softmax_bwd(g, lp, ld_out, gscale=None, fused_identity=False, out_dtype=torch.bfloat16):
    """g, lp [..., C] fp32 -> d logits bf16 (fp32 with ``out_dtype=torch.float32``) [..., ld_out] (zero padded)."""
    C = g.shape[-1]
    rows = g.numel() // C
    g = g.contiguous()
    out = torch.empty(g.shape[:-1] + (ld_out,), dtype=out_dtype, device=g.device)
    fn = _lib.load().w2l_log_softmax_bwd_f32 if out_dtype == torch.float32 else _lib.load().w2l_log_softmax_bwd
    with _on(g.device):
        _lib.check(fn(_ptr(g), _ptr(lp), _ptr(gscale), _ptr(out), ld_out, rows, C, int(fused_identity), _stream()), "log_softmax_bwd")
    return out


def colsum(x, C):
    ld = x.shape[-1]
    rows = x.numel() // ld
    out = torch.zeros((C,), dtype=torch.float32, device=x.device)
    fn = _lib.load().w2l_colsum_f32 if x.dtype == torch.float32 else _lib.load().w2l_colsum
    with _on(x.device):
        _lib.check(fn(_ptr(x), rows, C, ld, _ptr(out), _stream()), "colsum")
    return out


def cast_bf16(src, dst=None):
    src = src.contiguous()
    if dst is None:
        dst = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    with _on(src.device):
        _lib.check(_lib.load().w2l_cast_bf16(_ptr(src), _ptr(dst), src.numel(), _stream()), "cast_bf16")
    return dst
