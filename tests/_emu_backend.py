"""TEST INFRASTRUCTURE ONLY -- ``wav2letter_pytorch_b200.functional`` entry points answered by the library's own CUDA-core
kernels executed on the host (tests/_kernel_emu.py), on CPU tensors.

Every function below has the signature of its namesake in functional.py and repeats, launch for launch, what the C wrapper
behind it does (grid / block / shared-memory arithmetic cited per function); the kernels themselves are compiled from the
.cu sources.  ``install(monkeypatch)`` layers these over tests/_host_sim.py, so that the real modules run on a machine
without a GPU with ONLY the tcgen05 GEMMs restated in torch -- layout changes, BatchNorm statistics / apply / backward,
dropout, halos, masks, depthwise convs, log_softmax, CTC and greedy decode are the shipped kernel code.

Nothing under wav2letter_pytorch_b200/ imports this module; the `-m gpu` suite never uses it."""
import contextlib
import ctypes
import functools

import torch

import _kernel_emu as KE

BF16 = torch.bfloat16
SMS = 148


def _p(t):
    return None if t is None else t.data_ptr()


def _grid_for(items, threads):
    """csrc/elementwise.cu grid_for"""
    return int(max(1, min((items + threads - 1) // threads, SMS * 16)))


def _rows_per_block_for(rows, col_blocks):
    """csrc/elementwise.cu rows_per_block_for"""
    target = max(1, SMS * 8 // max(col_blocks, 1))
    return int(max(32, (rows + target - 1) // target))


_BN_VARIANTS = ["%d, %s, %s" % (a, d, r) for a in (0, 1, 2) for d in ("false", "true") for r in ("false", "true")]


@functools.lru_cache(maxsize=None)
def elementwise():
    kernels = ["im2col_ncw_kernel<__nv_bfloat16>", "im2col_tm_kernel", "col2im_tm_kernel", "tm_to_ncw_kernel<__nv_bfloat16>", "tm_to_ncw_kernel<float>",
               "bn_stats_kernel", "bn_finalize_kernel", "log_softmax_kernel", "log_softmax_bwd_kernel<__nv_bfloat16>", "colsum_kernel<__nv_bfloat16>", "cast_bf16_kernel",
               "lens_chain_kernel", "reflect_halo_kernel<__nv_bfloat16>", "pack_wt_kernel<__nv_bfloat16>"]
    # (the BatchNorm / activation passes run through the C wrappers of tests/_emu_cabi.py: see _via_cabi below)
    return KE.build(["elementwise.cu"], kernels, drop=ELEMENTWISE_DROP, extra=ELEMENTWISE_PTX)


@functools.lru_cache(maxsize=None)
def depthwise():
    return KE.build(["depthwise.cu"], ["depthwise_corr_kernel<__nv_bfloat16>", "depthwise_dgrad_strided_kernel<__nv_bfloat16>", "depthwise_wgrad_kernel<__nv_bfloat16>"])


@functools.lru_cache(maxsize=None)
def decode():
    return KE.build(["decode.cu"], ["greedy_argmax_kernel", "greedy_compact_kernel"])


# host stand-ins for the inline PTX of csrc/elementwise.cu: max.NaN.f32 / min.NaN.f32 (NaN-propagating, as torch.relu / clamp are)
ELEMENTWISE_PTX = r"""
static inline float max_nan(float a, float b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }
static inline float min_nan(float a, float b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }
"""
ELEMENTWISE_DROP = ["max_nan", "min_nan"]

# host stand-ins for the inline PTX of csrc/ctc.cu (file:line of what each replaces)
CTC_PTX = r"""
static inline float fast_ex2(float x) { return std::exp2(x); }                                   // ctc.cu:36  ex2.approx.ftz.f32
static inline float fast_lg2(float x) { return std::log2(x); }                                   // ctc.cu:41  lg2.approx.ftz.f32
static inline void cp_async4(void* smem, const void* gmem) { std::memcpy(smem, gmem, 4); }       // ctc.cu:58  cp.async.ca 4 B (eager)
static inline void cp_async16(void* smem, const void* gmem) { std::memcpy(smem, gmem, 16); }     // ctc.cu:61  cp.async.cg 16 B (eager)
static inline void cp_async_commit() {}                                                          // ctc.cu:64
template <int N> static inline void cp_async_wait() {}                                           // ctc.cu:65
static inline void slot_put(float4* slot, float v0, float v1, int tag, float w = 0.f) {          // ctc.cu:295 st.volatile.shared.v4
  *slot = make_float4(v0, v1, __int_as_float(tag), w);
}
static inline float4 slot_load(const float4* slot) {                                             // ctc.cu:300 ld.volatile.shared.v4
  emu::yield_poll();                       // a poll lets the producing warp run (the hardware's warps run concurrently)
  return *slot;
}
"""
CTC_POST = r"""
extern "C" int emu_ctc_plan(long long N, long long T, long long S, long long C, long long* out) {
  w2l::CtcPlan p;
  if (!w2l::make_plan(N, T, S, C, &p)) return 1;
  long long v[] = {p.R, p.threads, p.Lp, p.Cp, p.parallel, (long long)p.off_lp2, (long long)p.off_alpha, (long long)p.off_aoff,
                   (long long)p.off_beta, (long long)p.off_boff, (long long)p.off_meta, (long long)p.total, w2l::kRing, w2l::kBlk,
                   w2l::kGradFrames, (long long)sizeof(w2l::CtcMeta)};
  for (int i = 0; i < 16; ++i) out[i] = v[i];
  return 0;
}
"""


@functools.lru_cache(maxsize=None)
def ctc():
    kernels = ["ctc_prep_kernel", "ctc_grad_kernel", "ctc_finish_kernel"]
    for r in (2, 4, 8):
        kernels += ["ctc_lattice_kernel<%d>" % r, "ctc_alpha_kernel<%d>" % r, "ctc_beta_grad_kernel<%d>" % r]
    return KE.build(["ctc.cu"], kernels, drop=["fast_ex2", "fast_lg2", "cp_async4", "cp_async16", "cp_async_commit", "cp_async_wait",
                                               "slot_put", "slot_load", "launch_ctc"], extra=CTC_PTX, post=CTC_POST)


# ------------------------------------------------------------------------------------------------ layout (elementwise.cu)
def im2col_ncw(x, rows, k, stride, dilation, pad_left, pad_mode, lens=None, out_dtype=torch.bfloat16):
    assert out_dtype == torch.bfloat16, "the fp32-faithful mode is covered by the C-ABI emulation (tests/_emu_cabi.py) and the -m gpu tests"
    """w2l_im2col_ncw (elementwise.cu:713-733)"""
    x = x.contiguous().float()
    B, F, T = x.shape
    out = torch.full((B, rows, k * F), float("nan"), dtype=BF16)
    span = 31 * stride + (k - 1) * dilation + 1
    pitch = span | 1
    lens = None if lens is None else lens.to(torch.int32).contiguous()
    elementwise().launch("im2col_ncw_kernel<__nv_bfloat16>", ((rows + 31) // 32, B), 256, _p(x), _p(out), F, T, rows, k, stride, dilation, pad_left, pad_mode,
                         _p(lens), span, pitch, smem=F * pitch * 4)
    return out


def tm_to_ncw(x, T, C, x_rows=None, x_row_offset=0):
    """w2l_tm_to_ncw (elementwise.cu:765-777)"""
    x = x.contiguous()
    B, rows, ld = x.shape
    out = torch.full((B, C, T), float("nan"))
    name = "tm_to_ncw_kernel<__nv_bfloat16>" if x.dtype == BF16 else "tm_to_ncw_kernel<float>"
    elementwise().launch(name, ((T + 31) // 32, (C + 31) // 32, B), (32, 8), _p(x), _p(out), T, C, rows * ld, x_row_offset, ld, T)
    return out


def im2col_tm(x, T_out, k, stride, dilation, pad_left):
    """w2l_im2col_tm (elementwise.cu:735-746)"""
    assert x.dtype == BF16 and x.is_contiguous()
    B, rows, C = x.shape
    out = torch.full((B, T_out, k * C), float("nan"), dtype=BF16)
    elementwise().launch("im2col_tm_kernel", _grid_for(B * T_out * k * (C // 8), 256), 256, _p(x), _p(out), B, rows, C, T_out, k, stride,
                         dilation, pad_left)
    return out


def col2im_tm(dcol, x_rows, C, k, stride, dilation, pad_left):
    """w2l_col2im_tm (elementwise.cu:748-759)"""
    assert dcol.dtype == BF16 and dcol.is_contiguous()
    B, T_out, _ = dcol.shape
    dx = torch.full((B, x_rows, C), float("nan"), dtype=BF16)
    elementwise().launch("col2im_tm_kernel", _grid_for(B * x_rows * (C // 8), 256), 256, _p(dcol), _p(dx), B, x_rows, C, T_out, k, stride,
                         dilation, pad_left)
    return dx


def cast_bf16(src, dst=None):
    """w2l_cast_bf16 (elementwise.cu:889-896)"""
    src = src.contiguous().float()
    if dst is None:
        dst = torch.empty(src.shape, dtype=BF16)
    assert dst.is_contiguous()
    n = src.numel()
    if n:
        elementwise().launch("cast_bf16_kernel", _grid_for(n // 8 + 1, 256), 256, _p(src), _p(dst), n)
    return dst


def pack_wt(w_store, wt, cout, cin):
    """w2l_pack_wt (elementwise.cu:949-956)"""
    w_store = w_store.contiguous()
    k = w_store.shape[0]
    co_pad, ci_pad = wt.shape[2], wt.shape[1]
    elementwise().launch("pack_wt_kernel<__nv_bfloat16>", ((ci_pad + 31) // 32, (co_pad + 31) // 32, k), (32, 8), _p(w_store), _p(wt), k, cout, cin, co_pad,
                         ci_pad)
    return wt


def reflect_halo(y, T, pad_left, pad_right):
    """w2l_reflect_halo (elementwise.cu:919-928)"""
    B, rows, C = y.shape
    if pad_left + pad_right:
        elementwise().launch("reflect_halo_kernel<__nv_bfloat16>", _grid_for(B * (pad_left + pad_right) * (C // 8), 256), 256, _p(y), B, T, C, pad_left,
                             pad_right)
    return y


class _LensChain(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32), ("k", ctypes.c_int16 * 192), ("s", ctypes.c_int16 * 192), ("d", ctypes.c_int16 * 192),
                ("p", ctypes.c_int16 * 192)]


def lens_chain(lens, conv_params):
    """w2l_lens_chain (elementwise.cu:958-976)"""
    if lens.dtype not in (torch.int32, torch.int64):
        lens = lens.to(torch.int64)
    lens = lens.contiguous()
    B, n = lens.numel(), len(conv_params)
    rows = torch.full((n + 1, B), -1, dtype=torch.int32)
    final = torch.full((B,), -1, dtype=torch.int64)
    c = _LensChain()
    c.n = n
    for j, (k, s, d, p) in enumerate(conv_params):
        c.k[j], c.s[j], c.d[j], c.p[j] = int(k), int(s), int(d), int(p)
    elementwise().launch("lens_chain_kernel", (B + 127) // 128, 128, _p(lens), int(lens.dtype == torch.int64), B, ctypes.addressof(c),
                         _p(rows), _p(final))
    return rows, final


# ------------------------------------------------------------------------------------------------ BatchNorm + activation
def bn_stats(z, C, out=None):
    """w2l_bn_stats (elementwise.cu:779-787); accumulates into ``out`` when given, as the kernel does"""
    z = z.contiguous()
    rows = z.numel() // C
    stats = torch.zeros((2 * C,)) if out is None else out
    col_blocks = (C + 255) // 256
    rpb = _rows_per_block_for(rows, col_blocks)
    elementwise().launch("bn_stats_kernel", (col_blocks, (rows + rpb - 1) // rpb), (32, 8), _p(z), rows, C, _p(stats), rpb)
    return stats


def bn_finalize(stats, rows, C, gamma, beta, conv_bias, eps, momentum, running_mean, running_var, num_batches_tracked=None):
    """w2l_bn_finalize (elementwise.cu:789-800)"""
    out = torch.full((4, C), float("nan"))
    f32 = lambda t: None if t is None else t.detach()                       # noqa: E731
    elementwise().launch("bn_finalize_kernel", (C + 127) // 128, 128, _p(stats), rows, C, _p(f32(gamma)), _p(f32(beta)), _p(f32(conv_bias)),
                         float(eps), float(momentum), _p(running_mean), _p(running_var), _p(out[0]), _p(out[1]), _p(out[2]), _p(out[3]),
                         _p(num_batches_tracked))
    return out


# The three BatchNorm / activation passes own their launch geometry in the C wrappers since round 2 (row-looping grids, argument
# structs, the finalize fold), so they are no longer repeated here: these run the UNMODIFIED functional.py wrapper over the library's
# own extern "C" entry points compiled for the host (tests/_emu_cabi.py) -- the kernels' source executes exactly as before.
def _via_cabi(name):
    def call(*a, **k):
        import _emu_cabi
        from wav2letter_pytorch_b200 import _lib
        from wav2letter_pytorch_b200 import functional as Freal
        saved = {"load": _lib.load, "_need_cuda": Freal._need_cuda, "_stream": Freal._stream, "torch": Freal.torch, "device": torch.cuda.device}
        lib = _emu_cabi.library()
        try:
            _lib.load = lambda: lib
            Freal._need_cuda = lambda *ts: None
            Freal._stream = lambda: None
            Freal.torch = _emu_cabi._TorchProxy()
            torch.cuda.device = lambda *aa, **kk: contextlib.nullcontext()
            return _ORIG[name](*a, **k)
        finally:
            _lib.load, Freal._need_cuda, Freal._stream, Freal.torch = saved["load"], saved["_need_cuda"], saved["_stream"], saved["torch"]
            torch.cuda.device = saved["device"]
    call.__name__ = name
    return call


from wav2letter_pytorch_b200 import functional as _Freal  # noqa: E402
_ORIG = {n: getattr(_Freal, n) for n in ("bn_act_pad", "bn_finalize_act_pad", "bn_act_bwd")}
bn_act_pad = _via_cabi("bn_act_pad")
bn_finalize_act_pad = _via_cabi("bn_finalize_act_pad")
bn_act_bwd = _via_cabi("bn_act_bwd")


# ------------------------------------------------------------------------------------------------ head
def log_softmax(logits, C, mode=0, nan_flag=None):
    """w2l_log_softmax (elementwise.cu:868-873)"""
    logits = logits.contiguous()
    ld = logits.shape[-1]
    rows = logits.numel() // ld
    out = torch.full(logits.shape[:-1] + (C,), float("nan"))
    elementwise().launch("log_softmax_kernel", (rows + 7) // 8, 256, _p(logits), ld, _p(out), rows, C, mode, _p(nan_flag))
    return out


def log_softmax_bwd(g, lp, ld_out, gscale=None, fused_identity=False, out_dtype=torch.bfloat16):
    assert out_dtype == torch.bfloat16
    """w2l_log_softmax_bwd (elementwise.cu:875-883)"""
    C = g.shape[-1]
    rows = g.numel() // C
    g = g.contiguous()
    lp = None if lp is None else lp.contiguous()
    out = torch.full(g.shape[:-1] + (ld_out,), float("nan"), dtype=BF16)
    elementwise().launch("log_softmax_bwd_kernel<__nv_bfloat16>", (rows + 7) // 8, 256, _p(g), _p(lp), _p(gscale), _p(out), ld_out, rows, C,
                         int(fused_identity))
    return out


def colsum(x, C):
    """w2l_colsum (elementwise.cu:885-893)"""
    x = x.contiguous()
    ld = x.shape[-1]
    rows = x.numel() // ld
    out = torch.zeros((C,))
    col_blocks = (C + 31) // 32
    rpb = _rows_per_block_for(rows, col_blocks)
    elementwise().launch("colsum_kernel<__nv_bfloat16>", (col_blocks, (rows + rpb - 1) // rpb), (32, 8), _p(x), rows, C, ld, _p(out), rpb)
    return out


# ------------------------------------------------------------------------------------------------ depthwise (depthwise.cu)
def depthwise_fwd(x, w, T_out, k, stride, dilation, pad, out_lens=None):
    """w2l_depthwise_fwd (depthwise.cu:161-171)"""
    x, w = x.contiguous(), w.contiguous()
    B, T, C = x.shape
    y = torch.full((B, T_out, C), float("nan"), dtype=BF16)
    ol = None if out_lens is None else out_lens.to(torch.int32).contiguous()
    depthwise().launch("depthwise_corr_kernel<__nv_bfloat16>", ((C // 8 + 31) // 32, (B * T_out + 7) // 8), (32, 8), _p(x), _p(w), _p(y), B, T, T_out, C, k,
                       stride, dilation, -pad, 0, None, _p(ol))
    return y


def depthwise_dgrad(dy, w, T, k, dilation, pad, dy_lens=None, stride=1):
    """w2l_depthwise_dgrad / w2l_depthwise_dgrad_strided (depthwise.cu:173-196)"""
    dy, w = dy.contiguous(), w.contiguous()
    B, T_out, C = dy.shape
    dx = torch.full((B, T, C), float("nan"), dtype=BF16)
    dl = None if dy_lens is None else dy_lens.to(torch.int32).contiguous()
    grid = ((C // 8 + 31) // 32, (B * T + 7) // 8)
    if stride == 1:
        depthwise().launch("depthwise_corr_kernel<__nv_bfloat16>", grid, (32, 8), _p(dy), _p(w), _p(dx), B, T_out, T, C, k, 1, dilation,
                           pad - (k - 1) * dilation, 1, _p(dl), None)
    else:
        depthwise().launch("depthwise_dgrad_strided_kernel<__nv_bfloat16>", grid, (32, 8), _p(dy), _p(w), _p(dx), B, T, T_out, C, k, stride, dilation, pad,
                           _p(dl))
    return dx


def depthwise_wgrad(dy, x, k, stride, dilation, pad, dy_lens=None):
    """w2l_depthwise_wgrad (depthwise.cu:198-211)"""
    dy, x = dy.contiguous(), x.contiguous()
    B, T_out, C = dy.shape
    T = x.shape[1]
    dw = torch.zeros((k, C))
    dl = None if dy_lens is None else dy_lens.to(torch.int32).contiguous()
    rows = B * T_out
    rpb = max(64, (rows + SMS - 1) // SMS)
    depthwise().launch("depthwise_wgrad_kernel<__nv_bfloat16>", ((C // 8 + 31) // 32, (rows + rpb - 1) // rpb, (k + 3) // 4), (32, 8), _p(dy), _p(x), _p(dw),
                       B, T, T_out, C, k, stride, dilation, pad, _p(dl), rpb)
    return dw


# ------------------------------------------------------------------------------------------------ CTC (ctc.cu)
def ctc_loss_raw(x, targets, input_lengths, target_lengths, blank=0, zero_infinity=True, reduction_mean=True, from_logits=False,
                 need_grad=True, serial=False, return_plan=False):
    """w2l_ctc_loss + launch_ctc (ctc.cu:820-918), launch for launch.  ``serial`` forces the alpha -> beta+grad schedule."""
    x = x.float()
    if x.stride(2) != 1:
        x = x.contiguous()
    N, T, C = x.shape
    targets = targets.to(torch.int32).contiguous()
    il, tl = input_lengths.to(torch.int32).contiguous(), target_lengths.to(torch.int32).contiguous()
    S = targets.shape[1]
    K = ctc()
    out = (ctypes.c_longlong * 16)()
    if K.lib.emu_ctc_plan(ctypes.c_longlong(N), ctypes.c_longlong(T), ctypes.c_longlong(S), ctypes.c_longlong(C), out) != 0:
        raise RuntimeError("ctc_loss: target length %d not supported" % S)
    R, threads, Lp, Cp, parallel, o_lp2, o_alpha, o_aoff, o_beta, o_boff, o_meta, total, kRing, kBlk, kGradFrames, meta_sz = list(out)
    assert meta_sz == 16
    ws = torch.full((total + 256,), 0xFF, dtype=torch.uint8)              # poisoned workspace, 256-byte aligned base
    base = (ws.data_ptr() + 255) // 256 * 256
    nll, loss = torch.full((N,), float("nan")), torch.full((1,), float("nan"))
    grad = torch.full((N, T, C), float("nan")) if need_grad else None
    p = lambda off: base + off                                            # noqa: E731
    zi, rm = int(zero_infinity), int(reduction_mean)
    K.launch("ctc_prep_kernel", (N * T + 7) // 8, 256, _p(x), int(from_logits), N, T, C, x.stride(0), x.stride(1), _p(il), p(o_lp2), Cp)
    tgp = _p(targets) if S > 0 else None
    if parallel and need_grad and not serial:
        n_blk = (T + kBlk - 1) // kBlk
        smem_l = 2 * kBlk * Cp * 4 + 33 * kBlk * 16 + (32 + 2) * 4
        K.launch("ctc_lattice_kernel<%d>" % R, 2 * N, threads, p(o_lp2), N, T, Cp, tgp, S, _p(il), _p(tl), blank, p(o_alpha), p(o_aoff),
                 p(o_beta), p(o_boff), n_blk, p(o_meta), Lp, smem=smem_l)
        smem_g = 8 * Cp * 8 + (2 * S + 1 + 15)
        K.launch("ctc_grad_kernel", ((T + kGradFrames - 1) // kGradFrames, N), 256, p(o_lp2), T, C, Cp, tgp, S, _p(il), _p(tl), blank,
                 p(o_alpha), p(o_aoff), p(o_beta), p(o_boff), n_blk, p(o_meta), Lp, zi, rm, N, _p(grad), smem=smem_g)
    else:
        smem_a = (kRing * Cp + 2 * 32 * 2 + 32 + 2) * 4
        smem_b = (kRing * Cp + kRing * threads * R + 2 * 32 * 2 + 32 + 2 * Cp) * 4
        K.launch("ctc_alpha_kernel<%d>" % R, N, threads, p(o_lp2), T, Cp, tgp, S, _p(il), _p(tl), blank, p(o_alpha), p(o_aoff), p(o_meta), Lp,
                 0, smem=smem_a)
        if need_grad:
            K.launch("ctc_beta_grad_kernel<%d>" % R, N, threads, p(o_lp2), T, C, Cp, tgp, S, _p(il), _p(tl), blank, p(o_alpha), p(o_aoff),
                     p(o_meta), Lp, zi, rm, N, _p(grad), smem=smem_b)
    K.launch("ctc_finish_kernel", 1, 256, p(o_meta), _p(tl), S, N, zi, rm, _p(nll), _p(loss))
    if return_plan:
        return loss, nll, grad, dict(R=R, threads=threads, parallel=bool(parallel))
    return loss, nll, grad


# ------------------------------------------------------------------------------------------------ greedy decode (decode.cu)
def greedy_decode(scores, sizes=None, blank=0):
    """w2l_greedy_decode (decode.cu:148-177), launch for launch; any [N, T, C] view with unit class stride"""
    scores = scores.float()
    if scores.stride(2) != 1:
        scores = scores.contiguous()
    N, T, C = scores.shape
    chunk = 256
    nchunks = max(1, (T + chunk - 1) // chunk)
    am = torch.full((N, T), -7, dtype=torch.int32)
    tok, off = torch.full((N, T), -7, dtype=torch.int32), torch.full((N, T), -7, dtype=torch.int32)
    cnt, cc = torch.full((N,), -7, dtype=torch.int32), torch.full((N * nchunks,), -7, dtype=torch.int32)
    sz = None if sizes is None else torch.as_tensor(sizes).to(torch.int32).contiguous()
    if N == 0:
        return am, tok, off, cnt
    if T == 0:
        cnt.zero_()
        return am, tok, off, cnt
    D = decode()
    D.launch("greedy_argmax_kernel", (nchunks, N), chunk, _p(scores), T, C, scores.stride(0), scores.stride(1), _p(sz), blank, _p(am), _p(cc),
             nchunks, smem=(4 + chunk * C) * 4)
    D.launch("greedy_compact_kernel", (nchunks, N), chunk, _p(am), T, _p(sz), blank, _p(cc), nchunks, _p(tok), _p(off), _p(cnt))
    return am, tok, off, cnt


# ------------------------------------------------------------------------------------------------ NovoGrad (novograd.cu)
@functools.lru_cache(maxsize=None)
def novograd():
    return KE.build(["novograd.cu"], ["novograd_norm_kernel", "novograd_moment_kernel", "novograd_update_kernel"])


def novograd_step(params, grads, exp_avg, exp_avg_sq, max_exp_avg_sq, shadows, lr, beta1, beta2, eps, weight_decay, grad_averaging):
    """w2l_novograd_step (novograd.cu:126-145) over lists of fp32 tensors; ``exp_avg_sq`` / ``max_exp_avg_sq`` are fp32 [n_tensors]
    (one second moment per tensor), ``shadows`` an optional list of bf16 tensors refreshed with the new parameters.  The chunk
    table is built as novograd.py's ``_plan`` builds it: kNgChunk elements per CTA."""
    chunk = 16384
    n = len(params)
    numel = torch.tensor([p.numel() for p in params], dtype=torch.int64)
    per = [(int(c) + chunk - 1) // chunk for c in numel]
    prefix = torch.tensor([sum(per[:i]) for i in range(n)], dtype=torch.int32)
    table = lambda ts: torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64)      # noqa: E731
    tp, tg, tm = table(params), table(grads), table(exp_avg)
    ts = None if shadows is None else table(shadows)
    norms = torch.zeros(n)
    K = novograd()
    K.launch("novograd_norm_kernel", sum(per), 256, _p(tg), _p(numel), _p(prefix), n, _p(norms))
    K.launch("novograd_moment_kernel", (n + 127) // 128, 128, _p(exp_avg_sq), _p(max_exp_avg_sq), _p(norms), n, float(beta2), float(eps))
    K.launch("novograd_update_kernel", sum(per), 256, _p(tp), _p(tg), _p(tm), _p(norms), _p(ts), _p(numel), _p(prefix), n, float(lr),
             float(beta1), float(weight_decay), int(grad_averaging))


# ------------------------------------------------------------------------------------------------ WER / CER (metrics.cu)
@functools.lru_cache(maxsize=None)
def metrics():
    return KE.build(["metrics.cu"], ["metrics_split_kernel", "edit_distance_kernel<int32_t, 8>", "edit_distance_kernel<long long, 8>",
                                     "edit_distance_kernel<int32_t, 32>", "edit_distance_kernel<long long, 32>", "metrics_finalize_kernel"])


def string_metrics(tokens, counts, space_index, ref_ids, ref_lens, cer_den, wer_den, len_den):
    """w2l_string_metrics (metrics.cu:141-181): tokens [N, T] / ref_ids [N, S] int32 -> fp32 [3] = (cer, wer, len_ratio)"""
    tokens, counts = tokens.to(torch.int32).contiguous(), counts.to(torch.int32).contiguous()
    ref_ids, ref_lens = ref_ids.to(torch.int32).contiguous(), ref_lens.to(torch.int32).contiguous()
    N, T = tokens.shape
    S = ref_ids.shape[1]
    assert 1 <= S <= 1023
    i64 = lambda *shape: torch.full(shape, -1, dtype=torch.int64)          # noqa: E731
    i32 = lambda *shape: torch.full(shape, -1, dtype=torch.int32)          # noqa: E731
    h_words, r_words, h_chars, r_chars = i64(N, T), i64(N, S), i32(N, T), i32(N, S)
    h_nc, h_nw, r_nc, r_nw, cer_d, wer_d = (i32(N) for _ in range(6))
    ratios = torch.full((3,), float("nan"))
    K = metrics()
    # one CTA of 128 threads per utterance (kSplitThreads); one warp per pair, 4 warps per CTA (kEditWarps), 8 / 32 columns per lane
    K.launch("metrics_split_kernel", N, 128, _p(tokens), _p(counts), N, T, space_index, _p(h_chars), _p(h_nc), _p(h_words), _p(h_nw))
    K.launch("metrics_split_kernel", N, 128, _p(ref_ids), _p(ref_lens), N, S, space_index, _p(r_chars), _p(r_nc), _p(r_words), _p(r_nw))
    kk = 8 if S <= 256 else 32
    K.launch("edit_distance_kernel<int32_t, %d>" % kk, (N + 3) // 4, 128, _p(h_chars), _p(h_nc), T, _p(r_chars), _p(r_nc), S, _p(cer_d), N)
    K.launch("edit_distance_kernel<long long, %d>" % kk, (N + 3) // 4, 128, _p(h_words), _p(h_nw), T, _p(r_words), _p(r_nw), S, _p(wer_d), N)
    K.launch("metrics_finalize_kernel", 1, 256, _p(cer_d), _p(wer_d), _p(counts), N, float(cer_den), float(wer_den), float(len_den),
             _p(ratios))
    return ratios, cer_d, wer_d


# ------------------------------------------------------------------------------------------------ feature front-end (features.cu)
@functools.lru_cache(maxsize=None)
def features():
    return KE.build(["features.cu"], ["logmel_kernel", "feat_norm_kernel"])


def logmel_features(audio, lens, noise, window, fb, n_fft, win_length, hop, dither=1e-5, preemph=0.97, log_guard=2.0 ** -24, eps=1e-5):
    """w2l_logmel_features (features.cu:180-214): audio [B, Lmax] fp32 zero padded, lens int32 [B], noise [B, Lmax] | None, window
    [win_length], fb [n_mels, n_fft/2+1] -> (features [B, n_mels, T_max] fp32 zero padded, frames per utterance)"""
    audio = audio.float().contiguous()
    lens = lens.to(torch.int32).contiguous()
    noise = None if noise is None else noise.float().contiguous()
    window, fb = window.float().contiguous(), fb.float().contiguous()
    B, Lmax = audio.shape
    n_mels = fb.shape[-2]
    T_max = 1 + int(lens.max()) // hop
    feats = torch.full((B, T_max, n_mels), float("nan"))
    out = torch.full((B, n_mels, T_max), float("nan"))
    log2_fft = n_fft.bit_length() - 1
    n_bins = n_fft // 2 + 1
    fb_pitch = n_bins | 1
    warps = 8
    smem = (n_mels * fb_pitch + ((n_mels * fb_pitch) & 1)) * 4 + (n_fft // 2) * 8 + (win_length + (win_length & 1)) * 4 + warps * n_fft * 8 + n_mels * 8
    bx = max(1, min((2 * SMS + B - 1) // B, (T_max + 4 * warps - 1) // (4 * warps)))
    K = features()
    K.launch("logmel_kernel", (bx, B), warps * 32, _p(audio), audio.stride(0), _p(noise), _p(lens), n_fft, log2_fft, win_length, hop,
             _p(window), _p(fb), n_mels, float(dither), float(preemph), float(log_guard), _p(feats), T_max, smem=smem)
    K.launch("feat_norm_kernel", ((n_mels + 31) // 32, B), (32, 8), _p(feats), _p(lens), hop, T_max, n_mels, float(eps), _p(out))
    return out, (lens // hop + 1).to(torch.int32)


# ------------------------------------------------------------------------------------------------ conv GEMM (conv_gemm.cu)
# the two inline-PTX statements inside conv_gemm_kernel's body, replaced textually (everything else in that kernel is C++ over the
# wrappers of common.cuh, whose functional stand-ins live in tests/kernel_emu_runtime.h)
GEMM_SUBS = [(r'asm volatile\("bar\.sync 1, 128;" ::: "memory"\);', "emu::named_barrier(1, 128);"),
             (r'asm volatile\("red\.global\.add\.v4\.f32 \[%0\], \{%1, %2, %3, %4\};" ::"l"\((.*?)\), "f"\((.*?)\),\s*"f"\((.*?)\), '
              r'"f"\((.*?)\), "f"\((.*?)\)\s*: "memory"\);', r"emu_red_add_v4(\1, \2, \3, \4, \5);")]
# the CTA-pair kernel (cluster of 2, tcgen05 cta_group::2) is outside the emulated surface: the emulated library always takes the
# single-CTA kernel, as the real one does under W2L_CG2=0; the pair kernel's parity is the -m gpu tests' job
GEMM_DROP = {"conv_gemm_cg2_kernel": "", "conv_wgrad_cg2_kernel": "",
             "launch_wgrad_cg2": "\nstatic int launch_wgrad_cg2(const GemmParams&, int, cudaStream_t) { set_error(\"no CTA-pair kernel on the host\"); return W2L_ERR_CUDA; }",
             "launch_gemm_cg2": "\nstatic int launch_gemm_cg2(const GemmParams&, cudaStream_t) { set_error(\"no CTA-pair kernel on the host\"); return W2L_ERR_CUDA; }",
             "cg2_wanted": "\nstatic bool cg2_wanted() { return false; }"}
GEMM_POST = r"""
extern "C" const char* emu_last_error() { return w2l::g_err; }
extern "C" long long emu_launch_count() { return w2l::g_launches; }
extern "C" void emu_set_sm_budget(int sms) { w2l::g_sm_budget = sms; }
"""


@functools.lru_cache(maxsize=None)
def gemm():
    """conv_gemm.cu WITH its extern "C" wrappers (descriptor checks, tile / ring / stream-K / tail-split planning, tensor maps):
    the emulated library exports w2l_conv1d_fwd / _dgrad / _dgrad_wt / _wgrad / _wgrad_splits / w2l_set_gemm_scratch themselves"""
    from wav2letter_pytorch_b200 import _lib
    K = KE.build(["conv_gemm.cu"], [], helpers_from_common=("pack_bf16x2", "make_smem_desc", "make_idesc_bf16"), subs=GEMM_SUBS,
                 drop=GEMM_DROP, c_abi=True, post=GEMM_POST, opt="-O2")
    for name in ("w2l_conv1d_fwd", "w2l_conv1d_dgrad", "w2l_conv1d_dgrad_wt", "w2l_conv1d_dgrad_wt_bnred", "w2l_conv1d_wgrad", "w2l_conv1d_wgrad_splits",
                 "w2l_set_gemm_scratch", "w2l_conv1d_fwd_tail_parts"):
        res, args = _lib.SIGNATURES[name]
        fn = getattr(K.lib, name)
        fn.restype, fn.argtypes = res, args
    K.lib.emu_last_error.restype = ctypes.c_char_p
    K.lib.emu_launch_count.restype = ctypes.c_longlong
    return K


def _gemm_check(rc, what):
    if rc:
        raise RuntimeError("emulated %s failed (code %d): %s" % (what, rc, gemm().lib.emu_last_error().decode()))


_scratch = []


def ensure_gemm_scratch(device=None):
    """functional.ensure_gemm_scratch: the zero-filled fp32 scratch of the forward tail split"""
    if not _scratch:
        nbytes = 4096 + 148 * 128 * 256 * 4
        buf = torch.zeros((nbytes + 256,), dtype=torch.uint8)
        base = (buf.data_ptr() + 255) // 256 * 256
        _gemm_check(gemm().lib.w2l_set_gemm_scratch(base, nbytes), "set_gemm_scratch")
        _scratch.append(buf)


def conv1d_fwd(x, w, desc, y, bias=None, scale=None, shift=None, bn_stats=None):
    ensure_gemm_scratch()
    det = lambda t: None if t is None else t.detach()                       # noqa: E731
    _gemm_check(gemm().lib.w2l_conv1d_fwd(_p(x), _p(w), _p(det(bias)), _p(det(scale)), _p(det(shift)), _p(bn_stats), _p(y), ctypes.byref(desc),
                                          None), "conv1d_fwd")
    return y


def conv1d_dgrad(dy, w, desc, dx):
    _gemm_check(gemm().lib.w2l_conv1d_dgrad(_p(dy), _p(w), _p(dx), ctypes.byref(desc), None), "conv1d_dgrad")
    return dx


def conv1d_dgrad_wt(dy, wt, desc, dx, bnred=None):
    if bnred is None:
        _gemm_check(gemm().lib.w2l_conv1d_dgrad_wt(_p(dy), _p(wt), _p(dx), ctypes.byref(desc), None), "conv1d_dgrad_wt")
        return dx
    from wav2letter_pytorch_b200._lib import BnReduce
    q = lambda t: None if t is None else t.data_ptr()       # noqa: E731
    r = BnReduce(q(bnred["z"]), q(bnred.get("mask")), q(bnred["scale"]), q(bnred["shift"]), q(bnred["mean"]), q(bnred.get("lens")), q(bnred["red"]),
                 bnred["B"], bnred["T"], bnred["pad_left"], bnred["pad_right"], bnred["act"], float(bnred.get("drop_p", 0.0)))
    _gemm_check(gemm().lib.w2l_conv1d_dgrad_wt_bnred(_p(dy), _p(wt), _p(dx), ctypes.byref(desc), ctypes.byref(r), None), "conv1d_dgrad_wt_bnred")
    return dx


def conv1d_wgrad(dy, x, desc, dw):
    lib = gemm().lib
    if lib.w2l_conv1d_wgrad_splits(ctypes.byref(desc)) > 1:
        dw.zero_()
    _gemm_check(lib.w2l_conv1d_wgrad(_p(dy), _p(x), _p(dw), ctypes.byref(desc), None), "conv1d_wgrad")
    return dw


# ------------------------------------------------------------------------------------------------ install
_NAMES = ["im2col_ncw", "tm_to_ncw", "im2col_tm", "col2im_tm", "cast_bf16", "pack_wt", "bn_stats", "bn_finalize", "lens_chain",
          "bn_act_pad", "bn_finalize_act_pad", "reflect_halo", "bn_act_bwd", "log_softmax", "log_softmax_bwd", "colsum", "depthwise_fwd", "depthwise_dgrad",
          "depthwise_wgrad", "ctc_loss_raw", "greedy_decode"]
_GEMM_NAMES = ["conv1d_fwd", "conv1d_dgrad", "conv1d_dgrad_wt", "conv1d_wgrad", "ensure_gemm_scratch"]


def install(monkeypatch, gemm_too=True):
    """Every ``functional`` entry point the models call answered by the library's own kernel source on the host; with
    ``gemm_too=False`` the tcgen05 GEMMs stay with the torch restatement of tests/_host_sim.py"""
    import _host_sim
    F = _host_sim.install(monkeypatch)
    if gemm_too:
        from wav2letter_pytorch_b200 import layers
        monkeypatch.setattr(layers.FusedBnReduce, "enabled", True)        # this backend's backward-data GEMM has the fused reduction
    for n in _NAMES + (_GEMM_NAMES if gemm_too else []):
        assert hasattr(F, n), n
        monkeypatch.setattr(F, n, globals()[n])
    return F
