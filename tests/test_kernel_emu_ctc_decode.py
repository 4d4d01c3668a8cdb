"""The CTC loss + gradient kernels and the greedy decode kernels (kernels 2 and 3), executed on the host from
their own source text (tests/_kernel_emu.py) in the launch sequences of ``w2l_ctc_loss`` / ``w2l_greedy_decode`` and held to
the oracle: transcripts bit-exact, CTC loss within 1e-4 relative, gradient within 2e-3 of its largest element -- the bars of
the `-m gpu` tests in tests/test_gpu_kernels.py, on cases small enough for a fiber-per-thread emulation.

Both CTC schedules run: the parallel one (alpha and beta lattices in separate CTAs with the tagged-slot wavefront across
warps, then the per-frame gradient kernel) and the serial one (alpha, then beta fused with the gradient).  The inline-PTX
helpers of ctc.cu are replaced by host functions below; everything else is the library's code, including ``make_plan``."""
import ctypes

import numpy as np
import pytest
import torch

import _kernel_emu as KE
from oracle import w2l_oracle as O

pytestmark = pytest.mark.skipif(not KE.available(), reason="needs g++ and the CUDA headers")

# host stand-ins for the inline PTX of csrc/ctc.cu (file:line of what each replaces)
CTC_PTX = r"""
static inline float fast_ex2(float x) { return std::exp2(x); }                                   // ctc.cu:36  ex2.approx.ftz.f32
static inline float fast_lg2(float x) { return std::log2(x); }                                   // ctc.cu:41  lg2.approx.ftz.f32
static inline void cp_async4(void* smem, const void* gmem) { std::memcpy(smem, gmem, 4); }       // ctc.cu:58  cp.async.ca 4 B (eager)
static inline void cp_async16(void* smem, const void* gmem) { std::memcpy(smem, gmem, 16); }     // ctc.cu:61  cp.async.cg 16 B (eager)
static inline void cp_async_commit() {}                                                          // ctc.cu:64
template <int N> static inline void cp_async_wait() {}                                           // ctc.cu:65
static inline void slot_put(float4* slot, float v0, float v1, int tag) {                         // ctc.cu:292 st.volatile.shared.v4
  *slot = make_float4(v0, v1, __int_as_float(tag), 0.f);
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


@pytest.fixture(scope="module")
def ctc():
    kernels = ["ctc_prep_kernel", "ctc_grad_kernel", "ctc_finish_kernel"]
    for r in (2, 4, 8):
        kernels += ["ctc_lattice_kernel<%d>" % r, "ctc_alpha_kernel<%d>" % r, "ctc_beta_grad_kernel<%d>" % r]
    return KE.build(["ctc.cu"], kernels, drop=["fast_ex2", "fast_lg2", "cp_async4", "cp_async16", "cp_async_commit", "cp_async_wait",
                                               "slot_put", "slot_load", "launch_ctc"], extra=CTC_PTX, post=CTC_POST)


def emu_ctc_loss(ctc, x, targets, il, tl, from_logits=False, serial=False, blank=0, zero_infinity=1, reduction_mean=1):
    """csrc/ctc.cu w2l_ctc_loss + launch_ctc, launch for launch"""
    x = x.float().contiguous()
    N, T, C = x.shape
    targets, il, tl = targets.int().contiguous(), il.int().contiguous(), tl.int().contiguous()
    S = targets.shape[1]
    out = (ctypes.c_longlong * 16)()
    assert ctc.lib.emu_ctc_plan(ctypes.c_longlong(N), ctypes.c_longlong(T), ctypes.c_longlong(S), ctypes.c_longlong(C), out) == 0
    R, threads, Lp, Cp, parallel, o_lp2, o_alpha, o_aoff, o_beta, o_boff, o_meta, total, kRing, kBlk, kGradFrames, meta_sz = list(out)
    assert meta_sz == 16
    ws = torch.full((total + 256,), 0xFF, dtype=torch.uint8)              # poisoned workspace, 256-byte aligned base
    base = (ws.data_ptr() + 255) // 256 * 256
    nll, loss = torch.full((N,), float("nan")), torch.full((1,), float("nan"))
    grad = torch.full((N, T, C), float("nan"))
    p = lambda off: base + off                                            # noqa: E731
    ctc.launch("ctc_prep_kernel", (N * T + 7) // 8, 256, x.data_ptr(), int(from_logits), N, T, C, x.stride(0), x.stride(1), il.data_ptr(),
               p(o_lp2), Cp)
    tgp = targets.data_ptr() if S > 0 else None
    if parallel and not serial:
        n_blk = (T + kBlk - 1) // kBlk
        smem_l = 2 * kBlk * Cp * 4 + 33 * kBlk * 16 + (32 + 2) * 4
        ctc.launch("ctc_lattice_kernel<%d>" % R, 2 * N, threads, p(o_lp2), N, T, Cp, tgp, S, il.data_ptr(), tl.data_ptr(), blank,
                   p(o_alpha), p(o_aoff), p(o_beta), p(o_boff), n_blk, p(o_meta), Lp, smem=smem_l)
        smem_g = 8 * Cp * 8 + (2 * S + 1 + 15)
        ctc.launch("ctc_grad_kernel", ((T + kGradFrames - 1) // kGradFrames, N), 256, p(o_lp2), T, C, Cp, tgp, S, il.data_ptr(), tl.data_ptr(),
                   blank, p(o_alpha), p(o_aoff), p(o_beta), p(o_boff), n_blk, p(o_meta), Lp, zero_infinity, reduction_mean, N,
                   grad.data_ptr(), smem=smem_g)
    else:
        smem_a = (kRing * Cp + 2 * 32 * 2 + 32 + 2) * 4
        smem_b = (kRing * Cp + kRing * threads * R + 2 * 32 * 2 + 32 + 2 * Cp) * 4
        ctc.launch("ctc_alpha_kernel<%d>" % R, N, threads, p(o_lp2), T, Cp, tgp, S, il.data_ptr(), tl.data_ptr(), blank, p(o_alpha), p(o_aoff),
                   p(o_meta), Lp, 0, smem=smem_a)
        ctc.launch("ctc_beta_grad_kernel<%d>" % R, N, threads, p(o_lp2), T, C, Cp, tgp, S, il.data_ptr(), tl.data_ptr(), blank, p(o_alpha),
                   p(o_aoff), p(o_meta), Lp, zero_infinity, reduction_mean, N, grad.data_ptr(), smem=smem_b)
    ctc.launch("ctc_finish_kernel", 1, 256, p(o_meta), tl.data_ptr(), S, N, zero_infinity, reduction_mean, nll.data_ptr(), loss.data_ptr())
    return loss, nll, grad, dict(R=R, threads=threads, parallel=bool(parallel))


def check_ctc(ctc, lp, tg, il, tl, from_logits=False, serial=False, loss_tol=1e-4, grad_tol=2e-3):
    lp = torch.as_tensor(lp, dtype=torch.float32)
    tg, il, tl = (torch.as_tensor(v, dtype=torch.int32) for v in (tg, il, tl))
    ref_in = torch.log_softmax(lp.double(), -1) if from_logits else lp.double()
    loss_ref, grad_ref = O.ctc_loss_torch(ref_in, tg, il, tl, dtype=torch.float64)
    if from_logits:
        x = lp.double().clone().requires_grad_(True)
        l = torch.nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)(torch.log_softmax(x, -1).transpose(0, 1), tg, il, tl)
        (grad_ref,) = torch.autograd.grad(l, x)
    loss, nll, grad, plan = emu_ctc_loss(ctc, lp, tg, il, tl, from_logits=from_logits, serial=serial)
    assert abs(loss.item() - loss_ref.item()) <= loss_tol * max(1.0, abs(loss_ref.item())), (loss.item(), loss_ref.item())
    assert not torch.isnan(grad).any()                                    # every gradient element written
    gmax = grad_ref.abs().max().item() + 1e-12
    err = (grad.double() - grad_ref).abs().max().item()
    assert err <= grad_tol * gmax, (err, gmax)
    for n in range(lp.shape[0]):
        assert (grad[n, int(il[n]):] == 0).all()
    return nll, plan


@pytest.mark.parametrize("name", ["ragged", "infeasible", "single"])
@pytest.mark.parametrize("serial", [False, True])
def test_ctc_source_on_reference_fixtures(ctc, golden, name, serial):
    """tests/golden/ctc.npz: losses and gradients frozen from the reference's nn.CTCLoss(blank=0, 'mean', zero_infinity=True)"""
    g = golden("ctc")
    nll, _ = check_ctc(ctc, g[name + ":lp"], g[name + ":tg"], g[name + ":il"], g[name + ":tl"], serial=serial)
    np.testing.assert_allclose(nll.numpy(), g[name + ":nll"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("N,T,S,C,want_threads", [(3, 70, 12, 29, (2, 32)), (2, 150, 100, 29, (2, 128)), (2, 90, 40, 5, (2, 64)),
                                                  (2, 40, 1, 29, (2, 32)), (1, 170, 150, 29, (4, 96)), (1, 340, 300, 29, (8, 96))])
@pytest.mark.parametrize("from_logits", [False, True])
def test_ctc_source_random(ctc, N, T, S, C, want_threads, from_logits):
    """ragged lengths, repeated labels, an empty target; lattices of one to four warps (the wavefront across warps) with 2, 4
    and 8 states per lane"""
    g = torch.Generator().manual_seed(T + S)
    x = torch.randn(N, T, C, generator=g) * 1.5
    lp = x if from_logits else torch.log_softmax(x, -1)
    tg = torch.randint(1, C, (N, S), generator=g, dtype=torch.int32)
    tg[:, 1::3] = tg[:, 0::3][:, : tg[:, 1::3].shape[1]]
    il = torch.randint(max(1, T // 2), T + 1, (N,), generator=g, dtype=torch.int32)
    tl = torch.randint(0, S + 1, (N,), generator=g, dtype=torch.int32)
    il[0], tl[0] = T, S
    if N > 1:
        tl[1] = 0
    for n in range(N):
        tg[n, tl[n]:] = 0
    _, plan = check_ctc(ctc, lp, tg, il, tl, from_logits=from_logits)
    assert (plan["R"], plan["threads"]) == want_threads and plan["parallel"]
    if not from_logits:
        check_ctc(ctc, lp, tg, il, tl, serial=True)


def test_ctc_source_zero_length_input_and_renorm(ctc):
    """an utterance with input length 0 (grad exactly 0, feasible only for an empty target) and peaked frames long enough to
    cross several re-centring blocks"""
    g = torch.Generator().manual_seed(2)
    N, T, S, C = 3, 130, 9, 7
    lp = torch.log_softmax(torch.randn(N, T, C, generator=g) * 6, -1)
    tg = torch.randint(1, C, (N, S), generator=g, dtype=torch.int32)
    il = torch.tensor([130, 0, 97], dtype=torch.int32)
    tl = torch.tensor([9, 0, 4], dtype=torch.int32)
    tg[1] = 0
    tg[2, 4:] = 0
    for serial in (False, True):
        loss, nll, grad, _ = emu_ctc_loss(ctc, lp, tg, il, tl, serial=serial)
        assert nll[1].item() == 0.0 and (grad[1] == 0).all()
        ref = torch.nn.functional.ctc_loss(lp[[0, 2]].double().transpose(0, 1), tg[[0, 2]], il[[0, 2]], tl[[0, 2]], reduction="none")
        np.testing.assert_allclose(nll[[0, 2]].numpy(), ref.numpy(), rtol=1e-4)


# ------------------------------------------------------------------------------------------------ greedy decode
@pytest.fixture(scope="module")
def dec():
    return KE.build(["decode.cu"], ["greedy_argmax_kernel", "greedy_compact_kernel"])


def emu_greedy_decode(dec, scores, sizes=None, blank=0):
    """csrc/decode.cu w2l_greedy_decode, launch for launch; scores may be any [N, T, C] view with unit class stride"""
    N, T, C = scores.shape
    assert scores.stride(2) == 1
    chunk = 256
    nchunks = max(1, (T + chunk - 1) // chunk)
    am = torch.full((N, T), -7, dtype=torch.int32)
    tok, off = torch.full((N, T), -7, dtype=torch.int32), torch.full((N, T), -7, dtype=torch.int32)
    cnt, cc = torch.full((N,), -7, dtype=torch.int32), torch.full((N * nchunks,), -7, dtype=torch.int32)
    sz = None if sizes is None else torch.as_tensor(sizes, dtype=torch.int32)
    szp = None if sz is None else sz.data_ptr()
    dec.launch("greedy_argmax_kernel", (nchunks, N), chunk, scores.data_ptr(), T, C, scores.stride(0), scores.stride(1), szp, blank,
               am.data_ptr(), cc.data_ptr(), nchunks, smem=(4 + chunk * C) * 4)
    dec.launch("greedy_compact_kernel", (nchunks, N), chunk, am.data_ptr(), T, szp, blank, cc.data_ptr(), nchunks, tok.data_ptr(),
               off.data_ptr(), cnt.data_ptr())
    return am, tok, off, cnt


def check_decode(dec, lp, sizes):
    am, tok, off, cnt = emu_greedy_decode(dec, lp, sizes)
    want_am = O.greedy_argmax(lp.numpy())
    assert np.array_equal(am.numpy(), want_am)
    toks, offs = O.greedy_collapse(want_am, sizes)
    for n in range(lp.shape[0]):
        k = int(cnt[n])
        assert tok[n, :k].tolist() == toks[n] and off[n, :k].tolist() == offs[n]           # bit-exact transcripts and offsets
        assert (tok[n, k:] == -1).all() and (off[n, k:] == -1).all()                       # deterministic tail


def test_decode_source_reference_vectors(dec):
    """the reference's own decoder vectors: unit_tests/decoder_test.py:40-42 and decoder.py:305-311"""
    labels = ["_", "A", "B", " "]
    probs = torch.tensor([[[0.8, 0.2, 0, 0], [0.6, 0.4, 0, 0]]])
    _, tok, _, cnt = emu_greedy_decode(dec, probs)
    assert "".join(labels[i] for i in tok[0, :int(cnt[0])].tolist()) == ""
    probs = torch.tensor([[[0.1, 0.7, 0.1, 0.1], [0.7, 0.1, 0.1, 0.1], [0.1, 0.1, 0.7, 0.1], [0.1, 0.7, 0.1, 0.1]],
                          [[0.1, 0.1, 0.1, 0.7], [0.1, 0.1, 0.1, 0.7], [0.7, 0.1, 0.1, 0.1], [0.7, 0.1, 0.1, 0.1]]])
    am, tok, off, cnt = emu_greedy_decode(dec, probs)
    lab = ["_", "a", "b", " "]
    assert ["".join(lab[i] for i in tok[n, :int(cnt[n])].tolist()) for n in range(2)] == ["aba", " "]
    assert off[0, :3].tolist() == [0, 2, 3] and off[1, :1].tolist() == [0]
    # ties -> lowest index, NaN wins the argmax (torch.max semantics)
    probs = torch.tensor([[[0.5, 0.5, 0, 0], [0.1, float("nan"), 0.9, 0.0], [0.2, 0.3, 0.3, 0.2]]])
    am, _, _, _ = emu_greedy_decode(dec, probs)
    assert am[0].tolist() == [0, 1, 1]


@pytest.mark.parametrize("N,T,C", [(3, 50, 29), (2, 256, 29), (2, 257, 5), (2, 700, 29), (1, 1, 2), (4, 300, 31)])
def test_decode_source_random(dec, N, T, C):
    g = torch.Generator().manual_seed(N * 1000 + T)
    lp = torch.log_softmax(torch.randn(N, T, C, generator=g) * 2, -1)
    lp[:, :, 0] += 1.0
    lp = torch.round(lp * 4) / 4                                           # plenty of exact ties and repeats
    sizes = torch.randint(0, T + 1, (N,), generator=g).tolist()
    sizes[0] = T
    check_decode(dec, lp, sizes)
    check_decode(dec, lp, None)
    v = lp.transpose(0, 1).contiguous().transpose(0, 1)                    # [N,T,C] view of a [T,N,C] tensor: the strided path
    am, _, _, _ = emu_greedy_decode(dec, v)
    assert np.array_equal(am.numpy(), O.greedy_argmax(lp.numpy()))
    w = torch.zeros(N * T * C + 3)[3:].view(N, T, C).copy_(lp)             # base 12 bytes off 16: the peeled, vectorised copy
    check_decode(dec, w, sizes)
