"""Fused NovoGrad, device-side WER/CER and the log-mel front-end (csrc/novograd.cu, metrics.cu,
features.cu) executed on the host from their source text (tests/_emu_backend.py) against the fixtures frozen from the
reference (tests/golden/novograd.npz, features.npz) and the reference's host arithmetic."""
import random

import numpy as np
import pytest
import torch

import _emu_backend as E
import _kernel_emu as KE
from oracle import ref_loader as rl
from oracle import w2l_oracle as O

pytestmark = pytest.mark.skipif(not KE.available(), reason="needs g++ and the CUDA headers")


@pytest.mark.parametrize("amsgrad", [False, True])
def test_novograd_source_on_reference_fixture(golden, amsgrad):
    """novograd.py:52-114 through three steps (second moment per tensor, weight decay, momentum), bf16 shadow refreshed"""
    g = golden("novograd")
    p = [torch.from_numpy(g["p0:0"]).clone(), torch.from_numpy(g["p0:1"]).clone()]
    m = [torch.zeros_like(q) for q in p]
    shadow = [torch.zeros(q.shape, dtype=torch.bfloat16) for q in p]
    v, vmax = torch.zeros(2), (torch.zeros(2) if amsgrad else None)
    for step in range(3):
        grads = [torch.from_numpy(g["g%d:%d" % (step, i)]).clone() * ((0.5 ** step) if amsgrad else 1.0) for i in range(2)]
        E.novograd_step(p, grads, m, v, vmax, shadow, lr=0.01, beta1=0.95, beta2=0.5, eps=1e-8, weight_decay=1e-3, grad_averaging=False)
        for i in range(2):
            np.testing.assert_allclose(p[i].numpy(), g[("ams_p%d:%d" if amsgrad else "p%d:%d") % (step + 1, i)], rtol=1e-5, atol=1e-6)
            assert torch.equal(shadow[i], p[i].to(torch.bfloat16))
    if amsgrad:
        assert (vmax >= v).all()


def test_novograd_source_many_tensors_and_chunks():
    """tensors from 1 element to several 16384-element chunks, odd sizes (vector body + scalar tail), grad_averaging, vs the oracle"""
    gen = torch.Generator().manual_seed(0)
    sizes = [1, 7, 16384, 16385, 40001, 3 * 16384, 29]
    p = [torch.randn(n, generator=gen) for n in sizes]
    p_ref = [q.clone() for q in p]
    m = [torch.zeros_like(q) for q in p]
    v = torch.zeros(len(p))
    state = [{} for _ in p]
    for step in range(2):
        grads = [torch.randn(n, generator=gen) * (0.1 + i) for i, n in enumerate(sizes)]
        E.novograd_step(p, grads, m, v, None, None, lr=0.02, beta1=0.9, beta2=0.25, eps=1e-8, weight_decay=1e-2, grad_averaging=True)
        O.novograd_step(p_ref, [x.clone() for x in grads], state, lr=0.02, betas=(0.9, 0.25), eps=1e-8, weight_decay=1e-2, grad_averaging=True)
        for a, b in zip(p, p_ref):
            np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=2e-5, atol=1e-6)


def _lev(a, b):
    return rl._levenshtein(a, b)


def test_string_metrics_source_vs_reference_arithmetic():
    """decoder.py:31-66 + base_asr_models.py:58-69: CER on the strings without spaces, WER on the word lists, len ratio"""
    labels = O.ENGLISH_LOWERCASE
    space = labels.index(" ")
    rnd = random.Random(5)
    N, T = 9, 120
    hyps = ["".join(rnd.choice(labels[1:]) for _ in range(rnd.randint(0, 100))) for _ in range(N)]
    hyps[3] = ""                                                 # an empty hypothesis
    hyps[4] = " lead and trail  double "
    texts = ["".join(rnd.choice(labels[1:]) for _ in range(rnd.randint(1, 90))) for _ in range(N)]
    texts[1] = "  " + texts[1] + "  a "
    texts[5] = "word"
    texts[6] = hyps[6] or "same"                                 # an exact match
    hyps[6] = texts[6]
    tokens, counts = torch.full((N, T), -1, dtype=torch.int32), torch.zeros(N, dtype=torch.int32)
    for n, h in enumerate(hyps):
        ids = [labels.index(c) for c in h]
        tokens[n, :len(ids)] = torch.tensor(ids, dtype=torch.int32) if ids else tokens[n, :0]
        counts[n] = len(ids)
    S = max(len(t) for t in texts)
    ref_ids, ref_lens = torch.zeros((N, S), dtype=torch.int32), torch.zeros(N, dtype=torch.int32)
    for n, t in enumerate(texts):
        ref_ids[n, :len(t)] = torch.tensor([labels.index(c) for c in t], dtype=torch.int32)
        ref_lens[n] = len(t)
    cer_den = sum(len(t.replace(" ", "")) for t in texts)
    wer_den = sum(len(t.split()) for t in texts)
    len_den = sum(map(len, texts))
    ratios, cer_d, wer_d = E.string_metrics(tokens, counts, space, ref_ids, ref_lens, cer_den, wer_den, len_den)
    want_c = [_lev(t.replace(" ", ""), h.replace(" ", "")) for t, h in zip(texts, hyps)]
    want_w = []
    for t, h in zip(texts, hyps):                                # decoder.py:31-48: words -> symbols, then Levenshtein
        vocab = {w: i for i, w in enumerate(set(t.split() + h.split()))}
        want_w.append(_lev([vocab[w] for w in t.split()], [vocab[w] for w in h.split()]))
    assert cer_d.tolist() == want_c and wer_d.tolist() == want_w
    np.testing.assert_allclose(ratios.numpy(), [sum(want_c) / cer_den, sum(want_w) / wer_den, sum(map(len, hyps)) / len_den], rtol=1e-6)


def test_features_source_on_reference_fixture_and_ragged_batch(golden):
    """tests/golden/features.npz (the reference's SpectrogramExtractor.extract, same dither noise) and a ragged batch vs the oracle"""
    g = golden("features")
    win, hop, n_fft = 320, 160, 512
    window = torch.hamming_window(win, periodic=False)
    fb = torch.from_numpy(g["fb"])
    for name in ("a", "b"):
        sig, noise = torch.from_numpy(g[name + ":signal"]).float(), torch.from_numpy(g[name + ":noise"]).float()
        out, nfr = E.logmel_features(sig[None], torch.tensor([sig.numel()]), noise[None], window, fb, n_fft, win, hop)
        want = g[name + ":feats"]
        assert out[0].shape == want.shape and int(nfr[0]) == want.shape[1]
        assert np.abs(out[0].numpy() - want).max() < 2e-4
    rs = np.random.RandomState(3)
    sigs = [(0.2 * rs.randn(n)).astype(np.float32) for n in (4000, 1611, 3200, 400)]
    lens = torch.tensor([len(s) for s in sigs], dtype=torch.int32)
    audio = torch.zeros(len(sigs), int(lens.max()))
    for i, s in enumerate(sigs):
        audio[i, :len(s)] = torch.from_numpy(s)
    noise = torch.randn(len(sigs), int(lens.max()), generator=torch.Generator().manual_seed(5))
    out, nfr = E.logmel_features(audio, lens, noise, window, fb, n_fft, win, hop)
    want, want_lens = O.collate_features([O.spectrogram_extract(s, noise=noise[i, :len(s)].numpy()) for i, s in enumerate(sigs)])
    assert nfr.tolist() == want_lens.tolist() == [26, 11, 21, 3]
    assert out.shape == want.shape and (out - want).abs().max() < 2e-4
    for i, n in enumerate(nfr.tolist()):
        assert (out[i, :, n:] == 0).all()
    out2, _ = E.logmel_features(audio, lens, None, window, fb, n_fft, win, hop)       # no dither: same features to within its effect
    assert (out2 - out).abs().max() < 0.2
