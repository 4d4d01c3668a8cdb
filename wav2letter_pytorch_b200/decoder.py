"""Decoders with the reference's interface (decoder.py:11-145): ``Decoder(labels, blank_index)``, ``decode(probs,
sizes=None[, return_offsets])``, ``wer`` / ``cer`` / ``wer_ratio`` / ``cer_ratio``.

``GreedyDecoder`` runs argmax + collapse as one coalesced CUDA pass (csrc/decode.cu) instead of the reference's
per-frame Python loop with two ``.item()`` device syncs per frame; a single device->host copy of the compacted
tokens then builds the strings."""
import ctypes

import numpy as np
import torch

from . import _lib
from . import functional as F
from . import label_sets


def _as_ids(seq, vocab):
    """characters -> code points (zero Python loop); words -> dense ids through ``vocab``."""
    if isinstance(seq, str):
        return np.frombuffer(seq.encode("utf-32-le"), dtype=np.int32)
    return np.fromiter((vocab.setdefault(s, len(vocab)) for s in seq), dtype=np.int32, count=len(seq))


def edit_distances(pairs):
    """Levenshtein distance of every (a, b) in ``pairs`` (strings, or sequences of hashable symbols such as word
    lists) in ONE call of the library's batched bit-parallel host routine (the reference calls the python-Levenshtein
    C extension once per pair: decoder.py:31-60, base_asr_models.py:58-69)."""
    n = len(pairs)
    if n == 0:
        return []
    vocab = {}
    a = [_as_ids(p[0], vocab) for p in pairs]
    b = [_as_ids(p[1], vocab) for p in pairs]
    a_off = np.zeros(n + 1, dtype=np.int64)
    b_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(x) for x in a], out=a_off[1:])
    np.cumsum([len(x) for x in b], out=b_off[1:])
    a_cat = np.ascontiguousarray(np.concatenate(a)) if a_off[-1] else np.zeros(1, dtype=np.int32)
    b_cat = np.ascontiguousarray(np.concatenate(b)) if b_off[-1] else np.zeros(1, dtype=np.int32)
    out = np.zeros(n, dtype=np.int64)
    P = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    rc = _lib.load().w2l_edit_distance_batch_host(P(a_cat), P(a_off), P(b_cat), P(b_off), n, P(out), 0)
    if rc != 0:
        raise RuntimeError("w2l_edit_distance_batch_host failed")
    return out.tolist()


def _edit_distance(a, b):
    return edit_distances([(a, b)])[0]


class Decoder(object):
    def __init__(self, labels, blank_index=0):
        # decoder.py:22-29; NB the reference builds int_to_char from the *argument* even when it is a label-set
        # name -- callers always pass the list (config.yaml:16), and so must users of this class.
        self.labels = label_sets.labels_map[labels] if type(labels) is str else labels
        self.int_to_char = dict(enumerate(self.labels))
        self.blank_index = blank_index
        self.space_index = self.labels.index(" ") if " " in self.labels else len(self.labels)

    def wer(self, s1, s2):
        return _edit_distance(s1.split(), s2.split())

    def cer(self, s1, s2):
        return _edit_distance(s1.replace(" ", ""), s2.replace(" ", ""))

    def cer_ratio(self, expected, predicted):
        return self.cer(expected, predicted), len(expected.replace(" ", ""))

    def wer_ratio(self, expected, predicted):
        return self.wer(expected, predicted), len(expected.split())

    def error_ratio_sums(self, expected, predicted):
        """(sum cer, sum cer denominators, sum wer, sum wer denominators) over a batch -- the quantities
        base_asr_models.py:58-67 accumulates pair by pair -- with two batched library calls."""
        cer = edit_distances([(e.replace(" ", ""), p.replace(" ", "")) for e, p in zip(expected, predicted)])
        wer = edit_distances([(e.split(), p.split()) for e, p in zip(expected, predicted)])
        return (sum(cer), sum(len(e.replace(" ", "")) for e in expected), sum(wer), sum(len(e.split()) for e in expected))

    def decode(self, probs, sizes=None):
        raise NotImplementedError


class GreedyDecoder(Decoder):
    def decode_tokens(self, probs, sizes=None):
        """Device-side part: returns (tokens [N,T], offsets [N,T], counts [N]) int32 CUDA tensors."""
        if not torch.is_tensor(probs):
            probs = torch.as_tensor(probs, dtype=torch.float32)
        if probs.dim() == 2:
            probs = probs.unsqueeze(0)
        if not probs.is_cuda:
            probs = F.to_cuda(probs, "GreedyDecoder")              # host scores are uploaded; without a CUDA device this raises
        if sizes is not None and not torch.is_tensor(sizes):
            sizes = torch.as_tensor([int(s) for s in sizes], dtype=torch.int32)
        if sizes is not None:
            sizes = sizes.to(probs.device)
        _, tokens, offsets, counts = F.greedy_decode(probs.detach(), sizes, self.blank_index)
        return tokens, offsets, counts

    # ---- device-side scoring (no host sync): CER / WER / length ratio as a CUDA tensor
    def _encode_refs(self, texts, into=None):
        """texts -> (ids [N,S] int32 pinned, lens [N] int32 pinned, cer_den, wer_den, len_den) or None when the device path
        does not apply (multi-character labels, exotic whitespace, references longer than 1023).  ``into`` = (ids, lens): caller-owned
        pinned tensors to encode into instead of the rotating internal ones (graph_step.py keeps one fixed pair per captured step)."""
        C = len(self.labels)
        if getattr(self, "_lut", None) is None:
            if any(len(l) != 1 for l in self.labels) or " " not in self.labels:
                self._lut = False
            else:
                lut = np.arange(0x10000, dtype=np.int64) + C
                for i, ch in enumerate(self.labels):
                    if ord(ch) < 0x10000:
                        lut[ord(ch)] = i
                self._lut = lut.astype(np.int32)
                ws = np.zeros(0x10000, dtype=bool)               # str.isspace() code points other than ' ' (all in the BMP)
                for c in range(0x10000):
                    if c != 32 and chr(c).isspace():
                        ws[c] = True
                self._other_ws = ws
                self._pins = []
        if self._lut is False:
            return None
        n = len(texts)
        lens_l = [len(t) for t in texts]
        smax = max(1, max(lens_l))
        if smax > 1023:
            return None
        # the whole batch in ONE pass over its code points (a Python loop per transcript -- split twice, encode, look up -- was 0.8 ms
        # per step for 64 x 225 characters): the transcripts joined by single spaces, so that no word runs across two of them
        cp = np.frombuffer(" ".join(texts).encode("utf-32-le"), dtype=np.int32)
        low = np.minimum(cp, 0xFFFF)
        if self._other_ws[low][cp < 0x10000].any():             # whitespace other than ' ' (str.split() would cut there): host path
            return None
        is_sp = cp == 32
        words = int(np.count_nonzero(~is_sp[1:] & is_sp[:-1]) + (0 if cp.size == 0 or is_sp[0] else 1))
        slot = None
        if into is not None:
            if into[0].shape[0] < n or into[0].shape[1] < smax:
                return None
            slot = [into[0], into[1], None]
        for s in self._pins if slot is None else ():            # rotate pinned staging buffers (CPU may run ahead of the GPU)
            if s[0].shape[0] >= n and s[0].shape[1] >= smax and (s[2] is None or s[2].query()):
                slot = s
                break
        if slot is None:
            slot = [torch.zeros((n, max(smax, 256)), dtype=torch.int32).pin_memory(), torch.zeros((n,), dtype=torch.int32).pin_memory(), None]
            self._pins.append(slot)
        ids, lens = slot[0].numpy(), slot[1].numpy()
        lens_a = np.asarray(lens_l, dtype=np.int64)
        lens[:n] = lens_a
        total = int(lens_a.sum())
        if total:
            starts = np.cumsum(lens_a) - lens_a                  # offset of transcript i among the characters proper ...
            rows = np.repeat(np.arange(n), lens_a)
            cols = np.arange(total) - np.repeat(starts, lens_a)
            src = np.arange(total) + rows                        # ... and in the joined array (i separators precede it)
            c = cp[src]
            ids[rows, cols] = np.where(c < 0x10000, self._lut[np.minimum(c, 0xFFFF)], c + C)
        n_spaces = int(np.count_nonzero(is_sp)) - (n - 1 if n > 0 else 0)
        return slot, smax, total - n_spaces, words, total

    def error_ratios_device(self, probs, sizes, texts):
        """Greedy-decodes ``probs`` and scores the transcripts against ``texts`` entirely on the device.  Returns a CUDA
        fp32 tensor [3] = (cer, wer, len_ratio) -- the three numbers ConvCTCASR.add_string_metrics logs -- or None when the
        device path does not apply.  No device->host synchronisation."""
        enc = self._encode_refs(texts)
        if enc is None:
            return None
        slot, smax, cer_den, wer_den, len_den = enc
        if cer_den == 0 or wer_den == 0 or len_den == 0:
            raise ZeroDivisionError("division by zero")            # what the reference's host arithmetic raises
        dev = probs.device if torch.is_tensor(probs) and probs.is_cuda else torch.device("cuda", torch.cuda.current_device())
        ref_ids = slot[0].to(dev, non_blocking=True)
        ref_lens = slot[1].to(dev, non_blocking=True)
        slot[2] = torch.cuda.Event()
        slot[2].record()
        return self.score_device(probs, sizes, ref_ids, ref_lens, (cer_den, wer_den, len_den))

    def score_device(self, probs, sizes, ref_ids, ref_lens, dens=(1.0, 1.0, 1.0)):
        """Device part of ``error_ratios_device`` over references that are already encoded and resident (ids [N,S] / lens [N] int32
        CUDA): greedy decode + edit distances, returns fp32 [3] = (character errors, word errors, hypothesis characters) each divided
        by its entry of ``dens``.  With the default denominators these are the raw sums, which is what a captured step needs -- the
        denominators change with every batch's texts and would be frozen into the graph as by-value arguments."""
        tokens, _offsets, counts = self.decode_tokens(probs, sizes)
        dev = tokens.device
        N, T = tokens.shape
        S = ref_ids.shape[1]
        lib = _lib.load()
        ws = torch.empty((lib.w2l_string_metrics_workspace_bytes(N, T, S) + 7) // 8, dtype=torch.int64, device=dev)
        ratios = torch.empty(3, dtype=torch.float32, device=dev)
        with F._on(dev):
            _lib.check(lib.w2l_string_metrics(F._ptr(tokens), F._ptr(counts), N, T, self.space_index, F._ptr(ref_ids), F._ptr(ref_lens), S,
                                              float(dens[0]), float(dens[1]), float(dens[2]), F._ptr(ratios), F._ptr(ws), ws.numel() * 8,
                                              F._stream()), "string_metrics")
        return ratios

    def decode(self, probs, sizes=None, return_offsets=False):
        """probs [N,T,C] (or [T,C]) -> list[str] (and list[[IntTensor]] of frame offsets), decoder.py:121-145."""
        tokens, offsets, counts = self.decode_tokens(probs, sizes)
        N, T = tokens.shape
        counts_h = counts.cpu()
        width = int(counts_h.max().item()) if N > 0 else 0
        packed = torch.stack([tokens[:, :width], offsets[:, :width]]).cpu() if width > 0 else None    # one D2H copy
        strings, offs = [], []
        for n in range(N):
            c = int(counts_h[n])
            ids = packed[0, n, :c].tolist() if c else []
            if ids and self.space_index >= len(self.labels):
                raise IndexError("list index out of range")          # decoder.py:113 with no ' ' in labels
            strings.append("".join(self.int_to_char[i] for i in ids))
            if return_offsets:
                offs.append([packed[1, n, :c].to(torch.int32) if c else torch.IntTensor([])])
        if return_offsets:
            return strings, offs
        return strings


# ------------------------------------------------------------------------------------------------ prefix beam search
_LM_CB = ctypes.CFUNCTYPE(ctypes.c_double, ctypes.POINTER(ctypes.c_int32), ctypes.c_int64, ctypes.c_void_p)


def _beam_search_batch(probs, frames, labels, blank_index, lm, k, alpha, beta, prune, end_char, threads=0):
    """probs [N, T, F] float64 numpy (C-contiguous), frames [N] or None -> (list[str], list[float])."""
    import re
    if any(len(l) != 1 for l in labels):
        raise NotImplementedError("prefix beam search: multi-character labels are not supported")
    N, T, F_ = probs.shape
    word = np.array([1 if re.fullmatch(r"\w", c) else 0 for c in labels], dtype=np.uint8)
    term = np.array([1 if re.fullmatch(r"[\s|>]", c) else 0 for c in labels], dtype=np.uint8)
    space_id = labels.index(" ") if " " in labels else -1
    end_id = labels.index(end_char) if end_char in labels else -1
    out_ids = np.zeros((N, max(T, 1)), dtype=np.int32)
    out_len = np.zeros(N, dtype=np.int64)
    out_score = np.zeros(N, dtype=np.float64)
    fr = None if frames is None else np.ascontiguousarray(frames, dtype=np.int64)
    cb = ctypes.cast(None, _LM_CB)
    if lm is not None:
        def _call(ids, n, _user):
            return float(lm("".join(labels[ids[i]] for i in range(n))))
        cb = _LM_CB(_call)
    P = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    rc = _lib.load().w2l_prefix_beam_search_host(P(probs), P(fr), N, T, F_, blank_index, space_id, end_id, P(word), P(term), int(k),
                                                float(alpha), float(beta), float(prune), ctypes.cast(cb, ctypes.c_void_p), None,
                                                P(out_ids), out_ids.shape[1], P(out_len), P(out_score), threads)
    if rc != 0:
        raise RuntimeError("w2l_prefix_beam_search_host failed (code %d)" % rc)
    strings = ["".join(labels[i] for i in out_ids[n, :out_len[n]]) for n in range(N)]
    return strings, out_score.tolist()


def prefix_beam_search(ctc, labels, blank_index=0, lm=None, k=5, alpha=0.3, beta=5, prune=0.001, end_char=">", return_weights=False):
    """Drop-in for the reference function (decoder.py:147-231): ``ctc`` [T, F] probabilities -> best prefix (and its score).
    Same float64 arithmetic, candidate order and tie-breaking, run by the library's host routine."""
    if torch.is_tensor(ctc):
        ctc = ctc.detach().cpu().numpy()
    ctc = np.asarray(ctc)
    assert (ctc.shape[1] == len(labels)), "ctc size:%d, labels: %d" % (ctc.shape[1], len(labels))
    assert ctc.shape[0] > 1, "ctc length: %d was too short" % ctc.shape[0]
    assert (ctc >= 0).all(), "ctc output contains negative numbers"
    probs = np.ascontiguousarray(ctc, dtype=np.float64)[None]
    strings, scores = _beam_search_batch(probs, None, list(labels), blank_index, lm, k, alpha, beta, prune, end_char)
    if return_weights:
        return strings[0], scores[0]
    return strings[0]


class PrefixBeamSearchLMDecoder(Decoder):
    """decoder.py:233-267: same constructor and ``decode``; a batch is decoded by ONE library call spread over host threads
    (the reference loops over utterances in Python)."""

    def __init__(self, lm_path, labels, blank_index=0, k=5, alpha=0.3, beta=5, prune=1e-3):
        super(PrefixBeamSearchLMDecoder, self).__init__(labels, blank_index)
        if lm_path:
            import kenlm                                   # same optional dependency as the reference
            self.lm = kenlm.Model(lm_path)
            self.lm_weigh = lambda f: 10 ** (self.lm.score(f))
        else:
            self.lm_weigh = None                           # the reference's `lambda s: 1`: no callback needed
        self.k, self.alpha, self.beta, self.prune = k, alpha, beta, prune

    def decode(self, probs, sizes=None, return_offsets=False):
        if return_offsets:
            raise NotImplementedError("Prefix beam search does not support offsets (yet).")
        if torch.is_tensor(probs):
            probs = probs.detach().float().cpu().numpy()
        probs = np.asarray(probs)
        if len(probs.shape) == 2:
            return prefix_beam_search(probs, self.labels, self.blank_index, self.lm_weigh, self.k, self.alpha, self.beta, self.prune)
        if len(probs.shape) == 3:
            assert probs.shape[2] == len(self.labels), "ctc size:%d, labels: %d" % (probs.shape[2], len(self.labels))
            assert probs.shape[1] > 1, "ctc length: %d was too short" % probs.shape[1]
            assert (probs >= 0).all(), "ctc output contains negative numbers"
            # NB the reference ignores `sizes` here (every utterance is decoded over all T frames); so does this class
            strings, _ = _beam_search_batch(np.ascontiguousarray(probs, dtype=np.float64), None, list(self.labels), self.blank_index,
                                            self.lm_weigh, self.k, self.alpha, self.beta, self.prune, ">")
            return strings
        raise RuntimeError("Decoding with wrong shape: %s, expected either [Batch X Frames X Labels] or [Frames X Labels]" % str(probs.shape))


def get_time_per_word(predictions, offsets, ratio=1.0):
    """(word, start, end) for every word of a greedy transcript, from the per-character frame offsets that
    ``GreedyDecoder.decode(..., return_offsets=True)`` returns (decoder.py:269-300; ``ratio`` = seconds per output frame).
    As upstream, a word's end is the FIRST frame of its last character."""
    assert len(predictions) == len(offsets)
    words, current, start, end = [], "", -1, -1
    for letter, offset in zip(predictions, offsets):
        if letter == " ":
            if current:
                words.append((current, start, end))
                current, start, end = "", -1, -1
            continue
        t = offset * ratio
        if current:
            current, end = current + letter, t
        else:
            current, start, end = letter, t, t
    if current:
        words.append((current, start, end))
    return words
