"""ConvCTCASR -- the model base class with the reference's surface (base_asr_models.py:16-94): it owns the CTC
criterion and the decoder, and provides training_step / validation_step / configure_optimizers /
add_string_metrics / compute_output_lengths.  pytorch_lightning is optional: when importable the class derives from
``LightningModule`` (so ``Trainer.fit`` works as in train.py:34-37), otherwise from ``nn.Module`` with the two
Lightning hooks the steps use (``log_dict``, ``optimizers``)."""
import random

import torch
import torch.nn as nn

from .config import instantiate
from .ctc_loss import CTCLoss

try:  # pragma: no cover - not installed in the build image
    import pytorch_lightning as _ptl
    _Base = _ptl.LightningModule
except Exception:  # noqa: BLE001
    class _Base(nn.Module):
        def log_dict(self, logs, *args, **kwargs):
            # detached, as Lightning's logger stores them: a kept loss would keep the step's autograd graph (and its AccumulateGrad
            # nodes, bound to the stream of the step that made them) alive into the next step
            self.logged = {k: v.detach() if torch.is_tensor(v) else v for k, v in logs.items()}

        def optimizers(self):
            return self._optimizer


class ConvCTCASR(_Base):
    def __init__(self, cfg):
        super().__init__()
        self._cfg = cfg
        self.audio_conf = cfg.audio_conf
        self.labels = cfg.labels
        self.ctc_decoder = instantiate(cfg.decoder)
        # base_asr_models.py:23 -- same constructor arguments, CUDA kernel behind it
        self.criterion = CTCLoss(blank=0, reduction="mean", zero_infinity=True)
        self.print_decoded_prob = cfg.get("print_decoded_prob", 0)
        self.example_input_array = self.create_example_input_array()

    def create_example_input_array(self):
        """(features [4, input_size, 200], lengths [4]) -- what Lightning uses for its model summary."""
        n, lo, hi = 4, 100, 200
        lengths = torch.randint(lo, hi, (n,))            # drawn first, like the reference, so seeded inits line up
        return torch.rand(n, self._cfg.input_size, hi), lengths

    # ---- length bookkeeping: subclasses define scaling_factor (product of the strides)
    @property
    def scaling_factor(self):
        raise NotImplementedError()

    def compute_output_lengths(self, input_lengths):
        """Floor division by the total stride; dtype/device follow ``input_lengths`` (int32 from the collator)."""
        return input_lengths // self.scaling_factor

    def forward(self, inputs, input_lengths):
        """-> (scores [B, T', n_labels], output_lengths [B])"""
        raise NotImplementedError()

    # ---- metrics: greedy transcripts (CUDA argmax+collapse) scored on the host against the reference texts
    def add_string_metrics(self, out, output_lengths, texts, prefix):
        dec = self.ctc_decoder
        show = random.random() < self.print_decoded_prob
        if not show and out.is_cuda and hasattr(dec, "error_ratios_device"):
            # decode + CER/WER entirely on the device: the logged values are 0-dim CUDA tensors (log_dict accepts them)
            # and the step never waits for the GPU, unlike the reference's per-frame .item() loop + host Levenshtein
            ratios = dec.error_ratios_device(out, output_lengths, texts)
            if ratios is not None:
                return {prefix + "_cer": ratios[0], prefix + "_wer": ratios[1], prefix + "_len_ratio": ratios[2]}
        hyps = dec.decode(out, output_lengths)
        if show:
            print("reference: %s\ndecoded  : %s" % (texts[0], hyps[0]))
        if hasattr(dec, "error_ratio_sums"):          # one batched, threaded host call instead of 2*B scalar ones
            cer_num, cer_den, wer_num, wer_den = dec.error_ratio_sums(texts, hyps)
        else:
            cer_pairs = [dec.cer_ratio(ref, hyp) for ref, hyp in zip(texts, hyps)]
            wer_pairs = [dec.wer_ratio(ref, hyp) for ref, hyp in zip(texts, hyps)]
            cer_num, cer_den = sum(p[0] for p in cer_pairs), sum(p[1] for p in cer_pairs)
            wer_num, wer_den = sum(p[0] for p in wer_pairs), sum(p[1] for p in wer_pairs)
        # ZeroDivisionError on empty references, as upstream
        return {prefix + "_cer": cer_num / cer_den, prefix + "_wer": wer_num / wer_den,
                prefix + "_len_ratio": sum(len(h) for h in hyps) / sum(len(t) for t in texts)}

    # ---- the hooks a Lightning Trainer (or bench.py's plain loop) drives
    def configure_optimizers(self):
        self._optimizer = instantiate(self._cfg.optimizer, params=self.parameters())
        if hasattr(self._optimizer, "attach_model"):          # fused NovoGrad also refreshes the bf16 weight shadows
            self._optimizer.attach_model(self)
        return [self._optimizer], [instantiate(self._cfg.scheduler, optimizer=self._optimizer)]

    def _step(self, batch, prefix):
        inputs, input_lengths, targets, target_lengths, _paths, texts = batch
        scores, out_lens = self.forward(inputs, input_lengths)
        # the criterion takes [T, N, C]; this is a strided view, the CTC kernel reads it in place
        loss = self.criterion(scores.transpose(0, 1), targets, out_lens, target_lengths)
        # decode + WER/CER feed only the logger.  They run inline, behind the CTC kernels: on a side stream BESIDE them (rounds 1-2) the
        # barrier-heavy edit-distance CTAs shared SMs with the latency-bound lattice recursion and cost it 0.2 ms for the 0.13 ms they
        # take (profiles/r2_graph_step.md, call 32).  graph_step.GraphedTrainStep forks them as a branch that overlaps the backward pass.
        return loss, self.add_string_metrics(scores, out_lens, texts, prefix)

    def training_step(self, batch, batch_idx):
        loss, metrics = self._step(batch, "train")
        self.log_dict({"train_loss": loss, "learning_rate": self.optimizers().param_groups[0]["lr"], **metrics})
        return loss

    def validation_step(self, batch, batch_idx):
        loss, metrics = self._step(batch, "val")
        self.log_dict({"val_loss": loss, **metrics})
        return loss
