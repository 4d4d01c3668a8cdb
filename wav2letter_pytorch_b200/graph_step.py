"""GraphedTrainStep -- one whole training step of a ConvCTCASR model (forward, CTC loss, greedy decode + CER/WER, backward,
optimizer) captured ONCE in a CUDA graph and replayed per batch.

Why: the step of a small configuration -- the reference's literal default, ``mid_layers: 1`` (configuration/model/wav2letter.yaml) --
is a few dozen short kernels; launched one by one from Python the host is the bottleneck (1.7 ms per step for 0.6 ms of kernels).
Replayed as a graph the step costs one launch.  The library is capturable by construction: every entry point only enqueues work on
the stream it is given, workspaces are caller-owned, and nothing reads device memory back.

What the reference does per step (base_asr_models.py:78-85 ``training_step`` + Lightning's backward / optimizer.step) and what
changes under capture:
  * by-value kernel arguments are frozen into the graph.  Batch data therefore lives in static device tensors the caller's batch is
    copied into; the dropout seed is advanced by a device-resident epoch counter bumped inside the graph
    (``functional.set_dropout_epoch``), so every replay draws a fresh mask like nn.Dropout does per call; the metric denominators
    (reference character / word counts) are device operands; optimizer hyper-parameters are by-value, so a changed learning rate
    (scheduler) re-captures the graph -- once per scheduler step, not per batch.
  * shapes are frozen: every batch must have the example batch's shapes (pad to the bucket's maximum as the collator does anyway,
    data/data_loader.py:149-158); true lengths stay dynamic (they are device operands).
  * the ``warmup`` eager steps that precede the capture are real training steps on the example batch.

Single process only: the data-parallel gradient reducers exchange flags with peer GPUs and are not captured."""
import torch

from . import functional as F


_epochs = {}


def _dropout_epoch(dev):
    """The device-resident step counter the library folds into every dropout seed, registered once per process and never freed
    (the library keeps a raw pointer to it; one device per process, like the GEMM scratch): captured steps bump it inside their graph."""
    t = _epochs.get(dev)
    if t is None:
        if _epochs:
            raise RuntimeError("wav2letter_pytorch_b200 runs one process per GPU: the dropout epoch already lives on %s" % next(iter(_epochs)))
        t = _epochs[dev] = torch.zeros(1, dtype=torch.int64, device=dev)
        F.set_dropout_epoch(t)
    return t


class GraphedTrainStep:
    def __init__(self, model, optimizer, batch, warmup=2, max_text_len=None):
        """``batch`` = (inputs [B,F,T] fp32, input_lengths [B], targets [B,S], target_lengths [B], paths, texts) -- the collator's
        tuple (data/data_loader.py:158); its shapes become the graph's.  ``max_text_len``: longest reference transcript any later
        batch may carry (default: the width of ``targets``, which is what the collator pads to)."""
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            raise RuntimeError("GraphedTrainStep: single-process only (the peer gradient reducer is not capturable)")
        inputs, input_lengths, targets, target_lengths, _paths, texts = batch
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep: the model must live on a CUDA device (there is no CPU path)")
        self.model, self.optimizer, self.device = model, optimizer, dev
        self._static = (torch.empty(inputs.shape, dtype=torch.float32, device=dev),
                        torch.empty(input_lengths.shape, dtype=torch.int32, device=dev),
                        torch.empty(targets.shape, dtype=torch.int32, device=dev),
                        torch.empty(target_lengths.shape, dtype=torch.int32, device=dev))
        n = inputs.shape[0]
        s_ref = max(int(max_text_len or targets.shape[1]), max((len(t) for t in texts), default=1), 4)
        s_ref = (s_ref + 3) // 4 * 4
        n4 = (n + 3) // 4 * 4
        # one pinned block / one device block per step for everything the metrics need: [3 denominators + pad | lens | ids]
        self._ref_host = torch.zeros(4 + n4 + n * s_ref, dtype=torch.int32).pin_memory()
        self._ref_dev = torch.zeros(4 + n4 + n * s_ref, dtype=torch.int32, device=dev)

        def views(buf):
            return buf[:4].view(torch.float32)[:3], buf[4:4 + n], buf[4 + n4:].view(n, s_ref)
        self._den_h, self._lens_h, self._ids_h = views(self._ref_host)
        self._den_d, self._lens_d, self._ids_d = views(self._ref_dev)
        self._metric_stream = torch.cuda.Stream(device=dev)
        self._epoch = _dropout_epoch(dev)
        self._metrics_on = self._load(batch)             # False: this decoder / these texts have no device scoring path
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)             # warm up off the default stream, as CUDA graph capture requires
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, int(warmup))):
                optimizer.zero_grad(set_to_none=True)
                loss, _ = self._body()
                loss.backward()
                optimizer.step()
                self._join()
                del loss                                 # no autograd graph of an eager step may outlive it (see _capture)
        cur.wait_stream(side)
        self.graph, self._key = None, None
        self._capture()

    # ---- host side of one step: the caller's batch into the static operands
    def _load(self, batch):
        inputs, input_lengths, targets, target_lengths, _paths, texts = batch
        for dst, src in zip(self._static, (inputs, input_lengths, targets, target_lengths)):
            if tuple(src.shape) != tuple(dst.shape):
                raise ValueError("GraphedTrainStep: batch tensor of shape %s, the captured step takes %s (pad every batch to the "
                                 "example batch's shapes)" % (tuple(src.shape), tuple(dst.shape)))
            dst.copy_(src, non_blocking=True)
        dec = self.model.ctc_decoder
        enc = dec._encode_refs(texts, into=(self._ids_h, self._lens_h)) if hasattr(dec, "_encode_refs") else None
        if enc is None:
            return False
        _slot, _smax, cer_den, wer_den, len_den = enc
        if cer_den == 0 or wer_den == 0 or len_den == 0:
            raise ZeroDivisionError("division by zero")  # what the reference's host arithmetic raises on empty references
        self._den_h[0], self._den_h[1], self._den_h[2] = float(cer_den), float(wer_den), float(len_den)
        self._ref_dev.copy_(self._ref_host, non_blocking=True)
        return True

    # ---- device side: what gets captured (the arithmetic of ConvCTCASR._step, operands all device-resident)
    def _body(self):
        x, il, tg, tl = self._static
        m = self.model
        scores, out_lens = m.forward(x, il)
        if not self._metrics_on:
            return m.criterion(scores.transpose(0, 1), tg, out_lens, tl), None
        # decode + CER/WER feed only the logger: their own branch, forked behind the CTC kernels (beside them the edit-distance CTAs
        # slow the latency-bound lattice recursion down, call 32).  In the graph nothing waits for the branch before the very end of
        # the step, so it runs beside the backward pass -- at the default yaml's size it is a fifth of the step's kernel time.
        # `scores` / `out_lens` are kept referenced until the join: memory freed on the compute stream could otherwise be handed to a
        # backward kernel while the branch still reads it.
        main = torch.cuda.current_stream(self.device)
        loss = m.criterion(scores.transpose(0, 1), tg, out_lens, tl)
        self._metric_stream.wait_stream(main)
        with torch.cuda.stream(self._metric_stream):
            ratios = m.ctc_decoder.score_device(scores, out_lens, self._ids_d, self._lens_d) / self._den_d
        self._branch_keep = (scores, out_lens)
        return loss, ratios

    def _join(self):
        if self._metrics_on:
            torch.cuda.current_stream(self.device).wait_stream(self._metric_stream)
        self._branch_keep = None

    def _hyper(self):
        return tuple(tuple(sorted((k, repr(v)) for k, v in g.items() if k != "params")) for g in self.optimizer.param_groups)

    def _capture(self):
        if self.graph is not None:
            self.sync_optimizer_state()
        if hasattr(self.optimizer, "prepare_capture"):
            self.optimizer.prepare_capture()
        # The captured backward allocates the gradients in the graph's own pool.  Autograd binds a parameter's AccumulateGrad node to the
        # stream of the step that created it and keeps the node for as long as any earlier autograd graph is referenced (a kept loss
        # tensor): such a node would drag its old stream -- for an eager step the default stream -- into the capture and invalidate it.
        # The package's own steps keep none (log_dict stores detached values); callers must not hold an earlier step's loss either.
        self.optimizer.zero_grad(set_to_none=True)
        torch.cuda.synchronize(self.device)
        self.graph = None                                # a stale graph's pool goes before the new one is built
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._epoch.add_(1)
            loss, ratios = self._body()
            loss.backward()
            self.optimizer.step()
            self._join()
        self.graph, self._loss, self._ratios, self._key = g, loss.detach(), ratios, self._hyper()
        self._mutated = list(self.model.parameters()) + list(self.model.buffers())
        self._replays = -1                               # the optimizer's host-side step counters: the capture itself counted one step
        self.sync_optimizer_state()                      # that never ran
        self._fresh_convs = [c for plan in getattr(self.optimizer, "_plans", {}).values() for c in plan.get("convs", ())]
        # the graph holds raw pointers into the fused optimizer's plans (moments, pointer tables): keep them alive, and notice when the
        # optimizer drops them (load_state_dict / add_param_group) -- its state then lives in new tensors the graph knows nothing about
        self._plans_held = list(getattr(self.optimizer, "_plans", {}).values())
        # Jasper's NaN assertion (jasper.py:474): the captured forward left its device flag here instead of reading it
        self._nan_flag = self.model.__dict__.pop("_nan_flag_graph", None)
        self._nan_pending = []

    def __call__(self, batch, batch_idx=0):
        """One training step on ``batch``; returns the loss (a fresh 0-dim CUDA tensor) and logs what ``training_step`` logs."""
        if self._load(batch) != self._metrics_on:
            raise RuntimeError("GraphedTrainStep: this batch's transcripts %s device scoring but the captured step %s it"
                               % (("allow", "lacks") if not self._metrics_on else ("rule out", "contains")))
        plans = getattr(self.optimizer, "_plans", None)
        if plans is not None and (len(plans) != len(self._plans_held) or any(a is not b for a, b in zip(plans.values(), self._plans_held))):
            raise RuntimeError("GraphedTrainStep: the optimizer's state was replaced after the capture (load_state_dict / add_param_group); "
                               "close() this step and build a new one")
        if self._hyper() != self._key:                   # a scheduler moved the learning rate: by-value operand, capture again
            self._capture()
        self.graph.replay()
        self._replays += 1
        # the replay rewrote parameters, optimizer moments and BatchNorm statistics behind autograd's back: bump their version counters
        # as the eager step does, so that every cache keyed on them (the eval-mode BatchNorm fold, the operand copies of the weights)
        # is rebuilt by its next eager user; the bf16 operand copies the fused optimizer maintains inside the graph stay marked fresh
        torch.autograd.graph.increment_version(self._mutated)
        for conv in self._fresh_convs:
            conv.mark_shadow_fresh()
        if self._nan_flag is not None:
            self.check_nan(block=False)
            host = self._nan_free.pop() if getattr(self, "_nan_free", None) else torch.zeros(1, dtype=torch.int32).pin_memory()
            host.copy_(self._nan_flag, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._nan_pending.append((host, ev))
        loss = self._loss.clone()
        logs = {"train_loss": loss, "learning_rate": self.optimizer.param_groups[0]["lr"]}
        if self._metrics_on:
            r = self._ratios.clone()
            logs.update({"train_cer": r[0], "train_wer": r[1], "train_len_ratio": r[2]})
        self.model.log_dict(logs)
        return loss

    def sync_optimizer_state(self):
        """Adds the replayed steps to the optimizer's host-side per-parameter ``step`` counters (the reference's state layout,
        novograd.py:85-88; nothing on the device reads them).  Called by ``close()`` and before every re-capture; call it before
        ``optimizer.state_dict()`` if the counters matter."""
        n, self._replays = self._replays, 0
        if n:
            for st in self.optimizer.state.values():
                if "step" in st:
                    st["step"] += n

    def check_nan(self, block=True):
        """Raises AssertionError if a replayed forward produced a NaN (the reference asserts inside forward, jasper.py:474; a replay
        cannot, so the flag is copied out behind every replay and examined here -- without waiting unless ``block``)."""
        if not hasattr(self, "_nan_free"):
            self._nan_free = []
        while self._nan_pending:
            host, ev = self._nan_pending[0]
            if block:
                ev.synchronize()
            elif not ev.query():
                return
            self._nan_pending.pop(0)
            bad = bool(host.item())
            self._nan_free.append(host)
            assert not bad  # is there any NAN in result?

    def close(self):
        """Raises a pending NaN assertion and drops the graph with its memory pool.  (The dropout epoch stays registered: eager steps
        fold the -- then constant -- counter into their per-call seeds, which changes nothing about their statistics.)"""
        if self._nan_flag is not None:
            self.check_nan(block=True)
        if self.graph is not None:
            self.sync_optimizer_state()
        self.graph = None
