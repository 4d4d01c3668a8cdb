"""Data-parallel plumbing: one process per GPU, gradients averaged with NCCL over NVLink, overlapped with backward.

The reference has no distributed code of its own; it delegates to Lightning's DDP (README.md:40, config.yaml:21),
i.e. bucketed gradient all-reduce (mean) overlapped with backward, per-replica BatchNorm statistics (no SyncBN) and
rank-0 buffer broadcast.  ``GradientReducer`` gives the same semantics for the modules of this package:

  * every parameter gets a post-accumulate-grad hook; the moment a layer's wgrad kernel has produced the gradient the
    all-reduce for that tensor is enqueued on NCCL's stream (it waits for the wgrad through a stream event), so the
    reduction of layer i overlaps the dgrad/wgrad kernels of layers i-1, i-2, ... that are still running;
  * ``finish()`` makes the compute stream wait for all pending reductions (no host synchronisation) -- call it before
    ``optimizer.step()``;
  * conv weights are permuted views over kernel-layout storage; the reducer all-reduces the dense storage view, so no
    gradient is copied or re-laid-out for the wire.
"""
import os

import torch
import torch.distributed as dist

_comm_ctas = [0]          # CTAs per NCCL collective chosen by init_process_group (0 = NCCL default: nothing reserved)


def dense_view(t):
    """A contiguous view of a dense, non-overlapping tensor (dims sorted by stride) sharing its memory."""
    if t.is_contiguous():
        return t
    order = sorted(range(t.dim()), key=lambda d: t.stride(d), reverse=True)
    v = t.permute(*order)
    if not v.is_contiguous():
        raise RuntimeError("gradient is not a dense tensor; cannot all-reduce in place")
    return v


class GradientReducer:
    """``small_numel``: gradients below this size (biases, BatchNorm affine parameters) are packed into one flat buffer
    and reduced with a single collective in ``finish()`` instead of ~60 latency-bound launches per step."""

    def __init__(self, model, process_group=None, small_numel=65536, comm_sms=None):
        if not dist.is_initialized():
            raise RuntimeError("GradientReducer needs an initialised torch.distributed process group")
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.use_avg = dist.get_backend(process_group) == "nccl"
        # the collective's CTAs need SMs of their own: a persistent GEMM CTA fills an SM's shared memory, so a grid of
        # #SM CTAs launched while NCCL holds a few SMs runs as two waves (measured: wgrad 11 -> 17 ms/step at 8 GPUs)
        if comm_sms is None:
            comm_sms = int(os.environ.get("W2L_COMM_SMS", str(_comm_ctas[0])))
        if self.use_avg and self.world > 1 and comm_sms > 0:
            from . import _lib
            lib = _lib.load()
            lib.w2l_set_sm_budget(0)
            lib.w2l_set_sm_budget(max(1, lib.w2l_get_sm_budget() - comm_sms))
        self.small_numel = small_numel
        self.pending, self.small, self.hooks = [], [], []
        for p in model.parameters():
            if p.requires_grad:
                self.hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _reduce(self, t):
        op = dist.ReduceOp.AVG if self.use_avg else dist.ReduceOp.SUM      # gloo (CPU tests): sum now, scale in finish()
        if t.is_cuda:
            from .layers import WgradStream
            if WgradStream.enabled:                    # weight gradients are produced on the wgrad side stream
                side = WgradStream.side(t.device)
                side.wait_stream(torch.cuda.current_stream(t.device))
                with torch.cuda.stream(side):
                    return dist.all_reduce(t, op=op, group=self.group, async_op=True)
        return dist.all_reduce(t, op=op, group=self.group, async_op=True)

    def _on_grad(self, p):
        g = dense_view(p.grad)
        if g.numel() < self.small_numel:
            self.small.append(g)
        else:
            self.pending.append((self._reduce(g), g))

    def finish(self):
        flat = None
        if self.small:
            flat = torch.cat([g.reshape(-1) for g in self.small])
            self.pending.append((self._reduce(flat), flat))
        for work, g in self.pending:
            work.wait()
            if not self.use_avg:
                g.div_(self.world)
        if flat is not None:
            torch._foreach_copy_([g.reshape(-1) for g in self.small], list(torch.split(flat, [g.numel() for g in self.small])))
        self.pending.clear()
        self.small.clear()

    def remove(self):
        for h in self.hooks:
            h.remove()
        self.hooks.clear()


def init_process_group(backend="nccl", device=None, max_ctas=4, **kwargs):
    """``dist.init_process_group`` with NCCL limited to a few CTAs per collective: the gradient all-reduce of one layer
    (<= 93 MB) has the whole remaining backward pass to hide behind, while every SM it occupies slows the persistent
    one-CTA-per-SM GEMM kernels running next to it."""
    if backend == "nccl" and max_ctas:
        try:
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = int(max_ctas)
            opts.config.min_ctas = 1
            kwargs.setdefault("pg_options", opts)
            _comm_ctas[0] = int(max_ctas)
        except Exception:  # noqa: BLE001  (older torch: fall back to NCCL's defaults)
            pass
    if device is not None:
        kwargs.setdefault("device_id", device)
    return dist.init_process_group(backend, **kwargs)


def broadcast_buffers(model, src=0, process_group=None):
    """DDP's rank-0 buffer broadcast (BatchNorm running statistics are per-replica during training)."""
    for b in model.buffers():
        dist.broadcast(b, src=src, group=process_group)


def shard_batch(batch, rank, world):
    """Utterances [rank*B/W, (rank+1)*B/W) of a collated batch (inputs, input_lengths, targets, target_lengths, paths, texts)."""
    n = batch[0].shape[0]
    lo, hi = rank * n // world, (rank + 1) * n // world
    return tuple(b[lo:hi] if b is not None else None for b in batch)
