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


class PeerGradientReducer:
    """Gradient averaging over NVLink peer memory with the library's own kernel (csrc/comm.cu) instead of NCCL collectives.

    All large gradients live in ONE symmetric arena (same offsets on every rank; allocated and exchanged through
    ``torch.distributed._symmetric_memory``, which is plumbing only): every conv layer's wgrad kernel writes straight into its slice
    (``ConvParams._grad_buffer``), so the tensor autograd stores in ``.grad`` *is* the arena slice and nothing is copied.  From the
    post-accumulate-grad hook -- i.e. right behind that layer's wgrad on the side stream -- one ``w2l_grad_allreduce`` launch per
    layer is enqueued on a communication stream: barrier, in-switch reduction of the slice this rank owns (NVLS ``multimem``; or
    peer loads in rank order when no multicast mapping exists), broadcast of the mean, barrier.  The kernel uses a few CTAs and no
    shared memory, so it co-resides with the persistent GEMM CTAs of the remaining backward pass.  Small gradients (biases,
    BatchNorm affine parameters) are staged through one packed region in ``finish()``.  Semantics = ``GradientReducer`` (DDP
    mean); every rank ends with bit-identical gradients."""

    def __init__(self, model, process_group=None, small_numel=65536, ctas=None, use_multicast=True):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        if not dist.is_initialized():
            raise RuntimeError("PeerGradientReducer needs an initialised torch.distributed process group")
        self.group = process_group if process_group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.lib = _lib.load()
        self.ctas = int(os.environ.get("W2L_COMM_CTAS", "4")) if ctas is None else int(ctas)
        params = [p for p in model.parameters() if p.requires_grad]
        if not params or not params[0].is_cuda:
            raise RuntimeError("PeerGradientReducer: CUDA parameters required")
        self.device = params[0].device
        # ---- arena layout: [large tensors, 512-byte aligned][packed small tensors]
        self.entries, self.small_params, off = {}, [], 0
        for p in params:
            if p.numel() >= small_numel:
                self.entries[id(p)] = (off, p.numel())
                off = (off + p.numel() + 127) // 128 * 128
            else:
                self.small_params.append(p)
        self.small_off = off
        self.small_slices, so = {}, 0
        for p in self.small_params:
            self.small_slices[id(p)] = (so, p.numel())
            so += (p.numel() + 3) // 4 * 4
        self.small_numel = so
        total = max(off + so, 4)
        with torch.cuda.device(self.device):
            self.arena = symm_mem.empty(total, dtype=torch.float32, device=self.device)
            self.flags = symm_mem.empty(64 * 16, dtype=torch.int32, device=self.device)
            self.arena.zero_()
            self.flags.zero_()
            torch.cuda.synchronize(self.device)
            self.h_arena = symm_mem.rendezvous(self.arena, group=self.group)
            self.h_flags = symm_mem.rendezvous(self.flags, group=self.group)
        dist.barrier(group=self.group)                                     # every rank's flags are zero before the first kernel
        import ctypes
        Arr = ctypes.c_void_p * self.world
        self._peer_data = Arr(*[int(p) for p in self.h_arena.buffer_ptrs])
        self._peer_flags = Arr(*[int(p) for p in self.h_flags.buffer_ptrs])
        mc = int(getattr(self.h_arena, "multicast_ptr", 0) or 0)
        self.multicast = ctypes.c_void_p(mc) if (use_multicast and mc and os.environ.get("W2L_COMM_P2P", "0") != "1") else None
        self.seq = 0
        self.comm = torch.cuda.Stream(device=self.device)
        # ---- hand the conv layers their arena slices; views shaped like the parameters for everything else
        from .layers import ConvParams
        conv_of = {id(m.weight): m for m in model.modules() if isinstance(m, ConvParams)}
        self.views = {}
        for p in params:
            ent = self.entries.get(id(p))
            if ent is None:
                continue
            flat = self.arena[ent[0]:ent[0] + ent[1]]
            conv = conv_of.get(id(p))
            if conv is not None:
                buf = flat.view(conv.k_eff, conv.out_channels, conv.cin_eff)
                buf._w2l_arena = True
                conv._grad_buffer = buf
                self.views[id(p)] = conv.grad_view(buf)
            else:
                self.views[id(p)] = flat.view(dense_view(p).shape)
        self.pending_small, self.hooks = [], []
        for p in params:
            self.hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self._model = model

    # ---- one library call: average arena[offset : offset+numel] over the ranks, on the communication stream
    def _allreduce(self, offset, numel):
        numel = (numel + 3) // 4 * 4
        self.seq = (self.seq + 2) & 0xFFFFFFFF
        from . import _lib
        import ctypes
        with torch.cuda.device(self.device):
            _lib.check(self.lib.w2l_grad_allreduce(self._peer_data, self._peer_flags, self.multicast, offset, numel, self.rank, self.world,
                                                   self.seq, self.ctas, ctypes.c_void_p(self.comm.cuda_stream)), "grad_allreduce")

    def _on_grad(self, p):
        ent = self.entries.get(id(p))
        if ent is None:
            self.pending_small.append(p)
            return
        from .layers import WgradStream
        main = torch.cuda.current_stream(self.device)
        self.comm.wait_stream(main)
        if WgradStream.enabled:
            self.comm.wait_stream(WgradStream.side(self.device))           # the wgrad that produced this gradient
        view = self.views[id(p)]
        g = p.grad if p.grad.shape == view.shape else dense_view(p.grad)
        staged = g.data_ptr() != view.data_ptr()                           # not produced in place (e.g. accumulated gradients)
        with torch.cuda.stream(self.comm):
            if staged:
                view.copy_(g)
            self._allreduce(ent[0], ent[1])
            if staged:
                g.copy_(view)

    def finish(self):
        """Call before ``optimizer.step()``: reduces the packed small gradients and joins the communication stream."""
        main = torch.cuda.current_stream(self.device)
        if self.pending_small:
            smalls = self.pending_small
            dst = [self.arena[self.small_off + self.small_slices[id(p)][0]: self.small_off + self.small_slices[id(p)][0] + p.numel()]
                   for p in smalls]
            src = [dense_view(p.grad).reshape(-1) for p in smalls]     # views (depthwise weights are permuted over dense storage)
            self.arena[self.small_off:self.small_off + self.small_numel].zero_()
            torch._foreach_copy_(dst, src)
            self.comm.wait_stream(main)
            self._allreduce(self.small_off, self.small_numel)
            main.wait_stream(self.comm)
            torch._foreach_copy_(src, dst)
            self.pending_small = []
        else:
            main.wait_stream(self.comm)

    def remove(self):
        for h in self.hooks:
            h.remove()
        self.hooks.clear()
        from .layers import ConvParams
        for m in self._model.modules():
            if isinstance(m, ConvParams) and hasattr(m, "_grad_buffer"):
                del m._grad_buffer


def make_gradient_reducer(model, process_group=None, **kwargs):
    """``PeerGradientReducer`` on CUDA with an NCCL group (falls back to the NCCL ``GradientReducer`` when symmetric memory cannot
    be set up, and for CPU/gloo groups); ``W2L_REDUCER=nccl`` forces the NCCL path."""
    want = os.environ.get("W2L_REDUCER", "peer")
    cuda = any(p.is_cuda for p in model.parameters())
    if want == "peer" and cuda and dist.get_backend(process_group) == "nccl":
        try:
            return PeerGradientReducer(model, process_group, **kwargs)
        except Exception as e:  # noqa: BLE001
            import warnings
            warnings.warn("PeerGradientReducer unavailable (%r); using the NCCL GradientReducer" % (e,))
    return GradientReducer(model, process_group)
