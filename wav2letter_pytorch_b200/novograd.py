"""NovoGrad with the reference's constructor and update rule (novograd.py:12-114), executed as ONE fused
multi-tensor CUDA step (csrc/novograd.cu) instead of ~8 elementwise kernels plus a host sync per parameter:

    v_t   = ||g||^2                       if v == 0 else  b2*v + (1-b2)*||g||^2      (per tensor)
    g'    = g / (sqrt(v_t) + eps) + wd*p  [* (1-b1) when grad_averaging]
    m_t   = b1*m + g'
    p    -= lr * m_t

The same pass refreshes the bf16 weight shadow the conv kernels read (see ``attach_model``)."""
import torch
from torch.optim import Optimizer

from . import _lib
from . import functional as F


class Novograd(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.95, 0), eps=1e-8, weight_decay=0, grad_averaging=False, amsgrad=False):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay,
                                      grad_averaging=grad_averaging, amsgrad=amsgrad))
        self._conv_of = {}         # id(weight param) -> ConvParams module (bf16 shadow owner)
        self._plans = {}

    def attach_model(self, model):
        """Lets the step refresh each conv layer's packed bf16 weights in the same pass."""
        from .layers import ConvParams
        for m in model.modules():
            if isinstance(m, ConvParams):
                self._conv_of[id(m.weight)] = m
        self._plans.clear()
        return self

    def load_state_dict(self, state_dict):
        """Replaced optimizer state invalidates every cached step plan (they alias the old moments)."""
        self._plans.clear()
        super().load_state_dict(state_dict)
        self._plans.clear()

    def __setstate__(self, state):
        super().__setstate__(state)
        self.__dict__.setdefault("_conv_of", {})
        self._plans = {}

    def _plan_is_current(self, gi, plan, params):
        """a cached plan may only be reused while the optimizer state it aliases is still the state: same exp_avg tensors, and
        exp_avg_sq / max_exp_avg_sq still the 0-dim views into the plan's own vectors (a plan for another gradient set, or
        state replaced by hand, breaks that)"""
        ams = self.param_groups[gi]["amsgrad"]
        for i, p in enumerate(params):
            st = self.state.get(p)
            if st is None or st.get("exp_avg") is not plan["m"][i]:
                return False
            sq = st.get("exp_avg_sq")
            if not torch.is_tensor(sq) or sq.data_ptr() != plan["v"].data_ptr() + 4 * i:
                return False
            if ams:
                mx = st.get("max_exp_avg_sq")
                if not torch.is_tensor(mx) or mx.data_ptr() != plan["vmax"].data_ptr() + 4 * i:
                    return False
            sh = plan["shadow_of"][i]
            if sh is not None and sh.packed().data_ptr() != plan["shadow_ptr"][i]:
                return False
        return True

    def _plan(self, gi, params):
        key = (gi, tuple(id(p) for p in params))
        plan = self._plans.get(key)
        if plan is not None:
            if self._plan_is_current(gi, plan, params):
                return plan
            del self._plans[key]
        dev = params[0].device
        n = len(params)
        v = torch.zeros(n, dtype=torch.float32, device=dev)
        vmax = torch.zeros(n, dtype=torch.float32, device=dev)
        shadows, convs, shadow_of = [], [], []
        for i, p in enumerate(params):
            st = self.state[p]
            if "exp_avg" not in st:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            else:
                m = st["exp_avg"]
                if m.stride() != p.stride() or m.dtype != torch.float32 or m.device != p.device:
                    # e.g. a checkpoint written by the reference: contiguous [Cout,Cin,k] moments, while the kernel walks the
                    # parameter's own (kernel-layout) storage order -- re-lay the moments out like the parameter
                    st["exp_avg"] = torch.empty_like(p, memory_format=torch.preserve_format).copy_(m)
                if "exp_avg_sq" in st:
                    v[i] = torch.as_tensor(st["exp_avg_sq"], dtype=torch.float32).to(dev)
            st["exp_avg_sq"] = v[i]                      # 0-dim view, as in the reference's state layout
            if self.param_groups[gi]["amsgrad"]:
                if "max_exp_avg_sq" in st:
                    vmax[i] = torch.as_tensor(st["max_exp_avg_sq"], dtype=torch.float32).to(dev)
                st["max_exp_avg_sq"] = vmax[i]
            conv = self._conv_of.get(id(p))
            if conv is not None and conv.cout_pad == conv.out_channels and not getattr(conv, "f32", False) and not conv.padded:   # (fp32 mode: no bf16 shadow; padded widths: re-packed by ConvParams.packed)
                shadows.append(conv.packed().data_ptr())
                convs.append(conv)
                shadow_of.append(conv)
            else:
                shadows.append(0)
                shadow_of.append(None)
                if conv is not None:
                    convs.append(None)
        # pointer table rows: params, grads (rewritten every step), exp_avg, bf16 shadows, numel.  The pinned host
        # copies rotate so that a CPU running ahead of the GPU never overwrites a table still waiting to be uploaded.
        hosts = []
        for _ in range(4):
            host = torch.empty((5, n), dtype=torch.int64).pin_memory()
            host[0] = torch.tensor([p.data_ptr() for p in params], dtype=torch.int64)
            host[2] = torch.tensor([self.state[p]["exp_avg"].data_ptr() for p in params], dtype=torch.int64)
            host[3] = torch.tensor(shadows, dtype=torch.int64)
            host[4] = torch.tensor([p.numel() for p in params], dtype=torch.int64)
            hosts.append([host, None])
        chunk = _lib.load().w2l_novograd_chunk()
        counts = torch.tensor([(p.numel() + chunk - 1) // chunk for p in params], dtype=torch.int64)
        prefix = (torch.cumsum(counts, 0) - counts).to(torch.int32)
        plan = dict(hosts=hosts, turn=0, dev=torch.empty((5, n), dtype=torch.int64, device=dev), v=v,
                    chunk_prefix=prefix.to(dev), n_chunks=int(counts.sum()), vmax=vmax,
                    ws=torch.empty(n, dtype=torch.float32, device=dev), convs=[c for c in convs if c is not None],
                    m=[self.state[p]["exp_avg"] for p in params], shadow_of=shadow_of, shadow_ptr=shadows)
        self._plans[key] = plan
        return plan

    def prepare_capture(self):
        """Call between the eager warm-up steps and the capture of ``step()`` into a CUDA graph: gives every fused-step plan a pinned
        pointer table of its own for the capture (see ``step``).  One call per capture."""
        for plan in self._plans.values():
            n = plan["dev"].shape[1]
            plan["capture_host"] = torch.empty((5, n), dtype=torch.int64).pin_memory()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            grads = []
            for p in params:
                if not p.is_cuda:
                    raise RuntimeError("Novograd (fused): CUDA parameters required")
                g = p.grad
                if g.is_sparse:
                    raise RuntimeError("Sparse gradients are not supported.")
                if g.dtype != torch.float32 or g.stride() != p.stride():
                    g = torch.empty_like(p, memory_format=torch.preserve_format).copy_(g)
                grads.append(g)
            plan = self._plan(gi, params)
            dev = plan["dev"]
            if F.capturing():
                # the step is being captured into a CUDA graph (graph_step.py): the upload becomes a memcpy node that re-reads ITS pinned
                # table on every replay, so the table is one that no later eager step rewrites (allocated by prepare_capture -- pinned
                # allocations are not allowed while capturing); the gradients it points at live in the graph's private pool
                cap = plan.get("capture_host")
                if cap is None:
                    raise RuntimeError("Novograd: call prepare_capture() after the warm-up steps and before capturing step() in a CUDA graph")
                plan["capture_host"] = None               # owned by this capture from here on
                plan.setdefault("captured_hosts", []).append(cap)
                cap.copy_(plan["hosts"][0][0])
                cap[1] = torch.tensor([g.data_ptr() for g in grads], dtype=torch.int64)
                dev.copy_(cap, non_blocking=True)
            else:
                slot = plan["hosts"][plan["turn"] % len(plan["hosts"])]
                plan["turn"] += 1
                if slot[1] is not None:
                    slot[1].synchronize()
                slot[0][1] = torch.tensor([g.data_ptr() for g in grads], dtype=torch.int64)
                dev.copy_(slot[0], non_blocking=True)
                slot[1] = torch.cuda.Event()
                slot[1].record()
            b1, b2 = group["betas"]
            P = F._ptr
            with F._on(dev.device):
                _lib.check(lib.w2l_novograd_step(P(dev[0]), P(dev[1]), P(dev[2]), P(plan["v"]),
                                                 P(plan["vmax"]) if group["amsgrad"] else None, P(dev[3]), P(dev[4]), P(plan["chunk_prefix"]),
                                                 len(params), plan["n_chunks"], float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                                 float(group["weight_decay"]), int(bool(group["grad_averaging"])), P(plan["ws"]),
                                                 F._stream()), "novograd_step")
            for p in params:
                self.state[p]["step"] += 1
                torch.autograd.graph.increment_version(p)
            for conv in plan["convs"]:
                conv.mark_shadow_fresh()
        return loss
