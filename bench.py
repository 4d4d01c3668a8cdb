#!/usr/bin/env python
"""bench.py -- throughput of the Wav2Letter train step on B200 (the metric BASELINE.json names).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port) on the host cores

A "step" is one full training step of Wav2Letter through the reference-shaped API: ``training_step`` (forward over
the conv stack, CTC loss, greedy decode + WER/CER as base_asr_models.py:78-85 does every step), ``backward`` and
``Novograd.step`` -- on a synthetic collated batch (B=64 utterances of 15 s per GPU, 64 mel bins, 225 labels each).

Prints ONE JSON line (rank 0): value = audio-seconds/second with inputs resident in HBM, timed with CUDA events
(max over ranks); e2e = same with pinned-host inputs copied in and the loss read back each step; roofline = the
conv implicit-GEMM kernels' achieved TFLOP/s (algorithmic FLOPs / CUDA-event time of those launches inside the timed
region) against the measured bf16 peak; cpu_baseline = the oracle port timed on this box's host cores.
"""
import argparse
import ctypes
import inspect
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UTT_SEC = 15
BATCH = 64
LEGS = None                    # test hook: [(key, arch, mid_layers, ragged)] instead of the config3 + ragged legs
CONFIG5_POINTS = ((1, 200, 50), (512, 200, 50), (64, 750, 225), (64, 3000, 600), (512, 3000, 50), (512, 3000, 600))   # (N, T, S)


# ------------------------------------------------------------------------------------------------ helpers
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def conv_flops_per_utt(specs, t_in):
    """2*T_out*Cout*Cin*k per conv per pass; fwd + wgrad for every layer, dgrad for all but the first."""
    fwd, total, t = 0.0, 0.0, t_in
    for i, s in enumerate(specs):
        from oracle.w2l_oracle import reflect_pad_amounts
        pl, pr = reflect_pad_amounts(s["cin"], s["k"], s["stride"], s["dilation"])
        t = (t + pl + pr - s["dilation"] * (s["k"] - 1) - 1) // s["stride"] + 1
        f = 2.0 * t * s["cout"] * s["cin"] * s["k"]
        fwd += f
        total += f * (2 if i == 0 else 3)
    return fwd, total


def conv_flops_model(model, t_in):
    """Algorithmic conv FLOPs per utterance (fwd, fwd+dgrad+wgrad) of a Wav2Letter / Jasper module tree: every dense conv counts
    2*T_out*Cout*Cin*k per pass, depthwise convs 2*T_out*C*k; no dgrad for the very first conv."""
    from wav2letter_pytorch_b200.layers import ConvParams, DepthwiseParams
    fwd, total, t, first = 0.0, 0.0, t_in, True
    convs = [m for m in model.modules() if isinstance(m, (ConvParams, DepthwiseParams))]
    res_ids = {id(m) for n, m in model.named_modules() if ".res." in n and isinstance(m, ConvParams)}
    for m in convs:
        k, s, d, p = m.kernel_size[0], m.stride[0], m.dilation[0], m.padding[0]
        if id(m) in res_ids:
            t_out = t                                    # 1x1 branch over the block input: same length as the block output
        elif hasattr(model, "conv1ds"):                  # Wav2Letter: reflection padding rule
            from oracle.w2l_oracle import reflect_pad_amounts
            pl, pr = reflect_pad_amounts(m.in_channels, k, s, d)
            t_out = (t + pl + pr - d * (k - 1) - 1) // s + 1
        else:
            t_out = (t + 2 * p - d * (k - 1) - 1) // s + 1
        cin = 1 if isinstance(m, DepthwiseParams) else m.in_channels
        f = 2.0 * t_out * m.out_channels * cin * k
        fwd += f
        total += f * (2 if first else 3)
        first = False
        if id(m) not in res_ids:
            t = t_out
    return fwd, total


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed regions.  A thread polls NVML in-process every 10 ms (nvidia-smi's
    start-up alone outlasts a short timed region); without NVML it reads a `nvidia-smi -lms 50` pipe.  Every sample carries a host
    timestamp and only those inside a window handed to ``mark()`` (the host interval that brackets a timed region, GPU idle on both
    sides) count: ``sm_mhz`` is their median."""

    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0, pci_bus_id=None, period_s=0.010):
        self.rows, self.windows = [], []               # rows: (host time, sm MHz, reason names)
        self.gpu, self.bus, self.period = gpu_index, pci_bus_id, period_s
        self.proc = self.thread = self.nvml = None
        self.sm_max, self.source, self._stop = None, None, threading.Event()

    # ---- NVML, in-process
    def _nvml_open(self):
        import pynvml as nv
        nv.nvmlInit()
        try:
            h = nv.nvmlDeviceGetHandleByPciBusId(self.bus.encode() if hasattr(self.bus, "encode") else self.bus) if self.bus else None
        except Exception:  # noqa: BLE001
            h = None
        if h is None:
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
        self.sm_max = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(get(h))       # fail here, not in the thread
        return nv, h, get

    def _nvml_loop(self, nv, h, get):
        while not self._stop.is_set():
            try:
                t = time.perf_counter()
                sm, mask = int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(get(h))
                self.rows.append((t, sm, tuple(n for bit, n in self.REASONS if mask & bit)))
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.period)

    # ---- nvidia-smi pipe
    def _smi_loop(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            if len(r) >= 9 and r[1].replace(".", "").isdigit():
                if r[2].replace(".", "").isdigit():
                    self.sm_max = int(float(r[2]))
                self.rows.append((time.perf_counter(), int(float(r[1])),
                                  tuple(n for i, n in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"))
                                        if r[5 + i].lower() == "active")))

    def start(self):
        try:
            args = self._nvml_open()
            self.source = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=args, daemon=True)
            self.thread.start()
            return
        except Exception:  # noqa: BLE001
            pass
        try:
            sel = ["-i", str(self.bus or self.gpu)]
            self.proc = subprocess.Popen(["nvidia-smi"] + sel + ["--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._smi_loop, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def mark(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable (no NVML, no nvidia-smi)"], "samples": 0}
        self._stop.set()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:  # noqa: BLE001
                self.proc.kill()
        self.thread.join(timeout=2)
        rows = list(self.rows)
        inside = [r for r in rows if any(a <= r[0] <= b for a, b in self.windows)]
        sm = sorted(r[1] for r in inside)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max, "reasons": sorted({n for r in inside for n in r[2]}),
                "samples": len(inside), "samples_total": len(rows), "source": self.source,
                "sm_mhz_min_max": [sm[0], sm[-1]] if sm else None,
                "window": "samples inside the timed regions of this run (device-resident, serialized-roofline and e2e legs)"}


RAGGED = False                                           # --ragged: SURVEY 8d's second run (masks / lengths exercised)


def synthetic_batch(B, seconds, seed):
    from oracle.w2l_oracle import ENGLISH_LOWERCASE
    g = torch.Generator().manual_seed(seed)
    T, S = 1 + 100 * seconds, 15 * seconds
    x = torch.randn(B, 64, T, generator=g)
    tg = torch.randint(1, 29, (B, S), generator=g, dtype=torch.int32)
    il = torch.full((B,), T, dtype=torch.int32)
    tl = torch.full((B,), S, dtype=torch.int32)
    if RAGGED:                                           # input lengths uniform in [0.6 T, T], targets in [S/2, S], zero padded as the collator does
        il = torch.randint(int(0.6 * T), T + 1, (B,), generator=g, dtype=torch.int32)
        tl = torch.randint(S // 2, S + 1, (B,), generator=g, dtype=torch.int32)
        il[0], tl[0] = T, S
        for n in range(B):
            x[n, :, il[n]:] = 0
            tg[n, tl[n]:] = 0
    texts = ["".join(ENGLISH_LOWERCASE[c] for c in row[:int(n)].tolist()) for row, n in zip(tg, tl)]
    return x, il, tg, tl, texts


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_train_step_factory(mid_layers, B, seconds, seed=0, arch="wav2letter"):
    """The reference's CPU path restated (oracle/w2l_oracle.py): fwd + CTC + greedy decode + backward + NovoGrad."""
    from oracle import w2l_oracle as O
    if arch != "wav2letter":
        return cpu_jasper_step_factory(arch, B, seconds, seed)
    specs = O.w2l_layer_specs(mid_layers)
    sd = O.w2l_init_state_dict(specs, seed=seed)
    names = [k for k, v in sd.items() if v.is_floating_point() and "running" not in k]
    for k in names:
        sd[k].requires_grad_(True)
    x, il, tg, tl, texts = synthetic_batch(B, seconds, seed)
    crit = torch.nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)
    state = [{} for _ in names]

    def step():
        lp, ol = O.w2l_forward(x, il, sd, specs, training=True)
        loss = crit(lp.transpose(0, 1), tg, ol, tl)
        am = lp.detach().argmax(-1).numpy()                       # torch.max(probs, 2), decoder.py:136
        O.greedy_collapse(am, ol.numpy())                         # the per-frame Python loop, decoder.py:104-119
        grads = torch.autograd.grad(loss, [sd[k] for k in names])
        with torch.no_grad():
            O.novograd_step([sd[k] for k in names], list(grads), state, lr=1e-3, weight_decay=1e-3)
        return float(loss.detach())

    return step


def cpu_jasper_step_factory(arch, B, seconds, seed=0):
    """Same for Jasper (jasper.py:154-475 as restated by oracle.jasper_forward): the weights are those of this package's module
    constructed on the host (reference-shaped state_dict, xavier-uniform init), the arithmetic is plain torch CPU fp32."""
    from oracle import w2l_oracle as O
    from wav2letter_pytorch_b200 import config
    from wav2letter_pytorch_b200.jasper import Jasper
    ov = ["model=%s" % arch, "optimizer=novograd"] + (["model.mid_layers=15"] if arch == "jasper" else [])
    cfg = config.compose(overrides=ov).model
    torch.manual_seed(seed)
    module = Jasper(cfg)
    blocks = [dict(b) for b in list(cfg.jasper_blocks)[:int(cfg.mid_layers)]]
    for b in blocks:
        b["dropout"] = 0
    specs = O.jasper_block_specs(blocks)
    sd = {k: v.detach().clone().contiguous() for k, v in module.state_dict().items()}
    del module
    names = [k for k, v in sd.items() if v.is_floating_point() and "running" not in k]
    for k in names:
        sd[k].requires_grad_(True)
    x, il, tg, tl, _texts = synthetic_batch(B, seconds, seed)
    crit = torch.nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)
    state = [{} for _ in names]

    def step():
        lp, ol = O.jasper_forward(x, il, sd, specs, True)
        loss = crit(lp.transpose(0, 1), tg, ol, tl)
        O.greedy_collapse(lp.detach().argmax(-1).numpy(), ol.numpy())
        grads = torch.autograd.grad(loss, [sd[k] for k in names], allow_unused=True)
        live = [(sd[k], g) for k, g in zip(names, grads) if g is not None]
        with torch.no_grad():
            O.novograd_step([p for p, _ in live], [g for _, g in live], state[:len(live)], lr=1e-3, weight_decay=1e-3)
        return float(loss.detach())

    return step


def config1_cpu_ms(mid_layers, reps=3):
    """BASELINE configs[0] on the host: Wav2Letter default config, eval-mode forward + CTCLoss + greedy decode, batch 8,
    synthetic 10 s utterances -- the oracle port of the reference's CPU path, best of ``reps`` after one warm-up."""
    from oracle import w2l_oracle as O
    specs = O.w2l_layer_specs(mid_layers)
    sd = O.w2l_init_state_dict(specs, seed=0)
    x, il, tg, tl, _texts = synthetic_batch(8, 10, seed=0)
    crit = torch.nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)
    best = 1e30
    with torch.no_grad():
        for r in range(reps + 1):
            t0 = time.perf_counter()
            lp, ol = O.w2l_forward(x, il, sd, specs, training=False)
            crit(lp.transpose(0, 1), tg, ol, tl)
            O.greedy_decode(lp.numpy(), ol.numpy())
            if r > 0:
                best = min(best, time.perf_counter() - t0)
    return best * 1e3


def run_cpu_arm(args, as_reference):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = args.cpu_batch
    step = cpu_train_step_factory(args.mid_layers, B, UTT_SEC, arch=args.model)
    for _ in range(args.warmup if as_reference else 1):
        step()
    n = args.steps if as_reference else 4                 # cpu_baseline leg of the GPU arm: 1 warm-up + 4 steps, about 10 s of CPU work
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    dt = (time.perf_counter() - t0) / n
    value = B * UTT_SEC / dt
    what = "Wav2Letter mid_layers=%d" % args.mid_layers if args.model == "wav2letter" else args.model
    sample = ("oracle port (torch CPU fp32 + Python greedy loop): %s train step, B=%d x %d s%s, mean of %d steps after %d warm-up"
              % (what, B, UTT_SEC, ", ragged lengths" if RAGGED else "", n, args.warmup if as_reference else 1))
    return dict(value=value, unit="audio-s/s", cores=cores, kind="port", sample=sample, ms_per_step=dt * 1e3)


# ------------------------------------------------------------------------------------------------ GPU arm
def reducer_report(reducer, dev, world, rank):
    """Which gradient exchange this run uses, and -- for the library's own NVLink kernel -- a parity statement made on the very
    reducer that is timed afterwards: the arena is filled with rank-dependent integers, every region goes through
    ``w2l_grad_allreduce``, and the result must equal the exact mean bit for bit on every rank (DDP semantics, README.md:40)."""
    import torch.distributed as dist
    from wav2letter_pytorch_b200.distributed import PeerGradientReducer
    if not isinstance(reducer, PeerGradientReducer):
        return "nccl", "not applicable (NCCL all-reduce)"
    kind = "peer-nvls" if reducer.multicast is not None else "peer-p2p"
    try:
        regions = list(reducer.entries.values()) + ([(reducer.small_off, reducer.small_numel)] if reducer.small_numel else [])
        n = reducer.arena.numel()
        idx = torch.arange(n, device=dev, dtype=torch.float32)
        covered = torch.zeros(n, dtype=torch.bool, device=dev)
        for off, numel in regions:
            covered[off:off + numel] = True
        bad = 0
        for rep in range(2):
            reducer.arena.copy_(((idx * 7 + rep) % 1021 - 510) * (rank + 1) * world)
            torch.cuda.synchronize()
            dist.barrier()
            for off, numel in regions:
                reducer.comm.wait_stream(torch.cuda.current_stream())
                reducer._allreduce(off, numel)
            reducer.comm.synchronize()
            dist.barrier()
            want = ((idx * 7 + rep) % 1021 - 510) * float(sum(range(1, world + 1)))
            bad += int(((reducer.arena != want) & covered).sum())
        reducer.arena.zero_()
        t = torch.tensor([bad], device=dev, dtype=torch.int64)
        dist.all_reduce(t)
        bad = int(t.item())
        return kind, ("ok: exact integer mean over %d ranks, %d arena regions, 2 rounds, bitwise on every rank" % (world, len(regions))
                      if bad == 0 else "FAILED: %d elements differ from the exact mean" % bad)
    except Exception as e:  # noqa: BLE001
        return kind, "check raised: %r" % (e,)


def _bound(fn, a, k):
    """the call's arguments by name, defaults filled in"""
    ba = inspect.signature(fn).bind(*a, **k)
    ba.apply_defaults()
    return ba.arguments


class KernelTimer:
    """CUDA-event timing of individual library calls on the launching stream (recorded inside the timed region)."""

    # algorithmic bytes of one call: CTC reads N*T*C fp32 once and writes the gradient once (+ targets, lengths,
    # nll); greedy decode reads N*T*C fp32 and writes <= T int32 tokens + a count per utterance
    BYTES = {"ctc_loss_raw": lambda a: a[0].numel() * 8 + a[1].numel() * 4 + 12 * a[0].shape[0],
             "greedy_decode": lambda a: a[0].numel() * 4 + a[0].shape[0] * a[0].shape[1] * 4 + a[0].shape[0] * 4}

    # elementwise companions of the GEMMs, from the call's bound arguments: BatchNorm apply + dropout + clamp/ReLU
    # + next layer's halo reads z (+ the residual) and writes the padded activation; its backward reads dy (padded) and z twice (reduce,
    # apply) and writes dz (+ the residual branch's gradient)
    BYTES_BOUND = {
        "bn_act_pad": lambda b: 2 * b["B"] * b["C"] * (b["T"] * (2 if b["res"] is not None else 1) + b["pad_left"] + b["T"] + b["pad_right"]),
        "bn_finalize_act_pad": lambda b: 2 * b["B"] * b["C"] * (b["T"] * (2 if b["res"] is not None else 1) + b["pad_left"] + b["T"] + b["pad_right"]),
        "bn_act_bwd": lambda b: 2 * b["B"] * b["C"] * (2 * (b["pad_left"] + b["T"] + b["pad_right"] + b["T"])
                                                     + (b["dz_rows"] or b["T"]) + (b["T"] if b["want_g"] else 0)),
    }

    def __init__(self):
        self.spans = []
        self.bytes = {}

    def reset(self):
        self.spans = []
        self.bytes = {}

    def wrap(self, F, names):
        self._orig = {n: getattr(F, n) for n in names}
        for n in names:
            def make(fn, tag):
                def timed(*a, **k):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    if tag in self.BYTES:
                        self.bytes[tag] = self.bytes.get(tag, 0) + self.BYTES[tag](a)
                    elif tag in self.BYTES_BOUND:
                        self.bytes[tag] = self.bytes.get(tag, 0) + self.BYTES_BOUND[tag](_bound(fn, a, k))
                    e0.record()
                    r = fn(*a, **k)
                    e1.record()
                    d = a[2] if len(a) > 2 and isinstance(a[2], ctypes.Structure) else None   # the conv calls' w2l_conv_desc
                    self.spans.append((tag, e0, e1, (d.B, d.T_out, d.Cin, d.Cout, d.k, d.dilation) if d is not None else None))
                    return r
                return timed
            setattr(F, n, make(self._orig[n], n))
        self._F = F

    def unwrap(self):
        for n, fn in self._orig.items():
            setattr(self._F, n, fn)

    def totals_ms(self):
        out = {}
        for tag, e0, e1, _shape in self.spans:
            out[tag] = out.get(tag, 0.0) + e0.elapsed_time(e1)
        return out

    def by_layer(self, steps, peak_tf):
        """Per (pass, layer shape): calls, ms and TFLOP/s per step from the per-launch spans -- meaningful when nothing overlaps (the
        serialized run).  FLOPs = 2*B*T_out*Cout*Cin*k of the call's own descriptor (the unfolded first layer has k folded into Cin;
        backward-data over the flat row space counts its halo rows, +2..3 %)."""
        acc = {}
        for tag, e0, e1, shape in self.spans:
            if shape is None:
                continue
            B, T_out, Cin, Cout, k, dil = shape
            a = acc.setdefault((tag, Cin, Cout, k, dil), [0, 0.0, 0.0])
            a[0] += 1
            a[1] += e0.elapsed_time(e1)
            a[2] += 2.0 * B * T_out * Cout * Cin * k
        rows = []
        for (tag, Cin, Cout, k, dil), (n, ms, flops) in sorted(acc.items(), key=lambda kv: (kv[0][0], kv[0][1] * kv[0][3], kv[0][2])):
            tf = flops / (ms / 1e3) / 1e12 if ms > 0 else 0.0
            rows.append({"pass": tag, "Cin": Cin, "Cout": Cout, "k": k, "dilation": dil, "calls_per_step": n / steps, "ms_per_step": ms / steps,
                         "tflops": tf, "frac": tf / peak_tf if peak_tf else None})
        return rows

    def union_ms(self, prefix):
        """Time during which at least one call whose name starts with ``prefix`` was in flight.  Weight-gradient GEMMs run on
        a side stream next to the compute stream, so per-call spans overlap (a span also covers the wait for SMs); the union
        is the wall time the GEMM kernels had the machine."""
        if not self.spans:
            return 0.0
        base = self.spans[0][1]
        iv = sorted((base.elapsed_time(e0), base.elapsed_time(e1)) for tag, e0, e1, _shape in self.spans if tag.startswith(prefix))
        total, cur0, cur1 = 0.0, None, None
        for a, b in iv:
            if cur1 is None or a > cur1:
                if cur1 is not None:
                    total += cur1 - cur0
                cur0, cur1 = a, b
            else:
                cur1 = max(cur1, b)
        if cur1 is not None:
            total += cur1 - cur0
        return total


def quick_leg(env, arch, mid, ragged, steps):
    """A short secondary measurement inside the same process: ms/step and audio-s/s of a train step, the conv kernels' roofline
    fraction (union of the CUDA-event spans, as the headline computes it) and -- from a serialized re-run -- the BatchNorm / activation
    passes' achieved HBM GB/s."""
    global RAGGED
    from wav2letter_pytorch_b200.layers import WgradStream
    F, build, one_step, barrier, dev, rank = (env[k] for k in ("F", "build", "one_step", "barrier", "dev", "rank"))
    was, RAGGED = RAGGED, ragged
    try:
        xb, ilb, tgb, tlb, txt = synthetic_batch(BATCH, UTT_SEC, seed=rank)
    finally:
        RAGGED = was
    batch = tuple(t.to(dev) for t in (xb, ilb, tgb, tlb))
    model, opt, reducer = build(mid, arch)
    flops = conv_flops_model(model, 1 + 100 * UTT_SEC)[1] * BATCH
    peaks = measured_peaks()
    conv_names = ["conv1d_fwd", "conv1d_dgrad", "conv1d_dgrad_wt", "conv1d_wgrad"]

    def run(n, names):
        kt = KernelTimer()
        kt.wrap(F, names)
        try:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for it in range(n):
                loss = one_step(model, opt, reducer, batch, it, txt)
            e1.record()
            barrier()
        finally:
            kt.unwrap()
        return kt, e0.elapsed_time(e1) / n, float(loss.item())

    for it in range(3):
        one_step(model, opt, reducer, batch, it, txt)
    barrier()
    kt, ms, loss = run(steps, conv_names + ["ctc_loss_raw", "greedy_decode"])
    if not (loss == loss and abs(loss) < 1e30):
        raise RuntimeError("non-finite loss %r" % loss)
    conv_ms = kt.union_ms("conv1d") / steps
    tot = {k: v / steps for k, v in kt.totals_ms().items()}
    out = {"workload": "%s train step, B=%d x %d s, %s lengths" % (arch if arch != "wav2letter" else "Wav2Letter mid_layers=%d" % mid, BATCH, UTT_SEC,
                                                                   "ragged [0.6T, T] / [S/2, S]" if ragged else "full"),
           "steps": steps, "warmup": 3, "ms_per_step": ms, "value": BATCH * UTT_SEC / (ms / 1e3), "unit": "audio-s/s", "loss_last_step": loss,
           "conv": {"flops_per_step": flops, "kernel_ms_per_step": conv_ms, "achieved_tflops": flops / (conv_ms / 1e3) / 1e12 if conv_ms else None,
                    "frac_sustained": flops / (conv_ms / 1e3) / 1e12 / peaks["tf_sustained"] if conv_ms else None,
                    "frac_burst": flops / (conv_ms / 1e3) / 1e12 / peaks["tf_burst"] if conv_ms else None}}
    for tag in ("ctc_loss_raw", "greedy_decode"):
        if tot.get(tag):
            out[tag] = {"ms_per_step": tot[tag], "GBps": kt.bytes[tag] / steps / (tot[tag] / 1e3) / 1e9,
                        "frac_hbm": kt.bytes[tag] / steps / (tot[tag] / 1e3) / 1e9 / peaks["hbm"]}
    enabled = WgradStream.enabled
    WgradStream.enabled = False                            # serialized: the BatchNorm passes' spans overlap nothing
    try:
        one_step(model, opt, reducer, batch, 0, txt)
        barrier()
        kt2, _ms2, _ = run(3, ["bn_act_pad", "bn_finalize_act_pad", "bn_act_bwd"])
    finally:
        WgradStream.enabled = enabled
    tot2, by2 = kt2.totals_ms(), dict(kt2.bytes)
    if "bn_finalize_act_pad" in tot2:
        tot2["bn_act_pad"] = tot2.pop("bn_finalize_act_pad") + tot2.get("bn_act_pad", 0.0)
        by2["bn_act_pad"] = by2.pop("bn_finalize_act_pad", 0) + by2.get("bn_act_pad", 0)
    for tag, v in tot2.items():
        by = by2.get(tag, 0) / 3
        out[tag] = {"ms_per_step": v / 3, "GBps": by / (v / 3 / 1e3) / 1e9, "frac_hbm": by / (v / 3 / 1e3) / 1e9 / peaks["hbm"]}
    del model, opt, reducer
    torch.cuda.empty_cache()
    return out


def tf32_leg(env, mid, steps=3):
    """The fp32-faithful mode (model precision 'tf32': fp32 activations / weights / gradients in memory, kind::tf32 GEMMs, fp32
    BatchNorm passes; north_star's "bf16/fp32 activations") on the headline workload: ms/step, and the first step's loss next to the
    default bf16 model's from the same initialisation and batch (they must agree to 2e-2: same network, different roundings)."""
    build, one_step, barrier, dev, rank = (env[k] for k in ("build", "one_step", "barrier", "dev", "rank"))
    xb, ilb, tgb, tlb, txt = synthetic_batch(BATCH, UTT_SEC, seed=rank)
    batch = tuple(t.to(dev) for t in (xb, ilb, tgb, tlb))
    first = {}
    for prec in ("bf16", "tf32"):
        model, opt, reducer = build(mid, "wav2letter", precision=prec)
        first[prec] = float(one_step(model, opt, reducer, batch, 0, txt).item())
        if prec == "tf32":
            one_step(model, opt, reducer, batch, 1, txt)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for it in range(steps):
                loss = one_step(model, opt, reducer, batch, it, txt)
            e1.record()
            barrier()
            ms, last = e0.elapsed_time(e1) / steps, float(loss.item())
        del model, opt, reducer
        torch.cuda.empty_cache()
    rel = abs(first["tf32"] - first["bf16"]) / abs(first["bf16"])
    if not (last == last and rel < 2e-2):
        raise RuntimeError("tf32 leg: first-step loss %r vs bf16 %r (rel %.3g), last %r" % (first["tf32"], first["bf16"], rel, last))
    return {"workload": "Wav2Letter mid_layers=%d train step, B=%d x %d s, precision=tf32 (fp32 storage, tf32 multiply, fp32 accumulate)" % (mid, BATCH, UTT_SEC),
            "steps": steps, "ms_per_step": ms, "value": BATCH * UTT_SEC / (ms / 1e3), "unit": "audio-s/s", "loss_first_step": first["tf32"],
            "loss_first_step_bf16": first["bf16"], "rel_diff": rel,
            "note": "parity mode: single-CTA kind::tf32 GEMMs, weight gradient over transposed operands -- not the throughput path"}


def config5_corners(F, dev):
    """Six corners of BASELINE config 5 (standalone CTC loss+grad and greedy decode over T x S x N, C=29, fp32 log-probs): this
    library's kernels (ms, algorithmic GB/s of SURVEY 8d against the measured HBM peak) next to torch's own CUDA ``ctc_loss`` fwd+bwd
    as the library yard-stick; L2 flushed between repetitions, best of 3."""
    import torch.nn.functional as TF
    peaks = measured_peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    C = 29

    def best_ms(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    rows = []
    for N, T, S in CONFIG5_POINTS:
        g = torch.Generator(device=dev).manual_seed(N + T)
        lp = torch.log_softmax(torch.randn(N, T, C, generator=g, device=dev), -1)
        il = torch.full((N,), T, dtype=torch.int32, device=dev)
        tg = torch.randint(1, C, (N, S), generator=g, device=dev, dtype=torch.int32)
        tl = torch.full((N,), S, dtype=torch.int32, device=dev)
        t_ctc = best_ms(lambda: F.ctc_loss_raw(lp, tg, il, tl))
        t_dec = best_ms(lambda: F.greedy_decode(lp, il))
        lpr = lp.detach().clone().requires_grad_(True)

        def torch_ctc():
            lpr.grad = None
            TF.ctc_loss(lpr.transpose(0, 1), tg, il, tl, blank=0, reduction="mean", zero_infinity=True).backward()
        try:
            t_torch = best_ms(torch_ctc)
        except RuntimeError:
            t_torch = None
        ctc_b, dec_b = N * T * C * 8 + N * S * 4 + 12 * N, N * T * C * 4 + N * T * 4 + N * 4
        rows.append({"N": N, "T": T, "S": S, "ctc_ms": t_ctc, "ctc_GBps": ctc_b / t_ctc / 1e6, "ctc_frac_hbm": ctc_b / t_ctc / 1e6 / peaks["hbm"],
                     "torch_cuda_ctc_ms": t_torch, "ctc_speedup_vs_torch_cuda": (t_torch / t_ctc) if t_torch else None,
                     "decode_ms": t_dec, "decode_GBps": dec_b / t_dec / 1e6, "decode_frac_hbm": dec_b / t_dec / 1e6 / peaks["hbm"]})
        del lp, lpr
    return {"workload": "CTC loss+grad and greedy decode, C=29 fp32, six corners of T x S x N", "hbm_peak_GBps": peaks["hbm"], "rows": rows}


def loss_check(args, build, dev):
    """Step-0 loss of the benchmarked architecture (same seed-0 weights, dropout off, train-mode BatchNorm) on two 15 s utterances:
    the CUDA path against the CPU oracle with the same storage precision emulated.  Bound 1e-2 relative; a mismatch ends the run."""
    from oracle import w2l_oracle as O
    from wav2letter_pytorch_b200 import config
    model, _opt, _red = build(args.mid_layers)
    for m in model.modules():                              # dropout off for the check (the timed steps keep the yaml values)
        if hasattr(m, "drop_out_prob"):
            m.drop_out_prob = -1.0
        if hasattr(m, "dropout_p"):
            m.dropout_p = 0.0
    sd = {k: v.detach().cpu().clone().contiguous() for k, v in model.state_dict().items()}
    x, il, tg, tl = O.synthetic_batch(2, UTT_SEC, seed=123, ragged=True)
    out, ol = model(x.to(dev), il.to(dev))
    got = float(model.criterion(out.transpose(0, 1), tg.to(dev), ol, tl.to(dev)).item())
    with torch.no_grad():
        if args.model == "wav2letter":
            lp, ol_ref = O.w2l_forward_bf16emu(x, il, sd, O.w2l_layer_specs(args.mid_layers, dropout=0.0), True)
        else:
            cfg = config.compose(overrides=["model=%s" % args.model] + (["model.mid_layers=15"] if args.model == "jasper" else [])).model
            blocks = [dict(b) for b in list(cfg.jasper_blocks)[:int(cfg.mid_layers)]]
            for b in blocks:
                b["dropout"] = 0
            lp, ol_ref = O.jasper_forward(x, il, sd, O.jasper_block_specs(blocks), True, emu=True)
        want = float(torch.nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)(lp.transpose(0, 1), tg, ol_ref, tl))
    rel = abs(got - want) / abs(want)
    del model
    torch.cuda.empty_cache()
    if not rel < 1e-2:
        raise SystemExit("bench.py: step-0 loss %.6f differs from the oracle's %.6f (rel %.2e > 1e-2)" % (got, want, rel))
    return {"cuda": got, "oracle_bf16emu": want, "rel": rel, "bound": 1e-2,
            "what": "train-mode forward + CTC loss, 2 x %d s ragged utterances, seed-0 weights, dropout off" % UTT_SEC}


def run_gpu_arm(args):
    import torch.distributed as dist
    from oracle import w2l_oracle as O
    from wav2letter_pytorch_b200 import _lib, config
    from wav2letter_pytorch_b200 import functional as F
    from wav2letter_pytorch_b200.wav2letter import Wav2Letter

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    global BATCH
    if args.strong:                                      # strong scaling: the global batch stays 64, every rank takes 64 / world
        BATCH = max(1, BATCH // world)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this arm has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        from wav2letter_pytorch_b200.distributed import init_process_group
        init_process_group("nccl", device=dev, max_ctas=int(os.environ.get("W2L_NCCL_MAX_CTAS", "4")))

    def build(mid_layers, arch=None, precision=None):
        arch = arch or args.model
        if arch == "wav2letter":
            cfg = config.compose(overrides=["model.mid_layers=%d" % mid_layers, "optimizer=novograd"]).model
            if precision:
                cfg["precision"] = precision
            cls = Wav2Letter
        else:                                            # jasper10x5 (BASELINE config 3) | jasper (the shipped separable yaml)
            from wav2letter_pytorch_b200.jasper import Jasper
            ov = ["model=%s" % arch, "optimizer=novograd"] + (["model.mid_layers=15"] if arch == "jasper" else [])
            cfg = config.compose(overrides=ov).model
            cls = Jasper
        torch.manual_seed(0)
        model = cls(cfg).to(dev).train()
        (opt,), _ = model.configure_optimizers()
        reducer = None
        if world > 1:
            from wav2letter_pytorch_b200.distributed import make_gradient_reducer
            reducer = make_gradient_reducer(model)
        return model, opt, reducer

    x, il, tg, tl, texts = synthetic_batch(BATCH, UTT_SEC, seed=rank)
    host = [t.pin_memory() for t in (x, il, tg, tl)]
    resident = tuple(t.to(dev) for t in host)
    h2d_bytes = sum(t.numel() * t.element_size() for t in host)

    def one_step(model, opt, reducer, batch, it, txt=None):
        opt.zero_grad(set_to_none=True)
        loss = model.training_step(batch + (None, texts if txt is None else txt), it)
        loss.backward()
        if reducer is not None:
            reducer.finish()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    props = torch.cuda.get_device_properties(local)
    bus = None
    if all(hasattr(props, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):   # NVML/nvidia-smi order ignores CUDA_VISIBLE_DEVICES
        bus = "%08X:%02X:%02X.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
    sampler, sampling = ClockSampler(local, bus), [True]
    step_spread = []                                      # (min, median, max) per-step ms of every timed region, in call order
    loss_trace = []                                       # loss of the last step of every timed region (asserted finite)

    def timed(model, opt, reducer, steps, warmup, from_host):
        # e2e leg: every step's inputs come from pinned host memory through the package's DevicePrefetcher (the copy of step i+1's
        # batch is enqueued on a side stream when step i's batch is handed out: it runs beside step i's kernels), and every step's loss
        # is read back; all of it inside the timed region
        from wav2letter_pytorch_b200.data_loader import DevicePrefetcher
        feed = DevicePrefetcher((host for _ in range(warmup + steps)), dev) if from_host else None
        loss_slots = [torch.zeros(1).pin_memory() for _ in range(2)] if from_host else None
        host_losses = []

        def batch():
            return next(feed) if from_host else resident
        for it in range(warmup):
            l = one_step(model, opt, reducer, batch(), it)
            if from_host:
                l.item()
        barrier()
        if steps == 0:
            return 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = []
        host_t0 = time.perf_counter()                     # GPU idle here (barrier above) and again after the barrier below
        e0.record()
        pending = None
        for it in range(steps):
            l = one_step(model, opt, reducer, batch(), it)
            if from_host:
                # device->host read of EVERY step's loss, one step late: step i's loss is copied to pinned memory on the compute stream
                # behind step i and read by the host while step i+1 is already enqueued (asynchronous logging, as a training loop
                # does it); the last one is read before the region closes
                if pending is not None:
                    pending[1].synchronize()
                    host_losses.append(float(pending[0]))
                slot = loss_slots[it % 2]
                slot.copy_(l.detach().reshape(1), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                pending = (slot, ev)
            m = torch.cuda.Event(enable_timing=True)
            m.record()
            marks.append(m)
        if pending is not None:
            pending[1].synchronize()
            host_losses.append(float(pending[0]))
        e1.record()
        barrier()
        if sampling[0]:
            sampler.mark(host_t0, time.perf_counter())
        last = float(l.item())                            # after the timed region: a step that went numerically wrong must not print a number
        if not (last == last and abs(last) < 1e30):
            raise SystemExit("bench.py: non-finite loss %r inside a timed region" % last)
        loss_trace.append(last)
        ms = e0.elapsed_time(e1) / steps
        per_step = [a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks)]
        step_spread.append((min(per_step), sorted(per_step)[len(per_step) // 2], max(per_step)))
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    from wav2letter_pytorch_b200 import reserve_device_memory
    reserve_device_memory(dev, gib=int(os.environ.get("W2L_RESERVE_GIB", "64")))      # no cudaMalloc inside the timed steps
    model, opt, reducer = build(args.mid_layers)
    reducer_kind, reducer_parity = reducer_report(reducer, dev, world, rank) if world > 1 else ("none (single GPU)", None)
    if args.model == "wav2letter":
        fwd_flops, train_flops = conv_flops_per_utt(O.w2l_layer_specs(args.mid_layers), 1 + 100 * UTT_SEC)
        assert abs(conv_flops_model(model, 1 + 100 * UTT_SEC)[1] / train_flops - 1) < 1e-9
    else:
        fwd_flops, train_flops = conv_flops_model(model, 1 + 100 * UTT_SEC)

    # ---- device-resident throughput (the headline `value`), conv kernels timed live with CUDA events
    timer = KernelTimer()
    if rank == 0:
        sampler.start()                                   # polls from here on; only samples inside the timed windows are kept
    # allocator settling (set-up, not warm-up): the caching allocator keeps one pool per stream, and with the host running several
    # steps ahead of the device the side streams (metrics, weight-shadow prefetch) need a second generation of blocks before the first is
    # released -- run call 21 saw those two cudaMalloc calls land inside the timed region once (47 instead of 37 ms/step).  Eight
    # back-to-back steps put them in front of the warm-up instead.
    timed(model, opt, reducer, 0, 8, False)
    timed(model, opt, reducer, 0, args.warmup, False)
    timer.wrap(F, ["conv1d_fwd", "conv1d_dgrad", "conv1d_dgrad_wt", "conv1d_wgrad", "ctc_loss_raw", "greedy_decode"])
    remeasured = None
    for attempt in range(2):
        timer.reset()
        launches0 = _lib.launch_count()
        seg0 = torch.cuda.memory_stats(dev).get("segment.all.allocated", 0)
        ms = timed(model, opt, reducer, args.steps, 0, False)
        new_segments = torch.cuda.memory_stats(dev).get("segment.all.allocated", 0) - seg0    # cudaMalloc calls inside the timed region
        if new_segments == 0 or attempt == 1:
            break
        # a cudaMalloc inside the timed region stalls the device pipeline: that is the allocator growing, not the step -- time the K steps again
        remeasured = "first timed region discarded: %d cudaMalloc call(s) inside it (%.2f ms/step)" % (new_segments, ms)
        step_spread.pop()
        loss_trace.pop()
    launches_total = _lib.launch_count() - launches0
    launches = launches_total // max(args.steps, 1)
    timer.unwrap()
    all_ms = {k: v / args.steps for k, v in timer.totals_ms().items()}
    conv_ms = {k: v for k, v in all_ms.items() if k.startswith("conv1d")}
    # ---- the same GEMM launches without overlap (wgrad back on the compute stream): per-kernel quality, spans do not overlap
    iso_ms, iso_layers, iso_elem = None, None, {}
    if not args.profile and os.environ.get("W2L_BENCH_SKIP_ISO", "0") != "1":
        from wav2letter_pytorch_b200.layers import WgradStream
        was = WgradStream.enabled
        WgradStream.enabled = False
        iso_timer = KernelTimer()
        timed(model, opt, reducer, 0, 1, False)
        iso_timer.wrap(F, ["conv1d_fwd", "conv1d_dgrad", "conv1d_dgrad_wt", "conv1d_wgrad", "bn_act_pad", "bn_finalize_act_pad", "bn_act_bwd"])
        iso_steps = max(3, min(args.steps, 6))
        timed(model, opt, reducer, iso_steps, 0, False)
        iso_timer.unwrap()
        WgradStream.enabled = was
        iso_all = {k: v / iso_steps for k, v in iso_timer.totals_ms().items()}
        iso_ms = {k: v for k, v in iso_all.items() if k.startswith("conv1d")}
        iso_elem = {k: (v, iso_timer.bytes.get(k, 0) / iso_steps) for k, v in iso_all.items() if not k.startswith("conv1d")}
        if "bn_finalize_act_pad" in iso_elem:              # training runs the pass with the finalize fold: one row for both entry points
            a_, b_ = iso_elem.pop("bn_finalize_act_pad"), iso_elem.get("bn_act_pad", (0.0, 0.0))
            iso_elem["bn_act_pad"] = (a_[0] + b_[0], a_[1] + b_[1])
        try:                                               # a secondary table must never take the headline line down
            iso_layers = iso_timer.by_layer(iso_steps, None)
        except Exception as e:  # noqa: BLE001
            iso_layers = {"error": repr(e)[:200]}
    # ---- end to end: pinned host inputs in, loss out, every step
    ms_e2e = ms if args.profile else timed(model, opt, reducer, args.steps, 1, True)
    sampling[0] = False                                   # the secondary tables below are launch-bound: not "under load"
    clocks = sampler.stop() if rank == 0 else None

    # ---- the literal default config (mid_layers=1), reported beside the full stack
    extra = None
    if args.model == "wav2letter" and args.mid_layers != 1 and not args.skip_default:
        del model, opt, reducer
        torch.cuda.empty_cache()
        m1, o1, r1 = build(1)
        t_host = time.perf_counter()
        ms1 = timed(m1, o1, r1, max(args.steps, 10), max(args.warmup, 3), False)
        t_host = (time.perf_counter() - t_host) * 1e3 / (max(args.steps, 10) + max(args.warmup, 3))
        extra = {"workload": "Wav2Letter mid_layers=1 (literal yaml default) train step, B=%d/GPU x %d s" % (BATCH, UTT_SEC),
                 "ms_per_step": ms1, "value": world * BATCH * UTT_SEC / (ms1 / 1e3), "unit": "audio-s/s",
                 "host_wall_ms_per_step_incl_warmup": t_host, "path": "eager (one Python launch per kernel)"}
        if world == 1 and not args.profile:
            # the same step replayed from a CUDA graph (graph_step.GraphedTrainStep): this configuration is launch-bound when eager
            try:
                from wav2letter_pytorch_b200.graph_step import GraphedTrainStep
                full = resident + (None, texts)
                gstep = GraphedTrainStep(m1, o1, full, warmup=2)
                n_g = 5 * max(args.steps, 10)
                for it in range(5):
                    gstep(full, it)
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for it in range(n_g):
                    last = gstep(full, it)
                e1.record()
                barrier()
                ms_g = e0.elapsed_time(e1) / n_g
                lv = float(last)
                assert math.isfinite(lv), "graphed default-config step produced a non-finite loss"
                gstep.close()
                extra.update({"eager_ms_per_step": ms1, "ms_per_step": ms_g, "value": world * BATCH * UTT_SEC / (ms_g / 1e3), "steps": n_g,
                              "loss_last_step": lv,
                              "path": "GraphedTrainStep: the whole step (fwd + CTC + decode + CER/WER + bwd + NovoGrad) replayed from one CUDA "
                                      "graph; per step the host copies the batch into the static operands, encodes the transcripts and "
                                      "launches the graph (all inside the timed region)"})
                del gstep
            except Exception as e:  # noqa: BLE001  (a secondary table must never take the headline line down)
                extra["graphed_error"] = repr(e)[:300]
        del m1, o1, r1

    # ---- secondary legs the driver's single line must carry (N=1 only: the scaling runs stay lean): BASELINE config 3 (Jasper 10x5),
    # SURVEY 8d's ragged run of the headline model, six corners of config 5's CTC / decode sweep, and the step-0 loss against the oracle
    legs = {}
    if world == 1 and not args.profile and not args.skip_legs:
        try:
            del model, opt, reducer
        except NameError:
            pass
        torch.cuda.empty_cache()
        want = [("ragged", args.model, args.mid_layers, True)]
        if args.model == "wav2letter":
            want.insert(0, ("config3", "jasper10x5", 0, False))
        if LEGS is not None:
            want = LEGS
        for key, arch, mid, ragged in want:
            try:
                legs[key] = quick_leg(dict(F=F, build=build, one_step=one_step, barrier=barrier, dev=dev, rank=rank), arch, mid, ragged,
                                      steps=5)
            except Exception as e:  # noqa: BLE001  (a secondary table must never take the headline line down)
                legs[key] = {"error": repr(e)[:300]}
        if args.model == "wav2letter":
            try:
                legs["precision_tf32"] = tf32_leg(dict(build=build, one_step=one_step, barrier=barrier, dev=dev, rank=rank), args.mid_layers)
            except Exception as e:  # noqa: BLE001
                legs["precision_tf32"] = {"error": repr(e)[:300]}
        try:
            legs["config5"] = config5_corners(F, dev)
        except Exception as e:  # noqa: BLE001
            legs["config5"] = {"error": repr(e)[:300]}
        if not args.skip_cpu:
            legs["loss_check"] = loss_check(args, build, dev)     # raises SystemExit on a mismatch: no number without parity

    # ---- BASELINE configs[0] (the reference's own CPU-runnable case): forward + CTCLoss + greedy decode to strings, batch 8 x 10 s,
    # eval mode, mid_layers 1 (literal default) and 20; ours on the GPU (host tensors in, strings out) next to the oracle port on the CPU
    config1 = None
    if world == 1 and not args.skip_cpu and args.model == "wav2letter":
        try:
            config1 = {"workload": "Wav2Letter default config, eval forward + CTCLoss + greedy decode (strings), B=8 x 10 s", "gpu_ms": {},
                       "cpu_ms": {}, "cpu_cores": os.cpu_count()}
            x1, il1, tg1, tl1, _t1 = synthetic_batch(8, 10, seed=0)
            for mid in (1, 20):
                m1, _o1, _r1 = build(mid, "wav2letter")
                m1.eval()
                best = 1e30
                with torch.no_grad():
                    for r in range(6):
                        torch.cuda.synchronize()
                        t0 = time.perf_counter()
                        out, ol = m1(x1.to(dev, non_blocking=True), il1.to(dev, non_blocking=True))
                        m1.criterion(out.transpose(0, 1), tg1.to(dev), ol, tl1.to(dev)).item()
                        m1.ctc_decoder.decode(out, ol)
                        torch.cuda.synchronize()
                        if r > 0:
                            best = min(best, time.perf_counter() - t0)
                config1["gpu_ms"][str(mid)] = best * 1e3
                config1["cpu_ms"][str(mid)] = config1_cpu_ms(mid)
                del m1
        except Exception as e:  # noqa: BLE001  (a secondary table must never take the headline line down)
            config1 = {"error": repr(e)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    conv_total_ms = timer.union_ms("conv1d") / args.steps     # == the sum of the spans when nothing overlaps
    achieved = train_flops * BATCH / (conv_total_ms / 1e3) / 1e12 if conv_total_ms > 0 else 0.0
    value = world * BATCH * UTT_SEC / (ms / 1e3)
    line = {
        "metric": "audio-sec/sec per train step", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": ("Wav2Letter mid_layers=%d (full layers: list of the default yaml)" % args.mid_layers if args.model == "wav2letter"
                                else {"jasper10x5": "Jasper 10x5 (dense, configuration/model/jasper10x5.yaml)",
                                      "jasper": "Jasper separable (shipped model/jasper.yaml, mid_layers=15)"}[args.model])
                               + " train step: fwd+CTC+greedy decode+WER/CER+bwd+NovoGrad, B=%d/GPU x %d s utterances, 64 mel bins, 225 labels"
                               % (BATCH, UTT_SEC),
                   "global_batch": world * BATCH, "parallelism": "dp%d" % world,
                   "lengths": "ragged: inputs uniform in [0.6 T, T], targets in [S/2, S]; audio seconds counted as padded" if RAGGED else "full",
                   "l2": "inputs+activations per step (>3 GB) exceed the 126 MB L2; no explicit flush",
                   "ctc_schedule": "log-space (alpha || beta CTAs + parallel gradient pass)",
                   "precision": "bf16 operands and activations, fp32 accumulation / statistics / master weights (the fp32-faithful tf32 "
                                "mode is measured in the precision_tf32 leg)",
                   "gemm": "CTA pairs (tcgen05 cta_group::2): conv_gemm_cg2_kernel fwd/dgrad, conv_wgrad_cg2_kernel; W2L_CG2=0 = single CTA"},
        "e2e": {"value": world * BATCH * UTT_SEC / (ms_e2e / 1e3), "unit": "audio-s/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4 + BATCH * 4,
                "h2d": "pinned host batch -> device through data_loader.DevicePrefetcher, every step, inside the timed region: the copy of "
                       "step i+1's inputs runs on a side stream beside step i's kernels",
                "d2h": "every step's loss is copied to pinned host memory behind the step and read by the host one step later (the last one "
                       "before the region closes): the host never drains the GPU between steps"},
        "gpu_launches": int(launches_total), "gpu_launches_per_step": int(launches),
        "step_ms_min_median_max": {"value": step_spread[0], "e2e": step_spread[2] if len(step_spread) > 2 else None},
        "cuda_mallocs_in_timed_region": int(new_segments), "remeasured": remeasured,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv_gemm_cg2_kernel (fwd, dgrad) + conv_wgrad_cg2_kernel: tcgen05 cta_group::2 implicit GEMMs", "achieved": achieved,
                     "peak": peaks["tf_sustained"], "peak_source": peaks["source"] + " bf16_tflops_sustained", "unit": "TFLOP/s",
                     "frac": achieved / peaks["tf_sustained"], "traffic": None,
                     "flops_per_step": train_flops * BATCH, "kernel_ms_per_step": conv_total_ms, "by_pass_span_ms": conv_ms,
                     "timing": "CUDA events on the launching streams; kernel_ms_per_step = union of the conv calls' spans (wgrad runs on a "
                               "side stream beside dgrad/BN-backward, so the per-pass spans overlap and include waiting for SMs)",
                     "share_of_step": conv_total_ms / ms},
    }
    # the two HBM-side kernels BASELINE.json's metric names, timed live inside the same steps (CUDA events around the library call)
    line["hbm_kernels"] = {}
    for tag, label in (("ctc_loss_raw", "ctc_prep + ctc_lattice (alpha || beta) + ctc_grad + ctc_finish (CTC loss + gradient)"), ("greedy_decode", "greedy_argmax+compact")):
        if all_ms.get(tag):
            gbs = timer.bytes[tag] / args.steps / (all_ms[tag] / 1e3) / 1e9
            line["hbm_kernels"][tag] = {"kernel": label, "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                                        "ms_per_step": all_ms[tag], "algorithmic_bytes_per_step": timer.bytes[tag] // args.steps,
                                        "note": "CTC is serial in T (latency/MUFU bound at this batch), see DESIGN.md 3.2" if tag == "ctc_loss_raw" else
                                                "5.6 MB per call: launch-latency bound at this size; profiles/ has the size sweep"}
    for tag, label in (("bn_act_pad", "bn_act_pad_kernel x layers (BatchNorm finalize + apply + dropout + clamp/ReLU + residual + next layer's halo)"),
                       ("bn_act_bwd", "bn_act_bwd_reduce + bn_act_bwd_apply x layers (activation/dropout/BatchNorm backward + halo fold)")):
        if tag in iso_elem and iso_elem[tag][0] > 0:       # timed in the serialized run (nothing overlaps them there), all layers together
            ms_t, by_t = iso_elem[tag]
            gbs = by_t / (ms_t / 1e3) / 1e9
            line["hbm_kernels"][tag] = {"kernel": label, "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                                        "ms_per_step": ms_t, "algorithmic_bytes_per_step": int(by_t),
                                        "note": "spans of the host calls in the serialized run (the call also launches its small torch fills)"}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):                                # dram bytes per launch from the committed ncu --set full capture
        with open(tr) as f:
            line["roofline"]["traffic"] = json.load(f).get(args.model)
    if iso_ms:
        tot = sum(iso_ms.values())
        a = train_flops * BATCH / (tot / 1e3) / 1e12
        line["roofline"]["serialized"] = {"achieved": a, "frac": a / peaks["tf_sustained"], "kernel_ms_per_step": tot, "by_pass_ms": iso_ms,
                                          "note": "same launches with wgrad on the compute stream (no overlap): sum of per-launch CUDA-event spans"}
        if isinstance(iso_layers, list):                   # per layer shape and pass: L0 (k folded into Cin) and the k=1 head stand apart
            for row in iso_layers:
                row["frac"] = row["tflops"] / peaks["tf_sustained"]
        line["roofline"]["serialized"]["by_layer"] = iso_layers
        line["roofline"]["serialized"]["frac_vs_burst"] = a / peaks["tf_burst"]
        if isinstance(iso_layers, list):                   # the record the driver keeps truncates long tails: worst rows up front
            big = [r for r in iso_layers if r["ms_per_step"] > 0.05]
            line["worst_layers"] = sorted(big, key=lambda r: r["tflops"])[:3]
    line["reducer"] = reducer_kind
    if reducer_parity is not None:
        line["reducer_parity"] = reducer_parity
    line["loss_last_step"] = {"value_leg": loss_trace[0] if loss_trace else None, "all_timed_regions": loss_trace}
    line["roofline"]["frac_vs_burst"] = achieved / peaks["tf_burst"]
    line["roofline"]["peak_burst"] = peaks["tf_burst"]
    for key in ("config3", "ragged", "precision_tf32", "config5", "loss_check"):
        if key in legs:
            line[key] = legs[key]
    if config1:
        line["config1"] = config1
    if extra:
        line["default_config"] = extra
    if world == 1 and not args.skip_cpu:
        line["cpu_baseline"] = run_cpu_arm(args, as_reference=False)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mid-layers", dest="mid_layers", type=int, default=20)
    ap.add_argument("--model", default="wav2letter", choices=["wav2letter", "jasper10x5", "jasper"],
                    help="wav2letter = BASELINE config 2 (headline); jasper10x5 = config 3; jasper = the shipped separable yaml")
    ap.add_argument("--cpu-batch", dest="cpu_batch", type=int, default=4, help="utterances in the bounded CPU sample")
    ap.add_argument("--strong", action="store_true", help="strong scaling: global batch 64 split over the ranks (default: weak, 64 per GPU)")
    ap.add_argument("--ragged", action="store_true", help="input lengths uniform in [0.6 T, T], target lengths in [S/2, S] (default: all full)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-default", action="store_true")
    ap.add_argument("--skip-legs", dest="skip_legs", action="store_true", help="no config3 / ragged / config5 / loss-check legs")
    ap.add_argument("--profile", action="store_true", help="for runs under ncu: no warm-up floor, no e2e/default/CPU passes")
    args = ap.parse_args()
    global RAGGED
    RAGGED = bool(args.ragged)
    if args.profile:
        args.skip_cpu = args.skip_default = args.skip_legs = True
    else:
        args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        r = run_cpu_arm(args, as_reference=True)
        line = {"impl": "reference", "metric": "audio-sec/sec per train step", "value": r["value"], "unit": "audio-s/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "%s train step on the host CPU (bounded sample B=%d x %d s)"
                                       % ("Wav2Letter mid_layers=%d" % args.mid_layers if args.model == "wav2letter" else args.model,
                                          args.cpu_batch, UTT_SEC),
                           "lengths": "ragged" if RAGGED else "full"},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return
    run_gpu_arm(args)


if __name__ == "__main__":
    main()
