"""torchrun worker for tests/test_gpu_multi.py: PeerGradientReducer (NVLS and peer-to-peer paths) against NCCL.

Three statements per path:
  1. integers: arena filled with rank-dependent integers -> the mean over the ranks is exact and must match bit for bit;
  2. in situ: inside real training steps every launch's input is captured, the expected mean rebuilt with an NCCL all-reduce,
     and compared exactly (world 2: an fp32 sum of two addends has one rounding whatever the order);
  3. wiring: weight gradients are produced IN the arena (no copy), every rank ends with bit-identical gradients, and the result
     agrees (rel-L2) with the NCCL mean of a separate plain step.  That last comparison is loose on purpose: the step itself is
     reproducible only to fp32-atomics level (BN reductions), and a clamp gate flipping on a near-zero pre-activation moves single
     gradient elements by O(1)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wav2letter_pytorch_b200 import config, layers  # noqa: E402
from wav2letter_pytorch_b200.distributed import PeerGradientReducer, dense_view, init_process_group  # noqa: E402
from wav2letter_pytorch_b200.wav2letter import Wav2Letter  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    init_process_group("nccl", device=dev, max_ctas=0)
    cfg = config.compose(overrides=["model.mid_layers=3", "optimizer=novograd"]).model
    torch.manual_seed(0)
    model = Wav2Letter(cfg).to(dev).train()
    g = torch.Generator().manual_seed(100 + rank)                       # every rank its own utterances
    B, T, S = 4, 301, 20
    x = torch.randn(B, 64, T, generator=g).to(dev)
    il = torch.full((B,), T, dtype=torch.int32, device=dev)
    tg = torch.randint(1, 29, (B, S), generator=g, dtype=torch.int32).to(dev)
    tl = torch.full((B,), S, dtype=torch.int32, device=dev)

    def run(reducer):
        layers._seed_counter[0] = 0                                     # same dropout masks in every run
        model.zero_grad(set_to_none=True)
        out, ol = model(x, il)
        loss = model.criterion(out.transpose(0, 1), tg, ol, tl)
        loss.backward()
        if reducer is not None:
            reducer.finish()
        torch.cuda.synchronize()
        return [p.grad.detach().clone() for p in model.parameters()]

    ref = run(None)
    for t in ref:
        dist.all_reduce(dense_view(t), op=dist.ReduceOp.SUM)            # conv weights are permuted views of dense storage
        t.div_(world)
    results = {}
    for name, kw in (("nvls", dict(use_multicast=True)), ("p2p", dict(use_multicast=False))):
        red = PeerGradientReducer(model, small_numel=4096, **kw)
        if name == "nvls" and red.multicast is None:
            results[name] = "skipped (no multicast mapping)"
            red.remove()
            continue
        regions = list(red.entries.values()) + [(red.small_off, red.small_numel)]
        # ---- 1. integers
        n_arena = red.arena.numel()
        idx = torch.arange(n_arena, device=dev, dtype=torch.float32)
        covered = torch.zeros(n_arena, dtype=torch.bool, device=dev)
        for off, numel in regions:
            covered[off:off + numel] = True
        for rep in range(2):
            red.arena.copy_(((idx * 7 + rep) % 1021 - 510) * (rank + 1) * world)
            torch.cuda.synchronize()
            dist.barrier()
            for off, numel in regions:
                red.comm.wait_stream(torch.cuda.current_stream())
                red._allreduce(off, numel)
            red.comm.synchronize()
            dist.barrier()
            want = ((idx * 7 + rep) % 1021 - 510) * float(sum(range(1, world + 1)))
            bad = int(((red.arena != want) & covered).sum())
            assert bad == 0, "%s: %d arena elements differ from the exact mean (rep %d)" % (name, bad, rep)
        # ---- 2. in situ
        captured, orig = {}, red._allreduce

        def spy(off, numel):
            with torch.cuda.stream(red.comm):
                captured[off] = red.arena[off:off + numel].clone()
            orig(off, numel)
        red._allreduce = spy
        for it in range(3):                                             # repeated steps: flags / sequence numbers keep working
            captured.clear()
            got = run(red)
            for off in sorted(captured):
                exp = captured[off]
                dist.all_reduce(exp, op=dist.ReduceOp.SUM)
                exp.div_(world)
                have = red.arena[off:off + exp.numel()]
                bad = int((have != exp).sum()) if world == 2 else int(((have - exp).abs() > 1e-6 * exp.abs().max()).sum())
                assert bad == 0, "%s: step %d, arena offset %d: %d elements differ from the NCCL mean" % (name, it, off, bad)
        red._allreduce = orig
        # ---- 3. wiring
        in_place = sum(1 for m in model.modules() if getattr(m, "_grad_buffer", None) is not None
                       and m.weight.grad.data_ptr() == m._grad_buffer.data_ptr())
        worst = max(float((a - b).norm() / (b.norm() + 1e-20)) for a, b in zip(got, ref))
        sig = torch.stack([dense_view(t).view(torch.int32).to(torch.int64).sum() for t in got])
        sigs = [torch.empty_like(sig) for _ in range(world)]
        dist.all_gather(sigs, sig)
        same = all(torch.equal(s, sigs[0]) for s in sigs)
        red.remove()
        assert worst < 5e-2, (name, worst)
        assert same, name + ": ranks disagree bitwise"
        assert in_place >= 1, name + ": no weight gradient was produced in the arena"
        results[name] = "ok rel_l2_vs_separate_step=%.2e in_place=%d" % (worst, in_place)
    if rank == 0:
        print("DP_WORKER_RESULT", results, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
