"""Probe (run under torchrun): symmetric-memory / multicast availability and NCCL all-reduce time of the W2L gradient set for a
few CTA limits."""
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def log(*a):
    if rank == 0:
        print(*a, flush=True)


try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
    log("symm_mem ok: world", hdl.world_size, "multicast_ptr", hex(hdl.multicast_ptr), "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs][:3],
        "signal_pad_ptrs", len(hdl.signal_pad_ptrs), "signal_pad_size", getattr(hdl, "signal_pad_size", None))
    t.fill_(rank + 1)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (4,), torch.float32)
    log("peer read", peer.tolist())
    hdl.barrier()
except Exception as e:  # noqa: BLE001
    log("symm_mem FAILED:", repr(e)[:400])

sizes = [896 * 896 * 29] * 2 + [896 * 768 * 29, 768 * 768 * 25, 768 * 768 * 25, 768 * 640 * 25, 640 * 640 * 21, 640 * 640 * 21, 640 * 512 * 21,
                                512 * 512 * 17, 512 * 512 * 17, 512 * 384 * 17, 384 * 384 * 13, 384 * 384 * 13, 384 * 256 * 13, 256 * 256 * 11,
                                256 * 256 * 11, 256 * 256 * 11, 256 * 704, 1024 * 896, 29 * 1024]
bufs = [torch.randn(n, device=dev) for n in sizes]
tot = sum(sizes) * 4 / 1e6
for ctas in [0, 2, 4, 8, 16]:
    if ctas:
        opts = dist.ProcessGroupNCCL.Options()
        opts.config.max_ctas = ctas
        opts.config.min_ctas = 1
        pg = dist.new_group(pg_options=opts)
    else:
        pg = dist.group.WORLD
    for b in bufs[:3]:
        dist.all_reduce(b, group=pg)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        for b in bufs:
            dist.all_reduce(b, op=dist.ReduceOp.AVG, group=pg)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    log("max_ctas %2d: %.2f ms for %.0f MB in %d all-reduces  (%.1f GB/s algbw)" % (ctas, ms, tot, len(bufs), tot / ms))
dist.destroy_process_group()
