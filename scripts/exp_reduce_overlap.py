#!/usr/bin/env python
"""Multi-GPU experiment: where does a sharded frame's time go for different ways of issuing the reduce?
modes: none (no reduce), serial (reduce_begin + reduce_end on the render stream), overlap (side stream, host joins frame i-1 after enqueuing
frame i), slots (side stream, no host wait: the render stream waits for the reduce that last read the sum buffer), torch (torch.distributed.reduce on the render stream, round-1 bench). Run under torchrun; rank 0 prints one JSON line per mode."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from merian_b200 import abi, sharding  # noqa: E402
from merian_b200.context import Context  # noqa: E402
from merian_b200.scene import upload_scene  # noqa: E402


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.cuda.set_device(local)
    tables = bench.build_scene(False)
    ctx = Context(local)
    upload_scene(ctx, tables)
    W, H, spp = 1920, 1080, int(os.environ.get("EXP_SPP", "64"))
    cam = tables.desc.camera.to_abi(0)
    tex1 = ctx.primary_visibility_device(cam, W, H)
    sums = [torch.empty((H, W, 4), dtype=torch.float32, device=dev) for _ in range(2)]
    outs = [torch.empty((H, W, 4), dtype=torch.float32, device=dev) for _ in range(2)]

    def exchange(raw):
        t = torch.zeros(abi.COMM_UNIQUE_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
        dist.broadcast(t, src=0)
        return bytes(t.cpu().numpy().tobytes())
    comm = sharding.ShardComm(local, rank, world, exchange)
    P = lambda f: abi.RenderParams(W, H, f, spp, 5, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY | abi.RENDER_OUTPUT_SUM, rank, world, 0.0)
    steps = 8
    for mode in ("none", "serial", "overlap", "slots", "torch", "overlap", "slots", "none"):
        pend = [False]
        pslot = [False, False]

        def step(i):
            b = i % 2
            if pslot[b]:
                comm.reduce_end(slot=b)
                pslot[b] = False
            ctx.render_device(cam, P(i), tex1=tex1, out=sums[b])
            if mode == "serial":
                comm.reduce_begin(sums[b], outs[b], W, H, spp * world)
                comm.reduce_end()
            elif mode == "overlap":
                if pend[0]:
                    comm.reduce_wait()
                comm.reduce_begin(sums[b], outs[b], W, H, spp * world)
                pend[0] = True
            elif mode == "slots":  # no host wait: the render stream waits for the reduce that last read sums[b]
                comm.reduce_begin(sums[b], outs[b], W, H, spp * world, slot=b)
                pslot[b] = True
            elif mode == "torch":
                dist.reduce(sums[b], dst=0, op=dist.ReduceOp.SUM)
        for i in range(3):
            step(i)
        if pend[0]:
            comm.reduce_wait()
            pend[0] = False
        for b in range(2):
            if pslot[b]:
                comm.reduce_wait(slot=b)
                pslot[b] = False
        dist.barrier()
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        t_host = []
        evs[0].record()
        t0 = time.perf_counter()
        for i in range(steps):
            step(3 + i)
            evs[i + 1].record()
            t_host.append(time.perf_counter() - t0)
        if pend[0]:
            comm.reduce_end()
            pend[0] = False
        for b in range(2):
            if pslot[b]:
                comm.reduce_end(slot=b)
                pslot[b] = False
        end = torch.cuda.Event(enable_timing=True)
        end.record()
        torch.cuda.synchronize()
        per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
        tot = torch.tensor([evs[0].elapsed_time(end)], device=dev)
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"mode": mode, "world": world, "spp_per_gpu": spp, "ms_per_step_max_over_ranks": float(tot.item()) / steps,
                              "rank0_per_step_ms": [round(x, 2) for x in per], "rank0_host_enqueue_s": [round(x, 4) for x in t_host]}), flush=True)
        dist.barrier()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
