#!/usr/bin/env python
"""Where does the pipelined host-buffer path (mpt_render_host_begin / _end, two iterations in flight) spend its time? Per-step host
timestamps of the loop bench.py times as `e2e`, next to the device-timed step. One JSON line.  usage: python scripts/e2e_probe.py [steps]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from merian_b200 import abi  # noqa: E402
from merian_b200.context import Context  # noqa: E402
from merian_b200.scene import upload_scene  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    W, H, spp, L = bench.WIDTH, bench.HEIGHT, bench.SPP, bench.MAX_PATH_LENGTH
    tables = bench.build_scene(False)
    ctx = Context(0)
    upload_scene(ctx, tables)
    cam = tables.desc.camera.to_abi(0)
    dev = torch.device("cuda", 0)
    tex1 = ctx.primary_visibility_device(cam, W, H)
    out = torch.empty((H, W, 4), dtype=torch.float32, device=dev)
    P = lambda f: abi.RenderParams(W, H, f, spp, L, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0)
    for i in range(3):
        ctx.render_device(cam, P(i), tex1=tex1, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.render_device(cam, P(3), tex1=tex1, out=out)
    e0.record()
    for i in range(steps):
        ctx.render_device(cam, P(4 + i), tex1=tex1, out=out)
    e1.record()
    torch.cuda.synchronize()
    device_ms = e0.elapsed_time(e1) / steps
    h_tex1 = tex1.cpu().pin_memory()
    h_outs = [torch.empty((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(4)]
    if os.environ.get("PROBE_GC_OFF"):
        import gc
        gc.collect(); gc.freeze(); gc.disable()
    ctx.render_host(cam, P(0), h_tex1=h_tex1, h_out=h_outs[0])
    res = {"device_ms_per_step": device_ms}
    for name, in_flight in (("two_in_flight", 2), ("four_in_flight", 4), ("one_in_flight", 1)):
        for s_ in range(in_flight):
            ctx.render_host_begin(cam, P(s_), h_outs[s_], s_, h_tex1=h_tex1)
        for s_ in range(in_flight):
            ctx.render_host_end(s_)
        torch.cuda.synchronize()
        stamps, begin_ms, end_ms = [], [], []
        t0 = time.perf_counter()
        for i in range(steps):
            if i >= in_flight:
                ta = time.perf_counter()
                ctx.render_host_end(i % in_flight)
                end_ms.append((time.perf_counter() - ta) * 1e3)
            ta = time.perf_counter()
            ctx.render_host_begin(cam, P(100 + i), h_outs[i % in_flight], i % in_flight, h_tex1=h_tex1)
            begin_ms.append((time.perf_counter() - ta) * 1e3)
            stamps.append((time.perf_counter() - t0) * 1e3)
        for i in range(max(steps - in_flight, 0), steps):
            ctx.render_host_end(i % in_flight)
        total = (time.perf_counter() - t0) * 1e3
        res[name] = {"ms_per_step": total / steps, "total_ms": total, "begin_ms_max": max(begin_ms), "begin_ms_mean": sum(begin_ms) / len(begin_ms), "begin_ms": [round(x, 2) for x in begin_ms],
                     "end_ms": [round(x, 2) for x in end_ms], "loop_stamps_ms": [round(x, 1) for x in stamps]}
    # the same pipelined loop, copies only (no render): what the PCIe legs cost by themselves
    d = torch.empty_like(tex1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        d.copy_(h_tex1, non_blocking=True)
        h_outs[i % 2].copy_(out, non_blocking=True)
    torch.cuda.synchronize()
    res["copies_only_ms_per_step"] = (time.perf_counter() - t0) * 1e3 / steps
    print(json.dumps(res))


if __name__ == "__main__":
    main()
