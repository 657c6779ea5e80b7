#!/usr/bin/env python
"""Throughput + full-size property checks for the BASELINE.json configs other than the headline (bench.py = C3).

  C1  Cornell box, 512x512, 16 spp, L=5                               (the reference's CPU-runnable case)
  C2  primary visibility, 1M-triangle terrain, 1080p: coherent camera rays and incoherent rays, Mrays/s
  C4  2M triangles, 64 glTF materials x 4 textures (4096^2 RGBA8, ~17 GB), 1080p, spp/GPU configurable
  C5  10M instanced triangles (1000 x ~10k-triangle BLAS under a TLAS), 3840x2160

At these sizes the oracle cannot run; correctness is checked through size-independent properties: a frame renders
bit-identically twice, the shards of a frame sum to the frame, mpt_trace of the camera rays reproduces
mpt_primary_visibility, the image is finite. One JSON line per config on stdout.

  python scripts/bench_configs.py [--configs C1,C2,C4,C5] [--spp4 16] [--spp5 16] [--tex 4096]
Multi-GPU (torchrun): every rank renders its sample shard, NCCL reduce to rank 0, max-over-ranks device time.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from merian_b200 import abi, sharding, synthetic  # noqa: E402
from merian_b200.context import Context  # noqa: E402
from merian_b200.scene import SceneTables, upload_scene  # noqa: E402


def timed(fn, steps, warmup, world, dev):
    for _ in range(warmup):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()  # untimed lead-in step: the timed ones are enqueued while it runs (bench.py explains why)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def render_config(name, tables, W, H, spp, L, rank, world, dev, steps, warmup, extra):
    ctx = Context(dev.index)
    t0 = time.time()
    upload_scene(ctx, tables)
    upload_s = time.time() - t0
    cam = tables.desc.camera.to_abi(0)
    bvh = ctx.bvh_info()
    tex1 = ctx.primary_visibility_device(cam, W, H)
    out = torch.empty((H, W, 4), dtype=torch.float32, device=dev)
    frame = [0]

    def P(f):
        base = abi.RenderParams(W, H, f, spp * world, L, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0)
        return sharding.shard_params(base, rank, world)[0] if world > 1 else base

    def step():
        ctx.render_device(cam, P(frame[0]), tex1=tex1, out=out)
        if world > 1:
            dist.reduce(out, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                out.mul_(1.0 / (spp * world))
        frame[0] += 1

    # ---- properties at full size
    checks = {}
    a = ctx.render_device(cam, abi.RenderParams(W, H, 3, min(spp, 4), L, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0), tex1=tex1).clone()
    b = ctx.render_device(cam, abi.RenderParams(W, H, 3, min(spp, 4), L, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0), tex1=tex1)
    checks["deterministic"] = bool(torch.equal(a, b))
    checks["finite"] = bool(torch.isfinite(a).all())
    acc = torch.zeros_like(a)
    base = abi.RenderParams(W, H, 3, min(spp, 4), L, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0)
    for r in range(2):
        p, tot = sharding.shard_params(base, r, 2)
        acc += ctx.render_device(cam, p, tex1=tex1)
    checks["shards_sum_to_frame"] = bool(torch.allclose(acc / tot, a, rtol=1e-4, atol=1e-5))
    hits = tex1.cpu().numpy().view(abi.PACKED_HIT_DTYPE).reshape(H, W)
    checks["hit_fraction"] = float((hits["instance_id"] != 0xFFFFFFFF).mean())

    # traversal work per ray (one counting frame) and the extend / shade split (CUDA events inside libmpt)
    ctx.set_profiling(1)
    ctx.reset_stats()
    ctx.render_device(cam, abi.RenderParams(W, H, 5, min(spp, 4), L, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0), tex1=tex1)
    cs = ctx.stats()
    work = {"nodes_per_ray": cs.nodes_visited / max(cs.rays_continuation, 1), "tris_per_ray": cs.tris_tested / max(cs.rays_continuation, 1)}
    ctx.set_profiling(2)
    ctx.reset_stats()
    ms = timed(step, steps, warmup, world, dev)
    st = ctx.stats()
    work["extend_share"] = st.extend_ms / (steps + warmup + 1) / ms
    work["shade_share"] = st.shade_ms / (steps + warmup + 1) / ms
    ctx.set_profiling(0)
    rays = torch.tensor([float(st.rays_continuation)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(rays)
    total_steps = steps + warmup + 1
    line = {"config": name, "n_gpus": world, "width": W, "height": H, "spp_per_gpu": spp, "spp_total": spp * world, "max_path_length": L,
            "triangles": tables.triangle_count, "instances": len(tables.instances), "ms_per_frame": ms,
            "msamples_per_s": W * H * spp * world / (ms * 1e-3) / 1e6,
            "mrays_per_s": float(rays.item()) / total_steps / (ms * 1e-3) / 1e6,
            "bvh_nodes": bvh.node_count, "bvh_mb": (bvh.node_bytes + 2 * bvh.triangle_bytes) / 1e6, "bvh_build_ms": bvh.build_ms,
            "upload_s": upload_s, "upload_split_s": getattr(tables, "upload_timing", None), "work": work, "checks": checks, "mean_luminance": float((out[..., :3] * torch.tensor([0.299, 0.587, 0.114], device=dev)).sum(-1).mean()) if rank == 0 else None}
    line["generator"] = synthetic.GENERATOR_VERSION
    line["seed"] = tables.desc.seed
    line.update(extra)
    ctx.close()
    return line


def primary_config(dev, steps, warmup):
    """C2: Mrays/s of closest-hit queries; ids of the two entry points must agree."""
    tables = SceneTables(synthetic.terrain(nx=1000, ny=500, seed=1, env_size=(64, 32)))
    ctx = Context(dev.index)
    upload_scene(ctx, tables)
    W, H = 1920, 1080
    cam = tables.desc.camera.to_abi(0)
    tex1 = ctx.primary_visibility_device(cam, W, H)
    ms_coh = timed(lambda: ctx.primary_visibility_device(cam, W, H, out=tex1), steps, warmup, 1, dev)
    # the same camera rays through mpt_trace must give the same ids (self-consistency of the two kernels)
    px, py = np.meshgrid(np.arange(W), np.arange(H))
    rays = np.zeros(W * H, abi.RAY_DTYPE)
    f32 = np.float32
    U, V, Wv = (np.array(cam.U[:], f32), np.array(cam.V[:], f32), np.array(cam.W[:], f32))
    nx = (f32(2) * ((px.reshape(-1).astype(f32) + f32(0.5)) / f32(W)) - f32(1)).astype(f32)
    ny = (f32(-2) * ((py.reshape(-1).astype(f32) + f32(0.5)) / f32(H)) + f32(1)).astype(f32)
    d = (nx[:, None] * U[None] + ny[:, None] * V[None] + Wv[None]).astype(f32)
    d = (d / np.sqrt((d * d).sum(1, dtype=f32))[:, None]).astype(f32)
    rays["origin"] = np.array(cam.position[:], f32)[None]
    rays["direction"] = d
    rays["tmin"] = 0.0
    rays["tmax"] = 1e7
    d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).to(dev)
    hits = ctx.trace_device(d_rays)
    h = hits.cpu().numpy().view(abi.HIT_DTYPE).reshape(H, W)
    t1 = tex1.cpu().numpy().view(abi.PACKED_HIT_DTYPE).reshape(H, W)
    agree = float(((h["primitive_id"] == t1["primitive_id"]) & (h["instance_id"] == t1["instance_id"])).mean())
    ms_trace_coh = timed(lambda: ctx.trace_device(d_rays, hits=hits), steps, warmup, 1, dev)
    inc = synthetic.incoherent_rays(cam, W, H)
    inc["origin"] = np.array([0.0, 30.0, 0.0], np.float32)  # above the terrain, cosine-distributed about the view axis
    d_inc = torch.from_numpy(inc.view(np.float32).reshape(-1, 8)).to(dev)
    hits_inc = ctx.trace_device(d_inc)
    hit_frac_inc = float((hits_inc.cpu().numpy().view(abi.HIT_DTYPE)["instance_id"] != 0xFFFFFFFF).mean())
    ms_inc = timed(lambda: ctx.trace_device(d_inc, hits=hits_inc), steps, warmup, 1, dev)
    n = W * H
    line = {"config": "C2 primary visibility, 1M-triangle terrain, 1920x1080", "n_gpus": 1, "triangles": tables.triangle_count,
            "coherent_primary_mrays_per_s": n / (ms_coh * 1e-3) / 1e6, "coherent_trace_mrays_per_s": n / (ms_trace_coh * 1e-3) / 1e6,
            "incoherent_trace_mrays_per_s": n / (ms_inc * 1e-3) / 1e6, "incoherent_hit_fraction": hit_frac_inc,
            "checks": {"trace_ids_equal_primary_ids": agree, "hit_fraction": float((t1["instance_id"] != 0xFFFFFFFF).mean())}}
    ctx.close()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C1,C2,C4,C5")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--spp4", type=int, default=16, help="C4 samples per pixel per GPU (config: 256 in total)")
    ap.add_argument("--spp5", type=int, default=16, help="C5 samples per pixel per GPU (config: 1024 in total over 8 GPUs)")
    ap.add_argument("--tex", type=int, default=4096)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfgs = args.configs.split(",")
    out = []
    if "C1" in cfgs:
        t = SceneTables(synthetic.cornell())
        out.append(render_config("C1 Cornell box 512x512 16 spp L5", t, 512, 512, 16, 5, rank, world, dev, args.steps * 4, 3, {}))
    if "C2" in cfgs and world == 1:
        out.append(primary_config(dev, args.steps * 4, 3))
    if "C4" in cfgs:
        t = SceneTables(synthetic.terrain(nx=1000, ny=1000, seed=4, patches=(8, 8), textured=True, tiled_textures=True, texture_size=args.tex,
                                          env_size=(4096, 2048)))
        out.append(render_config(f"C4 2M triangles, 64 materials x 4 textures {args.tex}^2, 1920x1080, L5", t, 1920, 1080, args.spp4, 5, rank, world,
                                 dev, args.steps, args.warmup, {"texture_gb": 256 * args.tex * args.tex * 4 / 1e9}))
    if "C5" in cfgs:
        t = SceneTables(synthetic.instanced_rocks(n_instances=1000, rock_res=71, seed=5, env_size=(4096, 2048)))
        out.append(render_config("C5 10M instanced triangles (1000 instances), 3840x2160, L5", t, 3840, 2160, args.spp5, 5, rank, world, dev,
                                 args.steps, args.warmup, {}))
    if rank == 0:
        for l in out:
            print(json.dumps(l))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
