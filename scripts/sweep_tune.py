#!/usr/bin/env python
"""Tuning sweep of the extend kernel's run-time knobs on C3 in ONE process (the scene is generated once; every setting gets
a fresh context because libmpt reads the MPT_* knobs in mpt_create). Prints one JSON line per setting.
usage: python scripts/sweep_tune.py "MPT_TRI_VOTE=8 MPT_PREFETCH=1" "MPT_TRI_VOTE=0" ...   (an empty string = defaults)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from merian_b200 import abi  # noqa: E402
from merian_b200.context import Context  # noqa: E402
from merian_b200.scene import upload_scene  # noqa: E402

KNOBS = ("MPT_EXTEND_LEGACY", "MPT_POOLED_REFILL_THRESHOLD", "MPT_POOLED_DRAIN_FIRST", "MPT_L2_PERSIST", "MPT_L2_PERSIST_MB", "MPT_TRI_VOTE", "MPT_PREFETCH", "MPT_REFILL_THRESHOLD", "MPT_POOL_MSLOTS", "MPT_EXTEND_BLOCKS_PER_SM", "MPT_BVH_MAX_LEAF", "MPT_BVH_TRAV_COST", "MPT_BVH_COLLAPSE", "MPT_BVH_CPRIM", "MPT_BVH_CPRIM_TLAS", "MPT_SHADE_BLOCKS_PER_SM")


def main():
    settings = sys.argv[1:] or [""]
    scene = os.environ.get("SWEEP_SCENE", "c3")  # c3 (bench workload) | c1 (Cornell box) | c5 (instanced rocks, two-level) | small (200x100 terrain)
    W, H, spp, L = bench.WIDTH, bench.HEIGHT, bench.SPP, bench.MAX_PATH_LENGTH
    if scene == "c3":
        tables = bench.build_scene(False)
    else:
        from merian_b200 import synthetic
        from merian_b200.scene import SceneTables
        if scene == "c1":
            tables, W, H, spp = SceneTables(synthetic.cornell()), 1024, 1024, 64
        elif scene == "c5":
            tables, W, H, spp = SceneTables(synthetic.instanced_rocks(n_instances=1000, rock_res=71, seed=5, env_size=(1024, 512))), 3840, 2160, 8
        else:
            tables, spp = SceneTables(synthetic.terrain(nx=200, ny=100, env_size=(256, 128))), 32
    dev = torch.device("cuda", 0)
    ref_img = None
    for s in settings:
        for k in KNOBS:
            os.environ.pop(k, None)
        for kv in s.split():
            k, v = kv.split("=")
            os.environ[k] = v
        ctx = Context(0)
        upload_scene(ctx, tables)
        cam = tables.desc.camera.to_abi(0)
        tex1 = ctx.primary_visibility_device(cam, W, H)
        out = torch.empty((H, W, 4), dtype=torch.float32, device=dev)
        P = lambda f: abi.RenderParams(W, H, f, spp, L, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0)
        ctx.set_profiling(1)
        ctx.reset_stats()
        ctx.render_device(cam, P(0), tex1=tex1, out=out)
        cs = ctx.stats()
        img = out.clone()
        npr, tpr = cs.nodes_visited / max(cs.rays_continuation, 1), cs.tris_tested / max(cs.rays_continuation, 1)
        ctx.set_profiling(0)
        for i in range(3):
            ctx.render_device(cam, P(i), tex1=tex1, out=out)
        torch.cuda.synchronize()
        ctx.set_profiling(2)
        ctx.reset_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.render_device(cam, P(2), tex1=tex1, out=out)  # lead-in frame: the timed ones are enqueued while it runs (see bench.py)
        e0.record()
        n = 4
        for i in range(n):
            ctx.render_device(cam, P(3 + i), tex1=tex1, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        st = ctx.stats()
        diff = None
        if ref_img is None:
            ref_img = img
        else:
            diff = float((img[..., :3] - ref_img[..., :3]).abs().mean() / ref_img[..., :3].abs().mean())
        print(json.dumps({"scene": scene, "setting": s, "msamples_per_s": W * H * spp / ms / 1e3, "ms_per_frame": ms, "extend_ms_per_frame": st.extend_ms / (n + 1),
                          "shade_ms_per_frame": st.shade_ms / (n + 1), "nodes_per_ray": npr, "tris_per_ray": tpr, "rel_mean_abs_diff_vs_first": diff}), flush=True)
        ctx.close()
        del ctx


if __name__ == "__main__":
    main()
