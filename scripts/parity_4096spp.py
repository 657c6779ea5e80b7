#!/usr/bin/env python
"""Radiance parity at 4096 spp (north_star: per-pixel relative MSE and mean-luminance difference at 4096 spp), protocol of SURVEY.md §8(c):
the procedural Cornell box is written as glTF, loaded back through the glTF loader, and rendered by the node chain
gbuffer -> pt -> accumulate (256 iterations x 16 spp, alpha 1, infinite history) on the GPU; the converged image is written as PFM
(the Image Write leg) and compared with the CPU oracle's 4096-spp image of the same scene rendered with DIFFERENT random numbers
(frame offset), so the two estimates are statistically independent. The reference's own Vulkan `pt` cannot run in this image
(SURVEY §0.8); when it can, its `accumulate.out` PFM takes the oracle's place in the same comparison.

Stated tolerance: mean-luminance difference < 0.5 %, per-pixel relMSE (eps = 1e-2) < 0.02.
  python scripts/parity_4096spp.py [--size 256] [--out gpurun_out/parity_4096spp.json]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from merian_b200 import abi, gltf, image_io, nodes, synthetic  # noqa: E402
from merian_b200.context import Context  # noqa: E402
from merian_b200.scene import SceneTables, upload_scene  # noqa: E402


def lum(a):
    return a[..., 0] * 0.299 + a[..., 1] * 0.587 + a[..., 2] * 0.114


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--iterations", type=int, default=256)
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_4096spp.json"))
    args = ap.parse_args()
    S = args.size
    tmp = tempfile.mkdtemp()
    path = gltf.save_gltf(synthetic.cornell(), os.path.join(tmp, "cornell.gltf"))
    desc = gltf.load_gltf(path)
    desc.camera.jitter_sequence = "None"  # same primary rays every iteration, as the oracle's single frame (replicate or disable)
    t = SceneTables(desc)
    ctx = Context(0)
    upload_scene(ctx, t)
    gb, pt, acc = nodes.GBufferRT(ctx), nodes.RenderPT(ctx), nodes.Accumulate(ctx)
    gb.load_config({"width": S, "height": S})
    pt.load_config({"samples per pixel": args.spp, "max path length": 5})
    pt.describe_outputs((S, S))
    acc.load_config({"alpha": 1.0, "inf history": True, "use motion vectors": False, "gbuffer check": "never", "f32 history": False})
    import torch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(args.iterations):
        cam = desc.camera.to_abi(it)
        g = gb.process(cam)
        out, hist, _ = acc.process_fused(pt, cam, it, it, g)
    torch.cuda.synchronize()
    gpu_s = time.perf_counter() - t0
    pfm = os.path.join(tmp, "accumulate_out.pfm")
    ctx.image_write(pfm, out)
    got = image_io.load_pfm(pfm, 3)
    total_spp = args.iterations * args.spp

    from oracle.oracle_py import OracleContext
    o = OracleContext()
    upload_scene(o, t)
    cam = desc.camera.to_abi(0)
    t0 = time.perf_counter()
    ref = o.render(cam, abi.RenderParams(S, S, 100000, total_spp, 5, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0))[..., :3]
    cpu_s = time.perf_counter() - t0
    lg, lr = float(lum(got).mean()), float(lum(ref).mean())
    relmse = float(np.mean((lum(got) - lum(ref)) ** 2 / (lum(ref) ** 2 + 1e-2)))
    res = {"scene": "procedural Cornell box via glTF round trip", "size": S, "spp_total": total_spp, "iterations": args.iterations,
           "mean_luminance_gpu": lg, "mean_luminance_oracle": lr, "mean_luminance_rel_diff": abs(lg - lr) / lr, "rel_mse": relmse,
           "tolerance": {"mean_luminance_rel_diff": 5e-3, "rel_mse": 0.02},
           "pass": bool(abs(lg - lr) / lr < 5e-3 and relmse < 0.02),
           "history_counter": float(np.float16(min(args.iterations, 2048))), "gpu_seconds": gpu_s, "oracle_seconds": cpu_s, "oracle_threads": o.num_threads(),
           "gpu_msamples_per_s_incl_gbuffer_and_accumulate": S * S * total_spp / gpu_s / 1e6}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))
    sys.exit(0 if res["pass"] else 1)


if __name__ == "__main__":
    main()
