#!/usr/bin/env python
"""CUDA-vs-oracle parity at the BASELINE.json sizes (round-1 review, "full-size parity"): one JSON document, written to --out.

  C2   1M-triangle terrain, 1920x1080: mpt_primary_visibility vs the oracle's BVH2 traversal (ids equal, shared-edge ties enumerated and
       < 1e-4 of pixels, barycentrics bit-exact as half2) and 1M incoherent rays through mpt_trace (ids + fp32 barycentrics bit-exact)
  C1   Cornell box 512x512 at 4096 spp: mean-luminance difference < 0.5 %, relMSE < 0.02 against the oracle's 4096-spp image rendered
       with different random numbers (frame offset), so the two estimates are independent
  C3   the bench workload's scene and camera at 256x144, 4096 spp. The sky's 1-degree sun (radiance 5e4) makes two independent 4096-spp
       estimates differ by far more than the tolerance, so bias and noise are separated: same random numbers on both arms (tolerance as
       above) + independent random numbers against the GPU-vs-GPU noise floor
  C4   2M triangles, 64 textured glTF materials (textures at --tex, default 1024^2: the oracle keeps them in host memory), 256x144, 4096 spp,
       same protocol as C3
  C5   10M instanced triangles (1000 instances under a TLAS): primary ids on a 960x540 view

The oracle is the CPU restatement (parity unpinned against the Vulkan driver: SURVEY.md §8(c)); what this pins is that the CUDA path
equals it at full size, not only on the seconds-sized cases of tests/.
  python scripts/parity_fullsize.py [--configs C1,C2,C3,C4,C5] [--spp 4096] [--out profiles/r02_parity_fullsize.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from merian_b200 import abi, synthetic  # noqa: E402
from merian_b200.context import Context  # noqa: E402
from merian_b200.scene import SceneTables, upload_scene  # noqa: E402

from helpers import compare_hits, edge_ties  # noqa: E402

TOL = {"mean_luminance_rel_diff": 5e-3, "rel_mse": 0.02}


def lum(a):
    return a[..., 0] * 0.299 + a[..., 1] * 0.587 + a[..., 2] * 0.114


def oracle_for(t):
    from oracle.oracle_py import OracleContext
    o = OracleContext()
    try:
        o.set_num_threads(len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    upload_scene(o, t)
    return o


def visibility(ctx, t, w, h, name, incoherent=0):
    upload_scene(ctx, t)
    o = oracle_for(t)
    cam = t.desc.camera.to_abi(0)
    t0 = time.perf_counter()
    ref = o.primary_visibility(cam, w, h)
    cpu_s = time.perf_counter() - t0
    got = ctx.primary_visibility(cam, w, h)
    mism, bary_err, same = compare_hits(got, ref)
    n_mis, n_ties = edge_ties(got, ref, same)
    res = {"config": name, "pixels": w * h, "hit_fraction": float((ref["instance_id"] != 0xFFFFFFFF).mean()), "mismatching_pixels": n_mis,
           "shared_edge_ties": n_ties, "mismatch_fraction": mism, "max_barycentric_diff_on_matches": bary_err, "oracle_seconds": cpu_s,
           "pass": bool(n_mis == n_ties and mism < 1e-4 and bary_err == 0.0)}
    if incoherent:
        rays = synthetic.incoherent_rays(cam, w, h)[:incoherent]
        rays["origin"] = np.array([0.0, 30.0, 0.0], np.float32)  # above the terrain, cosine-distributed about the view axis (bench_configs C2)
        r_ref, _ = o.trace(rays, return_t=True)
        r_got = ctx.trace(rays)
        same_r = (r_ref["primitive_id"] == r_got["primitive_id"]) & (r_ref["geometry_index"] == r_got["geometry_index"]) & (r_ref["instance_id"] == r_got["instance_id"])
        bits = (r_ref["barycentrics"].view(np.uint32) == r_got["barycentrics"].view(np.uint32)).all(axis=1)
        res["incoherent"] = {"rays": int(len(rays)), "hit_fraction": float((r_ref["instance_id"] != 0xFFFFFFFF).mean()), "id_mismatches": int((~same_r).sum()),
                             "barycentrics_bit_exact_on_matches": bool(bits[same_r].all())}
        res["pass"] = bool(res["pass"] and (~same_r).mean() < 1e-4 and bits[same_r].all())
    o.close()
    return res


def rel_mse(a, b):
    return float(np.mean((lum(a) - lum(b)) ** 2 / (lum(b) ** 2 + 1e-2)))


def radiance(ctx, t, w, h, spp, name, same_seed=False):
    """Independent-seed comparison against the stated tolerance. With same_seed (scenes whose variance at 4096 spp is far above the
    tolerance: the C3 / C4 sky has a 1-degree sun of radiance 5e4 that only BSDF sampling can find) two more comparisons separate bias
    from noise: the SAME random numbers on both arms (path by path the same estimate, up to fp32 round-off flipping a decision), and
    the noise floor = two GPU renders with different random numbers."""
    import torch
    upload_scene(ctx, t)
    o = oracle_for(t)
    cam = t.desc.camera.to_abi(0)
    tex1 = ctx.primary_visibility_device(cam, w, h)
    P = lambda frame: abi.RenderParams(w, h, frame, spp, 5, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    got = ctx.render_device(cam, P(0), tex1=tex1).cpu().numpy()[..., :3]
    gpu_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref = o.render(cam, P(100000))[..., :3]
    cpu_s = time.perf_counter() - t0
    lg, lr = float(lum(got).mean()), float(lum(ref).mean())
    res = {"config": name, "width": w, "height": h, "spp": spp, "mean_luminance_gpu": lg, "mean_luminance_oracle": lr,
           "mean_luminance_rel_diff": abs(lg - lr) / lr, "rel_mse": rel_mse(got, ref), "finite": bool(np.isfinite(got).all()), "tolerance": TOL,
           "gpu_seconds": gpu_s, "oracle_seconds": cpu_s, "oracle_msamples_per_s": w * h * spp / cpu_s / 1e6}
    res["pass"] = bool(res["mean_luminance_rel_diff"] < TOL["mean_luminance_rel_diff"] and res["rel_mse"] < TOL["rel_mse"])
    if same_seed:
        ref0 = o.render(cam, P(0))[..., :3]
        got2 = ctx.render_device(cam, P(200000), tex1=tex1).cpu().numpy()[..., :3]
        l0 = float(lum(ref0).mean())
        d = np.abs(lum(got) - lum(ref0)) / (np.abs(lum(ref0)) + 1e-2)
        res["same_random_numbers"] = {"mean_luminance_rel_diff": abs(lg - l0) / l0, "rel_mse": rel_mse(got, ref0),
                                      "pixels_differing_by_more_than_1e-3_relative": float((d > 1e-3).mean())}
        res["noise_floor_gpu_vs_gpu"] = {"mean_luminance_rel_diff": abs(lg - float(lum(got2).mean())) / lg, "rel_mse": rel_mse(got, got2)}
        ss, nf = res["same_random_numbers"], res["noise_floor_gpu_vs_gpu"]
        # bias would show with the same random numbers; with different ones the arms may differ by what two GPU renders differ by
        res["pass"] = bool(ss["mean_luminance_rel_diff"] < TOL["mean_luminance_rel_diff"] and ss["rel_mse"] < TOL["rel_mse"] and
                           res["mean_luminance_rel_diff"] < 3.0 * nf["mean_luminance_rel_diff"] + TOL["mean_luminance_rel_diff"] and
                           res["rel_mse"] < 2.0 * nf["rel_mse"] + TOL["rel_mse"])
        res["criterion"] = ("same random numbers: mean-luminance diff < 0.5 %, relMSE < 0.02; independent random numbers: within 3x / 2x of the "
                            "GPU-vs-GPU noise floor (+ the tolerance)")
    o.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C2,C1,C3,C4,C5")
    ap.add_argument("--spp", type=int, default=4096)
    ap.add_argument("--tex", type=int, default=1024)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_parity_fullsize.json"))
    args = ap.parse_args()
    cfgs = args.configs.split(",")
    ctx = Context(0)
    out = {"generator": synthetic.GENERATOR_VERSION, "results": []}

    def emit(r):
        out["results"].append(r)
        print(json.dumps(r), flush=True)
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)

    terrain = None
    if "C2" in cfgs or "C3" in cfgs:
        terrain = SceneTables(synthetic.terrain(nx=1000, ny=500, seed=1, env_size=(4096, 2048)))
    if "C2" in cfgs:
        emit(visibility(ctx, terrain, 1920, 1080, "C2 primary visibility, 1M-triangle terrain, 1920x1080", incoherent=1_000_000))
    if "C1" in cfgs:
        emit(radiance(ctx, SceneTables(synthetic.cornell()), 512, 512, args.spp, "C1 Cornell box 512x512"))
    if "C3" in cfgs:
        emit(radiance(ctx, terrain, 256, 144, args.spp, "C3 scene and camera (1M triangles + 4096x2048 env) at 256x144", same_seed=True))
    terrain = None
    if "C4" in cfgs:
        t4 = SceneTables(synthetic.terrain(nx=1000, ny=1000, seed=4, patches=(8, 8), textured=True, tiled_textures=True, texture_size=args.tex,
                                           env_size=(4096, 2048)))
        emit(radiance(ctx, t4, 256, 144, args.spp, f"C4 scene (2M triangles, 64 textured materials, textures {args.tex}^2) at 256x144", same_seed=True))
        t4 = None
    if "C5" in cfgs:
        t5 = SceneTables(synthetic.instanced_rocks(n_instances=1000, rock_res=71, seed=5, env_size=(1024, 512)))
        emit(visibility(ctx, t5, 960, 540, "C5 primary visibility, 10M instanced triangles (1000 instances), 960x540"))
    out["pass"] = all(r["pass"] for r in out["results"])
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    ctx.close()
    sys.exit(0 if out["pass"] else 1)


if __name__ == "__main__":
    main()
