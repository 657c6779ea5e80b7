#!/usr/bin/env python
"""CPU-only: shape of the product's BVH8 on a synthetic scene and the oracle's per-ray work counters on it.

Builds the scene's acceleration structure with the product's host builder (mpt_host_accel_build), casts the camera rays and one
generation of cosine-hemisphere bounce rays from the primary hit points (a stand-in for the continuation rays of the first bounce)
through the oracle's BVH8 walker, and prints nodes / triangles per ray, the deepest stack and a children-per-node histogram.
Used to evaluate builder changes without a GPU (env knobs MPT_BVH_* of bvh8_build.cpp apply).

  python scripts/bvh8_cpu_stats.py [--scene c3|c5|cornell] [--width 480 --height 270]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from merian_b200 import abi, synthetic  # noqa: E402
from merian_b200.host_accel import HostAccel  # noqa: E402
from merian_b200.scene import SceneTables  # noqa: E402
from oracle.oracle_py import OracleContext  # noqa: E402


def camera_rays(cam, w, h):
    o = OracleContext()
    rays = np.zeros(w * h, abi.RAY_DTYPE)
    import ctypes as C
    for y in range(h):
        for x in range(w):
            r = abi.Ray()
            o.l.oracle_camera_ray(C.byref(cam), C.c_uint32(x), C.c_uint32(y), C.c_uint32(w), C.c_uint32(h), C.byref(r))
            i = y * w + x
            rays["origin"][i] = r.origin[:]
            rays["tmin"][i] = r.tmin
            rays["direction"][i] = r.direction[:]
            rays["tmax"][i] = r.tmax
    return rays


def bounce_rays(acc, rays, hits, t, records, seed=7):
    """Cosine-hemisphere rays from the hit points, about the geometric normal on the side the ray came from."""
    ok = hits["instance_id"] != 0xFFFFFFFF
    idx = np.nonzero(ok)[0]
    tri = acc.tris[records[idx]]  # (n, 3, 4)
    v0, v1, v2 = tri[:, 0, :3], tri[:, 1, :3], tri[:, 2, :3]
    ng = np.cross(v1 - v0, v2 - v0)
    ng /= np.linalg.norm(ng, axis=1, keepdims=True)
    d_in = rays["direction"][idx]
    flip = (ng * d_in).sum(1) > 0
    ng[flip] *= -1
    u, v = hits["barycentrics"][idx, 0:1], hits["barycentrics"][idx, 1:2]
    p = (1 - u - v) * v0 + u * v1 + v * v2
    rng = np.random.default_rng(seed)
    u0, u1 = rng.random(len(idx), np.float32), rng.random(len(idx), np.float32)
    su = np.sqrt(u0)
    lx, ly, lz = su * np.cos(2 * np.pi * u1), su * np.sin(2 * np.pi * u1), np.sqrt(1 - u0)
    a = np.where(np.abs(ng[:, 0:1]) > 0.9, np.array([[0, 1, 0]], np.float32), np.array([[1, 0, 0]], np.float32))
    tx = np.cross(ng, a)
    tx /= np.linalg.norm(tx, axis=1, keepdims=True)
    ty = np.cross(ng, tx)
    d = lx[:, None] * tx + ly[:, None] * ty + lz[:, None] * ng
    out = np.zeros(len(idx), abi.RAY_DTYPE)
    out["origin"] = (p + ng * 1e-3).astype(np.float32)
    out["direction"] = d.astype(np.float32)
    out["tmin"] = 0.0
    out["tmax"] = 1e7
    return out, records[idx]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="c3")
    ap.add_argument("--width", type=int, default=480)
    ap.add_argument("--height", type=int, default=270)
    args = ap.parse_args()
    if args.scene == "c3":
        desc = synthetic.terrain(nx=1000, ny=500, seed=1, env_size=(64, 32))
    elif args.scene == "c5":
        desc = synthetic.instanced_rocks(env_size=(64, 32))
    else:
        desc = synthetic.cornell()
    tables = SceneTables(desc)
    t0 = time.perf_counter()
    acc = HostAccel(tables)
    wall = time.perf_counter() - t0
    nodes = acc.nodes
    ninner = np.array([bin(int(m)).count("1") for m in nodes["imask"]])
    lb = nodes["leaf_bits"].astype(np.uint32)
    nchild = ninner + sum((((lb >> np.uint32(3 * s)) & np.uint32(7)) != 0).astype(np.int64) for s in range(8))
    o = OracleContext()
    cam = desc.camera.to_abi(0)
    pr = camera_rays(cam, args.width, args.height)
    single = acc.view.two_level == 0
    h0, t0_, rec0, c0 = o.bvh8_trace(acc.view, pr)
    br, skip = bounce_rays(acc, pr, h0, t0_, rec0) if single else (None, None)
    res = {"scene": args.scene, "nodes": int(acc.view.node_count), "triangles": int(acc.view.triangle_count), "build_ms": acc.view.build_ms,
           "build_wall_ms": wall * 1e3, "max_blas_depth": acc.view.max_blas_depth, "tlas_depth": acc.view.tlas_depth,
           "children_per_node_mean": float(nchild.mean()), "children_hist": np.bincount(nchild, minlength=9).tolist(),
           "inner_children_mean": float(ninner.mean()),
           "primary": {k: (v / c0["rays"] if k in ("nodes", "tris", "instance_entries") else v) for k, v in c0.items()}}
    if br is not None:
        h1, t1, rec1, c1 = o.bvh8_trace(acc.view, br, skip_records=skip)
        res["bounce"] = {k: (v / c1["rays"] if k in ("nodes", "tris", "instance_entries") else v) for k, v in c1.items()}
        res["bounce"]["hit_fraction"] = float((h1["instance_id"] != 0xFFFFFFFF).mean())
    print(json.dumps(res))


if __name__ == "__main__":
    main()
