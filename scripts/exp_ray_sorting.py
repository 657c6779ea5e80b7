#!/usr/bin/env python
"""Experiment: how much does ray ordering matter for the BVH8 traversal? Traces the same diffuse secondary rays (origins = primary
hit points of the C3 terrain in pixel order, cosine-hemisphere directions about +y) through mpt_trace in four orders."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from merian_b200 import abi, synthetic
from merian_b200.context import Context
from merian_b200.scene import SceneTables, upload_scene

t = SceneTables(synthetic.terrain(nx=1000, ny=500, seed=1, env_size=(64, 32)))
ctx = Context(0); upload_scene(ctx, t)
W, H = 1920, 1080
cam = t.desc.camera.to_abi(0)
dev = ctx.torch_device
# primary rays -> hit positions
inc = synthetic.incoherent_rays(cam, W, H)
px, py = np.meshgrid(np.arange(W), np.arange(H))
f32 = np.float32
U, V, Wv = (np.array(cam.U[:], f32), np.array(cam.V[:], f32), np.array(cam.W[:], f32))
nx = (2 * ((px.reshape(-1) + 0.5) / W) - 1).astype(f32); ny = (-2 * ((py.reshape(-1) + 0.5) / H) + 1).astype(f32)
d = nx[:, None] * U[None] + ny[:, None] * V[None] + Wv[None]; d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(f32)
rays = np.zeros(W * H, abi.RAY_DTYPE); rays["origin"] = np.array(cam.position[:], f32)[None]; rays["direction"] = d; rays["tmax"] = 1e7
R = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).to(dev)
hits = ctx.trace_device(R)
hit = hits[:, 0] != -1
# hit position: we do not get t back; approximate by re-tracing: use the terrain height is unknown -> use barycentric? simpler: march: origin + dir * t where t from solving y... use oracle-free trick:
# sample points: intersect with plane y=0 then offset up by 13 (above terrain max 12*~0.9)? Not on the surface. Instead generate origins on the surface from the mesh itself:
P = t.desc.meshes[0].positions
idx = t.desc.meshes[0].indices
prim = hits[hit][:, 2].long().cpu().numpy()
b = hits[hit][:, 4:6].view(torch.float32).cpu().numpy()
tri = idx[prim]
pos = (1 - b[:, :1] - b[:, 1:2]) * P[tri[:, 0]] + b[:, :1] * P[tri[:, 1]] + b[:, 1:2] * P[tri[:, 2]]
n = len(pos)
rng = np.random.default_rng(0)
u0, u1 = rng.random(n, dtype=f32), rng.random(n, dtype=f32)
su = np.sqrt(u0); dl = np.stack([su * np.cos(2 * np.pi * u1), np.sqrt(1 - u0), su * np.sin(2 * np.pi * u1)], 1).astype(f32)
sec = np.zeros(n, abi.RAY_DTYPE); sec["origin"] = pos + np.array([0, 1e-3, 0], f32); sec["direction"] = dl; sec["tmax"] = 1e7
S = torch.from_numpy(sec.view(np.float32).reshape(-1, 8)).to(dev)

def bench(rays_t, label):
    out = ctx.trace_device(rays_t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ctx.trace_device(rays_t, hits=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({"order": label, "rays": int(rays_t.shape[0]), "ms": ms, "mrays_per_s": rays_t.shape[0] / ms / 1e3, "hit_frac": float((out[:, 0] != -1).float().mean())}))

bench(S, "pixel order (coherent origins, random directions)")
perm = torch.randperm(n, device=dev)
bench(S[perm].contiguous(), "random order")
octant = ((S[:, 4] < 0).long() | ((S[:, 5] < 0).long() << 1) | ((S[:, 6] < 0).long() << 2))
order = torch.sort(octant, stable=True).indices
bench(S[order].contiguous(), "stable sort by direction octant")
# blocks of 4096 rays (pixel order) sorted by octant inside the block
blk = torch.arange(n, device=dev) // 4096
order2 = torch.sort(blk * 8 + octant, stable=True).indices
bench(S[order2].contiguous(), "octant sort inside blocks of 4096 rays")
blk = torch.arange(n, device=dev) // 512
order3 = torch.sort(blk * 8 + octant, stable=True).indices
bench(S[order3].contiguous(), "octant sort inside blocks of 512 rays")
# finer direction bins: 8 octants x 4 by dominant axis / elevation
elev = (S[:, 5].abs() > 0.7071).long()
order4 = torch.sort(blk * 16 + octant * 2 + elev, stable=True).indices
bench(S[order4].contiguous(), "octant x elevation sort inside blocks of 512 rays")
