#!/usr/bin/env python
"""Generates tests/golden/*.npz. Run from the repo root: `python tests/golden/make_golden.py`.

The reference (Slang + Vulkan ray queries) cannot be built or run in this image, and its own tests
hold no fixed-value vectors for visibility or radiance (§8(c)); these fixtures therefore come from
  * rng.npz       an independent numpy restatement of hash.slang:138-146 / random.slang:108-128 (merian_b200.synthetic)
  * cornell.npz   the CPU oracle (oracle/oracle.cpp) on the C1 Cornell box: GBuffer hit texture 64x64 and a 48x48,
                  8 spp, max path length 5 frame.  Regression anchors: they freeze the oracle, they do not pin it to
                  the reference ("parity unpinned" for visibility/radiance).
  * terrain.npz   the oracle's GBuffer hit texture of a 40x20x2-triangle terrain (indexed mesh, u32 indices), 96x54.
  * frame_nodes.npz  the oracle's GBuffer (Raytraced) texels (tex0-tex3, emission) of the Cornell box under a moved previous camera,
                  and the Accumulate node's out / history after three iterations (GBuffer check on, hard clamp 2).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from merian_b200 import abi, synthetic  # noqa: E402
from merian_b200.scene import SceneTables, upload_scene  # noqa: E402
from oracle.oracle_py import OracleContext  # noqa: E402


def main():
    rng = np.random.default_rng(42)
    v = rng.integers(0, 2 ** 32, size=(64, 4), dtype=np.uint64).astype(np.uint32)
    v[0] = 0
    seeds = synthetic.pcg4d16_np(v[:, 0], v[:, 1], v[:, 2], v[:, 3])
    floats, st = [], seeds.copy()
    for _ in range(8):
        f, st = synthetic.pcg_next_float_np(st)
        floats.append(f)
    np.savez_compressed(os.path.join(HERE, "rng.npz"), inputs=v, pcg4d16=seeds, floats=np.stack(floats, 1))

    t = SceneTables(synthetic.cornell())
    o = OracleContext()
    upload_scene(o, t)
    cam = t.desc.camera.to_abi(0)
    P = abi.RenderParams(48, 48, 0, 8, 5, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0)
    np.savez_compressed(os.path.join(HERE, "cornell.npz"), tex1=o.primary_visibility(cam, 64, 64).view(np.uint32).reshape(64, 64, 4),
                        frame=o.render(cam, P))

    t = SceneTables(synthetic.terrain(nx=40, ny=20, env_size=(64, 32)))
    o = OracleContext()
    upload_scene(o, t)
    np.savez_compressed(os.path.join(HERE, "terrain.npz"),
                        tex1=o.primary_visibility(t.desc.camera.to_abi(0), 96, 54).view(np.uint32).reshape(54, 96, 4))
    make_frame_nodes()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))




def make_frame_nodes():
    """GBuffer (Raytraced) texels of the Cornell box under a moved previous camera, and three Accumulate iterations on top of oracle frames."""
    from merian_b200.scene import CameraDesc
    from oracle.oracle_py import accumulate
    t = SceneTables(synthetic.cornell())
    o = OracleContext()
    upload_scene(o, t)
    cd = t.desc.camera
    cam = cd.to_abi(0)
    prev = CameraDesc(position=tuple(np.array(cd.position) + [0.05, 0.02, -0.03]), target=cd.target, up=cd.up, yfov_deg=cd.yfov_deg, aspect=cd.aspect,
                      near=cd.near, far=cd.far).to_abi(0)
    g = o.gbuffer(cam, 48, 48, prev_cam=prev)
    out = hist = None
    for it in range(3):
        frame = o.render(cam, abi.RenderParams(48, 48, it, 4, 5, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0))
        A = abi.AccumulateParams(48, 48, 1.0, float("inf"), 2.0, 0.8, 0.1, int(it == 0), abi.ACC_GBUFFER_CHECK)
        out, hist = accumulate(A, frame, out, hist, g["tex0"], g["tex2"])
    np.savez_compressed(os.path.join(HERE, "frame_nodes.npz"), tex0=g["tex0"], tex1=g["tex1"].view(np.uint32).reshape(48, 48, 4), tex2=g["tex2"], tex3=g["tex3"],
                        emission=g["emission"], acc_out=out, acc_history=hist)


if __name__ == "__main__":
    main()
