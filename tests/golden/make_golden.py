#!/usr/bin/env python
"""Generates tests/golden/*.npz. Run from the repo root: `python tests/golden/make_golden.py`.

The reference (Slang + Vulkan ray queries) cannot be built or run in this image, and its own tests
hold no fixed-value vectors for visibility or radiance (§8(c)); these fixtures therefore come from
  * rng.npz       an independent numpy restatement of hash.slang:138-146 / random.slang:108-128 (merian_b200.synthetic)
  * cornell.npz   the CPU oracle (oracle/oracle.cpp) on the C1 Cornell box: GBuffer hit texture 64x64 and a 48x48,
                  8 spp, max path length 5 frame.  Regression anchors: they freeze the oracle, they do not pin it to
                  the reference ("parity unpinned" for visibility/radiance).
  * terrain.npz   the oracle's GBuffer hit texture of a 40x20x2-triangle terrain (indexed mesh, u32 indices), 96x54.
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
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
