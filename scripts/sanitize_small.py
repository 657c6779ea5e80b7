#!/usr/bin/env python
"""Small renders for compute-sanitizer (memcheck / racecheck): single-level (Cornell), two-level (instanced rocks), NEE, GBuffer with
moving instances. usage: compute-sanitizer --tool racecheck python scripts/sanitize_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from merian_b200 import abi, synthetic  # noqa: E402
from merian_b200.context import Context  # noqa: E402
from merian_b200.scene import SceneTables, upload_scene  # noqa: E402

ctx = Context(0)
for name, desc, (w, h) in (("cornell", synthetic.cornell(), (96, 96)),
                           ("rocks", synthetic.instanced_rocks(n_instances=40, rock_res=9, seed=5, env_size=(64, 32)), (128, 72)),
                           ("terrain", synthetic.terrain(nx=60, ny=30, env_size=(64, 32)), (128, 72))):
    t = SceneTables(desc)
    upload_scene(ctx, t)
    cam = t.desc.camera.to_abi(0)
    for flags in (abi.RENDER_EMISSION_ON_PRIMARY, abi.RENDER_EMISSION_ON_PRIMARY | abi.RENDER_NEE):
        img = ctx.render(cam, abi.RenderParams(w, h, 0, 8, 5, 0xFF, flags, 0, 1, 0.0))
        assert np.isfinite(img).all()
    prev = np.tile(np.eye(4, dtype=np.float32), (len(t.instances), 1, 1))
    ctx.set_prev_transforms(prev)
    g = ctx.gbuffer(cam, w, h)
    ctx.set_prev_transforms(None)
    print(name, "ok", float(img[..., :3].mean()), int((g["tex1"]["instance_id"] != 0xFFFFFFFF).sum()))
ctx.close()
