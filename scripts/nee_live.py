#!/usr/bin/env python
"""Live (not profiled) NEE frames of C3 with libmpt's own CUDA-event timing of the extend and shade launches."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from merian_b200 import abi  # noqa: E402
from merian_b200.context import Context  # noqa: E402
from merian_b200.scene import upload_scene  # noqa: E402

t = bench.build_scene(False)
ctx = Context(0)
upload_scene(ctx, t)
cam = t.desc.camera.to_abi(0)
W, H = bench.WIDTH, bench.HEIGHT
tex1 = ctx.primary_visibility_device(cam, W, H)
out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
P = lambda f, fl: abi.RenderParams(W, H, f, 64, 5, 0xFF, fl, 0, 1, 0.0)
for name, fl in (("plain", abi.RENDER_EMISSION_ON_PRIMARY), ("nee", abi.RENDER_EMISSION_ON_PRIMARY | abi.RENDER_NEE), ("plain_again", abi.RENDER_EMISSION_ON_PRIMARY)):
    for f in range(3):
        ctx.render_device(cam, P(f, fl), tex1=tex1, out=out)
    torch.cuda.synchronize()
    ctx.set_profiling(2)
    ctx.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 4
    for f in range(n):
        ctx.render_device(cam, P(10 + f, fl), tex1=tex1, out=out)
    e1.record()
    torch.cuda.synchronize()
    st = ctx.stats()
    ctx.set_profiling(0)
    print(json.dumps({"mode": name, "ms_per_frame": e0.elapsed_time(e1) / n, "extend_ms": st.extend_ms / n, "shade_ms": st.shade_ms / n,
                      "extend_launches": st.extend_launches / n}))
