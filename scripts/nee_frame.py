#!/usr/bin/env python
"""Two frames of C3 with MPT_RENDER_NEE (for an ncu launch list: which kernel carries the NEE frame)."""
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
for f in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    ctx.render_device(cam, abi.RenderParams(W, H, f, 64, 5, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY | abi.RENDER_NEE, 0, 1, 0.0), tex1=tex1, out=out)
torch.cuda.synchronize()
print("ok")
