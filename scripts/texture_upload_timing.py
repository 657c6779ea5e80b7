#!/usr/bin/env python
"""Upload rate of mpt_texture_upload_ex from pageable host memory: 16 RGBA8 textures of 4096^2 (67 MB each), with and without a mip chain."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from merian_b200 import abi  # noqa: E402
from merian_b200.context import Context  # noqa: E402

ctx = Context(0)
import ctypes, torch
px = np.random.default_rng(0).integers(0, 256, (4096, 4096, 4), dtype=np.uint8)
out = {}
for name, mode in (("no_mips", abi.MIPMAP_NONE), ("mip_chain", abi.MIPMAP_LINEAR)):
    ctx.texture_upload(0, 4096, 4096, abi.TEX_RGBA8_SRGB, abi.WRAP_REPEAT, abi.WRAP_REPEAT, abi.FILTER_LINEAR, px, mode)  # warm
    t0 = time.perf_counter()
    n = 16 if mode == abi.MIPMAP_NONE else 4
    for i in range(n):
        ctx.texture_upload(i, 4096, 4096, abi.TEX_RGBA8_SRGB, abi.WRAP_REPEAT, abi.WRAP_REPEAT, abi.FILTER_LINEAR, px, mode)
    dt = time.perf_counter() - t0
    out[name] = {"textures": n, "ms_per_texture": dt / n * 1e3, "gb_per_s": n * px.nbytes / dt / 1e9}
# reference points: a plain memcpy of the same bytes, and a pinned -> device copy
dst = np.empty_like(px)
t0 = time.perf_counter()
for i in range(8):
    np.copyto(dst, px)
out["host_memcpy_gb_per_s"] = 8 * px.nbytes / (time.perf_counter() - t0) / 1e9
hp = torch.from_numpy(px).pin_memory()
d = torch.empty_like(hp, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(8):
    d.copy_(hp, non_blocking=True)
torch.cuda.synchronize()
out["pinned_h2d_gb_per_s"] = 8 * px.nbytes / (time.perf_counter() - t0) / 1e9
out["direct"] = bool(os.environ.get("MPT_UPLOAD_DIRECT"))
print(json.dumps(out))
