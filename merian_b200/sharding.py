"""Sample-index sharding of one `pt` frame over the GPUs of one box.

The reference has no multi-device path; its raygen shader loops `for s in 0..spp` per pixel with the RNG seeded by
pcg4d16(pixel, frame, s) (render_pt.slang:113-114). Because the seed depends only on (pixel, frame, sample index), ANY
partition of the sample indices yields the same set of paths: rank r of N renders indices {r, r+N, r+2N, ...} for all
pixels into a private sum buffer (MPT_RENDER_OUTPUT_SUM), the buffers are summed onto rank 0 (one NCCL reduce over
NVLink of 16 B/pixel per frame — the only exchange step of the path) and normalised once by the total sample count.
fp32 summation order differs with N, so results agree to rounding, not bit for bit.

This module is backend-agnostic host logic: `render` is any callable(cam, params) -> tensor/array of per-rank sums
(merian_b200.Context.render_device on the GPU; the CPU oracle in the gloo tests).
"""
from __future__ import annotations

import copy
from typing import Tuple

from . import abi


def shard_counts(spp_total: int, world: int) -> list:
    """Samples per rank for interleaved sharding: rank r gets ceil((spp_total - r) / world)."""
    if world < 1 or spp_total < 0:
        raise ValueError("world must be >= 1 and spp_total >= 0")
    return [max(0, (spp_total - r + world - 1) // world) for r in range(world)]


def shard_params(params: abi.RenderParams, rank: int, world: int) -> Tuple[abi.RenderParams, int]:
    """Per-rank render parameters for a frame whose TOTAL sample count is params.spp.
    Returns (rank params, total spp). The rank's output holds sums (not means)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    if params.sample_stride != 1 or params.sample_offset != 0:
        raise ValueError("shard_params expects the unsharded frame parameters (offset 0, stride 1)")
    total = int(params.spp)
    p = copy.copy(params)
    p.spp = shard_counts(total, world)[rank]
    p.sample_offset = rank
    p.sample_stride = world
    p.flags = params.flags | abi.RENDER_OUTPUT_SUM
    return p, total


def render_sharded(render, cam, params: abi.RenderParams, rank: int, world: int, dst: int = 0, group=None):
    """One sharded frame. `render(cam, rank_params)` must return a torch tensor (H, W, 4) of sums on the device the
    process group reduces on. Returns the normalised image on rank `dst` (None elsewhere)."""
    import torch
    import torch.distributed as dist

    p, total = shard_params(params, rank, world)
    out = render(cam, p)
    if not isinstance(out, torch.Tensor):
        out = torch.from_numpy(out)
    if world > 1:
        dist.reduce(out, dst=dst, op=dist.ReduceOp.SUM, group=group)
    if rank != dst:
        return None
    if total > 0 and not (params.flags & abi.RENDER_OUTPUT_SUM):
        out.mul_(1.0 / total)
    return out
