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


class ShardComm:
    """The library's own NCCL path (mpt_comm_* / mpt_reduce_* of include/mpt.h): one communicator + side stream per process.

    `exchange_id(id_bytes_or_None) -> bytes` distributes rank 0's NCCL unique id to the other ranks out of band (torch.distributed
    broadcast in bench.py; a C++ host would use its own channel). The reduce runs on the library's side stream behind the render."""

    def __init__(self, device: int, rank: int, world: int, exchange_id):
        import ctypes as C

        from ._lib import MptError, lib
        self.l, self.rank, self.world = lib(), rank, world
        buf = (C.c_uint8 * abi.COMM_UNIQUE_ID_BYTES)()
        if rank == 0:
            st = self.l.mpt_comm_unique_id(buf)
            if st != abi.MPT_OK:
                raise MptError(st, self.l.mpt_comm_last_error(None).decode())
        raw = exchange_id(bytes(buf) if rank == 0 else None)
        buf = (C.c_uint8 * abi.COMM_UNIQUE_ID_BYTES).from_buffer_copy(raw)
        self.h = C.c_void_p()
        st = self.l.mpt_comm_create(device, buf, rank, world, C.byref(self.h))
        if st != abi.MPT_OK:
            raise MptError(st, self.l.mpt_comm_last_error(None).decode())

    def _check(self, st):
        if st != abi.MPT_OK:
            from ._lib import MptError
            raise MptError(st, self.l.mpt_comm_last_error(self.h).decode())

    def reduce_begin(self, partial_sum, out, width, height, spp_total, root=0, flags=0, stream=None, h_out=None, slot=0):
        """partial_sum / out: torch CUDA tensors (H, W, 4) float32; h_out: pinned host tensor or None."""
        import ctypes as C

        import torch
        st = stream if stream is not None else torch.cuda.current_stream(partial_sum.device).cuda_stream
        self._check(self.l.mpt_reduce_begin(self.h, C.c_void_p(partial_sum.data_ptr()), C.c_void_p(out.data_ptr()) if out is not None else None,
                                            C.c_void_p(h_out.data_ptr()) if h_out is not None else None, width, height, spp_total, root, flags,
                                            C.c_void_p(st), slot))

    def reduce_end(self, slot=0, stream=None):
        import ctypes as C

        import torch
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        self._check(self.l.mpt_reduce_end(self.h, slot, C.c_void_p(st)))

    def reduce_wait(self, slot=0):
        self._check(self.l.mpt_reduce_wait(self.h, slot))

    def close(self):
        if getattr(self, "h", None):
            self.l.mpt_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
