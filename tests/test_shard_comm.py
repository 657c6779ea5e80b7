"""The library's own exchange step (mpt_comm_* / mpt_reduce_*, include/mpt.h): NCCL inside libmpt.so. On one GPU the communicator has
one rank, which still exercises the whole path a C++ host would drive: unique id, ncclCommInitRank, the side stream behind the render
stream, reduce -> normalise -> copy to a pinned host buffer, and the two ways of joining it."""
import numpy as np
import pytest

from merian_b200 import abi, sharding
from merian_b200.scene import upload_scene

from helpers import tables

pytestmark = pytest.mark.gpu


def test_single_rank_reduce_normalises_and_lands_in_host_memory(gpu_ctx):
    import torch
    t = tables("cornell")
    upload_scene(gpu_ctx, t)
    cam = t.desc.camera.to_abi(0)
    w = h = 64
    comm = sharding.ShardComm(0, 0, 1, lambda raw: raw)
    spp = 8
    base = abi.RenderParams(w, h, 1, spp, 5, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0)
    want = gpu_ctx.render_device(cam, base).clone()
    p, total = sharding.shard_params(base, 0, 1)
    part = gpu_ctx.render_device(cam, p).clone()  # sums
    out = torch.empty_like(part)
    h_out = torch.empty((h, w, 4), dtype=torch.float32).pin_memory()
    comm.reduce_begin(part, out, w, h, total, root=0, h_out=h_out)
    comm.reduce_wait()
    assert torch.equal(out, want)  # sum / spp with the division k_resolve uses: bit-identical to the unsharded frame
    assert np.array_equal(h_out.numpy(), want.cpu().numpy())
    # stream-side join, in place, un-normalised
    comm.reduce_begin(part, part, w, h, total, root=0, flags=abi.REDUCE_KEEP_SUM, slot=1)
    comm.reduce_end(slot=1)
    torch.cuda.synchronize()
    assert torch.allclose(part / total, want, rtol=1e-6, atol=0)
    # two reduces in flight in different slots; the same slot twice is refused; protocol errors are reported, not swallowed
    a, b = part.clone(), part.clone()
    comm.reduce_begin(a, a, w, h, total, slot=0)
    comm.reduce_begin(b, b, w, h, total, slot=1)
    with pytest.raises(Exception):
        comm.reduce_begin(b, b, w, h, total, slot=1)
    comm.reduce_wait(slot=0)
    comm.reduce_wait(slot=1)
    assert torch.equal(a, b)
    with pytest.raises(Exception):
        comm.reduce_wait(slot=0)
    comm.close()
