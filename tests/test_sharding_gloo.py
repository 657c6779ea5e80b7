"""N>1 path on CPU: two gloo processes shard the sample indices of one frame, reduce their sum buffers and must
reproduce the single-process frame (same set of paths, different fp32 summation order). The renderer behind the host
logic is the CPU oracle (allowed in tests); on the GPU box the same code runs with Context.render_device + NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from merian_b200 import abi, sharding, synthetic
from merian_b200.scene import SceneTables, upload_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, SPP, L = 48, 48, 7, 5  # odd spp: uneven shards (4 + 3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.oracle_py import OracleContext
        t = SceneTables(synthetic.cornell())
        o = OracleContext()
        upload_scene(o, t)
        cam = t.desc.camera.to_abi(0)
        P = abi.RenderParams(W, H, 3, SPP, L, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0)
        img = sharding.render_sharded(lambda c, p: torch.from_numpy(o.render(c, p)), cam, P, rank, world)
        if rank == 0:
            ret["img"] = img.numpy().copy()
    finally:
        dist.destroy_process_group()


def test_shard_counts_partition_all_samples():
    for spp in (0, 1, 7, 64, 1024):
        for world in (1, 2, 3, 8):
            c = sharding.shard_counts(spp, world)
            assert sum(c) == spp and max(c) - min(c) <= 1
            idx = sorted(r + i * world for r in range(world) for i in range(c[r]))
            assert idx == list(range(spp))  # every sample index exactly once


def test_shard_params_validate():
    P = abi.RenderParams(8, 8, 0, 10, 5, 0xFF, 0, 0, 1, 0.0)
    p, total = sharding.shard_params(P, 1, 4)
    assert (p.spp, p.sample_offset, p.sample_stride, total) == (3, 1, 4, 10) and p.flags & abi.RENDER_OUTPUT_SUM
    with pytest.raises(ValueError):
        sharding.shard_params(P, 4, 4)
    with pytest.raises(ValueError):
        sharding.shard_params(p, 0, 2)  # already sharded


@pytest.mark.timeout(300)
def test_two_ranks_reproduce_single_process_frame(oracle_lib):
    from oracle.oracle_py import OracleContext
    t = SceneTables(synthetic.cornell())
    o = OracleContext()
    upload_scene(o, t)
    ref = o.render(t.desc.camera.to_abi(0), abi.RenderParams(W, H, 3, SPP, L, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0))
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
        img = ret["img"]
    assert img.shape == ref.shape
    # identical sample set; only the order of the fp32 additions differs
    assert np.allclose(img[..., :3], ref[..., :3], rtol=1e-5, atol=1e-6)
    assert np.allclose(img[..., 3], ref[..., 3], rtol=1e-4, atol=1e-6)
