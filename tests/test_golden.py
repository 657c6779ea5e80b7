"""Committed fixtures (tests/golden, made by tests/golden/make_golden.py): the oracle must keep reproducing them on
the CPU, and the CUDA path must match them on the GPU."""
import os

import numpy as np
import pytest

from merian_b200 import abi, synthetic
from merian_b200.scene import SceneTables, upload_scene

from helpers import compare_hits, edge_ties

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cornell():
    t = SceneTables(synthetic.cornell())
    return t, t.desc.camera.to_abi(0), abi.RenderParams(48, 48, 0, 8, 5, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0)


def test_rng_golden(oracle_lib):
    import ctypes as C
    g = np.load(os.path.join(G, "rng.npz"))
    for row, seed, fl in zip(g["inputs"], g["pcg4d16"], g["floats"]):
        assert oracle_lib.oracle_pcg4d16(*[int(x) for x in row]) == int(seed)
        u = np.zeros(8, np.uint32)
        f = np.zeros(8, np.float32)
        oracle_lib.oracle_pcg_stream(C.c_uint32(int(seed)), C.c_uint32(8), C.c_void_p(u.ctypes.data), C.c_void_p(f.ctypes.data))
        assert (f == fl).all()


def test_oracle_reproduces_golden(oracle_lib):
    from oracle.oracle_py import OracleContext
    g = np.load(os.path.join(G, "cornell.npz"))
    t, cam, P = _cornell()
    o = OracleContext()
    upload_scene(o, t)
    assert (o.primary_visibility(cam, 64, 64).view(np.uint32).reshape(64, 64, 4) == g["tex1"]).all()
    assert np.allclose(o.render(cam, P), g["frame"], rtol=1e-5, atol=1e-6)  # libm differences only
    g = np.load(os.path.join(G, "terrain.npz"))
    t = SceneTables(synthetic.terrain(nx=40, ny=20, env_size=(64, 32)))
    o = OracleContext()
    upload_scene(o, t)
    assert (o.primary_visibility(t.desc.camera.to_abi(0), 96, 54).view(np.uint32).reshape(54, 96, 4) == g["tex1"]).all()


@pytest.mark.gpu
def test_cuda_matches_golden(gpu_ctx):
    g = np.load(os.path.join(G, "cornell.npz"))
    t, cam, P = _cornell()
    upload_scene(gpu_ctx, t)
    ref = np.ascontiguousarray(g["tex1"]).view(abi.PACKED_HIT_DTYPE).reshape(64, 64)
    got = gpu_ctx.primary_visibility(cam, 64, 64)
    mism, bary_err, same = compare_hits(got, ref)
    n_mis, n_ties = edge_ties(got, ref, same)
    assert n_mis == n_ties and mism < 1e-4 and bary_err == 0.0
    img = gpu_ctx.render(cam, P)
    d = np.abs(img[..., :3] - g["frame"][..., :3]).max(axis=-1) / np.maximum(np.abs(g["frame"][..., :3]).max(axis=-1), 1e-3)
    assert (d > 1e-3).mean() < 0.05
    assert abs(img[..., :3].mean() - g["frame"][..., :3].mean()) / g["frame"][..., :3].mean() < 5e-3
    g = np.load(os.path.join(G, "terrain.npz"))
    t = SceneTables(synthetic.terrain(nx=40, ny=20, env_size=(64, 32)))
    upload_scene(gpu_ctx, t)
    ref = np.ascontiguousarray(g["tex1"]).view(abi.PACKED_HIT_DTYPE).reshape(54, 96)
    got = gpu_ctx.primary_visibility(t.desc.camera.to_abi(0), 96, 54)
    mism, bary_err, same = compare_hits(got, ref)
    n_mis, n_ties = edge_ties(got, ref, same)
    assert n_mis == n_ties and mism < 1e-4 and bary_err == 0.0
