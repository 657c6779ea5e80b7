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


def _frame_nodes_inputs():
    from merian_b200.scene import CameraDesc
    t = SceneTables(synthetic.cornell())
    cd = t.desc.camera
    prev = CameraDesc(position=tuple(np.array(cd.position) + [0.05, 0.02, -0.03]), target=cd.target, up=cd.up, yfov_deg=cd.yfov_deg, aspect=cd.aspect,
                      near=cd.near, far=cd.far).to_abi(0)
    return t, cd.to_abi(0), prev


def _acc_params(it):
    return abi.AccumulateParams(48, 48, 1.0, float("inf"), 2.0, 0.8, 0.1, int(it == 0), abi.ACC_GBUFFER_CHECK)


def test_oracle_reproduces_frame_node_golden(oracle_lib):
    """GBuffer (Raytraced) and Accumulate rows: the oracle keeps reproducing the committed texels."""
    from oracle.oracle_py import OracleContext, accumulate
    g = np.load(os.path.join(G, "frame_nodes.npz"))
    t, cam, prev = _frame_nodes_inputs()
    o = OracleContext()
    upload_scene(o, t)
    got = o.gbuffer(cam, 48, 48, prev_cam=prev)
    assert np.array_equal(got["tex1"].view(np.uint32).reshape(48, 48, 4), g["tex1"])
    assert (got["tex0"] != g["tex0"]).mean() < 1e-3  # libm differences may move an encoded normal by one step
    for k in ("tex2", "tex3", "emission"):
        np.testing.assert_allclose(got[k], g[k], rtol=1e-5, atol=1e-6)
    out = hist = None
    for it in range(3):
        frame = o.render(cam, abi.RenderParams(48, 48, it, 4, 5, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0))
        out, hist = accumulate(_acc_params(it), frame, out, hist, g["tex0"], g["tex2"])
    assert np.array_equal(hist, g["acc_history"])
    np.testing.assert_allclose(out, g["acc_out"], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_cuda_matches_frame_node_golden(gpu_ctx):
    import torch
    g = np.load(os.path.join(G, "frame_nodes.npz"))
    t, cam, prev = _frame_nodes_inputs()
    upload_scene(gpu_ctx, t)
    got = gpu_ctx.gbuffer(cam, 48, 48, prev_cam=prev)
    assert np.array_equal(got["tex1"].view(np.uint32).reshape(48, 48, 4), g["tex1"])
    a, b = got["tex0"][..., 0], g["tex0"][..., 0]
    assert (a != b).mean() < 0.02 and np.abs((a & 0xFFFF).astype(np.int16).astype(np.int32) - (b & 0xFFFF).astype(np.int16).astype(np.int32)).max() <= 1
    np.testing.assert_allclose(got["tex2"][..., 2:], g["tex2"][..., 2:], rtol=1e-5, atol=1e-4)
    assert np.abs(got["tex2"][..., :2] - g["tex2"][..., :2]).max() < 2e-3
    np.testing.assert_allclose(got["tex3"], g["tex3"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(got["emission"], g["emission"], rtol=2e-4, atol=1e-6)
    # Accumulate: three iterations of CUDA frames through the CUDA node against the committed history counters; the radiance agrees with
    # the golden image within the path-by-path tolerance of test_cuda_matches_golden
    dev = gpu_ctx.torch_device
    tex0 = torch.from_numpy(g["tex0"].view(np.int32)).to(dev)
    tex2 = torch.from_numpy(g["tex2"]).to(dev)
    out = torch.zeros((48, 48, 4), dtype=torch.float32, device=dev)
    hist = torch.zeros((48, 48, 2), dtype=torch.int32, device=dev)
    for it in range(3):
        frame = gpu_ctx.render_device(cam, abi.RenderParams(48, 48, it, 4, 5, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0))
        gpu_ctx.accumulate_device(_acc_params(it), frame, out, hist, out, hist, tex0, tex2)
    assert np.array_equal(hist.cpu().numpy().view(np.uint32), g["acc_history"])
    o = out.cpu().numpy()
    d = np.abs(o[..., :3] - g["acc_out"][..., :3]).max(-1) / np.maximum(np.abs(g["acc_out"][..., :3]).max(-1), 1e-3)
    assert (d > 1e-3).mean() < 0.05
