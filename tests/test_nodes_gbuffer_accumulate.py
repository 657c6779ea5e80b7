"""Host side of the GBuffer (Raytraced) and Accumulate nodes (include/mpt_node.h) against the reference's property keys and defaults
(src/merian-graph/nodes/gbuffer_rt/gbuffer.cpp:155-181, include/.../gbuffer.hpp:64-70; src/merian-graph/nodes/accumulate/accumulate.cpp:172-250,
include/.../accumulate.hpp:32-36), and the node chain gbuffer -> pt -> accumulate on the device."""
import math

import numpy as np
import pytest

from merian_b200 import abi, nodes
from merian_b200._lib import MptError

from helpers import tables


def test_gbuffer_node_config():
    n = nodes.GBufferRT()
    d = n.dump_config()
    assert d == {**{str(i): True for i in range(8)}, "emission format": "R32G32B32A32Sfloat", "height": 1080, "width": 1920}
    assert list(d) == sorted(d)
    assert n.describe_inputs() == [("scene", "PtrIn<Scene>")]
    assert n.describe_outputs() == [("gbuffer", "GBufferObject", (1920, 1080)), ("emission", "R32G32B32A32Sfloat", (1920, 1080))]
    assert n.load_config({"width": 640, "height": 360, "5": False, "unknown": [1, 2]}) == nodes.NEEDS_RECONNECT  # gbuffer.cpp:160-161,176-178
    assert n.load_config({"width": 640}) == 0
    assert n.dump_config()["5"] is False and n.describe_outputs()[0][2] == (640, 360)
    assert n.set_resolution(640, 360) == 0 and n.set_resolution(800, 600) == nodes.NEEDS_RECONNECT          # pre_process, gbuffer.cpp:104-115
    for bad in ({"width": "wide"}, {"0": 1}, {"width": 0}):
        with pytest.raises(MptError):
            n.load_config(bad)
    with pytest.raises(MptError) as e:
        n.load_config({"emission format": "R16G16B16A16Sfloat"})
    assert e.value.status == abi.MPT_ERR_UNSUPPORTED


def test_accumulate_node_config():
    n = nodes.Accumulate()
    d = n.dump_config()
    assert list(d) == sorted(d)
    # defaults of accumulate.hpp:32-36: alpha 0, infinite history, no hard clamp, normal_reject_cos 0.8, motion vectors on
    assert d["alpha"] == 0 and d["inf history"] is True and d["inf clamp"] is True and d["use motion vectors"] is True and d["gbuffer check"] == "always"
    assert d["normal threshold"] == pytest.approx(math.acos(0.8), rel=1e-6) and d["firefly filter enable"] is False and d["f32 history"] is False
    # examples/gltf.json:46-67 style block: alpha 1, infinite history, firefly filter off
    st = n.load_config({"alpha": 1.0, "inf history": True, "firefly filter enable": False, "use motion vectors": False, "gbuffer check": "never",
                        "hard clamp": 100.0, "max history": 64.0, "extended search": False})
    assert st == nodes.NEEDS_RECONNECT   # "use motion vectors" / "gbuffer check" / "extended search" rebuild the pipeline (accumulate.cpp:188-213)
    d = n.dump_config()
    assert d["alpha"] == 1.0 and d["inf history"] is True and d["hard clamp"] == 100.0 and d["inf clamp"] is False and d["gbuffer check"] == "never"
    n.load_config({"inf history": False, "max history": 64.0, "inf clamp": True})
    d = n.dump_config()
    assert d["inf history"] is False and d["max history"] == 64.0 and d["inf clamp"] is True
    m = nodes.Accumulate()
    m.load_config(d)
    assert m.dump_config() == d
    for bad in ({"alpha": 2.0}, {"alpha": "x"}, {"clear": 1}, {"gbuffer check": 3}):
        with pytest.raises(MptError):
            n.load_config(bad)


@pytest.mark.gpu
def test_node_chain_gbuffer_pt_accumulate(gpu_ctx):
    """Three iterations of gbuffer -> pt -> accumulate through the node objects, fused and unfused, equal the direct C-ABI calls."""
    import torch
    from merian_b200.scene import upload_scene
    from oracle.oracle_py import accumulate as oracle_accumulate
    from test_accumulate import acc_params
    t = tables("cornell")
    upload_scene(gpu_ctx, t)
    W, H = 80, 60
    gb = nodes.GBufferRT(gpu_ctx)
    gb.load_config({"width": W, "height": H})
    pt = nodes.RenderPT(gpu_ctx)
    pt.load_config({"samples per pixel": 4, "max path length": 4})
    pt.describe_outputs((W, H))
    acc_a, acc_b = nodes.Accumulate(gpu_ctx), nodes.Accumulate(gpu_ctx)
    props = {"alpha": 1.0, "inf history": True, "use motion vectors": False, "gbuffer check": "always", "normal threshold": 0.6, "depth threshold": 0.05}
    acc_a.load_config(props)
    acc_b.load_config(props)
    with pytest.raises(MptError) as e:      # the reference default (motion-vector reprojection) is not what this node does: it says so
        nodes.Accumulate(gpu_ctx).process(0, torch.zeros((H, W, 4), device=gpu_ctx.torch_device))
    assert e.value.status == abi.MPT_ERR_UNSUPPORTED
    ref_out = ref_hist = None
    for it in range(3):
        cam = t.desc.camera.to_abi(it)       # R2 jitter would apply for glTF cameras; the synthetic camera has none
        g = gb.process(cam)
        assert torch.equal(g["tex1"], gpu_ctx.primary_visibility_device(cam, W, H))
        irr, _ = pt.process(cam, it, gbuffer_hits=g["tex1"])
        out_a, hist_a, ev_a = acc_a.process(it, irr, g)
        out_b, hist_b, ev_b = acc_b.process_fused(pt, cam, it, it, g)
        assert ev_a == (nodes.EVENT_CLEAR if it == 0 else 0) and (ev_b & nodes.EVENT_CLEAR) == ev_a
        assert torch.equal(out_a, out_b) and torch.equal(hist_a, hist_b)
        A = acc_params(W, H, clear=int(it == 0), flags=abi.ACC_GBUFFER_CHECK, normal_reject_cos=math.cos(np.float32(0.6)), z_reject_percent=0.05)
        A.normal_reject_cos = np.float32(np.cos(np.float32(0.6)))
        ref_out, ref_hist = oracle_accumulate(A, irr.cpu().numpy(), ref_out, ref_hist, g["tex0"].cpu().numpy().view(np.uint32), g["tex2"].cpu().numpy())
        assert np.array_equal(hist_a.cpu().numpy().view(np.uint32), ref_hist)
        np.testing.assert_allclose(out_a.cpu().numpy(), ref_out, rtol=1e-6, atol=1e-7)
    assert ((hist_a.cpu().numpy().view(np.uint32)[..., 1] >> 16) == int(np.float16(3.0).view(np.uint16))).all()  # static view: every pixel kept its history
    acc_a.request_clear()
    _, hist_a, ev = acc_a.process(3, irr, g)
    assert ev == nodes.EVENT_CLEAR and ((hist_a.cpu().numpy().view(np.uint32)[..., 1] >> 16) == int(np.float16(1.0).view(np.uint16))).all()
