"""The node's host side (include/mpt_node.h, merian_b200.nodes.RenderPT) against the reference's behaviour:
property keys/defaults (render_pt.cpp:147-187, render_pt.hpp:53-67), JSON load semantics
(properties_json_load.cpp:126-180 — the mechanism behind `--spp`, `--max-path-length`, tests/test_graph_load_store.cpp),
connectors (render_pt.cpp:26-34), events (render_pt.cpp:102-105). CPU part needs no device; process() is -m gpu."""
import json
import os

import numpy as np
import pytest

from merian_b200 import abi, nodes
from merian_b200._lib import MptError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULTS = {"0": True, "1": True, "2": True, "3": True, "4": True, "5": True, "6": True, "7": True, "demodulate albedo": False,
            "emission on primary": True, "irradiance format": "R32G32B32A32Sfloat", "max path length": 5, "next event estimation": False,
            "samples per pixel": 1,
            "shader execution reordering": False}


def test_defaults_match_the_reference_node():
    n = nodes.RenderPT()
    assert n.dump_config() == DEFAULTS
    assert list(n.dump_config().keys()) == sorted(DEFAULTS)  # nlohmann::json objects are key-sorted
    assert n.describe_inputs() == [("scene", "PtrIn<Scene>"), ("gbuffer", "ShaderObjectIn<GBufferObject>")]
    assert n.describe_outputs((640, 360)) == [("irradiance", "R32G32B32A32Sfloat", (640, 360))]


def test_loads_the_reference_renderer_fragment():
    """The "properties" block of examples/renderers/pt.json loads unchanged (the fragment our pt_cuda.json copies)."""
    frag = json.load(open(os.path.join(ROOT, "adapter", "pt_cuda.json")))
    props = frag["nodes"]["render"]["properties"]
    assert frag["nodes"]["render"]["type"] == nodes.TYPE_NAME
    assert frag["nodes"]["render"]["outputs"] == ["irradiance->accumulate.src"]
    assert frag["cli"]["spp"]["pointer"] == "/nodes/render/properties/samples per pixel"
    assert frag["cli"]["max-path-length"]["pointer"] == "/nodes/render/properties/max path length"
    n = nodes.RenderPT()
    assert n.load_config(props) == 0
    d = n.dump_config()
    assert d["samples per pixel"] == 2 and d["max path length"] == 5 and d["emission on primary"] is True


def test_load_semantics():
    n = nodes.RenderPT()
    # missing keys keep their value, unknown keys are ignored, UI ranges (1..16) are not applied on load
    assert n.load_config({"samples per pixel": 64, "some other node's key": 3, "nested": {"a": [1, 2, {"b": "}"}]}}) == 0
    assert n.dump_config()["samples per pixel"] == 64 and n.dump_config()["max path length"] == 5
    n.load_config({"3": False, "7": False, "demodulate albedo": True, "shader execution reordering": True})
    d = n.dump_config()
    assert (d["3"], d["7"], d["0"], d["demodulate albedo"], d["shader execution reordering"]) == (False, False, True, True, True)
    # the irradiance format is the one property that needs a reconnect (render_pt.cpp:180-185)
    assert n.load_config({"irradiance format": "R16G16B16A16Sfloat"}) == nodes.NEEDS_RECONNECT
    assert n.load_config({"irradiance format": "R16G16B16A16Sfloat"}) == 0
    assert n.describe_outputs((8, 4)) == [("irradiance", "R16G16B16A16Sfloat", (8, 4))]
    # wrong JSON types are errors and leave the node untouched (nlohmann get<T>() throws in the reference)
    before = n.dump_config()
    for bad in ({"samples per pixel": "many"}, {"emission on primary": 1}, {"0": "yes"}, {"irradiance format": 5}):
        with pytest.raises(MptError) as e:
            n.load_config(bad)
        assert e.value.status == abi.MPT_ERR_INVALID_ARGUMENT
    with pytest.raises(MptError) as e:
        n.load_config({"irradiance format": "B10G11R11UfloatPack32"})
    assert e.value.status == abi.MPT_ERR_UNSUPPORTED
    assert n.dump_config() == before
    # values no frame can be rendered with, and numbers an int32 cannot hold, are refused at load time (not later in process())
    for bad in ({"samples per pixel": -1}, {"max path length": 0}, {"samples per pixel": 1e12}, {"max path length": -3e9}):
        with pytest.raises(MptError) as e:
            n.load_config(bad)
        assert e.value.status == abi.MPT_ERR_INVALID_ARGUMENT
    assert n.dump_config() == before
    # round trip
    m = nodes.RenderPT()
    m.load_config(n.dump_config())
    assert m.dump_config() == n.dump_config()


def test_process_without_device_context_fails_loudly():
    import ctypes as C
    n = nodes.RenderPT()
    cam = abi.Camera()
    ev = C.c_uint32()
    buf = (C.c_float * 4)()
    st = n.l.mptn_process(n.h, C.byref(cam), 0, None, None, buf, None, C.byref(ev))
    assert st == abi.MPT_ERR_NOT_READY


@pytest.mark.gpu
def test_process_matches_mpt_render(gpu_ctx):
    import torch
    from merian_b200 import synthetic
    from merian_b200.scene import SceneTables, upload_scene
    from helpers import render_params
    t = SceneTables(synthetic.cornell())
    upload_scene(gpu_ctx, t)
    cam = t.desc.camera.to_abi(0)
    n = nodes.RenderPT(gpu_ctx)
    n.load_config({"samples per pixel": 8, "max path length": 4})
    n.describe_outputs((64, 48))
    hits = gpu_ctx.primary_visibility_device(cam, 64, 48)
    out, ev = n.process(cam, 7, gbuffer_hits=hits)
    assert ev == nodes.EVENT_BOUNCES_CHANGED  # 5 -> 4 (render_pt.cpp:102-105), sent once
    out2, ev2 = n.process(cam, 7, gbuffer_hits=hits)
    assert ev2 == 0 and torch.equal(out, out2)
    ref = gpu_ctx.render(cam, render_params(64, 48, 8, L=4, frame=7))
    assert np.array_equal(out.cpu().numpy(), ref)
    # instance mask bits: nothing is visible to an empty mask -> no scattering, primary emission only
    n.load_config({str(b): False for b in range(8)})
    dark, _ = n.process(cam, 7, gbuffer_hits=hits)
    assert float(dark[..., :3].sum()) <= float(out[..., :3].sum())
    n.load_config({str(b): True for b in range(8)})
    # fp16 output format
    assert n.load_config({"irradiance format": "R16G16B16A16Sfloat"}) == nodes.NEEDS_RECONNECT
    n.describe_outputs((64, 48))
    h16, _ = n.process(cam, 7, gbuffer_hits=hits)
    assert h16.dtype == torch.float16 and torch.equal(h16, out.to(torch.float16))
    # demodulate albedo needs the albedo texture
    n.load_config({"demodulate albedo": True})
    with pytest.raises(MptError):
        n.process(cam, 7, gbuffer_hits=hits)
    # sharded node: sums of the shards = the frame
    n.load_config({"demodulate albedo": False, "irradiance format": "R32G32B32A32Sfloat", "samples per pixel": 9})
    n.describe_outputs((64, 48))
    full, _ = n.process(cam, 3, gbuffer_hits=hits)
    acc = torch.zeros_like(full)
    for r in range(2):
        n.set_shard(r, 2)
        part, _ = n.process(cam, 3, gbuffer_hits=hits)
        acc += part
    n.set_shard(0, 1)
    assert torch.allclose(acc / 9.0, full, rtol=2e-5, atol=1e-6)
    # scene not ready: process is a no-op returning OK (render_pt.cpp:98-99)
    from merian_b200.context import Context
    fresh = Context(0)
    m = nodes.RenderPT(fresh)
    m.describe_outputs((8, 8))
    sentinel = torch.full((8, 8, 4), 3.0, device=fresh.torch_device)
    o, _ = m.process(cam, 0, out=sentinel)
    assert float(o.min()) == 3.0
    fresh.close()
