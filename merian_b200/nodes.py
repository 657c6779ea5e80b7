"""Python mirror of merian-graph's `Render (Path-traced)` node (host side).

Same surface as merian::RenderPT (include/merian-graph/nodes/render_pt/render_pt.hpp:26-87): describe_inputs /
describe_outputs / properties (JSON load + dump with the reference's keys) / process. All logic lives in libmpt.so
(csrc/node.cu, include/mpt_node.h); this class only marshals. The merian C++ plugin (adapter/) binds the same entry
points.
"""
import ctypes as C
import json

from . import abi
from ._lib import MptError, lib

TYPE_NAME = "Render (Path-traced, CUDA)"  # the built-in name cannot be registered twice (node_registry.cpp:95)
NEEDS_RECONNECT, RESET_IN_FLIGHT_DATA, REMOVE_NODE = 0x1, 0x2, 0x4
EVENT_BOUNCES_CHANGED = 0x1
VK_FORMAT = {97: "R16G16B16A16Sfloat", 109: "R32G32B32A32Sfloat"}


class RenderPT:
    def __init__(self, context=None):
        """`context`: merian_b200.Context (initialize(context, allocator) of the reference) or None for a node that
        only carries configuration."""
        self.l = lib()
        self.context = context
        self.h = C.c_void_p()
        st = self.l.mptn_create(context.h if context is not None else None, C.byref(self.h))
        if st != abi.MPT_OK:
            raise MptError(st, "mptn_create failed")
        self.extent = (1920, 1080)

    def close(self):
        if getattr(self, "h", None):
            self.l.mptn_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != abi.MPT_OK:
            raise MptError(st, self.l.mptn_last_error(self.h).decode())

    # ---- Node interface
    @staticmethod
    def describe_inputs():
        """render_pt.cpp:26-28."""
        return [("scene", "PtrIn<Scene>"), ("gbuffer", "ShaderObjectIn<GBufferObject>")]

    def describe_outputs(self, gbuffer_extent):
        """render_pt.cpp:30-34: [("irradiance", format, extent)] with the GBuffer's extent."""
        fmt, w, h = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._check(self.l.mptn_describe_outputs(self.h, gbuffer_extent[0], gbuffer_extent[1], C.byref(fmt), C.byref(w), C.byref(h)))
        self.extent = (w.value, h.value)
        return [("irradiance", VK_FORMAT[fmt.value], self.extent)]

    def load_config(self, properties: dict) -> int:
        """Node::load_config (node.cpp:11-16). Returns the NodeStatusFlags (NEEDS_RECONNECT when the format changed)."""
        st = C.c_uint32()
        self._check(self.l.mptn_load_config(self.h, json.dumps(properties).encode(), C.byref(st)))
        return st.value

    def dump_config(self) -> dict:
        """Node::dump_config (node.cpp:17-21)."""
        n = C.c_size_t()
        self._check(self.l.mptn_dump_config(self.h, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value + 1)
        self._check(self.l.mptn_dump_config(self.h, buf, n.value + 1, C.byref(n)))
        return json.loads(buf.value.decode())

    def set_shard(self, rank, world):
        self._check(self.l.mptn_set_shard(self.h, rank, world))

    def process(self, camera, frame, gbuffer_hits=None, gbuffer_albedo=None, out=None):
        """RenderPT::process (render_pt.cpp:94-124) on torch CUDA tensors. Returns (irradiance tensor, events)."""
        import torch
        w, h = self.extent
        fmt = self.dump_config()["irradiance format"]
        if out is None:
            out = torch.empty((h, w, 4), dtype=torch.float16 if fmt == "R16G16B16A16Sfloat" else torch.float32, device=self.context.torch_device)
        ev = C.c_uint32()
        self._check(self.l.mptn_process(self.h, C.byref(camera), frame, C.c_void_p(gbuffer_hits.data_ptr()) if gbuffer_hits is not None else None,
                                        C.c_void_p(gbuffer_albedo.data_ptr()) if gbuffer_albedo is not None else None, C.c_void_p(out.data_ptr()),
                                        self.context._stream(), C.byref(ev)))
        return out, ev.value


class _ConfigNode:
    """Shared marshalling for the mptn_<kind>_* objects of include/mpt_node.h."""
    KIND = ""

    def __init__(self, context=None):
        self.l = lib()
        self.context = context
        self.h = C.c_void_p()
        st = getattr(self.l, f"mptn_{self.KIND}_create")(context.h if context is not None else None, C.byref(self.h))
        if st != abi.MPT_OK:
            raise MptError(st, f"mptn_{self.KIND}_create failed")

    def close(self):
        if getattr(self, "h", None):
            getattr(self.l, f"mptn_{self.KIND}_destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != abi.MPT_OK:
            raise MptError(st, getattr(self.l, f"mptn_{self.KIND}_last_error")(self.h).decode())

    def load_config(self, properties: dict) -> int:
        st = C.c_uint32()
        self._check(getattr(self.l, f"mptn_{self.KIND}_load_config")(self.h, json.dumps(properties).encode(), C.byref(st)))
        return st.value

    def dump_config(self) -> dict:
        n = C.c_size_t()
        f = getattr(self.l, f"mptn_{self.KIND}_dump_config")
        self._check(f(self.h, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value + 1)
        self._check(f(self.h, buf, n.value + 1, C.byref(n)))
        return json.loads(buf.value.decode())


class GBufferRT(_ConfigNode):
    """merian::GBufferRTNode (include/merian-graph/nodes/gbuffer_rt/gbuffer.hpp, src/.../gbuffer_rt/gbuffer.cpp)."""
    KIND = "gbuffer"
    TYPE_NAME = "GBuffer (Raytraced, CUDA)"

    @staticmethod
    def describe_inputs():
        return [("scene", "PtrIn<Scene>")]  # gbuffer.cpp:29-31

    def describe_outputs(self):
        """gbuffer.cpp:33-42: [("gbuffer", extent), ("emission", format, extent)]."""
        w, h, fmt = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._check(self.l.mptn_gbuffer_describe_outputs(self.h, C.byref(w), C.byref(h), C.byref(fmt)))
        return [("gbuffer", "GBufferObject", (w.value, h.value)), ("emission", VK_FORMAT[fmt.value], (w.value, h.value))]

    def set_resolution(self, width, height) -> int:
        """pre_process: the scene camera's resolution wins over the properties (gbuffer.cpp:104-115)."""
        st = C.c_uint32()
        self._check(self.l.mptn_gbuffer_set_resolution(self.h, width, height, C.byref(st)))
        return st.value

    def process(self, camera, prev_camera=None):
        """Returns the dict of device tensors Context.gbuffer_device returns (tex0..tex3, emission)."""
        import torch
        (_, _, (w, h)), _ = self.describe_outputs()
        dev = self.context.torch_device
        g = {"tex0": torch.empty((h, w, 2), dtype=torch.int32, device=dev), "tex1": torch.empty((h, w, 4), dtype=torch.int32, device=dev),
             "tex2": torch.empty((h, w, 4), dtype=torch.float32, device=dev), "tex3": torch.empty((h, w, 4), dtype=torch.float32, device=dev),
             "emission": torch.empty((h, w, 4), dtype=torch.float32, device=dev)}
        imgs = abi.GBufferImages(*[C.c_void_p(g[k].data_ptr()) for k in ("tex0", "tex1", "tex2", "tex3", "emission")])
        self._check(self.l.mptn_gbuffer_process(self.h, C.byref(camera), C.byref(prev_camera) if prev_camera is not None else None, C.byref(imgs),
                                                self.context._stream()))
        return g


EVENT_CLEAR = 0x2


class Accumulate(_ConfigNode):
    """merian::Accumulate, plain mode (src/merian-graph/nodes/accumulate/accumulate.cpp). Owns the feedback images (out, history)."""
    KIND = "accumulate"
    TYPE_NAME = "Accumulate (CUDA)"

    def __init__(self, context=None):
        super().__init__(context)
        self.out = self.history = None

    @staticmethod
    def describe_inputs():
        return [("src", "VkImageIn"), ("gbuffer", "ShaderObjectIn<GBufferObject>"), ("prev_out", "VkImageIn, delay 1"), ("prev_history", "VkImageIn, delay 1")]

    def request_clear(self):
        self._check(self.l.mptn_accumulate_request_clear(self.h))

    def _buffers(self, w, h, gbuffer):
        import torch
        if self.out is None or tuple(self.out.shape[:2]) != (h, w):
            self.out = torch.zeros((h, w, 4), dtype=torch.float32, device=self.context.torch_device)
            self.history = torch.zeros((h, w, 2), dtype=torch.int32, device=self.context.torch_device)
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        g = gbuffer or {}
        return abi.AccumulateBuffers(p(self.out), p(self.history), p(self.out), p(self.history), p(g.get("tex0")), p(g.get("tex2")))

    def process(self, iteration, src, gbuffer=None):
        """Accumulate::process: returns (out, history, events). `gbuffer`: the dict of GBufferRT.process (tex0 / tex2 are read)."""
        h, w = src.shape[:2]
        b = self._buffers(w, h, gbuffer)
        ev = C.c_uint32()
        self._check(self.l.mptn_accumulate_process(self.h, iteration, w, h, C.c_void_p(src.data_ptr()), C.byref(b), self.context._stream(), C.byref(ev)))
        return self.out, self.history, ev.value

    def process_fused(self, pt: "RenderPT", camera, frame, iteration, gbuffer):
        """pt + accumulate in one call (mptn_process_accumulate): the irradiance image between the two nodes is never materialised."""
        w, h = pt.extent
        b = self._buffers(w, h, gbuffer)
        ev = C.c_uint32()
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        self._check(self.l.mptn_process_accumulate(pt.h, self.h, C.byref(camera), frame, iteration, p(gbuffer.get("tex1")), p(gbuffer.get("tex3")),
                                                   C.byref(b), self.context._stream(), C.byref(ev)))
        return self.out, self.history, ev.value
