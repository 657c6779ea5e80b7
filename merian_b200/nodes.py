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
