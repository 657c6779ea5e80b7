"""HostAccel: the acceleration structure exactly as mpt_scene_upload lays it out in HBM, built on the host (no device needed).

Thin ctypes face of mpt_host_accel_build / _get / _free (include/mpt.h). Used by tests of the BVH8 builder and handed to the CPU
oracle's BVH8 walker, which counts nodes and triangles per ray on the SAME tree the device walks."""
import ctypes as C

import numpy as np

from . import abi
from ._lib import MptError, lib

NODE_DTYPE = np.dtype([("p", "<f4", 3), ("e", "u1", 3), ("imask", "u1"), ("child_base", "<u4"), ("prim_base", "<u4"), ("leaf_bits", "<u4"), ("reserved", "<u4"),
                       ("qlo", "u1", (3, 8)), ("qhi", "u1", (3, 8))])
assert NODE_DTYPE.itemsize == 80


class HostAccel:
    def __init__(self, tables):
        self.l = lib()
        self.h = C.c_void_p()
        self._desc = tables.scene_desc()  # keeps the host arrays alive
        st = self.l.mpt_host_accel_build(C.byref(self._desc), C.byref(self.h))
        if st != abi.MPT_OK:
            raise MptError(st, self.l.mpt_last_error(None).decode())
        self.view = abi.HostAccelView()
        st = self.l.mpt_host_accel_get(self.h, C.byref(self.view))
        if st != abi.MPT_OK:
            raise MptError(st, self.l.mpt_last_error(None).decode())

    def _array(self, ptr, count, dtype):
        if not count:
            return np.zeros(0, dtype)
        buf = (C.c_char * (int(count) * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=int(count))

    @property
    def nodes(self):
        return self._array(self.view.nodes, self.view.node_count, NODE_DTYPE)

    @property
    def tris(self):  # (n, 3, 4) float32: v0|primitive_id, v1|geometry_index, v2|flags
        return self._array(self.view.tris, self.view.triangle_count * 12, np.float32).reshape(-1, 3, 4)

    @property
    def prim_to_tri(self):
        return self._array(self.view.prim_to_tri, self.view.prim_count, np.uint32)

    def close(self):
        if getattr(self, "h", None):
            self.l.mpt_host_accel_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
