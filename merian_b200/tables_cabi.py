"""CTables: the scene tables built by libmpt's own C++ restatement of the merian::Scene host logic (include/mpt_tables.h, mptt_*),
from the same SceneDescription the numpy mirror (scene.SceneTables) takes. What a C or C++ host without a merian::Scene would call;
here it exists so that the two restatements can be compared (tests/test_tables_cabi.py) and so that Python callers can skip the numpy
packing of large meshes."""
import ctypes as C

import numpy as np

from . import abi
from ._lib import MptError, lib


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


class CTables:
    def __init__(self, desc):
        self.l = lib()
        self.desc = desc
        keep = []
        meshes = (abi.SrcMesh * max(len(desc.meshes), 1))()
        for i, m in enumerate(desc.meshes):
            pos = np.ascontiguousarray(m.positions, np.float32)
            nrm = np.ascontiguousarray(m.normals, np.float32)
            uvs = None if m.uvs is None else np.ascontiguousarray(m.uvs, np.float32)
            tan = None if m.tangents is None else np.ascontiguousarray(m.tangents, np.float32)
            idx = None if m.indices is None else np.ascontiguousarray(m.indices)
            keep += [pos, nrm, uvs, tan, idx]
            s = meshes[i]
            s.positions, s.normals, s.uvs, s.tangents, s.vertex_count = _ptr(pos), _ptr(nrm), _ptr(uvs), _ptr(tan), len(pos)
            s.indices = _ptr(idx)
            s.primitive_count = len(idx) if idx is not None else len(pos) // 3
            s.index_size = idx.dtype.itemsize if idx is not None else 0
            s.material_id, s.flags, s.instance_mask = m.material_id, m.flags, m.instance_mask
        nodes = (abi.SrcNode * max(len(desc.nodes), 1))()
        for i, n in enumerate(desc.nodes):
            t = np.ascontiguousarray(n.transform, np.float32).reshape(16)
            ids = np.ascontiguousarray(n.meshes, np.uint32)
            keep.append(ids)
            nodes[i].transform[:] = t.tolist()
            nodes[i].meshes, nodes[i].mesh_count = _ptr(ids), len(ids)
        self.h = C.c_void_p()
        st = self.l.mptt_build(C.cast(meshes, C.c_void_p), len(desc.meshes), C.cast(nodes, C.c_void_p), len(desc.nodes), C.byref(self.h))
        if st != abi.MPT_OK:
            raise MptError(st, self.l.mpt_last_error(None).decode())
        self.sd = self.l.mptt_scene_desc(self.h).contents

    def scene_desc(self):
        return self.sd

    # ---- numpy views of the result (valid until close)
    def _arr(self, ptr, count, dtype):
        if not ptr or not count:
            return np.zeros(0, dtype)
        return np.frombuffer((C.c_char * (int(count) * np.dtype(dtype).itemsize)).from_address(ptr), dtype=dtype, count=int(count))

    def mesh_descs(self):
        return (abi.MeshDesc * self.sd.mesh_count).from_address(self.sd.meshes) if self.sd.mesh_count else []

    def mesh_vertices(self, i):
        m = self.mesh_descs()[i]
        return self._arr(m.vertices, m.vertex_count, abi.PACKED_VERTEX_DTYPE)

    @property
    def geometries(self):
        g = (abi.GeometryDesc * self.sd.geometry_count).from_address(self.sd.geometries) if self.sd.geometry_count else []
        return [(x.mesh, x.material_id, x.flags, x.primitive_count) for x in g]

    @property
    def instances(self):
        a = (abi.InstanceDesc * self.sd.instance_count).from_address(self.sd.instances) if self.sd.instance_count else []
        return [(np.array(x.transform[:], np.float32), x.instance_custom_index, x.mask, x.flags, x.blas) for x in a]

    @property
    def groups(self):
        b = (abi.BlasDesc * self.sd.blas_count).from_address(self.sd.blas) if self.sd.blas_count else []
        return [self._arr(x.meshes, x.mesh_count, np.uint32).tolist() for x in b]

    @property
    def instance_transforms(self):
        return self._arr(self.sd.instance_transforms, self.sd.instance_count * 16, np.float32).reshape(-1, 4, 4)

    @property
    def inv_t_instance_transforms(self):
        return self._arr(self.sd.inverse_transposed_instance_transforms, self.sd.instance_count * 16, np.float32).reshape(-1, 4, 4)

    def close(self):
        if getattr(self, "h", None):
            self.l.mptt_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pack_vertices(positions, normals, uvs=None, tangents=None, transform=None):
    """mptt_pack_vertices: the C++ counterpart of scene.pack_vertices."""
    pos = np.ascontiguousarray(positions, np.float32)
    nrm = np.ascontiguousarray(normals, np.float32)
    uv = None if uvs is None else np.ascontiguousarray(uvs, np.float32)
    tan = None if tangents is None else np.ascontiguousarray(tangents, np.float32)
    M = None if transform is None else np.ascontiguousarray(transform, np.float32).reshape(16)
    out = np.zeros(len(pos), abi.PACKED_VERTEX_DTYPE)
    st = lib().mptt_pack_vertices(_ptr(pos), _ptr(nrm), _ptr(uv), _ptr(tan), len(pos), _ptr(M), _ptr(out))
    if st != abi.MPT_OK:
        raise MptError(st, lib().mpt_last_error(None).decode())
    return out
