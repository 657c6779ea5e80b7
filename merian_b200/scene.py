"""Host-side scene tables: the part of merian::Scene that defines what the renderer sees.

Restates, in numpy, the host logic the reference runs before `pt`:
  * vertex packing            src/merian-shaders/scene/gltf_scene.cpp:107-118, src/merian/utils/normal_encoding.cpp:9-25
  * the GPU pre-transform     src/merian-shaders/scene/transform_vertex.slang:12-40 (always runs, even with identity)
  * mesh grouping -> BLAS     src/merian-shaders/scene/scene.cpp:1023-1105
  * GeometryData order        src/merian-shaders/scene/scene.cpp:1236-1291
  * TLAS instances / ids      src/merian-shaders/scene/scene.cpp:1808-1857
  * material buffer entries   src/merian-shaders/shading/materials/material_system.cpp:140-195,
                              include/merian-shaders/shading/materials/gltf_material.hpp:93-146
  * camera basis              src/merian/utils/camera/camera.cpp:263-266,296-298,318-368

The result (`SceneTables`) is what `mpt_scene_upload` / `mpt_materials_upload` consume; a merian node adapter
would assemble the same tables from `Scene::get_mesh_infos()` and the export shader.
"""
from __future__ import annotations

import ctypes as C
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import abi

# merian::Scene::MeshFlags, include/merian-shaders/scene/scene.hpp:39-58
MESH_IS_MORPHED = 0x1
MESH_HAS_VARIABLE_TOPOLOGY = 0x2
MESH_IS_OPAQUE = 0x4
MESH_FLIP_FACING = 0x8
MESH_TWO_SIDED = 0x10
MESH_HAS_TANGENTS = 0x20
MESH_FLAT_SHADING = 0x40
MESH_USE_ENV_MAP = 0x80
GROUP_SPLIT_MASK = MESH_FLIP_FACING | MESH_TWO_SIDED  # scene.hpp:350-351


# ----------------------------------------------------------------------------------------------- encodings
def _oct_project(n: np.ndarray) -> np.ndarray:
    """octahedral_project, include/merian-shaders/utils/encoding.slang:17-23 (fp32)."""
    n = n.astype(np.float32)
    inv = np.float32(1.0) / ((np.abs(n[:, 0]) + np.abs(n[:, 1])) + np.abs(n[:, 2]))
    px = n[:, 0] * inv
    py = n[:, 1] * inv
    neg = n[:, 2] < 0
    fx = np.copysign(np.float32(1.0) - np.abs(n[:, 1] * inv), n[:, 0])
    fy = np.copysign(np.float32(1.0) - np.abs(n[:, 0] * inv), n[:, 1])
    return np.stack([np.where(neg, fx, px), np.where(neg, fy, py)], axis=1).astype(np.float32)


def host_encode_normal(n: np.ndarray) -> np.ndarray:
    """merian::encode_normal on the host, normal_encoding.cpp:9-25: x32768, roundf, clamp to int16."""
    p = _oct_project(n)
    e = np.clip(p * np.float32(32768.0), -32768.0, 32767.0)
    # roundf = half away from zero
    r = np.where(e >= 0, np.floor(e + np.float32(0.5)), np.ceil(e - np.float32(0.5))).astype(np.int32)
    return ((r[:, 0] & 0xFFFF) | ((r[:, 1] & 0xFFFF) << 16)).astype(np.uint32)


def device_decode_normal(enc: np.ndarray) -> np.ndarray:
    """decode_normal, encoding.slang:6-13 (unpackSnorm2x16: int16/32767 clamped)."""
    a = (enc & 0xFFFF).astype(np.uint16).view(np.int16).astype(np.float32)
    b = (enc >> 16).astype(np.uint16).view(np.int16).astype(np.float32)
    x = np.clip(a / np.float32(32767.0), -1, 1).astype(np.float32)
    y = np.clip(b / np.float32(32767.0), -1, 1).astype(np.float32)
    z = (np.float32(1.0) - np.abs(x)) - np.abs(y)
    neg = z < 0
    nx = np.where(neg, np.copysign(np.float32(1.0) - np.abs(y), x), x)
    ny = np.where(neg, np.copysign(np.float32(1.0) - np.abs(x), y), y)
    v = np.stack([nx, ny, z], axis=1).astype(np.float32)
    return (v / np.sqrt(np.sum(v * v, axis=1, keepdims=True, dtype=np.float32))).astype(np.float32)


def device_encode_normal(n: np.ndarray) -> np.ndarray:
    """encode_normal, encoding.slang:28-30 (packSnorm2x16: round(clamp(v,-1,1)*32767), ties to even)."""
    p = np.clip(_oct_project(n), -1, 1)
    r = np.rint(p * np.float32(32767.0)).astype(np.int32)
    return ((r[:, 0] & 0xFFFF) | ((r[:, 1] & 0xFFFF) << 16)).astype(np.uint32)


def _normalize(v: np.ndarray) -> np.ndarray:
    v = v.astype(np.float32)
    return (v / np.sqrt(np.sum(v * v, axis=1, keepdims=True, dtype=np.float32))).astype(np.float32)


def pack_vertices(positions, normals, uvs=None, tangents=None, transform=None) -> np.ndarray:
    """Host packing (gltf_scene.cpp:107-118) followed by the device transform pass
    (transform_vertex.slang:12-40) with `transform` (4x4 row-major, None = identity): the bits the renderer reads."""
    positions = np.ascontiguousarray(positions, dtype=np.float32)
    n = positions.shape[0]
    normals = np.ascontiguousarray(normals, dtype=np.float32)
    uvs = np.zeros((n, 2), np.float32) if uvs is None else np.ascontiguousarray(uvs, dtype=np.float32)
    M = np.eye(4, dtype=np.float32) if transform is None else np.asarray(transform, dtype=np.float32)
    M_it = np.linalg.inv(M.astype(np.float64)).T.astype(np.float32)

    host_enc_n = host_encode_normal(normals)
    if tangents is not None:
        tangents = np.ascontiguousarray(tangents, dtype=np.float32)
        l = np.sqrt(np.sum(tangents[:, :3] ** 2, axis=1, keepdims=True))
        tn = np.where(l > 0, tangents[:, :3] / np.maximum(l, 1e-30), np.array([[0, 0, 1]], np.float32))
        host_enc_t = (host_encode_normal(tn) & np.uint32(0xFFFFFFFE)) | (tangents[:, 3] < 0).astype(np.uint32)
    else:
        host_enc_t = np.zeros(n, np.uint32)  # gltf_scene.cpp:116

    # device pass
    out = np.zeros(n, dtype=abi.PACKED_VERTEX_DTYPE)
    p4 = np.concatenate([positions, np.ones((n, 1), np.float32)], axis=1)
    out["position"] = (p4 @ M.T)[:, :3].astype(np.float32)
    nd = device_decode_normal(host_enc_n)
    out["encoded_normal"] = device_encode_normal(_normalize(nd @ M_it[:3, :3].T))
    out["uv"] = uvs.astype(np.float16).view(np.uint16)
    td = device_decode_normal(host_enc_t & np.uint32(0xFFFFFFFE))
    tw = _normalize(td @ M[:3, :3].T)
    out["encoded_tangent"] = (device_encode_normal(tw) & np.uint32(0xFFFFFFFE)) | (host_enc_t & np.uint32(1))
    return out


# ----------------------------------------------------------------------------------------------- materials
@dataclass
class DiffuseMaterial:
    """DiffuseMaterial, include/merian-shaders/shading/materials/diffuse-material.slang:8-12."""
    base_color_factor: Tuple[float, float, float, float] = (1, 1, 1, 1)
    base_color_texture: int = abi.TEXTURE_NONE  # = header.alpha_texture_id

    def type_key(self):
        return (abi.MATERIAL_DIFFUSE, 0)

    def alpha_texture(self):
        return self.base_color_texture

    def payload(self) -> bytes:
        return struct.pack("<4f", *self.base_color_factor)


@dataclass
class GltfMaterial:
    """merian::GltfMaterial, include/merian-shaders/shading/materials/gltf_material.hpp:57-146."""
    base_color_factor: Tuple[float, float, float, float] = (1, 1, 1, 1)
    emissive_factor: Tuple[float, float, float] = (0, 0, 0)
    metallic_factor: float = 1.0
    roughness_factor: float = 1.0
    normal_scale: float = 1.0
    occlusion_strength: float = 1.0
    ior: float = 1.5
    base_color_texture: int = abi.TEXTURE_NONE  # header.alpha_texture_id
    metallic_roughness_texture: int = abi.TEXTURE_NONE
    normal_texture: int = abi.TEXTURE_NONE
    occlusion_texture: int = abi.TEXTURE_NONE
    emissive_texture: int = abi.TEXTURE_NONE
    transmission: Optional[dict] = None  # {weight}
    volume: Optional[dict] = None        # {absorption: (r,g,b)}
    clearcoat: Optional[dict] = None     # {weight, roughness, texture, roughness_texture}
    sheen: Optional[dict] = None         # {color, roughness, color_texture, roughness_texture}
    iridescence: Optional[dict] = None   # {factor, ior, thickness_min, thickness_max, texture, thickness_texture}
    anisotropy: Optional[dict] = None    # {strength, rotation, texture}

    def features(self) -> int:
        f = 0
        for bit, ext in ((abi.GLTF_TRANSMISSION, self.transmission), (abi.GLTF_VOLUME, self.volume),
                         (abi.GLTF_CLEARCOAT, self.clearcoat), (abi.GLTF_SHEEN, self.sheen),
                         (abi.GLTF_IRIDESCENCE, self.iridescence), (abi.GLTF_ANISOTROPY, self.anisotropy)):
            if ext is not None:
                f |= bit
        return f

    def type_key(self):
        return (abi.MATERIAL_GLTF, self.features())

    def alpha_texture(self):
        return self.base_color_texture

    def payload(self) -> bytes:  # gltf_material.hpp:93-146 serialize()
        N = abi.TEXTURE_NONE
        b = struct.pack("<4f3f5f4H", *self.base_color_factor, *self.emissive_factor, self.metallic_factor,
                        self.roughness_factor, self.normal_scale, self.occlusion_strength, self.ior,
                        self.metallic_roughness_texture, self.normal_texture, self.occlusion_texture,
                        self.emissive_texture)
        if self.transmission is not None:
            b += struct.pack("<f", self.transmission.get("weight", 0.0))
        if self.volume is not None:
            b += struct.pack("<3f", *self.volume.get("absorption", (0, 0, 0)))
        if self.clearcoat is not None:
            c = self.clearcoat
            b += struct.pack("<2f2H", c.get("weight", 0.0), c.get("roughness", 0.0), c.get("texture", N),
                             c.get("roughness_texture", N))
        if self.sheen is not None:
            s = self.sheen
            b += struct.pack("<3ff2H", *s.get("color", (0, 0, 0)), s.get("roughness", 0.0), s.get("color_texture", N),
                             s.get("roughness_texture", N))
        if self.iridescence is not None:
            i = self.iridescence
            b += struct.pack("<4f2H", i.get("factor", 0.0), i.get("ior", 1.3), i.get("thickness_min", 100.0),
                             i.get("thickness_max", 400.0), i.get("texture", N), i.get("thickness_texture", N))
        if self.anisotropy is not None:
            a = self.anisotropy
            b += struct.pack("<2fH", a.get("strength", 0.0), a.get("rotation", 0.0), a.get("texture", N))
        return b


def build_material_buffer(materials: Sequence) -> Tuple[bytes, int, List[Tuple[int, int]]]:
    """MaterialSystem::add_material / repack_host_buffer (material_system.cpp:140-195).
    Returns (entries, entry_size, type table) — type ids in first-registration order (:124-139)."""
    types: Dict[Tuple[int, int], int] = {}
    payloads = []
    for m in materials:
        key = m.type_key()
        if key not in types:
            types[key] = len(types)
        payloads.append((types[key], m.alpha_texture(), m.payload()))
    max_payload = max((len(p[2]) for p in payloads), default=0)
    entry_size = 4 + ((max_payload + 3) // 4) * 4  # header + payload_size_in_uints * 4
    buf = bytearray(entry_size * len(payloads))
    for i, (tid, atex, pl) in enumerate(payloads):
        struct.pack_into("<2H", buf, i * entry_size, tid, atex)
        buf[i * entry_size + 4:i * entry_size + 4 + len(pl)] = pl
    type_table = [k for k, _ in sorted(types.items(), key=lambda kv: kv[1])]
    return bytes(buf), entry_size, type_table


# ----------------------------------------------------------------------------------------------- scene graph
@dataclass
class Mesh:
    """merian::Scene::Mesh (scene.hpp:291-309): host vertex/index sources + flags."""
    positions: np.ndarray
    normals: np.ndarray
    uvs: Optional[np.ndarray] = None
    tangents: Optional[np.ndarray] = None
    indices: Optional[np.ndarray] = None  # (P,3) uint8/uint16/uint32, or None (IndexType::None)
    material_id: int = 0
    flags: int = MESH_IS_OPAQUE
    instance_mask: int = 0x01  # Scene::Mesh default (scene.hpp:193)


@dataclass
class Node:
    """One scene-graph node with a global transform and the meshes it instantiates."""
    meshes: List[int]
    transform: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))


@dataclass
class Texture:
    pixels: object  # (H,W,4) uint8 or float32 array, or a callable returning one (generated at upload time)
    srgb: bool = False
    wrap_u: int = abi.WRAP_REPEAT
    wrap_v: int = abi.WRAP_REPEAT
    filter: int = abi.FILTER_LINEAR
    mipmap: int = abi.MIPMAP_NONE  # a glTF sampler with a *_MIPMAP_* minFilter gets a mip chain (gltf_scene.cpp:212-248, 583-593)


@dataclass
class EnvMap:
    pixels: np.ndarray  # (H,W,4) float32 lat-long
    intensity: float = 1.0
    to_local: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))


@dataclass
class CameraDesc:
    """merian::Camera host state (camera.cpp)."""
    position: Tuple[float, float, float]
    target: Tuple[float, float, float]
    up: Tuple[float, float, float] = (0, 1, 0)
    yfov_deg: float = 45.0
    aspect: float = 1.0
    near: float = 0.01
    far: float = 1e4
    jitter_sequence: str = "None"  # "None" | "R2" (glTF cameras use R2 over 16 frames, gltf_scene.cpp:773-774)

    def jitter(self, frame: int) -> Tuple[float, float]:
        """Camera::advance_jitter, camera.cpp:318-347 (None / R2)."""
        if self.jitter_sequence == "R2":
            n = np.float32(frame % 16)
            ax, ay = np.float32(0.7548776662466927), np.float32(0.5698402909980532)
            jx = np.fmod(np.float32(0.5) + ax * n, np.float32(1.0)) - np.float32(0.5)
            jy = np.fmod(np.float32(0.5) + ay * n, np.float32(1.0)) - np.float32(0.5)
            return float(jx), float(jy)
        return 0.0, 0.0

    def to_abi(self, frame: int = 0) -> abi.Camera:
        """Camera::write_to, camera.cpp:349-368."""
        f32 = np.float32
        pos = np.array(self.position, f32)
        tgt = np.array(self.target, f32)
        up = np.array(self.up, f32)
        fwd = tgt - pos
        focus = f32(np.sqrt(np.sum(fwd * fwd, dtype=f32)))  # focus_distance_on_target (camera.cpp:255-260)
        fwd = (fwd / focus).astype(f32)
        rgt = np.cross(fwd, up).astype(f32)
        rgt = (rgt / np.sqrt(np.sum(rgt * rgt, dtype=f32))).astype(f32)
        half_tan = f32(np.tan(f32(np.deg2rad(self.yfov_deg)) * f32(0.5)))
        cam = abi.Camera()
        cam.position[:] = pos.tolist()
        cam.target[:] = tgt.tolist()
        cam.up[:] = up.tolist()
        cam.U[:] = (rgt * f32(half_tan * f32(self.aspect) * focus)).tolist()
        cam.V[:] = (np.cross(rgt, fwd).astype(f32) * f32(half_tan * focus)).tolist()
        cam.W[:] = (fwd * focus).tolist()
        cam.near_plane = self.near
        cam.far_plane = self.far
        cam.aspect_ratio = self.aspect
        cam.aperture_radius = 0.0
        jx, jy = self.jitter(frame)
        cam.jitter[:] = [jx, jy]
        return cam


@dataclass
class SceneDescription:
    """What a loader hands to merian::Scene: meshes, nodes, materials, textures, env, camera."""
    meshes: List[Mesh]
    nodes: List[Node]
    materials: List
    textures: List[Texture] = field(default_factory=list)
    env: Optional[EnvMap] = None
    camera: Optional[CameraDesc] = None
    name: str = "scene"
    seed: int = 0


class SceneTables:
    """The flat tables the renderer consumes, with ctypes views that keep the numpy storage alive."""

    def __init__(self, desc: SceneDescription):
        self.desc = desc
        self._build(desc)

    # scene.cpp:1023-1105 (grouping), 1236-1291 (geometry order), 1808-1857 (TLAS)
    def _build(self, d: SceneDescription):
        n_mesh = len(d.meshes)
        instances_of: List[List[int]] = [[] for _ in range(n_mesh)]
        for node_id, node in enumerate(d.nodes):
            for m in node.meshes:
                instances_of[m].append(node_id)  # SmallSet<NodeID>: ascending, unique (small_set.hpp:13)
        instances_of = [sorted(set(v)) for v in instances_of]

        groups: List[dict] = []
        groups_static: Dict[int, int] = {}
        groups_instanced: Dict[Tuple[int, ...], Dict[int, int]] = {}

        def get_or_create(table, key):
            if key not in table:
                table[key] = len(groups)
                groups.append({"meshes": [], "flags": 0, "mask": 0, "all_opaque": True, "instanced": False})
            return table[key]

        for mesh_id in range(n_mesh):  # ascending MeshID
            inst = instances_of[mesh_id]
            if not inst:
                continue
            mesh = d.meshes[mesh_id]
            key = (mesh.flags & GROUP_SPLIT_MASK) | (mesh.instance_mask << 16)
            if len(inst) > 1:
                gid = get_or_create(groups_instanced.setdefault(tuple(inst), {}), key)
                groups[gid]["instanced"] = True
            else:
                gid = get_or_create(groups_static, key)  # static, non-animated, non-morphed
            g = groups[gid]
            g["flags"] = key & GROUP_SPLIT_MASK
            g["mask"] = mesh.instance_mask
            g["all_opaque"] = g["all_opaque"] and bool(mesh.flags & MESH_IS_OPAQUE)
            g["meshes"].append(mesh_id)  # SmallSet<MeshID>: ascending because mesh_id ascends
            if len(g["meshes"]) == 1:
                g["instances"] = inst  # MeshGroup::get_instances = instances of the FIRST mesh (scene.hpp:380-383)
        self.groups = groups

        # ---- device vertex/index data per mesh (A13). Pretransformed groups bake the node transform.
        self.mesh_vertices: List[np.ndarray] = [None] * n_mesh
        self.mesh_indices: List[Optional[np.ndarray]] = [None] * n_mesh
        self.mesh_index_flag = [0] * n_mesh
        for g in groups:
            for mesh_id in g["meshes"]:
                mesh = d.meshes[mesh_id]
                # scene.cpp:1498-1501: a pretransformed mesh bakes the transform of ITS OWN (single) node
                M = None if g["instanced"] else d.nodes[instances_of[mesh_id][0]].transform
                self.mesh_vertices[mesh_id] = pack_vertices(mesh.positions, mesh.normals, mesh.uvs, mesh.tangents, M)
                if mesh.indices is not None:
                    idx = np.ascontiguousarray(mesh.indices)
                    flag = {1: abi.GEOM_INDEX_U8, 2: abi.GEOM_INDEX_U16, 4: abi.GEOM_INDEX_U32}[idx.dtype.itemsize]
                    self.mesh_indices[mesh_id] = idx
                    self.mesh_index_flag[mesh_id] = flag

        # ---- GeometryData[] + TLAS instances, same iteration order for both
        geoms, insts, xf, xf_it = [], [], [], []
        instance_id = 0
        ident = np.eye(4, dtype=np.float32)
        for gid, g in enumerate(groups):
            pre = not g["instanced"]
            for node_id in g["instances"]:
                T = d.nodes[node_id].transform.astype(np.float32)
                is_ident = bool(np.array_equal(T, ident))
                for mesh_id in g["meshes"]:
                    mesh = d.meshes[mesh_id]
                    f = self.mesh_index_flag[mesh_id]
                    if pre or is_ident:
                        f |= abi.GEOM_PRETRANSFORMED
                    if mesh.flags & MESH_FLIP_FACING:
                        f |= abi.GEOM_FLIP_FACING
                    if mesh.flags & MESH_HAS_TANGENTS:
                        f |= abi.GEOM_HAS_TANGENTS
                    if mesh.flags & MESH_FLAT_SHADING:
                        f |= abi.GEOM_FLAT_SHADING
                    if mesh.flags & MESH_USE_ENV_MAP:
                        f |= abi.GEOM_USE_ENV_MAP
                    if mesh.flags & MESH_TWO_SIDED:
                        f |= abi.GEOM_TWO_SIDED
                    prim = (len(mesh.indices) if mesh.indices is not None else len(mesh.positions) // 3)
                    geoms.append((mesh_id, mesh.material_id, f, prim))
                iflags = 0
                if g["flags"] & MESH_FLIP_FACING:
                    iflags |= abi.INSTANCE_FLIP_FACING
                if g["flags"] & MESH_TWO_SIDED:
                    iflags |= abi.INSTANCE_FACING_CULL_DISABLE
                if g["all_opaque"]:
                    iflags |= abi.INSTANCE_FORCE_OPAQUE
                Ti = ident if pre else T
                insts.append((Ti[:3, :].reshape(-1).copy(), instance_id, g["mask"], iflags, gid))
                xf.append(Ti.copy())
                xf_it.append(np.linalg.inv(Ti.astype(np.float64)).T.astype(np.float32))
                instance_id += len(g["meshes"])
        self.geometries = geoms
        self.instances = insts
        self.instance_transforms = np.ascontiguousarray(np.stack(xf), np.float32) if xf else np.zeros((0, 4, 4), np.float32)
        self.inv_t_instance_transforms = np.ascontiguousarray(np.stack(xf_it), np.float32) if xf_it else np.zeros((0, 4, 4), np.float32)

        self.material_entries, self.material_entry_size, self.material_types = build_material_buffer(d.materials)
        self.triangle_count = sum(g[3] for g in geoms)

    # ---- ctypes views -------------------------------------------------------------------------
    def scene_desc(self) -> abi.SceneDesc:
        d = self.desc
        n_mesh = len(d.meshes)
        self._keep = []
        meshes = (abi.MeshDesc * n_mesh)()
        for i in range(n_mesh):
            v = self.mesh_vertices[i]
            if v is None:  # unreferenced mesh
                continue
            meshes[i].vertices = v.ctypes.data
            meshes[i].vertex_count = len(v)
            idx = self.mesh_indices[i]
            meshes[i].indices = idx.ctypes.data if idx is not None else None
            meshes[i].primitive_count = len(idx) if idx is not None else len(v) // 3
            meshes[i].index_flag = self.mesh_index_flag[i]
            meshes[i].opaque = 1 if (d.meshes[i].flags & MESH_IS_OPAQUE) else 0
        blas = (abi.BlasDesc * len(self.groups))()
        for gi, g in enumerate(self.groups):
            arr = np.asarray(g["meshes"], dtype=np.uint32)
            self._keep.append(arr)
            blas[gi].meshes = arr.ctypes.data
            blas[gi].mesh_count = len(arr)
        geoms = (abi.GeometryDesc * len(self.geometries))()
        for i, (mesh_id, mat, f, prim) in enumerate(self.geometries):
            geoms[i].mesh, geoms[i].material_id, geoms[i].flags, geoms[i].primitive_count = mesh_id, mat, f, prim
        insts = (abi.InstanceDesc * len(self.instances))()
        for i, (t, iid, mask, fl, gid) in enumerate(self.instances):
            insts[i].transform[:] = t.tolist()
            insts[i].instance_custom_index, insts[i].mask, insts[i].flags, insts[i].blas = iid, mask, fl, gid
        sd = abi.SceneDesc()
        sd.meshes = C.cast(meshes, C.c_void_p)
        sd.mesh_count = n_mesh
        sd.blas = C.cast(blas, C.c_void_p)
        sd.blas_count = len(self.groups)
        sd.geometries = C.cast(geoms, C.c_void_p)
        sd.geometry_count = len(self.geometries)
        sd.instances = C.cast(insts, C.c_void_p)
        sd.instance_count = len(self.instances)
        sd.instance_transforms = self.instance_transforms.ctypes.data
        sd.inverse_transposed_instance_transforms = self.inv_t_instance_transforms.ctypes.data
        self._keep += [meshes, blas, geoms, insts]
        return sd

    def material_types_array(self):
        arr = (abi.MaterialType * len(self.material_types))()
        for i, (model, feat) in enumerate(self.material_types):
            arr[i].model, arr[i].features = model, feat
        return arr


def upload_scene(backend, tables: SceneTables):
    """Push tables, materials, textures and env through a backend exposing the mpt_* entry points
    (merian_b200.Context for the product; oracle_py.OracleContext in tests)."""
    import time
    d = tables.desc
    timing = {"generate_s": 0.0, "texture_upload_s": 0.0, "scene_upload_s": 0.0}  # harness-side texel synthesis vs the library's share
    for tid, t in enumerate(d.textures):
        t0 = time.perf_counter()
        px = np.ascontiguousarray(t.pixels() if callable(t.pixels) else t.pixels)  # callables defer large synthetic textures
        if px.dtype == np.uint8:
            fmt = abi.TEX_RGBA8_SRGB if t.srgb else abi.TEX_RGBA8_UNORM
        else:
            px = px.astype(np.float32)
            fmt = abi.TEX_RGBA32F
        t1 = time.perf_counter()
        backend.texture_upload(tid, px.shape[1], px.shape[0], fmt, t.wrap_u, t.wrap_v, t.filter, px, getattr(t, "mipmap", abi.MIPMAP_NONE))
        timing["generate_s"] += t1 - t0
        timing["texture_upload_s"] += time.perf_counter() - t1
    backend.materials_upload(tables.material_entries, tables.material_entry_size, len(d.materials),
                             tables.material_types_array(), len(tables.material_types))
    if d.env is not None:
        px = np.ascontiguousarray(d.env.pixels, dtype=np.float32)
        backend.env_set(abi.ENV_LATLONG, np.ascontiguousarray(d.env.to_local, np.float32), d.env.intensity, px.shape[1],
                        px.shape[0], px)
    else:
        backend.env_set(abi.ENV_EMPTY, None, 1.0, 0, 0, None)
    t0 = time.perf_counter()
    backend.scene_upload(tables.scene_desc())
    timing["scene_upload_s"] = time.perf_counter() - t0
    tables.upload_timing = timing
