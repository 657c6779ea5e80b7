"""CGltf: a glTF file loaded by libmpt's own C++ loader (include/mpt_tables.h, mptt_gltf_*), exposed as the same SceneDescription the
Python loader (gltf.load_gltf) builds — images decoded here with Pillow from the encoded bytes the C++ side hands over. Exists so that
the two restatements of GLTFScene::load can be compared (tests/test_gltf_cabi.py), and as the binding a Python host would use."""
import ctypes as C
import io
import os

import numpy as np

from . import abi
from ._lib import MptError, lib
from .scene import CameraDesc, DiffuseMaterial, GltfMaterial, Mesh, Node, SceneDescription, Texture


def _floats(ptr, n, width):
    if not ptr:
        return None
    return np.frombuffer((C.c_char * (n * width * 4)).from_address(ptr), np.float32, n * width).reshape(n, width).copy()


def _material(s):
    if s.model == abi.MATERIAL_DIFFUSE:
        return DiffuseMaterial(tuple(s.base_color_factor[:]), s.base_color_texture)
    m = GltfMaterial(base_color_factor=tuple(s.base_color_factor[:]), emissive_factor=tuple(s.emissive_factor[:]), metallic_factor=s.metallic_factor,
                     roughness_factor=s.roughness_factor, normal_scale=s.normal_scale, occlusion_strength=s.occlusion_strength, ior=s.ior,
                     base_color_texture=s.base_color_texture, metallic_roughness_texture=s.metallic_roughness_texture, normal_texture=s.normal_texture,
                     occlusion_texture=s.occlusion_texture, emissive_texture=s.emissive_texture)
    f = s.features
    if f & abi.GLTF_TRANSMISSION:
        m.transmission = {"weight": s.transmission_weight}
    if f & abi.GLTF_VOLUME:
        m.volume = {"absorption": tuple(s.volume_absorption[:])}
    if f & abi.GLTF_CLEARCOAT:
        m.clearcoat = {"weight": s.clearcoat_weight, "roughness": s.clearcoat_roughness, "texture": s.clearcoat_texture, "roughness_texture": s.clearcoat_roughness_texture}
    if f & abi.GLTF_SHEEN:
        m.sheen = {"color": tuple(s.sheen_color[:]), "roughness": s.sheen_roughness, "color_texture": s.sheen_color_texture, "roughness_texture": s.sheen_roughness_texture}
    if f & abi.GLTF_IRIDESCENCE:
        m.iridescence = {"factor": s.iridescence_factor, "ior": s.iridescence_ior, "thickness_min": s.iridescence_thickness_min,
                         "thickness_max": s.iridescence_thickness_max, "texture": s.iridescence_texture, "thickness_texture": s.iridescence_thickness_texture}
    if f & abi.GLTF_ANISOTROPY:
        m.anisotropy = {"strength": s.anisotropy_strength, "rotation": s.anisotropy_rotation, "texture": s.anisotropy_texture}
    return m


def load_gltf(path) -> SceneDescription:
    from PIL import Image
    l = lib()
    h = C.c_void_p()
    st = l.mptt_gltf_load(os.fsencode(str(path)), C.byref(h))
    if st != abi.MPT_OK:
        raise MptError(st, l.mpt_last_error(None).decode())
    try:
        v = abi.GltfView()
        st = l.mptt_gltf_get(h, C.byref(v))
        if st != abi.MPT_OK:
            raise MptError(st, l.mpt_last_error(None).decode())
        meshes = []
        for s in ((abi.SrcMesh * v.mesh_count).from_address(v.meshes) if v.mesh_count else []):
            n = s.vertex_count
            idx = None
            if s.indices and s.index_size:
                dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[s.index_size]
                idx = np.frombuffer((C.c_char * (s.primitive_count * 3 * s.index_size)).from_address(s.indices), dt, s.primitive_count * 3).reshape(-1, 3).copy()
            meshes.append(Mesh(positions=_floats(s.positions, n, 3), normals=_floats(s.normals, n, 3), uvs=_floats(s.uvs, n, 2), tangents=_floats(s.tangents, n, 4),
                               indices=idx, material_id=s.material_id, flags=s.flags, instance_mask=s.instance_mask))
        nodes = []
        for s in ((abi.SrcNode * v.node_count).from_address(v.nodes) if v.node_count else []):
            ids = np.frombuffer((C.c_char * (s.mesh_count * 4)).from_address(s.meshes), np.uint32, s.mesh_count).tolist() if s.mesh_count else []
            nodes.append(Node(meshes=ids, transform=np.array(s.transform[:], np.float32).reshape(4, 4)))
        materials = [_material(s) for s in ((abi.SrcMaterial * v.material_count).from_address(v.materials) if v.material_count else [])]
        textures = []
        for t in ((abi.GltfTexture * v.texture_count).from_address(v.textures) if v.texture_count else []):
            raw = bytes((C.c_char * t.encoded_size).from_address(t.encoded))
            px = np.ascontiguousarray(np.array(Image.open(io.BytesIO(raw)).convert("RGBA"), dtype=np.uint8))
            textures.append(Texture(px, srgb=bool(t.srgb), wrap_u=t.wrap_u, wrap_v=t.wrap_v, filter=t.filter, mipmap=t.mipmap))
        camera = None
        if v.has_camera:
            c = v.camera
            camera = CameraDesc(position=tuple(c.position[:]), target=tuple(c.target[:]), up=tuple(c.up[:]), yfov_deg=v.yfov_degrees, aspect=v.aspect,
                                near=v.znear, far=v.zfar, jitter_sequence="R2")
        desc = SceneDescription(meshes=meshes, nodes=nodes, materials=materials, textures=textures, env=None, camera=camera,
                                name=os.path.splitext(os.path.basename(str(path)))[0])
        desc.c_camera = abi.Camera.from_buffer_copy(v.camera) if v.has_camera else None  # the basis as mptt_camera_look_at computed it
        return desc
    finally:
        l.mptt_gltf_free(h)
