"""glTF 2.0 ingestion straight into the host scene description, and the matching writer.

`load_gltf` restates what GLTFScene::load hands to merian::Scene (src/merian-shaders/scene/gltf_scene.cpp):
  * materials incl. the KHR extension gates          gltf_scene.cpp:265-536   (load_materials)
  * lazy texture upload order -> TextureIDs, sRGB/linear per use   :538-604   (get_or_load_texture)
  * one Scene::Mesh per triangle primitive, flags, default material :607-694  (load_meshes)
  * node transforms (matrix or T*R*S), DFS node order, instances    :146-196, 696-716 (gltf_node_transform, load_node)
  * perspective camera from the node's global transform, R2 jitter  :719-775  (load_cameras)
The reference parses with tinygltf (un-vendored wrap) and decodes images with stb_image to RGBA8; here the container is parsed
with `json` and images are decoded with Pillow to RGBA8. `SceneTables(load_gltf(path))` then yields the flat tables
`mpt_scene_upload` consumes: no Vulkan Scene in between.

`save_gltf` writes a SceneDescription as .gltf + .bin (+ PNG textures) so that the synthetic scenes of BASELINE.json can be fed to
the reference's own `merian-graph-run gltf.json <scene.gltf>`.
"""
from __future__ import annotations

import base64
import io
import json
import os
import struct
from typing import List, Optional

import numpy as np

from . import abi
from .scene import (MESH_HAS_TANGENTS, MESH_IS_OPAQUE, MESH_TWO_SIDED, CameraDesc, GltfMaterial, Mesh, Node, SceneDescription, Texture)

_COMPONENT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_TYPE_WIDTH = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}
_WRAP = {33071: abi.WRAP_CLAMP_TO_EDGE, 33648: abi.WRAP_MIRRORED_REPEAT, 10497: abi.WRAP_REPEAT}


class GltfError(ValueError):
    pass


# ----------------------------------------------------------------------------------------------- container
def _read_container(path):
    """Returns (json dict, GLB binary chunk or None)."""
    with open(path, "rb") as f:
        data = f.read()
    if str(path).lower().endswith(".glb") or data[:4] == b"glTF":
        magic, version, length = struct.unpack_from("<4sII", data, 0)
        if magic != b"glTF" or version != 2:
            raise GltfError(f"{path}: not a glTF 2.0 binary container")
        off, doc, blob = 12, None, None
        while off + 8 <= min(length, len(data)):
            clen, ctype = struct.unpack_from("<II", data, off)
            chunk = data[off + 8:off + 8 + clen]
            if ctype == 0x4E4F534A:
                doc = json.loads(chunk.decode("utf-8"))
            elif ctype == 0x004E4942 and blob is None:
                blob = chunk
            off += 8 + clen
        if doc is None:
            raise GltfError(f"{path}: GLB without a JSON chunk")
        return doc, blob
    return json.loads(data.decode("utf-8")), None


def _load_uri(uri, base_dir):
    if uri.startswith("data:"):
        return base64.b64decode(uri.split(",", 1)[1])
    with open(os.path.join(base_dir, uri), "rb") as f:
        return f.read()


class _Doc:
    def __init__(self, path):
        self.path = str(path)
        self.base = os.path.dirname(os.path.abspath(self.path))
        self.j, glb = _read_container(path)
        self.buffers = []
        for i, b in enumerate(self.j.get("buffers", [])):
            if "uri" in b:
                self.buffers.append(_load_uri(b["uri"], self.base))
            elif i == 0 and glb is not None:
                self.buffers.append(glb)
            else:
                raise GltfError(f"buffer {i} has no data")

    def view_bytes(self, view_index):
        v = self.j["bufferViews"][view_index]
        off = v.get("byteOffset", 0)
        return self.buffers[v["buffer"]][off:off + v["byteLength"]], v.get("byteStride", 0)

    def accessor(self, index, expect_float=None):
        """get_accessor_base / get_accessor_byte_stride + read_strided (gltf_scene.cpp:30-60): a (count, width) array."""
        a = self.j["accessors"][index]
        dt = np.dtype(_COMPONENT[a["componentType"]])
        width = _TYPE_WIDTH[a["type"]]
        if expect_float is not None and dt != np.float32:
            raise GltfError(f"accessor {index} ({expect_float}) is not FLOAT: the reference reads these attributes as float (gltf_scene.cpp:127-146)")
        count = a["count"]
        if "bufferView" not in a:
            return np.zeros((count, width), dt)
        raw, stride = self.view_bytes(a["bufferView"])
        off = a.get("byteOffset", 0)
        elem = dt.itemsize * width
        stride = stride or elem
        if stride == elem:
            out = np.frombuffer(raw, dt, count * width, off).reshape(count, width)
        else:
            need = off + (count - 1) * stride + elem
            if need > len(raw):
                raise GltfError(f"accessor {index} overruns its buffer view")
            out = np.lib.stride_tricks.as_strided(np.frombuffer(raw, np.uint8, offset=off), (count, elem), (stride, 1)).copy().view(dt).reshape(count, width)
        return np.array(out)


# ----------------------------------------------------------------------------------------------- transforms
def _quat_to_matrix(q):
    """glm::mat4_cast (glm pinned by subprojects/glm.wrap; gltf_scene.cpp:173-180), fp32 arithmetic, returned row-major."""
    f = np.float32
    x, y, z, w = (f(v) for v in q)
    one, two = f(1), f(2)
    qxx, qyy, qzz, qxz, qxy, qyz, qwx, qwy, qwz = x * x, y * y, z * z, x * z, x * y, y * z, w * x, w * y, w * z
    R = np.eye(4, dtype=np.float32)
    R[0, 0] = one - two * (qyy + qzz); R[1, 0] = two * (qxy + qwz); R[2, 0] = two * (qxz - qwy)
    R[0, 1] = two * (qxy - qwz); R[1, 1] = one - two * (qxx + qzz); R[2, 1] = two * (qyz + qwx)
    R[0, 2] = two * (qxz + qwy); R[1, 2] = two * (qyz - qwx); R[2, 2] = one - two * (qxx + qyy)
    return R


def _mul4(a, b):
    """hlslpp mul(float4x4, float4x4) in fp32: rows of a dotted with columns of b, accumulated left to right."""
    out = np.zeros((4, 4), np.float32)
    for i in range(4):
        for j in range(4):
            acc = np.float32(a[i, 0]) * np.float32(b[0, j])
            for k in range(1, 4):
                acc = np.float32(acc + np.float32(a[i, k]) * np.float32(b[k, j]))
            out[i, j] = acc
    return out


def node_transform(node) -> np.ndarray:
    """gltf_node_transform, gltf_scene.cpp:146-196."""
    if len(node.get("matrix", [])) == 16:
        return np.array(node["matrix"], np.float64).reshape(4, 4).T.astype(np.float32)  # column-major in the file
    T, R, S = (np.eye(4, dtype=np.float32) for _ in range(3))
    if len(node.get("translation", [])) == 3:
        T[:3, 3] = np.array(node["translation"], np.float32)
    if len(node.get("rotation", [])) == 4:
        R = _quat_to_matrix(node["rotation"])
    if len(node.get("scale", [])) == 3:
        S[0, 0], S[1, 1], S[2, 2] = (np.float32(v) for v in node["scale"])
    return _mul4(_mul4(T, R), S)


# ----------------------------------------------------------------------------------------------- loader
def _decode_image(doc: _Doc, image_index):
    from PIL import Image
    img = doc.j["images"][image_index]
    if "uri" in img:
        raw = _load_uri(img["uri"], doc.base)
    elif "bufferView" in img:
        raw, _ = doc.view_bytes(img["bufferView"])
    else:
        return None
    try:
        return np.ascontiguousarray(np.array(Image.open(io.BytesIO(raw)).convert("RGBA"), dtype=np.uint8))
    except Exception:  # invalid image: the reference warns and leaves the slot empty (gltf_scene.cpp:556-559)
        return None


def load_gltf(path, force_mipmaps_color: bool = False) -> SceneDescription:
    doc = _Doc(path)
    j = doc.j
    textures: List[Texture] = []
    slots = {}  # (gltf texture index, linear) -> TextureID, in order of first use (get_or_load_texture)

    def get_or_load_texture(index, linear):
        if index is None or index < 0 or index >= len(j.get("textures", [])):
            return abi.TEXTURE_NONE
        key = (index, bool(linear))
        if key in slots:
            return slots[key]
        tex = j["textures"][index]
        src = tex.get("source", -1)
        if src < 0 or src >= len(j.get("images", [])):
            return abi.TEXTURE_NONE
        px = _decode_image(doc, src)
        if px is None or px.shape[0] <= 0 or px.shape[1] <= 0:
            return abi.TEXTURE_NONE
        filt, wu, wv, mip = abi.FILTER_LINEAR, abi.WRAP_REPEAT, abi.WRAP_REPEAT, abi.MIPMAP_NONE
        s = tex.get("sampler", -1)
        if 0 <= s < len(j.get("samplers", [])):
            sm = j["samplers"][s]
            if sm.get("magFilter", -1) == 9728:  # NEAREST (gltf_mag_filter_to_vk); libmpt has one filter per texture: mag wins
                filt = abi.FILTER_NEAREST
            # gltf_min_filter_to_vk (gltf_scene.cpp:212-248): the *_MIPMAP_* values ask for a chain, its second half is the mipmap mode.
            # GLTFScene builds the chain for colour AND data textures when the sampler asks for it (gltf_scene.cpp:583-593).
            mip = {9984: abi.MIPMAP_NEAREST, 9985: abi.MIPMAP_NEAREST, 9986: abi.MIPMAP_LINEAR, 9987: abi.MIPMAP_LINEAR}.get(sm.get("minFilter", -1), abi.MIPMAP_NONE)
            wu, wv = _WRAP.get(sm.get("wrapS", 10497), abi.WRAP_REPEAT), _WRAP.get(sm.get("wrapT", 10497), abi.WRAP_REPEAT)
        slots[key] = len(textures)
        textures.append(Texture(px, srgb=not linear, wrap_u=wu, wrap_v=wv, filter=filt, mipmap=mip))
        return slots[key]

    def tex_index(obj, key):
        t = obj.get(key)
        return t.get("index", -1) if isinstance(t, dict) else -1

    # ---- materials (load_materials)
    materials = []
    for gm in j.get("materials", []):
        pbr = gm.get("pbrMetallicRoughness", {})
        m = GltfMaterial(base_color_factor=tuple(float(v) for v in pbr.get("baseColorFactor", (1, 1, 1, 1))),
                         metallic_factor=float(pbr.get("metallicFactor", 1.0)), roughness_factor=float(pbr.get("roughnessFactor", 1.0)),
                         emissive_factor=tuple(float(v) for v in gm.get("emissiveFactor", (0, 0, 0))))
        if tex_index(pbr, "baseColorTexture") >= 0:
            m.base_color_texture = get_or_load_texture(tex_index(pbr, "baseColorTexture"), False)
        if tex_index(pbr, "metallicRoughnessTexture") >= 0:
            m.metallic_roughness_texture = get_or_load_texture(tex_index(pbr, "metallicRoughnessTexture"), True)
        if tex_index(gm, "normalTexture") >= 0:
            m.normal_texture = get_or_load_texture(tex_index(gm, "normalTexture"), True)
            m.normal_scale = float(gm["normalTexture"].get("scale", 1.0))
        if tex_index(gm, "emissiveTexture") >= 0:
            m.emissive_texture = get_or_load_texture(tex_index(gm, "emissiveTexture"), False)
        if tex_index(gm, "occlusionTexture") >= 0:
            m.occlusion_texture = get_or_load_texture(tex_index(gm, "occlusionTexture"), True)
            m.occlusion_strength = float(gm["occlusionTexture"].get("strength", 1.0))
        ext = gm.get("extensions", {})
        e = ext.get("KHR_materials_transmission")
        if e is not None and "transmissionFactor" in e and np.float32(e["transmissionFactor"]) > 0:
            m.transmission = {"weight": float(np.float32(e["transmissionFactor"]))}
        e = ext.get("KHR_materials_ior")
        if e is not None and "ior" in e:
            m.ior = float(e["ior"])
        e = ext.get("KHR_materials_volume")
        if e is not None:
            thickness = float(e.get("thicknessFactor", 0.0))
            col = [np.float32(c) for c in e.get("attenuationColor", (1, 1, 1))][:3]
            dist = float(e.get("attenuationDistance", np.inf))
            if thickness > 0.0 and np.isfinite(dist) and dist > 0.0:
                m.volume = {"absorption": tuple(float(np.float32(-np.log(np.float32(min(max(c, np.float32(1e-4)), np.float32(1.0)))) / dist)) for c in col)}
        e = ext.get("KHR_materials_clearcoat")
        if e is not None:
            c = {"weight": float(e.get("clearcoatFactor", 0.0)), "roughness": float(e.get("clearcoatRoughnessFactor", 0.0)),
                 "texture": abi.TEXTURE_NONE, "roughness_texture": abi.TEXTURE_NONE}
            if tex_index(e, "clearcoatTexture") >= 0:
                c["texture"] = get_or_load_texture(tex_index(e, "clearcoatTexture"), True)
            if tex_index(e, "clearcoatRoughnessTexture") >= 0:
                c["roughness_texture"] = get_or_load_texture(tex_index(e, "clearcoatRoughnessTexture"), True)
            if c["weight"] > 0.0:
                m.clearcoat = c
        e = ext.get("KHR_materials_sheen")
        if e is not None:
            col = tuple(float(v) for v in e.get("sheenColorFactor", (0, 0, 0)))[:3]
            s = {"color": col, "roughness": float(e.get("sheenRoughnessFactor", 0.0)), "color_texture": abi.TEXTURE_NONE,
                 "roughness_texture": abi.TEXTURE_NONE}
            if tex_index(e, "sheenColorTexture") >= 0:
                s["color_texture"] = get_or_load_texture(tex_index(e, "sheenColorTexture"), False)
            if tex_index(e, "sheenRoughnessTexture") >= 0:
                s["roughness_texture"] = get_or_load_texture(tex_index(e, "sheenRoughnessTexture"), True)
            if any(v > 0.0 for v in col):
                m.sheen = s
        e = ext.get("KHR_materials_iridescence")
        if e is not None:
            i = {"factor": float(e.get("iridescenceFactor", 0.0)), "ior": float(e.get("iridescenceIor", 1.3)),
                 "thickness_min": float(e.get("iridescenceThicknessMinimum", 100.0)), "thickness_max": float(e.get("iridescenceThicknessMaximum", 400.0)),
                 "texture": abi.TEXTURE_NONE, "thickness_texture": abi.TEXTURE_NONE}
            if tex_index(e, "iridescenceTexture") >= 0:
                i["texture"] = get_or_load_texture(tex_index(e, "iridescenceTexture"), True)
            if tex_index(e, "iridescenceThicknessTexture") >= 0:
                i["thickness_texture"] = get_or_load_texture(tex_index(e, "iridescenceThicknessTexture"), True)
            if i["factor"] > 0.0:
                m.iridescence = i
        e = ext.get("KHR_materials_anisotropy")
        if e is not None:
            a = {"strength": float(e.get("anisotropyStrength", 0.0)), "rotation": float(e.get("anisotropyRotation", 0.0)), "texture": abi.TEXTURE_NONE}
            if tex_index(e, "anisotropyTexture") >= 0:
                a["texture"] = get_or_load_texture(tex_index(e, "anisotropyTexture"), True)
            if a["strength"] > 0.0:
                m.anisotropy = a
        materials.append(m)
    if not materials:
        materials.append(GltfMaterial())  # default material for primitives without one (gltf_scene.cpp:530-535)

    # ---- meshes (load_meshes): one Scene::Mesh per TRIANGLES primitive with a POSITION attribute
    meshes: List[Mesh] = []
    mesh_map: List[List[int]] = []
    gmats = j.get("materials", [])
    for gmesh in j.get("meshes", []):
        ids = []
        for prim in gmesh.get("primitives", []):
            if prim.get("mode", 4) != 4 or "POSITION" not in prim.get("attributes", {}):
                continue
            at = prim["attributes"]
            pos = doc.accessor(at["POSITION"], "POSITION").astype(np.float32)[:, :3]
            n = len(pos)
            nrm = doc.accessor(at["NORMAL"], "NORMAL")[:, :3] if "NORMAL" in at else np.tile(np.array([[0, 1, 0]], np.float32), (n, 1))
            uv = doc.accessor(at["TEXCOORD_0"], "TEXCOORD_0")[:, :2] if "TEXCOORD_0" in at else None
            tan = doc.accessor(at["TANGENT"], "TANGENT")[:, :4] if "TANGENT" in at else None
            idx = None
            if prim.get("indices", -1) >= 0:
                raw = doc.accessor(prim["indices"]).reshape(-1)
                if raw.dtype not in (np.uint8, np.uint16, np.uint32):
                    raise GltfError("index accessor must be UNSIGNED_BYTE / UNSIGNED_SHORT / UNSIGNED_INT")
                idx = np.ascontiguousarray(raw[:(len(raw) // 3) * 3].reshape(-1, 3))
            mi = prim.get("material", -1)
            flags = 0
            if 0 <= mi < len(gmats):
                gm = gmats[mi]
                if gm.get("alphaMode", "OPAQUE") == "OPAQUE":
                    flags |= MESH_IS_OPAQUE
                if gm.get("doubleSided", False) or "KHR_materials_transmission" in gm.get("extensions", {}):
                    flags |= MESH_TWO_SIDED  # two-sided, or transmissive: refraction must hit the exit interface
            else:
                flags |= MESH_IS_OPAQUE
            if tan is not None:
                flags |= MESH_HAS_TANGENTS
            ids.append(len(meshes))
            meshes.append(Mesh(positions=np.ascontiguousarray(pos), normals=np.ascontiguousarray(nrm, np.float32),
                               uvs=None if uv is None else np.ascontiguousarray(uv, np.float32),
                               tangents=None if tan is None else np.ascontiguousarray(tan, np.float32), indices=idx,
                               material_id=mi if 0 <= mi < len(materials) else 0, flags=flags))
        mesh_map.append(ids)

    # ---- scene graph (load_node): NodeIDs in DFS pre-order over the default scene's roots
    nodes: List[Node] = []
    node_map = {}
    gnodes = j.get("nodes", [])
    scenes = j.get("scenes", [])
    if not scenes:
        raise GltfError("glTF has no scene")
    scene_index = j.get("scene", 0) if j.get("scene", -1) >= 0 else 0

    def load_node(gi, parent_global):
        gn = gnodes[gi]
        local = node_transform(gn)
        glob = local if parent_global is None else _mul4(parent_global, local)
        node_map[gi] = len(nodes)
        nodes.append(Node(meshes=list(mesh_map[gn["mesh"]]) if gn.get("mesh", -1) >= 0 else [], transform=glob))
        for c in gn.get("children", []):
            load_node(c, glob)

    for root in scenes[scene_index].get("nodes", []):
        load_node(root, None)

    # ---- camera (load_cameras): the first camera node becomes the active camera (add_camera activates the first one)
    camera: Optional[CameraDesc] = None
    for gi, gn in enumerate(gnodes):
        ci = gn.get("camera", -1)
        if ci < 0 or ci >= len(j.get("cameras", [])) or gi not in node_map:
            continue
        gc = j["cameras"][ci]
        if gc.get("type") != "perspective":
            continue
        M = nodes[node_map[gi]].transform
        f32 = np.float32
        up = M[:3, 1] / f32(np.sqrt(np.sum(M[:3, 1] * M[:3, 1], dtype=f32)))
        fwd = M[:3, 2] / f32(np.sqrt(np.sum(M[:3, 2] * M[:3, 2], dtype=f32)))
        eye = M[:3, 3]
        p = gc.get("perspective", {})
        aspect = float(p["aspectRatio"]) if p.get("aspectRatio", 0) > 0 else 1.0
        zfar = float(p["zfar"]) if p.get("zfar", 0) > 0 else 1000.0
        camera = CameraDesc(position=tuple(float(v) for v in eye), target=tuple(float(v) for v in (eye - fwd)), up=tuple(float(v) for v in up),
                            yfov_deg=float(np.rad2deg(f32(p.get("yfov", np.deg2rad(60.0))))), aspect=aspect, near=float(p.get("znear", 0.01)), far=zfar,
                            jitter_sequence="R2")
        break
    if camera is None:
        raise GltfError("glTF has no perspective camera: give one explicitly (the reference would fit a default camera to the scene's "
                        "bounding box, gltf_scene.cpp:760-771; SURVEY.md §8(c) asks for an explicit camera for parity)")
    return SceneDescription(meshes=meshes, nodes=nodes, materials=materials, textures=textures, env=None, camera=camera,
                            name=os.path.splitext(os.path.basename(str(path)))[0])


# ----------------------------------------------------------------------------------------------- writer
def _png_bytes(px):
    from PIL import Image
    b = io.BytesIO()
    Image.fromarray(np.ascontiguousarray(px, np.uint8), "RGBA").save(b, format="PNG")
    return b.getvalue()


def save_gltf(desc: SceneDescription, path, embed: bool = False):
    """Writes `desc` as glTF 2.0 (.gltf + .bin next to it, textures as PNG; `embed` = everything as data URIs).
    Node transforms are written as matrices; the camera becomes a camera node (looking down its -Z)."""
    path = str(path)
    base = os.path.splitext(path)[0]
    blob = bytearray()
    views, accessors = [], []

    def add_accessor(arr, ctype, typ, target=None, minmax=False):
        arr = np.ascontiguousarray(arr)
        while len(blob) % 4:
            blob.append(0)
        v = {"buffer": 0, "byteOffset": len(blob), "byteLength": arr.nbytes}
        if target:
            v["target"] = target
        blob.extend(arr.tobytes())
        views.append(v)
        a = {"bufferView": len(views) - 1, "componentType": ctype, "count": int(arr.shape[0]) if typ != "SCALAR" else int(arr.size), "type": typ}
        if minmax:
            a["min"], a["max"] = [float(x) for x in arr.min(0)], [float(x) for x in arr.max(0)]
        accessors.append(a)
        return len(accessors) - 1

    images, gtex, samplers = [], [], []
    tex_of = {}

    def gltf_texture(tid):
        if tid == abi.TEXTURE_NONE:
            return None
        if tid not in tex_of:
            t = desc.textures[tid]
            px = t.pixels() if callable(t.pixels) else t.pixels
            png = _png_bytes(px)
            if embed:
                images.append({"uri": "data:image/png;base64," + base64.b64encode(png).decode()})
            else:
                name = f"{os.path.basename(base)}_tex{tid}.png"
                with open(os.path.join(os.path.dirname(os.path.abspath(path)), name), "wb") as f:
                    f.write(png)
                images.append({"uri": name})
            inv = {v: k for k, v in _WRAP.items()}
            nearest = t.filter == abi.FILTER_NEAREST
            min_filter = {abi.MIPMAP_NONE: 9728 if nearest else 9729, abi.MIPMAP_NEAREST: 9984 if nearest else 9985,
                          abi.MIPMAP_LINEAR: 9986 if nearest else 9987}[getattr(t, "mipmap", abi.MIPMAP_NONE)]
            samplers.append({"magFilter": 9728 if nearest else 9729, "minFilter": min_filter,
                             "wrapS": inv[t.wrap_u], "wrapT": inv[t.wrap_v]})
            gtex.append({"source": len(images) - 1, "sampler": len(samplers) - 1})
            tex_of[tid] = len(gtex) - 1
        return {"index": tex_of[tid]}

    mats = []
    for m in desc.materials:
        if not isinstance(m, GltfMaterial):
            m = GltfMaterial(base_color_factor=tuple(m.base_color_factor), base_color_texture=m.base_color_texture, metallic_factor=0.0, roughness_factor=1.0)
        pbr = {"baseColorFactor": [float(v) for v in m.base_color_factor], "metallicFactor": float(m.metallic_factor), "roughnessFactor": float(m.roughness_factor)}
        g = {"pbrMetallicRoughness": pbr, "emissiveFactor": [float(v) for v in m.emissive_factor]}
        for key, tid, holder in (("baseColorTexture", m.base_color_texture, pbr), ("metallicRoughnessTexture", m.metallic_roughness_texture, pbr),
                                 ("normalTexture", m.normal_texture, g), ("emissiveTexture", m.emissive_texture, g), ("occlusionTexture", m.occlusion_texture, g)):
            ref = gltf_texture(tid)
            if ref is not None:
                holder[key] = ref
        if "normalTexture" in g:
            g["normalTexture"]["scale"] = float(m.normal_scale)
        ext = {}
        if m.ior != 1.5:
            ext["KHR_materials_ior"] = {"ior": float(m.ior)}
        if m.transmission is not None:
            ext["KHR_materials_transmission"] = {"transmissionFactor": float(m.transmission.get("weight", 0.0))}
        if m.volume is not None:
            ab = np.array(m.volume.get("absorption", (0, 0, 0)), np.float64)
            ext["KHR_materials_volume"] = {"thicknessFactor": 1.0, "attenuationDistance": 1.0, "attenuationColor": [float(v) for v in np.exp(-ab)]}
        if m.clearcoat is not None:
            c = m.clearcoat
            ext["KHR_materials_clearcoat"] = {"clearcoatFactor": float(c.get("weight", 0.0)), "clearcoatRoughnessFactor": float(c.get("roughness", 0.0))}
        if m.sheen is not None:
            s = m.sheen
            ext["KHR_materials_sheen"] = {"sheenColorFactor": [float(v) for v in s.get("color", (0, 0, 0))], "sheenRoughnessFactor": float(s.get("roughness", 0.0))}
        if m.iridescence is not None:
            i = m.iridescence
            ext["KHR_materials_iridescence"] = {"iridescenceFactor": float(i.get("factor", 0.0)), "iridescenceIor": float(i.get("ior", 1.3)),
                                                "iridescenceThicknessMinimum": float(i.get("thickness_min", 100.0)),
                                                "iridescenceThicknessMaximum": float(i.get("thickness_max", 400.0))}
        if m.anisotropy is not None:
            a = m.anisotropy
            ext["KHR_materials_anisotropy"] = {"anisotropyStrength": float(a.get("strength", 0.0)), "anisotropyRotation": float(a.get("rotation", 0.0))}
        if ext:
            g["extensions"] = ext
        mats.append(g)

    # one glTF primitive per Scene::Mesh; one glTF mesh per distinct mesh list of a node (a node instantiates all primitives of its
    # glTF mesh, gltf_scene.cpp:706-710), ordered by first mesh id so that load_meshes hands out the same MeshIDs again
    prims = []
    for i, m in enumerate(desc.meshes):
        at = {"POSITION": add_accessor(m.positions.astype(np.float32), 5126, "VEC3", 34962, minmax=True),
              "NORMAL": add_accessor(m.normals.astype(np.float32), 5126, "VEC3", 34962)}
        if m.uvs is not None:
            at["TEXCOORD_0"] = add_accessor(m.uvs.astype(np.float32), 5126, "VEC2", 34962)
        if m.tangents is not None:
            at["TANGENT"] = add_accessor(m.tangents.astype(np.float32), 5126, "VEC4", 34962)
        prim = {"attributes": at, "material": int(m.material_id), "mode": 4}
        if m.indices is not None:
            ct = {1: 5121, 2: 5123, 4: 5125}[m.indices.dtype.itemsize]
            prim["indices"] = add_accessor(np.ascontiguousarray(m.indices).reshape(-1), ct, "SCALAR", 34963)
        prims.append(prim)
        mg = mats[m.material_id]
        if not (m.flags & MESH_IS_OPAQUE):
            mg["alphaMode"] = "MASK"
        if (m.flags & MESH_TWO_SIDED) and "KHR_materials_transmission" not in mg.get("extensions", {}):
            mg["doubleSided"] = True
    lists = sorted({tuple(n.meshes) for n in desc.nodes if n.meshes}, key=lambda t: t[0])
    gmesh_of = {t: k for k, t in enumerate(lists)}
    gmeshes = [{"primitives": [prims[i] for i in t], "name": f"mesh{t[0]}"} for t in lists]
    flat = []
    for n in desc.nodes:
        gn = {"matrix": [float(v) for v in np.asarray(n.transform, np.float32).T.reshape(-1)]}
        if n.meshes:
            gn["mesh"] = gmesh_of[tuple(n.meshes)]
        flat.append(gn)
    roots = list(range(len(flat)))
    doc = {"asset": {"version": "2.0", "generator": "merian_b200.gltf"}, "scene": 0, "scenes": [{"nodes": roots}], "nodes": flat, "meshes": gmeshes,
           "materials": mats, "accessors": accessors, "bufferViews": views}
    used = sorted({e for g in mats for e in g.get("extensions", {})})
    if used:
        doc["extensionsUsed"] = used
    if gtex:
        doc["textures"], doc["images"], doc["samplers"] = gtex, images, samplers
    if desc.camera is not None:
        c = desc.camera
        pos, tgt, up = (np.array(v, np.float64) for v in (c.position, c.target, c.up))
        fwd = (tgt - pos) / np.linalg.norm(tgt - pos)
        right = np.cross(fwd, up)
        right /= np.linalg.norm(right)
        upv = np.cross(right, fwd)
        M = np.eye(4)
        M[:3, 0], M[:3, 1], M[:3, 2], M[:3, 3] = right, upv, -fwd, pos
        doc["cameras"] = [{"type": "perspective", "perspective": {"yfov": float(np.deg2rad(c.yfov_deg)), "aspectRatio": float(c.aspect), "znear": float(c.near),
                                                                   "zfar": float(c.far)}}]
        doc["nodes"].append({"camera": 0, "matrix": [float(v) for v in M.T.reshape(-1)]})
        doc["scenes"][0]["nodes"].append(len(doc["nodes"]) - 1)
    if embed:
        doc["buffers"] = [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(bytes(blob)).decode()}]
    else:
        with open(base + ".bin", "wb") as f:
            f.write(bytes(blob))
        doc["buffers"] = [{"byteLength": len(blob), "uri": os.path.basename(base) + ".bin"}]
    with open(path, "w") as f:
        json.dump(doc, f)
    return path
