"""include/mpt_tables.h mptt_gltf_*: libmpt's C++ glTF loader against the Python loader (merian_b200/gltf.py) — two independent
restatements of GLTFScene::load (gltf_scene.cpp) on the documents of tests/test_gltf.py and on the synthetic scenes written by
save_gltf (.gltf + .bin + PNG, and embedded). Meshes, flags, node transforms and order, materials (every extension gate), texture slots
and sampler state, camera: identical; the flat tables built from either are identical."""
import json

import numpy as np
import pytest

from merian_b200 import abi, gltf, gltf_cabi
from merian_b200._lib import MptError
from merian_b200.scene import GltfMaterial, SceneTables

import test_gltf as tg
from helpers import tables


def _same_material(a, b):
    assert type(a) is type(b)
    if not isinstance(a, GltfMaterial):
        assert a == b
        return
    for f in ("metallic_factor", "roughness_factor", "normal_scale", "occlusion_strength", "ior"):
        assert np.float32(getattr(a, f)) == np.float32(getattr(b, f)), f
    assert tuple(np.float32(a.base_color_factor)) == tuple(np.float32(b.base_color_factor)) and tuple(np.float32(a.emissive_factor)) == tuple(np.float32(b.emissive_factor))
    for f in ("base_color_texture", "metallic_roughness_texture", "normal_texture", "occlusion_texture", "emissive_texture"):
        assert getattr(a, f) == getattr(b, f), f
    assert a.features() == b.features()
    for ext in ("transmission", "volume", "clearcoat", "sheen", "iridescence", "anisotropy"):
        ea, eb = getattr(a, ext), getattr(b, ext)
        assert (ea is None) == (eb is None), ext
        if ea is None:
            continue
        for k, va in ea.items():
            vb = eb[k]
            if isinstance(va, (tuple, list)):
                np.testing.assert_allclose(np.float32(va), np.float32(vb), rtol=3e-7, atol=0)   # volume absorption: logf of libm vs numpy
            elif isinstance(va, float):
                assert np.float32(va) == np.float32(vb), (ext, k)
            else:
                assert va == vb, (ext, k)


def _same_scene(p, c):
    """p: the Python loader's SceneDescription, c: the C++ loader's."""
    assert len(p.meshes) == len(c.meshes) and len(p.nodes) == len(c.nodes) and len(p.materials) == len(c.materials) and len(p.textures) == len(c.textures)
    for a, b in zip(p.meshes, c.meshes):
        assert np.array_equal(a.positions, b.positions) and np.array_equal(a.normals, b.normals)
        for x, y in ((a.uvs, b.uvs), (a.tangents, b.tangents), (a.indices, b.indices)):
            assert (x is None) == (y is None)
            if x is not None:
                assert x.dtype == y.dtype and np.array_equal(x, y)
        assert (a.material_id, a.flags, a.instance_mask) == (b.material_id, b.flags, b.instance_mask)
    for a, b in zip(p.nodes, c.nodes):
        assert list(a.meshes) == list(b.meshes) and np.array_equal(np.asarray(a.transform, np.float32), b.transform)
    for a, b in zip(p.materials, c.materials):
        _same_material(a, b)
    for a, b in zip(p.textures, c.textures):
        pa = a.pixels() if callable(a.pixels) else a.pixels
        assert np.array_equal(pa, b.pixels) and (a.srgb, a.wrap_u, a.wrap_v, a.filter, getattr(a, "mipmap", 0)) == (b.srgb, b.wrap_u, b.wrap_v, b.filter, b.mipmap)
    assert (p.camera is None) == (c.camera is None)
    if p.camera is not None:
        for f in ("position", "target", "up"):
            np.testing.assert_allclose(np.float32(getattr(p.camera, f)), np.float32(getattr(c.camera, f)), rtol=0, atol=0)
        # yfov in degrees: the C++ side multiplies by 180 / pi in fp32 like glm::degrees (one rounding); numpy's rad2deg divides by pi / 180 (one ulp off at times)
        assert abs(np.float32(p.camera.yfov_deg) - np.float32(c.camera.yfov_deg)) <= np.spacing(np.float32(c.camera.yfov_deg)) and np.float32(p.camera.aspect) == np.float32(c.camera.aspect)
        assert p.camera.jitter_sequence == c.camera.jitter_sequence
        assert np.float32(p.camera.near) == np.float32(c.camera.near) and np.float32(p.camera.far) == np.float32(c.camera.far)
        ref = p.camera.to_abi(0)
        for f in ("U", "V", "W"):   # the basis mptt_camera_look_at computed
            np.testing.assert_allclose(np.array(getattr(c.c_camera, f)[:]), np.array(getattr(ref, f)[:]), rtol=3e-6, atol=1e-6)
    # and the flat tables the renderer consumes
    if p.camera is not None:
        tp, tc = SceneTables(p), SceneTables(c)
        assert tp.geometries == tc.geometries and [g["meshes"] for g in tp.groups] == [g["meshes"] for g in tc.groups]
        for va, vb in zip(tp.mesh_vertices, tc.mesh_vertices):
            assert (va is None) == (vb is None) and (va is None or va.tobytes() == vb.tobytes())
        if all(isinstance(m, GltfMaterial) and m.volume is None for m in p.materials):
            assert tp.material_entries == tc.material_entries and tp.material_types == tc.material_types


@pytest.mark.parametrize("scene,kw,embed", [("cornell", {}, False), ("cornell", {}, True), ("terrain", dict(nx=24, ny=12, env_size=(16, 8)), False),
                                            ("instanced_rocks", dict(n_instances=12, rock_res=9, env_size=(16, 8)), False),
                                            ("terrain", dict(nx=16, ny=8, patches=(2, 2), textured=True, texture_size=16, env_size=(16, 8)), False)])
def test_synthetic_scenes_load_identically(tmp_path, scene, kw, embed):
    t = tables(scene, **kw)
    path = gltf.save_gltf(t.desc, tmp_path / "scene.gltf", embed=embed)
    _same_scene(gltf.load_gltf(path), gltf_cabi.load_gltf(path))


def _write(tmp_path, doc, name="t.gltf"):
    p = tmp_path / name
    p.write_text(json.dumps(doc))
    return p


def test_material_extension_gates_and_mesh_flags(tmp_path):
    mats = [{"pbrMetallicRoughness": {"baseColorFactor": [0.5, 0.25, 0.125, 1.0], "metallicFactor": 0.25, "roughnessFactor": 0.75}, "emissiveFactor": [1, 2, 3],
             "doubleSided": True},
            {"alphaMode": "MASK", "extensions": {"KHR_materials_transmission": {"transmissionFactor": 0.0}, "KHR_materials_ior": {"ior": 1.33},
                                                   "KHR_materials_volume": {"thicknessFactor": 1.0, "attenuationDistance": 2.0, "attenuationColor": [0.5, 0.00001, 1.0]}}},
            {"extensions": {"KHR_materials_transmission": {"transmissionFactor": 0.6}, "KHR_materials_volume": {"thicknessFactor": 0.0, "attenuationDistance": 2.0},
                            "KHR_materials_clearcoat": {"clearcoatFactor": 0.0, "clearcoatRoughnessFactor": 0.3}, "KHR_materials_sheen": {"sheenColorFactor": [0, 0, 0]},
                            "KHR_materials_iridescence": {"iridescenceFactor": 0.5, "iridescenceIor": 1.4}, "KHR_materials_anisotropy": {"anisotropyStrength": 0.7, "anisotropyRotation": 0.3}}},
            {"extensions": {"KHR_materials_clearcoat": {"clearcoatFactor": 0.8}, "KHR_materials_sheen": {"sheenColorFactor": [0.1, 0.0, 0.0], "sheenRoughnessFactor": 0.4},
                            "KHR_materials_iridescence": {"iridescenceFactor": 0.0}, "KHR_materials_anisotropy": {"anisotropyStrength": 0.0}}}]
    for pm in (0, 1, 2, 3, 9):
        doc = tg._tiny_gltf(mats, primitives_material=pm)
        path = _write(tmp_path, doc, f"m{pm}.gltf")
        _same_scene(gltf.load_gltf(path), gltf_cabi.load_gltf(path))
    doc = tg._tiny_gltf([])        # no materials at all: the default material
    path = _write(tmp_path, doc, "nomat.gltf")
    _same_scene(gltf.load_gltf(path), gltf_cabi.load_gltf(path))


def test_node_transforms_camera_and_errors(tmp_path):
    extra = [{"rotation": [0.1825742, 0.3651484, 0.5477226, 0.7302967], "scale": [1, 2, 3], "translation": [4, 5, 6], "mesh": 0, "children": [3]},
             {"matrix": [1, 0, 0, 0, 0, 0, -1, 0, 0, 1, 0, 0, 7, 8, 9, 1], "mesh": 0}]
    doc = tg._tiny_gltf([{}], node={"translation": [1, 2, 3], "rotation": [0, 0.7071068, 0, 0.7071068]}, extra_nodes=extra)
    doc["scenes"][0]["nodes"] = [0, 2]
    doc["cameras"][0]["perspective"].update({"aspectRatio": 1.5, "zfar": 250.0})
    path = _write(tmp_path, doc)
    p, c = gltf.load_gltf(path), gltf_cabi.load_gltf(path)
    assert len(c.nodes) == 4
    _same_scene(p, c)
    # no camera: the C++ side reports it and leaves the choice to the host (the Python harness refuses)
    nocam = tg._tiny_gltf([{}], cameras=False)
    for n in nocam["nodes"]:
        n.pop("camera", None)
    assert gltf_cabi.load_gltf(_write(tmp_path, nocam, "nocam.gltf")).camera is None
    # malformed input is reported, not thrown across the ABI
    for bad in ("{", json.dumps({"scenes": []}), json.dumps(dict(tg._tiny_gltf([{}]), accessors=[{"bufferView": 0, "componentType": 5126, "count": 400, "type": "VEC3"}] * 2))):
        q = tmp_path / "bad.gltf"
        q.write_text(bad)
        with pytest.raises(MptError):
            gltf_cabi.load_gltf(q)
    with pytest.raises(MptError):
        gltf_cabi.load_gltf(tmp_path / "missing.gltf")


def test_glb_container_strided_accessors_and_textures(tmp_path):
    import base64
    import struct
    import io
    from PIL import Image
    # interleaved POSITION | NORMAL | TEXCOORD_0 (stride 32) + u32 indices + a PNG in a bufferView, all inside one GLB chunk
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], np.float32)
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (4, 1))
    uv = np.array([[0, 0], [1, 0], [0, 1], [1, 1]], np.float32)
    inter = np.concatenate([pos, nrm, uv], 1).astype(np.float32).tobytes()
    idx = np.array([0, 1, 2, 1, 3, 2], np.uint32).tobytes()
    buf = io.BytesIO()
    Image.fromarray(np.random.default_rng(0).integers(0, 256, (4, 4, 4), dtype=np.uint8), "RGBA").save(buf, "PNG")
    png = buf.getvalue()
    blob = inter + idx + png
    blob += b"\0" * (-len(blob) % 4)
    doc = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0, 1]}], "buffers": [{"byteLength": len(blob)}],
           "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": len(inter), "byteStride": 32}, {"buffer": 0, "byteOffset": len(inter), "byteLength": len(idx)},
                           {"buffer": 0, "byteOffset": len(inter) + len(idx), "byteLength": len(png)}],
           "accessors": [{"bufferView": 0, "byteOffset": 0, "componentType": 5126, "count": 4, "type": "VEC3"}, {"bufferView": 0, "byteOffset": 12, "componentType": 5126, "count": 4, "type": "VEC3"},
                         {"bufferView": 0, "byteOffset": 24, "componentType": 5126, "count": 4, "type": "VEC2"}, {"bufferView": 1, "componentType": 5125, "count": 6, "type": "SCALAR"}],
           "images": [{"bufferView": 2, "mimeType": "image/png"}, {"uri": "data:image/png;base64," + base64.b64encode(png).decode()}],
           "samplers": [{"magFilter": 9728, "minFilter": 9987, "wrapS": 33071, "wrapT": 33648}],
           "textures": [{"source": 0, "sampler": 0}, {"source": 1}],
           "materials": [{"pbrMetallicRoughness": {"baseColorTexture": {"index": 0}, "metallicRoughnessTexture": {"index": 0}}, "normalTexture": {"index": 1, "scale": 0.5},
                          "emissiveTexture": {"index": 0}, "occlusionTexture": {"index": 1, "strength": 0.25}}],
           "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "NORMAL": 1, "TEXCOORD_0": 2}, "indices": 3, "material": 0}]}],
           "nodes": [{"mesh": 0}, {"camera": 0, "translation": [0.5, 0.5, 3.0]}], "cameras": [{"type": "perspective", "perspective": {"yfov": 0.8, "znear": 0.1}}]}
    js = json.dumps(doc).encode()
    js += b" " * (-len(js) % 4)
    glb = struct.pack("<4sII", b"glTF", 2, 12 + 8 + len(js) + 8 + len(blob)) + struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(blob), 0x004E4942) + blob
    path = tmp_path / "t.glb"
    path.write_bytes(glb)
    p, c = gltf.load_gltf(path), gltf_cabi.load_gltf(path)
    assert len(c.textures) == 3 and [t.srgb for t in c.textures] == [True, False, False]   # (0, sRGB), (0, linear), (1, linear) in first-use order
    assert (c.textures[0].filter, c.textures[0].mipmap, c.textures[0].wrap_u, c.textures[0].wrap_v) == (abi.FILTER_NEAREST, abi.MIPMAP_LINEAR, abi.WRAP_CLAMP_TO_EDGE, abi.WRAP_MIRRORED_REPEAT)
    _same_scene(p, c)
