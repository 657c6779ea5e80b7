"""glTF ingestion: hand-written glTF documents against the rules of GLTFScene::load
(src/merian-shaders/scene/gltf_scene.cpp), and save -> load round trips of the synthetic scenes down to the flat tables."""
import base64
import json

import numpy as np
import pytest

from merian_b200 import abi, gltf, synthetic
from merian_b200.scene import MESH_HAS_TANGENTS, MESH_IS_OPAQUE, MESH_TWO_SIDED, GltfMaterial, SceneTables

from helpers import tables


def _tables_equal(a: SceneTables, b: SceneTables):
    assert a.geometries == b.geometries
    assert len(a.instances) == len(b.instances)
    for (ta, ia, ma, fa, ga), (tb, ib, mb, fb, gb) in zip(a.instances, b.instances):
        assert np.array_equal(ta, tb) and (ia, ma, fa, ga) == (ib, mb, fb, gb)
    assert [g["meshes"] for g in a.groups] == [g["meshes"] for g in b.groups]
    for va, vb in zip(a.mesh_vertices, b.mesh_vertices):
        assert (va is None) == (vb is None)
        if va is not None:
            assert va.tobytes() == vb.tobytes()
    for xa, xb in zip(a.mesh_indices, b.mesh_indices):
        assert (xa is None) == (xb is None)
        if xa is not None:
            assert xa.dtype == xb.dtype and np.array_equal(xa, xb)
    assert a.material_entries == b.material_entries and a.material_entry_size == b.material_entry_size and a.material_types == b.material_types
    assert np.array_equal(a.instance_transforms, b.instance_transforms)


@pytest.mark.parametrize("scene,kw,embed", [("cornell", {}, False), ("cornell", {}, True), ("terrain", dict(nx=24, ny=12, env_size=(16, 8)), False),
                                            ("instanced_rocks", dict(n_instances=12, rock_res=9, env_size=(16, 8)), False),
                                            ("terrain", dict(nx=16, ny=8, patches=(2, 2), textured=True, texture_size=16, env_size=(16, 8)), False)])
def test_round_trip_reproduces_the_flat_tables(tmp_path, scene, kw, embed):
    t = tables(scene, **kw)
    p = gltf.save_gltf(t.desc, tmp_path / "scene.gltf", embed=embed)
    d2 = gltf.load_gltf(p)
    t2 = SceneTables(d2)
    # glTF cannot say "instance mask": the loader applies Scene::Mesh's default (scene.hpp:193) like the reference does
    if all(m.instance_mask == 0x01 for m in t.desc.meshes):
        _tables_equal(t, t2)
    else:
        assert t2.geometries == t.geometries and [g["meshes"] for g in t2.groups] == [g["meshes"] for g in t.groups]
    # camera: same basis up to fp32 rounding of the look-at round trip
    c1, c2 = t.desc.camera.to_abi(0), d2.camera.to_abi(0)
    # (the loader puts the target one unit in front of the eye, gltf_scene.cpp:736: U, V, W scale with the focus distance, rays do not)
    np.testing.assert_allclose(np.array(c2.position[:]), np.array(c1.position[:]), rtol=1e-6, atol=1e-6)
    for k in ("U", "V", "W"):
        np.testing.assert_allclose(np.array(getattr(c2, k)[:]) / np.linalg.norm(c2.W[:]), np.array(getattr(c1, k)[:]) / np.linalg.norm(c1.W[:]), rtol=2e-5, atol=2e-5)
    assert d2.camera.jitter_sequence == "R2"  # gltf_scene.cpp:773-774
    # textures survive the PNG round trip bit for bit, in first-use order, sRGB/linear as the material slot dictates
    assert len(d2.textures) == len(t.desc.textures)
    for a, b in zip(t.desc.textures, d2.textures):
        pa = a.pixels() if callable(a.pixels) else a.pixels
        assert np.array_equal(pa, b.pixels) and a.srgb == b.srgb and (a.wrap_u, a.wrap_v, a.filter) == (b.wrap_u, b.wrap_v, b.filter)


def _tiny_gltf(materials, primitives_material=0, node=None, extra_nodes=(), cameras=True, alpha=None):
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], np.float32)
    idx = np.array([0, 1, 2, 1, 3, 2], np.uint16)
    blob = pos.tobytes() + idx.tobytes()
    doc = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}],
           "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}],
           "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 48}, {"buffer": 0, "byteOffset": 48, "byteLength": 12}],
           "accessors": [{"bufferView": 0, "componentType": 5126, "count": 4, "type": "VEC3"}, {"bufferView": 1, "componentType": 5123, "count": 6, "type": "SCALAR"}],
           "meshes": [{"primitives": [{"attributes": {"POSITION": 0}, "indices": 1, "material": primitives_material}, {"attributes": {"POSITION": 0}, "mode": 1},
                                      {"attributes": {"POSITION": 0}}]}],
           "materials": materials, "nodes": [dict(node or {}, mesh=0, children=[1])]}
    doc["nodes"].append({"camera": 0, "translation": [0.5, 0.5, 3.0]})
    doc["nodes"] += list(extra_nodes)
    if cameras:
        doc["cameras"] = [{"type": "perspective", "perspective": {"yfov": 0.8, "znear": 0.1}}]
    return doc


def _load(tmp_path, doc):
    p = tmp_path / "t.gltf"
    p.write_text(json.dumps(doc))
    return gltf.load_gltf(p)


def test_material_extension_gates_and_mesh_flags(tmp_path):
    mats = [{"pbrMetallicRoughness": {"baseColorFactor": [0.5, 0.25, 0.125, 1], "metallicFactor": 0.25, "roughnessFactor": 0.75}, "emissiveFactor": [1, 2, 3],
             "alphaMode": "MASK", "doubleSided": True,
             "extensions": {"KHR_materials_transmission": {"transmissionFactor": 0.0}, "KHR_materials_ior": {"ior": 1.33},
                            "KHR_materials_volume": {"thicknessFactor": 1.0, "attenuationDistance": 2.0, "attenuationColor": [0.5, 1.0, 0.0]},
                            "KHR_materials_clearcoat": {"clearcoatFactor": 0.0, "clearcoatRoughnessFactor": 0.3},
                            "KHR_materials_sheen": {"sheenColorFactor": [0, 0, 0.1], "sheenRoughnessFactor": 0.4},
                            "KHR_materials_iridescence": {"iridescenceFactor": 0.7},
                            "KHR_materials_anisotropy": {"anisotropyRotation": 1.0}}}]
    d = _load(tmp_path, _tiny_gltf(mats))
    m = d.materials[0]
    assert m.base_color_factor == (0.5, 0.25, 0.125, 1.0) and m.metallic_factor == 0.25 and m.roughness_factor == 0.75 and m.emissive_factor == (1.0, 2.0, 3.0)
    assert m.ior == pytest.approx(1.33)
    assert m.transmission is None          # weight 0 => no transmission payload (gltf_scene.cpp:352-357)
    assert m.clearcoat is None             # weight 0 (:409-411)
    assert m.anisotropy is None            # strength 0 (:489-491)
    assert m.sheen is not None and m.sheen["roughness"] == pytest.approx(0.4)     # any colour channel > 0 (:437-439)
    assert m.iridescence == {"factor": 0.7, "ior": 1.3, "thickness_min": 100.0, "thickness_max": 400.0, "texture": abi.TEXTURE_NONE,
                             "thickness_texture": abi.TEXTURE_NONE}
    # sigma_a = -ln(clamp(c, 1e-4, 1)) / distance (:380-386)
    np.testing.assert_allclose(m.volume["absorption"], [-np.log(0.5) / 2, 0.0, -np.log(1e-4) / 2], rtol=1e-6)
    assert m.features() == abi.GLTF_VOLUME | abi.GLTF_SHEEN | abi.GLTF_IRIDESCENCE
    # primitives: LINES skipped, the one without material is opaque + default material 0 (:652-668)
    assert len(d.meshes) == 2
    assert d.meshes[0].flags == MESH_TWO_SIDED                       # MASK => not opaque; doubleSided (or transmission ext present)
    assert d.meshes[1].flags == MESH_IS_OPAQUE and d.meshes[1].material_id == 0
    assert d.meshes[1].indices is None and d.meshes[0].indices.dtype == np.uint16 and d.meshes[0].indices.shape == (2, 3)
    assert np.array_equal(d.meshes[0].normals, np.tile([[0, 1, 0]], (4, 1)))  # missing NORMAL => (0,1,0) (:130-133)
    assert d.meshes[0].uvs is None and d.meshes[0].tangents is None
    assert d.nodes[0].meshes == [0, 1]
    # transmission extension present => TwoSided even when not doubleSided (:659-663)
    d2 = _load(tmp_path, _tiny_gltf([{"extensions": {"KHR_materials_transmission": {"transmissionFactor": 0.5}}}]))
    assert d2.meshes[0].flags == MESH_IS_OPAQUE | MESH_TWO_SIDED and d2.materials[0].transmission == {"weight": 0.5}
    # no materials at all => one default GltfMaterial (:530-535)
    doc = _tiny_gltf([])
    del doc["materials"]
    d3 = _load(tmp_path, doc)
    assert d3.materials == [GltfMaterial()] and all(m.flags == MESH_IS_OPAQUE for m in d3.meshes)


def test_node_transforms_and_camera(tmp_path):
    q = np.array([0.1, 0.2, 0.3, 0.9], np.float64)
    q /= np.linalg.norm(q)
    node = {"translation": [1, 2, 3], "rotation": q.tolist(), "scale": [2, 3, 4]}
    d = _load(tmp_path, _tiny_gltf([{}], node=node))
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    M = np.eye(4)
    M[:3, :3] = R @ np.diag([2, 3, 4])
    M[:3, 3] = [1, 2, 3]
    np.testing.assert_allclose(d.nodes[0].transform, M, rtol=1e-6, atol=1e-6)   # T * R * S (gltf_scene.cpp:195)
    # the camera node is a child: its global transform is parent * local; eye = translation column, looking down local -Z
    eye = M @ np.array([0.5, 0.5, 3.0, 1.0])
    np.testing.assert_allclose(d.camera.position, eye[:3], rtol=1e-5)
    fwd = -(M[:3, 2] / np.linalg.norm(M[:3, 2]))
    np.testing.assert_allclose(np.array(d.camera.target) - np.array(d.camera.position), fwd, rtol=1e-5, atol=1e-6)
    assert d.camera.yfov_deg == pytest.approx(np.rad2deg(0.8), rel=1e-6) and d.camera.aspect == 1.0 and d.camera.near == pytest.approx(0.1) and d.camera.far == 1000.0
    # a column-major matrix node
    Mc = np.arange(16, dtype=np.float32).reshape(4, 4)
    d2 = _load(tmp_path, _tiny_gltf([{}], node={"matrix": Mc.T.reshape(-1).tolist()}))
    assert np.array_equal(d2.nodes[0].transform, Mc)
    with pytest.raises(gltf.GltfError, match="camera"):
        _load(tmp_path, _tiny_gltf([{}], cameras=False))


def test_glb_container_and_strided_accessors(tmp_path):
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    nrm = np.array([[0, 0, 1]] * 3, np.float32)
    inter = np.concatenate([pos, nrm], 1).astype(np.float32)  # interleaved, byteStride 24
    blob = inter.tobytes()
    doc = {"asset": {"version": "2.0"}, "scenes": [{"nodes": [0, 1]}], "buffers": [{"byteLength": len(blob)}],
           "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": len(blob), "byteStride": 24}],
           "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"},
                         {"bufferView": 0, "byteOffset": 12, "componentType": 5126, "count": 3, "type": "VEC3"}],
           "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "NORMAL": 1}}]}], "nodes": [{"mesh": 0}, {"camera": 0}],
           "cameras": [{"type": "perspective", "perspective": {"yfov": 1.0, "znear": 0.01, "zfar": 50.0, "aspectRatio": 2.0}}]}
    js = json.dumps(doc).encode()
    js += b" " * (-len(js) % 4)
    bb = blob + b"\0" * (-len(blob) % 4)
    import struct
    glb = struct.pack("<4sII", b"glTF", 2, 12 + 8 + len(js) + 8 + len(bb)) + struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(bb), 0x004E4942) + bb
    p = tmp_path / "t.glb"
    p.write_bytes(glb)
    d = gltf.load_gltf(p)
    assert np.array_equal(d.meshes[0].positions, pos) and np.array_equal(d.meshes[0].normals, nrm)
    assert d.camera.aspect == 2.0 and d.camera.far == 50.0


@pytest.mark.gpu
def test_render_from_gltf_file_matches_render_from_generator(gpu_ctx, tmp_path):
    """C1 end to end: the procedural Cornell box written as glTF, loaded back, uploaded and rendered gives the image of the
    in-memory scene (tables are identical, so the images are bit-identical)."""
    import torch
    from merian_b200.scene import upload_scene
    from helpers import render_params
    t = tables("cornell")
    t2 = SceneTables(gltf.load_gltf(gltf.save_gltf(t.desc, tmp_path / "cornell.gltf")))
    cam = t.desc.camera.to_abi(0)
    P = render_params(96, 96, 8)
    upload_scene(gpu_ctx, t)
    a = gpu_ctx.render_device(cam, P)
    upload_scene(gpu_ctx, t2)
    b = gpu_ctx.render_device(cam, P)
    assert torch.equal(a, b)
