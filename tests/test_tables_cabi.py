"""include/mpt_tables.h (mptt_*): libmpt's C++ restatement of the merian::Scene host logic against the numpy
restatement in merian_b200/scene.py — two independent implementations of scene.cpp:1023-1105 / 1236-1291 / 1808-1857 and of the vertex
path gltf_scene.cpp:107-118 + transform_vertex.slang:12-40. Ids, flags, groups and instance records must be identical; vertex words are
identical up to the rounding of the matrix products (numpy's matmul does not promise an order of summation)."""
import numpy as np
import pytest

from merian_b200 import abi, synthetic
from merian_b200.scene import (MESH_FLIP_FACING, MESH_HAS_TANGENTS, MESH_IS_MORPHED, MESH_IS_OPAQUE, MESH_TWO_SIDED, DiffuseMaterial, Mesh, Node,
                               SceneDescription, SceneTables, pack_vertices)
from merian_b200.tables_cabi import CTables, pack_vertices as c_pack_vertices


def _snorm_delta(a, b):
    lo = np.abs((a & 0xFFFF).astype(np.uint16).view(np.int16).astype(np.int32) - (b & 0xFFFF).astype(np.uint16).view(np.int16).astype(np.int32))
    hi = np.abs((a >> 16).astype(np.uint16).view(np.int16).astype(np.int32) - (b >> 16).astype(np.uint16).view(np.int16).astype(np.int32))
    return np.maximum(lo, hi)


def _compare_vertices(c, p, exact):
    assert len(c) == len(p)
    assert np.array_equal(c["uv"], p["uv"])                                                    # float -> half, round to nearest even
    if exact:
        assert np.array_equal(c["position"], p["position"])
        assert np.array_equal(c["encoded_normal"], p["encoded_normal"]) and np.array_equal(c["encoded_tangent"], p["encoded_tangent"])
    else:
        np.testing.assert_allclose(c["position"], p["position"], rtol=3e-7, atol=4e-6)   # an ulp of the largest term of the row product
        assert _snorm_delta(c["encoded_normal"], p["encoded_normal"]).max() <= 1
        assert (c["encoded_tangent"] & 1 == p["encoded_tangent"] & 1).all()
        assert _snorm_delta(c["encoded_tangent"] & 0xFFFFFFFE, p["encoded_tangent"] & 0xFFFFFFFE).max() <= 2


def _compare_tables(desc, exact_vertices):
    p, c = SceneTables(desc), CTables(desc)
    assert c.groups == [g["meshes"] for g in p.groups]
    assert c.geometries == [tuple(int(v) for v in g) for g in p.geometries]
    assert len(c.instances) == len(p.instances)
    for (ct, cid, cmask, cfl, cgid), (pt, pid, pmask, pfl, pgid) in zip(c.instances, p.instances):
        assert np.array_equal(ct, pt) and (cid, cmask, cfl, cgid) == (pid, pmask, pfl, pgid)
    assert np.array_equal(c.instance_transforms, p.instance_transforms)
    np.testing.assert_allclose(c.inv_t_instance_transforms, p.inv_t_instance_transforms, rtol=2e-7, atol=1e-9)
    md = c.mesh_descs()
    for i, mesh in enumerate(desc.meshes):
        pv = p.mesh_vertices[i]
        if pv is None:
            assert md[i].vertex_count == 0 and not md[i].vertices
            continue
        assert md[i].index_flag == p.mesh_index_flag[i] and md[i].opaque == (1 if mesh.flags & MESH_IS_OPAQUE else 0)
        assert md[i].primitive_count == (len(mesh.indices) if mesh.indices is not None else len(mesh.positions) // 3)
        _compare_vertices(c.mesh_vertices(i), pv, exact_vertices)
        if mesh.indices is not None:
            n = md[i].primitive_count * 3
            got = np.frombuffer((__import__("ctypes").c_char * (n * mesh.indices.dtype.itemsize)).from_address(md[i].indices), dtype=mesh.indices.dtype, count=n)
            assert np.array_equal(got, np.ascontiguousarray(mesh.indices).reshape(-1))
    return p, c


@pytest.mark.parametrize("name", ["cornell", "terrain", "instanced_rocks"])
def test_tables_match_the_numpy_restatement_on_the_synthetic_scenes(name):
    desc = {"cornell": lambda: synthetic.cornell(), "terrain": lambda: synthetic.terrain(nx=60, ny=30, env_size=(64, 32)),
            "instanced_rocks": lambda: synthetic.instanced_rocks(n_instances=12, rock_res=9, seed=5, env_size=(64, 32))}[name]()
    # static scenes bake identity node transforms: every operation is exact; the rocks are instanced (identity pre-transform) as well
    _compare_tables(desc, exact_vertices=True)


def _quad(cx, flags, mask=0xFF, tangents=False, idx_dtype=np.uint16, mat=0):
    p = np.array([[cx - 0.4, -0.4, 0], [cx + 0.4, -0.4, 0], [cx + 0.4, 0.4, 0.2], [cx - 0.4, 0.4, 0.2]], np.float32)
    rng = np.random.default_rng(int(cx * 10) + 100)
    n = rng.normal(size=(4, 3)).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    t = None
    if tangents:
        t = np.concatenate([rng.normal(size=(4, 3)), np.array([[1], [-1], [1], [-1]])], 1).astype(np.float32)
    idx = None if idx_dtype is None else np.array([[0, 1, 2], [0, 2, 3]], idx_dtype)
    if idx is None:
        p, n = p[[0, 1, 2, 0, 2, 3]], n[[0, 1, 2, 0, 2, 3]]
        t = None if t is None else t[[0, 1, 2, 0, 2, 3]]
    uv = rng.random((len(p), 2)).astype(np.float32) * 3 - 1
    return Mesh(positions=p, normals=n, uvs=uv, tangents=t, indices=idx, material_id=mat, flags=flags | (MESH_HAS_TANGENTS if tangents else 0), instance_mask=mask)


def _xf(tx, ty, tz, angle, s=1.0):
    c, si = np.cos(angle) * s, np.sin(angle) * s
    return np.array([[c, -si, 0, tx], [si, c, 0, ty], [0, 0, s, tz], [0, 0, 0, 1]], np.float32)


def test_grouping_ids_and_flags_on_a_mixed_scene():
    """Static groups split by FlipFacing / TwoSided and by instance mask, instanced groups keyed by their instance set, an unreferenced
    mesh, a non-opaque mesh, every index width, identity and non-identity nodes."""
    meshes = [_quad(-2.0, MESH_IS_OPAQUE), _quad(-1.0, MESH_IS_OPAQUE | MESH_TWO_SIDED, tangents=True, idx_dtype=np.uint32),
              _quad(0.0, MESH_FLIP_FACING, idx_dtype=np.uint8, mat=1), _quad(1.0, MESH_IS_OPAQUE, mask=0x0F, idx_dtype=None),
              _quad(2.0, MESH_IS_OPAQUE, tangents=True), _quad(3.0, MESH_IS_OPAQUE | MESH_TWO_SIDED), _quad(4.0, MESH_IS_OPAQUE),
              _quad(5.0, MESH_IS_OPAQUE, mat=1)]
    nodes = [Node([0, 1], _xf(0.1, 0.2, 0.3, 0.4)), Node([2]), Node([3], _xf(0, 1, 0, -0.7, 1.5)),
             Node([4, 5], _xf(1, 0, 0, 0.2)), Node([4, 5], _xf(-1, 0, 2, 1.1, 0.5)), Node([4]), Node([7], _xf(0, 0, -1, 0.0))]   # mesh 6: unreferenced
    desc = SceneDescription(meshes=meshes, nodes=nodes, materials=[DiffuseMaterial(), DiffuseMaterial((0.5, 0.5, 0.5, 1))])
    p, c = _compare_tables(desc, exact_vertices=False)
    assert len(c.groups) >= 5 and any(len(g) > 1 for g in c.groups)
    flags = [g[2] for g in c.geometries]
    assert any(f & abi.GEOM_PRETRANSFORMED for f in flags) and any(not (f & abi.GEOM_PRETRANSFORMED) for f in flags)
    assert sum(1 for (_t, _i, _m, fl, _g) in c.instances if not (fl & abi.INSTANCE_FORCE_OPAQUE)) == 1   # the group of the non-opaque mesh 2
    c.close()


def test_pack_vertices_matches_numpy():
    rng = np.random.default_rng(3)
    n = 5000
    pos = (rng.random((n, 3)) * 20 - 10).astype(np.float32)
    nrm = rng.normal(size=(n, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm[:8] = [[0, 0, 1], [0, 0, -1], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0.70710678, 0, -0.70710678], [0, -0.6, -0.8]]
    uv = np.concatenate([rng.random((n - 6, 2)) * 8 - 4, [[0, 0], [1, 1], [65504, -65504], [1e-8, 6e-8], [0.33325195, 2049.0], [1e5, -1e-5]]]).astype(np.float32)
    tan = np.concatenate([rng.normal(size=(n, 3)), np.where(rng.random((n, 1)) < 0.5, -1.0, 1.0)], 1).astype(np.float32)
    tan[0, :3] = 0                                                                # degenerate tangent -> (0, 0, 1)
    _compare_vertices(c_pack_vertices(pos, nrm, uv, tan), pack_vertices(pos, nrm, uv, tan), exact=True)
    _compare_vertices(c_pack_vertices(pos, nrm), pack_vertices(pos, nrm), exact=True)
    M = _xf(0.3, -2.0, 5.0, 0.9, 1.7)
    M[2, 0], M[0, 2] = 0.2, -0.1
    _compare_vertices(c_pack_vertices(pos, nrm, uv, tan, M), pack_vertices(pos, nrm, uv, tan, M), exact=False)


def test_errors_are_reported_not_thrown():
    from merian_b200._lib import MptError
    bad = SceneDescription(meshes=[_quad(0.0, MESH_IS_OPAQUE | MESH_IS_MORPHED)], nodes=[Node([0])], materials=[DiffuseMaterial()])
    with pytest.raises(MptError):
        CTables(bad)
    out_of_range = SceneDescription(meshes=[_quad(0.0, MESH_IS_OPAQUE)], nodes=[Node([3])], materials=[DiffuseMaterial()])
    with pytest.raises(MptError):
        CTables(out_of_range)
    with pytest.raises(MptError):
        c_pack_vertices(np.zeros((3, 3), np.float32), np.ones((3, 3), np.float32), transform=np.zeros((4, 4), np.float32))   # singular


def test_c_tables_feed_the_host_accel_build():
    """The mpt_scene_desc of mptt_build goes straight into mpt_host_accel_build (what mpt_scene_upload runs on the host) and gives the
    acceleration structure the numpy tables give."""
    import ctypes as C
    from merian_b200._lib import lib
    from merian_b200.host_accel import HostAccel
    desc = synthetic.instanced_rocks(n_instances=12, rock_res=9, seed=5, env_size=(64, 32))
    ref = HostAccel(SceneTables(desc))
    c = CTables(desc)
    h = C.c_void_p()
    assert lib().mpt_host_accel_build(C.byref(c.scene_desc()), C.byref(h)) == abi.MPT_OK
    view = abi.HostAccelView()
    assert lib().mpt_host_accel_get(h, C.byref(view)) == abi.MPT_OK
    assert (view.node_count, view.triangle_count, view.prim_count) == (ref.view.node_count, ref.view.triangle_count, ref.view.prim_count)
    got = np.frombuffer((C.c_char * (view.node_count * 80)).from_address(view.nodes), dtype=np.uint8)
    assert np.array_equal(got, ref.nodes.view(np.uint8).reshape(-1))
    lib().mpt_host_accel_free(h)
    c.close()


def test_uv_half_conversion_matches_ieee_round_to_nearest_even():
    """half2(uv) of the vertex path (gltf_scene.cpp:113): the C++ float -> binary16 conversion against numpy's on every binary16 value, its
    two float32 neighbours, the midpoints between consecutive halves (the ties), the subnormal range, overflow and random bit patterns."""
    h = np.arange(0x10000, dtype=np.uint32).astype(np.uint16)
    exact = h.view(np.float16).astype(np.float32)
    exact = exact[np.isfinite(exact)]
    near = np.concatenate([np.nextafter(exact, np.float32(np.inf)), np.nextafter(exact, np.float32(-np.inf))])
    s = np.sort(exact)
    mid = ((s[:-1].astype(np.float64) + s[1:].astype(np.float64)) * 0.5).astype(np.float32)      # exactly representable in float32
    rng = np.random.default_rng(5)
    bits = rng.integers(0, 1 << 32, 400000, dtype=np.uint64).astype(np.uint32)
    rnd = bits.view(np.float32)
    rnd = rnd[np.isfinite(rnd)]
    small = (rng.random(100000).astype(np.float32) - np.float32(0.5)) * np.float32(2.5e-4)      # around the subnormal halves
    big = np.array([65504.0, 65519.9, 65520.0, 65536.0, 1e9, -65520.0, -1e9, 0.0, -0.0, 5.9604645e-08, 2.9802322e-08, 2.98023226e-08], np.float32)
    v = np.concatenate([exact, near, mid, rnd, small, big]).astype(np.float32)
    if len(v) % 2:
        v = v[:-1]
    uv = np.ascontiguousarray(v.reshape(-1, 2))
    n = len(uv)
    pos = np.zeros((n, 3), np.float32)
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (n, 1))
    got = c_pack_vertices(pos, nrm, uv)["uv"]
    with np.errstate(over="ignore"):
        ref = uv.astype(np.float16).view(np.uint16)
    assert np.array_equal(got, ref), np.argwhere(got != ref)[:5]


def _src_material(m):
    """A scene.py material as the mpt_src_material of include/mpt_tables.h."""
    from merian_b200.scene import GltfMaterial
    N = abi.TEXTURE_NONE
    s = abi.SrcMaterial()
    s.base_color_factor[:] = list(m.base_color_factor)
    s.base_color_texture = m.base_color_texture
    if not isinstance(m, GltfMaterial):
        s.model = abi.MATERIAL_DIFFUSE
        return s
    s.model = abi.MATERIAL_GLTF
    s.emissive_factor[:] = list(m.emissive_factor)
    s.metallic_factor, s.roughness_factor, s.normal_scale, s.occlusion_strength, s.ior = m.metallic_factor, m.roughness_factor, m.normal_scale, m.occlusion_strength, m.ior
    s.metallic_roughness_texture, s.normal_texture, s.occlusion_texture, s.emissive_texture = m.metallic_roughness_texture, m.normal_texture, m.occlusion_texture, m.emissive_texture
    s.features = m.features()
    if m.transmission is not None:
        s.transmission_weight = m.transmission.get("weight", 0.0)
    if m.volume is not None:
        s.volume_absorption[:] = list(m.volume.get("absorption", (0, 0, 0)))
    if m.clearcoat is not None:
        c = m.clearcoat
        s.clearcoat_weight, s.clearcoat_roughness, s.clearcoat_texture, s.clearcoat_roughness_texture = c.get("weight", 0.0), c.get("roughness", 0.0), c.get("texture", N), c.get("roughness_texture", N)
    if m.sheen is not None:
        c = m.sheen
        s.sheen_color[:] = list(c.get("color", (0, 0, 0)))
        s.sheen_roughness, s.sheen_color_texture, s.sheen_roughness_texture = c.get("roughness", 0.0), c.get("color_texture", N), c.get("roughness_texture", N)
    if m.iridescence is not None:
        c = m.iridescence
        s.iridescence_factor, s.iridescence_ior = c.get("factor", 0.0), c.get("ior", 1.3)
        s.iridescence_thickness_min, s.iridescence_thickness_max = c.get("thickness_min", 100.0), c.get("thickness_max", 400.0)
        s.iridescence_texture, s.iridescence_thickness_texture = c.get("texture", N), c.get("thickness_texture", N)
    if m.anisotropy is not None:
        c = m.anisotropy
        s.anisotropy_strength, s.anisotropy_rotation, s.anisotropy_texture = c.get("strength", 0.0), c.get("rotation", 0.0), c.get("texture", N)
    return s


def test_material_buffer_matches_the_numpy_restatement():
    """mptt_materials_build against scene.build_material_buffer (material_system.cpp:124-195, gltf_material.hpp:93-146): entry bytes, entry
    size and the type table in first-registration order, on a list that mixes both models and every extension block."""
    import ctypes as C
    from merian_b200._lib import lib
    from merian_b200.scene import GltfMaterial, build_material_buffer
    mats = [GltfMaterial(base_color_factor=(0.8, 0.3, 0.2, 1), metallic_factor=0.0, roughness_factor=0.5, emissive_factor=(0.1, 0.2, 0.3), base_color_texture=3),
            DiffuseMaterial((0.25, 0.5, 0.75, 1.0)),
            GltfMaterial(metallic_factor=1.0, roughness_factor=0.25, normal_texture=7, ior=1.33, clearcoat={"weight": 0.7, "roughness": 0.15, "texture": 2},
                         sheen={"color": (0.6, 0.5, 0.4), "roughness": 0.5, "roughness_texture": 9}),
            GltfMaterial(transmission={"weight": 0.9}, volume={"absorption": (0.1, 0.2, 0.3)}),
            GltfMaterial(iridescence={"factor": 0.6, "ior": 1.35, "thickness_min": 150.0, "thickness_max": 350.0, "thickness_texture": 4},
                         anisotropy={"strength": 0.4, "rotation": 1.2, "texture": 5}),
            DiffuseMaterial((1, 1, 1, 1), base_color_texture=1),
            GltfMaterial(base_color_factor=(0.1, 0.1, 0.1, 1), metallic_factor=0.3),                       # same type as the first
            GltfMaterial(transmission={"weight": 0.5}, volume={"absorption": (1, 2, 3)}, clearcoat={"weight": 1.0}, sheen={"color": (1, 1, 1)},
                         iridescence={"factor": 1.0}, anisotropy={"strength": 1.0})]
    ref_entries, ref_size, ref_types = build_material_buffer(mats)
    src = (abi.SrcMaterial * len(mats))(*[_src_material(m) for m in mats])
    h = C.c_void_p()
    assert lib().mptt_materials_build(C.cast(src, C.c_void_p), len(mats), C.byref(h)) == abi.MPT_OK
    view = abi.MaterialTableView()
    assert lib().mptt_materials_get(h, C.byref(view)) == abi.MPT_OK
    assert (view.entry_size, view.material_count, view.type_count) == (ref_size, len(mats), len(ref_types))
    got = bytes((C.c_char * (view.entry_size * view.material_count)).from_address(view.entries))
    assert got == ref_entries
    types = (abi.MaterialType * view.type_count).from_address(view.types)
    assert [(t.model, t.features) for t in types] == [tuple(t) for t in ref_types]
    lib().mptt_materials_free(h)
    bad = abi.SrcMaterial()
    bad.model = 7
    assert lib().mptt_materials_build(C.byref(bad), 1, C.byref(h)) != abi.MPT_OK


def test_camera_look_at_matches_the_python_mirror():
    """mptt_camera_look_at (Camera::write_to, camera.cpp:349-368) against scene.CameraDesc.to_abi: the same basis up to the last bit of
    tan / sqrt between libm and numpy."""
    import ctypes as C
    from merian_b200._lib import lib
    from merian_b200.scene import CameraDesc
    rng = np.random.default_rng(2)
    f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])
    for _ in range(20):
        pos, tgt = rng.normal(size=3) * 5, rng.normal(size=3) * 2
        fov, aspect = float(rng.uniform(20, 100)), float(rng.uniform(0.5, 2.5))
        ref = CameraDesc(position=tuple(pos), target=tuple(tgt), up=(0, 1, 0), yfov_deg=fov, aspect=aspect, near=0.05, far=500.0).to_abi(0)
        got = abi.Camera()
        assert lib().mptt_camera_look_at(f3(pos), f3(tgt), f3((0, 1, 0)), fov, aspect, 0.05, 500.0, 0.0, None, C.byref(got)) == abi.MPT_OK
        for name in ("position", "target", "up", "U", "V", "W"):
            np.testing.assert_allclose(np.array(getattr(got, name)[:]), np.array(getattr(ref, name)[:]), rtol=3e-6, atol=1e-6)
        assert (got.near_plane, got.far_plane, got.aperture_radius) == (ref.near_plane, ref.far_plane, 0.0) and abs(got.aspect_ratio - ref.aspect_ratio) < 1e-6
    bad = abi.Camera()
    assert lib().mptt_camera_look_at(f3((0, 0, 0)), f3((0, 0, 0)), f3((0, 1, 0)), 45.0, 1.0, 0.1, 10.0, 0.0, None, C.byref(bad)) != abi.MPT_OK
    assert lib().mptt_camera_look_at(f3((0, 0, 0)), f3((0, 2, 0)), f3((0, 1, 0)), 45.0, 1.0, 0.1, 10.0, 0.0, None, C.byref(bad)) != abi.MPT_OK
