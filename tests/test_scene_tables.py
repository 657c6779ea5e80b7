"""Host logic that defines what "instance id / geometry index / primitive id" mean:
mesh grouping -> BLAS (scene.cpp:1023-1105), GeometryData order (:1236-1291), TLAS instances (:1808-1857), vertex upload
path (transform_vertex.slang:12-40). Mirrors tests/test_scene.cpp's MeshGroupingStatic/Instanced cases and adds the id
assertions the reference's tests lack."""
import numpy as np

from merian_b200 import abi, synthetic
from merian_b200.scene import (MESH_FLIP_FACING, MESH_IS_OPAQUE, MESH_TWO_SIDED, CameraDesc, GltfMaterial, Mesh, Node,
                               SceneDescription, SceneTables, device_decode_normal, pack_vertices)


def _tri_mesh(offset=0.0, flags=MESH_IS_OPAQUE, mask=0xFF, mat=0, n_tris=1):
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]] * n_tris, np.float32) + np.float32(offset)
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (3 * n_tris, 1))
    return Mesh(positions=pos, normals=nrm, material_id=mat, flags=flags, instance_mask=mask)


def _T(tx=0.0, s=1.0):
    m = np.eye(4, dtype=np.float32) * s
    m[3, 3] = 1
    m[0, 3] = tx
    return m


def _scene(meshes, nodes):
    return SceneDescription(meshes=meshes, nodes=nodes, materials=[GltfMaterial(), GltfMaterial()], camera=CameraDesc((0, 0, 3), (0, 0, 0)))


def test_static_meshes_merge_into_one_blas():
    """MeshGroupingStatic (tests/test_scene.cpp:150): static single-instance meshes with equal (flip, two-sided, mask)
    share one BLAS, identity instance transform, vertices pretransformed by their node transform."""
    t = SceneTables(_scene([_tri_mesh(), _tri_mesh(), _tri_mesh()], [Node([0], _T(1.0)), Node([1], _T(2.0)), Node([2], _T(3.0, 2.0))]))
    assert len(t.groups) == 1 and t.groups[0]["meshes"] == [0, 1, 2]
    assert len(t.instances) == 1
    tr, iid, mask, flags, blas = t.instances[0]
    assert np.array_equal(tr.reshape(3, 4), np.eye(4, dtype=np.float32)[:3]) and iid == 0 and mask == 0xFF and blas == 0
    assert flags == abi.INSTANCE_FORCE_OPAQUE
    # geometry ids: group -> instance -> mesh in group; all pretransformed
    assert [(g[0], bool(g[2] & abi.GEOM_PRETRANSFORMED)) for g in t.geometries] == [(0, True), (1, True), (2, True)]
    assert np.allclose(t.mesh_vertices[0]["position"][1], [2, 0, 0])
    assert np.allclose(t.mesh_vertices[2]["position"][1], [5, 0, 0])  # scale 2 then translate 3


def test_group_split_by_flags_and_mask():
    """The group key is (FlipFacing | TwoSided, instance mask) (scene.hpp:350-351, scene.cpp:1046-1060); groups appear in
    first-seen order over ascending MeshID; instanceCustomIndex is the running sum of group sizes (scene.cpp:1823)."""
    meshes = [_tri_mesh(), _tri_mesh(flags=MESH_IS_OPAQUE | MESH_TWO_SIDED), _tri_mesh(), _tri_mesh(flags=MESH_IS_OPAQUE | MESH_FLIP_FACING),
              _tri_mesh(mask=0x01), _tri_mesh(flags=MESH_TWO_SIDED)]  # last one: two-sided and NOT opaque
    t = SceneTables(_scene(meshes, [Node([i]) for i in range(6)]))
    assert [g["meshes"] for g in t.groups] == [[0, 2], [1, 5], [3], [4]]
    assert [i[1] for i in t.instances] == [0, 2, 4, 5]
    assert [i[2] for i in t.instances] == [0xFF, 0xFF, 0xFF, 0x01]
    fl = [i[3] for i in t.instances]
    assert fl[0] == abi.INSTANCE_FORCE_OPAQUE
    assert fl[1] == abi.INSTANCE_FACING_CULL_DISABLE  # one non-opaque member => no ForceOpaque
    assert fl[2] == abi.INSTANCE_FLIP_FACING | abi.INSTANCE_FORCE_OPAQUE
    assert [g[0] for g in t.geometries] == [0, 2, 1, 5, 3, 4]
    assert bool(t.geometries[4][2] & abi.GEOM_FLIP_FACING) and bool(t.geometries[2][2] & abi.GEOM_TWO_SIDED)
    sd = t.scene_desc()
    assert sd.mesh_count == 6 and sd.blas_count == 4 and sd.instance_count == 4 and sd.geometry_count == 6


def test_instanced_meshes_group_by_instance_set():
    """MeshGroupingInstanced (tests/test_scene.cpp:176): meshes used by the same set of nodes share one BLAS that is
    instanced once per node (ascending NodeID) with the node transform; not pretransformed unless identity."""
    meshes = [_tri_mesh(), _tri_mesh(n_tris=2), _tri_mesh()]
    nodes = [Node([0, 1], _T(1.0)), Node([2]), Node([0, 1], _T(5.0)), Node([0, 1])]
    t = SceneTables(_scene(meshes, nodes))
    assert [g["meshes"] for g in t.groups] == [[0, 1], [2]]
    assert t.groups[0]["instances"] == [0, 2, 3] and t.groups[0]["instanced"]
    assert [i[1] for i in t.instances] == [0, 2, 4, 6] and [i[4] for i in t.instances] == [0, 0, 0, 1]
    assert t.instances[1][0].reshape(3, 4)[0, 3] == 5.0
    # GeometryData: (mesh, pretransformed?) in group -> instance -> mesh order
    assert [(g[0], bool(g[2] & abi.GEOM_PRETRANSFORMED), g[3]) for g in t.geometries] == \
        [(0, False, 1), (1, False, 2), (0, False, 1), (1, False, 2), (0, True, 1), (1, True, 2), (2, True, 1)]
    # instanced meshes keep object-space vertices
    assert np.allclose(t.mesh_vertices[0]["position"][1], [1, 0, 0])
    assert t.instance_transforms.shape == (4, 4, 4)
    assert np.allclose(t.inv_t_instance_transforms[1] @ t.instance_transforms[1].T, np.eye(4), atol=1e-6)


def test_vertex_upload_reencodes_normals():
    """A13: host encode (x32768, round half away) then the device pass decodes (int16/32767) and re-encodes
    (packSnorm2x16, x32767 ties-to-even); positions go through M; normals through the inverse transpose."""
    rng = np.random.default_rng(0)
    n = rng.normal(size=(64, 3)).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    p = rng.normal(size=(64, 3)).astype(np.float32)
    M = np.eye(4, dtype=np.float32)
    M[:3, :3] = np.diag([1.0, 2.0, 0.5]).astype(np.float32)
    M[:3, 3] = [1, 2, 3]
    v = pack_vertices(p, n, None, None, M)
    assert np.allclose(v["position"], p * np.array([1, 2, 0.5], np.float32) + np.array([1, 2, 3], np.float32), atol=1e-6)
    expect = n / np.array([1, 2, 0.5], np.float32)
    expect /= np.linalg.norm(expect, axis=1, keepdims=True)
    got = device_decode_normal(v["encoded_normal"])
    assert (np.sum(got * expect, axis=1) > 1 - 1e-4).all()
    vi = pack_vertices(p, n)
    assert np.array_equal(vi["position"], p)
    assert (vi["uv"] == 0).all() and (vi["encoded_tangent"] & 1 == 0).all()


def test_index_types_and_counts():
    d = synthetic.terrain(nx=8, ny=4, env=False)
    d.meshes[0].indices = d.meshes[0].indices.astype(np.uint16)
    t = SceneTables(d)
    assert t.triangle_count == 64 and t.mesh_index_flag[0] == abi.GEOM_INDEX_U16
    assert t.geometries[0][2] & abi.GEOM_INDEX_U16
    c = synthetic.cornell()
    tc = SceneTables(c)
    assert tc.triangle_count == 32 and tc.mesh_index_flag[0] == 0  # IndexType::None
    assert len(tc.groups) == 1 and len(tc.instances) == 1  # all static, same key => one BLAS
    assert [g[1] for g in tc.geometries] == [0, 1, 2, 3, 0, 0]


def test_synthetic_scenes_are_seeded():
    a, b = synthetic.terrain(nx=16, ny=8, seed=3, env=False), synthetic.terrain(nx=16, ny=8, seed=3, env=False)
    c = synthetic.terrain(nx=16, ny=8, seed=4, env=False)
    assert np.array_equal(a.meshes[0].positions, b.meshes[0].positions)
    assert not np.array_equal(a.meshes[0].positions, c.meshes[0].positions)
    r = synthetic.instanced_rocks(n_instances=9, rock_res=9, env_size=(16, 8))
    t = SceneTables(r)
    assert len(t.instances) == 10 and t.groups[0]["instanced"] and not t.groups[1]["instanced"]
