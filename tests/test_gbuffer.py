"""GBuffer (Raytraced) node: the oracle's restatement of gbuffer_rt.slang against independent closed forms,
and the CUDA kernel (mpt_gbuffer) against the oracle."""
import numpy as np
import pytest

from merian_b200 import synthetic
from merian_b200.scene import MESH_IS_OPAQUE, CameraDesc, DiffuseMaterial, GltfMaterial, Mesh, Node, SceneDescription, SceneTables, upload_scene

from helpers import compare_hits, make_oracle, tables

MISS = 0xFFFFFFFF


def _half2(word):
    w = np.ascontiguousarray(word, dtype=np.uint32)
    return np.stack([(w & 0xFFFF).astype(np.uint16).view(np.float16), (w >> 16).astype(np.uint16).view(np.float16)], -1).astype(np.float32)


def _decode_normal(enc):
    from merian_b200.scene import device_decode_normal
    e = np.ascontiguousarray(enc, dtype=np.uint32)
    return device_decode_normal(e.reshape(-1)).reshape(e.shape + (3,))


def _primary_dirs(cam, w, h):
    px, py = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    nx = 2 * (px + 0.5 + cam.jitter[0]) / w - 1
    ny = -2 * (py + 0.5 + cam.jitter[1]) / h + 1
    U, V, W = (np.array(v[:], np.float32) for v in (cam.U, cam.V, cam.W))
    d = nx[..., None] * U + ny[..., None] * V + W
    return d / np.linalg.norm(d, axis=-1, keepdims=True)


def _project(cam, pos, w, h):
    """camera.slang:58-70 in float64."""
    U, V, W, P = (np.array(v[:], np.float64) for v in (cam.U, cam.V, cam.W, cam.position))
    d = pos.astype(np.float64) - P
    t = (W @ W) / (d @ W)
    plane = t[..., None] * d - W
    ndc = np.stack([(plane @ U) / (U @ U), (plane @ V) / (V @ V)], -1)
    p = (ndc - np.array([-1.0, 1.0])) / np.array([2.0, -2.0])
    return p * np.array([w, h]) - 0.5


def _material_scene():
    """Three glTF quads (dielectric, metal, clearcoat + sheen) and a diffuse-model quad in front of the sky."""
    def quad(cx, mat):
        s = 0.45
        p = np.array([[cx - s, -s, 0], [cx + s, -s, 0], [cx + s, s, 0.3], [cx - s, s, 0.3]], np.float32)
        n = np.tile(np.array([[0, -0.287348, 0.957826]], np.float32), (4, 1))
        uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)
        return Mesh(positions=p, normals=n, uvs=uv, indices=np.array([[0, 1, 2], [0, 2, 3]], np.uint16), material_id=mat, flags=MESH_IS_OPAQUE)
    mats = [GltfMaterial(base_color_factor=(0.8, 0.3, 0.2, 1), metallic_factor=0.0, roughness_factor=0.5, emissive_factor=(0.1, 0.2, 0.3)),
            GltfMaterial(base_color_factor=(0.9, 0.8, 0.3, 1), metallic_factor=1.0, roughness_factor=0.25),
            GltfMaterial(base_color_factor=(0.2, 0.5, 0.7, 1), metallic_factor=0.2, roughness_factor=0.6, clearcoat={"weight": 0.7, "roughness": 0.15},
                         sheen={"color": (0.6, 0.5, 0.4), "roughness": 0.5}, iridescence={"factor": 0.6, "ior": 1.35, "thickness_min": 150.0, "thickness_max": 350.0}),
            DiffuseMaterial((0.25, 0.5, 0.75, 1.0))]
    meshes = [quad(-1.5 + i, i) for i in range(4)]
    env = synthetic.sky_env(64, 32, sun_peak=50.0, sun_sigma_deg=10.0, sun_dir=(0.2, 0.3, 1.0))
    return SceneDescription(meshes=meshes, nodes=[Node([i]) for i in range(4)], materials=mats, env=env,
                            camera=CameraDesc(position=(0.1, 0.2, 3.0), target=(0.0, 0.0, 0.0), yfov_deg=55.0, aspect=1.5, near=0.01, far=100.0))


# ------------------------------------------------------------------------------------------------- oracle vs closed forms
def test_oracle_gbuffer_static_view(oracle_lib):
    t = tables("cornell")
    o = make_oracle(t)
    cam = t.desc.camera.to_abi(0)
    w = h = 96
    g = o.gbuffer(cam, w, h)
    vis = o.primary_visibility(cam, w, h)
    assert np.array_equal(g["tex1"], vis)                       # store_hit, gbuffer.slang:167-169
    hit = vis["instance_id"] != MISS
    assert hit.mean() > 0.9
    # same camera for both frames: nothing moves on screen (mv = projection round trip - pixel) and depths agree
    assert np.abs(g["tex2"][..., :2]).max() < 2e-3
    assert (g["tex2"][..., 3][hit] == 0).all()
    assert (g["tex3"][..., 3] == 1).all() and (g["emission"][..., 3] == 0).all()
    # linear_z is the distance along the primary ray: o + d * z lies on the hit triangle's plane
    d = _primary_dirs(cam, w, h)
    z = g["tex2"][..., 2]
    pos = np.array(cam.position[:], np.float32) + d * z[..., None]
    assert np.abs(pos[hit]).max() < 1.01 * np.abs(np.concatenate([m.positions for m in t.desc.meshes])).max()
    # grad_z = screen-space derivative of linear_z (hit.slang:97-103): compare with finite differences inside one triangle
    gz = _half2(g["tex0"][..., 1])
    tri_key = vis["primitive_id"].astype(np.int64) | (vis["geometry_index"].astype(np.int64) << 32) | (vis["instance_index"].astype(np.int64) << 48)
    same_x = hit[:, :-1] & hit[:, 1:] & (tri_key[:, :-1] == tri_key[:, 1:])
    fd_x = (z[:, 1:] - z[:, :-1])[same_x]
    assert np.abs(gz[:, :-1, 0][same_x] - fd_x).max() < 2e-3 + 2e-3 * np.abs(fd_x).max()
    same_y = hit[:-1] & hit[1:] & (tri_key[:-1] == tri_key[1:])
    fd_y = (z[1:] - z[:-1])[same_y]
    assert np.abs(gz[:-1, :, 1][same_y] - fd_y).max() < 2e-3 + 2e-3 * np.abs(fd_y).max()
    # Lambert walls: albedo = base colour (brdf-lambert-diffuse.slang get_albedo); the light is the only emitter
    mats = t.desc.materials
    geo_mat = np.array([mat for (_mesh, mat, _flags, _prims) in t.geometries])
    mat_of = geo_mat[(vis["instance_id"] + vis["geometry_index"])[hit]]
    em = g["emission"][hit][:, :3]
    for mid in np.unique(mat_of):
        sel = mat_of == mid
        m = mats[mid]
        ef = np.array(getattr(m, "emissive_factor", (0, 0, 0)), np.float32) * np.float32(getattr(m, "emissive_strength", 1.0))
        np.testing.assert_allclose(em[sel], np.broadcast_to(ef, em[sel].shape), rtol=1e-6)
    # normals: unit length after decoding, facing the camera for these one-sided walls
    n = _decode_normal(g["tex0"][..., 0])
    assert np.abs(np.linalg.norm(n, axis=-1) - 1).max() < 1e-5
    assert ((n * -d).sum(-1)[hit] > -1e-3).all()


def test_oracle_gbuffer_moving_camera_and_miss(oracle_lib):
    t = SceneTables(_material_scene())
    o = make_oracle(t)
    w, h = 120, 80
    cam = t.desc.camera.to_abi(0)
    prev_desc = CameraDesc(position=(0.25, 0.1, 3.2), target=(0.05, 0.0, 0.0), yfov_deg=55.0, aspect=1.5, near=0.01, far=100.0)
    prev = prev_desc.to_abi(0)
    g = o.gbuffer(cam, w, h, prev_cam=prev)
    hit = g["tex1"]["instance_id"] != MISS
    assert 0.2 < hit.mean() < 0.9
    d = _primary_dirs(cam, w, h)
    z = g["tex2"][..., 2]
    P = np.array(cam.position[:], np.float32)
    pos = P + d * z[..., None]
    px, py = np.meshgrid(np.arange(w), np.arange(h))
    mv_ref = _project(prev, pos, w, h) - np.stack([px, py], -1)
    assert np.abs(g["tex2"][..., :2][hit] - mv_ref[hit]).max() < 5e-3       # compute_motion_vector, scene.slang:404-414
    dz_ref = np.linalg.norm(pos - np.array(prev.position[:], np.float32), axis=-1) - z
    assert np.abs(g["tex2"][..., 3][hit] - dz_ref[hit]).max() < 1e-4        # delta_z, gbuffer_rt.slang:38
    # miss: the environment shading point (gbuffer_rt.slang:82-93): depth clamps to the f16 ceiling, albedo = emission = sky radiance,
    # normal = -ray direction, no depth gradient
    miss = ~hit
    assert (z[miss] == 65504.0).all() and (g["tex0"][..., 1][miss] == 0).all()
    assert np.array_equal(g["tex3"][miss][:, :3], g["emission"][miss][:, :3]) and (g["emission"][miss][:, :3] > 0).all()
    n = _decode_normal(g["tex0"][..., 0])
    assert np.abs(n[miss] + d[miss]).max() < 2e-4
    # directional albedo: the conductor's is tinted by its base colour, a dielectric's lies between base colour and 1
    geo_mat = np.array([mat for (_mesh, mat, _flags, _prims) in t.geometries])
    mat_of = np.where(hit, geo_mat[np.where(hit, g["tex1"]["instance_id"] + g["tex1"]["geometry_index"], 0)], -1)
    alb = g["tex3"][..., :3]
    a_metal = alb[hit & (mat_of == 1)].mean(0)
    assert a_metal[0] > a_metal[1] > a_metal[2] and 0.2 < a_metal[2] < 0.5
    a_die = alb[hit & (mat_of == 0)]
    assert (a_die >= np.array([0.8, 0.3, 0.2]) - 1e-4).all() and (a_die <= 1.0).all()
    assert np.allclose(alb[hit & (mat_of == 3)], [0.25, 0.5, 0.75])
    np.testing.assert_allclose(g["emission"][hit & (mat_of == 0)][:, :3], np.broadcast_to([0.1, 0.2, 0.3], (int((hit & (mat_of == 0)).sum()), 3)), rtol=1e-6)


# ------------------------------------------------------------------------------------------------- CUDA vs oracle
def _compare_gbuffers(got, ref):
    mism, bary_err, same = compare_hits(got["tex1"], ref["tex1"])
    assert mism == 0.0 and bary_err == 0.0
    hit = ref["tex1"]["instance_id"] != MISS
    # encoded normals: identical up to one snorm16 step per component (fp32 contraction differs between nvcc and gcc)
    a, b = got["tex0"][..., 0], ref["tex0"][..., 0]
    dx = np.abs((a & 0xFFFF).astype(np.int16).astype(np.int32) - (b & 0xFFFF).astype(np.int16).astype(np.int32))
    dy = np.abs((a >> 16).astype(np.int16).astype(np.int32) - (b >> 16).astype(np.int16).astype(np.int32))
    assert max(dx.max(), dy.max()) <= 1 and (a != b).mean() < 0.02
    gz_g, gz_r = _half2(got["tex0"][..., 1]), _half2(ref["tex0"][..., 1])
    assert np.abs(gz_g - gz_r).max() <= 2e-3 * np.maximum(np.abs(gz_r).max(), 1.0)
    np.testing.assert_allclose(got["tex2"][..., 2], ref["tex2"][..., 2], rtol=2e-6)
    assert np.abs(got["tex2"][..., :2][hit] - ref["tex2"][..., :2][hit]).max() < 2e-3
    assert np.abs(got["tex2"][..., 3][hit] - ref["tex2"][..., 3][hit]).max() < 1e-4 * max(1.0, float(ref["tex2"][..., 2][hit].max()))
    np.testing.assert_allclose(got["tex3"], ref["tex3"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(got["emission"], ref["emission"], rtol=2e-4, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("scene", ["cornell", "materials", "terrain", "rocks"])
def test_cuda_gbuffer_matches_oracle(gpu_ctx, scene):
    if scene == "cornell":
        t, (w, h) = tables("cornell"), (128, 128)
    elif scene == "materials":
        t, (w, h) = SceneTables(_material_scene()), (150, 100)
    elif scene == "terrain":
        t, (w, h) = tables("terrain", nx=200, ny=100, env_size=(256, 128)), (240, 136)
    else:
        t, (w, h) = tables("instanced_rocks", n_instances=60, rock_res=17, env_size=(128, 64)), (160, 90)
    upload_scene(gpu_ctx, t)
    o = make_oracle(t)
    cam = t.desc.camera.to_abi(3)
    cd = t.desc.camera
    prev = CameraDesc(position=tuple(np.array(cd.position) + [0.05, 0.02, -0.03]), target=cd.target, up=cd.up, yfov_deg=cd.yfov_deg, aspect=cd.aspect,
                      near=cd.near, far=cd.far).to_abi(2)
    _compare_gbuffers(gpu_ctx.gbuffer(cam, w, h, prev_cam=prev), o.gbuffer(cam, w, h, prev_cam=prev))
    g = gpu_ctx.gbuffer(cam, w, h)  # static view
    assert np.array_equal(g["tex1"], gpu_ctx.primary_visibility(cam, w, h))
    hit = g["tex1"]["instance_id"] != MISS
    assert (g["tex2"][..., 3][hit] == 0).all()


@pytest.mark.gpu
def test_gbuffer_feeds_pt_with_demodulation(gpu_ctx):
    """The node chain gbuffer -> pt: tex1 and tex3 of mpt_gbuffer drive mpt_render with merian_render_demodulate_albedo
    (render_pt.slang:119-121); multiplying the albedo back gives the plain image."""
    import torch
    from merian_b200 import abi
    from helpers import render_params
    t = tables("cornell")
    upload_scene(gpu_ctx, t)
    cam = t.desc.camera.to_abi(0)
    w = h = 96
    g = gpu_ctx.gbuffer_device(cam, w, h)
    plain = gpu_ctx.render_device(cam, render_params(w, h, 16), tex1=g["tex1"])
    demod = gpu_ctx.render_device(cam, render_params(w, h, 16, flags=abi.RENDER_EMISSION_ON_PRIMARY | abi.RENDER_DEMODULATE_ALBEDO), tex1=g["tex1"], albedo=g["tex3"])
    back = demod[..., :3] * torch.clamp(g["tex3"][..., :3], min=1e-2)
    assert torch.allclose(back, plain[..., :3], rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------------------------------------------- moving instances
def _moving_instances_scene():
    """One quad instanced twice (an instanced mesh group keeps its node transforms: scene.cpp:1023-1105); the geometry of the SECOND
    instance is flagged IsAnimated, the way Scene::update marks the geometries of animated nodes."""
    from merian_b200 import abi
    s = 0.6
    p = np.array([[-s, -s, 0], [s, -s, 0], [s, s, 0.2], [-s, s, 0.2]], np.float32)
    n = np.tile(np.array([[0, -0.316228, 0.948683]], np.float32), (4, 1))
    quad = Mesh(positions=p, normals=n, uvs=np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32), indices=np.array([[0, 1, 2], [0, 2, 3]], np.uint16),
                material_id=0, flags=MESH_IS_OPAQUE)

    def xf(tx, ty, tz, angle):
        c, si = np.cos(angle), np.sin(angle)
        return np.array([[c, -si, 0, tx], [si, c, 0, ty], [0, 0, 1, tz], [0, 0, 0, 1]], np.float32)
    desc = SceneDescription(meshes=[quad], nodes=[Node([0], xf(-0.8, 0.0, 0.0, 0.2)), Node([0], xf(0.8, 0.1, -0.2, -0.3))],
                            materials=[DiffuseMaterial((0.6, 0.5, 0.4, 1.0))], env=synthetic.sky_env(64, 32, sun_peak=10.0, sun_sigma_deg=10.0),
                            camera=CameraDesc(position=(0.0, 0.1, 3.0), target=(0.0, 0.0, 0.0), yfov_deg=55.0, aspect=1.5, near=0.01, far=100.0))
    t = SceneTables(desc)
    assert len(t.instances) == 2 and len(t.geometries) == 2
    assert not (t.geometries[1][2] & abi.GEOM_PRETRANSFORMED)
    mesh_id, mat, flags, prims = t.geometries[1]
    t.geometries[1] = (mesh_id, mat, flags | abi.GEOM_IS_ANIMATED, prims)
    prev = t.instance_transforms.copy()
    prev[0] = xf(-0.5, 0.3, 0.1, 0.0)    # not animated: must be ignored
    prev[1] = xf(0.65, 0.02, -0.25, -0.1)
    return t, prev


def test_oracle_gbuffer_moving_instance(oracle_lib):
    """Scene::to_prev_world_position (scene.slang:261-270): the motion vector and delta_z of an animated instance follow the previous
    frame's instance transform; geometry without IsAnimated ignores it."""
    t, prev_xf = _moving_instances_scene()
    o = make_oracle(t)
    w, h = 150, 100
    cam = t.desc.camera.to_abi(0)
    static = o.gbuffer(cam, w, h)
    o.set_prev_transforms(prev_xf)
    g = o.gbuffer(cam, w, h)
    for k in ("tex0", "tex3", "emission"):
        assert np.array_equal(g[k], static[k])
    assert np.array_equal(g["tex1"], static["tex1"]) and np.array_equal(g["tex2"][..., 2], static["tex2"][..., 2])
    inst = g["tex1"]["instance_index"]
    hit = g["tex1"]["instance_id"] != MISS
    a, b = hit & (inst == 0), hit & (inst == 1)
    assert a.sum() > 500 and b.sum() > 500
    assert np.array_equal(g["tex2"][a], static["tex2"][a]) and np.array_equal(g["tex2"][~hit], static["tex2"][~hit])
    # instance 1: world position -> object space with the current matrix -> world space of the previous frame, in float64
    d = _primary_dirs(cam, w, h)
    z = g["tex2"][..., 2]
    pos = np.array(cam.position[:], np.float64) + d.astype(np.float64) * z[..., None]
    M, Mp = t.instance_transforms[1].astype(np.float64), prev_xf[1].astype(np.float64)
    obj = (np.concatenate([pos[b], np.ones((int(b.sum()), 1))], -1) @ np.linalg.inv(M).T)
    prev_pos = (obj @ Mp.T)[:, :3]
    px, py = np.meshgrid(np.arange(w), np.arange(h))
    mv_ref = _project(cam, prev_pos, w, h) - np.stack([px, py], -1)[b]
    assert np.abs(mv_ref).max() > 3.0                                            # the instance moved by several pixels
    assert np.abs(g["tex2"][..., :2][b] - mv_ref).max() < 5e-3
    dz_ref = np.linalg.norm(prev_pos - np.array(cam.position[:], np.float64), axis=-1) - z[b]
    assert np.abs(g["tex2"][..., 3][b] - dz_ref).max() < 1e-4
    o.set_prev_transforms(None)
    assert np.array_equal(o.gbuffer(cam, w, h)["tex2"], static["tex2"])


@pytest.mark.gpu
def test_cuda_gbuffer_moving_instance_matches_oracle(gpu_ctx):
    t, prev_xf = _moving_instances_scene()
    upload_scene(gpu_ctx, t)
    o = make_oracle(t)
    w, h = 150, 100
    cam = t.desc.camera.to_abi(1)
    cd = t.desc.camera
    prev = CameraDesc(position=tuple(np.array(cd.position) + [0.04, -0.02, 0.03]), target=cd.target, up=cd.up, yfov_deg=cd.yfov_deg, aspect=cd.aspect,
                      near=cd.near, far=cd.far).to_abi(0)
    static = gpu_ctx.gbuffer(cam, w, h, prev_cam=prev)
    gpu_ctx.set_prev_transforms(prev_xf)
    o.set_prev_transforms(prev_xf)
    got, ref = gpu_ctx.gbuffer(cam, w, h, prev_cam=prev), o.gbuffer(cam, w, h, prev_cam=prev)
    _compare_gbuffers(got, ref)
    b = (got["tex1"]["instance_id"] != MISS) & (got["tex1"]["instance_index"] == 1)
    assert np.abs(got["tex2"][..., :2][b] - static["tex2"][..., :2][b]).max() > 3.0
    assert np.array_equal(got["tex2"][~b], static["tex2"][~b])
    with pytest.raises(Exception):
        gpu_ctx.set_prev_transforms(prev_xf[:1])                                  # instance count must match the scene's
    gpu_ctx.set_prev_transforms(None)
    assert np.array_equal(gpu_ctx.gbuffer(cam, w, h, prev_cam=prev)["tex2"], static["tex2"])


def test_prev_vertices_are_refused(oracle_lib):
    """MPT_GEOM_HAS_PREV_VERTICES (morphed / skinned meshes) is outside the built scope: the host-side scene build says so."""
    from merian_b200 import abi
    from merian_b200.host_accel import HostAccel
    from merian_b200._lib import MptError
    t, _ = _moving_instances_scene()
    mesh_id, mat, flags, prims = t.geometries[0]
    t.geometries[0] = (mesh_id, mat, flags | abi.GEOM_HAS_PREV_VERTICES, prims)
    with pytest.raises(MptError):
        HostAccel(t)


# ------------------------------------------------------------------------------------------------- mip chains + ray differentials
def _minified_floor_scene(mipmap):
    """A large floor under a low camera with a high-frequency base-colour texture tiled 24x: towards the horizon a pixel covers many
    texels, which is what GradLODSampler + the mip chain exist for (gbuffer_rt.slang:41-66)."""
    from merian_b200 import abi
    from merian_b200.scene import Texture
    s = 20.0
    p = np.array([[-s, 0, -s], [s, 0, -s], [s, 0, s], [-s, 0, s]], np.float32)
    n = np.tile(np.array([[0, 1, 0]], np.float32), (4, 1))
    uv = np.array([[0, 0], [24, 0], [24, 24], [0, 24]], np.float32)
    floor = Mesh(positions=p, normals=n, uvs=uv, indices=np.array([[0, 2, 1], [0, 3, 2]], np.uint16), material_id=0, flags=MESH_IS_OPAQUE)
    rng = np.random.default_rng(11)
    px = rng.integers(0, 256, (64, 64, 4), dtype=np.uint8)
    px[..., 3] = 255
    tex = Texture(px, srgb=True, filter=abi.FILTER_LINEAR, mipmap=mipmap)
    mats = [GltfMaterial(base_color_factor=(1, 1, 1, 1), metallic_factor=0.0, roughness_factor=1.0, base_color_texture=0)]
    env = synthetic.sky_env(64, 32, sun_peak=10.0, sun_sigma_deg=10.0)
    return SceneDescription(meshes=[floor], nodes=[Node([0])], materials=mats, textures=[tex], env=env,
                            camera=CameraDesc(position=(0.0, 1.0, 12.0), target=(0.0, 0.0, 0.0), yfov_deg=50.0, aspect=1.5, near=0.01, far=200.0))


def test_oracle_gbuffer_selects_coarser_levels_towards_the_horizon(oracle_lib):
    from merian_b200 import abi
    w, h = 150, 100
    res = {}
    for mode in (abi.MIPMAP_NONE, abi.MIPMAP_LINEAR):
        t = SceneTables(_minified_floor_scene(mode))
        o = make_oracle(t)
        g = o.gbuffer(t.desc.camera.to_abi(0), w, h)
        res[mode] = g
    hit = res[abi.MIPMAP_NONE]["tex1"]["instance_id"] != MISS
    far = hit & (res[abi.MIPMAP_NONE]["tex2"][..., 2] > 15.0)   # strongly minified
    assert far.sum() > 200
    a0, a1 = res[abi.MIPMAP_NONE]["tex3"][..., :3], res[abi.MIPMAP_LINEAR]["tex3"][..., :3]
    # LOD 0 aliases (texel noise: large pixel-to-pixel variance); the mip chain converges to the texture's mean
    assert a1[far].std(axis=0).mean() < 0.35 * a0[far].std(axis=0).mean()
    assert np.abs(a1[far].mean(axis=0) - a0[hit].mean(axis=0)).max() < 0.05
    # (magnification selecting level 0 is pinned on the sampler itself: tests/test_texture_mips.py)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["nearest", "linear"])
def test_cuda_gbuffer_with_mip_chain_matches_oracle(gpu_ctx, mode):
    """tex2DGrad on a mip-mapped CUDA array against the oracle's SampleGrad (Vulkan LOD rules in fp32). The hardware quantises the LOD to 8
    fractional bits, the bilinear weights to 8, and may approximate the gradient length, so texels agree to a few percent; a kernel that
    ignored the gradients (round 1: LOD 0 only) would be far outside (asserted against the chain-less render below)."""
    from merian_b200 import abi
    m = abi.MIPMAP_NEAREST if mode == "nearest" else abi.MIPMAP_LINEAR
    t = SceneTables(_minified_floor_scene(m))
    upload_scene(gpu_ctx, t)
    o = make_oracle(t)
    cam = t.desc.camera.to_abi(0)
    w, h = 150, 100
    got, ref = gpu_ctx.gbuffer(cam, w, h), o.gbuffer(cam, w, h)
    mism, bary_err, _ = compare_hits(got["tex1"], ref["tex1"])
    assert mism == 0.0 and bary_err == 0.0
    hit = ref["tex1"]["instance_id"] != MISS
    d = np.abs(got["tex3"][..., :3] - ref["tex3"][..., :3])[hit]
    print(f"mip {mode}: albedo |diff| mean {d.mean():.4f} p99 {np.quantile(d, 0.99):.4f} max {d.max():.4f}")
    if m == abi.MIPMAP_LINEAR:
        assert d.mean() < 0.01 and np.quantile(d, 0.99) < 0.06
    else:
        assert d.mean() < 0.03  # nearest level: a rounding difference in lambda switches the level for a pixel
    # and the chain is really used: the LOD-0 image of the same scene is far away from both
    t0 = SceneTables(_minified_floor_scene(abi.MIPMAP_NONE))
    upload_scene(gpu_ctx, t0)
    lod0 = gpu_ctx.gbuffer(cam, w, h)
    far = hit & (ref["tex2"][..., 2] > 15.0)
    assert np.abs(lod0["tex3"][..., :3] - ref["tex3"][..., :3])[far].mean() > 5 * np.abs(got["tex3"][..., :3] - ref["tex3"][..., :3])[far].mean()
    # pt reads LOD 0 whatever the texture has (LOD0Sampler): the irradiance of the chained and the chain-less scene is the same image
    from helpers import render_params
    P = render_params(w, h, 4)
    upload_scene(gpu_ctx, t)
    a = gpu_ctx.render(cam, P)
    upload_scene(gpu_ctx, t0)
    b = gpu_ctx.render(cam, P)
    assert np.array_equal(a, b)


@pytest.mark.gpu
def test_texture_and_env_upload_from_device_memory(gpu_ctx):
    """mpt_texture_upload_device / mpt_env_set_device (the Vulkan -> CUDA import leg: texels that already live in device memory, e.g. an
    imported VkBuffer): same renders as the host uploads, with and without a row pitch."""
    import torch
    from merian_b200 import abi
    from helpers import render_params
    t = SceneTables(_minified_floor_scene(abi.MIPMAP_LINEAR))
    upload_scene(gpu_ctx, t)
    cam = t.desc.camera.to_abi(0)
    w, h = 120, 80
    want_g = gpu_ctx.gbuffer(cam, w, h)
    want = gpu_ctx.render(cam, render_params(w, h, 4))
    tex = t.desc.textures[0]
    px = np.ascontiguousarray(tex.pixels)
    H, W = px.shape[:2]
    padded = np.zeros((H, W + 16, 4), np.uint8)  # row pitch = (W + 16) * 4 bytes
    padded[:, :W] = px
    d_px = torch.from_numpy(padded).cuda()
    gpu_ctx.texture_upload_device(0, W, H, abi.TEX_RGBA8_SRGB, tex.wrap_u, tex.wrap_v, tex.filter, d_px, mipmap=abi.MIPMAP_LINEAR, row_pitch=(W + 16) * 4)
    env = t.desc.env
    d_env = torch.from_numpy(np.ascontiguousarray(env.pixels, np.float32)).cuda()
    gpu_ctx.env_set_device(abi.ENV_LATLONG, np.ascontiguousarray(env.to_local, np.float32), env.intensity, env.pixels.shape[1], env.pixels.shape[0], d_env)
    got_g = gpu_ctx.gbuffer(cam, w, h)
    assert np.array_equal(got_g["tex3"], want_g["tex3"]) and np.array_equal(got_g["emission"], want_g["emission"])
    assert np.array_equal(gpu_ctx.render(cam, render_params(w, h, 4)), want)
    # device-to-device path (no chain), tightly packed
    gpu_ctx.texture_upload_device(0, W, H, abi.TEX_RGBA8_SRGB, tex.wrap_u, tex.wrap_v, tex.filter, torch.from_numpy(px).cuda())
    gpu_ctx.texture_upload(1, W, H, abi.TEX_RGBA8_SRGB, tex.wrap_u, tex.wrap_v, tex.filter, px)
    assert np.array_equal(gpu_ctx.render(cam, render_params(w, h, 4)), want)  # pt reads level 0 either way
