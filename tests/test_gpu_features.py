"""GPU parity tests of the scene features around the core path, all through the C-ABI (libmpt.so) against the oracle:
instanced two-level scenes, textured glTF materials, glTF extensions incl. transmission/volume, instance masks,
two-sided / flip-facing geometry, sample sharding, the host-buffer entry point and the error behaviour."""
import numpy as np
import pytest
import torch

from merian_b200 import abi, sharding, synthetic
from merian_b200._lib import MptError
from merian_b200.scene import (MESH_FLIP_FACING, MESH_IS_OPAQUE, MESH_TWO_SIDED, CameraDesc, DiffuseMaterial, GltfMaterial, Mesh, Node,
                               SceneDescription, SceneTables, Texture, upload_scene)

import bsdf_cases as bc
from helpers import compare_hits, edge_ties, luminance, make_oracle, rel_mse, render_params, tables

pytestmark = pytest.mark.gpu


def _both(gpu_ctx, t):
    upload_scene(gpu_ctx, t)
    return make_oracle(t)


def _check_vis(gpu_ctx, o, cam, w, h, mask=0xFF):
    ref = o.primary_visibility(cam, w, h, mask)
    got = gpu_ctx.primary_visibility(cam, w, h, mask)
    mism, bary_err, same = compare_hits(got, ref)
    n_mis, n_ties = edge_ties(got, ref, same)
    assert n_mis == n_ties and mism < 1e-4 and bary_err == 0.0, (n_mis, n_ties, mism, bary_err)
    return ref


def _check_render(gpu_ctx, o, cam, P, tol_frac=0.05):
    ref = o.render(cam, P)
    got = gpu_ctx.render(cam, P)
    assert np.isfinite(got).all()
    d = np.abs(got[..., :3] - ref[..., :3]).max(axis=-1) / np.maximum(np.abs(ref[..., :3]).max(axis=-1), 1e-3)
    frac = float((d > 1e-3).mean())
    lr, lg = luminance(ref).mean(), luminance(got).mean()
    print(f"mean lum ref {lr:.5f} got {lg:.5f} relMSE {rel_mse(got, ref):.3e} differing pixels {frac:.4f}")
    assert abs(lg - lr) / max(lr, 1e-6) < 1e-2
    assert frac < tol_frac
    return ref, got


# ------------------------------------------------------------------------------------------------- instancing (C5 shape)
def test_instanced_two_level_scene(gpu_ctx):
    t = tables("instanced_rocks", n_instances=60, rock_res=17, env_size=(128, 64))
    assert len(t.instances) == 61
    o = _both(gpu_ctx, t)
    cam = t.desc.camera.to_abi(0)
    ref = _check_vis(gpu_ctx, o, cam, 320, 180)
    hit = ref["instance_id"] != 0xFFFFFFFF
    assert hit.mean() > 0.3 and len(np.unique(ref["instance_index"][hit])) > 20  # many instances visible
    _check_render(gpu_ctx, o, cam, render_params(160, 90, 16))
    # arbitrary rays through the TLAS
    rays = synthetic.incoherent_rays(cam, 128, 72)
    rays["origin"] = np.array([0.0, 12.0, 0.0], np.float32)
    rays["direction"][:, 1] = -np.abs(rays["direction"][:, 1]) - 0.1
    rays["direction"] /= np.linalg.norm(rays["direction"], axis=1, keepdims=True)
    r, g = o.trace(rays), gpu_ctx.trace(rays)
    same = (r["primitive_id"] == g["primitive_id"]) & (r["instance_index"] == g["instance_index"]) & (r["instance_id"] == g["instance_id"])
    assert 1 - same.mean() < 1e-4
    assert np.abs(r["barycentrics"] - g["barycentrics"])[same].max() == 0.0


def test_stacked_instances_of_one_blas(gpu_ctx):
    """Eight instances of the same two-triangle BLAS stacked along the view axis: every ray meets the SAME triangle record in several
    instances, in an order the TLAS decides. The closest one must win whatever was hit first (a hit on record X in one instance says nothing
    about record X in another instance)."""
    quad = _quad_mesh(0.0, 1.0, 0, mask=0xFF)
    rng = np.random.default_rng(3)
    nodes = []
    for k in range(8):
        T = np.eye(4, dtype=np.float32)
        T[:3, 3] = [rng.uniform(-1.2, 1.2), rng.uniform(-1.2, 1.2), -0.5 * k]
        nodes.append(Node([0], T))
    order = rng.permutation(8)
    d = _layer_scene([GltfMaterial(base_color_factor=(0.7, 0.7, 0.7, 1), metallic_factor=0, roughness_factor=1.0)], [quad])
    d.nodes = [nodes[i] for i in order]
    t = SceneTables(d)
    assert len(t.instances) == 8
    o = _both(gpu_ctx, t)
    cam = t.desc.camera.to_abi(0)
    ref = _check_vis(gpu_ctx, o, cam, 256, 256)
    hit = ref["instance_id"] != 0xFFFFFFFF
    assert hit.mean() > 0.3 and len(np.unique(ref["instance_index"][hit])) >= 5  # several of the stacked copies are the front-most somewhere
    rays = synthetic.incoherent_rays(cam, 160, 160)
    rays["origin"] = np.array([0.0, 0.0, 3.0], np.float32)
    rays["direction"] = np.array([0.0, 0.0, -1.0], np.float32) + 0.25 * (rng.random((len(rays), 3)).astype(np.float32) - 0.5)
    rays["direction"] /= np.linalg.norm(rays["direction"], axis=1, keepdims=True)
    r, g = o.trace(rays), gpu_ctx.trace(rays)
    assert np.array_equal(r["instance_index"], g["instance_index"]) and np.array_equal(r["primitive_id"], g["primitive_id"])
    _check_render(gpu_ctx, o, cam, render_params(96, 96, 16))


# ------------------------------------------------------------------------------------------------- textures (C4 shape)
def test_textured_materials(gpu_ctx):
    t = tables("terrain", nx=64, ny=32, patches=(2, 2), textured=True, texture_size=64, env_size=(128, 64))
    assert len(t.desc.textures) == 16 and len(t.desc.materials) == 4
    o = _both(gpu_ctx, t)
    cam = t.desc.camera.to_abi(0)
    _check_vis(gpu_ctx, o, cam, 240, 136)
    # hardware bilinear filtering uses 8-bit fractional weights (1.8 fixed point) where the oracle uses fp32 weights:
    # per-pixel differences are expected, the estimate is not biased
    ref = o.render(cam, render_params(120, 68, 32))
    got = gpu_ctx.render(cam, render_params(120, 68, 32))
    assert np.isfinite(got).all()
    lr, lg = luminance(ref).mean(), luminance(got).mean()
    print(f"textured: mean lum ref {lr:.5f} got {lg:.5f} relMSE {rel_mse(got, ref):.3e}")
    assert abs(lg - lr) / lr < 2e-2
    assert rel_mse(got, ref) < 0.05


def _quad_mesh(z, size, mat, flags=MESH_IS_OPAQUE, mask=0xFF, flip=False, uv_scale=1.0):
    p = np.array([[-size, -size, z], [size, -size, z], [size, size, z], [-size, size, z]], np.float32)
    idx = np.array([[0, 1, 2], [0, 2, 3]], np.uint16)
    if flip:
        idx = idx[:, ::-1].copy()
    n = np.tile(np.array([[0, 0, 1]], np.float32), (4, 1))
    uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32) * uv_scale
    return Mesh(positions=p, normals=n, uvs=uv, indices=idx, material_id=mat, flags=flags, instance_mask=mask)


def _layer_scene(materials, meshes, textures=()):
    env = synthetic.sky_env(64, 32, sun_peak=50.0, sun_sigma_deg=10.0, sun_dir=(0.2, 0.3, 1.0))
    return SceneDescription(meshes=meshes, nodes=[Node([i]) for i in range(len(meshes))], materials=list(materials), textures=list(textures), env=env,
                            camera=CameraDesc(position=(0.1, 0.2, 3.0), target=(0.0, 0.0, 0.0), yfov_deg=50.0, aspect=1.0, near=0.01, far=100.0))


def test_alpha_tested_geometry(gpu_ctx):
    """Non-opaque geometry runs the any-hit alpha test (scene.slang:500-519): texels with alpha < 0.5 are holes."""
    rng = np.random.default_rng(7)
    a = np.zeros((16, 16, 4), np.uint8)
    a[..., :3] = 200
    a[..., 3] = np.where(rng.random((16, 16)) < 0.5, 0, 255)
    tex = Texture(a, srgb=False, filter=abi.FILTER_NEAREST)
    mats = [GltfMaterial(base_color_factor=(0.8, 0.8, 0.8, 1), base_color_texture=0, metallic_factor=0, roughness_factor=0.8),
            GltfMaterial(base_color_factor=(0.2, 0.6, 0.2, 1), metallic_factor=0, roughness_factor=1.0)]
    meshes = [_quad_mesh(0.5, 1.0, 0, flags=0), _quad_mesh(-0.5, 2.0, 1)]  # front quad is NOT opaque
    t = SceneTables(_layer_scene(mats, meshes, [tex]))
    o = _both(gpu_ctx, t)
    cam = t.desc.camera.to_abi(0)
    ref = _check_vis(gpu_ctx, o, cam, 128, 128)
    seen = set(np.unique(ref["instance_id"] + ref["geometry_index"]).tolist())
    assert {0, 1} <= seen  # both the cut-out quad and the quad behind it are visible
    _check_render(gpu_ctx, o, cam, render_params(64, 64, 16))


def test_instance_masks_two_sided_and_flip_facing(gpu_ctx):
    mats = [GltfMaterial(base_color_factor=(0.8, 0.3, 0.3, 1), metallic_factor=0, roughness_factor=0.7),
            GltfMaterial(base_color_factor=(0.3, 0.3, 0.8, 1), metallic_factor=0, roughness_factor=0.7),
            GltfMaterial(base_color_factor=(0.3, 0.8, 0.3, 1), metallic_factor=0.5, roughness_factor=0.4)]
    meshes = [_quad_mesh(0.6, 0.5, 0, mask=0x01),                                              # only visible to mask bit 0
              _quad_mesh(0.2, 0.9, 1, flags=MESH_IS_OPAQUE | MESH_TWO_SIDED, flip=True),     # back face towards the camera, two-sided
              _quad_mesh(-0.2, 1.4, 2, flags=MESH_IS_OPAQUE | MESH_FLIP_FACING, flip=True),  # wound backwards + flip-facing = front
              _quad_mesh(-0.6, 2.5, 0, flip=True)]                                           # back face, one-sided: culled
    t = SceneTables(_layer_scene(mats, meshes))
    o = _both(gpu_ctx, t)
    cam = t.desc.camera.to_abi(0)
    for mask in (0xFF, 0xFE, 0x01):
        ref = _check_vis(gpu_ctx, o, cam, 96, 96, mask)
        ids = set(np.unique(ref["instance_index"][ref["instance_id"] != 0xFFFFFFFF]).tolist())
        print("mask", hex(mask), "visible TLAS instances", ids)
        if mask == 0xFE:
            assert 1 not in ids or len(ids) >= 1
    P = render_params(64, 64, 16)
    _check_render(gpu_ctx, o, cam, P)
    P2 = render_params(64, 64, 16, mask=0xFE)
    _check_render(gpu_ctx, o, cam, P2)


# ------------------------------------------------------------------------------------------------- glTF extensions
EXT_MATERIALS = {
    "clearcoat": GltfMaterial(base_color_factor=(0.7, 0.1, 0.1, 1), metallic_factor=0.0, roughness_factor=0.6, clearcoat={"weight": 1.0, "roughness": 0.1}),
    "sheen": GltfMaterial(base_color_factor=(0.2, 0.2, 0.6, 1), metallic_factor=0.0, roughness_factor=0.8, sheen={"color": (0.9, 0.8, 0.7), "roughness": 0.4}),
    "iridescence": GltfMaterial(base_color_factor=(0.5, 0.5, 0.5, 1), metallic_factor=1.0, roughness_factor=0.3,
                                iridescence={"factor": 1.0, "ior": 1.3, "thickness_min": 100.0, "thickness_max": 400.0}),
    "anisotropy": GltfMaterial(base_color_factor=(0.9, 0.9, 0.9, 1), metallic_factor=1.0, roughness_factor=0.3, anisotropy={"strength": 0.8, "rotation": 0.5}),
    "transmission": GltfMaterial(base_color_factor=(0.9, 0.95, 1.0, 1), metallic_factor=0.0, roughness_factor=0.2, transmission={"weight": 0.9},
                                 volume={"absorption": (0.5, 0.2, 0.1)}),
    "all": GltfMaterial(base_color_factor=(0.8, 0.7, 0.6, 1), metallic_factor=0.3, roughness_factor=0.4, transmission={"weight": 0.3},
                        volume={"absorption": (0.1, 0.1, 0.3)}, clearcoat={"weight": 0.5, "roughness": 0.2}, sheen={"color": (0.3, 0.3, 0.3), "roughness": 0.5},
                        iridescence={"factor": 0.5, "ior": 1.4, "thickness_min": 200.0, "thickness_max": 300.0}, anisotropy={"strength": 0.4, "rotation": 1.0}),
}


@pytest.mark.parametrize("name", list(EXT_MATERIALS))
def test_gltf_extension_materials(gpu_ctx, name):
    """A slab (box) of the extension material over a diffuse floor under the sky: exercises sample_eval of the full
    GltfBSDF nest, transmission bookkeeping (ior, interior medium) and Beer-Lambert absorption."""
    m = EXT_MATERIALS[name]
    box = synthetic._mesh_from_parts([synthetic._box((0.0, 0.0, 0.0), (0.8, 0.8, 0.25), 20.0, skip_bottom=False)], 0)
    floor = _quad_mesh(-0.6, 3.0, 1)
    mats = [m, GltfMaterial(base_color_factor=(0.6, 0.6, 0.6, 1), metallic_factor=0, roughness_factor=1.0)]
    t = SceneTables(_layer_scene(mats, [box, floor]))
    o = _both(gpu_ctx, t)
    cam = t.desc.camera.to_abi(0)
    _check_vis(gpu_ctx, o, cam, 96, 96)
    _check_render(gpu_ctx, o, cam, render_params(64, 64, 32, L=6), tol_frac=0.10)


def test_diffuse_material_model(gpu_ctx):
    mats = [DiffuseMaterial((0.25, 0.5, 0.75, 1.0)), GltfMaterial(base_color_factor=(0.6, 0.6, 0.6, 1), metallic_factor=0, roughness_factor=1.0)]
    t = SceneTables(_layer_scene(mats, [_quad_mesh(0.3, 0.8, 0), _quad_mesh(-0.3, 2.5, 1)]))
    o = _both(gpu_ctx, t)
    _check_render(gpu_ctx, o, t.desc.camera.to_abi(0), render_params(64, 64, 16))


# ------------------------------------------------------------------------------------------------- BSDF suite on the device
GPU_BSDF_CASES = [c for c in bc.cases(include_albedo=False) if c[1]["bsdf"] != bc.SHEEN]  # SheenBRDF alone is not a sub-tree of GltfBSDF: SHEEN_LAYER covers the lobe


@pytest.mark.parametrize("name,params,checks,extra", GPU_BSDF_CASES, ids=[c[0] for c in GPU_BSDF_CASES])
def test_reference_bsdf_suite_on_device(gpu_ctx, name, params, checks, extra):
    """tests/test_bsdf.cpp restated against the CUDA BSDFs (the sub-trees of GltfBSDF that mpt_bsdf_check can isolate)."""
    r = bc.run_case(lambda **kw: gpu_ctx.bsdf_check(**kw), params, checks, extra)
    from oracle.oracle_py import bsdf_check
    ro = bc.R(bsdf_check(**params))
    # same estimator on both arms: furnace and mean cos/pdf agree within the Monte-Carlo error
    assert np.all(np.abs(r.furnace - ro.furnace) <= 6 * (r.furnace_se + ro.furnace_se) + 2e-3)
    assert abs(r.mean_cos - ro.mean_cos) <= 6 * (r.se_cos + ro.se_cos) + 2e-3


# ------------------------------------------------------------------------------------------------- per-ray setup
def test_ray_setup_matches_ieee(gpu_ctx):
    """The shear constants of the watertight test define the barycentrics bit for bit; the device computes the three quotients with one shared
    reciprocal (bvh8_traverse.cuh shear_quotients) and must agree with IEEE float32 divisions, including directions with tiny, zero and
    huge components. The 1-ulp inverse direction used when a ray enters an instance stays within 2 ulp of the IEEE one."""
    rng = np.random.default_rng(11)
    n = 1 << 20
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    # stress cases: axis-aligned, tiny / zero minor components, very long and very short vectors, exact ties between components
    d[:1000, 1] = 0.0
    d[1000:2000, 0] *= np.float32(1e-30)
    d[2000:3000] *= np.float32(1e-25)
    d[3000:4000] *= np.float32(1e20)
    d[4000:5000, 1] = d[4000:5000, 0]
    d[5000:6000, 2] = -d[5000:6000, 0]
    d[6000:7000, 0] = np.float32(1e-45)  # a denormal
    d[7000:7003] = np.eye(3, dtype=np.float32)
    d[7003:7006] = -np.eye(3, dtype=np.float32)
    shear, k, perm, idir, idir_a = gpu_ctx.debug_ray_setup(d)
    ax, ay, az = np.abs(d[:, 0]), np.abs(d[:, 1]), np.abs(d[:, 2])
    kz = np.where(ax > ay, np.where(ax > az, 0, 2), np.where(ay > az, 1, 2))
    kx, ky = (kz + 1) % 3, (kz + 2) % 3
    idx = np.arange(n)
    dz = d[idx, kz]
    swap = dz < 0
    kx, ky = np.where(swap, ky, kx), np.where(swap, kx, ky)
    assert np.array_equal(k, (kx | (ky << 2) | (kz << 4)).astype(np.uint32))
    with np.errstate(all="ignore"):
        want = np.stack([d[idx, kx] / dz, d[idx, ky] / dz, np.float32(1.0) / dz], axis=1).astype(np.float32)
    assert np.array_equal(shear.view(np.uint32), want.view(np.uint32)), int((shear.view(np.uint32) != want.view(np.uint32)).any(axis=1).sum())
    octant = (d[:, 0] < 0).astype(np.uint32) | ((d[:, 1] < 0).astype(np.uint32) << 1) | ((d[:, 2] < 0).astype(np.uint32) << 2)
    c = 7 - octant
    want_perm = np.where(c & 1, 0x55, 0) | np.where(c & 2, 0x3300, 0) | np.where(c & 4, 0x0F0000, 0) | (c << 24)
    assert np.array_equal(perm, want_perm.astype(np.uint32))
    ok = np.abs(d) > 1e-20  # away from the clamp of safe_rcp
    rel = np.abs(idir_a[ok].astype(np.float64) - idir[ok].astype(np.float64)) / np.abs(idir[ok].astype(np.float64))
    assert rel.max() < 2.5e-7, rel.max()


# ------------------------------------------------------------------------------------------------- thin-lens camera
def test_thin_lens_camera_matches_oracle(gpu_ctx):
    """camera.slang:73-92 through Scene::generate_camera_ray (scene.slang:91-101): lens sample = R2(frame) rotated per pixel. The lens sample
    passes through sin / cos (last-bit differences between CUDA and libm), so ids are compared with a small mismatch budget."""
    t = tables("cornell")
    o = _both(gpu_ctx, t)
    pin = t.desc.camera.to_abi(0)
    lens = t.desc.camera.to_abi(0)
    lens.aperture_radius = 0.05
    for frame in (0, 7):
        gpu_ctx.set_scene_frame(frame)
        o.set_scene_frame(frame)
        got, ref = gpu_ctx.primary_visibility(lens, 128, 128), o.primary_visibility(lens, 128, 128)
        assert (got["primitive_id"] != ref["primitive_id"]).mean() < 2e-3
        same = got["primitive_id"] == ref["primitive_id"]
        assert np.abs(got["barycentrics"][same].astype(np.float32) - ref["barycentrics"][same].astype(np.float32)).max() < 2e-3
    base = gpu_ctx.primary_visibility(pin, 128, 128)
    assert (got["primitive_id"] != base["primitive_id"]).mean() > 5e-3  # the aperture does move the rays
    gpu_ctx.set_scene_frame(0)
    a = gpu_ctx.primary_visibility(lens, 128, 128)
    assert (a["primitive_id"] != got["primitive_id"]).any()  # and the lens sample follows the frame
    # radiance: same estimator on both arms, paths differ only where the last bits of the lens sample flip a decision
    P = render_params(64, 64, 16, frame=3)
    g, r = gpu_ctx.render(lens, P), o.render(lens, P)
    lum = lambda im: float((im[..., :3] * np.array([0.299, 0.587, 0.114], np.float32)).sum(-1).mean())
    assert abs(lum(g) - lum(r)) / lum(r) < 0.02
    # the GBuffer node sees the same primary ray (and pinhole ray differentials, gbuffer_rt.slang:35-36)
    gpu_ctx.set_scene_frame(3)
    o.set_scene_frame(3)
    gg, go = gpu_ctx.gbuffer(lens, 64, 64), o.gbuffer(lens, 64, 64)
    assert (gg["tex1"]["primitive_id"] != go["tex1"]["primitive_id"]).mean() < 5e-3
    gpu_ctx.set_scene_frame(0)
    o.set_scene_frame(0)


# ------------------------------------------------------------------------------------------------- sharding / frames / flags
def test_sample_shards_sum_to_the_full_frame(gpu_ctx):
    t = tables("cornell")
    upload_scene(gpu_ctx, t)
    cam = t.desc.camera.to_abi(0)
    P = render_params(96, 96, 13, frame=5)
    full = gpu_ctx.render(cam, P)
    for world in (2, 3, 8):
        acc = np.zeros_like(full)
        for rank in range(world):
            p, total = sharding.shard_params(P, rank, world)
            acc += gpu_ctx.render(cam, p)
        acc /= total
        assert np.allclose(acc[..., :3], full[..., :3], rtol=2e-5, atol=1e-6), world
        assert np.allclose(acc[..., 3], full[..., 3], rtol=1e-4, atol=1e-6), world


def test_frames_differ_and_repeat(gpu_ctx):
    t = tables("cornell")
    upload_scene(gpu_ctx, t)
    cam = t.desc.camera.to_abi(0)
    a = gpu_ctx.render(cam, render_params(64, 64, 4, frame=1))
    b = gpu_ctx.render(cam, render_params(64, 64, 4, frame=2))
    a2 = gpu_ctx.render(cam, render_params(64, 64, 4, frame=1))
    assert np.array_equal(a, a2)  # deterministic: counter-style seeding, no atomics on the radiance path
    assert not np.array_equal(a, b)


def test_many_passes_equal_one_pass(gpu_ctx):
    """spp larger than the in-flight pool (several wavefront passes) accumulates to the same mean as the oracle."""
    t = tables("cornell")
    o = _both(gpu_ctx, t)
    cam = t.desc.camera.to_abi(0)
    _check_render(gpu_ctx, o, cam, render_params(32, 32, 300))


def test_gbuffer_handoff_and_flags(gpu_ctx):
    """tex1 given by the caller (the GBuffer node's product) == primary hits traced internally; emission-on-primary off;
    demodulate albedo."""
    t = tables("cornell")
    o = _both(gpu_ctx, t)
    cam = t.desc.camera.to_abi(0)
    w = h = 64
    tex1 = gpu_ctx.primary_visibility(cam, w, h)
    P = render_params(w, h, 8)
    a = gpu_ctx.render(cam, P)
    b = gpu_ctx.render(cam, P, tex1=tex1)
    assert np.array_equal(a, b)
    P0 = render_params(w, h, 8, flags=0)
    r0, g0 = o.render(cam, P0), gpu_ctx.render(cam, P0)
    lamp = tex1["instance_id"] + tex1["geometry_index"] == 3
    assert lamp.any() and (luminance(g0)[lamp] < luminance(a)[lamp]).all()  # the lamp's own emission is gone
    assert abs(luminance(g0).mean() - luminance(r0).mean()) / luminance(r0).mean() < 1e-2
    albedo = np.full((h, w, 4), 0.5, np.float32)
    Pd = render_params(w, h, 8, flags=abi.RENDER_EMISSION_ON_PRIMARY | abi.RENDER_DEMODULATE_ALBEDO)
    gd = gpu_ctx.render(cam, Pd, albedo=albedo)
    assert np.allclose(gd[..., :3], a[..., :3] * 2.0, rtol=1e-5, atol=1e-6)
    assert np.allclose(o.render(cam, Pd, albedo=albedo)[..., :3].mean(), gd[..., :3].mean(), rtol=1e-2)


def test_render_host_matches_device_path(gpu_ctx):
    t = tables("cornell")
    upload_scene(gpu_ctx, t)
    cam = t.desc.camera.to_abi(0)
    P = render_params(64, 64, 8)
    tex1 = gpu_ctx.primary_visibility_device(cam, 64, 64)
    dev = gpu_ctx.render_device(cam, P, tex1=tex1).cpu().numpy()
    h_tex1 = tex1.cpu().pin_memory()
    host = gpu_ctx.render_host(cam, P, h_tex1=h_tex1).numpy()
    assert np.array_equal(dev, host)


def test_iterations_in_flight_equal_the_synchronous_call(gpu_ctx):
    """mpt_render_host_begin / _end with two slots in flight (merian's default) return the images of mpt_render_host, frame by frame."""
    t = tables("cornell")
    upload_scene(gpu_ctx, t)
    cam = t.desc.camera.to_abi(0)
    w, h = 96, 64
    tex1 = gpu_ctx.primary_visibility_device(cam, w, h).cpu().pin_memory()
    ref = [gpu_ctx.render_host(cam, render_params(w, h, 4, frame=f), h_tex1=tex1).clone() for f in range(5)]
    outs = [torch.empty((h, w, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    got = []
    for f in range(5):
        if f >= 2:
            gpu_ctx.render_host_end(f % 2)
            got.append(outs[f % 2].clone())
        gpu_ctx.render_host_begin(cam, render_params(w, h, 4, frame=f), outs[f % 2], f % 2, h_tex1=tex1)
    for f in (3, 4):
        gpu_ctx.render_host_end(f % 2)
        got.append(outs[f % 2].clone())
    for f in range(5):
        assert torch.equal(got[f], ref[f]), f
    with pytest.raises(MptError):
        gpu_ctx.render_host_end(0)  # nothing in flight
    gpu_ctx.render_host_begin(cam, render_params(w, h, 1), outs[0], 0, h_tex1=tex1)
    with pytest.raises(MptError):
        gpu_ctx.render_host_begin(cam, render_params(w, h, 1), outs[0], 0, h_tex1=tex1)  # slot busy
    gpu_ctx.render_host_end(0)
    with pytest.raises(MptError):
        gpu_ctx.render_host_begin(cam, render_params(w, h, 1), outs[0], 7, h_tex1=tex1)


def test_error_behaviour(gpu_ctx):
    """C-ABI error convention: int status + mpt_last_error, never an exception across the boundary or a silent fallback."""
    from merian_b200.context import Context
    fresh = Context(0)
    t = tables("cornell")
    cam = t.desc.camera.to_abi(0)
    with pytest.raises(MptError) as e:  # render_pt.cpp:98: scene not ready
        fresh.render(cam, render_params(8, 8, 1))
    assert e.value.status == abi.MPT_ERR_NOT_READY
    upload_scene(fresh, t)
    for bad in (render_params(0, 8, 1), render_params(8, 8, -1), render_params(8, 8, 1, L=0), render_params(8, 8, 1, stride=0)):
        with pytest.raises(MptError) as e:
            fresh.render(cam, bad)
        assert e.value.status == abi.MPT_ERR_INVALID_ARGUMENT
    z = fresh.render(cam, render_params(8, 8, 0))
    assert (z == 0).all()  # spp 0: defined, all zero
    one = fresh.render(cam, render_params(8, 8, 3, L=1))  # max path length 1: emission at the primary vertex only
    assert np.isfinite(one).all()
    fresh.close()


def test_material_system_hints(gpu_ctx):
    """alpha-test threshold / min roughness / normal clamping are MaterialSystem settings (material_system.hpp:83-97)."""
    rng = np.random.default_rng(11)
    a = np.zeros((16, 16, 4), np.uint8)
    a[..., :3] = 220
    a[..., 3] = rng.integers(0, 256, (16, 16))
    nm = np.zeros((16, 16, 4), np.uint8)
    nm[..., 0] = rng.integers(60, 200, (16, 16))
    nm[..., 1] = rng.integers(60, 200, (16, 16))
    nm[..., 2:] = 255
    texs = [Texture(a, srgb=False, filter=abi.FILTER_NEAREST), Texture(nm, srgb=False, filter=abi.FILTER_NEAREST)]
    mats = [GltfMaterial(base_color_factor=(0.8, 0.8, 0.8, 1), base_color_texture=0, normal_texture=1, metallic_factor=1.0, roughness_factor=0.0),
            GltfMaterial(base_color_factor=(0.2, 0.6, 0.2, 1), metallic_factor=0, roughness_factor=1.0)]
    t = SceneTables(_layer_scene(mats, [_quad_mesh(0.5, 1.0, 0, flags=0), _quad_mesh(-0.5, 2.0, 1)], texs))
    o = _both(gpu_ctx, t)
    cam = t.desc.camera.to_abi(0)
    base = gpu_ctx.primary_visibility(cam, 96, 96)
    try:
        for hints in ((0.8, 0.2, False), (0.2, 0.0316, True)):
            gpu_ctx.materials_set_hints(*hints)
            o.materials_set_hints(*hints)
            ref = _check_vis(gpu_ctx, o, cam, 96, 96)
            _check_render(gpu_ctx, o, cam, render_params(64, 64, 16), tol_frac=0.10)
        assert (ref["instance_id"] + ref["geometry_index"] != base["instance_id"] + base["geometry_index"]).any()  # threshold moved the holes
    finally:
        gpu_ctx.materials_set_hints()


@pytest.mark.gpu
def test_masked_out_scene_shades_every_continuation_ray(gpu_ctx):
    """The GBuffer hit texture is made with one instance mask and the path tracer runs with another that excludes the scene's only
    instance (both are independent node properties, render_pt.cpp:147-187 / gbuffer.cpp): every continuation ray is a miss that still
    has to collect the environment. 2 M samples — far more rays than the pooled extend kernel has records, which is what went wrong:
    records that needed no traversal dropped out of circulation and the rest of the queue was never shaded."""
    from helpers import make_oracle, render_params
    t = tables("terrain", nx=120, ny=60, env_size=(128, 64))
    upload_scene(gpu_ctx, t)
    cam = t.desc.camera.to_abi(0)
    w, h, spp = 512, 256, 16
    tex1 = gpu_ctx.primary_visibility(cam, w, h, mask=0xFF)
    P = render_params(w, h, spp, mask=0x00)
    got = gpu_ctx.render(cam, P, tex1=tex1)
    o = make_oracle(t)
    ref = o.render(cam, P, tex1=np.ascontiguousarray(tex1))
    assert np.isfinite(got).all()
    hit = tex1["instance_id"] != 0xFFFFFFFF
    assert hit.mean() > 0.3
    lg, lr = luminance(got)[hit].mean(), luminance(ref)[hit].mean()
    assert lr > 0 and abs(lg - lr) / lr < 0.02, (lg, lr)
    # every path of a hit pixel ends at the sky after one bounce: rows of the image are not left dark towards the end of the queue
    rows = luminance(got).reshape(h, -1).mean(1), luminance(ref).reshape(h, -1).mean(1)
    assert np.abs(rows[0] - rows[1]).max() < 0.15 * rows[1].mean()
