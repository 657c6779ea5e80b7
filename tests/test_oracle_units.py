"""Unit tests of the CPU oracle's building blocks (CPU only).

Pinned by the reference's own tests where they exist:
  * material entry byte layout + Lambert eval          tests/test_material_system.cpp:107-215
  * 1x1 RGBA8 nearest texture sample                   tests/test_texture_manager.cpp:105-158
and by independent numpy restatements of the Slang sources for everything else (RNG, encodings, camera,
ray offset, Vulkan culling semantics of the triangle test).
"""
import ctypes as C
import struct

import numpy as np
import pytest

from merian_b200 import abi, synthetic
from merian_b200.scene import (CameraDesc, DiffuseMaterial, GltfMaterial, Mesh, Node, SceneDescription, SceneTables, Texture,
                               build_material_buffer, device_decode_normal, device_encode_normal, upload_scene)


def _f3(*v):
    return (C.c_float * 3)(*v)


# ------------------------------------------------------------------------------------------------- RNG (hash.slang, random.slang)
def test_pcg4d16_matches_numpy(oracle_lib):
    rng = np.random.default_rng(0)
    v = rng.integers(0, 2 ** 32, size=(256, 4), dtype=np.uint64).astype(np.uint32)
    v[0] = 0
    v[1] = 0xFFFFFFFF
    ref = synthetic.pcg4d16_np(v[:, 0], v[:, 1], v[:, 2], v[:, 3])
    got = np.array([oracle_lib.oracle_pcg4d16(*[int(x) for x in row]) for row in v], np.uint32)
    assert (ref == got).all()


def test_pcg4d16_known_answer():
    # hand-evaluated: pcg4d16(0,0,0,0): v = 1013904223 each; hash.slang:138-146
    c = 1013904223
    M = 1 << 32
    x = y = z = w = c
    x = (x + y * w) % M
    y = (y + z * x) % M
    z = (z + x * y) % M
    w = (w + y * z) % M
    x = (x + y * w) % M
    assert int(synthetic.pcg4d16_np(np.uint32(0), np.uint32(0), np.uint32(0), np.uint32(0))) == x


def test_pcg_stream_matches_numpy(oracle_lib):
    n = 64
    u = np.zeros(n, np.uint32)
    f = np.zeros(n, np.float32)
    oracle_lib.oracle_pcg_stream(C.c_uint32(12345), C.c_uint32(n), C.c_void_p(u.ctypes.data), C.c_void_p(f.ctypes.data))
    st = np.uint32(12345)
    for i in range(n):
        fl, st = synthetic.pcg_next_float_np(st)
        assert f[i] == fl
        assert 0.0 <= f[i] < 1.0
    # uint_to_zero_one_float_rust, bit-twiddling.slang:23-27
    assert (f == (u >> 8).astype(np.float32) * np.float32(2.0 ** -24)).all()


# ------------------------------------------------------------------------------------------------- encodings (encoding.slang)
def test_normal_encoding_matches_numpy_and_round_trips(oracle_lib):
    rng = np.random.default_rng(1)
    n = rng.normal(size=(500, 3)).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    n[:6] = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], np.float32)
    ref_enc = device_encode_normal(n)
    for i in range(len(n)):
        e = oracle_lib.oracle_encode_normal(_f3(*n[i]))
        assert e == int(ref_enc[i])
        out = _f3()
        oracle_lib.oracle_decode_normal(C.c_uint32(e), out)
        d = np.array(out[:], np.float32)
        assert np.abs(d - device_decode_normal(np.array([e], np.uint32))[0]).max() < 1e-6
        assert np.dot(d, n[i]) > 1.0 - 1e-4  # 16-bit octahedral: < 0.01 rad


def test_half_conversion_matches_ieee(oracle_lib):
    vals = np.concatenate([np.linspace(-70000, 70000, 401), np.array([0.0, -0.0, 1e-8, 6.1e-5, 5.96e-8, 65504.0, 1.0 / 3.0])]).astype(np.float32)
    for v in vals:
        h = oracle_lib.oracle_float_to_half(C.c_float(v))
        assert h == int(np.float32(v).astype(np.float16).view(np.uint16)), v
    for h in list(range(0, 0x7C00, 37)) + list(range(0x8000, 0xFC00, 41)):
        assert oracle_lib.oracle_half_to_float(C.c_uint16(h)) == float(np.array([h], np.uint16).view(np.float16)[0])


# ------------------------------------------------------------------------------------------------- hit.slang:14-23
def test_offset_ray_origin(oracle_lib):
    def ref(p, n):
        p, n = np.float32(p), np.float32(n)
        of = (n * np.float32(256.0)).astype(np.int32)
        pi = p.view(np.int32) + np.where(p < 0, -of, of)
        return np.where(np.abs(p) < np.float32(1 / 16), p + n * np.float32(1 / 65536), pi.view(np.float32))
    rng = np.random.default_rng(2)
    for _ in range(200):
        p = (rng.normal(size=3) * rng.choice([0.01, 1.0, 100.0])).astype(np.float32)
        n = rng.normal(size=3).astype(np.float32)
        n /= np.linalg.norm(n)
        out = _f3()
        oracle_lib.oracle_offset_ray_origin(_f3(*p), _f3(*n), out)
        assert (np.array(out[:], np.float32) == ref(p, n)).all()
        # the offset point lies on the normal's side
        assert np.dot(np.array(out[:], np.float64) - p.astype(np.float64), n) > 0


# ------------------------------------------------------------------------------------------------- camera.slang:29-56
def test_camera_ray_matches_numpy(oracle_lib):
    cam = CameraDesc(position=(1, 2, 3), target=(0, 0.5, -1), yfov_deg=50.0, aspect=16 / 9, near=0.1, far=1000.0).to_abi(0)
    w, h = 64, 36
    f32 = np.float32
    U, V, W = (np.array(cam.U[:], f32), np.array(cam.V[:], f32), np.array(cam.W[:], f32))
    # the basis is orthogonal and scaled as camera.cpp:349-368 says
    assert abs(np.dot(U, V)) < 1e-5 and abs(np.dot(U, W)) < 1e-5 and abs(np.dot(V, W)) < 1e-5
    assert np.isclose(np.linalg.norm(U) / np.linalg.norm(V), 16 / 9, rtol=1e-5)
    for (x, y) in [(0, 0), (63, 35), (31, 17), (5, 30)]:
        r = abi.Ray()
        oracle_lib.oracle_camera_ray(C.byref(cam), C.c_uint32(x), C.c_uint32(y), C.c_uint32(w), C.c_uint32(h), C.byref(r))
        ndc = (f32(2) * ((f32(x) + f32(0.5)) / f32(w)) - f32(1), f32(-2) * ((f32(y) + f32(0.5)) / f32(h)) + f32(1))
        d = ndc[0] * U + ndc[1] * V + W
        d = d / np.linalg.norm(d)
        assert np.abs(np.array(r.direction[:]) - d).max() < 1e-6
        inv_cos = 1.0 / np.dot(W / np.linalg.norm(W), d)
        assert np.isclose(r.tmin, 0.1 * inv_cos, rtol=1e-5) and np.isclose(r.tmax, 1000.0 * inv_cos, rtol=1e-5)
        assert tuple(r.origin[:]) == (1.0, 2.0, 3.0)


# ------------------------------------------------------------------------------------------------- camera.slang:73-92, scene.slang:91-101
def test_thin_lens_ray_matches_numpy(oracle_lib):
    """Thin lens: origin on the aperture disk spanned by normalize(U), normalize(V); the ray passes through the pixel's point on the plane
    of perfect focus (position + ndc.x U + ndc.y V + W); lens sample = frac(R2(frame) + PCG(xxhash32(pixel)).next_float2())."""
    cam = CameraDesc(position=(1, 2, 3), target=(0, 0.5, -1), yfov_deg=50.0, aspect=16 / 9, near=0.1, far=1000.0).to_abi(0)
    cam.aperture_radius = 0.25
    w, h = 64, 36
    U, V, W, P = (np.array(v[:], np.float64) for v in (cam.U, cam.V, cam.W, cam.position))

    def xxhash32(px, py):
        M = 0xFFFFFFFF
        P2, P3, P4, P5 = 2246822519, 3266489917, 668265263, 374761393
        h32 = (py + P5 + px * P3) & M
        h32 = (P4 * (((h32 << 17) | (h32 >> 15)) & M)) & M
        h32 = (P2 * (h32 ^ (h32 >> 15))) & M
        h32 = (P3 * (h32 ^ (h32 >> 13))) & M
        return h32 ^ (h32 >> 16)

    def pcg_floats(seed, n):
        u = np.zeros(n, np.uint32)
        f = np.zeros(n, np.float32)
        oracle_lib.oracle_pcg_stream(C.c_uint32(seed), C.c_uint32(n), C.c_void_p(u.ctypes.data), C.c_void_p(f.ctypes.data))
        return f

    seen = set()
    for frame in (0, 1, 5):
        for (x, y) in [(0, 0), (63, 35), (31, 17)]:
            r = abi.Ray()
            oracle_lib.oracle_camera_ray_frame(C.byref(cam), C.c_uint32(x), C.c_uint32(y), C.c_uint32(w), C.c_uint32(h), C.c_uint32(frame), C.byref(r))
            o, d = np.array(r.origin[:], np.float64), np.array(r.direction[:], np.float64)
            xi = pcg_floats(xxhash32(x, y), 2).astype(np.float64)
            r2 = np.array([(0.5 + np.float32(0.7548776662466927) * frame) % 1.0, (0.5 + np.float32(0.5698402909980532) * frame) % 1.0])
            rnd = (r2 + xi) % 1.0
            disk = np.array([np.cos(2 * np.pi * rnd[0]), np.sin(2 * np.pi * rnd[0])]) * np.sqrt(rnd[1])
            off = 0.25 * (disk[0] * U / np.linalg.norm(U) + disk[1] * V / np.linalg.norm(V))
            assert np.abs(o - (P + off)).max() < 1e-5
            ndc = (2 * ((x + 0.5) / w) - 1, -2 * ((y + 0.5) / h) + 1)
            focus = P + ndc[0] * U + ndc[1] * V + W
            want = (focus - o) / np.linalg.norm(focus - o)
            assert np.abs(d - want).max() < 1e-5
            seen.add((frame, x, y, round(float(o[0]), 6)))
    assert len({s[3] for s in seen}) == len(seen)  # a different lens point for every (pixel, frame)
    # aperture 0 = the pinhole ray, whatever the frame
    cam.aperture_radius = 0.0
    a, b = abi.Ray(), abi.Ray()
    oracle_lib.oracle_camera_ray_frame(C.byref(cam), C.c_uint32(3), C.c_uint32(4), C.c_uint32(w), C.c_uint32(h), C.c_uint32(9), C.byref(a))
    oracle_lib.oracle_camera_ray(C.byref(cam), C.c_uint32(3), C.c_uint32(4), C.c_uint32(w), C.c_uint32(h), C.byref(b))
    assert bytes(a) == bytes(b)


def test_r2_jitter_sequence():
    cam = CameraDesc(position=(0, 0, 1), target=(0, 0, 0), jitter_sequence="R2")
    assert cam.jitter(0) == (0.0, 0.0)  # frame 0 => no jitter (camera.cpp:325-330)
    js = [cam.jitter(i) for i in range(32)]
    assert js[:16] == js[16:]  # period 16 (gltf_scene.cpp:773-774)
    assert all(-0.5 <= a < 0.5 and -0.5 <= b < 0.5 for a, b in js)
    assert len(set(js[:16])) == 16


# ------------------------------------------------------------------------------------------------- triangle test: Vulkan facing rules
def _isect(lib, o, d, v0, v1, v2, cull, tmin=0.0, tmax=1e7):
    tuv = _f3()
    hit = lib.oracle_intersect_triangle(_f3(*o), _f3(*d), _f3(*v0), _f3(*v1), _f3(*v2), C.c_int(cull), C.c_float(tmin), C.c_float(tmax), tuv)
    return bool(hit), tuple(tuv[:])


def test_triangle_facing_and_culling(oracle_lib):
    # counter-clockwise seen from +z: geometric normal cross(e01, e02) = +z
    v0, v1, v2 = (0, 0, 0), (1, 0, 0), (0, 1, 0)
    front = ((0.25, 0.25, 1.0), (0, 0, -1))  # ray arriving against the normal: front face
    back = ((0.25, 0.25, -1.0), (0, 0, 1))
    # cull 0 = RAY_FLAG_CULL_BACK_FACING_TRIANGLES (acceleration-structure.slang:64)
    hit, (t, u, v) = _isect(oracle_lib, *front, v0, v1, v2, 0)
    assert hit and t == 1.0 and abs(u - 0.25) < 1e-6 and abs(v - 0.25) < 1e-6  # u,v = weights of v1, v2
    assert not _isect(oracle_lib, *back, v0, v1, v2, 0)[0]
    # cull 1 = flip-facing instance (VK_GEOMETRY_INSTANCE_TRIANGLE_FLIP_FACING_BIT): the back face is the front
    assert not _isect(oracle_lib, *front, v0, v1, v2, 1)[0]
    assert _isect(oracle_lib, *back, v0, v1, v2, 1)[0]
    # cull 2 = TRIANGLE_FACING_CULL_DISABLE (two-sided groups)
    assert _isect(oracle_lib, *front, v0, v1, v2, 2)[0] and _isect(oracle_lib, *back, v0, v1, v2, 2)[0]
    # tmin/tmax are exclusive bounds
    assert not _isect(oracle_lib, *front, v0, v1, v2, 0, tmin=1.0)[0]
    assert not _isect(oracle_lib, *front, v0, v1, v2, 0, tmax=1.0)[0]
    # outside
    assert not _isect(oracle_lib, (2, 2, 1), (0, 0, -1), v0, v1, v2, 2)[0]


def test_triangle_watertight_shared_edge(oracle_lib):
    """Rays through the shared diagonal of a quad hit exactly one... at least one of the two triangles (no cracks)."""
    a, b, c, d = (0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0)
    rng = np.random.default_rng(3)
    for _ in range(2000):
        s = float(rng.random())
        p = np.array([s, s, 0.0])  # on the diagonal a-c
        o = np.array([rng.normal(), rng.normal(), 1.0 + rng.random() * 3])
        dr = p - o
        dr /= np.linalg.norm(dr)
        h1 = _isect(oracle_lib, o, dr, a, b, c, 2)[0]
        h2 = _isect(oracle_lib, o, dr, a, c, d, 2)[0]
        assert h1 or h2


# ------------------------------------------------------------------------------------------------- material system pins
def _one_tri_scene(materials, textures=()):
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (3, 1))
    m = Mesh(positions=pos, normals=nrm, uvs=np.zeros((3, 2), np.float32), material_id=0)
    return SceneDescription(meshes=[m], nodes=[Node(meshes=[0])], materials=list(materials), textures=list(textures),
                            camera=CameraDesc(position=(0.3, 0.3, 2), target=(0.3, 0.3, 0)))


def test_material_entry_layout_ReadMaterialOnGPU(oracle_lib):
    """tests/test_material_system.cpp:107-163: header = (type id u16, alpha texture id u16), payload = base colour floats."""
    from oracle.oracle_py import OracleContext
    t = SceneTables(_one_tri_scene([DiffuseMaterial((0.25, 0.5, 0.75, 1.0))]))
    assert t.material_entry_size == 4 + 16
    assert struct.unpack_from("<2H4f", t.material_entries, 0) == (0, 0xFFFF, 0.25, 0.5, 0.75, 1.0)
    o = OracleContext()
    upload_scene(o, t)
    out = np.zeros(7, np.uint32)
    oracle_lib.oracle_material_read(o.h, C.c_uint32(0), C.c_void_p(out.ctypes.data))
    assert out[0] == 0 and out[1] == 0xFFFF and out[6] == 1
    assert np.allclose(out[2:6].view(np.float32), [0.25, 0.5, 0.75, 1.0], atol=1e-5)


def test_lambert_eval_SampleMaterialOnGPU(oracle_lib):
    """tests/test_material_system.cpp:165-215: DiffuseMaterial(1,0,0,1).eval((0,0,1),(0,0,1)) = albedo / pi, +-1e-5."""
    from oracle.oracle_py import OracleContext
    o = OracleContext()
    upload_scene(o, SceneTables(_one_tri_scene([DiffuseMaterial((1.0, 0.0, 0.0, 1.0))])))
    out = _f3()
    oracle_lib.oracle_material_eval(o.h, C.c_uint32(0), _f3(0, 0, 1), _f3(0, 0, 1), (C.c_float * 2)(0, 0), out)
    assert np.allclose(out[:], [1.0 / np.pi, 0.0, 0.0], atol=1e-5)


def test_material_buffer_mixed_types_and_padding():
    """material_system.cpp:124-195: type ids in first-registration order; entry size = header + max payload rounded to uints."""
    mats = [GltfMaterial(), DiffuseMaterial(), GltfMaterial(anisotropy={"strength": 0.5}), GltfMaterial(clearcoat={"weight": 1.0})]
    buf, es, types = build_material_buffer(mats)
    assert types == [(abi.MATERIAL_GLTF, 0), (abi.MATERIAL_DIFFUSE, 0), (abi.MATERIAL_GLTF, abi.GLTF_ANISOTROPY), (abi.MATERIAL_GLTF, abi.GLTF_CLEARCOAT)]
    assert len(GltfMaterial().payload()) == 56  # gltf_material.hpp:102-112
    assert len(mats[2].payload()) == 66 and es == 4 + 68 and len(buf) == 4 * es
    assert [struct.unpack_from("<H", buf, i * es)[0] for i in range(4)] == [0, 1, 2, 3]


def test_texture_1x1_nearest_SampleTextureOnGPU(oracle_lib):
    """tests/test_texture_manager.cpp:105-158: 1x1 RGBA8 UNORM red, nearest/repeat -> (1,0,0,1) +-1e-3."""
    from oracle.oracle_py import OracleContext
    px = np.array([[[255, 0, 0, 255]]], np.uint8)
    tex = Texture(px, srgb=False, filter=abi.FILTER_NEAREST)
    o = OracleContext()
    upload_scene(o, SceneTables(_one_tri_scene([DiffuseMaterial()], [tex])))
    out = (C.c_float * 4)()
    for uv in [(0.5, 0.5), (0.0, 0.0), (3.7, -2.2)]:
        oracle_lib.oracle_sample_texture(o.h, C.c_uint32(0), C.c_float(uv[0]), C.c_float(uv[1]), out)
        assert np.allclose(out[:], [1, 0, 0, 1], atol=1e-3)


def test_texture_bilinear_wrap_and_srgb(oracle_lib):
    from oracle.oracle_py import OracleContext
    px = np.zeros((2, 2, 4), np.uint8)
    px[0, 0] = (255, 0, 0, 255)
    px[0, 1] = (0, 255, 0, 255)
    px[1, 0] = (0, 0, 255, 255)
    px[1, 1] = (255, 255, 255, 255)
    grey = np.full((1, 1, 4), 128, np.uint8)
    o = OracleContext()
    upload_scene(o, SceneTables(_one_tri_scene([DiffuseMaterial()], [Texture(px, srgb=False), Texture(px, srgb=False, wrap_u=abi.WRAP_CLAMP_TO_EDGE, wrap_v=abi.WRAP_CLAMP_TO_EDGE),
                                                                     Texture(grey, srgb=True)])))
    out = (C.c_float * 4)()

    def s(tid, u, v):
        oracle_lib.oracle_sample_texture(o.h, C.c_uint32(tid), C.c_float(u), C.c_float(v), out)
        return np.array(out[:])
    assert np.allclose(s(0, 0.25, 0.25), [1, 0, 0, 1])          # texel centre
    assert np.allclose(s(0, 0.5, 0.25), [0.5, 0.5, 0, 1])       # halfway between two texels
    assert np.allclose(s(0, 0.5, 0.5), [0.5, 0.5, 0.5, 1])      # all four
    assert np.allclose(s(0, 0.0, 0.25), [0.5, 0.5, 0, 1])       # repeat: wraps to the texel on the other side
    assert np.allclose(s(1, 0.0, 0.25), [1, 0, 0, 1])           # clamp to edge
    lin = ((128 / 255 + 0.055) / 1.055) ** 2.4
    assert np.allclose(s(2, 0.5, 0.5)[:3], lin, atol=1e-6) and np.isclose(s(2, 0.5, 0.5)[3], 128 / 255)  # alpha stays linear


def test_env_latlong_eval(oracle_lib):
    """environment-map.slang:23-35 + math.slang:243-249: uv = (atan2(x,-z)/2pi + 0.5, acos(y)/pi)."""
    from oracle.oracle_py import OracleContext
    d = synthetic.terrain(nx=4, ny=4, env_size=(64, 32))
    px = np.zeros((32, 64, 4), np.float32)
    px[..., 0] = np.arange(64, dtype=np.float32)[None, :]
    px[..., 1] = np.arange(32, dtype=np.float32)[:, None]
    d.env.pixels = px
    d.env.intensity = 2.0
    o = OracleContext()
    upload_scene(o, SceneTables(d))
    out = _f3()
    for dir_ in [(0, 0, -1), (1, 0, 0), (0.3, 0.5, 0.2), (-0.7, -0.1, 0.4)]:
        v = np.array(dir_, np.float64)
        v /= np.linalg.norm(v)
        oracle_lib.oracle_env_eval(o.h, _f3(*v), out)
        u = np.arctan2(v[0], -v[2]) / (2 * np.pi) + 0.5
        w = np.arccos(v[1]) / np.pi
        assert np.isclose(out[0], 2.0 * ((u * 64 - 0.5) % 64), atol=2e-3) or u * 64 < 0.5 or u * 64 > 63.5
        assert np.isclose(out[1], 2.0 * (w * 32 - 0.5), atol=2e-3)
