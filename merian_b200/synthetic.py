"""Seeded procedural scenes for BASELINE.json's configs (SURVEY.md §8(d) "Synthetic inputs").

One generator per config; every scene records its seed and GENERATOR_VERSION so results are reproducible.
  C1 cornell()                 ~32 triangles, Lambert-like glTF materials, emissive ceiling quad
  C2/C3 terrain()              displaced grid (default 1000 x 500 x 2 = 1M triangles) + lat-long env
  C4 terrain(patches=(8,8), textured=True)   2M triangles, 64 textured glTF materials
  C5 instanced_rocks()         one ~10k-triangle BLAS x 1000 TLAS instances
"""
from __future__ import annotations

import numpy as np

from . import abi
from .scene import (MESH_IS_OPAQUE, CameraDesc, EnvMap, GltfMaterial, Mesh, Node, SceneDescription, Texture)

GENERATOR_VERSION = 1


# ----------------------------------------------------------------------------------------------- helpers
def _value_noise(shape, cells, rng):
    """Smooth value noise on a (H, W) grid from a (cells+1)^2 random lattice."""
    H, W = shape
    lat = rng.random((cells + 1, cells + 1), dtype=np.float32)
    y = np.linspace(0, cells, H, dtype=np.float32)
    x = np.linspace(0, cells, W, dtype=np.float32)
    y0 = np.minimum(y.astype(np.int32), cells - 1)
    x0 = np.minimum(x.astype(np.int32), cells - 1)
    fy = y - y0
    fx = x - x0
    fy = fy * fy * (3 - 2 * fy)
    fx = fx * fx * (3 - 2 * fx)
    a = lat[y0][:, x0]
    b = lat[y0][:, x0 + 1]
    c = lat[y0 + 1][:, x0]
    d = lat[y0 + 1][:, x0 + 1]
    top = a + (b - a) * fx[None, :]
    bot = c + (d - c) * fx[None, :]
    return (top + (bot - top) * fy[:, None]).astype(np.float32)


def fbm(shape, seed, octaves=6, base_cells=4):
    rng = np.random.default_rng(seed)
    out = np.zeros(shape, np.float32)
    amp, cells = 1.0, base_cells
    for _ in range(octaves):
        out += amp * (_value_noise(shape, cells, rng) - 0.5)
        amp *= 0.5
        cells *= 2
    return out


def _quad(p0, p1, p2, p3):
    """Two CCW triangles; front face normal = cross(p1-p0, p3-p0)."""
    P = np.array([p0, p1, p2, p3], np.float32)
    n = np.cross(P[1] - P[0], P[3] - P[0])
    n = n / np.linalg.norm(n)
    pos = np.array([P[0], P[1], P[2], P[0], P[2], P[3]], np.float32)
    uv = np.array([[0, 0], [1, 0], [1, 1], [0, 0], [1, 1], [0, 1]], np.float32)
    return pos, np.tile(n.astype(np.float32), (6, 1)), uv


def _box(center, half, yaw_deg, skip_bottom=True):
    """Outward-facing box faces (5 faces = 10 triangles when the bottom is skipped)."""
    c, s = np.cos(np.deg2rad(yaw_deg)), np.sin(np.deg2rad(yaw_deg))
    R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float32)
    hx, hy, hz = half
    corners = {}
    for ix in (-1, 1):
        for iy in (-1, 1):
            for iz in (-1, 1):
                corners[(ix, iy, iz)] = (R @ np.array([ix * hx, iy * hy, iz * hz], np.float32)) + np.array(center, np.float32)
    k = corners
    faces = [
        (k[(1, -1, 1)], k[(1, -1, -1)], k[(1, 1, -1)], k[(1, 1, 1)]),      # +x
        (k[(-1, -1, -1)], k[(-1, -1, 1)], k[(-1, 1, 1)], k[(-1, 1, -1)]),  # -x
        (k[(-1, 1, 1)], k[(1, 1, 1)], k[(1, 1, -1)], k[(-1, 1, -1)]),      # +y
        (k[(-1, -1, 1)], k[(1, -1, 1)], k[(1, 1, 1)], k[(-1, 1, 1)]),      # +z
        (k[(1, -1, -1)], k[(-1, -1, -1)], k[(-1, 1, -1)], k[(1, 1, -1)]),  # -z
    ]
    if not skip_bottom:
        faces.append((k[(-1, -1, -1)], k[(1, -1, -1)], k[(1, -1, 1)], k[(-1, -1, 1)]))
    parts = [_quad(*f) for f in faces]
    return tuple(np.concatenate([p[i] for p in parts]) for i in range(3))


def _mesh_from_parts(parts, material_id):
    pos = np.concatenate([p[0] for p in parts])
    nrm = np.concatenate([p[1] for p in parts])
    uv = np.concatenate([p[2] for p in parts])
    return Mesh(positions=pos, normals=nrm, uvs=uv, indices=None, material_id=material_id, flags=MESH_IS_OPAQUE)


# ----------------------------------------------------------------------------------------------- C1
def cornell(seed: int = 0) -> SceneDescription:
    """C1: unit Cornell box, 5 walls + 2 boxes + ceiling light = 32 triangles, camera on +z looking down -z."""
    white = GltfMaterial(base_color_factor=(0.73, 0.73, 0.73, 1), metallic_factor=0.0, roughness_factor=1.0)
    red = GltfMaterial(base_color_factor=(0.65, 0.05, 0.05, 1), metallic_factor=0.0, roughness_factor=1.0)
    green = GltfMaterial(base_color_factor=(0.12, 0.45, 0.15, 1), metallic_factor=0.0, roughness_factor=1.0)
    light = GltfMaterial(base_color_factor=(0.78, 0.78, 0.78, 1), emissive_factor=(17, 12, 4), metallic_factor=0.0,
                         roughness_factor=1.0)
    # room [0,1]^3, open towards +z; all faces point inwards
    floor = _quad((0, 0, 1), (1, 0, 1), (1, 0, 0), (0, 0, 0))
    ceil = _quad((0, 1, 0), (1, 1, 0), (1, 1, 1), (0, 1, 1))
    back = _quad((0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0))
    left = _quad((0, 0, 1), (0, 0, 0), (0, 1, 0), (0, 1, 1))    # red, normal +x
    right = _quad((1, 0, 0), (1, 0, 1), (1, 1, 1), (1, 1, 0))   # green, normal -x
    lamp = _quad((0.35, 0.999, 0.35), (0.65, 0.999, 0.35), (0.65, 0.999, 0.65), (0.35, 0.999, 0.65))  # normal -y
    short = _box((0.67, 0.15, 0.63), (0.15, 0.15, 0.15), -18)
    tall = _box((0.33, 0.30, 0.33), (0.15, 0.30, 0.15), 18)
    meshes = [
        _mesh_from_parts([floor, ceil, back], 0),
        _mesh_from_parts([left], 1),
        _mesh_from_parts([right], 2),
        _mesh_from_parts([lamp], 3),
        _mesh_from_parts([short], 0),
        _mesh_from_parts([tall], 0),
    ]
    nodes = [Node(meshes=[i]) for i in range(len(meshes))]
    # slightly off-axis so that the quads' shared diagonals do not run exactly through pixel centres
    cam = CameraDesc(position=(0.5, 0.5, 2.4), target=(0.493, 0.508, 0.0), yfov_deg=39.3, aspect=1.0, near=0.01, far=100.0)
    return SceneDescription(meshes=meshes, nodes=nodes, materials=[white, red, green, light], camera=cam, name="cornell",
                            seed=seed)


# ----------------------------------------------------------------------------------------------- env
def sky_env(width: int = 4096, height: int = 2048, sun_peak: float = 5e4, sun_sigma_deg: float = 1.0,
            sun_dir=(0.35, 0.6, -0.72), sky: float = 0.5) -> EnvMap:
    """C3 env: constant sky + one Gaussian sun, lat-long RGBA32F (uv convention of math.slang:243-249)."""
    sd = np.array(sun_dir, np.float64)
    sd /= np.linalg.norm(sd)
    u = (np.arange(width, dtype=np.float64) + 0.5) / width
    v = (np.arange(height, dtype=np.float64) + 0.5) / height
    phi = np.pi * (2.0 * u - 1.0)
    theta = np.pi * v
    st, ct = np.sin(theta)[:, None], np.cos(theta)[:, None]
    # latlong_map_to_world, math.slang:254-259
    dx, dy, dz = st * np.sin(phi)[None, :], np.broadcast_to(ct, (height, width)), -st * np.cos(phi)[None, :]
    cosang = np.clip(dx * sd[0] + dy * sd[1] + dz * sd[2], -1, 1)
    ang = np.arccos(cosang)
    sig = np.deg2rad(sun_sigma_deg)
    sun = sun_peak * np.exp(-0.5 * (ang / sig) ** 2)
    px = np.empty((height, width, 4), np.float32)
    px[..., 0] = sky + sun
    px[..., 1] = sky + 0.95 * sun
    px[..., 2] = sky * 1.3 + 0.85 * sun
    px[..., 3] = 1.0
    return EnvMap(pixels=px, intensity=1.0)


# ----------------------------------------------------------------------------------------------- C2/C3/C4
def _noise_texture(size, seed, channels="rgb"):
    rng = np.random.default_rng(seed)
    base = np.stack([_value_noise((size, size), 16, rng) for _ in range(3)], axis=-1)
    fine = np.stack([_value_noise((size, size), 128, rng) for _ in range(3)], axis=-1)
    t = np.clip(0.7 * base + 0.3 * fine, 0, 1)
    px = np.empty((size, size, 4), np.uint8)
    px[..., :3] = (t * 255).astype(np.uint8)
    px[..., 3] = 255
    return px


class _TileTextures:
    """C4's 4 x 64 textures at full size without minutes of noise synthesis or 17 GB of host memory: four seeded base
    tiles per kind, and texture k = base tile (k mod 4) rolled by a seeded offset (tiling noise stays tiling) — made when
    the texture is uploaded, not when the scene is described."""

    def __init__(self, size, seed):
        self.size, self.seed, self._base = size, seed, {}

    def base(self, kind, i):
        if (kind, i) not in self._base:
            px = _noise_texture(self.size, self.seed * 7919 + 16 * kind + i)
            if kind == 2:  # normal map: small tilt around +z
                px[..., 0] = 128 + (px[..., 0].astype(np.int32) - 128) // 4
                px[..., 1] = 128 + (px[..., 1].astype(np.int32) - 128) // 4
                px[..., 2] = 255
            elif kind == 3:  # emissive: sparse bright spots
                px[..., :3] = np.where(px[..., :3] > 200, px[..., :3], 0)
            self._base[(kind, i)] = px
        return self._base[(kind, i)]

    def make(self, kind, mat_id):
        def gen():
            r = np.random.default_rng(self.seed * 104729 + 4 * mat_id + kind)
            dy, dx = (int(v) for v in r.integers(0, self.size, 2))
            return np.ascontiguousarray(np.roll(self.base(kind, mat_id % 4), (dy, dx), axis=(0, 1)))
        return gen


def terrain(nx: int = 1000, ny: int = 500, size: float = 100.0, height: float = 12.0, seed: int = 1, patches=(1, 1),
            textured: bool = False, texture_size: int = 4096, env: bool = True, env_size=(4096, 2048), width: int = 1920,
            height_px: int = 1080, tiled_textures: bool = False) -> SceneDescription:
    """Displaced grid of nx*ny*2 triangles over a size x size square (y up), fBm height.
    patches=(px,py) splits it into px*py meshes, one glTF material each (C4: 8x8 = 64 textured materials)."""
    H = fbm((ny + 1, nx + 1), seed, octaves=7) * height
    xs = np.linspace(-size / 2, size / 2, nx + 1, dtype=np.float32)
    zs = np.linspace(-size / 2, size / 2, ny + 1, dtype=np.float32)
    X, Z = np.meshgrid(xs, zs)
    P = np.stack([X, H, Z], axis=-1).astype(np.float32)
    dHdx = np.gradient(H, xs, axis=1)
    dHdz = np.gradient(H, zs, axis=0)
    N = np.stack([-dHdx, np.ones_like(H), -dHdz], axis=-1)
    N = (N / np.linalg.norm(N, axis=-1, keepdims=True)).astype(np.float32)
    tile = 8.0 if textured else 1.0
    UV = np.stack([(X / size + 0.5) * tile, (Z / size + 0.5) * tile], axis=-1).astype(np.float32)

    px_n, py_n = patches
    rng = np.random.default_rng(seed + 1000)
    tiles = _TileTextures(texture_size, seed) if (textured and tiled_textures) else None
    meshes, materials, textures = [], [], []
    for pj in range(py_n):
        for pi in range(px_n):
            x0, x1 = pi * nx // px_n, (pi + 1) * nx // px_n
            y0, y1 = pj * ny // py_n, (pj + 1) * ny // py_n
            w, h = x1 - x0, y1 - y0
            sub = (slice(y0, y1 + 1), slice(x0, x1 + 1))
            pos = P[sub].reshape(-1, 3)
            nrm = N[sub].reshape(-1, 3)
            uv = UV[sub].reshape(-1, 2)
            jj, ii = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
            v00 = (jj * (w + 1) + ii).reshape(-1)
            v10, v01, v11 = v00 + 1, v00 + (w + 1), v00 + (w + 2)
            # CCW seen from +y: (v00, v01, v11), (v00, v11, v10)
            idx = np.stack([np.stack([v00, v01, v11], 1), np.stack([v00, v11, v10], 1)], 1).reshape(-1, 3).astype(np.uint32)
            mat_id = len(materials)
            if textured and tiles is not None:
                tbase = len(textures)
                for kind, srgb in ((0, True), (1, False), (2, False), (3, True)):
                    textures.append(Texture(tiles.make(kind, mat_id), srgb=srgb))
                materials.append(GltfMaterial(base_color_factor=(1, 1, 1, 1), emissive_factor=(0.2, 0.2, 0.2),
                                              roughness_factor=float(rng.uniform(0.05, 1.0)),
                                              metallic_factor=float(rng.integers(0, 2)), base_color_texture=tbase,
                                              metallic_roughness_texture=tbase + 1, normal_texture=tbase + 2,
                                              emissive_texture=tbase + 3))
            elif textured:
                tbase = len(textures)
                textures.append(Texture(_noise_texture(texture_size, seed * 7919 + 4 * mat_id + 0), srgb=True))
                textures.append(Texture(_noise_texture(texture_size, seed * 7919 + 4 * mat_id + 1), srgb=False))
                nm = _noise_texture(texture_size, seed * 7919 + 4 * mat_id + 2)
                nm[..., 0] = 128 + (nm[..., 0].astype(np.int32) - 128) // 4
                nm[..., 1] = 128 + (nm[..., 1].astype(np.int32) - 128) // 4
                nm[..., 2] = 255
                textures.append(Texture(nm, srgb=False))
                em = _noise_texture(texture_size, seed * 7919 + 4 * mat_id + 3)
                em[..., :3] = np.where(em[..., :3] > 200, em[..., :3], 0)
                textures.append(Texture(em, srgb=True))
                materials.append(GltfMaterial(base_color_factor=(1, 1, 1, 1), emissive_factor=(0.2, 0.2, 0.2),
                                              roughness_factor=float(rng.uniform(0.05, 1.0)),
                                              metallic_factor=float(rng.integers(0, 2)), base_color_texture=tbase,
                                              metallic_roughness_texture=tbase + 1, normal_texture=tbase + 2,
                                              emissive_texture=tbase + 3))
            else:
                materials.append(GltfMaterial(base_color_factor=(0.5, 0.5, 0.5, 1), metallic_factor=0.0, roughness_factor=0.5))
            meshes.append(Mesh(positions=pos, normals=nrm, uvs=uv, indices=idx, material_id=mat_id, flags=MESH_IS_OPAQUE))
    nodes = [Node(meshes=[i]) for i in range(len(meshes))]
    cam = CameraDesc(position=(0.0, height * 1.6, size * 0.55), target=(0.0, -height * 0.2, 0.0), yfov_deg=45.0,
                     aspect=width / height_px, near=0.05, far=1e4)
    e = sky_env(*env_size) if env else None
    return SceneDescription(meshes=meshes, nodes=nodes, materials=materials, textures=textures, env=e, camera=cam,
                            name=f"terrain_{nx}x{ny}", seed=seed)


# ----------------------------------------------------------------------------------------------- C5
def instanced_rocks(n_instances: int = 1000, rock_res: int = 71, seed: int = 5, env_size=(1024, 512), aspect=16 / 9) -> SceneDescription:
    """C5: one displaced sphere ("rock", 2*(res-1)*res ~ 10k triangles) instanced n times with seeded TRS."""
    rng = np.random.default_rng(seed)
    nu, nv = rock_res, rock_res
    u = np.linspace(0, 2 * np.pi, nu + 1)[:-1]
    v = np.linspace(0, np.pi, nv)
    U, V = np.meshgrid(u, v)
    disp = 1.0 + 0.25 * fbm((nv, nu), seed, octaves=4)
    disp[0, :] = disp[0, :].mean()
    disp[-1, :] = disp[-1, :].mean()
    P = np.stack([np.sin(V) * np.cos(U), np.cos(V), np.sin(V) * np.sin(U)], -1) * disp[..., None]
    P = P.astype(np.float32).reshape(-1, 3)
    N = (P / np.linalg.norm(P, axis=1, keepdims=True)).astype(np.float32)
    UV = np.stack([U / (2 * np.pi), V / np.pi], -1).reshape(-1, 2).astype(np.float32)
    idx = []
    for j in range(nv - 1):
        for i in range(nu):
            a, b = j * nu + i, j * nu + (i + 1) % nu
            c, d = a + nu, b + nu
            idx.append((a, b, d))
            idx.append((a, d, c))
    idx = np.array(idx, np.uint32)
    # orient so that cross(e01,e02) points outwards
    e1, e2 = P[idx[:, 1]] - P[idx[:, 0]], P[idx[:, 2]] - P[idx[:, 0]]
    flip = np.sum(np.cross(e1, e2) * P[idx[:, 0]], axis=1) < 0
    idx[flip] = idx[flip][:, [0, 2, 1]]
    # drop degenerate pole triangles
    e1, e2 = P[idx[:, 1]] - P[idx[:, 0]], P[idx[:, 2]] - P[idx[:, 0]]
    idx = idx[np.linalg.norm(np.cross(e1, e2), axis=1) > 1e-9]
    rock = Mesh(positions=P, normals=N, uvs=UV, indices=idx, material_id=0, flags=MESH_IS_OPAQUE)
    ground = _mesh_from_parts([_quad((-200, 0, 200), (200, 0, 200), (200, 0, -200), (-200, 0, -200))], 1)
    nodes = []
    side = int(np.ceil(np.sqrt(n_instances)))
    for k in range(n_instances):
        gx, gz = k % side, k // side
        s = float(rng.uniform(0.6, 1.6))
        ang = float(rng.uniform(0, 2 * np.pi))
        tilt = float(rng.uniform(-0.3, 0.3))
        cy, sy = np.cos(ang), np.sin(ang)
        ct, st = np.cos(tilt), np.sin(tilt)
        Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        Rx = np.array([[1, 0, 0], [0, ct, -st], [0, st, ct]])
        T = np.eye(4)
        T[:3, :3] = (Ry @ Rx) * s
        T[:3, 3] = [(gx - side / 2) * 4.0 + rng.uniform(-1, 1), s * 0.6, (gz - side / 2) * 4.0 + rng.uniform(-1, 1)]
        nodes.append(Node(meshes=[0], transform=T.astype(np.float32)))
    nodes.append(Node(meshes=[1]))
    mats = [GltfMaterial(base_color_factor=(0.55, 0.5, 0.45, 1), metallic_factor=0.0, roughness_factor=0.6),
            GltfMaterial(base_color_factor=(0.3, 0.35, 0.25, 1), metallic_factor=0.0, roughness_factor=0.9)]
    ext = side * 2.0
    cam = CameraDesc(position=(0.0, ext * 0.5, ext * 1.2), target=(0, 0, 0), yfov_deg=45.0, aspect=aspect, near=0.05, far=1e4)
    return SceneDescription(meshes=[rock, ground], nodes=nodes, materials=mats, env=sky_env(*env_size), camera=cam,
                            name=f"rocks_{n_instances}", seed=seed)


# ----------------------------------------------------------------------------------------------- C2 rays
def incoherent_rays(cam: abi.Camera, width: int, height: int, seed_frame: int = 0) -> np.ndarray:
    """C2 incoherent set: camera origins, cosine-hemisphere directions (about -y... the camera forward axis)
    drawn from PCG seeded with pcg4d16(px, py, frame, 0)."""
    px, py = np.meshgrid(np.arange(width, dtype=np.uint32), np.arange(height, dtype=np.uint32))
    seeds = pcg4d16_np(px.reshape(-1), py.reshape(-1), np.uint32(seed_frame), np.uint32(0))
    u0, st = pcg_next_float_np(seeds)
    u1, st = pcg_next_float_np(st)
    su = np.sqrt(u0)
    lx, ly, lz = su * np.cos(2 * np.pi * u1), su * np.sin(2 * np.pi * u1), np.sqrt(1 - u0)
    W = np.array(cam.W[:], np.float32)
    W = W / np.linalg.norm(W)
    U = np.array(cam.U[:], np.float32)
    U = U / np.linalg.norm(U)
    V = np.cross(W, U)
    d = lx[:, None] * U[None] + ly[:, None] * V[None] + lz[:, None] * W[None]
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays = np.zeros(width * height, abi.RAY_DTYPE)
    rays["origin"] = np.array(cam.position[:], np.float32)[None]
    rays["tmin"] = 0.0
    rays["direction"] = d
    rays["tmax"] = 1e7
    return rays


def pcg4d16_np(x, y, z, w):
    """pcg4d16, include/merian-shaders/utils/hash.slang:138-146 (vectorised, uint32 wrap-around)."""
    with np.errstate(over="ignore"):
        x = np.asarray(x, np.uint32) * np.uint32(1664525) + np.uint32(1013904223)
        y = np.asarray(y, np.uint32) * np.uint32(1664525) + np.uint32(1013904223)
        z = np.asarray(z, np.uint32) * np.uint32(1664525) + np.uint32(1013904223)
        w = np.asarray(w, np.uint32) * np.uint32(1664525) + np.uint32(1013904223)
        x = x + y * w
        y = y + z * x
        z = z + x * y
        w = w + y * z
        x = x + y * w
    return x.astype(np.uint32)


def pcg_next_float_np(state):
    """PCG::next + uint_to_zero_one_float_rust (random.slang:108-128, bit-twiddling.slang:23-27); returns (float, new state)."""
    with np.errstate(over="ignore"):
        s = np.asarray(state, np.uint32)
        new = s * np.uint32(747796405) + np.uint32(2891336453)
        word = ((s >> ((s >> np.uint32(28)) + np.uint32(4))) ^ s) * np.uint32(277803737)
        out = (word >> np.uint32(22)) ^ word
    return (out >> np.uint32(8)).astype(np.float32) * np.float32(5.9604644775390625e-08), new.astype(np.uint32)
