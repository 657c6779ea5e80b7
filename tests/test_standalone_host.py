"""examples/standalone_host.c: libmpt.so driven from plain C (mptt_build -> mpt_scene_upload -> mpt_materials_upload -> mpt_env_set ->
mpt_render_host -> mpt_pfm_save), the route a host without merian or Python takes. CPU: it compiles against include/*.h, links, and fails
loudly without a device. GPU: its image equals the same scene rendered through the Python mirror (same tables, same kernels)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "examples", "standalone_host.c")
LIBDIR = os.path.join(ROOT, "merian_b200", "csrc")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    if not os.path.exists(os.path.join(LIBDIR, "libmpt.so")):
        import __graft_entry__ as g
        g.build()
    out = str(tmp_path_factory.mktemp("host") / "standalone_host")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), SRC, "-L" + LIBDIR, "-lmpt",
                           "-Wl,-rpath," + LIBDIR, "-lm", "-o", out])
    return out


def test_c_host_builds_and_refuses_to_run_without_a_device(exe, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([exe, str(tmp_path / "x.pfm")], capture_output=True, text=True)
    assert r.returncode == 2 and "mpt_create" in r.stderr          # no CPU fallback: the library says why it cannot run
    assert not (tmp_path / "x.pfm").exists()


def _python_scene():
    """The scene of standalone_host.c through the Python mirror."""
    from merian_b200 import synthetic  # noqa: F401
    from merian_b200.scene import MESH_IS_OPAQUE, CameraDesc, DiffuseMaterial, EnvMap, Mesh, Node, SceneDescription, SceneTables
    f32 = np.float32

    def quad(a, b, c, d, n):
        return np.array([a, b, c, d], f32), np.tile(np.array([n], f32), (4, 1))
    qi = np.array([[0, 1, 2], [0, 2, 3]], np.uint16)
    fp, fn_ = quad([-2, 0, -2], [-2, 0, 2], [2, 0, 2], [2, 0, -2], [0, 1, 0])
    wp, wn = quad([-2, 0, -2], [2, 0, -2], [2, 2.5, -2], [-2, 2.5, -2], [0, 0, 1])
    h = 0.5
    c = np.array([[-h, -h, -h], [h, -h, -h], [h, h, -h], [-h, h, -h], [-h, -h, h], [h, -h, h], [h, h, h], [-h, h, h]], f32)
    face = [[4, 5, 6, 7], [1, 0, 3, 2], [5, 1, 2, 6], [0, 4, 7, 3], [7, 6, 2, 3], [0, 1, 5, 4]]
    fnorm = [[0, 0, 1], [0, 0, -1], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0]]
    bp = np.concatenate([c[f] for f in face])
    bn = np.concatenate([np.tile(np.array([n], f32), (4, 1)) for n in fnorm])
    bi = np.concatenate([qi + 4 * k for k in range(6)]).astype(np.uint16)
    meshes = [Mesh(positions=fp, normals=fn_, uvs=None, indices=qi, material_id=0, flags=MESH_IS_OPAQUE),
              Mesh(positions=wp, normals=wn, uvs=None, indices=qi, material_id=1, flags=MESH_IS_OPAQUE),
              Mesh(positions=bp, normals=bn, uvs=None, indices=bi, material_id=2, flags=MESH_IS_OPAQUE)]
    cs, sn = f32(np.cos(f32(0.5))), f32(np.sin(f32(0.5)))
    place = np.array([[cs, 0, sn, 0.3], [0, 1, 0, 0.5], [-sn, 0, cs, 0.2], [0, 0, 0, 1]], f32)
    sky = np.tile(np.array([0.9, 1.0, 1.2, 1.0], f32), (4, 8, 1))
    return SceneTables(SceneDescription(
        meshes=meshes, nodes=[Node([0, 1]), Node([2], place)],
        materials=[DiffuseMaterial((0.7, 0.7, 0.7, 1.0)), DiffuseMaterial((0.2, 0.4, 0.7, 1.0)), DiffuseMaterial((0.8, 0.3, 0.2, 1.0))],
        env=EnvMap(sky), camera=CameraDesc(position=(0.0, 1.4, 4.0), target=(0.0, 0.6, 0.0), yfov_deg=45.0, aspect=160 / 120, near=0.01, far=1e4)))


@pytest.mark.gpu
def test_c_host_renders_what_the_python_mirror_renders(exe, tmp_path, gpu_ctx):
    from merian_b200 import abi
    from merian_b200.image_io import load_pfm
    from merian_b200.scene import upload_scene
    out = tmp_path / "frame.pfm"
    r = subprocess.run([exe, str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = load_pfm(str(out))
    assert got.shape[:2] == (120, 160) and np.isfinite(got).all()
    t = _python_scene()
    upload_scene(gpu_ctx, t)
    ref = gpu_ctx.render(t.desc.camera.to_abi(0), abi.RenderParams(160, 120, 0, 64, 5, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0))
    lum = lambda a: (a[..., :3] * np.array([0.299, 0.587, 0.114], np.float32)).sum(-1)
    assert lum(ref).mean() > 0.1
    # same tables (tests/test_tables_cabi.py), same RNG, same kernels: the images agree pixel by pixel up to the last bits of the camera basis
    # (tanf / cosf of libm against numpy); a flipped decision on a path shows up as an isolated pixel
    close = np.isclose(got[..., :3], ref[..., :3], rtol=2e-3, atol=2e-4).all(-1)
    assert close.mean() > 0.97, close.mean()
    assert abs(lum(got).mean() - lum(ref).mean()) / lum(ref).mean() < 5e-3
