"""The BVH8 builder's shape guarantees (host-only entry point mpt_bvh8_build_info): every primitive lands in exactly one leaf slot,
and the depth stays inside what the traversal stack holds even for inputs that make SAH degenerate."""
import ctypes as C

import numpy as np
import pytest

from merian_b200 import abi
from merian_b200._lib import lib

MAX_DEPTH = 14  # bvh8.h kMaxBvh8Depth


def build(boxes):
    boxes = np.ascontiguousarray(boxes, np.float32)
    info = abi.BvhInfo()
    order = np.zeros(len(boxes), np.uint32)
    assert lib().mpt_bvh8_build_info(C.c_void_p(boxes.ctypes.data), len(boxes), C.byref(info), C.c_void_p(order.ctypes.data)) == abi.MPT_OK
    return info, order


def grid_boxes(nx, ny):
    i, j = np.meshgrid(np.arange(nx), np.arange(ny))
    x, z = i.reshape(-1) * 0.1, j.reshape(-1) * 0.2
    y = np.sin(x * 1.3) * np.cos(z * 0.7)
    lo = np.stack([x, y - 0.05, z], -1)
    return np.concatenate([lo, lo + [0.1, 0.1, 0.2]], -1)


@pytest.mark.parametrize("name", ["grid", "chain", "coincident", "single", "random"])
def test_every_primitive_once_and_bounded_depth(name):
    rng = np.random.default_rng(0)
    if name == "grid":
        boxes = grid_boxes(300, 200)
    elif name == "chain":   # sizes grow geometrically along a line: SAH peels one primitive per level (depth 25 before the fallback)
        s = 1e-3 * 1.02 ** np.arange(2000)
        x = np.concatenate([[0], np.cumsum(s)[:-1]])
        boxes = np.stack([x, np.zeros_like(x), np.zeros_like(x), x + s, s, s], -1)
    elif name == "coincident":  # identical centroids: the builder must fall back to index splits
        boxes = np.tile(np.array([[0, 0, 0, 1, 1, 1]], np.float32), (1000, 1))
    elif name == "single":
        boxes = np.array([[0, 0, 0, 1, 2, 3]], np.float32)
    else:
        c = rng.random((50000, 3)) * 100
        r = (rng.random((50000, 1)) ** 3) * 5 + 1e-3
        boxes = np.concatenate([c - r, c + r], -1)
    info, order = build(boxes)
    assert info.triangle_count == len(boxes)
    assert np.array_equal(np.sort(order), np.arange(len(boxes)))     # a permutation: each primitive in exactly one leaf slot
    assert 1 <= info.max_depth <= MAX_DEPTH
    assert info.node_count >= max(1, len(boxes) // 24) and info.node_bytes == info.node_count * 80
    # 8-wide with <= 3 primitives per leaf slot: at most n internal-ish nodes, typically n / 5
    assert info.node_count <= len(boxes) + 1


def test_typical_scene_depth_is_small():
    info, _ = build(grid_boxes(500, 400))
    assert info.max_depth <= 9


@pytest.mark.gpu
def test_degenerate_geometry_renders_like_the_oracle(gpu_ctx):
    """Triangles of geometrically growing size along a line (the input that drove the SAH tree 25 levels deep) upload, and primary
    visibility stays bit-identical to the oracle."""
    from merian_b200.scene import MESH_IS_OPAQUE, CameraDesc, DiffuseMaterial, Mesh, Node, SceneDescription, SceneTables, upload_scene
    from helpers import compare_hits, make_oracle
    n = 1200
    s = (1e-3 * 1.012 ** np.arange(n)).astype(np.float32)
    x = np.concatenate([[0], np.cumsum(s)[:-1]]).astype(np.float32)
    P = np.zeros((n, 3, 3), np.float32)
    P[:, 0] = np.stack([x, np.zeros(n), np.zeros(n)], -1)
    P[:, 1] = np.stack([x + s, np.zeros(n), np.zeros(n)], -1)
    P[:, 2] = np.stack([x, s * 40, np.zeros(n)], -1)
    pos = P.reshape(-1, 3)
    mesh = Mesh(positions=pos, normals=np.tile([[0, 0, 1]], (len(pos), 1)).astype(np.float32), uvs=np.zeros((len(pos), 2), np.float32), indices=None,
                material_id=0, flags=MESH_IS_OPAQUE | 0x10, instance_mask=0xFF)
    xm = float(x[-1])
    t = SceneTables(SceneDescription(meshes=[mesh], nodes=[Node([0])], materials=[DiffuseMaterial((0.5, 0.5, 0.5, 1))],
                                     camera=CameraDesc(position=(xm * 0.5, xm * 0.3, xm * 1.2), target=(xm * 0.5, xm * 0.2, 0.0), yfov_deg=50.0, aspect=2.0,
                                                       near=1e-3, far=1e7)))
    upload_scene(gpu_ctx, t)
    assert gpu_ctx.bvh_info().max_depth <= MAX_DEPTH
    o = make_oracle(t)
    cam = t.desc.camera.to_abi(0)
    ref, got = o.primary_visibility(cam, 256, 128), gpu_ctx.primary_visibility(cam, 256, 128)
    mism, bary, _ = compare_hits(got, ref)
    assert (ref["instance_id"] != 0xFFFFFFFF).mean() > 0.02
    assert mism < 1e-4 and bary == 0.0


def test_tree_does_not_depend_on_the_thread_count():
    """The parallel builder (top of the tree split by all threads in fixed chunks, subtrees as tasks, node ids in per-thread blocks) must
    emit the same BVH8 whatever the number of threads: 150 k boxes (above the 131 072 threshold of Builder::split_top) built with 1, 3 and
    8 OpenMP threads in fresh processes."""
    import hashlib
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    prog = (
        "import sys, hashlib, ctypes as C, numpy as np\n"
        f"sys.path.insert(0, {root!r})\n"
        "from merian_b200 import abi\n"
        "from merian_b200._lib import lib\n"
        "rng = np.random.default_rng(7)\n"
        "c = rng.random((150000, 3)) * [200, 20, 100]\n"
        "r = (rng.random((150000, 1)) ** 4) * 3 + 1e-3\n"
        "boxes = np.ascontiguousarray(np.concatenate([c - r, c + r], -1), np.float32)\n"
        "info = abi.BvhInfo(); order = np.zeros(len(boxes), np.uint32)\n"
        "assert lib().mpt_bvh8_build_info(C.c_void_p(boxes.ctypes.data), len(boxes), C.byref(info), C.c_void_p(order.ctypes.data)) == abi.MPT_OK\n"
        "print(info.node_count, info.max_depth, hashlib.sha256(order.tobytes()).hexdigest())\n")
    outs = []
    for threads in ("1", "3", "8"):
        env = dict(os.environ, OMP_NUM_THREADS=threads)
        outs.append(subprocess.check_output([sys.executable, "-c", prog], env=env, text=True).strip())
    assert outs[0] == outs[1] == outs[2], outs
