"""The drop-in boundary: libmpt.so loads on a machine without a GPU, exports every entry point include/mpt.h declares
(and nothing is declared that the binding does not bind), and the ctypes mirrors agree with the C layouts.
No compute call is made here (those are the -m gpu tests)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from merian_b200 import abi
from merian_b200._lib import EXPORTS, LIB_PATH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mpt.h")
HEADERS = [HEADER, os.path.join(ROOT, "include", "mpt_node.h"), os.path.join(ROOT, "include", "mpt_tables.h")]


def _declared():
    src = "".join(open(h).read() for h in HEADERS)
    return sorted(set(re.findall(r"MPT_API\s+[\w\s\*]+?\b(mpt[nt]?_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def libmpt():
    if not os.path.exists(LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return C.CDLL(LIB_PATH)


def test_header_and_binding_list_agree():
    assert _declared() == sorted(EXPORTS)


def test_library_exports_every_declared_symbol(libmpt):
    for name in _declared():
        assert hasattr(libmpt, name), name
    libmpt.mpt_abi_version.restype = C.c_uint32
    assert libmpt.mpt_abi_version() == abi.MPT_ABI_VERSION


def test_only_mpt_symbols_are_exported():
    out = subprocess.check_output(["nm", "-D", "--defined-only", LIB_PATH], text=True)
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert sorted(s for s in syms if s.startswith(("mpt_", "mptn_", "mptt_"))) == sorted(EXPORTS)
    assert not [s for s in syms if not s.startswith(("mpt_", "mptn_", "mptt_")) and not s.startswith("_")], syms


def test_no_gpu_fails_loudly(libmpt):
    """There is no CPU fallback: without a CUDA device mpt_create must fail with an explanation."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    libmpt.mpt_last_error.restype = C.c_char_p
    libmpt.mpt_last_error.argtypes = [C.c_void_p]
    h = C.c_void_p()
    st = libmpt.mpt_create(0, C.byref(h))
    assert st != abi.MPT_OK and not h.value
    assert b"CUDA" in libmpt.mpt_last_error(None) or b"device" in libmpt.mpt_last_error(None)
    from merian_b200.context import Context
    with pytest.raises(RuntimeError):
        Context(0)


def test_struct_layouts_match_header():
    """Compile a tiny C program against include/mpt.h and compare sizeof/offsetof with the ctypes mirrors."""
    import tempfile
    checks = [("mpt_packed_vertex", abi.PackedVertex), ("mpt_mesh_desc", abi.MeshDesc), ("mpt_blas_desc", abi.BlasDesc),
              ("mpt_geometry_desc", abi.GeometryDesc), ("mpt_instance_desc", abi.InstanceDesc), ("mpt_scene_desc", abi.SceneDesc),
              ("mpt_material_type", abi.MaterialType), ("mpt_camera", abi.Camera), ("mpt_render_params", abi.RenderParams),
              ("mpt_ray", abi.Ray), ("mpt_hit", abi.Hit), ("mpt_stats", abi.Stats), ("mpt_bvh_info", abi.BvhInfo),
              ("mpt_bsdf_check_params", abi.BsdfCheckParams), ("mpt_accumulate_params", abi.AccumulateParams),
              ("mpt_accumulate_buffers", abi.AccumulateBuffers), ("mpt_gbuffer_images", abi.GBufferImages)]
    body = "".join(f'printf("{n} %zu\\n", sizeof({n}));\n' for n, _ in checks)
    for n, t in checks:
        for f, _ in t._fields_:
            body += f'printf("{n}.{f} %zu\\n", offsetof({n}, {f}));\n'
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "l.c")
        open(src, "w").write(f'#include <stdio.h>\n#include <stddef.h>\n#include "{HEADER}"\nint main(void){{\n{body}return 0;}}\n')
        exe = os.path.join(d, "l")
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-o", exe, src])  # the header is plain C
        got = dict(l.split() for l in subprocess.check_output([exe], text=True).splitlines())
    for n, t in checks:
        assert int(got[n]) == C.sizeof(t), n
        for f, _ in t._fields_:
            assert int(got[f"{n}.{f}"]) == getattr(t, f).offset, (n, f)
