"""adapter/merian_plugin_pt_cuda.cpp is checked code: it type-checks (g++ -std=c++20 -fsyntax-only) against tests/adapter_stubs/, and every
declaration in the stubs is tied to the reference header it restates (`@ref header: text` must occur in that header when /root/reference
is present) or to adapter/merian_getters.patch (`@patch name`). merian itself (Vulkan SDK, slang, meson) cannot be built in this image."""
import glob
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUBS = os.path.join(ROOT, "tests", "adapter_stubs")
REF_INCLUDE = "/root/reference/include"


def _json_include():
    for base in sys.path + [os.path.dirname(os.__file__)]:
        for hit in glob.glob(os.path.join(base, "**", "nlohmann", "json.hpp"), recursive=True) if os.path.isdir(base) and "site-packages" in base else []:
            return os.path.dirname(os.path.dirname(hit))
    return None


def _cuda_include():
    for p in ("/usr/local/cuda/include", os.path.join(os.environ.get("CUDA_HOME", ""), "include")):
        if os.path.exists(os.path.join(p, "cuda_runtime_api.h")):
            return p
    return None


def test_adapter_compiles_against_stubs():
    gxx = shutil.which("g++")
    js, cu = _json_include(), _cuda_include()
    if not gxx or not js or not cu:
        pytest.skip(f"needs g++ ({gxx}), nlohmann/json.hpp ({js}) and cuda_runtime_api.h ({cu})")
    cmd = [gxx, "-std=c++20", "-fsyntax-only", "-Wall", "-Wextra", "-Wno-unused-parameter", "-I", STUBS, "-I", os.path.join(ROOT, "include"), "-I", cu, "-I", js,
           os.path.join(ROOT, "adapter", "merian_plugin_pt_cuda.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0 and "warning" not in r.stderr, r.stderr[-4000:]


def test_adapter_binds_every_interop_and_node_entry_point():
    src = open(os.path.join(ROOT, "adapter", "merian_plugin_pt_cuda.cpp")).read()
    for sym in ("mpt_import_external_memory_fd", "mpt_release_external_memory", "mpt_import_external_semaphore_fd", "mpt_wait_external_semaphore",
                "mpt_signal_external_semaphore", "mpt_release_external_semaphore", "mptn_process", "mptn_describe_outputs", "mptn_load_config", "mptn_dump_config",
                "mpt_scene_upload", "mpt_materials_upload", "mpt_materials_set_hints", "mpt_texture_upload_device", "mpt_env_set_device"):
        assert re.search(r"\b" + sym + r"\(", src), sym
    # the gbuffer connector is read (tex1 = hits, tex3 = albedo), not ignored
    assert "get_image(1)" in src and "get_image(3)" in src
    assert "nullptr, /*d_albedo*/" not in src and "sync_to_cpu_and_back([&, cam" not in src  # the round-1 host round trip is gone
    hdr = open(os.path.join(ROOT, "include", "mpt.h")).read() + open(os.path.join(ROOT, "include", "mpt_node.h")).read()
    for call in set(re.findall(r"\b(mptn?_[a-z_0-9]+)\(", src)):
        assert re.search(r"\b" + call + r"\(", hdr), f"{call} is not declared in include/*.h"


def _norm(s):
    return re.sub(r"\s+", " ", s).strip()


def test_stub_declarations_exist_in_the_reference():
    if not os.path.isdir(REF_INCLUDE):
        pytest.skip("reference not present (GPU box)")
    patch = open(os.path.join(ROOT, "adapter", "merian_getters.patch")).read()
    n_ref = n_patch = 0
    cache = {}
    for line in open(os.path.join(STUBS, "merian_stub.hpp")):
        m = re.search(r"// @ref (\S+): (.+)$", line)
        if m:
            path, text = m.group(1), _norm(m.group(2))
            if path not in cache:
                full = os.path.join(REF_INCLUDE, path)
                assert os.path.exists(full), f"stub cites a header the reference does not have: {path}"
                cache[path] = _norm(open(full).read())
            assert text in cache[path], f"{path} does not contain: {text}"
            n_ref += 1
        m = re.search(r"// @patch (\w+)", line)
        if m:
            assert re.search(r"^\+.*\b" + m.group(1) + r"\(", patch, re.M), f"adapter/merian_getters.patch does not add {m.group(1)}()"
            n_patch += 1
    assert n_ref > 80 and n_patch >= 10, (n_ref, n_patch)
    # and the other way round: every getter the patch adds is declared in the stub
    stub = open(os.path.join(STUBS, "merian_stub.hpp")).read()
    for name in re.findall(r"^\+\s+[^/\n(]*?\b(get_\w+)\(", patch, re.M):  # declarations: the getter's name is in front of the line's first "("
        assert re.search(r"\b" + name + r"\(.*// @patch", stub), name
