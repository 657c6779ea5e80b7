"""Mip chains of material textures and the gradient sampler of the GBuffer node: the product's host-side chain
builder (mpt_mip_level: what mpt_texture_upload_ex uploads) against the oracle's restatement of cmd_generate_mipmaps
(include/merian/vk/utils/blits.hpp:101-168), the oracle's SampleGrad against closed forms of the Vulkan LOD rules, and the glTF sampler
-> mip-chain rule of gltf_scene.cpp:212-248, 583-593. CPU only."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

from merian_b200 import abi, gltf, synthetic
from merian_b200._lib import lib
from oracle.oracle_py import OracleContext


def mip_level(px, fmt, level):
    px = np.ascontiguousarray(px)
    h, w = px.shape[:2]
    lw, lh = C.c_uint32(), C.c_uint32()
    assert lib().mpt_mip_level(C.c_void_p(px.ctypes.data), w, h, fmt, level, None, 0, C.byref(lw), C.byref(lh)) == abi.MPT_OK
    out = np.zeros((lh.value, lw.value, 4), px.dtype)
    assert lib().mpt_mip_level(C.c_void_p(px.ctypes.data), w, h, fmt, level, C.c_void_p(out.ctypes.data), out.nbytes, C.byref(lw), C.byref(lh)) == abi.MPT_OK
    return out


def srgb_to_linear(c):
    c = np.asarray(c, np.float64)
    return np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)


@pytest.mark.parametrize("fmt", [abi.TEX_RGBA8_UNORM, abi.TEX_RGBA8_SRGB, abi.TEX_RGBA32F])
@pytest.mark.parametrize("shape", [(16, 16), (12, 40), (7, 5), (1, 9)])
def test_product_chain_equals_oracle_chain(fmt, shape):
    rng = np.random.default_rng(3)
    h, w = shape
    px = rng.random((h, w, 4), np.float32) if fmt == abi.TEX_RGBA32F else rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    o = OracleContext()
    o.texture_upload(0, w, h, fmt, abi.WRAP_REPEAT, abi.WRAP_REPEAT, abi.FILTER_LINEAR, px, mipmap=abi.MIPMAP_LINEAR)
    levels = int(np.floor(np.log2(max(w, h)))) + 1  # resource_allocator.cpp:303
    assert lib().mpt_mip_level(C.c_void_p(px.ctypes.data), w, h, fmt, levels, None, 0, None, None) == abi.MPT_ERR_INVALID_ARGUMENT
    for l in range(levels):
        got = mip_level(px, fmt, l)
        assert got.shape[:2] == (max(1, h >> l), max(1, w >> l))
        _, dims = o.texture_level_texel(0, l, 0, 0)
        assert dims == (got.shape[1], got.shape[0])
        for (x, y) in [(0, 0), (got.shape[1] - 1, got.shape[0] - 1), (got.shape[1] // 2, got.shape[0] // 2)]:
            ref, _ = o.texture_level_texel(0, l, x, y)  # decoded
            t = got[y, x].astype(np.float64)
            if fmt != abi.TEX_RGBA32F:
                t = t / 255.0
                if fmt == abi.TEX_RGBA8_SRGB:
                    t[:3] = srgb_to_linear(t[:3])
            assert np.allclose(t, ref, atol=1e-6), (l, x, y, t, ref)
    o.close()


def test_even_sized_levels_are_two_by_two_box_means_in_linear_space():
    rng = np.random.default_rng(5)
    px = rng.integers(0, 256, (8, 8, 4), dtype=np.uint8)
    l1 = mip_level(px, abi.TEX_RGBA8_UNORM, 1).astype(np.float64)
    box = px.astype(np.float64).reshape(4, 2, 4, 2, 4).mean(axis=(1, 3))
    assert np.abs(l1 - box).max() <= 0.5 + 1e-9  # rounded to 8 bit
    s1 = mip_level(px, abi.TEX_RGBA8_SRGB, 1).astype(np.float64) / 255.0
    lin = srgb_to_linear(px[..., :3] / 255.0).reshape(4, 2, 4, 2, 3).mean(axis=(1, 3))
    assert np.abs(srgb_to_linear(s1[..., :3]) - lin).max() < 0.02  # filtered in linear space, re-encoded (one 8-bit step near white is 0.01)
    assert np.abs(s1[..., 3] * 255.0 - box[..., 3]).max() <= 0.5 + 1e-9  # alpha is linear


def test_sample_grad_follows_the_vulkan_lod_rules():
    o = OracleContext()
    # level l of this chain is the constant l: R = level index / 8 (each level uploaded explicitly would need an API; use a texture whose
    # box-filtered levels are distinguishable instead: a 1-texel checker averages to 0.5 from level 1 on)
    n = 64
    yy, xx = np.mgrid[0:n, 0:n]
    px = np.zeros((n, n, 4), np.float32)
    px[..., 0] = (xx + yy) % 2
    px[..., 3] = 1
    for mode in (abi.MIPMAP_NEAREST, abi.MIPMAP_LINEAR):
        o.texture_upload(0, n, n, abi.TEX_RGBA32F, abi.WRAP_REPEAT, abi.WRAP_REPEAT, abi.FILTER_NEAREST, px, mipmap=mode)
        centre = ((10 + 0.5) / n, (10 + 0.5) / n)  # texel (10, 10) = 0
        assert o.sample_texture_grad(0, *centre, (0.5 / n, 0), (0, 0.5 / n))[0] == 0.0      # rho = 0.5: magnification, level 0
        assert o.sample_texture_grad(0, *centre, (1.0 / n, 0), (0, 1.0 / n))[0] == 0.0      # rho = 1: lambda = 0, still level 0
        assert o.sample_texture_grad(0, *centre, (4.0 / n, 0), (0, 1.0 / n))[0] == 0.5      # lambda = 2: level 2 (the longer gradient decides)
        r = o.sample_texture_grad(0, *centre, (2.0 ** 0.5 / n, 0), (0, 0))[0]              # lambda = 0.5
        if mode == abi.MIPMAP_NEAREST:
            assert r in (0.0, 0.5)  # lambda = 0.5 exactly: the tie may go either way
        else:
            assert abs(r - 0.25) < 1e-6  # half way between level 0 (0) and level 1 (0.5)
        assert o.sample_texture_grad(0, *centre, (1e6, 0), (0, 0))[0] == 0.5                # clamped to the last level
    # a texture without a chain ignores the gradients
    o.texture_upload(1, n, n, abi.TEX_RGBA32F, abi.WRAP_REPEAT, abi.WRAP_REPEAT, abi.FILTER_NEAREST, px)
    assert o.sample_texture_grad(1, (10 + 0.5) / n, (10 + 0.5) / n, (8.0 / n, 0), (0, 8.0 / n))[0] == 0.0
    o.close()


def test_gltf_sampler_min_filter_decides_the_mip_chain():
    from merian_b200.scene import GltfMaterial, Texture
    desc = synthetic.cornell()
    rng = np.random.default_rng(1)
    modes = [(abi.FILTER_LINEAR, abi.MIPMAP_LINEAR), (abi.FILTER_LINEAR, abi.MIPMAP_NEAREST), (abi.FILTER_NEAREST, abi.MIPMAP_LINEAR),
             (abi.FILTER_LINEAR, abi.MIPMAP_NONE)]
    desc.textures = [Texture(rng.integers(0, 256, (8, 8, 4), dtype=np.uint8), srgb=True, filter=f, mipmap=m) for f, m in modes]
    for i, m in enumerate(desc.materials[:len(modes)]):
        if isinstance(m, GltfMaterial):
            m.base_color_texture = i
    used = sorted({m.base_color_texture for m in desc.materials if isinstance(m, GltfMaterial) and getattr(m, "base_color_texture", abi.TEXTURE_NONE) != abi.TEXTURE_NONE})
    with tempfile.TemporaryDirectory() as d:
        back = gltf.load_gltf(gltf.save_gltf(desc, os.path.join(d, "s.gltf")))
    assert len(back.textures) == len(used) > 0
    for new_id, old_id in enumerate(used):  # texture ids are assigned in first-use order
        assert (back.textures[new_id].filter, back.textures[new_id].mipmap) == modes[old_id]
