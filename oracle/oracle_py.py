"""ctypes loader for the CPU oracle (oracle/_build/liboracle.so). TEST INFRASTRUCTURE ONLY.

Allowed importers: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
The product package (merian_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from merian_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")


def build(force: bool = False):
    srcs = [os.path.join(_HERE, f) for f in ("oracle.cpp", "oracle_math.hpp", "oracle_bsdf.hpp", "oracle_bvh8.hpp")]
    if not force and os.path.exists(_LIB_PATH) and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s"], env={**os.environ, "MAKEFLAGS": ""})
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_create.restype = C.c_void_p
        _lib.oracle_last_error.restype = C.c_char_p
        _lib.oracle_half_to_float.restype = C.c_float
        _lib.oracle_set_scene_frame.argtypes = [C.c_void_p, C.c_uint32]
        _lib.oracle_set_scene_frame.restype = None
        _lib.oracle_scene_set_prev_transforms.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        _lib.oracle_scene_set_prev_transforms.restype = None
        _lib.oracle_half_to_float.argtypes = [C.c_uint16]
        _lib.oracle_float_to_half.restype = C.c_uint16
        _lib.oracle_float_to_half.argtypes = [C.c_float]
        _lib.oracle_pcg4d16.restype = C.c_uint32
        _lib.oracle_pcg4d16.argtypes = [C.c_uint32] * 4
        _lib.oracle_encode_normal.restype = C.c_uint32
        _lib.oracle_encode_tangent.restype = C.c_uint32
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class OracleContext:
    """Same method surface as merian_b200.Context, on host arrays."""

    def __init__(self):
        self.l = lib()
        self.h = C.c_void_p(self.l.oracle_create())
        self.last_rays = 0

    def close(self):
        if self.h:
            self.l.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def scene_upload(self, scene_desc):
        assert self.l.oracle_scene_upload(self.h, C.byref(scene_desc)) == 0

    def materials_upload(self, entries, entry_size, count, types, type_count):
        assert self.l.oracle_materials_upload(self.h, entries, C.c_uint32(entry_size), C.c_uint32(count), types, C.c_uint32(type_count)) == 0

    def materials_set_hints(self, alpha_test_threshold=0.5, min_roughness=0.0316, clamp_normals=True):
        assert self.l.oracle_materials_set_hints(self.h, C.c_float(alpha_test_threshold), C.c_float(min_roughness), C.c_uint32(1 if clamp_normals else 0)) == 0

    def texture_upload(self, tid, w, h, fmt, wrap_u, wrap_v, filt, pixels, mipmap=abi.MIPMAP_NONE):
        assert self.l.oracle_texture_upload_ex(self.h, C.c_uint32(tid), C.c_uint32(w), C.c_uint32(h), C.c_uint32(fmt), C.c_uint32(wrap_u),
                                               C.c_uint32(wrap_v), C.c_uint32(filt), C.c_uint32(mipmap), _p(pixels)) == 0

    def texture_level_texel(self, tid, level, x, y):
        """(decoded texel, (width, height)) of one mip level (test hook)."""
        out, dims = np.zeros(4, np.float32), np.zeros(2, np.uint32)
        self.l.oracle_texture_level_texel(self.h, C.c_uint32(tid), C.c_uint32(level), C.c_uint32(x), C.c_uint32(y), _p(out), _p(dims))
        return out, (int(dims[0]), int(dims[1]))

    def sample_texture_grad(self, tid, u, v, ddx, ddy):
        out = np.zeros(4, np.float32)
        dx, dy = np.ascontiguousarray(ddx, np.float32), np.ascontiguousarray(ddy, np.float32)  # kept alive across the call
        self.l.oracle_sample_texture_grad(self.h, C.c_uint32(tid), C.c_float(u), C.c_float(v), _p(dx), _p(dy), _p(out))
        return out

    def env_set(self, typ, to_local, intensity, w, h, pixels):
        assert self.l.oracle_env_set(self.h, C.c_uint32(typ), _p(to_local), C.c_float(intensity), C.c_uint32(w), C.c_uint32(h), _p(pixels)) == 0

    def set_scene_frame(self, frame):
        self.l.oracle_set_scene_frame(self.h, C.c_uint32(int(frame)))

    def set_prev_transforms(self, transforms):
        """Scene::prev_instance_transforms: (instance_count, 4, 4) float32 row-major, None = static scene."""
        if transforms is None:
            self.l.oracle_scene_set_prev_transforms(self.h, None, 0)
            return
        t = np.ascontiguousarray(transforms, np.float32).reshape(-1, 16)
        self.l.oracle_scene_set_prev_transforms(self.h, C.c_void_p(t.ctypes.data), C.c_uint32(len(t)))

    def primary_visibility(self, cam, w, h, mask=0xFF):
        out = np.zeros((h, w), abi.PACKED_HIT_DTYPE)
        assert self.l.oracle_primary_visibility(self.h, C.byref(cam), C.c_uint32(w), C.c_uint32(h), C.c_uint32(mask), _p(out)) == 0
        return out

    def gbuffer(self, cam, w, h, prev_cam=None, mask=0xFF):
        g = {"tex0": np.zeros((h, w, 2), np.uint32), "tex1": np.zeros((h, w), abi.PACKED_HIT_DTYPE), "tex2": np.zeros((h, w, 4), np.float32),
             "tex3": np.zeros((h, w, 4), np.float32), "emission": np.zeros((h, w, 4), np.float32)}
        assert self.l.oracle_gbuffer(self.h, C.byref(cam), C.byref(prev_cam) if prev_cam is not None else None, C.c_uint32(w), C.c_uint32(h),
                                     C.c_uint32(mask), _p(g["tex0"]), _p(g["tex1"]), _p(g["tex2"]), _p(g["tex3"]), _p(g["emission"])) == 0
        return g

    def trace(self, rays, mask=0xFF, return_t=False):
        rays = np.ascontiguousarray(rays)
        hits = np.zeros(len(rays), abi.HIT_DTYPE)
        t = np.zeros(len(rays), np.float32)
        assert self.l.oracle_trace(self.h, _p(rays), C.c_uint64(len(rays)), C.c_uint32(mask), _p(hits), _p(t)) == 0
        return (hits, t) if return_t else hits

    def bvh8_trace(self, accel_view, rays, mask=0xFF, skip_records=None):
        """Closest hits over the PRODUCT's BVH8 (merian_b200.host_accel.HostAccel.view) with work counters.
        Returns (hits, t, records, counts) with counts = dict(rays, nodes, tris, instance_entries, max_stack)."""
        rays = np.ascontiguousarray(rays)
        n = len(rays)
        hits = np.zeros(n, abi.HIT_DTYPE)
        t = np.zeros(n, np.float32)
        rec = np.zeros(n, np.uint32)
        cnt = np.zeros(5, np.uint64)
        skip = None if skip_records is None else np.ascontiguousarray(skip_records, np.uint32)
        assert self.l.oracle_bvh8_trace(self.h, C.byref(accel_view), _p(rays), C.c_uint64(n), C.c_uint32(mask), _p(skip), _p(hits), _p(t), _p(rec), _p(cnt)) == 0
        return hits, t, rec, dict(zip(("rays", "nodes", "tris", "instance_entries", "max_stack"), (int(x) for x in cnt)))

    def render(self, cam, params, tex1=None, albedo=None):
        out = np.zeros((params.height, params.width, 4), np.float32)
        rays = C.c_uint64(0)
        assert self.l.oracle_render(self.h, C.byref(cam), C.byref(params), _p(tex1), _p(albedo), _p(out), C.byref(rays)) == 0
        self.last_rays = rays.value
        return out

    def num_threads(self):
        return self.l.oracle_num_threads()

    def set_num_threads(self, n):
        self.l.oracle_set_num_threads(C.c_int(int(n)))


def accumulate(acc_params, src, prev_out=None, prev_history=None, tex0=None, tex2=None):
    """Accumulate node, plain path (accumulate.slang). Returns (out float32 (H, W, 4), history uint32 (H, W, 2))."""
    h, w = acc_params.height, acc_params.width
    out = np.zeros((h, w, 4), np.float32)
    hist = np.zeros((h, w, 2), np.uint32)
    c = lambda a, dt: None if a is None else np.ascontiguousarray(a, dtype=dt)
    src, prev_out, prev_history, tex0, tex2 = c(src, np.float32), c(prev_out, np.float32), c(prev_history, np.uint32), c(tex0, np.uint32), c(tex2, np.float32)
    assert lib().oracle_accumulate(C.byref(acc_params), _p(src), _p(prev_out), _p(prev_history), _p(out), _p(hist), _p(tex0), _p(tex2)) == 0
    return out, hist


def bsdf_check(**kw):
    p = abi.BsdfCheckParams()
    p.bsdf = kw["bsdf"]
    p.wi[:] = kw["wi"]
    p.n_samples = kw.get("n_samples", 1 << 18)
    p.seed = kw.get("seed", 0x1234567)
    p.albedo[:] = kw.get("albedo", (1, 1, 1))
    p.alpha = kw.get("alpha", 0.3)
    p.f0[:] = kw.get("f0", (1, 1, 1))
    p.ior_f0 = kw.get("ior_f0", 0.04)
    p.alpha_bitangent = kw.get("alpha_bitangent", 0.3)
    p.iridescence = kw.get("iridescence", 0.0)
    out = np.zeros(16, np.float32)
    rc = lib().oracle_bsdf_check(C.byref(p), _p(out))
    assert rc == 0
    return out
