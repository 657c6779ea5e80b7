"""Context: the Python face of the C-ABI (include/mpt.h). PyTorch supplies device memory and the stream; every
computation happens in libmpt.so."""
import ctypes as C

import numpy as np
import torch

from . import abi
from ._lib import MptError, lib


def _hp(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Context:
    def __init__(self, device: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("merian_b200.Context needs a CUDA device (B200); there is no CPU path")
        self.l = lib()
        self.device = device
        self.torch_device = torch.device("cuda", device)
        torch.cuda.init()
        self.h = C.c_void_p()
        st = self.l.mpt_create(device, C.byref(self.h))
        if st != abi.MPT_OK:
            raise MptError(st, self.l.mpt_last_error(None).decode())

    # ------------------------------------------------------------------ plumbing
    def _check(self, st):
        if st != abi.MPT_OK:
            raise MptError(st, self.l.mpt_last_error(self.h).decode())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.torch_device).cuda_stream)

    def close(self):
        if getattr(self, "h", None):
            self.l.mpt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ uploads (host arrays)
    def scene_upload(self, scene_desc):
        self._check(self.l.mpt_scene_upload(self.h, C.byref(scene_desc)))

    def materials_upload(self, entries, entry_size, count, types, type_count):
        self._check(self.l.mpt_materials_upload(self.h, entries, entry_size, count, C.cast(types, C.c_void_p), type_count))

    def materials_set_hints(self, alpha_test_threshold=0.5, min_roughness=0.0316, clamp_normals=True):
        self._check(self.l.mpt_materials_set_hints(self.h, alpha_test_threshold, min_roughness, 1 if clamp_normals else 0))

    def texture_upload(self, tid, w, h, fmt, wrap_u, wrap_v, filt, pixels, mipmap=abi.MIPMAP_NONE):
        self._check(self.l.mpt_texture_upload_ex(self.h, tid, w, h, fmt, wrap_u, wrap_v, filt, mipmap, _hp(pixels)))

    def texture_upload_device(self, tid, w, h, fmt, wrap_u, wrap_v, filt, d_pixels, mipmap=abi.MIPMAP_NONE, row_pitch=0):
        """Texels already in device memory (a torch CUDA tensor): the Vulkan -> CUDA import leg (include/mpt.h)."""
        self._check(self.l.mpt_texture_upload_device(self.h, tid, w, h, fmt, wrap_u, wrap_v, filt, mipmap, C.c_void_p(d_pixels.data_ptr()), row_pitch,
                                                     self._stream()))

    def env_set_device(self, typ, to_local, intensity, w, h, d_pixels):
        self._check(self.l.mpt_env_set_device(self.h, typ, _hp(to_local), intensity, w, h, C.c_void_p(d_pixels.data_ptr())))

    def env_set(self, typ, to_local, intensity, w, h, pixels):
        self._check(self.l.mpt_env_set(self.h, typ, _hp(to_local), intensity, w, h, _hp(pixels)))

    def env_load(self, path, to_local, intensity=1.0):
        """Lat-long .pfm straight from disk (Scene::process_pending_env_load)."""
        import os
        self._check(self.l.mpt_env_load(self.h, os.fsencode(path), _hp(np.ascontiguousarray(to_local, np.float32)), intensity))

    def image_write(self, path, image):
        """PFM leg of the Image Write node for a float32 (H, W, 4) device tensor."""
        import os
        h, w = image.shape[:2]
        self._check(self.l.mpt_image_write(self.h, os.fsencode(path), C.c_void_p(image.data_ptr()), w, h, self._stream()))

    # ------------------------------------------------------------------ device entry points (torch tensors)
    def set_scene_frame(self, frame):
        """Scene::frame for primary_visibility / gbuffer (thin-lens cameras key their lens sample on it); render() takes it from its params."""
        self._check(self.l.mpt_set_scene_frame(self.h, C.c_uint32(int(frame))))

    def set_prev_transforms(self, transforms):
        """Scene::prev_instance_transforms for the GBuffer's motion vectors: (instance_count, 4, 4) float32 row-major, None = static scene."""
        if transforms is None:
            self._check(self.l.mpt_scene_set_prev_transforms(self.h, None, 0))
            return
        t = np.ascontiguousarray(transforms, np.float32).reshape(-1, 16)
        self._check(self.l.mpt_scene_set_prev_transforms(self.h, C.c_void_p(t.ctypes.data), C.c_uint32(len(t))))

    def primary_visibility_device(self, cam, w, h, mask=0xFF, out=None):
        if out is None:
            out = torch.empty((h, w, 4), dtype=torch.int32, device=self.torch_device)
        self._check(self.l.mpt_primary_visibility(self.h, C.byref(cam), w, h, mask, C.c_void_p(out.data_ptr()), self._stream()))
        return out

    def primary_visibility(self, cam, w, h, mask=0xFF):
        t = self.primary_visibility_device(cam, w, h, mask)
        return t.cpu().numpy().view(abi.PACKED_HIT_DTYPE).reshape(h, w)

    def gbuffer_device(self, cam, w, h, prev_cam=None, mask=0xFF):
        """The GBuffer (Raytraced) node: dict of device tensors tex0 (H,W,2 i32), tex1 (H,W,4 i32), tex2, tex3, emission (H,W,4 f32)."""
        dev = self.torch_device
        g = {"tex0": torch.empty((h, w, 2), dtype=torch.int32, device=dev), "tex1": torch.empty((h, w, 4), dtype=torch.int32, device=dev),
             "tex2": torch.empty((h, w, 4), dtype=torch.float32, device=dev), "tex3": torch.empty((h, w, 4), dtype=torch.float32, device=dev),
             "emission": torch.empty((h, w, 4), dtype=torch.float32, device=dev)}
        imgs = abi.GBufferImages(*[C.c_void_p(g[k].data_ptr()) for k in ("tex0", "tex1", "tex2", "tex3", "emission")])
        self._check(self.l.mpt_gbuffer(self.h, C.byref(cam), C.byref(prev_cam) if prev_cam is not None else None, w, h, mask, C.byref(imgs), self._stream()))
        return g

    def gbuffer(self, cam, w, h, prev_cam=None, mask=0xFF):
        g = {k: v.cpu().numpy() for k, v in self.gbuffer_device(cam, w, h, prev_cam, mask).items()}
        g["tex0"] = g["tex0"].view(np.uint32)
        g["tex1"] = g["tex1"].view(abi.PACKED_HIT_DTYPE).reshape(h, w)
        return g

    def trace_device(self, rays, mask=0xFF, hits=None):
        n = rays.shape[0]
        if hits is None:
            hits = torch.empty((n, 6), dtype=torch.int32, device=self.torch_device)
        self._check(self.l.mpt_trace(self.h, C.c_void_p(rays.data_ptr()), n, mask, C.c_void_p(hits.data_ptr()), self._stream()))
        return hits

    def trace(self, rays_np, mask=0xFF):
        rays = torch.from_numpy(np.ascontiguousarray(rays_np).view(np.float32).reshape(-1, 8)).to(self.torch_device)
        hits = self.trace_device(rays, mask)
        return hits.cpu().numpy().view(abi.HIT_DTYPE).reshape(-1)

    def render_device(self, cam, params, tex1=None, albedo=None, out=None):
        if out is None:
            out = torch.empty((params.height, params.width, 4), dtype=torch.float32, device=self.torch_device)
        self._check(self.l.mpt_render(self.h, C.byref(cam), C.byref(params), C.c_void_p(tex1.data_ptr()) if tex1 is not None else None,
                                      C.c_void_p(albedo.data_ptr()) if albedo is not None else None, C.c_void_p(out.data_ptr()), self._stream()))
        return out

    def render(self, cam, params, tex1=None, albedo=None):
        t1 = torch.from_numpy(np.ascontiguousarray(tex1).view(np.int32).reshape(params.height, params.width, 4)).to(self.torch_device) if tex1 is not None else None
        al = torch.from_numpy(np.ascontiguousarray(albedo, dtype=np.float32)).to(self.torch_device) if albedo is not None else None
        return self.render_device(cam, params, t1, al).cpu().numpy()

    def render_host(self, cam, params, h_tex1=None, h_albedo=None, h_out=None):
        """End-to-end call on HOST tensors (pinned recommended): H2D tex1, render, D2H irradiance, stream sync."""
        if h_out is None:
            h_out = torch.empty((params.height, params.width, 4), dtype=torch.float32).pin_memory()
        self._check(self.l.mpt_render_host(self.h, C.byref(cam), C.byref(params), C.c_void_p(h_tex1.data_ptr()) if h_tex1 is not None else None,
                                           C.c_void_p(h_albedo.data_ptr()) if h_albedo is not None else None, C.c_void_p(h_out.data_ptr()),
                                           self._stream()))
        return h_out

    def render_host_begin(self, cam, params, h_out, slot, h_tex1=None, h_albedo=None):
        """One iteration in flight slot `slot` (H2D, render, D2H enqueued; returns immediately). Pair with render_host_end(slot)."""
        self._check(self.l.mpt_render_host_begin(self.h, C.byref(cam), C.byref(params), C.c_void_p(h_tex1.data_ptr()) if h_tex1 is not None else None,
                                                 C.c_void_p(h_albedo.data_ptr()) if h_albedo is not None else None, C.c_void_p(h_out.data_ptr()), slot))

    def render_host_end(self, slot):
        self._check(self.l.mpt_render_host_end(self.h, slot))

    # ------------------------------------------------------------------ Accumulate node, plain mode (include/mpt.h)
    @staticmethod
    def _acc_buffers(prev_out, prev_history, out, history, tex0, tex2):
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        return abi.AccumulateBuffers(p(prev_out), p(prev_history), p(out), p(history), p(tex0), p(tex2))

    def accumulate_device(self, acc_params, src, prev_out, prev_history, out, history, tex0=None, tex2=None):
        """src/prev_out/out: float32 (H, W, 4); prev_history/history: int32 (H, W, 2) holding the uint2 texels."""
        b = self._acc_buffers(prev_out, prev_history, out, history, tex0, tex2)
        self._check(self.l.mpt_accumulate(self.h, C.byref(acc_params), C.c_void_p(src.data_ptr()), C.byref(b), self._stream()))
        return out, history

    def render_accumulate_device(self, cam, params, acc_params, prev_out, prev_history, out, history, tex1=None, albedo=None, tex0=None, tex2=None):
        b = self._acc_buffers(prev_out, prev_history, out, history, tex0, tex2)
        self._check(self.l.mpt_render_accumulate(self.h, C.byref(cam), C.byref(params), C.c_void_p(tex1.data_ptr()) if tex1 is not None else None,
                                                 C.c_void_p(albedo.data_ptr()) if albedo is not None else None, C.byref(acc_params), C.byref(b),
                                                 self._stream()))
        return out, history

    # ------------------------------------------------------------------ counters
    def set_profiling(self, flags):
        self._check(self.l.mpt_set_profiling(self.h, flags))

    def stats(self) -> abi.Stats:
        s = abi.Stats()
        self._check(self.l.mpt_get_stats(self.h, C.byref(s)))
        return s

    def reset_stats(self):
        self._check(self.l.mpt_reset_stats(self.h))

    def debug_ray_queue(self, bounce, capacity=None):
        """The continuation rays the last extend launch of the last frame consumed (bounce = max_path_length - 1) and the triangle
        record each of them starts on. Returns (rays RAY_DTYPE array, leaves uint32 array)."""
        n = C.c_uint64(0)
        self._check(self.l.mpt_debug_ray_queue(self.h, bounce, None, None, 0, C.byref(n)))
        m = int(n.value) if capacity is None else min(int(n.value), capacity)
        rays = np.zeros(m, abi.RAY_DTYPE)
        leaves = np.zeros(m, np.uint32)
        if m:
            self._check(self.l.mpt_debug_ray_queue(self.h, bounce, _hp(rays), _hp(leaves), m, C.byref(n)))
        return rays, leaves

    def debug_ray_setup(self, dirs):
        """Per-ray traversal setup of n directions (n, 3) float32: (shear (n, 3), k bits (n,), perm (n,), idir IEEE (n, 3), idir 1-ulp (n, 3))."""
        dirs = np.ascontiguousarray(dirs, np.float32)
        n = len(dirs)
        shear, kp, idir = np.zeros((n, 3), np.float32), np.zeros((n, 2), np.uint32), np.zeros((n, 6), np.float32)
        self._check(self.l.mpt_debug_ray_setup(self.h, _hp(dirs), C.c_uint64(n), _hp(shear), _hp(kp), _hp(idir)))
        return shear, kp[:, 0], kp[:, 1], idir[:, :3], idir[:, 3:]

    def bvh_info(self) -> abi.BvhInfo:
        b = abi.BvhInfo()
        self._check(self.l.mpt_bvh_get_info(self.h, C.byref(b)))
        return b

    def bsdf_check(self, **kw):
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
        self._check(self.l.mpt_bsdf_check(self.h, C.byref(p), _hp(out)))
        return out
