"""ctypes binding of libmpt.so (built in-tree from merian_b200/csrc). No fallback: a missing library or a
non-sm_100 device raises."""
import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# MPT_LIB_VARIANT=name loads csrc/libmpt_name.so (A/B builds for profiling sessions, see csrc/Makefile `variant`)
_VARIANT = os.environ.get("MPT_LIB_VARIANT", "")
LIB_PATH = os.path.join(_HERE, "csrc", f"libmpt_{_VARIANT}.so" if _VARIANT else "libmpt.so")
_lib = None

EXPORTS = ["mpt_abi_version", "mpt_create", "mpt_destroy", "mpt_last_error", "mpt_scene_upload", "mpt_materials_upload",
           "mpt_materials_set_hints", "mpt_texture_upload", "mpt_texture_upload_ex", "mpt_texture_upload_device", "mpt_mip_level", "mpt_env_set", "mpt_env_set_device", "mpt_set_scene_frame", "mpt_scene_set_prev_transforms", "mpt_primary_visibility", "mpt_gbuffer", "mpt_trace", "mpt_render", "mpt_render_host", "mpt_render_host_begin", "mpt_render_host_end",
           "mpt_accumulate", "mpt_render_accumulate", "mpt_pfm_info", "mpt_pfm_load", "mpt_pfm_save", "mpt_hdr_load", "mpt_env_load", "mpt_image_write",
           "mpt_set_profiling", "mpt_get_stats", "mpt_reset_stats", "mpt_debug_ray_queue", "mpt_debug_ray_setup", "mpt_bvh_get_info", "mpt_bvh8_build_info", "mpt_bsdf_check",
           "mpt_host_accel_build", "mpt_host_accel_get", "mpt_host_accel_free",
           "mpt_comm_unique_id", "mpt_comm_create", "mpt_comm_adopt", "mpt_comm_destroy", "mpt_comm_last_error", "mpt_comm_info",
           "mpt_reduce_begin", "mpt_reduce_end", "mpt_reduce_wait",
           "mpt_import_external_memory_fd", "mpt_release_external_memory", "mpt_import_external_semaphore_fd", "mpt_wait_external_semaphore",
           "mpt_signal_external_semaphore", "mpt_release_external_semaphore",
           # include/mpt_node.h
           "mptn_create", "mptn_destroy", "mptn_last_error", "mptn_load_config", "mptn_dump_config", "mptn_describe_outputs",
           "mptn_set_shard", "mptn_process",
           "mptn_gbuffer_create", "mptn_gbuffer_destroy", "mptn_gbuffer_last_error", "mptn_gbuffer_load_config", "mptn_gbuffer_dump_config",
           "mptn_gbuffer_set_resolution", "mptn_gbuffer_describe_outputs", "mptn_gbuffer_process",
           "mptn_accumulate_create", "mptn_accumulate_destroy", "mptn_accumulate_last_error", "mptn_accumulate_load_config",
           "mptn_accumulate_dump_config", "mptn_accumulate_request_clear", "mptn_accumulate_process", "mptn_process_accumulate",
           # include/mpt_tables.h
           "mptt_build", "mptt_scene_desc", "mptt_free", "mptt_pack_vertices", "mptt_materials_build", "mptt_materials_get", "mptt_materials_free", "mptt_camera_look_at", "mptt_gltf_load", "mptt_gltf_get", "mptt_gltf_free"]


class MptError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"libmpt status {status}: {message}")
        self.status = status


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU or PyTorch fallback for the path tracer)")
        l = C.CDLL(LIB_PATH)
        l.mpt_abi_version.restype = C.c_uint32
        l.mpt_last_error.restype = C.c_char_p
        l.mpt_last_error.argtypes = [C.c_void_p]
        l.mpt_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        l.mpt_destroy.argtypes = [C.c_void_p]
        l.mpt_destroy.restype = None
        l.mpt_scene_upload.argtypes = [C.c_void_p, C.POINTER(abi.SceneDesc)]
        l.mpt_materials_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32]
        l.mpt_materials_set_hints.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_uint32]
        l.mpt_texture_upload.argtypes = [C.c_void_p] + [C.c_uint32] * 7 + [C.c_void_p]
        l.mpt_texture_upload_ex.argtypes = [C.c_void_p] + [C.c_uint32] * 8 + [C.c_void_p]
        l.mpt_texture_upload_device.argtypes = [C.c_void_p] + [C.c_uint32] * 8 + [C.c_void_p, C.c_uint64, C.c_void_p]
        l.mpt_mip_level.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        l.mpt_env_set_device.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_float, C.c_uint32, C.c_uint32, C.c_void_p]
        l.mpt_env_set.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_float, C.c_uint32, C.c_uint32, C.c_void_p]
        l.mptt_build.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]
        l.mptt_scene_desc.argtypes = [C.c_void_p]
        l.mptt_scene_desc.restype = C.POINTER(abi.SceneDesc)
        l.mptt_free.argtypes = [C.c_void_p]
        l.mptt_free.restype = None
        l.mptt_materials_build.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]
        l.mptt_materials_get.argtypes = [C.c_void_p, C.c_void_p]
        l.mptt_materials_free.argtypes = [C.c_void_p]
        l.mptt_materials_free.restype = None
        l.mptt_camera_look_at.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.POINTER(abi.Camera)]
        l.mptt_gltf_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        l.mptt_gltf_get.argtypes = [C.c_void_p, C.c_void_p]
        l.mptt_gltf_free.argtypes = [C.c_void_p]
        l.mptt_gltf_free.restype = None
        l.mptt_pack_vertices.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        l.mpt_set_scene_frame.argtypes = [C.c_void_p, C.c_uint32]
        l.mpt_scene_set_prev_transforms.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        l.mpt_primary_visibility.argtypes = [C.c_void_p, C.POINTER(abi.Camera), C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        l.mpt_gbuffer.argtypes = [C.c_void_p, C.POINTER(abi.Camera), C.POINTER(abi.Camera), C.c_uint32, C.c_uint32, C.c_uint32,
                                  C.POINTER(abi.GBufferImages), C.c_void_p]
        l.mpt_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]
        l.mpt_render.argtypes = [C.c_void_p, C.POINTER(abi.Camera), C.POINTER(abi.RenderParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.mpt_render_host.argtypes = l.mpt_render.argtypes
        l.mpt_render_host_begin.argtypes = [C.c_void_p, C.POINTER(abi.Camera), C.POINTER(abi.RenderParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        l.mpt_render_host_end.argtypes = [C.c_void_p, C.c_uint32]
        l.mpt_accumulate.argtypes = [C.c_void_p, C.POINTER(abi.AccumulateParams), C.c_void_p, C.POINTER(abi.AccumulateBuffers), C.c_void_p]
        l.mpt_render_accumulate.argtypes = [C.c_void_p, C.POINTER(abi.Camera), C.POINTER(abi.RenderParams), C.c_void_p, C.c_void_p,
                                            C.POINTER(abi.AccumulateParams), C.POINTER(abi.AccumulateBuffers), C.c_void_p]
        l.mpt_pfm_info.argtypes = [C.c_char_p] + [C.POINTER(C.c_uint32)] * 3
        l.mpt_pfm_load.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint64]
        l.mpt_pfm_save.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        l.mpt_hdr_load.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_void_p, C.c_uint64]
        l.mpt_env_load.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_float]
        l.mpt_image_write.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        l.mpt_set_profiling.argtypes = [C.c_void_p, C.c_uint32]
        l.mpt_get_stats.argtypes = [C.c_void_p, C.POINTER(abi.Stats)]
        l.mpt_reset_stats.argtypes = [C.c_void_p]
        l.mpt_debug_ray_setup.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        l.mpt_debug_ray_queue.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        l.mpt_bvh_get_info.argtypes = [C.c_void_p, C.POINTER(abi.BvhInfo)]
        l.mpt_bvh8_build_info.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(abi.BvhInfo), C.c_void_p]
        l.mpt_host_accel_build.argtypes = [C.POINTER(abi.SceneDesc), C.POINTER(C.c_void_p)]
        l.mpt_host_accel_get.argtypes = [C.c_void_p, C.POINTER(abi.HostAccelView)]
        l.mpt_host_accel_free.argtypes = [C.c_void_p]
        l.mpt_host_accel_free.restype = None
        l.mpt_comm_unique_id.argtypes = [C.c_void_p]
        l.mpt_comm_create.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        l.mpt_comm_adopt.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        l.mpt_comm_destroy.argtypes = [C.c_void_p]
        l.mpt_comm_destroy.restype = None
        l.mpt_comm_last_error.argtypes = [C.c_void_p]
        l.mpt_comm_last_error.restype = C.c_char_p
        l.mpt_comm_info.argtypes = [C.POINTER(C.c_uint32)]
        l.mpt_reduce_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_void_p,
                                       C.c_uint32]
        l.mpt_reduce_end.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        l.mpt_reduce_wait.argtypes = [C.c_void_p, C.c_uint32]
        l.mpt_import_external_memory_fd.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        l.mpt_release_external_memory.argtypes = [C.c_void_p, C.c_void_p]
        l.mpt_import_external_semaphore_fd.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.POINTER(C.c_void_p)]
        l.mpt_wait_external_semaphore.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        l.mpt_signal_external_semaphore.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        l.mpt_release_external_semaphore.argtypes = [C.c_void_p, C.c_void_p]
        l.mpt_bsdf_check.argtypes = [C.c_void_p, C.POINTER(abi.BsdfCheckParams), C.c_void_p]
        l.mptn_create.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        l.mptn_destroy.argtypes = [C.c_void_p]
        l.mptn_destroy.restype = None
        l.mptn_last_error.restype = C.c_char_p
        l.mptn_last_error.argtypes = [C.c_void_p]
        l.mptn_load_config.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint32)]
        l.mptn_dump_config.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
        l.mptn_describe_outputs.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32] + [C.POINTER(C.c_uint32)] * 3
        l.mptn_set_shard.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        l.mptn_process.argtypes = [C.c_void_p, C.POINTER(abi.Camera), C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
        for kind in ("gbuffer", "accumulate"):
            getattr(l, f"mptn_{kind}_create").argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
            getattr(l, f"mptn_{kind}_destroy").argtypes = [C.c_void_p]
            getattr(l, f"mptn_{kind}_destroy").restype = None
            getattr(l, f"mptn_{kind}_last_error").argtypes = [C.c_void_p]
            getattr(l, f"mptn_{kind}_last_error").restype = C.c_char_p
            getattr(l, f"mptn_{kind}_load_config").argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint32)]
            getattr(l, f"mptn_{kind}_dump_config").argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
        l.mptn_gbuffer_set_resolution.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        l.mptn_gbuffer_describe_outputs.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint32)] * 3
        l.mptn_gbuffer_process.argtypes = [C.c_void_p, C.POINTER(abi.Camera), C.POINTER(abi.Camera), C.POINTER(abi.GBufferImages), C.c_void_p]
        l.mptn_accumulate_request_clear.argtypes = [C.c_void_p]
        l.mptn_accumulate_process.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(abi.AccumulateBuffers), C.c_void_p,
                                              C.POINTER(C.c_uint32)]
        l.mptn_process_accumulate.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(abi.Camera), C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p,
                                              C.POINTER(abi.AccumulateBuffers), C.c_void_p, C.POINTER(C.c_uint32)]
        if l.mpt_abi_version() != abi.MPT_ABI_VERSION:
            raise RuntimeError("libmpt.so ABI version mismatch")
        _lib = l
    return _lib
