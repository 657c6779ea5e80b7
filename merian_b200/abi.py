"""ctypes mirror of include/mpt.h (data formats only; no algorithm lives here).

Shared by the product loader (merian_b200._lib) and by the test-side oracle loader (oracle/oracle_py.py):
both sides consume the same scene tables, exactly as the reference's renderer and its Scene do.
"""
import ctypes as C

MPT_ABI_VERSION = 3

# mpt_status
MPT_OK, MPT_ERR_INVALID_ARGUMENT, MPT_ERR_CUDA, MPT_ERR_NOT_READY, MPT_ERR_UNSUPPORTED, MPT_ERR_LIMIT, MPT_ERR_IO = range(7)

# GeometryDataFlags, include/merian-shaders/scene/scene-data.slangh:15-42
GEOM_PRETRANSFORMED = 0x1
GEOM_HAS_PREV_VERTICES = 0x2
GEOM_IS_ANIMATED = 0x4
GEOM_FLIP_FACING = 0x8
GEOM_INDEX_U8 = 0x10
GEOM_INDEX_U16 = 0x20
GEOM_INDEX_U32 = 0x40
GEOM_FLAT_SHADING = 0x80
GEOM_HAS_TANGENTS = 0x100
GEOM_USE_ENV_MAP = 0x200
GEOM_TWO_SIDED = 0x400

INSTANCE_FACING_CULL_DISABLE = 0x1
INSTANCE_FLIP_FACING = 0x2
INSTANCE_FORCE_OPAQUE = 0x4

MATERIAL_DIFFUSE, MATERIAL_GLTF = 0, 1
GLTF_TRANSMISSION, GLTF_VOLUME, GLTF_CLEARCOAT, GLTF_SHEEN, GLTF_IRIDESCENCE, GLTF_ANISOTROPY = 1, 2, 4, 8, 16, 32

TEX_RGBA8_UNORM, TEX_RGBA8_SRGB, TEX_RGBA32F = 0, 1, 2
WRAP_REPEAT, WRAP_CLAMP_TO_EDGE, WRAP_MIRRORED_REPEAT = 0, 1, 2
FILTER_NEAREST, FILTER_LINEAR = 0, 1
MIPMAP_NONE, MIPMAP_NEAREST, MIPMAP_LINEAR = 0, 1, 2  # has a mip chain (cmd_generate_mipmaps) + VkSamplerMipmapMode
TEXTURE_NONE = 0xFFFF

ENV_EMPTY, ENV_LATLONG = 0, 1

RENDER_EMISSION_ON_PRIMARY = 0x1
RENDER_DEMODULATE_ALBEDO = 0x2
RENDER_OUTPUT_SUM = 0x4
RENDER_NEE = 0x8

EXTERNAL_MEMORY_OPAQUE_FD, EXTERNAL_MEMORY_CUDA_VMM_FD = 0, 1
EXTERNAL_MEMORY_DEDICATED = 0x1
EXTERNAL_SEMAPHORE_OPAQUE_FD, EXTERNAL_SEMAPHORE_TIMELINE_FD = 0, 1
COMM_UNIQUE_ID_BYTES = 128
REDUCE_ALL_RANKS, REDUCE_KEEP_SUM = 0x1, 0x2


class PackedVertex(C.Structure):  # scene-data.slangh:75-80, 24 bytes
    _pack_ = 4
    _fields_ = [("position", C.c_float * 3), ("encoded_normal", C.c_uint32), ("uv", C.c_uint16 * 2),
                ("encoded_tangent", C.c_uint32)]


class MeshDesc(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("indices", C.c_void_p), ("vertex_count", C.c_uint32),
                ("primitive_count", C.c_uint32), ("index_flag", C.c_uint32), ("opaque", C.c_uint32)]


class BlasDesc(C.Structure):
    _fields_ = [("meshes", C.c_void_p), ("mesh_count", C.c_uint32)]


class GeometryDesc(C.Structure):
    _fields_ = [("mesh", C.c_uint32), ("material_id", C.c_uint16), ("flags", C.c_uint16), ("primitive_count", C.c_uint32)]


class InstanceDesc(C.Structure):
    _fields_ = [("transform", C.c_float * 12), ("instance_custom_index", C.c_uint32), ("mask", C.c_uint32),
                ("flags", C.c_uint32), ("blas", C.c_uint32)]


class SceneDesc(C.Structure):
    _fields_ = [("meshes", C.c_void_p), ("mesh_count", C.c_uint32),
                ("blas", C.c_void_p), ("blas_count", C.c_uint32),
                ("geometries", C.c_void_p), ("geometry_count", C.c_uint32),
                ("instances", C.c_void_p), ("instance_count", C.c_uint32),
                ("instance_transforms", C.c_void_p), ("inverse_transposed_instance_transforms", C.c_void_p)]


class SrcMesh(C.Structure):  # mpt_src_mesh, include/mpt_tables.h
    _fields_ = [("positions", C.c_void_p), ("normals", C.c_void_p), ("uvs", C.c_void_p), ("tangents", C.c_void_p), ("vertex_count", C.c_uint32),
                ("indices", C.c_void_p), ("primitive_count", C.c_uint32), ("index_size", C.c_uint32), ("material_id", C.c_uint32),
                ("flags", C.c_uint32), ("instance_mask", C.c_uint32)]


class SrcMaterial(C.Structure):  # mpt_src_material, include/mpt_tables.h
    _fields_ = [("model", C.c_uint32), ("base_color_factor", C.c_float * 4), ("base_color_texture", C.c_uint32),
                ("emissive_factor", C.c_float * 3), ("metallic_factor", C.c_float), ("roughness_factor", C.c_float), ("normal_scale", C.c_float),
                ("occlusion_strength", C.c_float), ("ior", C.c_float),
                ("metallic_roughness_texture", C.c_uint32), ("normal_texture", C.c_uint32), ("occlusion_texture", C.c_uint32), ("emissive_texture", C.c_uint32),
                ("features", C.c_uint32), ("transmission_weight", C.c_float), ("volume_absorption", C.c_float * 3),
                ("clearcoat_weight", C.c_float), ("clearcoat_roughness", C.c_float), ("clearcoat_texture", C.c_uint32), ("clearcoat_roughness_texture", C.c_uint32),
                ("sheen_color", C.c_float * 3), ("sheen_roughness", C.c_float), ("sheen_color_texture", C.c_uint32), ("sheen_roughness_texture", C.c_uint32),
                ("iridescence_factor", C.c_float), ("iridescence_ior", C.c_float), ("iridescence_thickness_min", C.c_float), ("iridescence_thickness_max", C.c_float),
                ("iridescence_texture", C.c_uint32), ("iridescence_thickness_texture", C.c_uint32),
                ("anisotropy_strength", C.c_float), ("anisotropy_rotation", C.c_float), ("anisotropy_texture", C.c_uint32)]


class MaterialTableView(C.Structure):  # mpt_material_table_view
    _fields_ = [("entries", C.c_void_p), ("entry_size", C.c_uint32), ("material_count", C.c_uint32), ("types", C.c_void_p), ("type_count", C.c_uint32)]


class GltfTexture(C.Structure):  # mpt_gltf_texture
    _fields_ = [("encoded", C.c_void_p), ("encoded_size", C.c_uint64), ("srgb", C.c_uint32), ("wrap_u", C.c_uint32), ("wrap_v", C.c_uint32),
                ("filter", C.c_uint32), ("mipmap", C.c_uint32)]


class SrcNode(C.Structure):  # mpt_src_node
    _fields_ = [("transform", C.c_float * 16), ("meshes", C.c_void_p), ("mesh_count", C.c_uint32)]


class MaterialType(C.Structure):
    _fields_ = [("model", C.c_uint32), ("features", C.c_uint32)]


class Camera(C.Structure):  # camera.slang:6-24
    _fields_ = [("position", C.c_float * 3), ("target", C.c_float * 3), ("up", C.c_float * 3),
                ("U", C.c_float * 3), ("V", C.c_float * 3), ("W", C.c_float * 3),
                ("near_plane", C.c_float), ("far_plane", C.c_float), ("aspect_ratio", C.c_float),
                ("aperture_radius", C.c_float), ("jitter", C.c_float * 2)]



class GltfView(C.Structure):  # mpt_gltf_view
    _fields_ = [("meshes", C.c_void_p), ("mesh_count", C.c_uint32), ("nodes", C.c_void_p), ("node_count", C.c_uint32),
                ("materials", C.c_void_p), ("material_count", C.c_uint32), ("textures", C.c_void_p), ("texture_count", C.c_uint32),
                ("has_camera", C.c_uint32), ("camera", Camera), ("yfov_degrees", C.c_float), ("aspect", C.c_float), ("znear", C.c_float), ("zfar", C.c_float)]

class RenderParams(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("frame", C.c_uint32), ("spp", C.c_int32),
                ("max_path_length", C.c_int32), ("instance_mask", C.c_uint32), ("flags", C.c_uint32),
                ("sample_offset", C.c_uint32), ("sample_stride", C.c_uint32), ("scene_time", C.c_float)]


ACC_GBUFFER_CHECK, ACC_HISTORY_F32 = 0x1, 0x2


class AccumulateParams(C.Structure):  # push constants of accumulate.slang:80-97 that the plain path reads
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("alpha", C.c_float), ("max_history", C.c_float),
                ("hard_clamp", C.c_float), ("normal_reject_cos", C.c_float), ("z_reject_percent", C.c_float),
                ("clear", C.c_uint32), ("flags", C.c_uint32)]


class AccumulateBuffers(C.Structure):
    _fields_ = [("d_prev_out", C.c_void_p), ("d_prev_history", C.c_void_p), ("d_out", C.c_void_p), ("d_history", C.c_void_p),
                ("d_gbuf_tex0", C.c_void_p), ("d_gbuf_tex2", C.c_void_p)]


class GBufferImages(C.Structure):  # gbuffer.slang:63-82
    _fields_ = [("d_tex0", C.c_void_p), ("d_tex1", C.c_void_p), ("d_tex2", C.c_void_p), ("d_tex3", C.c_void_p), ("d_emission", C.c_void_p)]


class Ray(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("tmin", C.c_float), ("direction", C.c_float * 3), ("tmax", C.c_float)]


class Hit(C.Structure):
    _fields_ = [("instance_id", C.c_uint32), ("instance_index", C.c_uint32), ("primitive_id", C.c_uint32),
                ("geometry_index", C.c_uint32), ("barycentrics", C.c_float * 2)]


class Stats(C.Structure):
    _fields_ = [("rays_primary", C.c_uint64), ("rays_continuation", C.c_uint64), ("rays_shadow", C.c_uint64),
                ("shaded_hits", C.c_uint64), ("samples", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64), ("extend_ms", C.c_double),
                ("shade_ms", C.c_double), ("extend_launches", C.c_uint64), ("shade_launches", C.c_uint64)]


class BvhInfo(C.Structure):
    _fields_ = [("node_count", C.c_uint64), ("triangle_count", C.c_uint64), ("node_bytes", C.c_uint64),
                ("triangle_bytes", C.c_uint64), ("build_ms", C.c_double), ("max_depth", C.c_uint64)]


class HostAccelView(C.Structure):  # mpt_host_accel_view
    _fields_ = [("nodes", C.c_void_p), ("node_count", C.c_uint64), ("tris", C.c_void_p), ("triangle_count", C.c_uint64),
                ("instances", C.c_void_p), ("instance_count", C.c_uint64), ("tlas_prims", C.c_void_p), ("tlas_prim_count", C.c_uint64),
                ("prim_to_tri", C.c_void_p), ("prim_count", C.c_uint64), ("root", C.c_uint32), ("two_level", C.c_uint32),
                ("max_blas_depth", C.c_uint32), ("tlas_depth", C.c_uint32), ("build_ms", C.c_double)]


class BsdfCheckParams(C.Structure):
    _fields_ = [("bsdf", C.c_uint32), ("wi", C.c_float * 3), ("n_samples", C.c_uint32), ("seed", C.c_uint32),
                ("albedo", C.c_float * 3), ("alpha", C.c_float), ("f0", C.c_float * 3), ("ior_f0", C.c_float),
                ("alpha_bitangent", C.c_float), ("iridescence", C.c_float)]


# numpy dtypes of the wire formats
import numpy as np  # noqa: E402

PACKED_VERTEX_DTYPE = np.dtype([("position", "<f4", 3), ("encoded_normal", "<u4"), ("uv", "<u2", 2),
                                ("encoded_tangent", "<u4")])
assert PACKED_VERTEX_DTYPE.itemsize == 24 and C.sizeof(PackedVertex) == 24
# PackedHitInfo, include/merian-shaders/gbuffer.slang:40-61 (one R32G32B32A32_UINT texel of GBuffer tex1)
PACKED_HIT_DTYPE = np.dtype([("instance_id", "<u4"), ("primitive_id", "<u4"), ("geometry_index", "<u2"),
                             ("instance_index", "<u2"), ("barycentrics", "<f2", 2)])
assert PACKED_HIT_DTYPE.itemsize == 16
RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("tmin", "<f4"), ("direction", "<f4", 3), ("tmax", "<f4")])
HIT_DTYPE = np.dtype([("instance_id", "<u4"), ("instance_index", "<u4"), ("primitive_id", "<u4"),
                      ("geometry_index", "<u4"), ("barycentrics", "<f4", 2)])
assert RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 24
