/* mpt.h — C-ABI of libmpt.so, the B200-native path tracer behind merian-graph's `pt` node.
 *
 * This is the drop-in boundary. The reference has no FFI for this path: its
 * `Render (Path-traced)` node records ONE `vkCmdTraceRaysKHR` per iteration
 * (src/merian-graph/nodes/render_pt/render_pt.cpp:94-124) over tables owned by `merian::Scene`.
 * The entry points below are what a merian node adapter binds instead; each cites
 * the reference interface it replaces. Plain pointers and sizes only — no C++/torch types, no
 * exceptions; every call returns an `mpt_status` and `mpt_last_error()` explains failures.
 *
 * All pointers named `d_*` are device pointers on the context's device, everything else is host
 * memory. `stream` is a `cudaStream_t` passed as `void*` (0 = legacy default stream).
 */
#ifndef MPT_H
#define MPT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define MPT_API __declspec(dllexport)
#else
#define MPT_API __attribute__((visibility("default")))
#endif

#define MPT_ABI_VERSION 3u /* 2: + gbuffer, accumulate, PFM/HDR I/O, iterations in flight, NEE, bvh_info.max_depth
                              3: + external memory / semaphore import, mpt_comm_* / mpt_reduce_*, host accel export, ray-queue read-back */

typedef enum mpt_status {
    MPT_OK = 0,
    MPT_ERR_INVALID_ARGUMENT = 1,
    MPT_ERR_CUDA = 2,
    MPT_ERR_NOT_READY = 3, /* scene / materials not uploaded yet (render_pt.cpp:98: !scene->is_ready()) */
    MPT_ERR_UNSUPPORTED = 4,
    MPT_ERR_LIMIT = 5,
    MPT_ERR_IO = 6 /* file missing / malformed / short (image_io.cpp throws std::runtime_error for these) */
} mpt_status;

typedef struct mpt_context mpt_context;

/* ---- data formats shared with the reference ------------------------------------------------ */

/* GeometryDataFlags, include/merian-shaders/scene/scene-data.slangh:15-42 (same values). */
enum {
    MPT_GEOM_PRETRANSFORMED = 0x1,
    MPT_GEOM_HAS_PREV_VERTICES = 0x2,
    MPT_GEOM_IS_ANIMATED = 0x4,
    MPT_GEOM_FLIP_FACING = 0x8,
    MPT_GEOM_INDEX_U8 = 0x10,
    MPT_GEOM_INDEX_U16 = 0x20,
    MPT_GEOM_INDEX_U32 = 0x40,
    MPT_GEOM_FLAT_SHADING = 0x80,
    MPT_GEOM_HAS_TANGENTS = 0x100,
    MPT_GEOM_USE_ENV_MAP = 0x200,
    MPT_GEOM_TWO_SIDED = 0x400
};

/* VkGeometryInstanceFlagBitsKHR as set by Scene::build_tlas (src/merian-shaders/scene/scene.cpp:1836-1847). */
enum {
    MPT_INSTANCE_FACING_CULL_DISABLE = 0x1, /* eTriangleFacingCullDisable (TwoSided groups) */
    MPT_INSTANCE_FLIP_FACING = 0x2,         /* eTriangleFlipFacing                          */
    MPT_INSTANCE_FORCE_OPAQUE = 0x4         /* eForceOpaque (group.all_opaque)              */
};

/* PackedVertexData, scene-data.slangh:75-80: 24 bytes, 4-byte aligned.
 *   float3 position; uint32 encoded_normal (oct 2x snorm16); half2 uv; uint32 encoded_tangent. */
typedef struct mpt_packed_vertex {
    float position[3];
    uint32_t encoded_normal;
    uint16_t uv[2]; /* IEEE binary16 bits */
    uint32_t encoded_tangent;
} mpt_packed_vertex;

/* One mesh = one BLAS geometry: vertex + index arrays as Scene::upload_meshes stores them
 * (scene.cpp:1377-1670), i.e. AFTER the transform_vertex pass (transform_vertex.slang:12-40). */
typedef struct mpt_mesh_desc {
    const mpt_packed_vertex* vertices;
    const void* indices;      /* u8/u16/u32 triples, or NULL for sequential (IndexType::None) */
    uint32_t vertex_count;
    uint32_t primitive_count;
    uint32_t index_flag;      /* 0 | MPT_GEOM_INDEX_U8 | _U16 | _U32 */
    uint32_t opaque;          /* VK_GEOMETRY_OPAQUE_BIT: 0 => candidates run the alpha test (scene.slang:500-519) */
} mpt_mesh_desc;

/* One mesh group = one BLAS (scene.hpp:401-436). geometry_index of a hit = position in `meshes`. */
typedef struct mpt_blas_desc {
    const uint32_t* meshes;
    uint32_t mesh_count;
} mpt_blas_desc;

/* GeometryData, scene-data.slangh:119-129, in GeometryID order (scene.cpp:1236-1291), with the three
 * buffer device addresses replaced by a mesh index. */
typedef struct mpt_geometry_desc {
    uint32_t mesh;
    uint16_t material_id;
    uint16_t flags; /* MPT_GEOM_* */
    uint32_t primitive_count;
} mpt_geometry_desc;

/* VkAccelerationStructureInstanceKHR as filled by Scene::build_tlas (scene.cpp:1808-1857), in TLAS
 * order: InstanceIndex() = position in the array, InstanceID() = instance_custom_index. */
typedef struct mpt_instance_desc {
    float transform[12]; /* row-major 3x4 object->world (VkTransformMatrixKHR) */
    uint32_t instance_custom_index;
    uint32_t mask;  /* 8 bit */
    uint32_t flags; /* MPT_INSTANCE_* */
    uint32_t blas;
} mpt_instance_desc;

typedef struct mpt_scene_desc {
    const mpt_mesh_desc* meshes;
    uint32_t mesh_count;
    const mpt_blas_desc* blas;
    uint32_t blas_count;
    const mpt_geometry_desc* geometries;
    uint32_t geometry_count;
    const mpt_instance_desc* instances;
    uint32_t instance_count;
    /* Scene::instance_transforms / inverse_transposed_instance_transforms (scene.slang:33-38), one
     * row-major float4x4 per TLAS instance; may be NULL when every geometry is Pretransformed. */
    const float* instance_transforms;
    const float* inverse_transposed_instance_transforms;
} mpt_scene_desc;

/* Material model dispatch (material-system.slang:51-72): material_model_type_id -> model. The ids are
 * assigned at run time in registration order (material_system.cpp:124-139), so the adapter passes the table. */
enum { MPT_MATERIAL_DIFFUSE = 0, MPT_MATERIAL_GLTF = 1 };
enum { /* GltfMaterial<TRANSMISSION, VOLUME, CLEARCOAT, SHEEN, IRIDESCENCE, ANISOTROPY>, gltf-material.slang:64-69 */
    MPT_GLTF_TRANSMISSION = 1, MPT_GLTF_VOLUME = 2, MPT_GLTF_CLEARCOAT = 4,
    MPT_GLTF_SHEEN = 8, MPT_GLTF_IRIDESCENCE = 16, MPT_GLTF_ANISOTROPY = 32
};
typedef struct mpt_material_type {
    uint32_t model;    /* MPT_MATERIAL_* */
    uint32_t features; /* MPT_GLTF_* bits, 0 for diffuse */
} mpt_material_type;

/* Texture formats accepted by mpt_texture_upload (texture_manager.cpp:207-243 creates RGBA8 sRGB/UNORM;
 * the env map is RGBA32F, scene.cpp:780-795). */
enum { MPT_TEX_RGBA8_UNORM = 0, MPT_TEX_RGBA8_SRGB = 1, MPT_TEX_RGBA32F = 2 };
enum { MPT_WRAP_REPEAT = 0, MPT_WRAP_CLAMP_TO_EDGE = 1, MPT_WRAP_MIRRORED_REPEAT = 2 };
enum { MPT_FILTER_NEAREST = 0, MPT_FILTER_LINEAR = 1 };
#define MPT_TEXTURE_NONE 0xFFFFu

/* Camera, include/merian-shaders/scene/camera.slang:6-24, as written by Camera::write_to
 * (src/merian/utils/camera/camera.cpp:349-368). */
typedef struct mpt_camera {
    float position[3];
    float target[3];
    float up[3];
    float U[3];
    float V[3];
    float W[3];
    float near_plane;
    float far_plane;
    float aspect_ratio;
    float aperture_radius; /* 0 = pinhole; > 0 = thin lens (camera.slang:73-92; the hint merian_hint_enable_thin_lens follows it, scene.cpp:2019-2021) */
    float jitter[2];
} mpt_camera;

enum { MPT_ENV_EMPTY = 0, MPT_ENV_LATLONG = 1 };

/* Render flags: the compile-time constants of render_pt.slang:11-16 become run-time parameters. */
enum {
    MPT_RENDER_EMISSION_ON_PRIMARY = 0x1, /* merian_render_emission_on_primary */
    MPT_RENDER_DEMODULATE_ALBEDO = 0x2,   /* merian_render_demodulate_albedo (needs d_albedo) */
    MPT_RENDER_OUTPUT_SUM = 0x4,          /* write sums over the samples instead of means (multi-GPU reduce) */
    MPT_RENDER_NEE = 0x8                  /* next event estimation towards the lat-long environment: alias-table light samples + shadow rays,
                                             balance-heuristic MIS against the BSDF-sampled rays. Not in the reference's pt (it has no NEE):
                                             same expectation, lower variance; off by default and in every parity / bench number */
};

typedef struct mpt_render_params {
    uint32_t width, height;
    uint32_t frame;            /* Scene::frame (scene.slang:48), = info.get_total_iteration() */
    int32_t spp;               /* "samples per pixel"  (render_pt.cpp:151) */
    int32_t max_path_length;   /* "max path length"    (render_pt.cpp:154) */
    uint32_t instance_mask;    /* properties "0".."7"  (render_pt.cpp:127-131) */
    uint32_t flags;            /* MPT_RENDER_* */
    /* Sample sharding: this call renders sample indices s = sample_offset + i*sample_stride,
     * i in [0, spp). RNG seed is pcg4d16(px, py, frame, s) (render_pt.slang:114), so any partition of
     * the sample indices over GPUs yields the same set of paths. Single GPU: offset 0, stride 1. */
    uint32_t sample_offset;
    uint32_t sample_stride;
    float scene_time;          /* Scene::time (unused by the supported material models) */
} mpt_render_params;

/* Ray / hit records of mpt_trace. Ray = HLSL RayDesc (origin, tmin, direction, tmax). */
typedef struct mpt_ray {
    float origin[3];
    float tmin;
    float direction[3];
    float tmax;
} mpt_ray;

/* Hit = merian::Hit (hit.slang:25-66); miss = all ids 0xFFFFFFFF, barycentrics 0. */
typedef struct mpt_hit {
    uint32_t instance_id;
    uint32_t instance_index;
    uint32_t primitive_id;
    uint32_t geometry_index;
    float barycentrics[2];
} mpt_hit;

/* Work counters accumulated over calls since the last reset (bench.py: Mrays/s, roofline bytes). */
typedef struct mpt_stats {
    uint64_t rays_primary;      /* closest-hit queries issued for camera rays          */
    uint64_t rays_continuation; /* closest-hit queries issued for BSDF-sampled rays    */
    uint64_t rays_shadow;       /* any-hit (visibility) queries issued by NEE          */
    uint64_t shaded_hits;       /* shading points with geometry (get_shading_point)    */
    uint64_t samples;           /* paths started (pixels * spp)                        */
    uint64_t kernel_launches;   /* libmpt kernels launched                             */
    uint64_t nodes_visited;     /* BVH8 nodes popped   (only counted when profiling counters are on) */
    uint64_t tris_tested;       /* triangles tested    (only counted when profiling counters are on) */
    double extend_ms;           /* device time of the extend (traversal) kernels, CUDA events (when timing is on) */
    double shade_ms;            /* device time of the shade kernels                                            */
    uint64_t extend_launches;
    uint64_t shade_launches;
} mpt_stats;

/* ---- entry points -------------------------------------------------------------------------- */

MPT_API uint32_t mpt_abi_version(void);

/* Context per device; replaces RenderPT::initialize (render_pt.cpp:21-25). Fails (no CPU fallback)
 * when the device is not sm_100. */
MPT_API mpt_status mpt_create(int device, mpt_context** out_ctx);
MPT_API void mpt_destroy(mpt_context* ctx);
MPT_API const char* mpt_last_error(const mpt_context* ctx); /* ctx may be NULL: last create error */

/* Replaces Scene::upload_meshes / upload_geometry_data / build_blas / build_tlas
 * (scene.cpp:1377-1670, 1230-1311, 1672-1892): copies the tables into HBM and builds one compressed
 * 8-wide BVH per BLAS plus one over the TLAS instances. */
MPT_API mpt_status mpt_scene_upload(mpt_context* ctx, const mpt_scene_desc* scene);

/* Replaces MaterialSystem::update (material_system.cpp:234-273): `entries` is the reference's host
 * buffer, entry = MaterialHeader (4 B) + payload, entry_size bytes each (material_system.cpp:140-195). */
MPT_API mpt_status mpt_materials_upload(mpt_context* ctx, const void* entries, uint32_t entry_size,
                                        uint32_t material_count, const mpt_material_type* types,
                                        uint32_t type_count);

/* MaterialSystem hints, compile-time constants of the reference's composition (material_system.hpp:83-97,
 * material.slang:20-23, material-system.slang:7): alpha-test threshold (default 0.5), minimum roughness (0.0316) and
 * whether normal-mapped shading normals are clamped (Keller) to the visible hemisphere (true). */
MPT_API mpt_status mpt_materials_set_hints(mpt_context* ctx, float alpha_test_threshold, float min_roughness,
                                           uint32_t clamp_normals);

/* Replaces TextureManager texture creation (texture_manager.cpp:207-243). */
MPT_API mpt_status mpt_texture_upload(mpt_context* ctx, uint32_t texture_id, uint32_t width,
                                      uint32_t height, uint32_t format, uint32_t wrap_u,
                                      uint32_t wrap_v, uint32_t filter, const void* pixels);
/* The same with a mip chain: mipmap_mode != MPT_MIPMAP_NONE allocates floor(log2(max(w, h))) + 1 levels and fills them the way
 * create_texture_from_rgba8(..., generate_mipmaps) does (resource_allocator.cpp:298-313: cmd_generate_mipmaps = a linear blit per level,
 * include/merian/vk/utils/blits.hpp:101-168). Which textures get one is the loader's decision (GLTFScene: the glTF sampler's *_MIPMAP_*
 * minFilter, gltf_scene.cpp:212-248, 583-593). Only the GBuffer node samples with screen-space gradients (GradLODSampler,
 * gbuffer_rt.slang:41-66); `pt` reads LOD 0 of every texture (LOD0Sampler, textures.slang:338-366). */
enum { MPT_MIPMAP_NONE = 0, MPT_MIPMAP_NEAREST = 1, MPT_MIPMAP_LINEAR = 2 }; /* VkSamplerMipmapMode + "has a chain" */
MPT_API mpt_status mpt_texture_upload_ex(mpt_context* ctx, uint32_t texture_id, uint32_t width, uint32_t height, uint32_t format, uint32_t wrap_u,
                                         uint32_t wrap_v, uint32_t filter, uint32_t mipmap_mode, const void* pixels);
/* Host-only (no device needed): one level of the chain mpt_texture_upload_ex would build, for tests. out may be NULL (sizes only). */
MPT_API mpt_status mpt_mip_level(const void* level0, uint32_t width, uint32_t height, uint32_t format, uint32_t level, void* out, uint64_t capacity_bytes,
                                 uint32_t* level_width, uint32_t* level_height);
/* Texels that are already in device memory — inside merian: the VkImage copied to an exportable VkBuffer and imported with
 * mpt_import_external_memory_fd, which replaces the read-back of texture_manager.cpp:207-243 / scene.cpp:780-795 through the host.
 * row_pitch_bytes: 0 = tightly packed. Device-to-device for MPT_MIPMAP_NONE; a mip chain is filtered on the host. */
MPT_API mpt_status mpt_texture_upload_device(mpt_context* ctx, uint32_t texture_id, uint32_t width, uint32_t height, uint32_t format, uint32_t wrap_u,
                                             uint32_t wrap_v, uint32_t filter, uint32_t mipmap_mode, const void* d_pixels, uint64_t row_pitch_bytes,
                                             void* stream);

/* Replaces LatLongEnvMap::write_to (env_map.hpp:112-118). `to_local` row-major 4x4; pixels RGBA32F,
 * repeat addressing, bilinear (scene.cpp:786-789). type MPT_ENV_EMPTY ignores the rest. The alias table of
 * MPT_RENDER_NEE is built from these pixels on the first render that asks for it. */
MPT_API mpt_status mpt_env_set(mpt_context* ctx, uint32_t type, const float* to_local,
                               float intensity, uint32_t width, uint32_t height,
                               const float* rgba32f);
/* The lat-long texels already in device memory (see mpt_texture_upload_device). */
MPT_API mpt_status mpt_env_set_device(mpt_context* ctx, uint32_t type, const float* to_local, float intensity, uint32_t width, uint32_t height,
                                      const void* d_rgba32f);

/* Scene::frame (scene.slang:48) for the entry points that take no mpt_render_params (mpt_primary_visibility, mpt_gbuffer): a thin-lens
 * camera draws its lens sample from (pixel, frame) (Scene::generate_camera_ray, scene.slang:91-101), so every pass of one iteration must
 * see the same frame number. mpt_render* set it from params->frame. Pinhole cameras do not read it. */
MPT_API mpt_status mpt_set_scene_frame(mpt_context* ctx, uint32_t frame);

/* Scene::prev_instance_transforms (scene.slang:38, 136-139): the object -> world matrices of the PREVIOUS frame, instance_count x 16
 * floats in the layout of mpt_scene_desc::instance_transforms (row-major 4x4). Read by mpt_gbuffer for geometry flagged
 * MPT_GEOM_IS_ANIMATED and not MPT_GEOM_PRETRANSFORMED (Scene::to_prev_world_position, scene.slang:261-270): motion vector and
 * delta_z then follow the moving instance. NULL = drop them (every instance static, the state after mpt_scene_upload).
 * instance_count must equal the uploaded scene's. Previous-frame VERTICES (MPT_GEOM_HAS_PREV_VERTICES: morphed / skinned meshes,
 * transform_prev_vertex.slang) are not supported: mpt_scene_upload refuses that flag. */
MPT_API mpt_status mpt_scene_set_prev_transforms(mpt_context* ctx, const float* prev_instance_transforms, uint32_t instance_count);

/* Replaces the gbuffer_rt raygen for the hit texture (gbuffer_rt.slang:18-31, gbuffer.slang:40-61):
 * writes one PackedHitInfo (uint4) per pixel into d_tex1 (row-major, width*height*16 bytes). */
MPT_API mpt_status mpt_primary_visibility(mpt_context* ctx, const mpt_camera* camera, uint32_t width,
                                          uint32_t height, uint32_t instance_mask, void* d_tex1,
                                          void* stream);

/* The whole `GBuffer (Raytraced)` node: replaces the gbuffer_rt raygen
 * (include/merian-graph/nodes/gbuffer_rt/gbuffer_rt.slang:18-94) with the texel layouts of include/merian-shaders/gbuffer.slang:63-82,152-172.
 * Any image may be NULL (not written). prev_camera NULL = camera (static view: motion vector = -jitter, delta_z = 0).
 * Moving instances: mpt_scene_set_prev_transforms. Material textures are sampled with the ray-differential UV gradients over the
 * mip chains of mpt_texture_upload_ex (textures uploaded without a chain: LOD 0). */
typedef struct mpt_gbuffer_images {
    void* d_tex0;     /* uint2:  .x = encode_normal(material shading normal), .y = half2 grad_z          */
    void* d_tex1;     /* uint4:  PackedHitInfo (same as mpt_primary_visibility)                           */
    void* d_tex2;     /* float4: .xy = pixel-space motion vector, .z = min(linear_z, 65504), .w = delta_z */
    void* d_tex3;     /* float4: .rgb = BSDF::get_albedo(wi) (miss: environment radiance), .a = 1         */
    void* d_emission; /* float4: .rgb = emission of the primary vertex (the node's `emission` output)     */
} mpt_gbuffer_images;
MPT_API mpt_status mpt_gbuffer(mpt_context* ctx, const mpt_camera* camera, const mpt_camera* prev_camera,
                               uint32_t width, uint32_t height, uint32_t instance_mask,
                               const mpt_gbuffer_images* images, void* stream);

/* Closest-hit queries for arbitrary rays (Scene::trace_ray, scene.slang:62-76) with
 * RAY_FLAG_CULL_BACK_FACING_TRIANGLES (acceleration-structure.slang:64). d_rays: mpt_ray[n], d_hits: mpt_hit[n]. */
MPT_API mpt_status mpt_trace(mpt_context* ctx, const void* d_rays, uint64_t n, uint32_t instance_mask,
                             void* d_hits, void* stream);

/* Replaces RenderPT::process (render_pt.cpp:94-124) = the raygen shader render_pt.slang:92-130.
 * d_tex1: GBuffer hit texture (PackedHitInfo per pixel) or NULL to trace primary rays internally.
 * d_albedo: GBuffer tex3 (float4 per pixel) or NULL; d_irradiance: float4 per pixel (rgb, second moment). */
MPT_API mpt_status mpt_render(mpt_context* ctx, const mpt_camera* camera,
                              const mpt_render_params* params, const void* d_tex1,
                              const void* d_albedo, void* d_irradiance, void* stream);

/* Same, end to end from HOST memory (pinned or pageable): copies tex1 (if given) host->device, renders,
 * copies the irradiance image device->host, synchronises the stream. */
MPT_API mpt_status mpt_render_host(mpt_context* ctx, const mpt_camera* camera,
                                   const mpt_render_params* params, const void* h_tex1,
                                   const void* h_albedo, void* h_irradiance, void* stream);

/* The same with several iterations in flight, the way merian's graph runs (2 by default, examples/gltf.json:39; ring index =
 * info.get_in_flight_index(), node_process_info.hpp:42-46): _begin enqueues H2D, render and D2H of one iteration and returns;
 * _end(i) blocks until the irradiance of slot i is in its host buffer. Copies of neighbouring iterations overlap the render.
 * Host buffers must stay valid (and should be pinned) until _end. in_flight_index < 4. */
MPT_API mpt_status mpt_render_host_begin(mpt_context* ctx, const mpt_camera* camera, const mpt_render_params* params,
                                         const void* h_tex1, const void* h_albedo, void* h_irradiance,
                                         uint32_t in_flight_index);
MPT_API mpt_status mpt_render_host_end(mpt_context* ctx, uint32_t in_flight_index);

/* ---- Accumulate, plain mode ----------------------------------------------
 * Replaces the accumulate pass of the `Accumulate` node (include/merian-graph/nodes/accumulate/accumulate.slang:103-255,
 * src/merian-graph/nodes/accumulate/accumulate.cpp:153-167) for a signal that is reused at the SAME pixel
 * ("use motion vectors" off), firefly filter and adaptive alpha off:
 *   src: NaN/Inf -> 0, luminance clamped to `hard_clamp`;  clear or alpha == 0: out = src, history = 1;
 *   else history' = min(history + 1, max_history), out = lerp(prev_out, src, max(1 - alpha, 1 / history')).
 * History texels are the reference's (uint2: .x = GBuffer enc_normal, .y = f16 linear_z | f16 length << 16), so the f16
 * counter saturates at 2048 exactly like the reference; MPT_ACC_HISTORY_F32 stores the length as f32 bits in .y
 * instead (no depth, no cap). MPT_ACC_GBUFFER_CHECK = "gbuffer check: always" for the same pixel
 * (reprojection_valid, reprojection.slang:56-71) without the stochastic neighbourhood search; without the flag the
 * check mode is "never". prev and out may alias (in-place accumulation). */
enum { MPT_ACC_GBUFFER_CHECK = 0x1, MPT_ACC_HISTORY_F32 = 0x2 };
typedef struct mpt_accumulate_params {
    uint32_t width, height;
    float alpha;             /* "alpha", pc.accum_alpha (accumulate.cpp:175) */
    float max_history;       /* "max history"; INFINITY = "inf history" (accumulate.cpp:178-183) */
    float hard_clamp;        /* "hard clamp", pc.firefly_hard_clamp; INFINITY = off (accumulate.cpp:229-233) */
    float normal_reject_cos; /* pc.normal_reject_cos  (only with MPT_ACC_GBUFFER_CHECK) */
    float z_reject_percent;  /* pc.z_reject_percent   (only with MPT_ACC_GBUFFER_CHECK) */
    uint32_t clear;          /* pc.clear: first iteration or a clear request (accumulate.cpp:154-160) */
    uint32_t flags;          /* MPT_ACC_* */
} mpt_accumulate_params;
typedef struct mpt_accumulate_buffers {
    const void* d_prev_out;     /* float4 per pixel (in_prev_out);  may be NULL when clear */
    const void* d_prev_history; /* uint2 per pixel  (in_prev_history); may be NULL when clear */
    void* d_out;                /* float4 per pixel (out_out) */
    void* d_history;            /* uint2 per pixel  (out_history) */
    const void* d_gbuf_tex0;    /* GBuffer tex0, uint2 per pixel, or NULL (history normal = 0) */
    const void* d_gbuf_tex2;    /* GBuffer tex2, float4 per pixel, or NULL (history depth = 0) */
} mpt_accumulate_buffers;
/* d_src: float4 per pixel (rgb + second moment in a), e.g. the irradiance image of mpt_render. */
MPT_API mpt_status mpt_accumulate(mpt_context* ctx, const mpt_accumulate_params* params, const void* d_src,
                                  const mpt_accumulate_buffers* buffers, void* stream);
/* mpt_render with the accumulate step fused into the resolve kernel: the frame's irradiance image is never written
 * to or re-read from HBM (saves 2 x 16 B per pixel per iteration). Same arguments as mpt_render + mpt_accumulate. */
MPT_API mpt_status mpt_render_accumulate(mpt_context* ctx, const mpt_camera* camera, const mpt_render_params* params,
                                         const void* d_tex1, const void* d_albedo,
                                         const mpt_accumulate_params* acc_params,
                                         const mpt_accumulate_buffers* buffers, void* stream);

/* ---- PFM in / out -------------------------------------------------------
 * The float wire format at both ends of the path: lat-long env maps come in through image_load_f32
 * (src/merian-shaders/scene/scene.cpp:780-795) and the `Image Write` node's PFM leg goes out through image_save_f32
 * (src/merian-graph/nodes/image_write/image_write.cpp:23-30). Format per src/merian/io/image_io.cpp:53-168: "PF" (rgb) or
 * "Pf" (grey), "width height", scale (negative = little endian), ONE whitespace byte, bottom-up rows of f32. Images in
 * memory are top-down. Host-only calls take no context; their errors are read with mpt_last_error(NULL). */
MPT_API mpt_status mpt_pfm_info(const char* path, uint32_t* width, uint32_t* height, uint32_t* channels);
/* desired_channels 0 = as in the file, else 1..4 (grey replicated, alpha = 1); out holds width*height*channels floats. */
MPT_API mpt_status mpt_pfm_load(const char* path, uint32_t desired_channels, float* out, uint64_t capacity_floats);
/* channels 1, 3 or 4; RGBA is written as RGB (PFM has no alpha); little endian, scale -1. */
MPT_API mpt_status mpt_pfm_save(const char* path, const float* data, uint32_t width, uint32_t height, uint32_t channels);
/* Radiance RGBE (.hdr), decoded like stb_image's stbi_loadf which image_load_f32 calls (image_io.cpp:226-240): flat or new-style RLE
 * scanlines, component = mantissa * 2^(e - 136). desired_channels 3 or 4 (alpha = 1). out == NULL: only width / height are returned. */
MPT_API mpt_status mpt_hdr_load(const char* path, uint32_t desired_channels, uint32_t* width, uint32_t* height, float* out,
                                uint64_t capacity_floats);
/* Scene::process_pending_env_load for a lat-long .pfm or .hdr (scene.cpp:775-800): load as RGBA32F and mpt_env_set. */
MPT_API mpt_status mpt_env_load(mpt_context* ctx, const char* path, const float* to_local, float intensity);
/* The PFM leg of the `Image Write` node for a float4 device image (e.g. the accumulate output): D2H + mpt_pfm_save. */
MPT_API mpt_status mpt_image_write(mpt_context* ctx, const char* path, const void* d_rgba32f, uint32_t width,
                                   uint32_t height, void* stream);

/* Counters. profile_flags: bit0 = count BVH nodes/triangles per ray (slower kernels), bit1 = CUDA-event
 * timing of every extend/shade launch. */
MPT_API mpt_status mpt_set_profiling(mpt_context* ctx, uint32_t profile_flags);
MPT_API mpt_status mpt_get_stats(mpt_context* ctx, mpt_stats* out);
MPT_API mpt_status mpt_reset_stats(mpt_context* ctx);

/* Measurement hook (bench.py, tests): copies to the host the ray queue that the LAST extend launch of the last rendered frame consumed,
 * i.e. the continuation rays of bounce = max_path_length - 1 of that frame's last pass (earlier queues have been overwritten by then).
 * h_leaves (nullable) receives, per ray, the triangle record the ray starts on (the device does not test it; undefined for NEE frames).
 * *count = rays in that queue; at most `capacity` are copied. SURVEY.md §8(d): the oracle counts nodes / triangles per ray on these rays. */
MPT_API mpt_status mpt_debug_ray_queue(mpt_context* ctx, uint32_t bounce, mpt_ray* h_rays, uint32_t* h_leaves, uint64_t capacity, uint64_t* count);

/* Test hook: the per-ray setup of the traversal for n directions (3 floats each). h_shear: Sx, Sy, Sz of the watertight triangle test
 * (IEEE quotients d[kx]/d[kz], d[ky]/d[kz], 1/d[kz]: the device computes them with one shared reciprocal and must match a host division bit
 * for bit); h_k_perm: kx | ky << 2 | kz << 4, then the child visiting-order word; h_idir: the inverse direction as the single-level path
 * computes it (IEEE) followed by the 1-ulp form used when a ray enters an instance (6 floats per ray). */
MPT_API mpt_status mpt_debug_ray_setup(mpt_context* ctx, const float* h_dirs, uint64_t n, float* h_shear, uint32_t* h_k_perm, float* h_idir);

/* Introspection used by tests and DESIGN.md figures. */
typedef struct mpt_bvh_info {
    uint64_t node_count;     /* 80-byte BVH8 nodes over all BLAS + TLAS */
    uint64_t triangle_count; /* 48-byte triangle records */
    uint64_t node_bytes;
    uint64_t triangle_bytes;
    double build_ms;
    uint64_t max_depth;      /* wide levels on the longest path, deepest BLAS + TLAS (bounds the traversal stack) */
} mpt_bvh_info;
MPT_API mpt_status mpt_bvh_get_info(mpt_context* ctx, mpt_bvh_info* out);
/* Host-only: runs the BVH8 builder over `n` boxes (6 floats each: lo.xyz, hi.xyz) and reports its shape; prim_order (n entries, may be
 * NULL) receives the leaf order. No device needed. */
MPT_API mpt_status mpt_bvh8_build_info(const float* boxes, uint32_t n, mpt_bvh_info* out, uint32_t* prim_order);

/* ---- Vulkan <-> CUDA interop (north_star; SURVEY.md §8(b) "cross-API sync options" (2)) ------------------------------------------
 * The merian node shares its output with the rest of the graph without a host round trip: it allocates an exportable VkBuffer
 * (VkExportMemoryAllocateInfo, VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT; vkGetMemoryFdKHR) and two exportable TIMELINE semaphores
 * (VkExportSemaphoreCreateInfo, VK_EXTERNAL_SEMAPHORE_HANDLE_TYPE_OPAQUE_FD_BIT; vkGetSemaphoreFdKHR), imports them once, and per
 * iteration replaces `submission.sync_to_cpu_and_back` (include/merian/vk/command/submission.hpp:101-114) by: raw queue submit that
 * signals S_vk = n (include/merian/vk/command/queue.hpp:29-30) -> mpt_wait_external_semaphore(S_vk, n, stream) -> mpt_render(..., d_ptr,
 * stream) -> mpt_signal_external_semaphore(S_cu, n, stream) -> the graph's next submit waits S_cu = n. merian's own TimelineSemaphore
 * cannot be used: its handle type has a private, non-exportable constructor (include/merian/vk/sync/semaphore_timeline.hpp:10-29), so the
 * node owns its semaphores; the device extensions VK_KHR_external_memory_fd + VK_KHR_external_semaphore_fd are requested through
 * DeviceSupportInfo::check (include/merian/vk/extension/extension.hpp:70-78). A successful import takes ownership of the descriptor
 * (do not close it); a failed one leaves it with the caller. INTEGRATION.md §3 has the Vulkan side line by line. */
enum {
    MPT_EXTERNAL_MEMORY_OPAQUE_FD = 0,   /* vkGetMemoryFdKHR, VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT */
    MPT_EXTERNAL_MEMORY_CUDA_VMM_FD = 1  /* cuMemExportToShareableHandle(CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR): a CUDA-only host; size must
                                            be a multiple of the allocation granularity */
};
enum { MPT_EXTERNAL_MEMORY_DEDICATED = 0x1 /* the VkDeviceMemory is a dedicated allocation (VkMemoryDedicatedAllocateInfo) */ };
enum {
    MPT_EXTERNAL_SEMAPHORE_OPAQUE_FD = 0,  /* binary VkSemaphore exported as OPAQUE_FD */
    MPT_EXTERNAL_SEMAPHORE_TIMELINE_FD = 1 /* VK_SEMAPHORE_TYPE_TIMELINE exported as OPAQUE_FD */
};
typedef struct mpt_external_memory mpt_external_memory;
typedef struct mpt_external_semaphore mpt_external_semaphore;
MPT_API mpt_status mpt_import_external_memory_fd(mpt_context* ctx, int fd, uint64_t size, uint32_t handle_type, uint32_t flags,
                                                 mpt_external_memory** out, void** d_ptr);
MPT_API mpt_status mpt_release_external_memory(mpt_context* ctx, mpt_external_memory* memory);
MPT_API mpt_status mpt_import_external_semaphore_fd(mpt_context* ctx, int fd, uint32_t handle_type, mpt_external_semaphore** out);
/* `value`: the timeline value to wait for / to signal (ignored for binary semaphores). Enqueued on `stream`. */
MPT_API mpt_status mpt_wait_external_semaphore(mpt_context* ctx, mpt_external_semaphore* semaphore, uint64_t value, void* stream);
MPT_API mpt_status mpt_signal_external_semaphore(mpt_context* ctx, mpt_external_semaphore* semaphore, uint64_t value, void* stream);
MPT_API mpt_status mpt_release_external_semaphore(mpt_context* ctx, mpt_external_semaphore* semaphore);

/* ---- multi-GPU: the one exchange step of the path -------------------------------------------------------------
 * A frame is sharded by sample index (mpt_render_params.sample_offset / sample_stride + MPT_RENDER_OUTPUT_SUM): every GPU renders its
 * share of the samples of ALL pixels into a private sum buffer; the buffers are combined with ONE NCCL reduce over NVLink per frame and
 * normalised once. The reference has no multi-device path (render_pt.slang:113-122 loops over a pixel's samples in one thread); these
 * entry points are what lets a C++ host (one merian process per GPU) shard the node without Python.
 *
 * mpt_comm owns (or adopts) one ncclComm_t and a side stream. mpt_reduce_begin enqueues, BEHIND the work already submitted to
 * `render_stream`, on the side stream:  reduce(sum) -> divide by spp_total (receiving ranks) -> optional copy to the pinned host buffer
 * h_out. The caller may submit the next frame to render_stream right away (it overlaps the reduce) as long as it renders into a different
 * buffer. Up to 4 reduces may be in flight, one per `slot` (like in_flight_index of mpt_render_host_begin; the side stream runs them in
 * issue order). mpt_reduce_end makes a stream wait for a slot's result — e.g. the render stream, before the frame that reuses that slot's
 * sum buffer; mpt_reduce_wait blocks the host until it (and the host copy) is complete. Both release the slot.
 * NCCL is resolved at run time (dlopen libnccl.so.2; MPT_NCCL_LIBRARY overrides the name); without it the calls return
 * MPT_ERR_UNSUPPORTED. Errors: mpt_comm_last_error (NULL: errors of the calls that take no mpt_comm). */
#define MPT_COMM_UNIQUE_ID_BYTES 128
enum {
    MPT_REDUCE_ALL_RANKS = 0x1, /* ncclAllReduce: every rank receives the frame (each feeds its own Vulkan consumer) */
    MPT_REDUCE_KEEP_SUM = 0x2   /* do not divide by spp_total */
};
typedef struct mpt_comm mpt_comm;
MPT_API mpt_status mpt_comm_unique_id(uint8_t id[MPT_COMM_UNIQUE_ID_BYTES]); /* rank 0 calls it and hands the bytes to the other ranks */
MPT_API mpt_status mpt_comm_create(int device, const uint8_t id[MPT_COMM_UNIQUE_ID_BYTES], int rank, int world, mpt_comm** out);
MPT_API mpt_status mpt_comm_adopt(int device, void* nccl_comm, int rank, int world, mpt_comm** out); /* an ncclComm_t the host already has */
MPT_API void mpt_comm_destroy(mpt_comm* comm);
MPT_API const char* mpt_comm_last_error(const mpt_comm* comm);
MPT_API mpt_status mpt_comm_info(uint32_t* nccl_version);
/* d_partial_sum: this rank's RGBA32F sums (width*height float4). d_out: where the combined frame lands on the receiving rank(s) (may be
 * d_partial_sum itself); ignored elsewhere. h_out: optional pinned host buffer on the receiving rank(s). */
MPT_API mpt_status mpt_reduce_begin(mpt_comm* comm, const void* d_partial_sum, void* d_out, void* h_out, uint32_t width, uint32_t height,
                                    uint32_t spp_total, int root, uint32_t flags, void* render_stream, uint32_t slot);
MPT_API mpt_status mpt_reduce_end(mpt_comm* comm, uint32_t slot, void* stream);
MPT_API mpt_status mpt_reduce_wait(mpt_comm* comm, uint32_t slot);

/* Host-only (no device needed): the acceleration structure exactly as mpt_scene_upload lays it out in HBM — the BVH8 nodes of every
 * BLAS and of the TLAS, the triangle records in leaf order, the instance records. Replaces nothing in the reference (its BLAS/TLAS are
 * opaque driver objects, scene.cpp:1672-1892); it exists so that a CPU checker can walk the SAME tree the device walks (SURVEY.md
 * §8(d): nodes and triangles per ray are counted by the oracle on this structure) and so that the builder can be tested without a GPU.
 * The view's pointers stay valid until mpt_host_accel_free. Errors: mpt_last_error(NULL). */
typedef struct mpt_host_accel mpt_host_accel;
typedef struct mpt_host_accel_view {
    const void* nodes;           /* 80-byte BVH8 nodes (layout: merian_b200/csrc/bvh8.h) */
    uint64_t node_count;
    const void* tris;            /* 48-byte records: v0|primitive_id, v1|geometry_index, v2|flags (3 x float4) */
    uint64_t triangle_count;
    const void* instances;       /* 64-byte records: 3x4 world->object rows, BLAS root, custom index, mask|flags<<8, pad */
    uint64_t instance_count;
    const uint32_t* tlas_prims;  /* TLAS leaf record -> instance index */
    uint64_t tlas_prim_count;
    const uint32_t* prim_to_tri; /* geometry prim->record table */
    uint64_t prim_count;
    uint32_t root;               /* TLAS root (two_level) or the root of the only BLAS */
    uint32_t two_level;
    uint32_t max_blas_depth, tlas_depth;
    double build_ms;
} mpt_host_accel_view;
MPT_API mpt_status mpt_host_accel_build(const mpt_scene_desc* scene, mpt_host_accel** out);
MPT_API mpt_status mpt_host_accel_get(const mpt_host_accel* accel, mpt_host_accel_view* out);
MPT_API void mpt_host_accel_free(mpt_host_accel* accel);

/* Test hook for the BSDF property suite (tests/test_bsdf.cpp of the reference restated): runs
 * sample_eval / eval / pdf of one BSDF on the device. See merian_b200/csrc/bsdf_check.cu. */
typedef struct mpt_bsdf_check_params {
    uint32_t bsdf;        /* 0 Lambert, 1 GGX, 2 GGX aniso, 3 FresnelMix<GGX, Lambert> (= the clearcoat construction: ior_f0 must be 0.04),
                           * 4 IridescentFresnelMix, 5 IridescentRoughDielectric,
                           * 6 SheenBRDF alone (oracle only: on the device the sheen lobe exists only inside SheenLayer, MPT_ERR_UNSUPPORTED),
                           * 7 the whole GltfBSDF nest (metallic 0.5, clearcoat 0.5, sheen 0.5 * albedo, iridescence as given, no transmission),
                           * 8 Conductor, 9 SheenLayer<dielectric> with sheen colour = albedo */
    float wi[3];
    uint32_t n_samples;
    uint32_t seed;
    float albedo[3];
    float alpha;
    float f0[3];
    float ior_f0;
    float alpha_bitangent;
    float iridescence;
} mpt_bsdf_check_params;
MPT_API mpt_status mpt_bsdf_check(mpt_context* ctx, const mpt_bsdf_check_params* p, float* out16);

#ifdef __cplusplus
}
#endif
#endif /* MPT_H */
