/* mpt_tables.h — the host logic of merian::Scene that defines what `pt` sees, for hosts that are not
 * merian: source meshes + scene-graph nodes in, the flat mpt_scene_desc of mpt.h out (host-only, no CUDA call).
 *
 * Inside merian the plugin reads these tables from merian::Scene; a standalone C or C++ host has no Scene, so libmpt
 * restates the part of it the renderer depends on:
 *   vertex packing          src/merian-shaders/scene/gltf_scene.cpp:107-118, src/merian/utils/normal_encoding.cpp:9-25 (host encode)
 *   the device pre-transform src/merian-shaders/scene/transform_vertex.slang:12-40 (decode, transform, re-encode; runs even with identity)
 *   mesh grouping -> BLAS    src/merian-shaders/scene/scene.cpp:1023-1105 (split by FlipFacing / TwoSided and instance mask; meshes with
 *                            one instance form pre-transformed static groups, meshes with several share a group per instance set)
 *   GeometryData order       src/merian-shaders/scene/scene.cpp:1236-1291
 *   TLAS instances / ids     src/merian-shaders/scene/scene.cpp:1808-1857
 * Same names and meaning as the Python mirror merian_b200/scene.py (SceneTables, pack_vertices); tests/test_tables_cabi.py compares the two. */
#ifndef MPT_TABLES_H
#define MPT_TABLES_H

#include "mpt.h"

#ifdef __cplusplus
extern "C" {
#endif

/* merian::Scene::MeshFlags, include/merian-shaders/scene/scene.hpp:39-58 */
enum {
    MPT_MESH_IS_MORPHED = 0x1,
    MPT_MESH_HAS_VARIABLE_TOPOLOGY = 0x2,
    MPT_MESH_IS_OPAQUE = 0x4,
    MPT_MESH_FLIP_FACING = 0x8,
    MPT_MESH_TWO_SIDED = 0x10,
    MPT_MESH_HAS_TANGENTS = 0x20,
    MPT_MESH_FLAT_SHADING = 0x40,
    MPT_MESH_USE_ENV_MAP = 0x80
};

/* One source mesh as a loader hands it to Scene::add_mesh (gltf_scene.cpp:88-160): float attributes, optional indices. */
typedef struct mpt_src_mesh {
    const float* positions; /* vertex_count x 3 */
    const float* normals;   /* vertex_count x 3 */
    const float* uvs;       /* vertex_count x 2, or NULL (zeros) */
    const float* tangents;  /* vertex_count x 4 (xyz + handedness), or NULL */
    uint32_t vertex_count;
    const void* indices;    /* index triples, or NULL: sequential (vertex_count / 3 primitives) */
    uint32_t primitive_count;
    uint32_t index_size;    /* bytes per index: 1, 2 or 4 (ignored without indices) */
    uint32_t material_id;
    uint32_t flags;         /* MPT_MESH_* */
    uint32_t instance_mask; /* 8 bit */
} mpt_src_mesh;

/* One scene-graph node with its GLOBAL transform and the meshes it instantiates (Scene::Node, scene.hpp:300-330). */
typedef struct mpt_src_node {
    float transform[16]; /* row-major 4x4 object -> world */
    const uint32_t* meshes;
    uint32_t mesh_count;
} mpt_src_node;

typedef struct mpt_tables mpt_tables;

/* Builds the tables. Morphed / variable-topology meshes (animated vertex data) are refused with MPT_ERR_UNSUPPORTED.
 * The source arrays may be released after the call (indices are copied). Errors: mpt_last_error(NULL). */
MPT_API mpt_status mptt_build(const mpt_src_mesh* meshes, uint32_t mesh_count, const mpt_src_node* nodes, uint32_t node_count, mpt_tables** out);
/* The result, valid until mptt_free: pass it to mpt_scene_upload / mpt_host_accel_build. */
MPT_API const mpt_scene_desc* mptt_scene_desc(const mpt_tables* t);
MPT_API void mptt_free(mpt_tables* t);

/* The vertex path alone (A13): host packing followed by the device transform pass with `transform` (row-major 4x4, NULL = identity).
 * uvs / tangents may be NULL. Writes vertex_count records. */
MPT_API mpt_status mptt_pack_vertices(const float* positions, const float* normals, const float* uvs, const float* tangents, uint32_t vertex_count,
                                      const float* transform, mpt_packed_vertex* out);

/* ---- material buffer (the host side of SURVEY.md 8(a) A6) ---------------------------------------------------------------------------
 * MaterialSystem::add_material / repack_host_buffer (src/merian-shaders/shading/materials/material_system.cpp:124-195): every material
 * becomes one entry = MaterialHeader {u16 material_model_type_id, u16 alpha_texture_id} + the model's serialised payload, all entries
 * padded to the largest one; type ids are handed out in first-registration order of (model, feature set). Payloads:
 * DiffuseMaterial = float4 base colour (diffuse-material.slang:8-12); GltfMaterial = the core block followed by the blocks of the
 * enabled extensions in the order transmission, volume, clearcoat, sheen, iridescence, anisotropy
 * (include/merian-shaders/shading/materials/gltf_material.hpp:93-146). Texture ids: MPT_TEXTURE_NONE = no texture. */
typedef struct mpt_src_material {
    uint32_t model;                  /* MPT_MATERIAL_DIFFUSE | MPT_MATERIAL_GLTF */
    float base_color_factor[4];
    uint32_t base_color_texture;     /* = the header's alpha_texture_id */
    /* ---- glTF metallic-roughness only */
    float emissive_factor[3];
    float metallic_factor, roughness_factor, normal_scale, occlusion_strength, ior;
    uint32_t metallic_roughness_texture, normal_texture, occlusion_texture, emissive_texture;
    uint32_t features;               /* MPT_GLTF_* bits: which of the blocks below exist */
    float transmission_weight;
    float volume_absorption[3];
    float clearcoat_weight, clearcoat_roughness;
    uint32_t clearcoat_texture, clearcoat_roughness_texture;
    float sheen_color[3], sheen_roughness;
    uint32_t sheen_color_texture, sheen_roughness_texture;
    float iridescence_factor, iridescence_ior, iridescence_thickness_min, iridescence_thickness_max;
    uint32_t iridescence_texture, iridescence_thickness_texture;
    float anisotropy_strength, anisotropy_rotation;
    uint32_t anisotropy_texture;
} mpt_src_material;

typedef struct mpt_material_table mpt_material_table;
typedef struct mpt_material_table_view { /* the arguments of mpt_materials_upload */
    const void* entries;
    uint32_t entry_size, material_count;
    const mpt_material_type* types;
    uint32_t type_count;
} mpt_material_table_view;

MPT_API mpt_status mptt_materials_build(const mpt_src_material* materials, uint32_t material_count, mpt_material_table** out);
MPT_API mpt_status mptt_materials_get(const mpt_material_table* table, mpt_material_table_view* out);
MPT_API void mptt_materials_free(mpt_material_table* table);

/* ---- camera (the host side of SURVEY.md 8(a) A4) ------------------------------------------------------------------------------------
 * Camera::write_to (src/merian/utils/camera/camera.cpp:349-368): the basis the shaders read — W = forward x focus distance (the distance to
 * `target`, camera.cpp:255-260), U = right x tan(yfov / 2) x aspect x focus, V = up x tan(yfov / 2) x focus. jitter: sub-pixel offset in
 * pixels (Camera::advance_jitter, camera.cpp:318-347; NULL = none). aperture_radius 0 = pinhole. */
MPT_API mpt_status mptt_camera_look_at(const float position[3], const float target[3], const float up[3], float yfov_degrees, float aspect,
                                       float near_plane, float far_plane, float aperture_radius, const float jitter[2], mpt_camera* out);

/* ---- glTF ingestion -----------------------------------------------------------------------------------------------
 * GLTFScene::load (src/merian-shaders/scene/gltf_scene.cpp:831-897) without a Vulkan Scene: a .gltf / .glb file becomes the source tables
 * above — one mesh per TRIANGLES primitive with its flags (:607-694), nodes with GLOBAL transforms in DFS order (:146-196, 696-716),
 * materials with the KHR extension gates (:265-536), textures in first-use order with the sRGB / linear decision of the slot that uses
 * them (:538-604), the first perspective camera (:719-775). Images are NOT decoded (the reference uses stb_image, which is not vendored
 * here): every texture comes as the encoded bytes found in the file (PNG, JPEG, ...) plus its sampler state; the caller decodes to RGBA8
 * and calls mpt_texture_upload_ex(texture id = position in `textures`). Float attributes only, like the reference (:127-146).
 * The Python mirror is merian_b200/gltf.py (tests/test_gltf_cabi.py compares the two). */
typedef struct mpt_gltf mpt_gltf;
typedef struct mpt_gltf_texture {
    const uint8_t* encoded;  /* the image file as stored (bufferView, data URI or external file) */
    uint64_t encoded_size;
    uint32_t srgb;           /* 1: colour data (MPT_TEX_RGBA8_SRGB), 0: linear data (MPT_TEX_RGBA8_UNORM) */
    uint32_t wrap_u, wrap_v; /* MPT_WRAP_* */
    uint32_t filter;         /* MPT_FILTER_* */
    uint32_t mipmap;         /* MPT_MIPMAP_* */
} mpt_gltf_texture;
typedef struct mpt_gltf_view {
    const mpt_src_mesh* meshes;
    uint32_t mesh_count;
    const mpt_src_node* nodes;
    uint32_t node_count;
    const mpt_src_material* materials;
    uint32_t material_count;
    const mpt_gltf_texture* textures;
    uint32_t texture_count;
    uint32_t has_camera;     /* 0: the file has no perspective camera in its default scene (the caller supplies one) */
    mpt_camera camera;       /* jitter 0; glTF cameras jitter with the R2 sequence over 16 frames (gltf_scene.cpp:773-774): set camera.jitter per frame */
    float yfov_degrees, aspect, znear, zfar;
} mpt_gltf_view;

MPT_API mpt_status mptt_gltf_load(const char* path, mpt_gltf** out);
MPT_API mpt_status mptt_gltf_get(const mpt_gltf* gltf, mpt_gltf_view* out); /* pointers stay valid until mptt_gltf_free */
MPT_API void mptt_gltf_free(mpt_gltf* gltf);

#ifdef __cplusplus
}
#endif
#endif
