/* mpt_node.h — the `Render (Path-traced)` node, host side, as a C-ABI object on top of mpt.h.
 *
 * merian's node is a C++ class (include/merian-graph/nodes/render_pt/render_pt.hpp:26-87). Its toolchain (Vulkan >= 1.4.350,
 * slang, meson) is not present where this library is built, so the node's host logic — properties with the reference's
 * JSON keys and load/dump semantics, output description, the per-iteration process() — lives here behind plain C entry
 * points, and the merian plugin (adapter/merian_plugin_pt_cuda.cpp) is a thin `merian::Node` subclass
 * that forwards to them. Names follow the reference: one entry point per Node virtual.
 */
#ifndef MPT_NODE_H
#define MPT_NODE_H

#include "mpt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mptn_node mptn_node;

/* Node::NodeStatusFlagBits, include/merian-graph/graph/node.hpp:57-65 (same values). */
enum { MPTN_NEEDS_RECONNECT = 0x1, MPTN_RESET_IN_FLIGHT_DATA = 0x2, MPTN_REMOVE_NODE = 0x4 };
/* events process() asks the adapter to send with io.send_event (render_pt.cpp:102-105) */
enum { MPTN_EVENT_BOUNCES_CHANGED = 0x1 };
/* VkFormat values accepted for "irradiance format" (render_pt.hpp:53 default eR32G32B32A32Sfloat) */
enum { MPTN_FORMAT_R16G16B16A16_SFLOAT = 97, MPTN_FORMAT_R32G32B32A32_SFLOAT = 109 };

/* RenderPT::RenderPT + initialize (render_pt.cpp:11,21-25). The node renders on `ctx`'s device. */
MPT_API mpt_status mptn_create(mpt_context* ctx, mptn_node** out_node);
MPT_API void mptn_destroy(mptn_node* node);
MPT_API const char* mptn_last_error(const mptn_node* node);

/* RenderPT::properties through Node::load_config (src/merian-graph/graph/node.cpp:11-21 +
 * src/merian/utils/properties_json_load.cpp:126-180): `json` is the node's "properties" object, e.g.
 *   {"samples per pixel": 16, "max path length": 5, "emission on primary": true, "0": true, ...}
 * Keys (render_pt.cpp:147-187): "samples per pixel" (int), "max path length" (int), "emission on primary" (bool),
 * "shader execution reordering" (bool; accepted, no effect: there is no SER on this device), "demodulate albedo" (bool),
 * "0".."7" (bool, instance-mask bits), "irradiance format" (string: "R32G32B32A32Sfloat" | "R16G16B16A16Sfloat").
 * One key the reference does not have: "next event estimation" (bool, default false = the reference estimator): MPT_RENDER_NEE.
 * Missing keys keep their value, unknown keys are ignored, UI min/max are NOT applied (properties_json_load.cpp:126-133),
 * a value of the wrong JSON type is an error (nlohmann::json::get throws there). out_status: MPTN_NEEDS_RECONNECT when
 * the irradiance format changed (render_pt.cpp:180-185). */
MPT_API mpt_status mptn_load_config(mptn_node* node, const char* json, uint32_t* out_status);
/* Node::dump_config: writes the properties object (keys sorted like nlohmann::json) into buf; out_len = bytes needed
 * (excluding the terminator). */
MPT_API mpt_status mptn_dump_config(mptn_node* node, char* buf, size_t capacity, size_t* out_len);

/* RenderPT::describe_outputs (render_pt.cpp:30-34): the irradiance image has the GBuffer's extent and the configured
 * format. */
MPT_API mpt_status mptn_describe_outputs(mptn_node* node, uint32_t gbuffer_width, uint32_t gbuffer_height,
                                         uint32_t* out_vk_format, uint32_t* out_width, uint32_t* out_height);

/* Sample sharding for multi-GPU hosts (no reference counterpart): this node renders sample indices
 * rank, rank + world, ... of each frame's "samples per pixel" and writes SUMS; the host reduces and normalises. */
MPT_API mpt_status mptn_set_shard(mptn_node* node, uint32_t rank, uint32_t world);

/* RenderPT::process (render_pt.cpp:94-124). `frame` = info.get_total_iteration() (what Scene::update writes into
 * scene.frame, gltf_scene_node.cpp:59-60). d_tex1: GBuffer hit texture (gbuffer.slang:40-61) or NULL to trace primary
 * rays here; d_albedo: GBuffer tex3 or NULL; d_irradiance: width*height texels of the configured format.
 * Returns MPT_OK and does nothing when the scene is not ready (render_pt.cpp:98-99). out_events: MPTN_EVENT_*. */
MPT_API mpt_status mptn_process(mptn_node* node, const mpt_camera* camera, uint32_t frame, const void* d_tex1,
                                const void* d_albedo, void* d_irradiance, void* stream, uint32_t* out_events);

/* ---- `GBuffer (Raytraced)` node, host side -------------------------------------------------
 * GBufferRTNode (include/merian-graph/nodes/gbuffer_rt/gbuffer.hpp:20-75, src/merian-graph/nodes/gbuffer_rt/gbuffer.cpp).
 * Properties (gbuffer.cpp:155-181): "width", "height" (uint; NEEDS_RECONNECT on change), "0".."7" (instance-mask bits),
 * "emission format" ("R32G32B32A32Sfloat"). Outputs: "gbuffer" (tex0..tex3) and "emission", both of the node's extent. */
typedef struct mptn_gbuffer mptn_gbuffer;
MPT_API mpt_status mptn_gbuffer_create(mpt_context* ctx, mptn_gbuffer** out_node);
MPT_API void mptn_gbuffer_destroy(mptn_gbuffer* node);
MPT_API const char* mptn_gbuffer_last_error(const mptn_gbuffer* node);
MPT_API mpt_status mptn_gbuffer_load_config(mptn_gbuffer* node, const char* json, uint32_t* out_status);
MPT_API mpt_status mptn_gbuffer_dump_config(mptn_gbuffer* node, char* buf, size_t capacity, size_t* out_len);
/* GBufferRTNode::pre_process (gbuffer.cpp:104-115): the scene camera's resolution overrides the properties. */
MPT_API mpt_status mptn_gbuffer_set_resolution(mptn_gbuffer* node, uint32_t width, uint32_t height, uint32_t* out_status);
MPT_API mpt_status mptn_gbuffer_describe_outputs(mptn_gbuffer* node, uint32_t* out_width, uint32_t* out_height, uint32_t* out_emission_vk_format);
/* GBufferRTNode::process (gbuffer.cpp:117-146): one mpt_gbuffer over the node's extent; a no-op while the scene is not ready. */
MPT_API mpt_status mptn_gbuffer_process(mptn_gbuffer* node, const mpt_camera* camera, const mpt_camera* prev_camera,
                                        const mpt_gbuffer_images* images, void* stream);

/* ---- `Accumulate` node, plain mode, host side ----------------------------------------------
 * Accumulate (src/merian-graph/nodes/accumulate/accumulate.cpp:153-250). Properties with the reference's keys: "alpha", "max history",
 * "inf history", "clear", "use motion vectors" (must be false here: same-pixel reuse), "gbuffer check" ("always" | "only if moving" |
 * "never"), "normal threshold" (radians), "depth threshold", "extended search" (accepted, unused), "firefly filter enable" and
 * "adaptivity" (must stay off: they remain in the reference node), "hard clamp", "inf clamp"; plus "f32 history" (extension: no 2048 cap).
 * process(): clears on iteration 0 or after a clear request and reports MPTN_EVENT_CLEAR (accumulate.cpp:154-160). */
enum { MPTN_EVENT_CLEAR = 0x2 };
typedef struct mptn_accumulate mptn_accumulate;
MPT_API mpt_status mptn_accumulate_create(mpt_context* ctx, mptn_accumulate** out_node);
MPT_API void mptn_accumulate_destroy(mptn_accumulate* node);
MPT_API const char* mptn_accumulate_last_error(const mptn_accumulate* node);
MPT_API mpt_status mptn_accumulate_load_config(mptn_accumulate* node, const char* json, uint32_t* out_status);
MPT_API mpt_status mptn_accumulate_dump_config(mptn_accumulate* node, char* buf, size_t capacity, size_t* out_len);
MPT_API mpt_status mptn_accumulate_request_clear(mptn_accumulate* node);
MPT_API mpt_status mptn_accumulate_process(mptn_accumulate* node, uint64_t iteration, uint32_t width, uint32_t height,
                                           const void* d_src, const mpt_accumulate_buffers* buffers, void* stream,
                                           uint32_t* out_events);
/* The pt node and the Accumulate node of one graph fused into one call (mpt_render_accumulate): the irradiance connector between
 * them (`irradiance->accumulate.src`, examples/gltf.json:155-158) never touches HBM. */
MPT_API mpt_status mptn_process_accumulate(mptn_node* pt, mptn_accumulate* acc, const mpt_camera* camera, uint32_t frame,
                                           uint64_t iteration, const void* d_tex1, const void* d_albedo,
                                           const mpt_accumulate_buffers* buffers, void* stream, uint32_t* out_events);

#ifdef __cplusplus
}
#endif
#endif /* MPT_NODE_H */
