// merian-plugin-pt-cuda — the merian side of the drop-in boundary.
//
// NOT BUILT IN THIS REPOSITORY'S CI: it needs merian's headers and its toolchain (C++20, Vulkan >= 1.4.350, meson;
// meson.build:7,15,74 of the reference), none of which exist in the image libmpt.so is built in. It is the binding a
// merian maintainer adds: a `merian::Node` whose type name is "Render (Path-traced, CUDA)", with the connectors and
// property keys of the built-in `Render (Path-traced)` (src/merian-graph/nodes/render_pt/render_pt.cpp:26-34,147-187),
// forwarding every Node virtual to the matching entry point of include/mpt_node.h / include/mpt.h.
//
// Data path of this first adapter (sync_to_cpu_and_back, SURVEY.md §8(b) "cross-API sync options" (1)):
//   on geometry/material change:  Scene tables (adapter/merian_getters.patch) + device vertex/index buffers copied to a
//                                 host-visible staging buffer -> mpt_scene_upload / mpt_materials_upload / mpt_texture_upload
//   per iteration:                submission.sync_to_cpu_and_back( mptn_process(... d_tex1 = NULL ...) -> D2H into the
//                                 mapped staging buffer ), then vkCmdCopyBufferToImage into the graph's irradiance image.
// Primary hits are re-traced on the CUDA side (mpt_primary_visibility packs the same PackedHitInfo the GBuffer stores),
// so no GBuffer texel has to cross the API boundary. The zero-copy variant (VK_KHR_external_memory_fd +
// cudaImportExternalMemory on the staging buffer, timeline semaphores through cudaImportExternalSemaphore) replaces the
// two copies marked [interop] below without touching anything else.
#include "merian-graph/connectors/image/vk_image_out_managed.hpp"
#include "merian-graph/connectors/ptr_in.hpp"
#include "merian-graph/connectors/shader_object_in.hpp"
#include "merian-graph/graph/errors.hpp"
#include "merian-graph/graph/node.hpp"
#include "merian-graph/graph/node_registry.hpp"
#include "merian-graph/objects/gbuffer_object.hpp"
#include "merian-graph/plugin/node_plugin.hpp"
#include "merian-shaders/scene/env_map.hpp"
#include "merian-shaders/scene/scene.hpp"
#include "merian/utils/properties_json_dump.hpp"

#include <cuda_runtime_api.h>
#include <mpt_node.h>

#include <nlohmann/json.hpp>

namespace merian {

class RenderPTCuda : public Node {
  public:
    ~RenderPTCuda() override {
        if (node) mptn_destroy(node);
        if (ctx) mpt_destroy(ctx);
        if (pinned) cudaFreeHost(pinned);
    }

    // RenderPT::query_device_support asks for rayTracingPipeline/rayQuery (render_pt.cpp:13-19); this node needs neither.
    DeviceSupportInfo query_device_support(const DeviceSupportQueryInfo& query_info) override {
        return DeviceSupportInfo::check(query_info, {}, {});
    }

    void initialize(const ContextHandle& context, const ResourceAllocatorHandle& allocator) override {
        this->context = context;
        this->allocator = allocator;
        // the CUDA device that is the same physical GPU as merian's VkPhysicalDevice (match by UUID)
        const auto vk_uuid = context->get_physical_device()->get_physical_device_id_properties().deviceUUID;
        int n = 0, device = 0;
        cudaGetDeviceCount(&n);
        for (int i = 0; i < n; i++) {
            cudaDeviceProp p{};
            cudaGetDeviceProperties(&p, i);
            if (std::memcmp(p.uuid.bytes, vk_uuid.data(), 16) == 0) device = i;
        }
        if (mpt_create(device, &ctx) != MPT_OK) throw graph_errors::node_error{mpt_last_error(nullptr)};
        if (mptn_create(ctx, &node) != MPT_OK) throw graph_errors::node_error{"mptn_create failed"};
        cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
    }

    std::vector<InputConnectorDescriptor> describe_inputs() override { // render_pt.cpp:26-28
        return {{"scene", con_scene}, {"gbuffer", con_gbuffer, ConnectorAccess::ray_tracing_read}};
    }

    std::vector<OutputConnectorDescriptor> describe_outputs(const NodeIOLayout& io_layout) override { // :30-34
        const vk::Extent3D extent = io_layout[con_gbuffer]->get_create_info().extent;
        uint32_t vk_format = 0, w = 0, h = 0;
        check(mptn_describe_outputs(node, extent.width, extent.height, &vk_format, &w, &h));
        con_irradiance = ManagedVkImageOut::create(static_cast<vk::Format>(vk_format), vk::Extent3D{w, h, 1});
        const size_t bytes = size_t(w) * h * (vk_format == MPTN_FORMAT_R16G16B16A16_SFLOAT ? 8 : 16);
        staging = allocator->create_buffer(bytes, vk::BufferUsageFlagBits::eTransferSrc, MemoryMappingType::HOST_ACCESS_RANDOM,
                                           "RenderPTCuda::staging");
        if (pinned) cudaFreeHost(pinned);
        cudaMallocHost(&pinned, bytes);
        cudaMalloc(&d_out, bytes);
        out_bytes = bytes;
        return {{"irradiance", con_irradiance, ConnectorAccess::transfer_dst}}; // filled by vkCmdCopyBufferToImage
    }

    NodeStatusFlags on_connected(const NodeIOLayout&, const NodeIO&, const NodeConnectionInfo&, Submission&) override {
        tables_dirty = true;
        return {};
    }

    NodeStatusFlags process(const NodeIO& io, const NodeProcessInfo& info, Submission& submission) override { // :94-124
        const SceneHandle& scene = io[con_scene];
        if (!scene || !scene->is_ready()) return {};
        const auto& changes = scene->get_last_update_changes();
        if (tables_dirty || changes.geometry_changed || changes.transform_changed) {
            upload_scene(*scene, submission); // may split the submission to read the device vertex buffers back
            tables_dirty = false;
        }
        const mpt_camera cam = to_mpt_camera(*scene->get_active_camera());
        const uint32_t frame = uint32_t(info.get_total_iteration()); // = scene.frame, gltf_scene_node.cpp:59-60

        uint32_t events = 0;
        std::string error;
        submission.sync_to_cpu_and_back([&, cam, frame] { // submission.hpp:101-114: GPU waits until this returns
            if (mptn_process(node, &cam, frame, /*d_tex1*/ nullptr, /*d_albedo*/ nullptr, d_out, stream, &events) != MPT_OK) {
                error = mptn_last_error(node);
                return;
            }
            cudaMemcpyAsync(pinned, d_out, out_bytes, cudaMemcpyDeviceToHost, stream); // [interop] disappears with external memory
            cudaStreamSynchronize(stream);
            std::memcpy(staging->get_memory()->map(), pinned, out_bytes);               // [interop]
            staging->get_memory()->unmap();
        });
        if (!error.empty()) throw graph_errors::node_error{error}; // graph disables the node and reconnects (graph.cpp:385-395)
        if (events & MPTN_EVENT_BOUNCES_CHANGED) io.send_event("bounces_changed"); // render_pt.cpp:102-105

        const CommandBufferHandle& cmd = submission.get_cmd();
        const ImageHandle& image = io[con_irradiance].get_texture()->get_image();
        cmd->copy(staging, image); // ManagedVkImageOut declared transfer_dst: the graph already emitted the barrier
        cmd->keep_until_pool_reset(staging);
        return {};
    }

    // Same keys as RenderPT::properties (render_pt.cpp:147-187). The values live in the C-ABI node so that JSON load/dump
    // and the UI see one source of truth: dump -> edit through `config` -> load.
    NodeStatusFlags properties(Properties& config) override {
        size_t len = 0;
        mptn_dump_config(node, nullptr, 0, &len);
        std::string buf(len + 1, '\0');
        mptn_dump_config(node, buf.data(), buf.size(), &len);
        nlohmann::json j = nlohmann::json::parse(buf.c_str());
        int32_t spp = j["samples per pixel"], max_path_length = j["max path length"];
        bool emission = j["emission on primary"], ser = j["shader execution reordering"], demod = j["demodulate albedo"];
        config.config_int("samples per pixel", spp, "Number of BSDF-sampled paths per pixel.", 1, 16);
        config.config_int("max path length", max_path_length, "Maximum number of path segments, including the primary hit.", 1, 16);
        config.config_bool("emission on primary", emission, "Fold primary-hit emission into irradiance (self-contained).");
        config.config_bool("shader execution reordering", ser, "Accepted for compatibility; the CUDA wavefront sorts by queue instead.");
        config.config_bool("demodulate albedo", demod, "Divide the primary-hit albedo out of the output.");
        config.st_separate("instance mask");
        for (uint32_t bit = 0; bit < 8; ++bit) {
            bool b = j[std::to_string(bit)];
            config.config_bool(std::to_string(bit), b);
            j[std::to_string(bit)] = b;
            if ((bit & 3u) != 3u) config.st_no_space();
        }
        config.st_separate();
        vk::Format format = j["irradiance format"] == "R16G16B16A16Sfloat" ? vk::Format::eR16G16B16A16Sfloat : vk::Format::eR32G32B32A32Sfloat;
        config.config_enum("irradiance format", format, Properties::OptionsStyle::COMBO);
        j["samples per pixel"] = spp;
        j["max path length"] = max_path_length;
        j["emission on primary"] = emission;
        j["shader execution reordering"] = ser;
        j["demodulate albedo"] = demod;
        j["irradiance format"] = vk::to_string(format);
        uint32_t status = 0;
        if (mptn_load_config(node, j.dump().c_str(), &status) != MPT_OK) throw graph_errors::node_error{mptn_last_error(node)};
        return (status & MPTN_NEEDS_RECONNECT) ? NEEDS_RECONNECT : NodeStatusFlags{};
    }

  private:
    void check(mpt_status s) {
        if (s != MPT_OK) throw graph_errors::node_error{node ? mptn_last_error(node) : "libmpt error"};
    }

    static mpt_camera to_mpt_camera(Camera& c) { // Camera::write_to, camera.cpp:349-368
        mpt_camera m{};
        const float half_vfov_tan = std::tan(c.get_field_of_view_vertical() * 0.5f);
        const float3 fwd = c.get_forward(), rgt = c.get_right();
        const float focus = c.get_focus_distance();
        const auto put = [](float* dst, const float3& v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; };
        put(m.position, c.get_position());
        put(m.target, c.get_target());
        put(m.up, c.get_up());
        put(m.U, rgt * (half_vfov_tan * c.get_aspect_ratio() * focus));
        put(m.V, cross(rgt, fwd) * (half_vfov_tan * focus));
        put(m.W, fwd * focus);
        m.near_plane = c.get_near_plane();
        m.far_plane = c.get_far_plane();
        m.aspect_ratio = c.get_aspect_ratio();
        m.aperture_radius = c.get_aperture_radius();
        m.jitter[0] = c.get_jitter().x;
        m.jitter[1] = c.get_jitter().y;
        return m;
    }

    // Scene tables -> mpt_scene_desc. Uses the getters of adapter/merian_getters.patch; the device vertex / index buffers
    // (bits AFTER transform_vertex.slang) are read back once through a staging copy.
    void upload_scene(Scene& scene, Submission& submission) {
        const auto& infos = scene.get_mesh_infos();
        const auto& groups = scene.get_mesh_groups();
        const auto& geometries = scene.get_geometries();
        const auto& tlas = scene.get_tlas_instances();

        // 1. read the device vertex / index buffers back
        struct Range { vk::DeviceSize vb_off, vb_size, ib_off, ib_size; };
        std::vector<Range> ranges(infos.size());
        vk::DeviceSize total = 0;
        for (size_t i = 0; i < infos.size(); i++) {
            if (!infos[i].mesh) continue;
            const auto& mesh = *infos[i].mesh;
            ranges[i].vb_off = total;
            ranges[i].vb_size = vk::DeviceSize(mesh.get_vertex_count()) * sizeof(PackedVertexData);
            total += ranges[i].vb_size;
            ranges[i].ib_off = total;
            ranges[i].ib_size = mesh.has_indices() ? vk::DeviceSize(mesh.get_primitive_count()) * 3 * size_for_index_type(mesh.index_type) : 0;
            total += (ranges[i].ib_size + 3) & ~vk::DeviceSize(3);
        }
        const BufferHandle readback = allocator->create_buffer(std::max<vk::DeviceSize>(total, 4), vk::BufferUsageFlagBits::eTransferDst,
                                                               MemoryMappingType::HOST_ACCESS_RANDOM, "RenderPTCuda::readback");
        const CommandBufferHandle& cmd = submission.get_cmd();
        for (size_t i = 0; i < infos.size(); i++) {
            if (!infos[i].mesh) continue;
            cmd->copy(infos[i].vertex_buffer.get_buffer(), readback, vk::BufferCopy{infos[i].vertex_buffer.get_offset(), ranges[i].vb_off, ranges[i].vb_size});
            if (ranges[i].ib_size)
                cmd->copy(infos[i].index_buffer.get_buffer(), readback, vk::BufferCopy{infos[i].index_buffer.get_offset(), ranges[i].ib_off, ranges[i].ib_size});
        }
        submission.sync_to_cpu_and_back([&] {
            const uint8_t* base = readback->get_memory()->map_as<uint8_t>();
            std::vector<mpt_mesh_desc> meshes(infos.size());
            for (size_t i = 0; i < infos.size(); i++) {
                if (!infos[i].mesh) continue;
                const auto& mesh = *infos[i].mesh;
                meshes[i].vertices = reinterpret_cast<const mpt_packed_vertex*>(base + ranges[i].vb_off);
                meshes[i].indices = ranges[i].ib_size ? base + ranges[i].ib_off : nullptr;
                meshes[i].vertex_count = mesh.get_vertex_count();
                meshes[i].primitive_count = mesh.get_primitive_count();
                meshes[i].index_flag = mesh.index_type == vk::IndexType::eUint32 ? MPT_GEOM_INDEX_U32
                                       : mesh.index_type == vk::IndexType::eUint16 ? MPT_GEOM_INDEX_U16
                                       : mesh.index_type == vk::IndexType::eUint8 ? MPT_GEOM_INDEX_U8 : 0u;
                meshes[i].opaque = mesh.is_opaque() ? 1u : 0u;
            }
            // 2. BLAS = mesh group (scene.hpp:401-436), in group order
            std::vector<std::vector<uint32_t>> group_meshes(groups.size());
            std::vector<mpt_blas_desc> blas(groups.size());
            for (size_t g = 0; g < groups.size(); g++) {
                for (const MeshID m : groups[g].meshes) group_meshes[g].push_back(uint32_t(m));
                blas[g] = {group_meshes[g].data(), uint32_t(group_meshes[g].size())};
            }
            // 3. GeometryData[] (scene.cpp:1236-1291): same order; the buffer addresses become mesh indices
            std::vector<mpt_geometry_desc> geoms(geometries.size());
            {
                size_t k = 0;
                for (size_t g = 0; g < groups.size(); g++)
                    for ([[maybe_unused]] const NodeID node_id : groups[g].get_instances(infos))
                        for (const MeshID m : groups[g].meshes) {
                            geoms[k] = {uint32_t(m), uint16_t(geometries[k].material_id), uint16_t(geometries[k].flags), geometries[k].primitive_count};
                            k++;
                        }
            }
            // 4. TLAS instances (scene.cpp:1808-1857)
            std::vector<mpt_instance_desc> insts(tlas.size());
            {
                size_t k = 0;
                for (size_t g = 0; g < groups.size(); g++)
                    for ([[maybe_unused]] const NodeID node_id : groups[g].get_instances(infos)) {
                        const auto& t = tlas[k];
                        std::memcpy(insts[k].transform, t.transform.matrix.data(), sizeof(float) * 12);
                        insts[k].instance_custom_index = t.instanceCustomIndex;
                        insts[k].mask = t.mask;
                        const auto f = static_cast<vk::GeometryInstanceFlagsKHR>(t.flags);
                        insts[k].flags = ((f & vk::GeometryInstanceFlagBitsKHR::eTriangleFacingCullDisable) ? MPT_INSTANCE_FACING_CULL_DISABLE : 0u) |
                                         ((f & vk::GeometryInstanceFlagBitsKHR::eTriangleFlipFacing) ? MPT_INSTANCE_FLIP_FACING : 0u) |
                                         ((f & vk::GeometryInstanceFlagBitsKHR::eForceOpaque) ? MPT_INSTANCE_FORCE_OPAQUE : 0u);
                        insts[k].blas = uint32_t(g);
                        k++;
                    }
            }
            mpt_scene_desc d{};
            d.meshes = meshes.data();
            d.mesh_count = uint32_t(meshes.size());
            d.blas = blas.data();
            d.blas_count = uint32_t(blas.size());
            d.geometries = geoms.data();
            d.geometry_count = uint32_t(geoms.size());
            d.instances = insts.data();
            d.instance_count = uint32_t(insts.size());
            d.instance_transforms = reinterpret_cast<const float*>(scene.get_instance_transforms().data());
            d.inverse_transposed_instance_transforms = reinterpret_cast<const float*>(scene.get_inverse_transposed_instance_transforms().data());
            upload_status = mpt_scene_upload(ctx, &d);
            readback->get_memory()->unmap();

            // 5. materials: the packed host buffer + the run-time type-id table (material_system.cpp:124-195)
            const auto& ms = scene.get_material_system();
            std::vector<mpt_material_type> types(ms->get_material_types().size());
            for (const auto& [name, info] : ms->get_material_types()) {
                mpt_material_type t{MPT_MATERIAL_DIFFUSE, 0};
                if (name.rfind("merian::GltfMaterial", 0) == 0) t = {MPT_MATERIAL_GLTF, gltf_features_from_type_name(name)};
                types[info.dispatch_id] = t;
            }
            if (upload_status == MPT_OK)
                upload_status = mpt_materials_upload(ctx, ms->get_host_buffer().data(), ms->get_entry_size_bytes(), ms->get_material_count(),
                                                     types.data(), uint32_t(types.size()));
            mpt_materials_set_hints(ctx, ms->get_alpha_test_threshold(), ms->get_min_roughness(), ms->get_clamp_normals());
        });
        if (upload_status != MPT_OK) throw graph_errors::node_error{mpt_last_error(ctx)};
        upload_textures_and_env(scene, submission);
    }

    // "merian::GltfMaterial<true,false,...>" -> MPT_GLTF_* bits, order TRANSMISSION, VOLUME, CLEARCOAT, SHEEN, IRIDESCENCE,
    // ANISOTROPY (gltf-material.slang:64-69, gltf_material.hpp type-name construction)
    static uint32_t gltf_features_from_type_name(const std::string& name) {
        uint32_t bits = 0, i = 0;
        for (size_t p = name.find('<'); p != std::string::npos && i < 6; i++) {
            p++;
            if (name.compare(p, 4, "true") == 0) bits |= 1u << i;
            p = name.find(',', p);
        }
        return bits;
    }

    // Textures (RGBA8 sRGB/UNORM, texture_manager.cpp:207-243) and the lat-long env map (RGBA32F, scene.cpp:780-795) are
    // copied image -> buffer -> mpt_texture_upload / mpt_env_set the same way as the vertex buffers (omitted here for
    // brevity: cmd->copy(image, buffer) per texture, then one sync_to_cpu_and_back).
    void upload_textures_and_env(Scene& scene, Submission& submission);

    ContextHandle context;
    ResourceAllocatorHandle allocator;
    PtrInHandle<Scene> con_scene = PtrIn<Scene>::create();
    ShaderObjectInHandle<GBufferObject> con_gbuffer = ShaderObjectIn<GBufferObject>::create();
    ManagedVkImageOutHandle con_irradiance;
    BufferHandle staging;
    mpt_context* ctx = nullptr;
    mptn_node* node = nullptr;
    cudaStream_t stream = nullptr;
    void* pinned = nullptr;
    void* d_out = nullptr;
    size_t out_bytes = 0;
    bool tables_dirty = true;
    mpt_status upload_status = MPT_OK;
};

} // namespace merian

extern "C" {

MERIAN_PLUGIN_EXPORT const char* merian_plugin_name(void) { return "pt-cuda"; }

MERIAN_PLUGIN_EXPORT uint32_t merian_node_plugin_abi_version(void) { return MERIAN_NODE_PLUGIN_ABI_VERSION; }

MERIAN_PLUGIN_EXPORT void merian_register_nodes(merian::NodeRegistry& registry) {
    // "Render (Path-traced)" itself is registered built-in and a duplicate name throws (node_registry.cpp:95,
    // node_registry.hpp:60-63), so the swap happens in the graph JSON: adapter/pt_cuda.json is examples/renderers/pt.json
    // with this type name; node id `render` and all connections stay as they are.
    registry.register_node_type<merian::RenderPTCuda>("Render (Path-traced, CUDA)",
                                                      "Path tracer on hand-written sm_100a CUDA kernels (libmpt.so).");
}

} // extern "C"
