// merian-plugin-pt-cuda — the merian side of the drop-in boundary.
//
// A `merian::Node` whose type name is "Render (Path-traced, CUDA)", with the connectors and property keys of the built-in
// `Render (Path-traced)` (src/merian-graph/nodes/render_pt/render_pt.cpp:26-34,147-187), forwarding every Node virtual to the matching
// entry point of include/mpt_node.h / include/mpt.h. It needs merian's headers and toolchain (C++20, Vulkan >= 1.4.350, meson;
// meson.build:7,15,74 of the reference), none of which exist in the image libmpt.so is built in: tests/test_adapter_syntax.py compiles
// this file against tests/adapter_stubs/ (the declarations used here, each citing the reference header it restates) so that it is
// checked code, not prose; a maintainer builds it with adapter/meson.build inside a merian checkout with adapter/merian_getters.patch applied.
//
// Data path per iteration n (zero host round trips; SURVEY.md §8(b) "cross-API sync options" (2)):
//   Vulkan  vkCmdCopyImageToBuffer  gbuffer tex1 (PackedHitInfo) [+ tex3 (albedo) with "demodulate albedo"] -> exported buffer `in[slot]`
//           submit; empty batch signalling the exported timeline semaphore S_vk = n         (queue.hpp:29-30)
//   CUDA    mpt_wait_external_semaphore(S_vk, n) -> mptn_process(in[slot] -> out[slot]) -> mpt_signal_external_semaphore(S_cu, n)
//   Vulkan  empty batch waiting S_cu = n (covers everything later in submission order), vkCmdCopyBufferToImage out[slot] -> irradiance
// `slot` = info.get_in_flight_index(): the graph guarantees that the iteration that last used the slot has finished
// (node_process_info.hpp:42-46), so CUDA never overwrites a buffer Vulkan still reads. merian's own TimelineSemaphore cannot be exported
// (private constructor, semaphore_timeline.hpp:10-29) and its allocator has no export flags, so the node owns these few raw Vulkan objects.
//
// On geometry / material change (rare): Scene tables (adapter/merian_getters.patch) + device vertex / index buffers copied to a
// host-visible buffer -> mpt_scene_upload / mpt_materials_upload (the BVH8 build runs on the host); textures and the lat-long env map
// image -> exported buffer -> mpt_texture_upload_device / mpt_env_set_device (device to device).
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

#include <cuda_runtime_api.h>
#include <mpt_node.h>

#include <nlohmann/json.hpp>

#include <unistd.h>

#include <array>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace merian {

namespace {

constexpr uint32_t kMaxInFlight = 4; // mpt_render_host_begin's bound as well; merian's default ring is 2 (examples/gltf.json:39)

uint32_t find_device_local_memory_type(const vk::PhysicalDeviceMemoryProperties& props, const uint32_t type_bits) {
    for (uint32_t i = 0; i < props.memoryTypeCount; i++)
        if ((type_bits & (1u << i)) && (props.memoryTypes[i].propertyFlags & vk::MemoryPropertyFlagBits::eDeviceLocal)) return i;
    throw graph_errors::node_error{"no device-local memory type for an exportable buffer"};
}

// A VkBuffer on its own exportable VkDeviceMemory (dedicated allocation, OPAQUE_FD), imported into CUDA once. The graph's
// ResourceAllocator cannot be used: VkExportMemoryAllocateInfo has to be chained into the allocation (include/mpt.h "Vulkan <-> CUDA interop").
class ExportedBuffer {
  public:
    ExportedBuffer(const ContextHandle& context, mpt_context* mpt, const vk::DeviceSize bytes, const vk::BufferUsageFlags usage)
        : device(context->get_device()->get_device()), mpt(mpt), size(bytes) {
        const vk::ExternalMemoryBufferCreateInfo external_info{vk::ExternalMemoryHandleTypeFlagBits::eOpaqueFd};
        vk::BufferCreateInfo buffer_info{{}, bytes, usage, vk::SharingMode::eExclusive};
        buffer_info.pNext = &external_info;
        buffer = device.createBuffer(buffer_info);
        const vk::MemoryRequirements req = device.getBufferMemoryRequirements(buffer);
        const vk::MemoryDedicatedAllocateInfo dedicated_info{{}, buffer};
        vk::ExportMemoryAllocateInfo export_info{vk::ExternalMemoryHandleTypeFlagBits::eOpaqueFd};
        export_info.pNext = &dedicated_info;
        vk::MemoryAllocateInfo alloc_info{
            req.size, find_device_local_memory_type(context->get_physical_device()->get_memory_properties().memoryProperties, req.memoryTypeBits)};
        alloc_info.pNext = &export_info;
        memory = device.allocateMemory(alloc_info);
        device.bindBufferMemory(buffer, memory, 0);
        const int fd = device.getMemoryFdKHR(vk::MemoryGetFdInfoKHR{memory, vk::ExternalMemoryHandleTypeFlagBits::eOpaqueFd});
        // a successful import owns the descriptor (include/mpt.h); a failed one leaves it with us
        if (mpt_import_external_memory_fd(mpt, fd, req.size, MPT_EXTERNAL_MEMORY_OPAQUE_FD, MPT_EXTERNAL_MEMORY_DEDICATED, &imported, &d_ptr) != MPT_OK) {
            ::close(fd);
            destroy_vulkan();
            throw graph_errors::node_error{mpt_last_error(mpt)};
        }
    }
    ~ExportedBuffer() {
        if (imported) mpt_release_external_memory(mpt, imported);
        destroy_vulkan();
    }
    ExportedBuffer(const ExportedBuffer&) = delete;
    ExportedBuffer& operator=(const ExportedBuffer&) = delete;

    vk::Buffer buffer;
    void* d_ptr = nullptr; // the same bytes as a CUDA device pointer

  private:
    void destroy_vulkan() {
        if (buffer) device.destroyBuffer(buffer);
        if (memory) device.freeMemory(memory);
        buffer = nullptr;
        memory = nullptr;
    }
    vk::Device device;
    mpt_context* mpt;
    vk::DeviceSize size;
    vk::DeviceMemory memory;
    mpt_external_memory* imported = nullptr;
};

// A timeline VkSemaphore exported as OPAQUE_FD and imported into CUDA. Signalled / waited on the Vulkan side with raw, empty queue batches.
class ExportedTimeline {
  public:
    ExportedTimeline(const ContextHandle& context, mpt_context* mpt) : device(context->get_device()->get_device()), mpt(mpt) {
        const vk::ExportSemaphoreCreateInfo export_info{vk::ExternalSemaphoreHandleTypeFlagBits::eOpaqueFd};
        vk::SemaphoreTypeCreateInfo type_info{vk::SemaphoreType::eTimeline, 0};
        type_info.pNext = &export_info;
        vk::SemaphoreCreateInfo create_info{};
        create_info.pNext = &type_info;
        semaphore = device.createSemaphore(create_info);
        const int fd = device.getSemaphoreFdKHR(vk::SemaphoreGetFdInfoKHR{semaphore, vk::ExternalSemaphoreHandleTypeFlagBits::eOpaqueFd});
        if (mpt_import_external_semaphore_fd(mpt, fd, MPT_EXTERNAL_SEMAPHORE_TIMELINE_FD, &imported) != MPT_OK) {
            device.destroySemaphore(semaphore);
            throw graph_errors::node_error{mpt_last_error(mpt)};
        }
    }
    ~ExportedTimeline() {
        if (imported) mpt_release_external_semaphore(mpt, imported);
        device.destroySemaphore(semaphore);
    }
    ExportedTimeline(const ExportedTimeline&) = delete;
    ExportedTimeline& operator=(const ExportedTimeline&) = delete;

    // An empty batch that signals `value` once everything submitted to the queue so far has finished (signal operations cover all
    // commands earlier in submission order).
    void queue_signal(const QueueHandle& queue, const uint64_t value) const {
        const vk::TimelineSemaphoreSubmitInfo values{0, nullptr, 1, &value};
        vk::SubmitInfo submit{0, nullptr, nullptr, 0, nullptr, 1, &semaphore};
        submit.pNext = &values;
        queue->submit(submit, vk::Fence{});
    }
    // An empty batch that waits for `value`: the wait's second scope includes every command later in submission order on this queue.
    void queue_wait(const QueueHandle& queue, const uint64_t value, const vk::PipelineStageFlags stage) const {
        const vk::TimelineSemaphoreSubmitInfo values{1, &value, 0, nullptr};
        vk::SubmitInfo submit{1, &semaphore, &stage, 0, nullptr, 0, nullptr};
        submit.pNext = &values;
        queue->submit(submit, vk::Fence{});
    }

    vk::Semaphore semaphore;
    mpt_external_semaphore* imported = nullptr;

  private:
    vk::Device device;
    mpt_context* mpt;
};

vk::BufferImageCopy whole_image(const vk::Extent3D& extent, const vk::DeviceSize buffer_offset) {
    return vk::BufferImageCopy{buffer_offset, 0, 0, vk::ImageSubresourceLayers{vk::ImageAspectFlagBits::eColor, 0, 0, 1}, vk::Offset3D{0, 0, 0}, extent};
}

} // namespace

class RenderPTCuda : public Node {
  public:
    ~RenderPTCuda() override {
        release_interop();
        if (node) mptn_destroy(node);
        if (ctx) mpt_destroy(ctx);
        if (stream) cudaStreamDestroy(stream);
    }

    // RenderPT::query_device_support asks for rayTracingPipeline / rayQuery (render_pt.cpp:13-19); this node needs neither, only the
    // two export extensions (extension.hpp:70-78: required features, optional features, required extensions).
    DeviceSupportInfo query_device_support(const DeviceSupportQueryInfo& query_info) override {
        return DeviceSupportInfo::check(query_info, {"timelineSemaphore"}, {}, {"VK_KHR_external_memory_fd", "VK_KHR_external_semaphore_fd"});
    }

    void initialize(const ContextHandle& context, const ResourceAllocatorHandle& allocator) override {
        this->context = context;
        this->allocator = allocator;
        // the CUDA device that is the same physical GPU as merian's VkPhysicalDevice (match by UUID)
        const vk::PhysicalDeviceIDProperties& id = context->get_physical_device()->get_properties().get_id_properties();
        int n = 0, device = -1;
        cudaGetDeviceCount(&n);
        for (int i = 0; i < n; i++) {
            cudaDeviceProp p{};
            cudaGetDeviceProperties(&p, i);
            if (std::memcmp(p.uuid.bytes, id.deviceUUID.data(), 16) == 0) device = i;
        }
        if (device < 0) throw graph_errors::node_error{"no CUDA device with the UUID of merian's Vulkan device"};
        if (mpt_create(device, &ctx) != MPT_OK) throw graph_errors::node_error{mpt_last_error(nullptr)};
        if (mptn_create(ctx, &node) != MPT_OK) throw graph_errors::node_error{"mptn_create failed"};
        cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
    }

    std::vector<InputConnectorDescriptor> describe_inputs() override { // render_pt.cpp:26-28; the hit / albedo images are copied, not bound
        return {{"scene", con_scene}, {"gbuffer", con_gbuffer, ConnectorAccess::transfer_src}};
    }

    std::vector<OutputConnectorDescriptor> describe_outputs(const NodeIOLayout& io_layout) override { // :30-34
        const vk::Extent3D gbuffer_extent = io_layout[con_gbuffer]->get_create_info().extent;
        uint32_t vk_format = 0, w = 0, h = 0;
        check(mptn_describe_outputs(node, gbuffer_extent.width, gbuffer_extent.height, &vk_format, &w, &h));
        extent = vk::Extent3D{w, h, 1};
        texel_bytes = vk_format == MPTN_FORMAT_R16G16B16A16_SFLOAT ? 8 : 16;
        con_irradiance = ManagedVkImageOut::create(static_cast<vk::Format>(vk_format), extent);
        return {{"irradiance", con_irradiance, ConnectorAccess::transfer_dst}}; // filled by vkCmdCopyBufferToImage
    }

    NodeStatusFlags on_connected(const NodeIOLayout&, const NodeIO&, const NodeConnectionInfo&, Submission&) override {
        // all in-flight work has finished here (node.hpp:171-176): the exported objects can be rebuilt for the new extent
        release_interop();
        const vk::DeviceSize pixels = vk::DeviceSize(extent.width) * extent.height;
        for (uint32_t slot = 0; slot < kMaxInFlight; slot++) {
            in[slot] = std::make_unique<ExportedBuffer>(context, ctx, pixels * 32, vk::BufferUsageFlagBits::eTransferDst);  // tex1 | tex3
            out[slot] = std::make_unique<ExportedBuffer>(context, ctx, pixels * texel_bytes, vk::BufferUsageFlagBits::eTransferSrc);
        }
        vk_done = std::make_unique<ExportedTimeline>(context, ctx);
        cuda_done = std::make_unique<ExportedTimeline>(context, ctx);
        timeline_value = 0;
        tables_dirty = true;
        return {};
    }

    NodeStatusFlags process(const NodeIO& io, const NodeProcessInfo& info, Submission& submission) override { // :94-124
        const SceneHandle& scene = io[con_scene];
        if (!scene || !scene->is_ready()) return {};
        const Scene::UpdateChanges& changes = scene->get_last_update_changes();
        if (tables_dirty || changes.geometry_changed || changes.transform_changed || scene->get_material_system()->version() != materials_version) {
            upload_scene(*scene, submission); // splits the submission to read the device vertex buffers back
            tables_dirty = false;
        }
        const mpt_camera cam = to_mpt_camera(*scene->get_active_camera());
        const uint32_t frame = uint32_t(info.get_total_iteration()); // = scene.frame, gltf_scene_node.cpp:59-60
        const uint32_t slot = info.get_in_flight_index() % kMaxInFlight;
        const vk::DeviceSize pixels = vk::DeviceSize(extent.width) * extent.height;
        const bool demodulate = demodulate_albedo();

        // 1. GBuffer hit texture (and albedo) -> exported input buffer. The images stay in eGeneral (gbuffer_object.hpp:57-60).
        const auto gbuf = io[con_gbuffer];
        const vk::CommandBuffer& raw = submission.get_cmd()->get_command_buffer();
        raw.copyImageToBuffer(gbuf->get_image(1)->get_image(), vk::ImageLayout::eGeneral, in[slot]->buffer, whole_image(extent, 0));
        if (demodulate)
            raw.copyImageToBuffer(gbuf->get_image(3)->get_image(), vk::ImageLayout::eGeneral, in[slot]->buffer, whole_image(extent, pixels * 16));

        // 2. hand over to CUDA: S_vk = n after everything recorded so far
        const uint64_t n = ++timeline_value;
        submission.submit();
        vk_done->queue_signal(submission.get_queue(), n);

        // 3. CUDA: wait, render into the exported output buffer, signal. Nothing here blocks the host.
        uint32_t events = 0;
        check(mpt_wait_external_semaphore(ctx, vk_done->imported, n, stream));
        const char* d_in = static_cast<const char*>(in[slot]->d_ptr);
        if (mptn_process(node, &cam, frame, d_in, demodulate ? d_in + pixels * 16 : nullptr, out[slot]->d_ptr, stream, &events) != MPT_OK)
            throw graph_errors::node_error{mptn_last_error(node)}; // the graph disables the node and reconnects (graph.cpp:385-395)
        check(mpt_signal_external_semaphore(ctx, cuda_done->imported, n, stream));
        if (events & MPTN_EVENT_BOUNCES_CHANGED) io.send_event("bounces_changed"); // render_pt.cpp:102-105

        // 4. back to Vulkan: everything submitted from here on waits for S_cu = n; copy the result into the graph's image
        cuda_done->queue_wait(submission.get_queue(), n, vk::PipelineStageFlagBits::eTransfer);
        const ImageHandle& image = io[con_irradiance].get_image();
        submission.get_cmd()->get_command_buffer().copyBufferToImage(out[slot]->buffer, image->get_image(), vk::ImageLayout::eTransferDstOptimal,
                                                                     whole_image(extent, 0)); // transfer_dst: the graph emitted the barrier
        return {};
    }

    // Same keys as RenderPT::properties (render_pt.cpp:147-187). The values live in the C-ABI node so that JSON load / dump and the UI
    // see one source of truth: dump -> edit through `config` -> load.
    NodeStatusFlags properties(Properties& config) override {
        nlohmann::json j = dump_properties();
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
        int format = j["irradiance format"] == "R16G16B16A16Sfloat" ? 1 : 0;
        config.config_options("irradiance format", format, {"R32G32B32A32Sfloat", "R16G16B16A16Sfloat"}, Properties::OptionsStyle::COMBO);
        j["samples per pixel"] = spp;
        j["max path length"] = max_path_length;
        j["emission on primary"] = emission;
        j["shader execution reordering"] = ser;
        j["demodulate albedo"] = demod;
        j["irradiance format"] = format == 1 ? "R16G16B16A16Sfloat" : "R32G32B32A32Sfloat";
        uint32_t status = 0;
        if (mptn_load_config(node, j.dump().c_str(), &status) != MPT_OK) throw graph_errors::node_error{mptn_last_error(node)};
        return (status & MPTN_NEEDS_RECONNECT) ? NodeStatusFlags{NEEDS_RECONNECT} : NodeStatusFlags{};
    }

  private:
    void check(const mpt_status s) {
        if (s != MPT_OK) throw graph_errors::node_error{node ? mptn_last_error(node) : "libmpt error"};
    }

    nlohmann::json dump_properties() {
        size_t len = 0;
        mptn_dump_config(node, nullptr, 0, &len);
        std::string buf(len + 1, '\0');
        mptn_dump_config(node, buf.data(), buf.size(), &len);
        return nlohmann::json::parse(buf.c_str());
    }
    bool demodulate_albedo() { return dump_properties()["demodulate albedo"].get<bool>(); }

    void release_interop() {
        if (stream) cudaStreamSynchronize(stream);
        for (uint32_t slot = 0; slot < kMaxInFlight; slot++) {
            in[slot].reset();
            out[slot].reset();
        }
        vk_done.reset();
        cuda_done.reset();
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
        m.aperture_radius = c.get_aperture_radius(); // > 0: thin lens, the lens sample is keyed on `frame` (scene.slang:91-101)
        m.jitter[0] = c.get_jitter().x;
        m.jitter[1] = c.get_jitter().y;
        return m;
    }

    // Scene tables -> mpt_scene_desc. Uses the getters of adapter/merian_getters.patch; the device vertex / index buffers
    // (bits AFTER transform_vertex.slang) are read back once through a host-visible copy.
    void upload_scene(Scene& scene, Submission& submission) {
        const auto& infos = scene.get_mesh_infos();
        const auto& groups = scene.get_mesh_groups();
        const auto& geometries = scene.get_geometries();
        const auto& tlas = scene.get_tlas_instances();

        // 1. read the device vertex / index buffers back
        struct Range {
            vk::DeviceSize vb_off = 0, vb_size = 0, ib_off = 0, ib_size = 0;
        };
        std::vector<Range> ranges(infos.size());
        vk::DeviceSize total = 0;
        for (size_t i = 0; i < infos.size(); i++) {
            if (!infos[i].mesh) continue;
            const Scene::Mesh& mesh = *infos[i].mesh;
            ranges[i].vb_off = total;
            ranges[i].vb_size = vk::DeviceSize(mesh.get_vertex_count()) * sizeof(mpt_packed_vertex);
            total += ranges[i].vb_size;
            ranges[i].ib_off = total;
            ranges[i].ib_size = mesh.has_indices() ? vk::DeviceSize(mesh.get_primitive_count()) * 3 * index_bytes(mesh.index_type) : 0;
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
        cmd->keep_until_pool_reset(readback);
        mpt_status upload_status = MPT_OK;
        submission.sync_to_cpu_and_back([&] { // submission.hpp:101-114: the GPU waits until this returns
            const uint8_t* base = readback->get_memory()->map_as<uint8_t>();
            std::vector<mpt_mesh_desc> meshes(infos.size());
            for (size_t i = 0; i < infos.size(); i++) {
                if (!infos[i].mesh) continue;
                const Scene::Mesh& mesh = *infos[i].mesh;
                meshes[i].vertices = reinterpret_cast<const mpt_packed_vertex*>(base + ranges[i].vb_off);
                meshes[i].indices = ranges[i].ib_size ? base + ranges[i].ib_off : nullptr;
                meshes[i].vertex_count = mesh.get_vertex_count();
                meshes[i].primitive_count = mesh.get_primitive_count();
                meshes[i].index_flag = !mesh.has_indices()                            ? 0u
                                       : mesh.index_type == vk::IndexType::eUint32    ? uint32_t(MPT_GEOM_INDEX_U32)
                                       : mesh.index_type == vk::IndexType::eUint16    ? uint32_t(MPT_GEOM_INDEX_U16)
                                                                                      : uint32_t(MPT_GEOM_INDEX_U8);
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
                        const vk::AccelerationStructureInstanceKHR& t = tlas[k];
                        std::memcpy(insts[k].transform, &t.transform.matrix, sizeof(float) * 12);
                        insts[k].instance_custom_index = t.instanceCustomIndex;
                        insts[k].mask = t.mask;
                        const auto f = static_cast<vk::GeometryInstanceFlagsKHR>(t.flags);
                        insts[k].flags = ((f & vk::GeometryInstanceFlagBitsKHR::eTriangleFacingCullDisable) ? uint32_t(MPT_INSTANCE_FACING_CULL_DISABLE) : 0u) |
                                         ((f & vk::GeometryInstanceFlagBitsKHR::eTriangleFlipFacing) ? uint32_t(MPT_INSTANCE_FLIP_FACING) : 0u) |
                                         ((f & vk::GeometryInstanceFlagBitsKHR::eForceOpaque) ? uint32_t(MPT_INSTANCE_FORCE_OPAQUE) : 0u);
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
            const MaterialSystemHandle& ms = scene.get_material_system();
            std::vector<mpt_material_type> types(ms->get_material_types().size());
            for (const auto& [name, type_info] : ms->get_material_types()) {
                mpt_material_type t{MPT_MATERIAL_DIFFUSE, 0};
                if (name.rfind("merian::GltfMaterial", 0) == 0) t = {MPT_MATERIAL_GLTF, gltf_features_from_type_name(name)};
                types[type_info.dispatch_id] = t;
            }
            if (upload_status == MPT_OK)
                upload_status = mpt_materials_upload(ctx, ms->get_host_buffer().data(), ms->get_entry_size_bytes(), ms->get_material_count(), types.data(),
                                                     uint32_t(types.size()));
            mpt_materials_set_hints(ctx, ms->get_alpha_test_threshold(), ms->get_min_roughness(), ms->get_clamp_normals() ? 1u : 0u);
            materials_version = ms->version();
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

    static uint32_t index_bytes(const vk::IndexType t) { return t == vk::IndexType::eUint32 ? 4u : (t == vk::IndexType::eUint16 ? 2u : 1u); }

    static uint32_t to_mpt_wrap(const vk::SamplerAddressMode m) {
        return m == vk::SamplerAddressMode::eRepeat ? uint32_t(MPT_WRAP_REPEAT)
               : m == vk::SamplerAddressMode::eMirroredRepeat ? uint32_t(MPT_WRAP_MIRRORED_REPEAT)
                                                               : uint32_t(MPT_WRAP_CLAMP_TO_EDGE);
    }

    // Material textures (RGBA8 sRGB / UNORM, texture_manager.cpp:207-243) and the lat-long env map (RGBA32F, scene.cpp:780-795):
    // level 0 of every image -> one exported buffer -> mpt_texture_upload_device / mpt_env_set_device, device to device. The Vulkan copies
    // and the CUDA uploads are ordered by the two exported timeline semaphores; only the final stream synchronize blocks the host, and
    // only on the iterations where the scene's resources changed.
    void upload_textures_and_env(Scene& scene, Submission& submission) {
        const TextureManagerHandle& textures = scene.get_texture_manager();
        struct Item {
            TextureHandle texture;
            uint32_t id = 0;
            vk::DeviceSize offset = 0;
            bool is_env = false;
        };
        std::vector<Item> items;
        vk::DeviceSize total = 0;
        const auto add = [&](const TextureHandle& t, const uint32_t id, const bool is_env) {
            if (!t) return;
            const vk::Extent3D e = t->get_image()->get_extent();
            items.push_back({t, id, total, is_env});
            total += (vk::DeviceSize(e.width) * e.height * (is_env ? 16 : 4) + 255) & ~vk::DeviceSize(255);
        };
        for (uint32_t id = 0; id < textures->get_capacity(); id++) add(textures->get_texture(id), id, false);
        const auto latlong = std::dynamic_pointer_cast<LatLongEnvMap>(scene.get_env());
        if (latlong) add(latlong->get_texture(), 0, true);
        if (!latlong) check(mpt_env_set(ctx, MPT_ENV_EMPTY, nullptr, 0.f, 0, 0, nullptr)); // EmptyEnv (env_map.hpp:86-99); cube / equal-area: not built
        if (items.empty()) return;

        ExportedBuffer staging(context, ctx, total, vk::BufferUsageFlagBits::eTransferDst);
        const CommandBufferHandle& cmd = submission.get_cmd();
        std::vector<vk::ImageMemoryBarrier2> to_src, back;
        for (const Item& it : items) {
            const ImageHandle& image = it.texture->get_image();
            const vk::ImageLayout old_layout = image->get_current_layout();
            to_src.push_back(image->barrier2(vk::ImageLayout::eTransferSrcOptimal));
            back.push_back(image->barrier2(old_layout)); // the barriers are recorded in this order below, so the tracked state matches
        }
        cmd->barrier(to_src);
        for (const Item& it : items)
            cmd->get_command_buffer().copyImageToBuffer(it.texture->get_image()->get_image(), vk::ImageLayout::eTransferSrcOptimal, staging.buffer,
                                                        whole_image(it.texture->get_image()->get_extent(), it.offset));
        cmd->barrier(back);
        const uint64_t n = ++timeline_value;
        submission.submit();
        vk_done->queue_signal(submission.get_queue(), n);
        check(mpt_wait_external_semaphore(ctx, vk_done->imported, n, stream));
        for (const Item& it : items) {
            const ImageHandle& image = it.texture->get_image();
            const vk::Extent3D e = image->get_extent();
            const char* d_pixels = static_cast<const char*>(staging.d_ptr) + it.offset;
            if (it.is_env) {
                const EnvMapHandle& env = scene.get_env();
                const float4x4 to_local = env->get_transform();
                check(mpt_env_set_device(ctx, MPT_ENV_LATLONG, reinterpret_cast<const float*>(&to_local), env->get_intensity(), e.width, e.height, d_pixels));
            } else {
                const vk::SamplerCreateInfo& s = it.texture->get_sampler()->get_create_info(); // getter: adapter/merian_getters.patch
                const bool srgb = image->get_format() == vk::Format::eR8G8B8A8Srgb;
                const bool mips = image->get_mip_levels() > 1; // the chain is rebuilt the way cmd_generate_mipmaps fills it (include/mpt.h)
                check(mpt_texture_upload_device(ctx, it.id, e.width, e.height, srgb ? MPT_TEX_RGBA8_SRGB : MPT_TEX_RGBA8_UNORM, to_mpt_wrap(s.addressModeU),
                                                to_mpt_wrap(s.addressModeV), s.magFilter == vk::Filter::eNearest ? MPT_FILTER_NEAREST : MPT_FILTER_LINEAR,
                                                !mips                                                 ? uint32_t(MPT_MIPMAP_NONE)
                                                : s.mipmapMode == vk::SamplerMipmapMode::eNearest     ? uint32_t(MPT_MIPMAP_NEAREST)
                                                                                                      : uint32_t(MPT_MIPMAP_LINEAR),
                                                d_pixels, 0, stream));
            }
        }
        check(mpt_signal_external_semaphore(ctx, cuda_done->imported, n, stream));
        cuda_done->queue_wait(submission.get_queue(), n, vk::PipelineStageFlagBits::eTopOfPipe);
        cudaStreamSynchronize(stream); // `staging` is released on return
    }

    ContextHandle context;
    ResourceAllocatorHandle allocator;
    PtrInHandle<Scene> con_scene = PtrIn<Scene>::create();
    ShaderObjectInHandle<GBufferObject> con_gbuffer = ShaderObjectIn<GBufferObject>::create();
    ManagedVkImageOutHandle con_irradiance;
    vk::Extent3D extent{};
    uint32_t texel_bytes = 16;
    std::array<std::unique_ptr<ExportedBuffer>, kMaxInFlight> in, out;
    std::unique_ptr<ExportedTimeline> vk_done, cuda_done; // S_vk: Vulkan finished writing `in`; S_cu: CUDA finished writing `out`
    uint64_t timeline_value = 0;
    mpt_context* ctx = nullptr;
    mptn_node* node = nullptr;
    cudaStream_t stream = nullptr;
    bool tables_dirty = true;
    uint64_t materials_version = ~uint64_t(0);
};

} // namespace merian

extern "C" {

MERIAN_PLUGIN_EXPORT const char* merian_plugin_name(void) { return "pt-cuda"; }

MERIAN_PLUGIN_EXPORT uint32_t merian_node_plugin_abi_version(void) { return MERIAN_NODE_PLUGIN_ABI_VERSION; }

MERIAN_PLUGIN_EXPORT void merian_register_nodes(merian::NodeRegistry& registry) {
    // "Render (Path-traced)" itself is registered built-in and a duplicate name throws (node_registry.hpp:60-63), so the swap happens in
    // the graph JSON: adapter/pt_cuda.json is examples/renderers/pt.json with this type name; node id `render` and all connections stay.
    registry.register_node_type<merian::RenderPTCuda>("Render (Path-traced, CUDA)", "Path tracer on hand-written sm_100a CUDA kernels (libmpt.so).");
}

} // extern "C"
