// Stub of the merian declarations adapter/merian_plugin_pt_cuda.cpp uses. TEST INFRASTRUCTURE: tests/test_adapter_syntax.py type-checks the
// adapter against it (g++ -std=c++20 -fsyntax-only) in an image that has neither merian's dependencies nor the Vulkan SDK.
// Every declaration restates one of the reference; the `@ref file:` tags name the header (relative to /root/reference/include) and the
// text the test looks for there when the reference is present, so a stub that drifts from the reference fails the test.
// `@patch` marks the getters adapter/merian_getters.patch adds (they are looked up in the patch instead).
#pragma once
#include <vulkan/vulkan.hpp>

#include <any>
#include <cstdint>
#include <functional>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#define MERIAN_PLUGIN_EXPORT __attribute__((visibility("default"))) // @ref merian/plugin/plugin_export.hpp: MERIAN_PLUGIN_EXPORT
#define MERIAN_NODE_PLUGIN_ABI_VERSION 1u                           // @ref merian-graph/plugin/node_plugin.hpp: #define MERIAN_NODE_PLUGIN_ABI_VERSION 1u

namespace merian {

// ---- math (merian/utils/vector_matrix.hpp: HLSL-style vectors) ---------------------------------------------------------------------
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4x4 { float m[16]; };
inline float3 operator*(const float3& a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float3 cross(const float3& a, const float3& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// ---- Vulkan core wrappers -------------------------------------------------------------------------------------------------------------
class VulkanProperties {
  public:
    const vk::PhysicalDeviceIDProperties& get_id_properties() const; // @ref merian/vk/utils/vulkan_properties.hpp: const vk::PhysicalDeviceIDProperties& get_id_properties() const;
};
class PhysicalDevice {
  public:
    const VulkanProperties& get_properties() const;                             // @ref merian/vk/physical_device.hpp: const VulkanProperties& get_properties() const {
    const vk::PhysicalDeviceMemoryProperties2& get_memory_properties() const;   // @ref merian/vk/physical_device.hpp: const vk::PhysicalDeviceMemoryProperties2& get_memory_properties() const {
};
using PhysicalDeviceHandle = std::shared_ptr<PhysicalDevice>;
class Device {
  public:
    const vk::Device& get_device() const; // @ref merian/vk/device.hpp: const vk::Device& get_device() const {
};
using DeviceHandle = std::shared_ptr<Device>;
class Context {
  public:
    const PhysicalDeviceHandle& get_physical_device() const; // @ref merian/vk/context.hpp: const PhysicalDeviceHandle& get_physical_device() const;
    const DeviceHandle& get_device() const;                  // @ref merian/vk/context.hpp: const DeviceHandle& get_device() const;
};
using ContextHandle = std::shared_ptr<Context>;

class Object {};
using ConstObjectHandle = std::shared_ptr<const Object>;
class MemoryAllocation {
  public:
    template <class T> T* map_as() { return (T*)map(); } // @ref merian/vk/memory/memory_allocator.hpp: template <class T> T* map_as() {
    virtual void* map();                                 // @ref merian/vk/memory/memory_allocator.hpp: virtual void* map() {
    virtual void unmap();                                // @ref merian/vk/memory/memory_allocator.hpp: virtual void unmap() {
    virtual ~MemoryAllocation() = default;
};
using MemoryAllocationHandle = std::shared_ptr<MemoryAllocation>;
enum class MemoryMappingType { NONE, HOST_ACCESS_RANDOM, HOST_ACCESS_SEQUENTIAL_WRITE }; // @ref merian/vk/memory/memory_allocator.hpp: HOST_ACCESS_RANDOM
class Buffer : public Object {
  public:
    const MemoryAllocationHandle& get_memory() const; // @ref merian/vk/memory/resource_allocations.hpp: const MemoryAllocationHandle& get_memory() const {
};
using BufferHandle = std::shared_ptr<Buffer>;
class Image : public Object {
  public:
    const vk::Image& get_image() const;           // @ref merian/vk/memory/resource_allocations.hpp: const vk::Image& get_image() const {
    const vk::Extent3D& get_extent() const;       // @ref merian/vk/memory/resource_allocations.hpp: const vk::Extent3D& get_extent() const {
    const vk::Format& get_format() const;         // @ref merian/vk/memory/resource_allocations.hpp: const vk::Format& get_format() const {
    uint32_t get_mip_levels() const;              // @ref merian/vk/memory/resource_allocations.hpp: uint32_t get_mip_levels() const {
    const vk::ImageLayout& get_current_layout() const; // @ref merian/vk/memory/resource_allocations.hpp: const vk::ImageLayout& get_current_layout() const {
    [[nodiscard]] vk::ImageMemoryBarrier2 barrier2(const vk::ImageLayout new_layout, const bool transition_from_undefined = false); // @ref merian/vk/memory/resource_allocations.hpp: [[nodiscard]] vk::ImageMemoryBarrier2 barrier2(const vk::ImageLayout new_layout,
};
using ImageHandle = std::shared_ptr<Image>;
class Sampler {
  public:
    const vk::SamplerCreateInfo& get_create_info() const; // @patch get_create_info
};
using SamplerHandle = std::shared_ptr<Sampler>;
class Texture {
  public:
    const ImageHandle& get_image() const;     // @ref merian/vk/memory/resource_allocations.hpp: const ImageHandle& get_image() const {
    const SamplerHandle& get_sampler() const; // @ref merian/vk/memory/resource_allocations.hpp: const SamplerHandle& get_sampler() const {
};
using TextureHandle = std::shared_ptr<Texture>;

class ResourceAllocator {
  public:
    BufferHandle create_buffer(const vk::DeviceSize size_, const vk::BufferUsageFlags usage_ = {}, const MemoryMappingType mapping_type = MemoryMappingType::NONE,
                               const std::string& debug_name = {}, const std::optional<vk::DeviceSize> min_alignment = std::nullopt); // @ref merian/vk/memory/resource_allocator.hpp: BufferHandle create_buffer(const vk::DeviceSize size_,
};
using ResourceAllocatorHandle = std::shared_ptr<ResourceAllocator>;

class CommandBuffer {
  public:
    const vk::CommandBuffer& get_command_buffer() const noexcept; // @ref merian/vk/command/command_buffer.hpp: const vk::CommandBuffer& get_command_buffer() const noexcept {
    void keep_until_pool_reset(const ConstObjectHandle& object);  // @ref merian/vk/command/command_buffer.hpp: void keep_until_pool_reset(const ConstObjectHandle& object) {
    void copy(const BufferHandle& src_buffer, const BufferHandle& dst_buffer, const vk::ArrayProxy<const vk::BufferCopy>& regions); // @ref merian/vk/command/command_buffer.hpp: void copy(const BufferHandle& src_buffer,
    void barrier(const vk::ArrayProxy<const vk::ImageMemoryBarrier2>& image_memory_barriers); // @ref merian/vk/command/command_buffer.hpp: void barrier(const vk::ArrayProxy<const vk::ImageMemoryBarrier2>& image_memory_barriers) {
};
using CommandBufferHandle = std::shared_ptr<CommandBuffer>;
class Queue {
  public:
    void submit(const vk::ArrayProxy<vk::SubmitInfo>& submit_infos, vk::Fence fence = VK_NULL_HANDLE); // @ref merian/vk/command/queue.hpp: void submit(const vk::ArrayProxy<vk::SubmitInfo>& submit_infos,
};
using QueueHandle = std::shared_ptr<Queue>;
class Submission {
  public:
    const QueueHandle& get_queue() const noexcept;                        // @ref merian/vk/command/submission.hpp: const QueueHandle& get_queue() const noexcept {
    const CommandBufferHandle& get_cmd();                                 // @ref merian/vk/command/submission.hpp: const CommandBufferHandle& get_cmd() {
    void sync_to_cpu_and_back(const std::function<void()>& callback);     // @ref merian/vk/command/submission.hpp: void sync_to_cpu_and_back(const std::function<void()>& callback) {
    void submit(const vk::Fence& fence = VK_NULL_HANDLE);                 // @ref merian/vk/command/submission.hpp: void submit(const vk::Fence& fence = VK_NULL_HANDLE) {
};

// ---- device support ---------------------------------------------------------------------------------------------------------------------
struct DeviceSupportQueryInfo;
struct DeviceSupportInfo {
    bool supported = true;
    static DeviceSupportInfo check(const DeviceSupportQueryInfo& query_info, const std::vector<const char*>& required_features = {},
                                   const std::vector<const char*>& optional_features = {}, const std::vector<const char*>& required_extensions = {},
                                   const std::vector<const char*>& optional_extensions = {}); // @ref merian/vk/extension/extension.hpp: static DeviceSupportInfo check(const DeviceSupportQueryInfo& query_info,
};

// ---- properties -------------------------------------------------------------------------------------------------------------------------
class Properties {
  public:
    enum class OptionsStyle { DONT_CARE, RADIO_BUTTON, COMBO, LIST_BOX }; // @ref merian/utils/properties.hpp: enum class OptionsStyle {
    virtual ~Properties() = default;
    virtual void st_separate(const std::string& label = "") = 0; // @ref merian/utils/properties.hpp: virtual void st_separate(const std::string& label = "") = 0;
    virtual void st_no_space() = 0;                              // @ref merian/utils/properties.hpp: virtual void st_no_space() = 0;
    bool config_int(const std::string& id, int32_t& value, const std::string& desc = "", const std::optional<int32_t>& min = std::nullopt,
                    const std::optional<int32_t>& max = std::nullopt); // @ref merian/utils/properties.hpp: bool config_int(const std::string& id,
    virtual bool config_bool(const std::string& id, bool& value, const std::string& desc = "") = 0; // @ref merian/utils/properties.hpp: virtual bool config_bool(const std::string& id, bool& value, const std::string& desc = "") = 0;
    virtual bool config_options(const std::string& id, int& selected, const std::vector<std::string>& options, const OptionsStyle style = OptionsStyle::DONT_CARE,
                                const std::string& desc = "") = 0; // @ref merian/utils/properties.hpp: virtual bool config_options(const std::string& id,
};

// ---- graph ------------------------------------------------------------------------------------------------------------------------------
namespace graph_errors {
class node_error : public std::runtime_error { // @ref merian-graph/graph/errors.hpp: class node_error
  public:
    node_error(const std::string& reason) : std::runtime_error(reason) {}
};
} // namespace graph_errors

struct ConnectorAccess { // @ref merian-graph/graph/connector_access.hpp: struct ConnectorAccess {
    vk::PipelineStageFlags2 stages{};
    vk::AccessFlags2 access{};
    vk::ImageUsageFlags image_usage{};
    static const ConnectorAccess transfer_src;      // @ref merian-graph/graph/connector_access.hpp: static const ConnectorAccess transfer_src;
    static const ConnectorAccess transfer_dst;      // @ref merian-graph/graph/connector_access.hpp: static const ConnectorAccess transfer_dst;
    static const ConnectorAccess ray_tracing_read;  // @ref merian-graph/graph/connector_access.hpp: static const ConnectorAccess ray_tracing_read;
};
class InputConnector { public: virtual ~InputConnector() = default; };
class OutputConnector { public: virtual ~OutputConnector() = default; };
using InputConnectorHandle = std::shared_ptr<InputConnector>;
using OutputConnectorHandle = std::shared_ptr<OutputConnector>;
struct InputConnectorDescriptor { // @ref merian-graph/graph/node.hpp: struct InputConnectorDescriptor {
    std::string name;
    InputConnectorHandle connector;
    ConnectorAccess access{};
    uint32_t delay = 0;
    bool optional = false;
};
struct OutputConnectorDescriptor { // @ref merian-graph/graph/node.hpp: struct OutputConnectorDescriptor {
    std::string name;
    OutputConnectorHandle connector;
    ConnectorAccess access{};
    bool disabled = false;
};

// PtrIn<T>: io[con] is the transported std::shared_ptr<T> (merian-graph/connectors/ptr_in.hpp)
template <typename T> class PtrIn : public InputConnector {
  public:
    using resource_access_type = const std::shared_ptr<T>&;
    static std::shared_ptr<PtrIn<T>> create(const uint32_t delay = 0); // @ref merian-graph/connectors/ptr_in.hpp: static PtrInHandle<T> create(
};
template <typename T> using PtrInHandle = std::shared_ptr<PtrIn<T>>;

// ShaderObjectIn<T>: io[con] is a ShaderObjectAccess<T> (operator-> reaches the T), io_layout[con] the ShaderObjectOut<T>
template <typename T> class ShaderObjectAccess { // @ref merian-graph/connectors/shader_object_out.hpp: template <typename T> class ShaderObjectAccess {
  public:
    T* operator->() const; // @ref merian-graph/connectors/shader_object_out.hpp: T* operator->() const {
};
template <typename T> class ShaderObjectOut : public OutputConnector {
  public:
    const typename T::CreateInfo& get_create_info() const; // @ref merian-graph/connectors/shader_object_out.hpp: get_create_info() const
};
template <typename T> using ShaderObjectOutHandle = std::shared_ptr<ShaderObjectOut<T>>;
template <typename T> class ShaderObjectIn : public InputConnector {
  public:
    using resource_access_type = ShaderObjectAccess<T>;
    using output_connector_type = ShaderObjectOutHandle<T>;
    static std::shared_ptr<ShaderObjectIn<T>> create(); // @ref merian-graph/connectors/shader_object_in.hpp: static ShaderObjectInHandle<T> create() {
};
template <typename T> using ShaderObjectInHandle = std::shared_ptr<ShaderObjectIn<T>>;

class ImageArrayResource {
  public:
    virtual const ImageHandle& get_image(const uint32_t index = 0) const = 0; // @ref merian-graph/resources/image_array_resource.hpp: virtual const merian::ImageHandle& get_image(const uint32_t index = 0) const = 0;
    virtual ~ImageArrayResource() = default;
};
class ManagedVkImageOut : public OutputConnector {
  public:
    using resource_access_type = const ImageArrayResource&;
    static std::shared_ptr<ManagedVkImageOut> create(const vk::Format format, const vk::Extent3D extent, const bool persistent = false); // @ref merian-graph/connectors/image/vk_image_out_managed.hpp: create(const vk::Format format, const vk::Extent3D extent, const bool persistent = false);
};
using ManagedVkImageOutHandle = std::shared_ptr<ManagedVkImageOut>;

class GBufferObject { // @ref merian-graph/objects/gbuffer_object.hpp: class GBufferObject : public GraphShaderObject {
  public:
    struct CreateInfo {
        vk::Extent3D extent;
    };
    const ImageHandle& get_image(uint32_t index) const; // @patch get_image
};

class NodeIOLayout {
  public:
    template <typename T> typename T::output_connector_type operator[](const std::shared_ptr<T>& input_connector) const; // @ref merian-graph/graph/node_io.hpp: OutputConnectorType operator[](const std::shared_ptr<T>& input_connector) const {
};
class NodeIO {
  public:
    template <typename T> typename T::resource_access_type operator[](const std::shared_ptr<T>& connector) const; // @ref merian-graph/graph/node_io.hpp: ResourceAccessType operator[](const std::shared_ptr<T>& input_connector) const {
    void send_event(const std::string& event_name, const std::any& event_data = {}, const bool notify_all = true) const; // @ref merian-graph/graph/node_io.hpp: void send_event(const std::string& event_name,
};
class NodeProcessInfo {
  public:
    const uint64_t& get_total_iteration() const noexcept; // @ref merian-graph/graph/node_process_info.hpp: const uint64_t& get_total_iteration() const noexcept {
    const uint32_t& get_in_flight_index() const noexcept; // @ref merian-graph/graph/node_process_info.hpp: const uint32_t& get_in_flight_index() const noexcept {
};
struct NodeConnectionInfo {};

class Node : public std::enable_shared_from_this<Node> { // @ref merian-graph/graph/node.hpp: class Node : public std::enable_shared_from_this<Node> {
  public:
    using NodeStatusFlags = uint32_t; // @ref merian-graph/graph/node.hpp: using NodeStatusFlags = uint32_t;
    enum NodeStatusFlagBits { NEEDS_RECONNECT = 0b1, RESET_IN_FLIGHT_DATA = 0b10, REMOVE_NODE = 0b100 }; // @ref merian-graph/graph/node.hpp: NEEDS_RECONNECT = 0b1,
    virtual ~Node() {}
    virtual DeviceSupportInfo query_device_support(const DeviceSupportQueryInfo& query_info); // @ref merian-graph/graph/node.hpp: virtual DeviceSupportInfo query_device_support(const DeviceSupportQueryInfo& /*query_info*/) {
    virtual void initialize(const ContextHandle& context, const ResourceAllocatorHandle& allocator); // @ref merian-graph/graph/node.hpp: virtual void initialize(const ContextHandle& /*context*/,
    virtual std::vector<InputConnectorDescriptor> describe_inputs(); // @ref merian-graph/graph/node.hpp: virtual std::vector<InputConnectorDescriptor> describe_inputs() {
    virtual std::vector<OutputConnectorDescriptor> describe_outputs(const NodeIOLayout& io_layout); // @ref merian-graph/graph/node.hpp: describe_outputs([[maybe_unused]] const NodeIOLayout& io_layout) {
    virtual NodeStatusFlags on_connected(const NodeIOLayout& io_layout, const NodeIO& io, const NodeConnectionInfo& info, Submission& submission); // @ref merian-graph/graph/node.hpp: virtual NodeStatusFlags on_connected([[maybe_unused]] const NodeIOLayout& io_layout,
    virtual NodeStatusFlags process(const NodeIO& io, const NodeProcessInfo& info, Submission& submission); // @ref merian-graph/graph/node.hpp: virtual NodeStatusFlags process([[maybe_unused]] const NodeIO& io,
    virtual NodeStatusFlags properties(Properties& props); // @ref merian-graph/graph/node.hpp: virtual NodeStatusFlags properties([[maybe_unused]] Properties& props) {
};

class NodeRegistry {
  public:
    template <typename NODE_TYPE> void register_node_type(const std::string& node_type_name, const std::string& description, const bool add_default_config = true); // @ref merian-graph/graph/node_registry.hpp: void register_node_type(const std::string& node_type_name,
};

// ---- scene ------------------------------------------------------------------------------------------------------------------------------
using MeshID = uint32_t;      // @ref merian-shaders/scene/scene.hpp: MeshID
using NodeID = uint32_t;      // @ref merian-shaders/scene/scene.hpp: NodeID
using MaterialID = uint32_t;
using TextureID = uint32_t;

class Camera {
  public:
    const float3& get_position() const noexcept;        // @ref merian/utils/camera/camera.hpp: const float3& get_position() const noexcept;
    const float3& get_target() const noexcept;          // @ref merian/utils/camera/camera.hpp: const float3& get_target() const noexcept;
    const float3& get_up() const noexcept;              // @ref merian/utils/camera/camera.hpp: const float3& get_up() const noexcept;
    float get_field_of_view_vertical() const noexcept;  // @ref merian/utils/camera/camera.hpp: float get_field_of_view_vertical() const noexcept;
    float get_aspect_ratio() const noexcept;            // @ref merian/utils/camera/camera.hpp: float get_aspect_ratio() const noexcept;
    float get_near_plane() const noexcept;              // @ref merian/utils/camera/camera.hpp: float get_near_plane() const noexcept;
    float get_far_plane() const noexcept;               // @ref merian/utils/camera/camera.hpp: float get_far_plane() const noexcept;
    float get_aperture_radius() const noexcept;         // @ref merian/utils/camera/camera.hpp: float get_aperture_radius() const noexcept;
    float get_focus_distance() noexcept;                // @ref merian/utils/camera/camera.hpp: float get_focus_distance() noexcept;
    const float3& get_forward() noexcept;               // @ref merian/utils/camera/camera.hpp: const float3& get_forward() noexcept;
    const float3& get_right() noexcept;                 // @ref merian/utils/camera/camera.hpp: const float3& get_right() noexcept;
    const float2& get_jitter() const noexcept;          // @ref merian/utils/camera/camera.hpp: const float2& get_jitter() const noexcept;
};
using CameraHandle = std::shared_ptr<Camera>;

class EnvMap { // @ref merian-shaders/scene/env_map.hpp: class EnvMap {
  public:
    virtual ~EnvMap() = default;
    float get_intensity() const;           // @ref merian-shaders/scene/env_map.hpp: float get_intensity() const {
    const float4x4& get_transform() const; // @ref merian-shaders/scene/env_map.hpp: const float4x4& get_transform() const {
};
using EnvMapHandle = std::shared_ptr<EnvMap>;
class LatLongEnvMap : public EnvMap { // @ref merian-shaders/scene/env_map.hpp: class LatLongEnvMap : public EnvMap {
  public:
    const TextureHandle& get_texture() const; // @ref merian-shaders/scene/env_map.hpp: const TextureHandle& get_texture() const {
};

class TextureManager {
  public:
    const TextureHandle& get_texture(TextureID id) const; // @ref merian-shaders/utils/texture_manager.hpp: const TextureHandle& get_texture(TextureID id) const {
    uint32_t get_capacity() const;                        // @ref merian-shaders/utils/texture_manager.hpp: uint32_t get_capacity() const {
};
using TextureManagerHandle = std::shared_ptr<TextureManager>;

class MaterialSystem {
  public:
    struct MaterialTypeInfo { // @ref merian-shaders/shading/materials/material_system.hpp: struct MaterialTypeInfo {
        uint32_t dispatch_id; // @ref merian-shaders/shading/materials/material_system.hpp: MaterialModelID dispatch_id;
    };
    uint32_t get_material_count() const;        // @ref merian-shaders/shading/materials/material_system.hpp: uint32_t get_material_count() const {
    float get_alpha_test_threshold() const;     // @ref merian-shaders/shading/materials/material_system.hpp: float get_alpha_test_threshold() const {
    bool get_clamp_normals() const;             // @ref merian-shaders/shading/materials/material_system.hpp: bool get_clamp_normals() const {
    float get_min_roughness() const;            // @ref merian-shaders/shading/materials/material_system.hpp: float get_min_roughness() const {
    uint64_t version() const;                   // @ref merian-shaders/shading/materials/material_system.hpp: uint64_t version() const {
    const std::vector<uint8_t>& get_host_buffer();                                          // @patch get_host_buffer
    uint32_t get_entry_size_bytes() const;                                                  // @patch get_entry_size_bytes
    const std::unordered_map<std::string, MaterialTypeInfo>& get_material_types() const;   // @patch get_material_types
};
using MaterialSystemHandle = std::shared_ptr<MaterialSystem>;

struct GeometryData { // scene-data.slangh:119-129 as the host mirrors it
    uint64_t vertex_buffer, prev_vertex_buffer, index_buffer;
    uint16_t material_id, flags;
    uint32_t primitive_count;
};

class Scene {
  public:
    class Mesh { // @ref merian-shaders/scene/scene.hpp: class Mesh {
      public:
        vk::IndexType index_type = vk::IndexType::eUint32;   // @ref merian-shaders/scene/scene.hpp: vk::IndexType index_type = vk::IndexType::eUint32;
        virtual ~Mesh() = default;
        virtual uint32_t get_vertex_count() const = 0;       // @ref merian-shaders/scene/scene.hpp: virtual uint32_t get_vertex_count() const = 0;
        virtual uint32_t get_primitive_count() const = 0;    // @ref merian-shaders/scene/scene.hpp: virtual uint32_t get_primitive_count() const = 0;
        bool is_opaque() const;                              // @ref merian-shaders/scene/scene.hpp: bool is_opaque() const {
        bool has_indices() const;                            // @ref merian-shaders/scene/scene.hpp: bool has_indices() const {
    };
    using MeshHandle = std::shared_ptr<Mesh>;
    struct MeshBufferRegion { // @ref merian-shaders/scene/scene.hpp: struct MeshBufferRegion {
        vk::DeviceSize size = 0;
        vk::DeviceSize get_offset() const;  // @ref merian-shaders/scene/scene.hpp: vk::DeviceSize get_offset() const {
        BufferHandle get_buffer() const;    // @patch get_buffer
    };
    class MeshInfo { // @ref merian-shaders/scene/scene.hpp: class MeshInfo {
      public:
        MeshHandle mesh;
        MeshBufferRegion vertex_buffer;
        MeshBufferRegion index_buffer;
    };
    struct MeshGroup {
        std::vector<MeshID> meshes;                                                            // @ref merian-shaders/scene/scene.hpp: meshes;
        const std::vector<NodeID>& get_instances(const std::vector<MeshInfo>& mesh_infos) const; // @ref merian-shaders/scene/scene.hpp: get_instances(const std::vector<MeshInfo>& mesh_infos) const {
    };
    struct UpdateChanges { // @ref merian-shaders/scene/scene.hpp: struct UpdateChanges {
        bool geometry_changed = false;
        bool transform_changed = false;
        bool camera_changed = false;
    };
    virtual ~Scene() = default;
    virtual bool is_ready() const;                              // @ref merian-shaders/scene/scene.hpp: virtual bool is_ready() const {
    const UpdateChanges& get_last_update_changes() const;       // @ref merian-shaders/scene/scene.hpp: const UpdateChanges& get_last_update_changes() const {
    const MaterialSystemHandle& get_material_system() const;    // @ref merian-shaders/scene/scene.hpp: const MaterialSystemHandle& get_material_system() const {
    const EnvMapHandle& get_env() const;                        // @ref merian-shaders/scene/scene.hpp: const EnvMapHandle& get_env() const {
    const TextureManagerHandle& get_texture_manager() const;    // @ref merian-shaders/scene/scene.hpp: const TextureManagerHandle& get_texture_manager() const {
    CameraHandle get_active_camera() const;                     // @ref merian-shaders/scene/scene.hpp: CameraHandle get_active_camera() const;
    const std::vector<MeshInfo>& get_mesh_infos() const;        // @ref merian-shaders/scene/scene.hpp: const std::vector<MeshInfo>& get_mesh_infos() const {
    const std::vector<MeshGroup>& get_mesh_groups() const;                                                      // @patch get_mesh_groups
    const std::vector<GeometryData>& get_geometries() const;                                                    // @patch get_geometries
    const std::vector<vk::AccelerationStructureInstanceKHR>& get_tlas_instances() const;                        // @patch get_tlas_instances
    const std::vector<float4x4>& get_instance_transforms() const;                                               // @patch get_instance_transforms
    const std::vector<float4x4>& get_inverse_transposed_instance_transforms() const;                            // @patch get_inverse_transposed_instance_transforms
};
using SceneHandle = std::shared_ptr<Scene>;

} // namespace merian
