// Stub of the part of Vulkan-Hpp (vulkan/vulkan.hpp, Khronos) that adapter/merian_plugin_pt_cuda.cpp uses. TEST INFRASTRUCTURE: it exists
// so that tests/test_adapter_syntax.py can type-check the adapter in an image without the Vulkan SDK. Struct constructors take their
// members in declaration order followed by pNext, as the generated vulkan_structs.hpp does; only members the adapter touches exist.
#pragma once
#include <array>
#include <cstdint>
#include <initializer_list>
#include <vector>

#define VK_NULL_HANDLE nullptr
#define VK_QUEUE_FAMILY_IGNORED (~0u)

namespace vk {

using DeviceSize = uint64_t;
using DeviceAddress = uint64_t;

template <typename BitType> class Flags {
  public:
    constexpr Flags() = default;
    constexpr Flags(BitType bit) : mask(static_cast<uint32_t>(bit)) {}
    constexpr explicit Flags(uint32_t flags) : mask(flags) {} // VULKAN_HPP_CONSTEXPR explicit Flags( MaskType flags )
    constexpr Flags operator|(Flags o) const { Flags r; r.mask = mask | o.mask; return r; }
    constexpr Flags operator&(Flags o) const { Flags r; r.mask = mask & o.mask; return r; }
    constexpr explicit operator bool() const { return mask != 0; }
    constexpr bool operator==(Flags o) const { return mask == o.mask; }
    uint32_t mask = 0;
};
template <typename BitType> constexpr Flags<BitType> operator&(BitType a, Flags<BitType> b) { return Flags<BitType>(a) & b; }

template <typename T> class ArrayProxy {
  public:
    ArrayProxy() = default;
    ArrayProxy(std::nullptr_t) {}
    ArrayProxy(T& v) : n(1), p(&v) {}
    ArrayProxy(const std::initializer_list<std::remove_const_t<T>>& l) : n(uint32_t(l.size())), p(l.begin()) {}
    template <typename V> ArrayProxy(V& v) : n(uint32_t(v.size())), p(v.data()) {}
    uint32_t size() const { return n; }
    const T* data() const { return p; }
  private:
    uint32_t n = 0;
    const T* p = nullptr;
};

#define MPT_STUB_HANDLE(Name)                                  \
    class Name {                                               \
      public:                                                  \
        Name() = default;                                      \
        Name(std::nullptr_t) {}                                \
        explicit operator bool() const { return h != nullptr; } \
        Name& operator=(std::nullptr_t) { h = nullptr; return *this; } \
      private:                                                 \
        void* h = nullptr;                                     \
    };
MPT_STUB_HANDLE(Buffer)
MPT_STUB_HANDLE(DeviceMemory)
MPT_STUB_HANDLE(Semaphore)
MPT_STUB_HANDLE(Fence)
MPT_STUB_HANDLE(Image)
MPT_STUB_HANDLE(ImageView)
MPT_STUB_HANDLE(Sampler)
#undef MPT_STUB_HANDLE

enum class Format { eUndefined = 0, eR8G8B8A8Unorm = 37, eR8G8B8A8Srgb = 43, eR16G16B16A16Sfloat = 97, eR32G32Uint = 101, eR32G32B32A32Uint = 107, eR32G32B32A32Sfloat = 109 };
enum class ImageLayout { eUndefined = 0, eGeneral = 1, eShaderReadOnlyOptimal = 5, eTransferSrcOptimal = 6, eTransferDstOptimal = 7 };
enum class SharingMode { eExclusive = 0, eConcurrent = 1 };
enum class IndexType { eUint16 = 0, eUint32 = 1, eNoneKHR = 1000165000, eUint8 = 1000265000 };
enum class SemaphoreType { eBinary = 0, eTimeline = 1 };
enum class SamplerAddressMode { eRepeat = 0, eMirroredRepeat = 1, eClampToEdge = 2, eClampToBorder = 3 };
enum class Filter { eNearest = 0, eLinear = 1 };
enum class SamplerMipmapMode { eNearest = 0, eLinear = 1 };

enum class MemoryPropertyFlagBits : uint32_t { eDeviceLocal = 0x1, eHostVisible = 0x2 };
using MemoryPropertyFlags = Flags<MemoryPropertyFlagBits>;
enum class BufferUsageFlagBits : uint32_t { eTransferSrc = 0x1, eTransferDst = 0x2, eStorageBuffer = 0x20 };
using BufferUsageFlags = Flags<BufferUsageFlagBits>;
enum class BufferCreateFlagBits : uint32_t {};
using BufferCreateFlags = Flags<BufferCreateFlagBits>;
enum class ImageUsageFlagBits : uint32_t { eTransferSrc = 0x1, eTransferDst = 0x2, eSampled = 0x4, eStorage = 0x8 };
using ImageUsageFlags = Flags<ImageUsageFlagBits>;
enum class ImageAspectFlagBits : uint32_t { eColor = 0x1 };
using ImageAspectFlags = Flags<ImageAspectFlagBits>;
enum class PipelineStageFlagBits : uint32_t { eTopOfPipe = 0x1, eTransfer = 0x1000, eAllCommands = 0x10000 };
using PipelineStageFlags = Flags<PipelineStageFlagBits>;
enum class PipelineStageFlagBits2 : uint64_t { eAllCommands = 0x10000 };
using PipelineStageFlags2 = Flags<PipelineStageFlagBits2>;
enum class AccessFlagBits2 : uint64_t { eMemoryRead = 0x8000, eMemoryWrite = 0x10000 };
using AccessFlags2 = Flags<AccessFlagBits2>;
enum class ExternalMemoryHandleTypeFlagBits : uint32_t { eOpaqueFd = 0x1 };
using ExternalMemoryHandleTypeFlags = Flags<ExternalMemoryHandleTypeFlagBits>;
enum class ExternalSemaphoreHandleTypeFlagBits : uint32_t { eOpaqueFd = 0x1 };
using ExternalSemaphoreHandleTypeFlags = Flags<ExternalSemaphoreHandleTypeFlagBits>;
enum class SemaphoreCreateFlagBits : uint32_t {};
using SemaphoreCreateFlags = Flags<SemaphoreCreateFlagBits>;
enum class GeometryInstanceFlagBitsKHR : uint32_t { eTriangleFacingCullDisable = 0x1, eTriangleFlipFacing = 0x2, eForceOpaque = 0x4, eForceNoOpaque = 0x8 };
using GeometryInstanceFlagsKHR = Flags<GeometryInstanceFlagBitsKHR>;

struct Extent3D { uint32_t width = 0, height = 0, depth = 0; };
struct Offset3D { int32_t x = 0, y = 0, z = 0; };
struct MemoryType { MemoryPropertyFlags propertyFlags; uint32_t heapIndex = 0; };
struct PhysicalDeviceMemoryProperties { uint32_t memoryTypeCount = 0; std::array<MemoryType, 32> memoryTypes; };
struct PhysicalDeviceMemoryProperties2 { const void* pNext = nullptr; PhysicalDeviceMemoryProperties memoryProperties; };
struct PhysicalDeviceIDProperties { const void* pNext = nullptr; std::array<uint8_t, 16> deviceUUID; };
struct MemoryRequirements { DeviceSize size = 0, alignment = 0; uint32_t memoryTypeBits = 0; };

struct ExternalMemoryBufferCreateInfo {
    ExternalMemoryBufferCreateInfo(ExternalMemoryHandleTypeFlags handleTypes_ = {}, const void* pNext_ = nullptr) : pNext(pNext_), handleTypes(handleTypes_) {}
    const void* pNext;
    ExternalMemoryHandleTypeFlags handleTypes;
};
struct BufferCreateInfo {
    BufferCreateInfo(BufferCreateFlags flags_ = {}, DeviceSize size_ = 0, BufferUsageFlags usage_ = {}, SharingMode sharingMode_ = SharingMode::eExclusive,
                     uint32_t queueFamilyIndexCount_ = 0, const uint32_t* pQueueFamilyIndices_ = nullptr, const void* pNext_ = nullptr)
        : pNext(pNext_), flags(flags_), size(size_), usage(usage_), sharingMode(sharingMode_), queueFamilyIndexCount(queueFamilyIndexCount_), pQueueFamilyIndices(pQueueFamilyIndices_) {}
    const void* pNext;
    BufferCreateFlags flags;
    DeviceSize size;
    BufferUsageFlags usage;
    SharingMode sharingMode;
    uint32_t queueFamilyIndexCount;
    const uint32_t* pQueueFamilyIndices;
};
struct MemoryDedicatedAllocateInfo {
    MemoryDedicatedAllocateInfo(Image image_ = {}, Buffer buffer_ = {}, const void* pNext_ = nullptr) : pNext(pNext_), image(image_), buffer(buffer_) {}
    const void* pNext;
    Image image;
    Buffer buffer;
};
struct ExportMemoryAllocateInfo {
    ExportMemoryAllocateInfo(ExternalMemoryHandleTypeFlags handleTypes_ = {}, const void* pNext_ = nullptr) : pNext(pNext_), handleTypes(handleTypes_) {}
    const void* pNext;
    ExternalMemoryHandleTypeFlags handleTypes;
};
struct MemoryAllocateInfo {
    MemoryAllocateInfo(DeviceSize allocationSize_ = 0, uint32_t memoryTypeIndex_ = 0, const void* pNext_ = nullptr) : pNext(pNext_), allocationSize(allocationSize_), memoryTypeIndex(memoryTypeIndex_) {}
    const void* pNext;
    DeviceSize allocationSize;
    uint32_t memoryTypeIndex;
};
struct MemoryGetFdInfoKHR {
    MemoryGetFdInfoKHR(DeviceMemory memory_ = {}, ExternalMemoryHandleTypeFlagBits handleType_ = ExternalMemoryHandleTypeFlagBits::eOpaqueFd, const void* pNext_ = nullptr)
        : pNext(pNext_), memory(memory_), handleType(handleType_) {}
    const void* pNext;
    DeviceMemory memory;
    ExternalMemoryHandleTypeFlagBits handleType;
};
struct ExportSemaphoreCreateInfo {
    ExportSemaphoreCreateInfo(ExternalSemaphoreHandleTypeFlags handleTypes_ = {}, const void* pNext_ = nullptr) : pNext(pNext_), handleTypes(handleTypes_) {}
    const void* pNext;
    ExternalSemaphoreHandleTypeFlags handleTypes;
};
struct SemaphoreTypeCreateInfo {
    SemaphoreTypeCreateInfo(SemaphoreType semaphoreType_ = SemaphoreType::eBinary, uint64_t initialValue_ = 0, const void* pNext_ = nullptr)
        : pNext(pNext_), semaphoreType(semaphoreType_), initialValue(initialValue_) {}
    const void* pNext;
    SemaphoreType semaphoreType;
    uint64_t initialValue;
};
struct SemaphoreCreateInfo {
    SemaphoreCreateInfo(SemaphoreCreateFlags flags_ = {}, const void* pNext_ = nullptr) : pNext(pNext_), flags(flags_) {}
    const void* pNext;
    SemaphoreCreateFlags flags;
};
struct SemaphoreGetFdInfoKHR {
    SemaphoreGetFdInfoKHR(Semaphore semaphore_ = {}, ExternalSemaphoreHandleTypeFlagBits handleType_ = ExternalSemaphoreHandleTypeFlagBits::eOpaqueFd, const void* pNext_ = nullptr)
        : pNext(pNext_), semaphore(semaphore_), handleType(handleType_) {}
    const void* pNext;
    Semaphore semaphore;
    ExternalSemaphoreHandleTypeFlagBits handleType;
};
struct TimelineSemaphoreSubmitInfo {
    TimelineSemaphoreSubmitInfo(uint32_t waitSemaphoreValueCount_ = 0, const uint64_t* pWaitSemaphoreValues_ = nullptr, uint32_t signalSemaphoreValueCount_ = 0,
                                const uint64_t* pSignalSemaphoreValues_ = nullptr, const void* pNext_ = nullptr)
        : pNext(pNext_), waitSemaphoreValueCount(waitSemaphoreValueCount_), pWaitSemaphoreValues(pWaitSemaphoreValues_), signalSemaphoreValueCount(signalSemaphoreValueCount_),
          pSignalSemaphoreValues(pSignalSemaphoreValues_) {}
    const void* pNext;
    uint32_t waitSemaphoreValueCount;
    const uint64_t* pWaitSemaphoreValues;
    uint32_t signalSemaphoreValueCount;
    const uint64_t* pSignalSemaphoreValues;
};
class CommandBuffer;
struct SubmitInfo {
    SubmitInfo(uint32_t waitSemaphoreCount_ = 0, const Semaphore* pWaitSemaphores_ = nullptr, const PipelineStageFlags* pWaitDstStageMask_ = nullptr, uint32_t commandBufferCount_ = 0,
               const CommandBuffer* pCommandBuffers_ = nullptr, uint32_t signalSemaphoreCount_ = 0, const Semaphore* pSignalSemaphores_ = nullptr, const void* pNext_ = nullptr)
        : pNext(pNext_), waitSemaphoreCount(waitSemaphoreCount_), pWaitSemaphores(pWaitSemaphores_), pWaitDstStageMask(pWaitDstStageMask_), commandBufferCount(commandBufferCount_),
          pCommandBuffers(pCommandBuffers_), signalSemaphoreCount(signalSemaphoreCount_), pSignalSemaphores(pSignalSemaphores_) {}
    const void* pNext;
    uint32_t waitSemaphoreCount;
    const Semaphore* pWaitSemaphores;
    const PipelineStageFlags* pWaitDstStageMask;
    uint32_t commandBufferCount;
    const CommandBuffer* pCommandBuffers;
    uint32_t signalSemaphoreCount;
    const Semaphore* pSignalSemaphores;
};
struct ImageSubresourceLayers {
    ImageSubresourceLayers(ImageAspectFlags aspectMask_ = {}, uint32_t mipLevel_ = 0, uint32_t baseArrayLayer_ = 0, uint32_t layerCount_ = 0)
        : aspectMask(aspectMask_), mipLevel(mipLevel_), baseArrayLayer(baseArrayLayer_), layerCount(layerCount_) {}
    ImageAspectFlags aspectMask;
    uint32_t mipLevel, baseArrayLayer, layerCount;
};
struct BufferImageCopy {
    BufferImageCopy(DeviceSize bufferOffset_ = 0, uint32_t bufferRowLength_ = 0, uint32_t bufferImageHeight_ = 0, ImageSubresourceLayers imageSubresource_ = {},
                    Offset3D imageOffset_ = {}, Extent3D imageExtent_ = {})
        : bufferOffset(bufferOffset_), bufferRowLength(bufferRowLength_), bufferImageHeight(bufferImageHeight_), imageSubresource(imageSubresource_), imageOffset(imageOffset_),
          imageExtent(imageExtent_) {}
    DeviceSize bufferOffset;
    uint32_t bufferRowLength, bufferImageHeight;
    ImageSubresourceLayers imageSubresource;
    Offset3D imageOffset;
    Extent3D imageExtent;
};
struct BufferCopy {
    BufferCopy(DeviceSize srcOffset_ = 0, DeviceSize dstOffset_ = 0, DeviceSize size_ = 0) : srcOffset(srcOffset_), dstOffset(dstOffset_), size(size_) {}
    DeviceSize srcOffset, dstOffset, size;
};
struct ImageCreateInfo { const void* pNext = nullptr; Format format = Format::eUndefined; Extent3D extent; ImageUsageFlags usage; };
struct ImageMemoryBarrier2 { const void* pNext = nullptr; ImageLayout oldLayout = ImageLayout::eUndefined, newLayout = ImageLayout::eUndefined; Image image; };
struct SamplerCreateInfo {
    const void* pNext = nullptr;
    Filter magFilter = Filter::eNearest, minFilter = Filter::eNearest;
    SamplerMipmapMode mipmapMode = SamplerMipmapMode::eNearest;
    SamplerAddressMode addressModeU = SamplerAddressMode::eRepeat, addressModeV = SamplerAddressMode::eRepeat, addressModeW = SamplerAddressMode::eRepeat;
};
struct TransformMatrixKHR { std::array<std::array<float, 4>, 3> matrix; };
struct AccelerationStructureInstanceKHR {
    TransformMatrixKHR transform;
    uint32_t instanceCustomIndex : 24;
    uint32_t mask : 8;
    uint32_t instanceShaderBindingTableRecordOffset : 24;
    uint32_t flags : 8; // VkGeometryInstanceFlagsKHR
    uint64_t accelerationStructureReference;
};

class CommandBuffer {
  public:
    void copyImageToBuffer(Image srcImage, ImageLayout srcImageLayout, Buffer dstBuffer, const ArrayProxy<const BufferImageCopy>& regions) const;
    void copyBufferToImage(Buffer srcBuffer, Image dstImage, ImageLayout dstImageLayout, const ArrayProxy<const BufferImageCopy>& regions) const;
};

class Device {
  public:
    Buffer createBuffer(const BufferCreateInfo& createInfo) const;
    void destroyBuffer(Buffer buffer) const;
    MemoryRequirements getBufferMemoryRequirements(Buffer buffer) const;
    DeviceMemory allocateMemory(const MemoryAllocateInfo& allocateInfo) const;
    void freeMemory(DeviceMemory memory) const;
    void bindBufferMemory(Buffer buffer, DeviceMemory memory, DeviceSize memoryOffset) const;
    int getMemoryFdKHR(const MemoryGetFdInfoKHR& getFdInfo) const;
    Semaphore createSemaphore(const SemaphoreCreateInfo& createInfo) const;
    void destroySemaphore(Semaphore semaphore) const;
    int getSemaphoreFdKHR(const SemaphoreGetFdInfoKHR& getFdInfo) const;
};
class PhysicalDevice {};

} // namespace vk
