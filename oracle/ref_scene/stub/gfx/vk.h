// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for the reference's include/gfx/vk.h so that the reference's own
// src/engine/resource/scene.cpp and material.cpp compile WHERE THEY LIE (oracle/Makefile target ref_scene), unmodified:
// every helios::vk class they name exists here with exactly the members they call, backed by host memory — a vk::Buffer
// is a malloc'ed block whose mapped_ptr() the reference fills with its Material / Light / Instance tables, a
// vk::DescriptorSet remembers nothing, vkUpdateDescriptorSets records which image views went into the texture array.
// Nothing here restates reference logic: the table build (scene.cpp:915-1311), the node transforms (:186-280), the light
// gathering (:540-690) and Material::is_emissive (material.cpp:71-77) all run from the reference's sources.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <string>
#include <vector>

// ---- the slice of vulkan_core.h those two files use
typedef uint64_t VkDeviceSize;
typedef uint64_t VkDeviceAddress;
typedef uint32_t VkFlags;
typedef uint32_t VkBool32;
typedef VkFlags  VkBufferUsageFlags;
typedef VkFlags  VkGeometryInstanceFlagsKHR;
typedef VkFlags  VkBuildAccelerationStructureFlagsKHR;
typedef VkFlags  VkDescriptorPoolCreateFlags;
typedef struct VkBuffer_T*                   VkBuffer;
typedef struct VkSampler_T*                  VkSampler;
typedef struct VkImageView_T*                VkImageView;
typedef struct VkDescriptorSet_T*            VkDescriptorSet;
typedef struct VkDevice_T*                   VkDevice;
typedef struct VkAccelerationStructureKHR_T* VkAccelerationStructureKHR;
#define VK_WHOLE_SIZE (~0ULL)
#define VK_FALSE 0U
enum VkStructureType
{
    VK_STRUCTURE_TYPE_WRITE_DESCRIPTOR_SET = 35,
    VK_STRUCTURE_TYPE_DESCRIPTOR_SET_VARIABLE_DESCRIPTOR_COUNT_ALLOCATE_INFO = 1000161003,
    VK_STRUCTURE_TYPE_WRITE_DESCRIPTOR_SET_ACCELERATION_STRUCTURE_KHR = 1000150007,
    VK_STRUCTURE_TYPE_ACCELERATION_STRUCTURE_GEOMETRY_INSTANCES_DATA_KHR = 1000150004,
    VK_STRUCTURE_TYPE_ACCELERATION_STRUCTURE_GEOMETRY_KHR = 1000150006,
    VK_STRUCTURE_TYPE_ACCELERATION_STRUCTURE_CREATE_INFO_KHR = 1000150017
};
enum VkImageLayout
{
    VK_IMAGE_LAYOUT_UNDEFINED = 0,
    VK_IMAGE_LAYOUT_SHADER_READ_ONLY_OPTIMAL = 5
};
enum VkDescriptorType
{
    VK_DESCRIPTOR_TYPE_COMBINED_IMAGE_SAMPLER = 1,
    VK_DESCRIPTOR_TYPE_STORAGE_BUFFER = 7,
    VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER_DYNAMIC = 8,
    VK_DESCRIPTOR_TYPE_ACCELERATION_STRUCTURE_KHR = 1000150000
};
enum VkBufferUsageFlagBits
{
    VK_BUFFER_USAGE_TRANSFER_SRC_BIT = 0x1,
    VK_BUFFER_USAGE_TRANSFER_DST_BIT = 0x2,
    VK_BUFFER_USAGE_STORAGE_BUFFER_BIT = 0x20,
    VK_BUFFER_USAGE_SHADER_DEVICE_ADDRESS_BIT = 0x20000,
    VK_BUFFER_USAGE_ACCELERATION_STRUCTURE_BUILD_INPUT_READ_ONLY_BIT_KHR = 0x80000
};
enum VkGeometryTypeKHR
{
    VK_GEOMETRY_TYPE_TRIANGLES_KHR = 0,
    VK_GEOMETRY_TYPE_INSTANCES_KHR = 2
};
enum VkAccelerationStructureTypeKHR
{
    VK_ACCELERATION_STRUCTURE_TYPE_TOP_LEVEL_KHR = 0,
    VK_ACCELERATION_STRUCTURE_TYPE_BOTTOM_LEVEL_KHR = 1
};
enum
{
    VK_GEOMETRY_INSTANCE_TRIANGLE_FACING_CULL_DISABLE_BIT_KHR = 0x1,
    VK_BUILD_ACCELERATION_STRUCTURE_ALLOW_UPDATE_BIT_KHR = 0x1,
    VK_BUILD_ACCELERATION_STRUCTURE_PREFER_FAST_TRACE_BIT_KHR = 0x4
};
struct VkDescriptorBufferInfo
{
    VkBuffer     buffer;
    VkDeviceSize offset, range;
};
struct VkDescriptorImageInfo
{
    VkSampler     sampler;
    VkImageView   imageView;
    VkImageLayout imageLayout;
};
struct VkWriteDescriptorSet
{
    VkStructureType               sType;
    const void*                   pNext;
    VkDescriptorSet               dstSet;
    uint32_t                      dstBinding, dstArrayElement, descriptorCount;
    VkDescriptorType              descriptorType;
    const VkDescriptorImageInfo*  pImageInfo;
    const VkDescriptorBufferInfo* pBufferInfo;
    const void*                   pTexelBufferView;
};
struct VkWriteDescriptorSetAccelerationStructureKHR
{
    VkStructureType                   sType;
    const void*                       pNext;
    uint32_t                          accelerationStructureCount;
    const VkAccelerationStructureKHR* pAccelerationStructures;
};
struct VkDescriptorSetVariableDescriptorCountAllocateInfo
{
    VkStructureType sType;
    const void*     pNext;
    uint32_t        descriptorSetCount;
    const uint32_t* pDescriptorCounts;
};
union VkDeviceOrHostAddressConstKHR
{
    VkDeviceAddress deviceAddress;
    const void*     hostAddress;
};
struct VkAccelerationStructureGeometryInstancesDataKHR
{
    VkStructureType               sType;
    const void*                   pNext;
    VkBool32                      arrayOfPointers;
    VkDeviceOrHostAddressConstKHR data;
};
struct VkAccelerationStructureGeometryTrianglesDataKHR
{
    VkStructureType sType;
    const void*     pNext;
    uint64_t        opaque[8];
};
union VkAccelerationStructureGeometryDataKHR
{
    VkAccelerationStructureGeometryTrianglesDataKHR triangles;
    VkAccelerationStructureGeometryInstancesDataKHR instances;
};
struct VkAccelerationStructureGeometryKHR
{
    VkStructureType                        sType;
    const void*                            pNext;
    VkGeometryTypeKHR                      geometryType;
    VkAccelerationStructureGeometryDataKHR geometry;
    VkFlags                                flags;
};
struct VkAccelerationStructureBuildSizesInfoKHR
{
    VkStructureType sType;
    const void*     pNext;
    VkDeviceSize    accelerationStructureSize, updateScratchSize, buildScratchSize;
};
struct VkAccelerationStructureCreateInfoKHR
{
    VkStructureType sType;
    const void*     pNext;
    VkFlags         createFlags;
    VkBuffer        buffer;
    VkDeviceSize    offset, size;
    int             type;
    VkDeviceAddress deviceAddress;
};
struct VkTransformMatrixKHR
{
    float matrix[3][4];
};
struct VkAccelerationStructureInstanceKHR
{
    VkTransformMatrixKHR       transform;
    uint32_t                   instanceCustomIndex : 24;
    uint32_t                   mask : 8;
    uint32_t                   instanceShaderBindingTableRecordOffset : 24;
    VkGeometryInstanceFlagsKHR flags : 8;
    uint64_t                   accelerationStructureReference;
};
enum VmaMemoryUsage
{
    VMA_MEMORY_USAGE_UNKNOWN = 0,
    VMA_MEMORY_USAGE_GPU_ONLY = 1,
    VMA_MEMORY_USAGE_CPU_ONLY = 2,
    VMA_MEMORY_USAGE_CPU_TO_GPU = 3
};
enum
{
    VMA_ALLOCATION_CREATE_MAPPED_BIT = 0x4
};
typedef VkFlags VmaAllocationCreateFlags;

// every image view that reaches the textures descriptor set, in array order: the harness reads it after Scene::update
namespace ref_scene_stub
{
std::vector<VkImageView>& texture_array();
}
inline void vkUpdateDescriptorSets(VkDevice, uint32_t n, const VkWriteDescriptorSet* w, uint32_t, const void*)
{
    for (uint32_t i = 0; i < n; i++)
        if (w[i].descriptorType == VK_DESCRIPTOR_TYPE_COMBINED_IMAGE_SAMPLER && w[i].dstBinding == 0 && w[i].descriptorCount >= 1 && w[i].pImageInfo && w[i].dstSet == (VkDescriptorSet)0x7E)
        {
            ref_scene_stub::texture_array().clear();
            for (uint32_t k = 0; k < w[i].descriptorCount; k++) ref_scene_stub::texture_array().push_back(w[i].pImageInfo[k].imageView);
        }
}

namespace helios
{
namespace vk
{
class Backend;
class Object
{
public:
    Object(std::shared_ptr<Backend> backend) : m_vk_backend(backend) {}
    virtual ~Object() {}
    inline std::weak_ptr<Backend> backend() { return m_vk_backend; }

protected:
    std::weak_ptr<Backend> m_vk_backend;
};
class Sampler
{
public:
    using Ptr = std::shared_ptr<Sampler>;
    VkSampler handle() { return (VkSampler)this; }
};
class Image
{
public:
    using Ptr = std::shared_ptr<Image>;
};
class ImageView
{
public:
    using Ptr = std::shared_ptr<ImageView>;
    VkImageView handle() { return (VkImageView)this; }
};
class Framebuffer
{
public:
    using Ptr = std::shared_ptr<Framebuffer>;
};
class RenderPass
{
public:
    using Ptr = std::shared_ptr<RenderPass>;
};
class GraphicsPipeline
{
public:
    using Ptr = std::shared_ptr<GraphicsPipeline>;
};
class PipelineLayout
{
public:
    using Ptr = std::shared_ptr<PipelineLayout>;
};
class CommandBuffer
{
public:
    using Ptr = std::shared_ptr<CommandBuffer>;
};
class DescriptorSetLayout
{
public:
    using Ptr = std::shared_ptr<DescriptorSetLayout>;
};
class DescriptorPool
{
public:
    using Ptr = std::shared_ptr<DescriptorPool>;
    struct Desc
    {
        Desc& set_max_sets(uint32_t) { return *this; }
        Desc& add_pool_size(VkDescriptorType, uint32_t) { return *this; }
        Desc& set_create_flags(VkDescriptorPoolCreateFlags) { return *this; }
    };
    static Ptr create(std::shared_ptr<Backend>, Desc) { return std::make_shared<DescriptorPool>(); }
};
class DescriptorSet
{
public:
    using Ptr = std::shared_ptr<DescriptorSet>;
    // the textures set is recognisable by its debug name (scene.cpp:838), which the harness needs to find the texture array
    static Ptr      create(std::shared_ptr<Backend>, DescriptorSetLayout::Ptr, DescriptorPool::Ptr, void* = nullptr) { return std::make_shared<DescriptorSet>(); }
    void            set_name(const std::string& n) { m_textures = n == "Textures Descriptor Set"; }
    VkDescriptorSet handle() { return m_textures ? (VkDescriptorSet)0x7E : (VkDescriptorSet)this; }

private:
    bool m_textures = false;
};
class Buffer
{
public:
    using Ptr = std::shared_ptr<Buffer>;
    static Ptr create(std::shared_ptr<Backend>, VkBufferUsageFlags, size_t size, VmaMemoryUsage, VmaAllocationCreateFlags, void* data = nullptr)
    {
        auto b = std::make_shared<Buffer>();
        b->m_size = size, b->m_data = std::calloc(size ? size : 1, 1);
        if (data) std::memcpy(b->m_data, data, size);
        return b;
    }
    ~Buffer() { std::free(m_data); }
    void*           mapped_ptr() { return m_data; }
    size_t          size() { return m_size; }
    VkBuffer        handle() { return (VkBuffer)this; }
    VkDeviceAddress device_address() { return (VkDeviceAddress)(uintptr_t)m_data; }
    void            set_name(const std::string&) {}

private:
    void*  m_data = nullptr;
    size_t m_size = 0;
};
class AccelerationStructure
{
public:
    using Ptr = std::shared_ptr<AccelerationStructure>;
    struct Desc
    {
        Desc& set_geometry_count(uint32_t) { return *this; }
        Desc& set_geometries(std::vector<VkAccelerationStructureGeometryKHR>) { return *this; }
        Desc& set_max_primitive_counts(std::vector<uint32_t>) { return *this; }
        Desc& set_type(VkAccelerationStructureTypeKHR) { return *this; }
        Desc& set_flags(VkBuildAccelerationStructureFlagsKHR) { return *this; }
    };
    static Ptr create(std::shared_ptr<Backend>, Desc) { return std::make_shared<AccelerationStructure>(); }
    const VkAccelerationStructureKHR&               handle() { return m_handle = (VkAccelerationStructureKHR)this; }
    VkDeviceAddress                                 device_address() { return (VkDeviceAddress)(uintptr_t)this; }
    const VkAccelerationStructureBuildSizesInfoKHR& build_sizes() { return m_sizes; }

private:
    VkAccelerationStructureKHR               m_handle = nullptr;
    VkAccelerationStructureBuildSizesInfoKHR m_sizes { VK_STRUCTURE_TYPE_ACCELERATION_STRUCTURE_CREATE_INFO_KHR, nullptr, 256, 256, 256 };
};
class BatchUploader
{
public:
    BatchUploader(std::shared_ptr<Backend>) {}
    void submit() {}
};
class Backend : public std::enable_shared_from_this<Backend>
{
public:
    using Ptr = std::shared_ptr<Backend>;
    static Ptr create() { return Ptr(new Backend()); }
    void       wait_idle() {}
    VkDevice   device() { return nullptr; }
    // (deferred destruction keyed by frame index in the reference, vk.cpp:4050-4071: here objects simply stay alive)
    template <class T>
    void queue_object_deletion(std::shared_ptr<T> object) { m_graveyard.push_back(std::static_pointer_cast<void>(object)); }
    Sampler::Ptr             trilinear_sampler() { return m_trilinear; }
    Sampler::Ptr             bilinear_sampler() { return m_bilinear; }
    ImageView::Ptr           default_cubemap() { return m_default_cubemap; }
    DescriptorSetLayout::Ptr scene_descriptor_set_layout() { return m_layout; }
    DescriptorSetLayout::Ptr buffer_array_descriptor_set_layout() { return m_layout; }
    DescriptorSetLayout::Ptr combined_sampler_array_descriptor_set_layout() { return m_layout; }

private:
    Backend() {}
    Sampler::Ptr                      m_trilinear = std::make_shared<Sampler>(), m_bilinear = std::make_shared<Sampler>();
    ImageView::Ptr                    m_default_cubemap = std::make_shared<ImageView>();
    DescriptorSetLayout::Ptr          m_layout          = std::make_shared<DescriptorSetLayout>();
    std::deque<std::shared_ptr<void>> m_graveyard;
};
} // namespace vk
} // namespace helios
