// gfx/vk.h — what is left of the reference's device backend (include/gfx/vk.h, src/engine/gfx/vk.cpp) on the
// path-trace pass once Vulkan is replaced: helios::vk::Backend owns one hl_context (include/helios_b200.h)
// instead of a VkDevice + swap chain, and the engine classes keep taking a vk::Backend::Ptr exactly as before
// (Renderer(vk::Backend::Ptr), Mesh::create(backend, ...), Scene::create(backend, ...)).
//   vk::Backend::create       <- vk.cpp:3300-3420 (instance/device/swap chain)   -> hl_context_create
//   swap_chain_extents()      <- vk.h Backend::swap_chain_extents                -> the headless image size
//   vk::BatchUploader         <- vk.cpp:3160-3235 (staging + BLAS batch)         -> no-op handle: hl_mesh_create
//                                                                                 uploads and builds directly
//   vk::CommandBuffer         <- per-frame command buffer                         -> tag type (one CUDA stream per context)
// Errors: a non-zero hl_status is logged with HELIOS_LOG_FATAL and thrown as std::runtime_error, the
// reference's convention for device failures.
#pragma once
#include <helios_b200.h>
#include <utility/logger.h>
#include <memory>
#include <stdexcept>
#include <string>

namespace helios
{
namespace vk
{
struct Extent2D
{
    uint32_t width, height;
};

class Backend : public std::enable_shared_from_this<Backend>
{
public:
    using Ptr = std::shared_ptr<Backend>;
    // device_ordinal selects the GPU (one process per GPU); width x height is the output image
    static Ptr create(int device_ordinal, uint32_t width, uint32_t height);
    // host-only mode for tools that inspect the scene tables without a GPU: every device call throws
    static Ptr create_without_device(uint32_t width, uint32_t height);
    ~Backend();

    hl_context context() { return m_ctx; }
    bool has_device() const { return m_ctx != nullptr; }
    Extent2D swap_chain_extents() const { return m_extents; }
    void resize(uint32_t width, uint32_t height);
    void wait_idle();
    // throws std::runtime_error(what + hl_last_error) when st != HL_OK
    void check(hl_status st, const char* what);
    hl_context require_device(const char* what);

private:
    Backend() = default;
    hl_context m_ctx = nullptr;
    Extent2D m_extents { 0, 0 };
};

class Object
{
public:
    Object(Backend::Ptr backend) : m_vk_backend(backend) {}
    virtual ~Object() {}

protected:
    std::weak_ptr<Backend> m_vk_backend;
};

class CommandBuffer
{
public:
    using Ptr = std::shared_ptr<CommandBuffer>;
};

class BatchUploader
{
public:
    BatchUploader(Backend::Ptr backend) : m_backend(backend) {}
    void submit() {} // uploads and BLAS builds are issued by Mesh::create itself
private:
    std::weak_ptr<Backend> m_backend;
};
} // namespace vk
} // namespace helios
