// backend.cpp — helios::vk::Backend on top of the C ABI (see include/gfx/vk.h for what it replaces).
#include <gfx/vk.h>

namespace helios
{
namespace vk
{
Backend::Ptr Backend::create(int device_ordinal, uint32_t width, uint32_t height)
{
    Ptr b(new Backend());
    b->m_extents = { width, height };
    hl_context ctx = nullptr;
    const hl_status st = hl_context_create(device_ordinal, width, height, &ctx);
    if (st != HL_OK)
    {
        const char*       why = hl_last_error(nullptr);
        const std::string msg = std::string("(Vulkan-free backend) failed to create the CUDA context: ") + (why ? why : "unknown error");
        HELIOS_LOG_FATAL(msg);
        throw std::runtime_error(msg);
    }
    b->m_ctx = ctx;
    return b;
}

Backend::Ptr Backend::create_without_device(uint32_t width, uint32_t height)
{
    Ptr b(new Backend());
    b->m_extents = { width, height };
    return b;
}

Backend::~Backend()
{
    if (m_ctx) hl_context_destroy(m_ctx);
}

void Backend::resize(uint32_t width, uint32_t height)
{
    m_extents = { width, height };
    if (m_ctx) check(hl_context_resize(m_ctx, width, height), "hl_context_resize");
}

void Backend::wait_idle()
{
    if (m_ctx) check(hl_synchronize(m_ctx), "hl_synchronize");
}

void Backend::check(hl_status st, const char* what)
{
    if (st == HL_OK) return;
    const char*       why = hl_last_error(m_ctx);
    const std::string msg = std::string(what) + " failed (status " + std::to_string(st) + "): " + (why ? why : "");
    HELIOS_LOG_FATAL(msg);
    throw std::runtime_error(msg);
}

hl_context Backend::require_device(const char* what)
{
    if (!m_ctx)
    {
        const std::string msg = std::string(what) + ": this backend was created without a device (there is no CPU path)";
        HELIOS_LOG_FATAL(msg);
        throw std::runtime_error(msg);
    }
    return m_ctx;
}
} // namespace vk
} // namespace helios
