// resource/texture.h — Texture2D / TextureCube (reference: include/resource/texture.h:8-57).  A Texture2D holds
// level 0 of an image (the path only samples textureLod(..., 0)); the texel data is kept on the host until
// Scene::update assigns the texture its slot in the global array (descriptor set 4) and uploads it with
// hl_texture2d_create.  A TextureCube holds six RGBA32F faces for IBLNode (hl_envmap_set).
#pragma once
#include <gfx/vk.h>
#include <vector>

namespace helios
{
class Texture : public vk::Object
{
public:
    using Ptr = std::shared_ptr<Texture>;

    Texture(vk::Backend::Ptr backend, const std::string& path);
    virtual ~Texture();
    uint32_t id() { return m_id; }
    std::string path() { return m_path; }

protected:
    std::string m_path;
    uint32_t m_id;
};

class Texture2D : public Texture
{
public:
    using Ptr = std::shared_ptr<Texture2D>;
    // format: HL_TEX_RGBA8_UNORM / _SRGB / _SNORM / HL_TEX_RGBA32F; texels = width * height * 4 components
    static Texture2D::Ptr create(vk::Backend::Ptr backend, int format, uint32_t width, uint32_t height, const void* level0_texels, const std::string& path = "");
    ~Texture2D();
    int format() const { return m_format; }
    uint32_t width() const { return m_width; }
    uint32_t height() const { return m_height; }
    const std::vector<uint8_t>& texels() const { return m_texels; }

private:
    Texture2D(vk::Backend::Ptr backend, int format, uint32_t width, uint32_t height, const void* texels, const std::string& path);
    int m_format;
    uint32_t m_width, m_height;
    std::vector<uint8_t> m_texels;
};

class TextureCube : public Texture
{
public:
    using Ptr = std::shared_ptr<TextureCube>;
    // faces +X,-X,+Y,-Y,+Z,-Z, each size * size RGBA32F
    static TextureCube::Ptr create(vk::Backend::Ptr backend, uint32_t size, const float* rgba32f_faces, const std::string& path = "");
    ~TextureCube();
    uint32_t size() const { return m_size; }
    const std::vector<float>& faces() const { return m_faces; }

private:
    TextureCube(vk::Backend::Ptr backend, uint32_t size, const float* faces, const std::string& path);
    uint32_t m_size;
    std::vector<float> m_faces;
};
} // namespace helios
