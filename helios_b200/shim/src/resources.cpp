// resources.cpp — Texture2D / TextureCube / Material / Mesh (reference: src/engine/resource/texture.cpp,
// material.cpp, mesh.cpp).  Ids come from per-class counters as in the reference (g_last_*_id).
#include <resource/material.h>
#include <resource/mesh.h>
#include <resource/texture.h>
#include <cstring>

namespace helios
{
static uint32_t g_last_texture_id  = 0;
static uint32_t g_last_material_id = 0;
static uint32_t g_last_mesh_id     = 0;

// ---- textures ------------------------------------------------------------------------------------
Texture::Texture(vk::Backend::Ptr backend, const std::string& path) : vk::Object(backend), m_path(path), m_id(g_last_texture_id++) {}
Texture::~Texture() {}

static size_t texel_bytes(int format) { return format == HL_TEX_RGBA32F ? 16 : 4; }

Texture2D::Ptr Texture2D::create(vk::Backend::Ptr backend, int format, uint32_t width, uint32_t height, const void* level0_texels, const std::string& path)
{
    if (!level0_texels || width == 0 || height == 0 || format < HL_TEX_RGBA8_UNORM || format > HL_TEX_RGBA32F)
    {
        HELIOS_LOG_ERROR("Texture2D::create: invalid image description for " + path);
        return nullptr; // recoverable load error: log + nullptr (core/resource_manager.cpp:206-221)
    }
    return std::shared_ptr<Texture2D>(new Texture2D(backend, format, width, height, level0_texels, path));
}
Texture2D::Texture2D(vk::Backend::Ptr backend, int format, uint32_t width, uint32_t height, const void* texels, const std::string& path) :
    Texture(backend, path), m_format(format), m_width(width), m_height(height)
{
    m_texels.resize((size_t)width * height * texel_bytes(format));
    std::memcpy(m_texels.data(), texels, m_texels.size());
}
Texture2D::~Texture2D() {}

TextureCube::Ptr TextureCube::create(vk::Backend::Ptr backend, uint32_t size, const float* rgba32f_faces, const std::string& path)
{
    if (!rgba32f_faces || size == 0)
    {
        HELIOS_LOG_ERROR("TextureCube::create: invalid image description for " + path);
        return nullptr;
    }
    return std::shared_ptr<TextureCube>(new TextureCube(backend, size, rgba32f_faces, path));
}
TextureCube::TextureCube(vk::Backend::Ptr backend, uint32_t size, const float* faces, const std::string& path) : Texture(backend, path), m_size(size)
{
    m_faces.assign(faces, faces + (size_t)6 * size * size * 4);
}
TextureCube::~TextureCube() {}

// ---- material ------------------------------------------------------------------------------------
Material::Ptr Material::create(vk::Backend::Ptr backend, MaterialType type, std::vector<std::shared_ptr<Texture2D>> textures, TextureInfo albedo_texture_info, TextureInfo normal_texture_info,
                               TextureInfo metallic_texture_info, TextureInfo roughness_texture_info, TextureInfo emissive_texture_info, glm::vec4 albedo_value, glm::vec4 emissive_value,
                               float metallic_value, float roughness_value, bool alpha_test, const std::string& path)
{
    return std::shared_ptr<Material>(new Material(backend, type, textures, albedo_texture_info, normal_texture_info, metallic_texture_info, roughness_texture_info, emissive_texture_info, albedo_value,
                                                  emissive_value, metallic_value, roughness_value, alpha_test, path));
}
Material::Material(vk::Backend::Ptr backend, MaterialType type, std::vector<std::shared_ptr<Texture2D>> textures, TextureInfo albedo, TextureInfo normal, TextureInfo metallic, TextureInfo roughness,
                   TextureInfo emissive, glm::vec4 albedo_value, glm::vec4 emissive_value, float metallic_value, float roughness_value, bool alpha_test, const std::string& path) :
    vk::Object(backend),
    m_type(type), m_textures(textures), m_albedo_texture_info(albedo), m_normal_texture_info(normal), m_metallic_texture_info(metallic), m_roughness_texture_info(roughness),
    m_emissive_texture_info(emissive), m_albedo_value(albedo_value), m_emissive_value(emissive_value), m_metallic_value(metallic_value), m_roughness_value(roughness_value),
    m_alpha_test(alpha_test), m_id(g_last_material_id++), m_path(path)
{
}
Material::~Material() {}
// material.cpp:71-77
bool Material::is_emissive()
{
    if (m_emissive_texture_info.array_index != -1) return true;
    return m_emissive_value.x > 0.0f || m_emissive_value.y > 0.0f || m_emissive_value.z > 0.0f;
}

// ---- mesh ----------------------------------------------------------------------------------------
Mesh::Ptr Mesh::create(vk::Backend::Ptr backend, std::vector<Vertex> vertices, std::vector<uint32_t> indices, std::vector<SubMesh> submeshes, std::vector<std::shared_ptr<Material>> materials,
                       vk::BatchUploader& uploader, const std::string& path)
{
    (void)uploader;
    return std::shared_ptr<Mesh>(new Mesh(backend, vertices, indices, submeshes, materials, path));
}
Mesh::Mesh(vk::Backend::Ptr backend, std::vector<Vertex>& vertices, std::vector<uint32_t>& indices, std::vector<SubMesh> submeshes, std::vector<std::shared_ptr<Material>> materials,
           const std::string& path) :
    vk::Object(backend), m_sub_meshes(submeshes), m_materials(materials), m_id(g_last_mesh_id++), m_path(path)
{
    // one BLAS geometry per submesh, opaque iff the material is MATERIAL_OPAQUE and not alpha tested (mesh.cpp:66-103)
    std::vector<hl_submesh> geometries(submeshes.size());
    for (size_t i = 0; i < submeshes.size(); i++)
    {
        if (submeshes[i].mat_idx >= materials.size() || !materials[submeshes[i].mat_idx])
        {
            const std::string msg = "Mesh::create: submesh " + std::to_string(i) + " of " + path + " refers to a missing material";
            HELIOS_LOG_FATAL(msg);
            throw std::runtime_error(msg);
        }
        Material::Ptr material   = materials[submeshes[i].mat_idx];
        geometries[i].base_index  = submeshes[i].base_index;
        geometries[i].index_count = submeshes[i].index_count;
        geometries[i].vertex_count = submeshes[i].vertex_count;
        geometries[i].opaque      = (material->type() == MATERIAL_OPAQUE && !material->is_alpha_tested()) ? 1u : 0u;
    }
    if (backend->has_device())
        backend->check(hl_mesh_create(backend->context(), reinterpret_cast<const hl_vertex*>(vertices.data()), (uint32_t)vertices.size(), indices.data(), (uint32_t)indices.size(), geometries.data(),
                                      (uint32_t)geometries.size(), &m_handle),
                       "hl_mesh_create");
}
Mesh::~Mesh()
{
    auto backend = m_vk_backend.lock();
    if (m_handle)
    {
        if (!backend)
        {
            // the reference throws when an object outlives its backend (vk.cpp:2416-2420); a destructor must not,
            // so the shim logs it — the context owned the handle and has already released it
            HELIOS_LOG_ERROR("Mesh destroyed after its backend: " + m_path);
            return;
        }
        hl_mesh_destroy(backend->context(), m_handle);
    }
}
hl_build_stats Mesh::build_stats()
{
    hl_build_stats s {};
    auto           backend = m_vk_backend.lock();
    if (backend && m_handle) backend->check(hl_mesh_build_stats(backend->context(), m_handle, &s), "hl_mesh_build_stats");
    return s;
}
} // namespace helios
