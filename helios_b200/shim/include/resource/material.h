// resource/material.h — Material (reference: include/resource/material.h:11-97, src/engine/resource/material.cpp).
// Same factory signature and accessors; the object is host-only data, turned into one 80-byte MaterialData row
// by Scene::update.
#pragma once
#include <gfx/vk.h>
#include <glm.hpp>
#include <memory>
#include <vector>

namespace helios
{
class Texture2D;

enum MaterialType
{
    MATERIAL_OPAQUE,
    MATERIAL_TRANSPARENT
};

struct TextureInfo
{
    int32_t array_index = -1;
    int32_t channel_index = -1;
};

class Material : public vk::Object
{
public:
    using Ptr = std::shared_ptr<Material>;

    static Material::Ptr create(vk::Backend::Ptr backend, MaterialType type, std::vector<std::shared_ptr<Texture2D>> textures, TextureInfo albedo_texture_info, TextureInfo normal_texture_info,
                                TextureInfo metallic_texture_info, TextureInfo roughness_texture_info, TextureInfo emissive_texture_info, glm::vec4 albedo_value = glm::vec4(0.0f),
                                glm::vec4 emissive_value = glm::vec4(0.0f), float metallic_value = 0.0f, float roughness_value = 0.0f, bool alpha_test = false, const std::string& path = "");
    ~Material();

    bool is_emissive();
    bool is_alpha_tested() { return m_alpha_test; }
    MaterialType type() { return m_type; }
    std::shared_ptr<Texture2D> albedo_texture() { return pick(m_albedo_texture_info); }
    std::shared_ptr<Texture2D> normal_texture() { return pick(m_normal_texture_info); }
    std::shared_ptr<Texture2D> metallic_texture() { return pick(m_metallic_texture_info); }
    std::shared_ptr<Texture2D> roughness_texture() { return pick(m_roughness_texture_info); }
    std::shared_ptr<Texture2D> emissive_texture() { return pick(m_emissive_texture_info); }
    TextureInfo albedo_texture_info() { return m_albedo_texture_info; }
    TextureInfo normal_texture_info() { return m_normal_texture_info; }
    TextureInfo metallic_texture_info() { return m_metallic_texture_info; }
    TextureInfo roughness_texture_info() { return m_roughness_texture_info; }
    TextureInfo emissive_texture_info() { return m_emissive_texture_info; }
    glm::vec4 albedo_value() { return m_albedo_value; }
    glm::vec4 emissive_value() { return m_emissive_value; }
    float metallic_value() { return m_metallic_value; }
    float roughness_value() { return m_roughness_value; }
    uint32_t id() { return m_id; }
    std::string path() { return m_path; }

private:
    Material(vk::Backend::Ptr backend, MaterialType type, std::vector<std::shared_ptr<Texture2D>> textures, TextureInfo albedo, TextureInfo normal, TextureInfo metallic, TextureInfo roughness,
             TextureInfo emissive, glm::vec4 albedo_value, glm::vec4 emissive_value, float metallic_value, float roughness_value, bool alpha_test, const std::string& path);
    std::shared_ptr<Texture2D> pick(const TextureInfo& i) { return i.array_index == -1 ? nullptr : m_textures[(size_t)i.array_index]; }

    MaterialType m_type = MATERIAL_OPAQUE;
    std::vector<std::shared_ptr<Texture2D>> m_textures;
    TextureInfo m_albedo_texture_info, m_normal_texture_info, m_metallic_texture_info, m_roughness_texture_info, m_emissive_texture_info;
    glm::vec4 m_albedo_value = glm::vec4(0.0f);
    glm::vec4 m_emissive_value = glm::vec4(0.0f);
    float m_metallic_value = 0.0f;
    float m_roughness_value = 0.0f;
    bool m_alpha_test = false;
    uint32_t m_id;
    std::string m_path;
};
} // namespace helios
