// core/resource_manager.h — ResourceManager (reference: include/core/resource_manager.h:12-30,
// src/engine/core/resource_manager.cpp:48-674): loads AssetCore textures, materials, meshes and scenes into the
// engine's resource classes, cached per path.  Same public signatures and failure behaviour (HELIOS_LOG_ERROR +
// nullptr); relative paths resolve to <asset root>/assets/<path> (utility::path_for_resource("assets/" + path),
// resource_manager.cpp:204), where the asset root defaults to the directory of the executable and can be set.
#pragma once
#include <gfx/vk.h>
#include <loader/loader.h>
#include <resource/material.h>
#include <resource/mesh.h>
#include <resource/scene.h>
#include <resource/texture.h>
#include <unordered_map>

namespace helios
{
class ResourceManager
{
public:
    ResourceManager(vk::Backend::Ptr backend);
    ~ResourceManager();

    Texture2D::Ptr load_texture_2d(const std::string& path, bool srgb = false);
    TextureCube::Ptr load_texture_cube(const std::string& path, bool srgb = false);
    Material::Ptr load_material(const std::string& path);
    Mesh::Ptr load_mesh(const std::string& path);
    Scene::Ptr load_scene(const std::string& path);

    // directory that contains "assets/" (the reference always uses the executable's directory)
    void set_asset_root(const std::string& dir) { m_asset_root = dir; }
    const std::string& asset_root() const { return m_asset_root; }

private:
    std::string full_path(const std::string& path) const;
    Texture2D::Ptr fetch_texture_2d(const std::string& path, bool srgb, vk::BatchUploader& uploader);
    TextureCube::Ptr fetch_texture_cube(const std::string& path, bool srgb, vk::BatchUploader& uploader);
    Material::Ptr fetch_material(const std::string& path, vk::BatchUploader& uploader);
    Mesh::Ptr fetch_mesh(const std::string& path, vk::BatchUploader& uploader);
    Node::Ptr node_from_description(std::shared_ptr<ast::SceneNode> ast_node, vk::BatchUploader& uploader);
    void fill_node(Node::Ptr node, std::shared_ptr<ast::SceneNode> ast_node, vk::BatchUploader& uploader);
    void fill_transform(TransformNode::Ptr node, std::shared_ptr<ast::SceneNode> ast_node);

    std::weak_ptr<vk::Backend> m_backend;
    std::string m_asset_root;
    std::unordered_map<std::string, Texture2D::Ptr> m_cache_2d;
    std::unordered_map<std::string, TextureCube::Ptr> m_cache_cube;
    std::unordered_map<std::string, Material::Ptr> m_materials;
    std::unordered_map<std::string, Mesh::Ptr> m_meshes;
};

// ImGuizmo::RecomposeMatrixFromComponents (external/ImGuizmo/ImGuizmo.cpp:2069-2097) restated: rows right / up /
// dir / position of Rx * Ry * Rz (row-vector convention, angles in degrees, fp32), rows scaled (|s| < FLT_EPSILON
// -> 0.001), translation in the last row.  The 16 floats are a column-major glm::mat4.
glm::mat4 recompose_matrix_from_components(const float translation[3], const float rotation_deg[3], const float scale[3]);
// texel conversion used by load_texture_2d (exposed for the loader tests): level 0 of `image` -> RGBA texels in one
// of the HL_TEX_* formats, following the reference's VkFormat choice (resource_manager.cpp:14-70).  Returns false
// for combinations the reference maps to VK_FORMAT_UNDEFINED and for BC6H / BC7.
bool convert_image_level0(const ast::Image& image, int array_slice, bool srgb, int& out_format, uint32_t& out_width, uint32_t& out_height, std::vector<uint8_t>& out_texels);
} // namespace helios
