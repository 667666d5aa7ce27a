// resource/mesh.h — Vertex, SubMesh, Mesh (reference: include/resource/mesh.h:10-81, src/engine/resource/mesh.cpp).
// Mesh::create(backend, vertices, indices, submeshes, materials, uploader, path) keeps its signature; where the
// reference creates a VBO/IBO and records a BLAS build per submesh geometry (mesh.cpp:61-116), this one hands
// the same arrays to hl_mesh_create, which uploads them and builds the 8-wide BVH on the GPU.
#pragma once
#include <gfx/vk.h>
#include <glm.hpp>
#include <memory>
#include <vector>

namespace helios
{
struct Vertex
{
    glm::vec4 position; // w = submesh index (core/resource_manager.cpp:467-473)
    glm::vec4 tex_coord;
    glm::vec4 normal;
    glm::vec4 tangent;
    glm::vec4 bitangent;
};
static_assert(sizeof(Vertex) == sizeof(hl_vertex), "Vertex must match the 80-byte shader ABI");

struct SubMesh
{
    std::string name;
    uint32_t mat_idx;
    uint32_t index_count;
    uint32_t vertex_count;
    uint32_t base_vertex;
    uint32_t base_index;
    glm::vec3 max_extents;
    glm::vec3 min_extents;
};

class Material;

class Mesh : public vk::Object
{
public:
    using Ptr = std::shared_ptr<Mesh>;

    static Mesh::Ptr create(vk::Backend::Ptr backend, std::vector<Vertex> vertices, std::vector<uint32_t> indices, std::vector<SubMesh> submeshes,
                            std::vector<std::shared_ptr<Material>> materials, vk::BatchUploader& uploader, const std::string& path = "");
    ~Mesh();

    const std::vector<std::shared_ptr<Material>>& materials() { return m_materials; }
    const std::vector<SubMesh>& sub_meshes() { return m_sub_meshes; }
    hl_mesh acceleration_structure() { return m_handle; } // the BLAS handle
    uint32_t id() { return m_id; }
    std::string path() { return m_path; }
    hl_build_stats build_stats();

private:
    Mesh(vk::Backend::Ptr backend, std::vector<Vertex>& vertices, std::vector<uint32_t>& indices, std::vector<SubMesh> submeshes, std::vector<std::shared_ptr<Material>> materials,
         const std::string& path);

    hl_mesh m_handle = nullptr;
    std::vector<SubMesh> m_sub_meshes;
    std::vector<std::shared_ptr<Material>> m_materials;
    uint32_t m_id;
    std::string m_path;
};
} // namespace helios
