// loader/loader.h — reader for the AssetCore asset formats the engine ingests:
// binary images (.ast: 2D textures and cube maps, optional mip chains, uncompressed or block compressed),
// binary meshes (.ast), material JSON, scene JSON.  Interface after the reference's
// external/AssetCore/include/loader/loader.h:9-13 (ast::load_image / load_mesh / load_material / load_scene
// returning false on failure); the data model holds the same fields with owning containers.
// File layout: external/AssetCore/include/common/{header,image,mesh}.h; reading order: src/loader/loader.cpp:39-164;
// JSON keys and defaults: src/loader/loader.cpp:166-568.
#pragma once
#include <glm.hpp>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace ast
{
// ---- images (include/common/image.h:9-47) -----------------------------------------------------------------------
enum CompressionType
{
    COMPRESSION_NONE = 0,
    COMPRESSION_BC1  = 1,
    COMPRESSION_BC1a = 2,
    COMPRESSION_BC2  = 3,
    COMPRESSION_BC3  = 4,
    COMPRESSION_BC3n = 5,
    COMPRESSION_BC4  = 6,
    COMPRESSION_BC5  = 7,
    COMPRESSION_BC6  = 8,
    COMPRESSION_BC7  = 9,
    COMPRESSION_ETC1 = 10,
    COMPRESSION_ETC2 = 11,
    COMPRESSION_PVR  = 12
};
enum PixelType
{
    PIXEL_TYPE_UNORM8  = 1,
    PIXEL_TYPE_FLOAT16 = 2,
    PIXEL_TYPE_FLOAT32 = 4
};
struct Image
{
    struct Level
    {
        uint32_t             width = 0, height = 0;
        std::vector<uint8_t> bytes;
    };
    std::string                     name;
    int                             components   = 0;
    int                             mip_slices   = 0;
    int                             array_slices = 0;
    PixelType                       type         = PIXEL_TYPE_UNORM8;
    CompressionType                 compression  = COMPRESSION_NONE;
    std::vector<std::vector<Level>> data; // [array slice][mip slice]
};

// ---- materials (include/common/material.h) ----------------------------------------------------------------------
enum TextureType
{
    TEXTURE_ALBEDO,
    TEXTURE_EMISSIVE,
    TEXTURE_DISPLACEMENT,
    TEXTURE_NORMAL,
    TEXTURE_METALLIC,
    TEXTURE_ROUGHNESS,
    TEXTURE_CUSTOM
};
enum PropertyType
{
    PROPERTY_ALBEDO,
    PROPERTY_EMISSIVE,
    PROPERTY_METALLIC,
    PROPERTY_ROUGHNESS
};
enum MaterialType
{
    MATERIAL_OPAQUE,
    MATERIAL_TRANSPARENT
};
enum ShadingModel
{
    SHADING_MODEL_STANDARD,
    SHADING_MODEL_CLOTH,
    SHADING_MODEL_SUBSURFACE
};
struct Texture
{
    TextureType type = TEXTURE_ALBEDO;
    std::string path;
    bool        srgb          = true;
    uint32_t    channel_index = 0;
};
struct MaterialProperty
{
    PropertyType type          = PROPERTY_ALBEDO;
    float        vec4_value[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
    float        float_value   = 0.0f;
};
struct Material
{
    std::string                   name;
    bool                          double_sided  = false;
    bool                          alpha_mask    = false;
    MaterialType                  material_type = MATERIAL_OPAQUE;
    ShadingModel                  shading_model = SHADING_MODEL_STANDARD;
    std::vector<Texture>          textures;
    std::vector<MaterialProperty> properties;
};

// ---- meshes (include/common/mesh.h) -----------------------------------------------------------------------------
struct Vertex
{
    float position[3], tex_coord[2], normal[3], tangent[3], bitangent[3];
}; // 56 B on disk
struct SkeletalVertex
{
    float   position[3], tex_coord[2], normal[3], tangent[3], bitangent[3];
    int32_t bone_indices[4];
    float   bone_weights[4];
}; // 88 B on disk
struct SubMesh
{
    uint32_t material_index, index_count, vertex_count, base_vertex, base_index;
    float    max_extents[3], min_extents[3];
    char     name[150];
}; // 196 B on disk
struct Mesh
{
    std::string                 name;
    std::vector<Vertex>         vertices;
    std::vector<SkeletalVertex> skeletal_vertices;
    std::vector<uint32_t>       indices;
    std::vector<SubMesh>        submeshes;
    std::vector<Material>       materials;
    std::vector<std::string>    material_paths;
    float                       max_extents[3] = { 0, 0, 0 }, min_extents[3] = { 0, 0, 0 };
};

// ---- scenes (include/common/scene.h) ----------------------------------------------------------------------------
enum SceneNodeType
{
    SCENE_NODE_MESH,
    SCENE_NODE_CAMERA,
    SCENE_NODE_DIRECTIONAL_LIGHT,
    SCENE_NODE_SPOT_LIGHT,
    SCENE_NODE_POINT_LIGHT,
    SCENE_NODE_IBL,
    SCENE_NODE_ROOT,
    SCENE_NODE_CUSTOM,
    SCENE_NODE_COUNT
};
// One node type with every per-kind field (the reference uses a class per kind; fields a kind does not have keep
// their defaults).  Members the reference leaves uninitialised when the JSON omits them start at neutral values
// here (position / rotation 0, scale 1, colour 1, intensity 1, radius 0, fov 60, planes 0.1 / 1000).
struct SceneNode
{
    SceneNodeType                           type = SCENE_NODE_CUSTOM;
    std::string                             name;
    std::vector<std::shared_ptr<SceneNode>> children;
    // transform nodes
    float position[3] = { 0, 0, 0 }, rotation[3] = { 0, 0, 0 }, scale[3] = { 1, 1, 1 }; // rotation: Euler angles, degrees
    // mesh
    std::string mesh, material_override;
    bool        casts_shadow = true;
    // lights
    float color[3] = { 1, 1, 1 }, intensity = 1.0f, radius = 0.0f, inner_cone_angle = 0.0f, outer_cone_angle = 0.0f;
    bool  casts_shadows = true;
    // camera
    float near_plane = 0.1f, far_plane = 1000.0f, fov = 60.0f;
    // IBL
    std::string image;
};
struct Scene
{
    std::string                name;
    std::shared_ptr<SceneNode> scene_graph;
};

bool load_image(const std::string& path, Image& image);
bool load_mesh(const std::string& path, Mesh& mesh);
bool load_material(const std::string& path, Material& material);
bool load_scene(const std::string& path, Scene& scene);
// directory part of a path including the trailing separator ("" when there is none): what relative material /
// texture paths are resolved against (external/AssetCore/src/common/filesystem.cpp:184-193)
std::string parent_directory(const std::string& path);
} // namespace ast
