// ast_loader.cpp — see loader/loader.h.  The binary files are little-endian PODs with natural alignment
// (SURVEY.md Appendix D); they are read field by field from one in-memory copy of the file with bounds checks, so
// a truncated or corrupt file makes the load fail instead of reading past the end.
#include <loader/loader.h>
#include <utility/json.h>
#include <cstring>
#include <fstream>
#include <iostream>

namespace ast
{
namespace
{
    struct Bytes
    {
        std::vector<uint8_t> buf;
        size_t               pos = 0;
        bool                 ok  = true;
        bool open(const std::string& path)
        {
            std::ifstream f(path, std::ios::binary | std::ios::ate);
            if (!f.is_open()) return false;
            const std::streamoff n = f.tellg();
            if (n < 0) return false;
            buf.resize((size_t)n);
            f.seekg(0);
            if (n > 0) f.read((char*)buf.data(), n);
            return (bool)f;
        }
        bool take(void* dst, size_t n)
        {
            if (!ok || n > buf.size() - pos) return ok = false;
            if (n) std::memcpy(dst, buf.data() + pos, n);
            pos += n;
            return true;
        }
        bool skip(size_t n)
        {
            if (!ok || n > buf.size() - pos) return ok = false;
            pos += n;
            return true;
        }
        template <class T>
        T get()
        {
            T v {};
            take(&v, sizeof(T));
            return v;
        }
    };
    // BINFileHeader, include/common/header.h:15-20: magic (not validated by the reference), version, type, 2 B padding
    constexpr size_t FILE_HEADER_BYTES = 8;
    constexpr size_t NAME_BYTES        = 150;

    std::string fixed_string(const char* p, size_t n)
    {
        size_t len = 0;
        while (len < n && p[len]) len++;
        return std::string(p, len);
    }
    template <size_t N>
    int index_of(const char* const (&table)[N], const std::string& s, int fallback)
    {
        for (size_t i = 0; i < N; i++)
            if (s == table[i]) return (int)i;
        return fallback;
    }
    const char* const kTextureType[]   = { "TEXTURE_ALBEDO", "TEXTURE_EMISSIVE", "TEXTURE_DISPLACEMENT", "TEXTURE_NORMAL", "TEXTURE_METALLIC", "TEXTURE_ROUGHNESS", "TEXTURE_CUSTOM" };
    const char* const kPropertyType[]  = { "PROPERTY_ALBEDO", "PROPERTY_EMISSIVE", "PROPERTY_METALLIC", "PROPERTY_ROUGHNESS" };
    const char* const kMaterialType[]  = { "MATERIAL_OPAQUE", "MATERIAL_TRANSPARENT" };
    const char* const kShadingModel[]  = { "SHADING_MODEL_STANDARD", "SHADING_MODEL_CLOTH", "SHADING_MODEL_SUBSURFACE" };
    const char* const kSceneNodeType[] = { "SCENE_NODE_MESH", "SCENE_NODE_CAMERA", "SCENE_NODE_DIRECTIONAL_LIGHT", "SCENE_NODE_SPOT_LIGHT", "SCENE_NODE_POINT_LIGHT", "SCENE_NODE_IBL", "SCENE_NODE_ROOT", "SCENE_NODE_CUSTOM" };
} // namespace

std::string parent_directory(const std::string& path)
{
    size_t cut = path.find_last_of('/');
    if (cut == std::string::npos) cut = path.find_last_of('\\');
    return cut == std::string::npos ? std::string() : path.substr(0, cut + 1);
}

// loader.cpp:39-87: file header, uint16 name length + name, BINImageHeader (8 B), then per array slice and mip
// level a BINMipSliceHeader (uint16 width, uint16 height, int32 size) followed by `size` bytes
bool load_image(const std::string& path, Image& image)
{
    Bytes f;
    if (!f.open(path)) return false;
    f.skip(FILE_HEADER_BYTES);
    const uint16_t name_len = f.get<uint16_t>();
    image.name.resize(name_len);
    f.take(name_len ? &image.name[0] : nullptr, name_len);
    uint8_t hdr[8] = {};
    f.take(hdr, sizeof(hdr));
    if (!f.ok) return false;
    uint16_t slices;
    std::memcpy(&slices, hdr + 4, 2);
    image.compression  = (CompressionType)hdr[0];
    image.type         = (PixelType)hdr[1];
    image.components   = hdr[2];
    image.array_slices = slices;
    image.mip_slices   = hdr[6];
    if (slices == 0) return false;
    image.data.assign(slices, std::vector<Image::Level>((size_t)image.mip_slices));
    for (int a = 0; a < image.array_slices; a++)
        for (int m = 0; m < image.mip_slices; m++)
        {
            Image::Level& L = image.data[(size_t)a][(size_t)m];
            L.width         = f.get<uint16_t>();
            L.height        = f.get<uint16_t>();
            const int32_t n = f.get<int32_t>();
            if (!f.ok || n < 0) return false;
            L.bytes.resize((size_t)n);
            if (!f.take(L.bytes.data(), (size_t)n)) return false;
        }
    return true;
}

// loader.cpp:89-164: file header, BINMeshFileHeader (196 B), Vertex[], SkeletalVertex[], uint32 indices[],
// SubMesh[], BINMeshMaterialJson[] (150-byte paths relative to the mesh file); every material is loaded too and a
// missing one fails the mesh
bool load_mesh(const std::string& path, Mesh& mesh)
{
    Bytes f;
    if (!f.open(path)) return false;
    f.skip(FILE_HEADER_BYTES);
    const uint32_t mesh_count = f.get<uint32_t>(), material_count = f.get<uint32_t>(), vertex_count = f.get<uint32_t>(), skeletal_count = f.get<uint32_t>(), index_count = f.get<uint32_t>();
    f.take(mesh.max_extents, 12), f.take(mesh.min_extents, 12);
    char name[NAME_BYTES + 2] = {};
    f.take(name, NAME_BYTES + 2); // 150 name bytes + 2 B tail padding of the 196-byte header
    if (!f.ok) return false;
    mesh.name = fixed_string(name, NAME_BYTES);
    static_assert(sizeof(Vertex) == 56 && sizeof(SkeletalVertex) == 88 && sizeof(SubMesh) == 196, "on-disk record sizes");
    // counts are checked against the bytes that are really there before anything is allocated
    auto read_array = [&](auto& vec, uint32_t count) {
        using T = typename std::remove_reference<decltype(vec)>::type::value_type;
        if (!f.ok || (uint64_t)count * sizeof(T) > f.buf.size() - f.pos) return f.ok = false;
        vec.resize(count);
        return f.take(vec.data(), (size_t)count * sizeof(T));
    };
    if (!read_array(mesh.vertices, vertex_count) || !read_array(mesh.skeletal_vertices, skeletal_count) || !read_array(mesh.indices, index_count) || !read_array(mesh.submeshes, mesh_count)) return false;
    if ((uint64_t)material_count * NAME_BYTES > f.buf.size() - f.pos) return false;
    mesh.materials.assign(material_count, Material());
    mesh.material_paths.assign(material_count, std::string());
    const std::string parent = parent_directory(path);
    for (uint32_t i = 0; i < material_count; i++)
    {
        char rel[NAME_BYTES];
        if (!f.take(rel, NAME_BYTES)) return false;
        const std::string relative = fixed_string(rel, NAME_BYTES);
        mesh.material_paths[i]     = parent + relative;
    }
    for (uint32_t i = 0; i < material_count; i++)
        if (!load_material(mesh.material_paths[i], mesh.materials[i]))
        {
            std::cout << "Failed to load material: " << mesh.material_paths[i].substr(parent.size()) << std::endl;
            return false;
        }
    return true;
}

// loader.cpp:166-357
bool load_material(const std::string& path, Material& material)
{
    helios::json::Value j;
    try
    {
        j = helios::json::parse_file(path);
    }
    catch (const std::exception&)
    {
        return false;
    }
    material.name          = j.get_string("name", "untitled");
    material.double_sided  = j.get_bool("double_sided", false);
    material.alpha_mask    = j.get_bool("alpha_mask", false);
    material.material_type = (MaterialType)index_of(kMaterialType, j.get_string("material_type", kMaterialType[0]), MATERIAL_OPAQUE);
    material.shading_model = (ShadingModel)index_of(kShadingModel, j.get_string("shading_model", kShadingModel[0]), SHADING_MODEL_STANDARD);
    const std::string parent = parent_directory(path);
    if (const helios::json::Value* textures = j.find("textures"))
        for (const helios::json::Value& jt : textures->array)
        {
            Texture t;
            t.srgb = jt.get_bool("srgb", true);
            if (const helios::json::Value* p = jt.find("path"))
                if (p->is_string()) t.path = parent + p->string;
            t.type          = (TextureType)index_of(kTextureType, jt.get_string("type", kTextureType[0]), TEXTURE_ALBEDO);
            t.channel_index = (uint32_t)jt.get_float("channel_index", 0.0f);
            material.textures.push_back(t);
        }
    if (const helios::json::Value* props = j.find("properties"))
        for (const helios::json::Value& jp : props->array)
        {
            const helios::json::Value* type = jp.find("type");
            if (!type || !type->is_string()) continue;
            const int k = index_of(kPropertyType, type->string, -1);
            if (k < 0) continue;
            MaterialProperty p;
            p.type = (PropertyType)k;
            if (k == PROPERTY_ALBEDO || k == PROPERTY_EMISSIVE)
            {
                // a vector property needs exactly four numbers (:290-297)
                if (!jp.get_vector("value", p.vec4_value, 4)) continue;
            }
            else if (k == PROPERTY_METALLIC)
            {
                if (!jp.find("value")) continue; // (:326-330: metallic without a value is dropped, roughness is kept)
                p.float_value = jp.get_float("value", 0.0f);
            }
            else
                p.float_value = jp.get_float("value", 0.0f);
            material.properties.push_back(p);
        }
    return true;
}

// loader.cpp:359-568
static std::shared_ptr<SceneNode> read_node(const helios::json::Value& j, int depth)
{
    const helios::json::Value* type = j.find("type");
    if (!type || !type->is_string() || depth > 512) return nullptr;
    const int k = index_of(kSceneNodeType, type->string, -1);
    if (k < 0) return nullptr;
    auto n  = std::make_shared<SceneNode>();
    n->type = (SceneNodeType)k;
    n->name = j.get_string("name", "");
    if (k != SCENE_NODE_IBL && k != SCENE_NODE_CUSTOM)
    {
        j.get_vector("position", n->position, 3);
        j.get_vector("rotation", n->rotation, 3);
        j.get_vector("scale", n->scale, 3);
    }
    switch (k)
    {
        case SCENE_NODE_MESH:
            n->mesh              = j.get_string("mesh", "");
            n->material_override = j.get_string("material_override", "");
            n->casts_shadow      = j.get_bool("casts_shadow", n->casts_shadow);
            break;
        case SCENE_NODE_SPOT_LIGHT:
            n->inner_cone_angle = j.get_float("inner_cone_angle", n->inner_cone_angle);
            n->outer_cone_angle = j.get_float("outer_cone_angle", n->outer_cone_angle);
            // fall through
        case SCENE_NODE_DIRECTIONAL_LIGHT:
        case SCENE_NODE_POINT_LIGHT:
            n->intensity     = j.get_float("intensity", n->intensity);
            n->radius        = j.get_float("radius", n->radius);
            n->casts_shadows = j.get_bool("casts_shadows", n->casts_shadows);
            j.get_vector("color", n->color, 3);
            break;
        case SCENE_NODE_CAMERA:
            n->near_plane = j.get_float("near_plane", n->near_plane);
            n->far_plane  = j.get_float("far_plane", n->far_plane);
            n->fov        = j.get_float("fov", n->fov);
            break;
        case SCENE_NODE_IBL: n->image = j.get_string("image", ""); break;
        default: break;
    }
    if (const helios::json::Value* children = j.find("children"))
        for (const helios::json::Value& c : children->array) n->children.push_back(read_node(c, depth + 1));
    return n;
}
bool load_scene(const std::string& path, Scene& scene)
{
    helios::json::Value j;
    try
    {
        j = helios::json::parse_file(path);
    }
    catch (const std::exception&)
    {
        return false;
    }
    scene.name = j.get_string("name", scene.name);
    if (const helios::json::Value* g = j.find("scene_graph")) scene.scene_graph = read_node(*g, 0);
    return true;
}
} // namespace ast
