// resource_manager.cpp — see core/resource_manager.h.
#include <core/resource_manager.h>
#include <utility/bc_decode.h>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <unistd.h>

namespace helios
{
namespace
{
    std::string executable_directory()
    {
        char          buf[4096];
        const ssize_t n = ::readlink("/proc/self/exe", buf, sizeof(buf) - 1);
        if (n <= 0) return ".";
        std::string       p(buf, (size_t)n);
        const size_t      cut = p.find_last_of('/');
        return cut == std::string::npos ? "." : p.substr(0, cut);
    }
    bool is_absolute(const std::string& p) { return !p.empty() && p[0] == '/'; }

    float half_to_float(uint16_t h)
    {
        const uint32_t sign = (uint32_t)(h >> 15) << 31;
        uint32_t       exp = (h >> 10) & 31u, man = h & 1023u, bits;
        if (exp == 0)
        {
            if (man == 0)
                bits = sign;
            else
            {
                // subnormal half: renormalise
                int e = -1;
                do
                {
                    man <<= 1;
                    e++;
                } while (!(man & 1024u));
                bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 1023u) << 13);
            }
        }
        else if (exp == 31)
            bits = sign | 0x7F800000u | (man << 13);
        else
            bits = sign | ((exp + 112u) << 23) | (man << 13);
        float f;
        std::memcpy(&f, &bits, 4);
        return f;
    }
} // namespace

glm::mat4 recompose_matrix_from_components(const float translation[3], const float rotation_deg[3], const float scale[3])
{
    const float DEG2RAD = 3.14159265358979323846f / 180.0f;
    // rot[i] = rotation about unit axis i (matrix_t::RotationAxis, ImGuizmo.cpp:528-566, row-vector form)
    float rot[3][16];
    for (int i = 0; i < 3; i++)
    {
        const float n[3] = { i == 0 ? 1.0f : 0.0f, i == 1 ? 1.0f : 0.0f, i == 2 ? 1.0f : 0.0f };
        const float a = rotation_deg[i] * DEG2RAD, s = std::sin(a), c = std::cos(a), k = 1.0f - c;
        const float xx = n[0] * n[0] * k + c, yy = n[1] * n[1] * k + c, zz = n[2] * n[2] * k + c;
        const float xy = n[0] * n[1] * k, yz = n[1] * n[2] * k, zx = n[2] * n[0] * k;
        const float xs = n[0] * s, ys = n[1] * s, zs = n[2] * s;
        const float m[16] = { xx, xy + zs, zx - ys, 0.0f, xy - zs, yy, yz + xs, 0.0f, zx + ys, yz - xs, zz, 0.0f, 0.0f, 0.0f, 0.0f, 1.0f };
        std::memcpy(rot[i], m, sizeof(m));
    }
    auto mul = [](const float* a, const float* b, float* r) {
        for (int row = 0; row < 4; row++)
            for (int col = 0; col < 4; col++) r[row * 4 + col] = a[row * 4 + 0] * b[0 + col] + a[row * 4 + 1] * b[4 + col] + a[row * 4 + 2] * b[8 + col] + a[row * 4 + 3] * b[12 + col];
    };
    float t[16], m[16];
    mul(rot[0], rot[1], t);
    mul(t, rot[2], m);
    for (int i = 0; i < 3; i++)
    {
        const float s = std::fabs(scale[i]) < FLT_EPSILON ? 0.001f : scale[i];
        for (int k = 0; k < 4; k++) m[i * 4 + k] *= s;
    }
    m[12] = translation[0], m[13] = translation[1], m[14] = translation[2], m[15] = 1.0f;
    glm::mat4 out;
    std::memcpy(&out, m, sizeof(m));
    return out;
}

bool convert_image_level0(const ast::Image& image, int array_slice, bool srgb, int& fmt, uint32_t& w, uint32_t& h, std::vector<uint8_t>& out)
{
    if (array_slice < 0 || array_slice >= (int)image.data.size() || image.data[(size_t)array_slice].empty()) return false;
    const ast::Image::Level& L = image.data[(size_t)array_slice][0];
    w = L.width, h = L.height;
    const size_t n = (size_t)w * h;
    if (n == 0) return false;
    if (image.compression != ast::COMPRESSION_NONE)
    {
        // kCompressedFormats: BC4 / BC5 / BC6H have no sRGB variant (VK_FORMAT_UNDEFINED)
        if (srgb && (image.compression == ast::COMPRESSION_BC4 || image.compression == ast::COMPRESSION_BC5 || image.compression == ast::COMPRESSION_BC6)) return false;
        if (image.compression == ast::COMPRESSION_BC6)
        {
            // VK_FORMAT_BC6H_UFLOAT_BLOCK (kCompressedFormats): HDR, expanded to RGBA32F texels (alpha 1)
            std::vector<uint16_t> half;
            if (!decode_bc6h(L.bytes.data(), L.bytes.size(), w, h, false, half)) return false;
            fmt = HL_TEX_RGBA32F;
            out.resize(n * 16);
            float* dst = (float*)out.data();
            for (size_t i = 0; i < n; i++)
            {
                for (int k = 0; k < 3; k++) dst[i * 4 + (size_t)k] = half_to_float(half[i * 3 + (size_t)k]);
                dst[i * 4 + 3] = 1.0f;
            }
            return true;
        }
        if (!decode_bc((int)image.compression, L.bytes.data(), L.bytes.size(), w, h, out)) return false;
        fmt = srgb ? HL_TEX_RGBA8_SRGB : HL_TEX_RGBA8_UNORM;
        return true;
    }
    const int c = image.components;
    if (c < 1 || c > 4) return false;
    if (image.type == ast::PIXEL_TYPE_UNORM8)
    {
        if (srgb && c < 3) return false; // kSRGBFormats: only R8G8B8_SRGB / R8G8B8A8_SRGB exist
        if (L.bytes.size() < n * (size_t)c) return false;
        // non-sRGB 8-bit data is created as *_SNORM by the reference (kNonSRGBFormats, the quirk of SURVEY A.8);
        // channels the file does not have read as 0, 0, 1 (Vulkan's conversion to RGBA)
        fmt                  = srgb ? HL_TEX_RGBA8_SRGB : HL_TEX_RGBA8_SNORM;
        const uint8_t one    = srgb ? 255 : 127;
        out.resize(n * 4);
        for (size_t i = 0; i < n; i++)
            for (int k = 0; k < 4; k++) out[i * 4 + (size_t)k] = k < c ? L.bytes[i * (size_t)c + (size_t)k] : (k == 3 ? one : 0);
        return true;
    }
    if (srgb) return false; // no sRGB float formats
    const size_t cs = image.type == ast::PIXEL_TYPE_FLOAT16 ? 2 : 4;
    if ((image.type != ast::PIXEL_TYPE_FLOAT16 && image.type != ast::PIXEL_TYPE_FLOAT32) || L.bytes.size() < n * (size_t)c * cs) return false;
    fmt = HL_TEX_RGBA32F;
    out.resize(n * 16);
    float* dst = (float*)out.data();
    for (size_t i = 0; i < n; i++)
        for (int k = 0; k < 4; k++)
        {
            float v = k == 3 ? 1.0f : 0.0f;
            if (k < c)
            {
                const uint8_t* src = L.bytes.data() + (i * (size_t)c + (size_t)k) * cs;
                if (cs == 4)
                    std::memcpy(&v, src, 4);
                else
                {
                    uint16_t hbits;
                    std::memcpy(&hbits, src, 2);
                    v = half_to_float(hbits);
                }
            }
            dst[i * 4 + (size_t)k] = v;
        }
    return true;
}

ResourceManager::ResourceManager(vk::Backend::Ptr backend) : m_backend(backend), m_asset_root(executable_directory()) {}
ResourceManager::~ResourceManager() {}

std::string ResourceManager::full_path(const std::string& path) const { return is_absolute(path) ? path : m_asset_root + "/assets/" + path; }

Texture2D::Ptr ResourceManager::load_texture_2d(const std::string& path, bool srgb)
{
    if (m_backend.expired()) return nullptr;
    vk::BatchUploader uploader(m_backend.lock());
    auto              resource = fetch_texture_2d(path, srgb, uploader);
    uploader.submit();
    return resource;
}
TextureCube::Ptr ResourceManager::load_texture_cube(const std::string& path, bool srgb)
{
    if (m_backend.expired()) return nullptr;
    vk::BatchUploader uploader(m_backend.lock());
    auto              resource = fetch_texture_cube(path, srgb, uploader);
    uploader.submit();
    return resource;
}
Material::Ptr ResourceManager::load_material(const std::string& path)
{
    if (m_backend.expired()) return nullptr;
    vk::BatchUploader uploader(m_backend.lock());
    auto              resource = fetch_material(path, uploader);
    uploader.submit();
    return resource;
}
Mesh::Ptr ResourceManager::load_mesh(const std::string& path)
{
    if (m_backend.expired()) return nullptr;
    vk::BatchUploader uploader(m_backend.lock());
    auto              resource = fetch_mesh(path, uploader);
    uploader.submit();
    return resource;
}
Scene::Ptr ResourceManager::load_scene(const std::string& path)
{
    if (m_backend.expired()) return nullptr;
    vk::Backend::Ptr  backend = m_backend.lock();
    vk::BatchUploader uploader(backend);
    ast::Scene        ast_scene;
    const std::string full = full_path(path);
    if (!ast::load_scene(full, ast_scene)) return nullptr;
    Node::Ptr root = ast_scene.scene_graph ? node_from_description(ast_scene.scene_graph, uploader) : nullptr;
    uploader.submit();
    return root ? Scene::create(backend, ast_scene.name, root, full) : nullptr;
}

Texture2D::Ptr ResourceManager::fetch_texture_2d(const std::string& path, bool srgb, vk::BatchUploader&)
{
    auto it = m_cache_2d.find(path);
    if (it != m_cache_2d.end()) return it->second;
    ast::Image        image;
    const std::string full = full_path(path);
    if (!ast::load_image(full, image))
    {
        HELIOS_LOG_ERROR("Failed to load Texture: " + path);
        return nullptr;
    }
    int                  fmt = 0;
    uint32_t             w = 0, h = 0;
    std::vector<uint8_t> texels;
    if (!convert_image_level0(image, 0, srgb, fmt, w, h, texels))
    {
        HELIOS_LOG_ERROR("Failed to load Texture: " + path + " (pixel format / compression has no image format)");
        return nullptr;
    }
    Texture2D::Ptr texture = Texture2D::create(m_backend.lock(), fmt, w, h, texels.data(), full);
    if (texture) m_cache_2d[path] = texture;
    return texture;
}
TextureCube::Ptr ResourceManager::fetch_texture_cube(const std::string& path, bool srgb, vk::BatchUploader&)
{
    auto it = m_cache_cube.find(path);
    if (it != m_cache_cube.end()) return it->second;
    ast::Image        image;
    const std::string full = full_path(path);
    if (!ast::load_image(full, image) || image.array_slices != 6)
    {
        HELIOS_LOG_ERROR("Failed to load Texture: " + path);
        return nullptr;
    }
    // six faces -> RGBA32F (8-bit faces go through the same decode tables as the device: sRGB / the SNORM quirk)
    std::vector<float> faces;
    uint32_t           size = 0;
    for (int f = 0; f < 6; f++)
    {
        int                  fmt = 0;
        uint32_t             w = 0, h = 0;
        std::vector<uint8_t> texels;
        if (!convert_image_level0(image, f, srgb, fmt, w, h, texels) || w != h || (f > 0 && w != size))
        {
            HELIOS_LOG_ERROR("Failed to load Texture: " + path + " (not a cube map the path can sample)");
            return nullptr;
        }
        size = w;
        faces.resize((size_t)6 * size * size * 4);
        float* dst = faces.data() + (size_t)f * size * size * 4;
        if (fmt == HL_TEX_RGBA32F)
            std::memcpy(dst, texels.data(), texels.size());
        else
            for (size_t i = 0; i < texels.size(); i++)
            {
                const double v = texels[i] / 255.0;
                if (fmt == HL_TEX_RGBA8_SNORM)
                    dst[i] = (float)std::max(-1.0, (int8_t)texels[i] / 127.0);
                else if (fmt == HL_TEX_RGBA8_SRGB && (i & 3) != 3)
                    dst[i] = (float)(v <= 0.04045 ? v / 12.92 : std::pow((v + 0.055) / 1.055, 2.4));
                else
                    dst[i] = (float)v;
            }
    }
    TextureCube::Ptr texture = TextureCube::create(m_backend.lock(), size, faces.data(), full);
    if (texture) m_cache_cube[path] = texture;
    return texture;
}
Material::Ptr ResourceManager::fetch_material(const std::string& path, vk::BatchUploader& uploader)
{
    auto it = m_materials.find(path);
    if (it != m_materials.end()) return it->second;
    ast::Material     am;
    const std::string full = full_path(path);
    if (!ast::load_material(full, am))
    {
        HELIOS_LOG_ERROR("Failed to load Material: " + path);
        return nullptr;
    }
    // resource_manager.cpp:306-411: one local texture list per material, textures shared between slots by path;
    // displacement / custom textures are not used by the path
    std::vector<Texture2D::Ptr>               textures;
    std::unordered_map<std::string, uint32_t> slot_of;
    TextureInfo                               albedo, emissive, normal, metallic, roughness;
    for (const ast::Texture& t : am.textures)
    {
        TextureInfo* info = t.type == ast::TEXTURE_ALBEDO ? &albedo : t.type == ast::TEXTURE_EMISSIVE ? &emissive : t.type == ast::TEXTURE_NORMAL ? &normal : t.type == ast::TEXTURE_METALLIC ? &metallic : t.type == ast::TEXTURE_ROUGHNESS ? &roughness : nullptr;
        if (!info) continue;
        if (slot_of.find(t.path) == slot_of.end())
        {
            slot_of[t.path] = (uint32_t)textures.size();
            textures.push_back(fetch_texture_2d(t.path, t.srgb, uploader));
        }
        info->array_index   = (int32_t)slot_of[t.path];
        info->channel_index = (int32_t)t.channel_index;
    }
    glm::vec4 albedo_value(0.0f), emissive_value(0.0f);
    float     metallic_value = 0.0f, roughness_value = 1.0f; // defaults, resource_manager.cpp:321-324
    for (const ast::MaterialProperty& p : am.properties)
    {
        if (p.type == ast::PROPERTY_ALBEDO) albedo_value = glm::vec4(p.vec4_value[0], p.vec4_value[1], p.vec4_value[2], p.vec4_value[3]);
        if (p.type == ast::PROPERTY_EMISSIVE) emissive_value = glm::vec4(p.vec4_value[0], p.vec4_value[1], p.vec4_value[2], p.vec4_value[3]);
        if (p.type == ast::PROPERTY_METALLIC) metallic_value = p.float_value;
        if (p.type == ast::PROPERTY_ROUGHNESS) roughness_value = p.float_value;
    }
    for (const Texture2D::Ptr& t : textures)
        if (!t)
        {
            // the reference would store a null texture and crash when the scene tables are built
            HELIOS_LOG_ERROR("Failed to load Material: " + path + " (a texture is missing)");
            return nullptr;
        }
    Material::Ptr material = Material::create(m_backend.lock(), am.material_type == ast::MATERIAL_OPAQUE ? MATERIAL_OPAQUE : MATERIAL_TRANSPARENT, textures, albedo, normal, metallic, roughness, emissive,
                                              albedo_value, emissive_value, metallic_value, roughness_value, am.alpha_mask, full);
    m_materials[path] = material;
    return material;
}
Mesh::Ptr ResourceManager::fetch_mesh(const std::string& path, vk::BatchUploader& uploader)
{
    auto it = m_meshes.find(path);
    if (it != m_meshes.end()) return it->second;
    ast::Mesh         am;
    const std::string full = full_path(path);
    if (!ast::load_mesh(full, am))
    {
        HELIOS_LOG_ERROR("Failed to load Mesh: " + path);
        return nullptr;
    }
    // resource_manager.cpp:440-477: widen the 56-byte vertices to the 80-byte shader layout, copy the submesh
    // table, tag every vertex with the index of the (last) submesh that references it in position.w
    std::vector<Vertex>  vertices(am.vertices.size());
    std::vector<SubMesh> submeshes(am.submeshes.size());
    for (size_t i = 0; i < vertices.size(); i++)
    {
        const ast::Vertex& v  = am.vertices[i];
        vertices[i].position  = glm::vec4(v.position[0], v.position[1], v.position[2], 0.0f);
        vertices[i].tex_coord = glm::vec4(v.tex_coord[0], v.tex_coord[1], 0.0f, 0.0f);
        vertices[i].normal    = glm::vec4(v.normal[0], v.normal[1], v.normal[2], 0.0f);
        vertices[i].tangent   = glm::vec4(v.tangent[0], v.tangent[1], v.tangent[2], 0.0f);
        vertices[i].bitangent = glm::vec4(v.bitangent[0], v.bitangent[1], v.bitangent[2], 0.0f);
    }
    for (size_t i = 0; i < submeshes.size(); i++)
    {
        const ast::SubMesh& s     = am.submeshes[i];
        submeshes[i].name         = std::string(s.name, strnlen(s.name, sizeof(s.name)));
        submeshes[i].mat_idx      = s.material_index;
        submeshes[i].index_count  = s.index_count;
        submeshes[i].vertex_count = s.vertex_count;
        submeshes[i].base_vertex  = s.base_vertex;
        submeshes[i].base_index   = s.base_index;
        submeshes[i].max_extents  = glm::vec3(s.max_extents[0], s.max_extents[1], s.max_extents[2]);
        submeshes[i].min_extents  = glm::vec3(s.min_extents[0], s.min_extents[1], s.min_extents[2]);
    }
    for (size_t s = 0; s < submeshes.size(); s++)
    {
        const SubMesh& sm = submeshes[s];
        if ((uint64_t)sm.base_index + sm.index_count > am.indices.size())
        {
            HELIOS_LOG_ERROR("Failed to load Mesh: " + path + " (submesh index range outside the index buffer)");
            return nullptr;
        }
        for (uint32_t i = sm.base_index; i < sm.base_index + sm.index_count; i++)
        {
            const uint64_t v = (uint64_t)sm.base_vertex + am.indices[i];
            if (v >= vertices.size())
            {
                HELIOS_LOG_ERROR("Failed to load Mesh: " + path + " (index outside the vertex buffer)");
                return nullptr;
            }
            vertices[(size_t)v].position.w = float(s);
        }
    }
    std::vector<Material::Ptr> materials(am.material_paths.size());
    for (size_t i = 0; i < materials.size(); i++) materials[i] = fetch_material(am.material_paths[i], uploader);
    for (const SubMesh& sm : submeshes)
        if (sm.mat_idx >= materials.size() || !materials[sm.mat_idx])
        {
            HELIOS_LOG_ERROR("Failed to load Mesh: " + path + " (a submesh has no material)");
            return nullptr;
        }
    Mesh::Ptr mesh = Mesh::create(m_backend.lock(), vertices, am.indices, submeshes, materials, uploader, full);
    m_meshes[path] = mesh;
    return mesh;
}

Node::Ptr ResourceManager::node_from_description(std::shared_ptr<ast::SceneNode> n, vk::BatchUploader& uploader)
{
    if (!n) return nullptr;
    switch (n->type)
    {
        case ast::SCENE_NODE_MESH:
        {
            MeshNode::Ptr node = std::shared_ptr<MeshNode>(new MeshNode(n->name));
            if (n->mesh != "")
            {
                Mesh::Ptr mesh = fetch_mesh(n->mesh, uploader);
                if (mesh)
                    node->set_mesh(mesh);
                else
                    HELIOS_LOG_ERROR("Failed to load mesh: " + n->mesh);
                if (n->material_override != "")
                {
                    Material::Ptr material_override = fetch_material(n->material_override, uploader);
                    if (!material_override) HELIOS_LOG_ERROR("Failed to load material override: " + n->material_override);
                    node->set_material_override(material_override);
                }
            }
            fill_transform(node, n);
            fill_node(node, n, uploader);
            return node;
        }
        case ast::SCENE_NODE_CAMERA:
        {
            CameraNode::Ptr node = std::shared_ptr<CameraNode>(new CameraNode(n->name));
            node->set_near_plane(n->near_plane);
            node->set_far_plane(n->far_plane);
            node->set_fov(n->fov);
            fill_transform(node, n);
            fill_node(node, n, uploader);
            return node;
        }
        case ast::SCENE_NODE_DIRECTIONAL_LIGHT:
        {
            DirectionalLightNode::Ptr node = std::shared_ptr<DirectionalLightNode>(new DirectionalLightNode(n->name));
            node->set_color(glm::vec3(n->color[0], n->color[1], n->color[2]));
            node->set_intensity(n->intensity);
            node->set_radius(n->radius);
            fill_transform(node, n);
            fill_node(node, n, uploader);
            return node;
        }
        case ast::SCENE_NODE_SPOT_LIGHT:
        {
            SpotLightNode::Ptr node = std::shared_ptr<SpotLightNode>(new SpotLightNode(n->name));
            node->set_color(glm::vec3(n->color[0], n->color[1], n->color[2]));
            node->set_intensity(n->intensity);
            node->set_radius(n->radius);
            node->set_inner_cone_angle(n->inner_cone_angle);
            node->set_outer_cone_angle(n->inner_cone_angle); // sic: the reference passes the INNER angle twice (resource_manager.cpp:591)
            fill_transform(node, n);
            fill_node(node, n, uploader);
            return node;
        }
        case ast::SCENE_NODE_POINT_LIGHT:
        {
            PointLightNode::Ptr node = std::shared_ptr<PointLightNode>(new PointLightNode(n->name));
            node->set_color(glm::vec3(n->color[0], n->color[1], n->color[2]));
            node->set_intensity(n->intensity);
            node->set_radius(n->radius);
            fill_transform(node, n);
            fill_node(node, n, uploader);
            return node;
        }
        case ast::SCENE_NODE_IBL:
        {
            IBLNode::Ptr node = std::shared_ptr<IBLNode>(new IBLNode(n->name));
            if (n->image != "")
            {
                TextureCube::Ptr cube = fetch_texture_cube(n->image, false, uploader);
                if (cube)
                    node->set_image(cube);
                else
                    HELIOS_LOG_ERROR("Failed to load cubemap: " + n->image);
            }
            fill_node(node, n, uploader);
            return node;
        }
        case ast::SCENE_NODE_ROOT:
        {
            RootNode::Ptr node = std::shared_ptr<RootNode>(new RootNode(n->name));
            fill_transform(node, n);
            fill_node(node, n, uploader);
            return node;
        }
        default: return nullptr; // SCENE_NODE_CUSTOM has no engine node (resource_manager.cpp:513)
    }
}
void ResourceManager::fill_node(Node::Ptr node, std::shared_ptr<ast::SceneNode> ast_node, vk::BatchUploader& uploader)
{
    for (auto& ast_child : ast_node->children)
    {
        Node::Ptr child = node_from_description(ast_child, uploader);
        if (child) node->add_child(child);
    }
}
void ResourceManager::fill_transform(TransformNode::Ptr node, std::shared_ptr<ast::SceneNode> ast_node)
{
    node->set_from_local_transform(recompose_matrix_from_components(ast_node->position, ast_node->rotation, ast_node->scale));
}
} // namespace helios
