// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product).
// ref_scene_tool: the reference's OWN scene code — src/engine/resource/scene.cpp (node hierarchy, transforms, light
// gathering, Scene::update, Scene::create_gpu_resources :915-1311) and src/engine/resource/material.cpp — compiled where
// they lie against the stand-in device layer of oracle/ref_scene/stub (see gfx/vk.h there), driven through the engine's
// public API exactly as an application would: build Texture2D / Material / Mesh / nodes, Scene::create, Scene::update.
// Whatever the reference writes into its Material / Light / Instance storage buffers and per-node submesh buffers is dumped
// in the format of `helios_headless --dump-tables`, so tests/test_ref_scene.py can diff it byte for byte against the C++
// host layer (helios_b200/shim/src/scene.cpp) and the Python host (helios_b200/scenes.py).  This pins SURVEY.md §8 row a17.
//
//   ref_scene_tool scene.hlsc tables.bin      (scene file format: helios_b200/scene_io.py)
//
// The only reference code NOT taken from the checkout are the constructors of the resource classes whose real bodies
// upload to Vulkan (Mesh, Texture*, HosekWilkieSkyModel: src/engine/resource/mesh.cpp, texture.cpp,
// gfx/hosek_wilkie_sky_model.cpp): below they just keep what the table build reads (sub-meshes, materials, ids).
#define private public // the dump reads Scene's storage buffers (private members); layout is unaffected
#define protected public
#include <resource/scene.h>
#include <resource/mesh.h>
#include <resource/material.h>
#include <resource/texture.h>
#undef private
#undef protected
#include <cstdio>
#include <fstream>
#include <stdexcept>

namespace ref_scene_stub
{
std::vector<VkImageView>& texture_array()
{
    static std::vector<VkImageView> v;
    return v;
}
} // namespace ref_scene_stub

namespace helios
{
// ---- resource classes: bodies that would talk to Vulkan ---------------------------------------------------------
static uint32_t g_next_mesh_id = 0, g_next_texture_id = 0;
Mesh::Ptr Mesh::create(vk::Backend::Ptr backend, std::vector<Vertex> vertices, std::vector<uint32_t> indices, std::vector<SubMesh> submeshes, std::vector<std::shared_ptr<Material>> materials,
                       vk::BatchUploader& uploader, const std::string& path)
{
    auto vbo = vk::Buffer::create(backend, VK_BUFFER_USAGE_STORAGE_BUFFER_BIT, sizeof(Vertex) * vertices.size(), VMA_MEMORY_USAGE_GPU_ONLY, 0, vertices.data());
    auto ibo = vk::Buffer::create(backend, VK_BUFFER_USAGE_STORAGE_BUFFER_BIT, 4 * indices.size(), VMA_MEMORY_USAGE_GPU_ONLY, 0, indices.data());
    return std::shared_ptr<Mesh>(new Mesh(backend, vbo, ibo, submeshes, materials, uploader, path));
}
Mesh::Mesh(vk::Backend::Ptr backend, vk::Buffer::Ptr vbo, vk::Buffer::Ptr ibo, std::vector<SubMesh> submeshes, std::vector<std::shared_ptr<Material>> materials, vk::BatchUploader&, const std::string& path) :
    vk::Object(backend), m_vbo(vbo), m_ibo(ibo), m_sub_meshes(submeshes), m_materials(materials), m_id(g_next_mesh_id++), m_path(path)
{
    m_blas = vk::AccelerationStructure::create(backend, vk::AccelerationStructure::Desc());
}
Mesh::~Mesh() {}
Texture::Texture(vk::Backend::Ptr backend, vk::Image::Ptr image, vk::ImageView::Ptr image_view, const std::string& path) : vk::Object(backend), m_image(image), m_image_view(image_view), m_path(path), m_id(g_next_texture_id++) {}
Texture::~Texture() {}
Texture2D::Ptr Texture2D::create(vk::Backend::Ptr backend, vk::Image::Ptr image, vk::ImageView::Ptr image_view, const std::string& path) { return std::shared_ptr<Texture2D>(new Texture2D(backend, image, image_view, path)); }
Texture2D::Texture2D(vk::Backend::Ptr backend, vk::Image::Ptr image, vk::ImageView::Ptr image_view, const std::string& path) : Texture(backend, image, image_view, path) {}
Texture2D::~Texture2D() {}
TextureCube::Ptr TextureCube::create(vk::Backend::Ptr backend, vk::Image::Ptr image, vk::ImageView::Ptr image_view, const std::string& path) { return std::shared_ptr<TextureCube>(new TextureCube(backend, image, image_view, path)); }
TextureCube::TextureCube(vk::Backend::Ptr backend, vk::Image::Ptr image, vk::ImageView::Ptr image_view, const std::string& path) : Texture(backend, image, image_view, path) {}
TextureCube::~TextureCube() {}
HosekWilkieSkyModel::HosekWilkieSkyModel(vk::Backend::Ptr) { m_cubemap_image_view = std::make_shared<vk::ImageView>(); }
HosekWilkieSkyModel::~HosekWilkieSkyModel() {}
void HosekWilkieSkyModel::update(vk::CommandBuffer::Ptr, glm::vec3) {} // (the fit is pinned separately: tests/test_ref_glsl.py)
} // namespace helios

using namespace helios;

namespace
{
// mirrors of the table rows declared inside scene.cpp (:25-52) — sizes only, the bytes are the reference's
struct Row80
{
    char b[80];
};
struct Row64
{
    char b[64];
};
struct Row144
{
    char b[144];
};
struct Reader
{
    std::vector<char> buf;
    size_t            pos = 0;
    explicit Reader(const std::string& path)
    {
        std::ifstream f(path, std::ios::binary);
        if (!f) throw std::runtime_error("cannot open " + path);
        buf.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    }
    void raw(void* dst, size_t n)
    {
        if (pos + n > buf.size()) throw std::runtime_error("scene file is truncated");
        std::memcpy(dst, buf.data() + pos, n);
        pos += n;
    }
    template <class T>
    T get()
    {
        T v;
        raw(&v, sizeof(T));
        return v;
    }
};
glm::vec3 read_vec3(Reader& r)
{
    const float x = r.get<float>(), y = r.get<float>(), z = r.get<float>();
    return glm::vec3(x, y, z);
}
glm::quat read_quat(Reader& r)
{
    const float w = r.get<float>(), x = r.get<float>(), y = r.get<float>(), z = r.get<float>();
    return glm::quat(w, x, y, z);
}
template <class T>
void write_rows(FILE* f, const void* p, uint32_t n)
{
    std::fwrite(&n, 4, 1, f);
    if (n) std::fwrite(p, sizeof(T), n, f);
}
} // namespace

int main(int argc, char** argv)
{
    if (argc != 3)
    {
        std::fprintf(stderr, "usage: ref_scene_tool scene.hlsc tables.bin\n");
        return 2;
    }
    try
    {
        Reader r(argv[1]);
        char   magic[8];
        r.raw(magic, 8);
        if (std::memcmp(magic, "HLSC0001", 8) != 0) throw std::runtime_error("not a HLSC0001 scene file");
        const uint32_t width = r.get<uint32_t>(), height = r.get<uint32_t>();
        r.get<uint32_t>(); // max bounces
        r.get<float>();    // shadow ray bias
        auto              backend = vk::Backend::create();
        vk::BatchUploader uploader(backend);

        std::vector<Texture2D::Ptr> textures(r.get<uint32_t>());
        for (auto& t : textures)
        {
            const int32_t  fmt = r.get<int32_t>();
            const uint32_t w = r.get<uint32_t>(), h = r.get<uint32_t>();
            r.pos += (size_t)w * h * (fmt == 3 ? 16 : 4); // texels: not part of the tables
            t = Texture2D::create(backend, nullptr, std::make_shared<vk::ImageView>(), "texture");
        }
        std::vector<Material::Ptr> materials(r.get<uint32_t>());
        for (size_t k = 0; k < materials.size(); k++)
        {
            const uint32_t type = r.get<uint32_t>(), alpha_test = r.get<uint32_t>();
            glm::vec4      albedo, emissive;
            r.raw(&albedo, 16), r.raw(&emissive, 16);
            const float metallic = r.get<float>(), roughness = r.get<float>();
            int32_t     tex[5]; // albedo, normal, metallic, roughness, emissive: indices into the file's texture list
            r.raw(tex, sizeof(tex));
            const int32_t rough_ch = r.get<int32_t>(), metal_ch = r.get<int32_t>();
            std::vector<Texture2D::Ptr> local; // each Material owns a texture list; TextureInfo::array_index points into it
            TextureInfo                 info[5];
            for (int s = 0; s < 5; s++)
                if (tex[s] >= 0)
                {
                    info[s].array_index = (int32_t)local.size();
                    local.push_back(textures.at((size_t)tex[s]));
                }
            info[2].channel_index = metal_ch, info[3].channel_index = rough_ch;
            materials[k] = Material::create(backend, type == 0 ? MATERIAL_OPAQUE : MATERIAL_TRANSPARENT, local, info[0], info[1], info[2], info[3], info[4], albedo, emissive, metallic, roughness, alpha_test != 0,
                                            "material" + std::to_string(k));
        }
        std::vector<Mesh::Ptr> meshes(r.get<uint32_t>());
        for (size_t k = 0; k < meshes.size(); k++)
        {
            const uint32_t        nv = r.get<uint32_t>(), ni = r.get<uint32_t>(), nsub = r.get<uint32_t>();
            std::vector<Vertex>   vertices(nv);
            std::vector<uint32_t> indices(ni);
            r.raw(vertices.data(), sizeof(Vertex) * (size_t)nv);
            r.raw(indices.data(), 4 * (size_t)ni);
            std::vector<SubMesh> subs(nsub);
            for (auto& s : subs)
            {
                s.mat_idx = r.get<uint32_t>(), s.index_count = r.get<uint32_t>(), s.vertex_count = r.get<uint32_t>(), s.base_vertex = r.get<uint32_t>(), s.base_index = r.get<uint32_t>();
                s.name = "submesh";
            }
            std::vector<Material::Ptr> mesh_materials(r.get<uint32_t>());
            for (auto& m : mesh_materials) m = materials.at(r.get<uint32_t>());
            meshes[k] = Mesh::create(backend, std::move(vertices), std::move(indices), subs, mesh_materials, uploader, "mesh" + std::to_string(k));
        }
        auto           root    = std::make_shared<RootNode>("root");
        const uint32_t n_nodes = r.get<uint32_t>();
        for (uint32_t k = 0; k < n_nodes; k++)
        {
            const uint32_t mesh = r.get<uint32_t>();
            glm::mat4      model;
            r.raw(&model, 64);
            auto node = std::make_shared<MeshNode>("mesh_node" + std::to_string(k));
            root->add_child(node);
            node->set_mesh(meshes.at(mesh));
            node->set_from_global_transform(model);
        }
        auto camera = std::make_shared<CameraNode>("camera");
        root->add_child(camera);
        camera->set_position(read_vec3(r));
        camera->set_orientation(read_quat(r));
        camera->set_fov(r.get<float>()), camera->set_near_plane(r.get<float>()), camera->set_far_plane(r.get<float>());
        camera->set_focal_length(r.get<float>()), camera->set_aperture_radius(r.get<float>());
        for (uint32_t k = 0, n = r.get<uint32_t>(); k < n; k++)
        {
            auto l = std::make_shared<DirectionalLightNode>("directional" + std::to_string(k));
            root->add_child(l);
            l->set_orientation(read_quat(r));
            l->set_color(read_vec3(r)), l->set_intensity(r.get<float>()), l->set_radius(r.get<float>());
        }
        for (uint32_t k = 0, n = r.get<uint32_t>(); k < n; k++)
        {
            auto l = std::make_shared<PointLightNode>("point" + std::to_string(k));
            root->add_child(l);
            l->set_position(read_vec3(r));
            l->set_color(read_vec3(r)), l->set_intensity(r.get<float>()), l->set_radius(r.get<float>());
        }
        for (uint32_t k = 0, n = r.get<uint32_t>(); k < n; k++)
        {
            auto l = std::make_shared<SpotLightNode>("spot" + std::to_string(k));
            root->add_child(l);
            l->set_position(read_vec3(r));
            l->set_orientation(read_quat(r));
            l->set_color(read_vec3(r)), l->set_intensity(r.get<float>()), l->set_radius(r.get<float>());
            l->set_inner_cone_angle(r.get<float>()), l->set_outer_cone_angle(r.get<float>());
        }
        if (const uint32_t cube = r.get<uint32_t>())
        {
            r.pos += (size_t)6 * cube * cube * 16;
            auto ibl = std::make_shared<IBLNode>("ibl");
            root->add_child(ibl);
            ibl->set_image(TextureCube::create(backend, nullptr, std::make_shared<vk::ImageView>(), "ibl"));
        }
        auto        scene = Scene::create(backend, "scene", root, argv[1]);
        RenderState rs;
        rs.setup(width, height, nullptr);
        scene->update(rs); // hierarchy update: the tables are built (scene.cpp:915-1311)
        const uint32_t n_inst = (uint32_t)rs.meshes().size();
        // num_lights lags one update behind the area-light count (scene.cpp:893 uses the previous m_num_area_lights): a second,
        // table-less update reports the count the integrator would see from the second frame on
        RenderState rs2;
        rs2.setup(width, height, nullptr);
        scene->update(rs2);
        const uint32_t n_lights = rs2.num_lights();
        uint32_t       n_mats   = 0;
        for (auto* node : rs.meshes())
        {
            const glm::uvec2* pairs = (const glm::uvec2*)node->material_indices_buffer()->mapped_ptr();
            for (size_t g = 0; g < node->mesh()->sub_meshes().size(); g++) n_mats = std::max(n_mats, pairs[g].y + 1u);
        }
        FILE* f = std::fopen(argv[2], "wb");
        if (!f) throw std::runtime_error(std::string("cannot write ") + argv[2]);
        write_rows<Row80>(f, scene->m_material_data_buffer->mapped_ptr(), n_mats);
        write_rows<Row144>(f, scene->m_instance_data_buffer->mapped_ptr(), n_inst);
        write_rows<Row64>(f, scene->m_light_data_buffer->mapped_ptr(), n_lights);
        std::fwrite(&n_inst, 4, 1, f);
        for (auto* node : rs.meshes()) write_rows<glm::uvec2>(f, node->material_indices_buffer()->mapped_ptr(), (uint32_t)node->mesh()->sub_meshes().size());
        // (no push constants / sky coefficients here: path_integrator.cpp and the sky fit are pinned elsewhere) — zeros keep the format
        char zero[192 + 160] = { 0 };
        std::fwrite(zero, 1, sizeof(zero), f);
        // trailer: the texture array (descriptor set 4) as indices into the file's texture list, in array order
        const auto&    arr = ref_scene_stub::texture_array();
        const uint32_t nt  = (uint32_t)arr.size();
        std::fwrite(&nt, 4, 1, f);
        for (VkImageView v : arr)
        {
            uint32_t idx = 0xFFFFFFFFu;
            for (size_t k = 0; k < textures.size(); k++)
                if ((VkImageView)textures[k]->image_view().get() == v) idx = (uint32_t)k;
            std::fwrite(&idx, 4, 1, f);
        }
        std::fclose(f);
        std::printf("{\"instances\": %u, \"materials\": %u, \"lights\": %u, \"textures\": %u, \"num_area_lights\": %u}\n", n_inst, n_mats, n_lights, nt, scene->m_num_area_lights);
        scene.reset();
    }
    catch (const std::exception& e)
    {
        std::fprintf(stderr, "ref_scene_tool: %s\n", e.what());
        return 1;
    }
    return 0;
}
