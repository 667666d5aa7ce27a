// helios_headless — headless driver of the path-trace pass through the engine's own interfaces
// (SURVEY.md §7 step 3).  It loads a scene description (helios_b200/scene_io.py), rebuilds it with
// Texture2D / Material / Mesh / MeshNode / ...LightNode / CameraNode / IBLNode, and runs the reference's frame
// loop (src/viewer/main.cpp:66-76):  render_state.setup(w, h, cmd) -> scene->update(render_state) ->
// renderer->render(render_state), once per sample.
//
//   helios_headless --ast-scene scene.json [--asset-root DIR] [--width W] [--height H] [--focal-length F] [--aperture A] ...
//       loads an AssetCore scene description (scene JSON -> mesh .ast -> material JSON -> image .ast) through
//       ResourceManager::load_scene, the reference's own asset route
//   helios_headless --scene file.hlsc [--spp N] [--device D] [--tiled] [--bounces B] [--exposure E]
//                   [--out image.png|.ppm|.pfm] [--dump-accum raw.f32] [--dump-tables tables.bin] [--no-device]
//
// --no-device builds the scene graph and the tables on the host only (for inspection); rendering needs a GPU.
#include <core/resource_manager.h>
#include <gfx/renderer.h>
#include <utility/png_writer.h>
#include <resource/material.h>
#include <resource/mesh.h>
#include <resource/scene.h>
#include <resource/texture.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

using namespace helios;

namespace
{
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
void write_vec(FILE* f, const std::vector<T>& v)
{
    const uint32_t n = (uint32_t)v.size();
    std::fwrite(&n, 4, 1, f);
    if (n) std::fwrite(v.data(), sizeof(T), n, f);
}
} // namespace

int main(int argc, char** argv)
{
    std::string scene_path, ast_scene_path, asset_root, out_path, accum_path, tables_path, output_buffer, output_buffer_path, ray_debug_path;
    int         ray_debug[3] = { 0, 0, 0 };
    uint32_t    spp = 16, bounces = 0, arg_width = 1280, arg_height = 720;
    float       focal_length = -1.0f, aperture = -1.0f;
    int         device = 0;
    std::vector<int> devices; // --gpus N / --devices a,b,...: one rank per entry (entries may repeat: ranks sharing a GPU)
    bool        tiled = false, no_device = false;
    int         jitter_seed = -1, jitter_after = 0; // --jitter-nodes SEED [--jitter-after K]: move every second mesh node after K frames
    float       exposure = 1.0f;
    for (int i = 1; i < argc; i++)
    {
        const std::string a = argv[i];
        auto              next = [&]() -> std::string {
            if (i + 1 >= argc) throw std::runtime_error("missing value for " + a);
            return argv[++i];
        };
        try
        {
            if (a == "--scene") scene_path = next();
            else if (a == "--ast-scene") ast_scene_path = next();
            else if (a == "--asset-root") asset_root = next();
            else if (a == "--width") arg_width = (uint32_t)std::stoul(next());
            else if (a == "--height") arg_height = (uint32_t)std::stoul(next());
            else if (a == "--focal-length") focal_length = std::stof(next());
            else if (a == "--aperture") aperture = std::stof(next());
            else if (a == "--spp") spp = (uint32_t)std::stoul(next());
            else if (a == "--device") device = std::stoi(next());
            else if (a == "--gpus") // samples sharded over GPUs 0..N-1 of this process, one reduction at the end
            {
                const int n = std::stoi(next());
                devices.clear();
                for (int g = 0; g < n; g++) devices.push_back(g);
            }
            else if (a == "--devices")
            {
                devices.clear();
                const std::string list = next();
                for (size_t b = 0; b < list.size();)
                {
                    const size_t e = list.find(',', b) == std::string::npos ? list.size() : list.find(',', b);
                    devices.push_back(std::stoi(list.substr(b, e - b)));
                    b = e + 1;
                }
            }
            else if (a == "--bounces") bounces = (uint32_t)std::stoul(next());
            else if (a == "--exposure") exposure = std::stof(next());
            else if (a == "--out") out_path = next();
            else if (a == "--dump-accum") accum_path = next();
            else if (a == "--dump-tables") tables_path = next();
            else if (a == "--output-buffer") output_buffer = next();           // albedo | normals | roughness | metallic | emissive
            else if (a == "--dump-output-buffer") output_buffer_path = next(); // raw RGBA32F of that debug view
            else if (a == "--ray-debug") // x y n : ray debug view through pixel (x, y) with n paths, gathered after the last frame
            {
                ray_debug[0] = std::atoi(next().c_str()), ray_debug[1] = std::atoi(next().c_str()), ray_debug[2] = std::atoi(next().c_str());
            }
            else if (a == "--dump-ray-debug") ray_debug_path = next(); // raw vertices, 8 floats each
            else if (a == "--jitter-nodes") jitter_seed = std::stoi(next()); // moves mesh nodes (seeded offsets): Scene::update turns a transform-only change into an instance-tree refit
            else if (a == "--jitter-after") jitter_after = std::stoi(next()); // ... after this many frames (0 = before the first one); the bake restarts, --spp frames follow
            else if (a == "--tiled") tiled = true;
            else if (a == "--no-device") no_device = true;
            else
            {
                std::fprintf(stderr, "unknown argument %s\n", a.c_str());
                return 2;
            }
        }
        catch (const std::exception& e)
        {
            std::fprintf(stderr, "%s\n", e.what());
            return 2;
        }
    }
    if (scene_path.empty() == ast_scene_path.empty())
    {
        std::fprintf(stderr, "usage: helios_headless (--scene file.hlsc | --ast-scene scene.json [--asset-root DIR] [--width W] [--height H]) [--spp N] [--out image.png] ...\n");
        return 2;
    }
    try
    {
      uint32_t         width = arg_width, height = arg_height, file_bounces = 8;
      float            bias  = 0.0f;
      vk::Backend::Ptr backend;
      Scene::Ptr       scene;
      std::vector<uint32_t> file_texture_ids; // Texture2D id of the k-th texture of the scene file (--dump-tables trailer)
      // builds backend + scene on GPU `device` (called once per rank for --gpus / --devices: the scene is replicated)
      auto make_scene = [&](int device) {
      if (!ast_scene_path.empty())
      {
        // the reference's own asset route: ResourceManager::load_scene on an AssetCore scene description
        // (src/viewer/main.cpp: m_resource_manager->load_scene(path)); camera lens values are not serialised
        backend = no_device ? vk::Backend::create_without_device(width, height) : vk::Backend::create(device, width, height);
        ResourceManager rm(backend);
        if (!asset_root.empty()) rm.set_asset_root(asset_root);
        scene = rm.load_scene(ast_scene_path);
        if (!scene) throw std::runtime_error("cannot load scene " + ast_scene_path);
        if (auto camera = scene->find_camera())
        {
            if (focal_length >= 0.0f) camera->set_focal_length(focal_length);
            if (aperture >= 0.0f) camera->set_aperture_radius(aperture);
        }
      }
      else
      {
        Reader r(scene_path);
        char   magic[8];
        r.raw(magic, 8);
        if (std::memcmp(magic, "HLSC0001", 8) != 0) throw std::runtime_error("not a HLSC0001 scene file");
        width = r.get<uint32_t>(), height = r.get<uint32_t>(), file_bounces = r.get<uint32_t>();
        bias  = r.get<float>();

        backend = no_device ? vk::Backend::create_without_device(width, height) : vk::Backend::create(device, width, height);
        vk::BatchUploader uploader(backend);

        std::vector<Texture2D::Ptr> textures(r.get<uint32_t>());
        for (auto& t : textures)
        {
            const int32_t        fmt = r.get<int32_t>();
            const uint32_t       w = r.get<uint32_t>(), h = r.get<uint32_t>();
            std::vector<uint8_t> texels((size_t)w * h * (fmt == HL_TEX_RGBA32F ? 16 : 4));
            r.raw(texels.data(), texels.size());
            t = Texture2D::create(backend, fmt, w, h, texels.data(), "texture");
            file_texture_ids.push_back(t->id());
        }
        std::vector<Material::Ptr> materials(r.get<uint32_t>());
        for (size_t k = 0; k < materials.size(); k++)
        {
            const uint32_t type = r.get<uint32_t>(), alpha_test = r.get<uint32_t>();
            glm::vec4      albedo, emissive;
            r.raw(&albedo, 16), r.raw(&emissive, 16);
            const float   metallic = r.get<float>(), roughness = r.get<float>();
            int32_t       tex[5];
            r.raw(tex, sizeof(tex));
            const int32_t rough_ch = r.get<int32_t>(), metal_ch = r.get<int32_t>();
            // each Material owns a local texture list; TextureInfo::array_index points into it
            std::vector<Texture2D::Ptr> local;
            TextureInfo                 info[5];
            for (int s = 0; s < 5; s++)
                if (tex[s] >= 0)
                {
                    info[s].array_index = (int32_t)local.size();
                    local.push_back(textures.at((size_t)tex[s]));
                }
            info[2].channel_index = metal_ch, info[3].channel_index = rough_ch;
            materials[k] = Material::create(backend, type == 0 ? MATERIAL_OPAQUE : MATERIAL_TRANSPARENT, local, info[0], info[1], info[2], info[3], info[4], albedo, emissive, metallic, roughness,
                                            alpha_test != 0, "material" + std::to_string(k));
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
        uploader.submit();

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
            std::vector<float> faces((size_t)6 * cube * cube * 4);
            r.raw(faces.data(), faces.size() * 4);
            auto ibl = std::make_shared<IBLNode>("ibl");
            root->add_child(ibl);
            ibl->set_image(TextureCube::create(backend, cube, faces.data(), "ibl"));
        }

        scene = Scene::create(backend, "scene", root, scene_path);
      }
      };
      if (devices.size() > 1)
      {
        // ---- samples-per-pixel sharding: one Backend / Scene / Renderer per rank, the reference's frame loop on each, one reduction
        if (no_device || tiled) throw std::runtime_error("--gpus / --devices needs a device and full-frame launches");
        std::vector<vk::Backend::Ptr> backends;
        std::vector<Scene::Ptr>       scenes;
        for (int d : devices)
        {
            make_scene(d);
            backends.push_back(backend), scenes.push_back(scene);
        }
        backend.reset(), scene.reset();
        const uint32_t G = (uint32_t)devices.size();
        if (spp % G != 0) throw std::runtime_error("--spp must be divisible by the number of ranks");
        {
            MultiGpuRenderer         group(backends);
            std::vector<RenderState> states(G);
            auto                     cmd = std::make_shared<vk::CommandBuffer>();
            for (uint32_t g = 0; g < G; g++)
            {
                auto pi = group.renderer(g)->path_integrator();
                pi->set_max_ray_bounces(bounces ? bounces : file_bounces);
                pi->set_shadow_ray_bias(bias);
                group.renderer(g)->set_exposure(exposure);
            }
            const auto t0 = std::chrono::steady_clock::now();
            for (uint32_t k = 0; k < spp / G; k++)
                for (uint32_t g = 0; g < G; g++) // asynchronous launches: one host thread keeps every GPU busy
                {
                    states[g].setup(width, height, cmd);
                    scenes[g]->update(states[g]);
                    group.renderer(g)->render(states[g]);
                }
            const std::vector<uint8_t> img = group.resolve(spp);
            const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (!out_path.empty())
            {
                const std::string& p = out_path;
                const bool ppm = p.size() > 4 && p.compare(p.size() - 4, 4, ".ppm") == 0;
                if (ppm)
                {
                    if (FILE* f = std::fopen(p.c_str(), "wb"))
                    {
                        std::fprintf(f, "P6\n%u %u\n255\n", width, height);
                        for (size_t i = 0; i < (size_t)width * height; i++) std::fwrite(&img[i * 4], 1, 3, f);
                        std::fclose(f);
                    }
                }
                else if (!write_png_rgba8(p, width, height, img.data(), (size_t)width * 4))
                    throw std::runtime_error("cannot write " + p);
            }
            if (!accum_path.empty())
            {
                const auto acc = group.renderer(0)->read_accumulation(); // the reduced SUM (divide by --spp for radiance)
                if (FILE* f = std::fopen(accum_path.c_str(), "wb"))
                {
                    std::fwrite(acc.data(), 4, acc.size(), f);
                    std::fclose(f);
                }
            }
            unsigned long long ext = 0, sh = 0;
            for (uint32_t g = 0; g < G; g++)
            {
                hl_counters c {};
                backends[g]->check(hl_get_counters(backends[g]->context(), &c), "hl_get_counters");
                ext += c.extension_rays, sh += c.shadow_rays;
            }
            std::printf("{\"scene\": \"%s\", \"width\": %u, \"height\": %u, \"ranks\": %u, \"launches\": %u, \"seconds\": %.6f, \"extension_rays\": %llu, \"shadow_rays\": %llu, \"mrays_per_s\": %.1f}\n",
                        (ast_scene_path.empty() ? scene_path : ast_scene_path).c_str(), width, height, G, spp, seconds, ext, sh, seconds > 0 ? double(ext + sh) / seconds / 1e6 : 0.0);
        }
        scenes.clear();
        return 0;
      }
      make_scene(devices.empty() ? device : devices[0]);
      if (scene_path.empty()) scene_path = ast_scene_path;
        RenderState render_state;
        auto        cmd = std::make_shared<vk::CommandBuffer>();

        if (no_device)
        {
            render_state.setup(width, height, cmd);
            scene->update(render_state);
        }
        std::unique_ptr<Renderer> renderer;
        double                    seconds = 0.0;
        if (!no_device)
        {
            renderer.reset(new Renderer(backend));
            renderer->path_integrator()->set_max_ray_bounces(bounces ? bounces : file_bounces);
            renderer->path_integrator()->set_shadow_ray_bias(bias);
            renderer->path_integrator()->set_tiled(tiled);
            renderer->set_exposure(exposure);
            if (tiled) renderer->path_integrator()->set_max_samples(spp); // spp per tile, tile after tile (path_integrator.cpp:48-84)
            auto       pi   = renderer->path_integrator();
            uint32_t   done = 0;
            const auto t0   = std::chrono::steady_clock::now();
            auto finished   = [&]() { return tiled ? (done > 0 && pi->tile_idx() * pi->max_samples() >= pi->num_target_samples()) : done == spp; };
            int  frames_before_jitter = jitter_seed >= 0 ? jitter_after : -1;
            while (!finished())
            {
                if (frames_before_jitter == 0)
                {
                    // seeded offsets for every second mesh node; set_position marks the transforms dirty, the next Scene::update
                    // reports a hierarchy change (the bake restarts) and only the instance matrices differ
                    uint32_t k = 0, state = 0x9E3779B9u * (uint32_t)(jitter_seed + 1);
                    auto     rnd = [&]() { state = state * 1664525u + 1013904223u; return float(state >> 8) / 16777216.0f - 0.5f; };
                    if (scene->root_node())
                        for (auto& child : scene->root_node()->children())
                            if (child->type() == NODE_MESH && (k++ & 1u))
                            {
                                auto node = std::static_pointer_cast<MeshNode>(child);
                                node->set_position(node->local_position() + glm::vec3(rnd() * 3.0f, rnd() * 0.5f, rnd() * 3.0f));
                            }
                    done = 0;
                }
                frames_before_jitter--;
                if (!tiled && !out_path.empty() && done + 1 == spp) renderer->save_image_to_disk(out_path);
                render_state.setup(width, height, cmd);
                scene->update(render_state);
                renderer->render(render_state);
                done++;
            }
            backend->wait_idle();
            seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (tiled && !out_path.empty())
            {
                // one more frame would restart nothing: the bake is complete, so only tone map + save
                const auto img = renderer->read_tone_mapped_image();
                if (FILE* f = std::fopen(out_path.c_str(), "wb"))
                {
                    std::fprintf(f, "P6\n%u %u\n255\n", width, height);
                    for (size_t i = 0; i < (size_t)width * height; i++) std::fwrite(&img[i * 4], 1, 3, f);
                    std::fclose(f);
                }
            }
            if (!accum_path.empty())
            {
                const auto acc = renderer->read_accumulation();
                if (FILE* f = std::fopen(accum_path.c_str(), "wb"))
                {
                    std::fwrite(acc.data(), 4, acc.size(), f);
                    std::fclose(f);
                }
            }
            if (ray_debug[2] > 0 && !ray_debug_path.empty())
            {
                // the editor's click (src/editor/main.cpp:516-523): add a view with the camera's matrices; the next render() gathers it
                renderer->add_ray_debug_view(glm::ivec2(ray_debug[0], ray_debug[1]), (uint32_t)ray_debug[2], render_state.camera()->view_matrix(), render_state.camera()->projection_matrix());
                render_state.setup(width, height, cmd);
                scene->update(render_state);
                renderer->render(render_state);
                const auto& v = renderer->ray_debug_vertices();
                if (FILE* f = std::fopen(ray_debug_path.c_str(), "wb"))
                {
                    std::fwrite(v.data(), sizeof(hl_debug_ray_vertex), v.size(), f);
                    std::fclose(f);
                }
            }
            if (!output_buffer.empty() && !output_buffer_path.empty())
            {
                static const char* names[] = { "albedo", "normals", "roughness", "metallic", "emissive" };
                int                which   = -1;
                for (int k = 0; k < 5; k++)
                    if (output_buffer == names[k]) which = k;
                if (which < 0) throw std::runtime_error("unknown output buffer " + output_buffer);
                renderer->set_current_output_buffer((OutputBuffer)which);
                const auto img = renderer->read_output_buffer(render_state);
                renderer->set_current_output_buffer(OUTPUT_BUFFER_FINAL);
                if (FILE* f = std::fopen(output_buffer_path.c_str(), "wb"))
                {
                    std::fwrite(img.data(), 4, img.size(), f);
                    std::fclose(f);
                }
            }
            hl_counters c {};
            backend->check(hl_get_counters(backend->context(), &c), "hl_get_counters");
            std::printf("{\"scene\": \"%s\", \"width\": %u, \"height\": %u, \"launches\": %u, \"seconds\": %.6f, \"extension_rays\": %llu, \"shadow_rays\": %llu, \"mrays_per_s\": %.1f}\n",
                        scene_path.c_str(), width, height, done, seconds, (unsigned long long)c.extension_rays, (unsigned long long)c.shadow_rays,
                        seconds > 0 ? double(c.extension_rays + c.shadow_rays) / seconds / 1e6 : 0.0);
        }
        if (!tables_path.empty())
        {
            // materials, instances, lights, per-instance submesh pairs, last push constants, sky coefficients
            const SceneTables& T = scene->tables();
            FILE*              f = std::fopen(tables_path.c_str(), "wb");
            if (!f) throw std::runtime_error("cannot write " + tables_path);
            write_vec(f, T.materials), write_vec(f, T.instances), write_vec(f, T.lights);
            const uint32_t ni = (uint32_t)T.submesh_info.size();
            std::fwrite(&ni, 4, 1, f);
            for (auto& v : T.submesh_info) write_vec(f, v);
            hl_push_constants pc;
            if (renderer)
                pc = renderer->path_integrator()->last_push_constants();
            else
            {
                PathIntegrator pi(backend);
                pi.set_max_ray_bounces(bounces ? bounces : file_bounces), pi.set_shadow_ray_bias(bias);
                pc = pi.make_push_constants(render_state, render_state.camera()->view_matrix(), render_state.camera()->projection_matrix(), glm::ivec2(0, 0), glm::ivec2(0, 0));
            }
            std::fwrite(&pc, sizeof(pc), 1, f);
            std::fwrite(scene->sky_model()->coefficients(), 4, 40, f);
            // trailer: the texture array as indices into the scene file's texture list (file order = creation order = ascending id)
            std::vector<uint32_t> ids = scene->texture_array_ids(), sorted_ids = file_texture_ids;
            const uint32_t        nt  = (uint32_t)ids.size();
            std::fwrite(&nt, 4, 1, f);
            for (uint32_t id : ids)
            {
                uint32_t idx = 0xFFFFFFFFu;
                for (size_t k = 0; k < sorted_ids.size(); k++)
                    if (sorted_ids[k] == id) idx = (uint32_t)k;
                std::fwrite(&idx, 4, 1, f);
            }
            std::fclose(f);
        }
        // release in dependency order: scene graph and resources before the backend
        renderer.reset();
        scene.reset();
    }
    catch (const std::exception& e)
    {
        std::fprintf(stderr, "helios_headless: %s\n", e.what());
        return 1;
    }
    return 0;
}
