// renderer.cpp — Renderer (reference: src/engine/gfx/renderer.cpp:106-330 render, :369-428 tone_map,
// :637-711 copy_and_save_tone_mapped_image, :715-726 on_window_resize).
#include <gfx/renderer.h>
#include <utility/png_writer.h>
#include <cstdio>

namespace helios
{
Renderer::Renderer(vk::Backend::Ptr backend) : m_backend(backend) { m_path_integrator = std::shared_ptr<PathIntegrator>(new PathIntegrator(backend)); }
Renderer::~Renderer() {}

void Renderer::render(RenderState& render_state)
{
    auto backend = m_backend.lock();
    if (!backend)
    {
        HELIOS_LOG_FATAL("Renderer::render: the backend was destroyed before the renderer");
        throw std::runtime_error("Renderer::render: the backend was destroyed before the renderer");
    }
    hl_context ctx = backend->require_device("Renderer::render");
    // (the top-level rebuild of :111-181 happened inside Scene::update -> hl_scene_set_tables)
    // restart branch, :205-223: any scene change, or a bake that is about to take its first sample
    if (m_output_image_recreated || render_state.m_scene_state != SCENE_STATE_READY || (m_path_integrator->num_accumulated_samples() == 0 && m_path_integrator->tile_idx() == 0))
    {
        backend->check(hl_accum_clear(ctx), "hl_accum_clear");
        m_output_image_recreated = false;
    }
    // the launch resolves accumulation + tone map itself (one fused pass for full-frame launches); only a frame
    // without a launch (bake complete) runs the stand-alone tone-map pass
    const int op = m_tone_map_operator == TONE_MAP_OPERATOR_ACES ? HL_TONE_MAP_ACES : HL_TONE_MAP_REINHARD;
    m_path_integrator->set_resolve_tone_map(true, m_exposure, op);
    if (render_state.m_scene) m_path_integrator->render(render_state);
    const bool resolved = render_state.m_scene && m_path_integrator->launched_last_render();
    if (m_ray_debug_view_added) // renderer.cpp:229-250
    {
        m_ray_debug_view_added = false;
        if (m_ray_debug_views.size() == 1) m_ray_debug_vertices.clear(); // the draw arguments are reset for the first view
        const auto& view = m_ray_debug_views.back();
        if (render_state.m_scene)
            m_path_integrator->gather_debug_rays(view.pixel_coord, view.num_debug_rays, view.view, view.projection, render_state, m_ray_debug_vertices, MAX_DEBUG_RAY_DRAW_COUNT * 2);
    }
    if (m_save_image_to_disk)
    {
        const auto           ext = backend->swap_chain_extents();
        std::vector<uint8_t> img((size_t)ext.width * ext.height * 4);
        if (resolved)
            backend->check(hl_read_rgba8(ctx, img.data()), "hl_read_rgba8");
        else
            tone_map(img.data());
        bool               ok  = false;
        const std::string& p   = m_image_save_path;
        auto               ext_is = [&](const char* e) { return p.size() > 4 && p.compare(p.size() - 4, 4, e) == 0; };
        if (ext_is(".pfm") || ext_is(".ppm"))
        {
            // headless extras: raw radiance (.pfm) / binary pixmap (.ppm)
            if (FILE* f = std::fopen(p.c_str(), "wb"))
            {
                if (ext_is(".pfm"))
                {
                    const std::vector<float> acc = read_accumulation();
                    std::fprintf(f, "PF\n%u %u\n-1.0\n", ext.width, ext.height);
                    // PFM rows run bottom-up, and so does the accumulation image (launch row 0 is the bottom of the
                    // view: the tone-map pass flips it, tone_map.frag + the negative-height viewport)
                    for (size_t i = 0; i < (size_t)ext.width * ext.height; i++) std::fwrite(&acc[i * 4], 4, 3, f);
                }
                else
                {
                    std::fprintf(f, "P6\n%u %u\n255\n", ext.width, ext.height);
                    for (size_t i = 0; i < (size_t)ext.width * ext.height; i++) std::fwrite(&img[i * 4], 1, 3, f);
                }
                ok = std::fclose(f) == 0;
            }
        }
        else
            // the reference's format, :651: 4-channel 8-bit PNG of the tone-mapped image
            ok = write_png_rgba8(p, ext.width, ext.height, img.data(), (size_t)ext.width * 4);
        if (!ok) HELIOS_LOG_ERROR("Renderer::save_image_to_disk: cannot write " + p);
        m_save_image_to_disk = false;
    }
    else if (!resolved)
        tone_map(nullptr);
}

void Renderer::tone_map(uint8_t* rgba8_host)
{
    auto backend = m_backend.lock();
    backend->check(hl_tonemap(backend->require_device("Renderer::tone_map"), m_exposure, m_tone_map_operator == TONE_MAP_OPERATOR_ACES ? HL_TONE_MAP_ACES : HL_TONE_MAP_REINHARD, 1.0f, rgba8_host),
                   "hl_tonemap");
}

void Renderer::on_window_resize()
{
    m_output_image_recreated = true;
    m_path_integrator->on_window_resize();
}

void Renderer::add_ray_debug_view(const glm::ivec2& pixel_coord, const uint32_t& num_debug_rays, const glm::mat4& view, const glm::mat4& projection)
{
    m_ray_debug_views.push_back({ pixel_coord, num_debug_rays, view, projection });
    m_ray_debug_view_added = true;
}

const std::vector<RayDebugView>& Renderer::ray_debug_views() { return m_ray_debug_views; }

void Renderer::clear_ray_debug_views() { m_ray_debug_views.clear(); }

void Renderer::save_image_to_disk(const std::string& path)
{
    if (path.length() == 0)
    {
        HELIOS_LOG_ERROR("A valid path is required to save an image to disk");
        return;
    }
    m_save_image_to_disk = true;
    m_image_save_path    = path;
}

std::vector<uint8_t> Renderer::read_tone_mapped_image()
{
    auto                 backend = m_backend.lock();
    const auto           ext     = backend->swap_chain_extents();
    std::vector<uint8_t> img((size_t)ext.width * ext.height * 4);
    tone_map(img.data());
    return img;
}

std::vector<float> Renderer::read_output_buffer(RenderState& render_state)
{
    if (m_current_output_buffer == OUTPUT_BUFFER_FINAL) return read_accumulation();
    auto backend = m_backend.lock();
    if (!render_state.camera())
    {
        HELIOS_LOG_FATAL("Renderer::read_output_buffer: the render state has no camera");
        throw std::runtime_error("Renderer::read_output_buffer: the render state has no camera");
    }
    const auto              ext = backend->swap_chain_extents();
    std::vector<float>      out((size_t)ext.width * ext.height * 4);
    const hl_push_constants pc  = m_path_integrator->make_push_constants(render_state, render_state.camera()->view_matrix(), render_state.camera()->projection_matrix(), glm::ivec2(0, 0), glm::ivec2(0, 0));
    static const int        map[] = { HL_OUTPUT_BUFFER_ALBEDO, HL_OUTPUT_BUFFER_NORMALS, HL_OUTPUT_BUFFER_ROUGHNESS, HL_OUTPUT_BUFFER_METALLIC, HL_OUTPUT_BUFFER_EMISSIVE };
    backend->check(hl_render_output_buffer(backend->require_device("Renderer::read_output_buffer"), &pc, map[m_current_output_buffer], out.data()), "hl_render_output_buffer");
    return out;
}

MultiGpuRenderer::MultiGpuRenderer(const std::vector<vk::Backend::Ptr>& backends) : m_backends(backends)
{
    if (backends.empty()) throw std::runtime_error("MultiGpuRenderer: no backend");
    std::vector<hl_context> ctxs;
    for (auto& b : m_backends) ctxs.push_back(b->require_device("MultiGpuRenderer"));
    if (hl_comm_init_all(ctxs.data(), (int)ctxs.size()) != HL_OK)
    {
        const std::string msg = std::string("hl_comm_init_all: ") + hl_comm_last_error();
        HELIOS_LOG_FATAL(msg);
        throw std::runtime_error(msg);
    }
    for (size_t g = 0; g < m_backends.size(); g++)
    {
        m_renderers.emplace_back(new Renderer(m_backends[g]));
        m_renderers.back()->path_integrator()->set_sample_sharding((uint32_t)g, (uint32_t)m_backends.size());
    }
}

std::vector<uint8_t> MultiGpuRenderer::resolve(uint32_t samples_total)
{
    std::vector<hl_context> ctxs;
    for (auto& b : m_backends) ctxs.push_back(b->require_device("MultiGpuRenderer::resolve"));
    const auto           ext = m_backends[0]->swap_chain_extents();
    std::vector<uint8_t> img((size_t)ext.width * ext.height * 4);
    Renderer*            r0 = m_renderers[0].get();
    const int            op = r0->tone_map_operator() == TONE_MAP_OPERATOR_ACES ? HL_TONE_MAP_ACES : HL_TONE_MAP_REINHARD;
    if (hl_multi_gpu_resolve(ctxs.data(), (int)ctxs.size(), 0, r0->exposure(), op, 1.0f / float(samples_total ? samples_total : 1u), img.data()) != HL_OK)
    {
        const std::string msg = std::string("hl_multi_gpu_resolve: ") + hl_comm_last_error();
        HELIOS_LOG_FATAL(msg);
        throw std::runtime_error(msg);
    }
    return img;
}

std::vector<float> Renderer::read_accumulation()
{
    auto               backend = m_backend.lock();
    const auto         ext     = backend->swap_chain_extents();
    std::vector<float> acc((size_t)ext.width * ext.height * 4);
    backend->check(hl_read_accum(backend->require_device("Renderer::read_accumulation"), acc.data()), "hl_read_accum");
    return acc;
}
} // namespace helios
