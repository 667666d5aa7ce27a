// gfx/renderer.h — Renderer (reference: include/gfx/renderer.h:35-120, src/engine/gfx/renderer.cpp:106-330
// render, :369-428 tone_map, :637-711 save path).  render(RenderState&) = (re)build the top level when the
// hierarchy changed, clear the accumulation when the bake restarts, one PathIntegrator iteration, tone map.
// The swap-chain copy and ImGui of the reference are outside the path; its ray-debug and G-buffer debug views are
// offered as data (segment vertices / channel images) instead of rasterised overlays.
#pragma once
#include <gfx/path_integrator.h>
#include <resource/scene.h>
#include <string>
#include <vector>

namespace helios
{
#define MAX_DEBUG_RAY_DRAW_COUNT 1024 // include/gfx/renderer.h:9

struct RayDebugView // include/gfx/renderer.h:13-19
{
    glm::ivec2 pixel_coord;
    uint32_t num_debug_rays;
    glm::mat4 view;
    glm::mat4 projection;
};

enum ToneMapOperator
{
    TONE_MAP_OPERATOR_ACES,
    TONE_MAP_OPERATOR_REINHARD
};

enum OutputBuffer
{
    OUTPUT_BUFFER_ALBEDO,
    OUTPUT_BUFFER_NORMALS,
    OUTPUT_BUFFER_ROUGHNESS,
    OUTPUT_BUFFER_METALLIC,
    OUTPUT_BUFFER_EMISSIVE,
    OUTPUT_BUFFER_FINAL
};

class Renderer
{
public:
    Renderer(vk::Backend::Ptr backend);
    ~Renderer();

    void set_tone_map_operator(const ToneMapOperator& tone_map) { m_tone_map_operator = tone_map; }
    void set_exposure(const float& exposure) { m_exposure = exposure; }
    void set_current_output_buffer(OutputBuffer buffer) { m_current_output_buffer = buffer; }
    PathIntegrator::Ptr path_integrator() { return m_path_integrator; }
    ToneMapOperator tone_map_operator() { return m_tone_map_operator; }
    OutputBuffer current_output_buffer() { return m_current_output_buffer; }
    float exposure() { return m_exposure; }

    void render(RenderState& render_state);
    void on_window_resize();
    // ray debug views (reference: renderer.cpp:229-250, :733-751): the view added last is gathered by the next render()
    // (PathIntegrator::gather_debug_rays); the first view after a clear resets the segment buffer, later ones append.
    void add_ray_debug_view(const glm::ivec2& pixel_coord, const uint32_t& num_debug_rays, const glm::mat4& view, const glm::mat4& projection);
    const std::vector<RayDebugView>& ray_debug_views();
    void clear_ray_debug_views();
    const std::vector<hl_debug_ray_vertex>& ray_debug_vertices() const { return m_ray_debug_vertices; } // what the reference draws as a line list
    // queues a save of the tone-mapped image; it is written at the end of the next render(), as in the reference
    // (8-bit RGBA PNG as in the reference, :651; a path ending in .ppm / .pfm selects those formats instead)
    void save_image_to_disk(const std::string& path);

    // headless read-backs (the reference presents to a swap chain instead)
    std::vector<uint8_t> read_tone_mapped_image(); // RGBA8, row 0 = top
    std::vector<float> read_accumulation(); // RGBA32F, row 0 = v 0
    // the debug view selected with set_current_output_buffer (albedo / normals / roughness / metallic / emissive of
    // the surfaces the camera sees; reference: renderer.cpp:459-547 + debug_visualization.frag), RGBA32F, row 0 = v 0.
    // OUTPUT_BUFFER_FINAL returns the accumulation image.
    std::vector<float> read_output_buffer(RenderState& render_state);

private:
    void tone_map(uint8_t* rgba8_host);

    std::vector<RayDebugView> m_ray_debug_views;
    bool m_ray_debug_view_added = false;
    std::vector<hl_debug_ray_vertex> m_ray_debug_vertices;
    std::weak_ptr<vk::Backend> m_backend;
    PathIntegrator::Ptr m_path_integrator;
    bool m_output_image_recreated = true;
    bool m_save_image_to_disk = false;
    std::string                m_image_save_path        = "";
    ToneMapOperator m_tone_map_operator = TONE_MAP_OPERATOR_ACES;
    float m_exposure = 1.0f;
    OutputBuffer m_current_output_buffer = OUTPUT_BUFFER_FINAL;
};

// Samples-per-pixel sharding across the GPUs of ONE process (the reference renders on one GPU): one
// Backend / Scene / Renderer per GPU, the scene replicated, renderer g set to shard (g, n) — its frame loop is the
// reference's own (scene->update, renderer->render) — and ONE reduction per finished picture: hl_multi_gpu_resolve sums the
// per-GPU images over peer memory (or NCCL), applies 1 / samples, exposure and the tone map, and hands back the image.
class MultiGpuRenderer
{
public:
    MultiGpuRenderer(const std::vector<vk::Backend::Ptr>& backends); // hl_comm_init_all; creates one Renderer per backend
    size_t size() const { return m_renderers.size(); }
    Renderer* renderer(size_t rank) { return m_renderers[rank].get(); }
    // RGBA8 image (row 0 = top) of sum / samples_total with renderer(0)'s exposure and tone map operator; the sum stays
    // in rank 0's accumulation image (renderer(0)->read_accumulation())
    std::vector<uint8_t> resolve(uint32_t samples_total);

private:
    std::vector<vk::Backend::Ptr> m_backends;
    std::vector<std::unique_ptr<Renderer>> m_renderers;
};
} // namespace helios
