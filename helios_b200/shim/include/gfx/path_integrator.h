// gfx/path_integrator.h — PathIntegrator (reference: include/gfx/path_integrator.h:9-60,
// src/engine/gfx/path_integrator.cpp:48-84 render, :125-200 launch_rays, :312-336 compute_tile_coords).
// Same public surface and the same sample / tile counters; launch_rays fills the 192-byte PushConstants block
// exactly as the reference does and calls hl_render_frame where the reference records vkCmdTraceRaysKHR.
#pragma once
#include <gfx/vk.h>
#include <resource/scene.h>
#include <vector>

namespace helios
{
class PathIntegrator
{
public:
    using Ptr = std::shared_ptr<PathIntegrator>;

    PathIntegrator(vk::Backend::Ptr backend);
    ~PathIntegrator();

    uint32_t max_ray_bounces() { return m_cfg.max_bounces; }
    uint32_t max_samples() { return m_cfg.max_samples; }
    uint32_t num_accumulated_samples() { return m_cfg.max_samples * m_bake.tile + m_bake.samples; }
    uint32_t num_target_samples() { return m_cfg.max_samples * (uint32_t)m_tiles.size(); }
    uint32_t tile_idx() { return m_bake.tile; }
    bool is_tiled() { return m_cfg.tiled; }
    float shadow_ray_bias() { return m_cfg.shadow_bias; }
    inline void restart_bake()
    {
        m_bake.samples = 0;
        m_bake.tile = 0;
    }
    void set_max_ray_bounces(const uint32_t& n) { m_cfg.max_bounces = n; }
    void set_max_samples(const uint32_t& n) { m_cfg.max_samples = n; }
    void set_shadow_ray_bias(const float& bias) { m_cfg.shadow_bias = bias; }

    void render(RenderState& render_state);
    // the ray debug view (reference: path_integrator.cpp:88-104): num_debug_rays paths through pixel_coord, one line segment
    // per secondary ray; the vertices land in `vertices` (appended, 2 per segment; capacity as the reference's buffer)
    void gather_debug_rays(const glm::ivec2& pixel_coord, const uint32_t& num_debug_rays, const glm::mat4& view, const glm::mat4& projection, RenderState& render_state,
                           std::vector<hl_debug_ray_vertex>& vertices, uint32_t max_vertices);
    void on_window_resize();
    void set_tiled(bool tiled);
    // samples-per-pixel sharding across GPUs (no counterpart in the reference): this integrator is rank
    // `rank` of `world` and renders frame indices rank + 1, rank + 1 + world, ... (frame 0 is discarded by the reference's
    // blend, path_trace_rgen.glsl:219-247, so it is skipped) into a per-GPU SUM image; MultiGpuRenderer combines the images.
    // world = 1 restores the reference's counting.  Full-frame launches only.
    void set_sample_sharding(uint32_t rank, uint32_t world);
    uint32_t shard_rank() const { return m_shard_rank; }
    uint32_t shard_world() const { return m_shard_world; }

    // Renderer::render hands the tone-map settings down so that the launch can resolve accumulation and tone map
    // in one fused pass (hl_render_frame_tonemapped); launched_last_render() tells it whether a launch happened
    void set_resolve_tone_map(bool enabled, float exposure, int tone_map_operator) { m_fuse_tone_map = enabled, m_fuse_exposure = exposure, m_fuse_operator = tone_map_operator; }
    bool launched_last_render() const { return m_launched; }

    // the block the last launch used (tests compare it with the Python host's restatement)
    const hl_push_constants& last_push_constants() const { return m_last_push_constants; }
    // fills a PushConstants block for (camera, counters) without launching
    hl_push_constants make_push_constants(RenderState& render_state, const glm::mat4& view, const glm::mat4& projection, const glm::ivec2& tile_coord, const glm::ivec2& pixel_coord);

private:
    void launch_rays(RenderState& render_state, const uint32_t& x, const uint32_t& y, const uint32_t& z, const glm::mat4& view, const glm::mat4& projection, const glm::ivec2& tile_coord,
                     const glm::ivec2& pixel_coord);
    void compute_tile_coords();
    // integrator settings (defaults of the reference: 7 bounces, 5000 samples, no bias, full-frame launches) and bake progress
    struct
    {
        bool tiled = false;
        uint32_t max_bounces = 7, max_samples = 5000;
        float shadow_bias = 0.0f;
    } m_cfg;
    struct
    {
        uint32_t samples = 0, tile = 0; // launches accumulated into the current tile; index of the tile being baked
    } m_bake;
    glm::uvec2 m_tile_extent;
    std::vector<glm::uvec2> m_tiles;
    std::weak_ptr<vk::Backend> m_backend;
    hl_push_constants m_last_push_constants {};
    uint32_t m_shard_rank = 0, m_shard_world = 1;
    bool m_fuse_tone_map = false, m_launched = false;
    float m_fuse_exposure = 1.0f;
    int m_fuse_operator = HL_TONE_MAP_ACES;
};
} // namespace helios
