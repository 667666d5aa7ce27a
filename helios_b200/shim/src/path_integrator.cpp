// path_integrator.cpp — PathIntegrator (reference: src/engine/gfx/path_integrator.cpp).
#include <gfx/path_integrator.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>

namespace helios
{
#define TILE_SIZE 128

static_assert(sizeof(hl_push_constants) == 192, "PushConstants must match the 192-byte block of path_trace_rgen.glsl:93-110");

PathIntegrator::PathIntegrator(vk::Backend::Ptr backend) : m_backend(backend) { compute_tile_coords(); }
PathIntegrator::~PathIntegrator() {}

// :48-84 — restart on any scene change; one launch per call; max_samples per tile, then the next tile
void PathIntegrator::render(RenderState& render_state)
{
    if (render_state.scene_state() != SCENE_STATE_READY)
    {
        m_bake.tile                = 0;
        m_bake.samples = 0;
    }
    m_launched = false;
    if (m_bake.tile < m_tiles.size())
    {
        if (!render_state.camera())
        {
            HELIOS_LOG_ERROR("PathIntegrator::render: the scene has no enabled camera");
            return;
        }
        const glm::uvec2 tile = m_tiles[m_bake.tile];
        launch_rays(render_state, m_tile_extent.x, m_tile_extent.y, 1, render_state.camera()->view_matrix(), render_state.camera()->projection_matrix(), glm::ivec2((int)tile.x, (int)tile.y),
                    glm::ivec2(0, 0));
        m_bake.samples++;
        m_launched = true;
    }
    if (m_bake.samples == m_cfg.max_samples)
    {
        m_bake.samples = 0;
        m_bake.tile++;
    }
}

void PathIntegrator::on_window_resize()
{
    restart_bake();
    compute_tile_coords();
}

void PathIntegrator::set_sample_sharding(uint32_t rank, uint32_t world)
{
    if (world == 0 || rank >= world) throw std::runtime_error("PathIntegrator::set_sample_sharding: rank outside the world");
    if (world > 1 && m_cfg.tiled) throw std::runtime_error("PathIntegrator::set_sample_sharding: sharded bakes use full-frame launches");
    auto backend  = m_backend.lock();
    m_shard_rank  = rank;
    m_shard_world = world;
    backend->check(hl_set_accum_mode(backend->require_device("PathIntegrator::set_sample_sharding"), world > 1 ? HL_ACCUM_SUM : HL_ACCUM_RUNNING_MEAN), "hl_set_accum_mode");
    restart_bake();
}

void PathIntegrator::set_tiled(bool tiled)
{
    if (tiled && m_shard_world > 1) throw std::runtime_error("PathIntegrator::set_tiled: sharded bakes use full-frame launches");
    m_cfg.tiled = tiled;
    compute_tile_coords();
}

// :136-161
void PathIntegrator::gather_debug_rays(const glm::ivec2& pixel_coord, const uint32_t& num_debug_rays, const glm::mat4& view, const glm::mat4& projection, RenderState& render_state,
                                       std::vector<hl_debug_ray_vertex>& vertices, uint32_t max_vertices)
{
    auto backend = m_backend.lock();
    // launch_rays(render_state, ray-debug pipeline, num_debug_rays, 1, 1, view, projection, tile (0, 0), pixel_coord)
    const hl_push_constants pc   = make_push_constants(render_state, view, projection, glm::ivec2(0, 0), pixel_coord);
    const size_t            have = vertices.size();
    const uint32_t          room = have < max_vertices ? max_vertices - (uint32_t)have : 0;
    vertices.resize(have + room);
    uint32_t count = 0;
    backend->check(hl_gather_debug_rays(backend->require_device("PathIntegrator::gather_debug_rays"), &pc, num_debug_rays, vertices.data() + have, room, &count), "hl_gather_debug_rays");
    vertices.resize(have + std::min(count, room));
}

hl_push_constants PathIntegrator::make_push_constants(RenderState& render_state, const glm::mat4& view, const glm::mat4& projection, const glm::ivec2& tile_coord, const glm::ivec2& pixel_coord)
{
    auto           backend = m_backend.lock();
    const auto     extents = backend->swap_chain_extents();
    CameraNode*    cam     = render_state.camera();
    const glm::vec3 right = cam->left(), up = cam->up(), forward = -cam->forward(), camera_pos = cam->global_position();
    const glm::vec3 focal_point = camera_pos + forward * cam->focal_length();
    glm::vec4       focal_plane = glm::vec4(-forward, 0.0f);
    focal_plane.w               = -(focal_plane.x * focal_point.x + focal_plane.y * focal_point.y + focal_plane.z * focal_point.z);

    hl_push_constants pc;
    std::memset(&pc, 0, sizeof(pc));
    const glm::mat4 vpi = glm::inverse(projection * view);
    std::memcpy(pc.view_proj_inverse, glm::value_ptr(vpi), 64);
    pc.camera_pos[0] = camera_pos.x, pc.camera_pos[1] = camera_pos.y, pc.camera_pos[2] = camera_pos.z, pc.camera_pos[3] = 0.0f;
    pc.up_direction[0] = up.x, pc.up_direction[1] = up.y, pc.up_direction[2] = up.z;
    pc.right_direction[0] = right.x, pc.right_direction[1] = right.y, pc.right_direction[2] = right.z;
    pc.focal_plane[0] = focal_plane.x, pc.focal_plane[1] = focal_plane.y, pc.focal_plane[2] = focal_plane.z, pc.focal_plane[3] = focal_plane.w;
    pc.ray_debug_pixel_coord[0] = pixel_coord.x, pc.ray_debug_pixel_coord[1] = (int32_t)extents.height - pixel_coord.y;
    pc.ray_debug_pixel_coord[2] = (int32_t)extents.width, pc.ray_debug_pixel_coord[3] = (int32_t)extents.height;
    pc.launch_id_size[0] = (uint32_t)tile_coord.x, pc.launch_id_size[1] = (uint32_t)tile_coord.y, pc.launch_id_size[2] = extents.width, pc.launch_id_size[3] = extents.height;
    pc.num_lights      = render_state.num_lights();
    pc.num_frames      = m_shard_world > 1 ? 1u + m_shard_rank + m_bake.samples * m_shard_world : m_bake.samples;
    pc.accumulation    = float(pc.num_frames) / float(pc.num_frames + 1);
    pc.debug_vis       = 0;
    pc.max_ray_bounces = m_cfg.max_bounces;
    pc.shadow_ray_bias = m_cfg.shadow_bias;
    pc.focal_length    = cam->focal_length();
    pc.aperture_radius = cam->aperture_radius();
    return pc;
}

// :125-200 — vkCmdTraceRaysKHR(x, y, z) becomes hl_render_frame over the same launch rectangle
void PathIntegrator::launch_rays(RenderState& render_state, const uint32_t& x, const uint32_t& y, const uint32_t& z, const glm::mat4& view, const glm::mat4& projection, const glm::ivec2& tile_coord,
                                 const glm::ivec2& pixel_coord)
{
    (void)z;
    auto backend          = m_backend.lock();
    m_last_push_constants = make_push_constants(render_state, view, projection, tile_coord, pixel_coord);
    hl_context ctx = backend->require_device("PathIntegrator::launch_rays");
    if (m_fuse_tone_map)
        backend->check(hl_render_frame_tonemapped(ctx, &m_last_push_constants, x, y, m_fuse_exposure, m_fuse_operator), "hl_render_frame_tonemapped");
    else
        backend->check(hl_render_frame(ctx, &m_last_push_constants, x, y), "hl_render_frame");
}

// :312-336
void PathIntegrator::compute_tile_coords()
{
    auto       backend = m_backend.lock();
    const auto extents = backend->swap_chain_extents();
    m_tiles.clear();
    if (m_cfg.tiled)
    {
        const uint32_t nx = (uint32_t)ceilf(float(extents.width) / float(TILE_SIZE)), ny = (uint32_t)ceilf(float(extents.height) / float(TILE_SIZE));
        for (uint32_t x = 0; x < nx; x++)
            for (uint32_t y = 0; y < ny; y++) m_tiles.push_back(glm::uvec2(x * TILE_SIZE, y * TILE_SIZE));
        m_tile_extent = glm::uvec2(TILE_SIZE, TILE_SIZE);
    }
    else
    {
        m_tiles.push_back(glm::uvec2(0, 0));
        m_tile_extent = glm::uvec2(extents.width, extents.height);
    }
}
} // namespace helios
