// hl_debug.h — the ray debug view: PathIntegrator::gather_debug_rays (gfx/path_integrator.cpp:88-104) launches the
// RAY_DEBUG_VIEW variant of the pipeline (:259-307) num_debug_rays x 1 x 1 wide.  What that variant changes
// (#if defined(RAY_DEBUG_VIEW) in path_trace_rgen.glsl / _rchit.glsl / _rmiss.glsl):
//   * rgen:193-195  three draws right after rng_init colour the path: next_float * 0.5 + 0.5 per channel
//   * rgen:137-147  every path starts through pixel ray_debug_pixel_coord.xy (+ jitter), scaled by .zw, whatever its launch id
//   * rchit:500-509 no Russian roulette (no draw, no throughput rescale): a path only ends at max_ray_bounces or on a miss
//   * rchit:548-567 / rmiss:40-58  every ray after the primary one appends a line segment (two vertices: origin, and hit point
//     or origin + direction * tmax) to DebugRayVertexBuffer through atomicAdd(DebugRayDrawArgs.count, 2); radiance is unused
// One path per lane, run to completion (the launch is at most a few thousand paths): extend, shade, extend, ...
#pragma once
#include "hl_bvh.h"
#include "hl_camera.h"
#include "hl_shade.h"

namespace hl
{
// A = appender: uint32_t A::alloc2() returns the index of the first of two consecutive vertices (may exceed the
// capacity: A::put then drops the vertex, the count keeps growing like the reference's draw argument).
template <class A>
HL_HD void debug_ray_path(const SceneView& s, const hl_push_constants& pc, uint32_t i, bool active, TravStack& st, A& out)
{
    const uint32_t lx = pc.launch_id_size[0] + i, ly = pc.launch_id_size[1];   // rgen:182
    bool           alive = active && lx < pc.launch_id_size[2] && ly < pc.launch_id_size[3]; // rgen:185
    Rng            rng   = rng_seed(lx, ly, pc.num_frames);
    f3             color;
    color.x = rand01(rng) * 0.5f + 0.5f;
    color.y = rand01(rng) * 0.5f + 0.5f;
    color.z = rand01(rng) * 0.5f + 0.5f;
    f3 o, d;
    primary_ray_at(pc, (float)pc.ray_debug_pixel_coord[0] + 0.5f, (float)pc.ray_debug_pixel_coord[1] + 0.5f, (float)pc.ray_debug_pixel_coord[2], (float)pc.ray_debug_pixel_coord[3], rng, o, d);
    ShadeParams prm;
    prm.num_lights = pc.num_lights, prm.max_ray_bounces = pc.max_ray_bounces, prm.shadow_ray_bias = pc.shadow_ray_bias, prm.ray_debug_view = 1;
    f3       T     = mk3(1.0f);
    uint32_t depth = 0;
    for (;;)
    {
        if (HL_WARP_BALLOT(alive) == 0) break;
        Hit h;
        trace_ray(s, alive, o, depth == 0 ? 0.001f : 0.0001f, d, 10000.0f, depth == 0 ? 0u : HL_RAY_OPAQUE, h, st);
        if (!alive) continue;
        const bool hit = h.instance != HL_MISS;
        if (depth > 0)
        {
            const uint32_t k   = out.alloc2();
            const f3       end = o + d * (hit ? h.t : 10000.0f);
            out.put(k, o, color), out.put(k + 1, end, color);
        }
        if (!hit)
        {
            alive = false;
            continue;
        }
        ShadeResult r;
        shade_hit(s, prm, depth, d, h, T, rng, r); // direct lighting still draws its random numbers; its shadow ray changes nothing here
        alive = r.continues;
        if (alive) o = r.next_o, d = r.next_d, T = r.T, depth++;
    }
}
} // namespace hl
