// hl_camera.h — the generate stage: per-pixel RNG seeding and the jittered thin-lens primary ray.
// Restates path_trace_rgen.glsl:132-174 (generate_ray) and :180-197 (payload init) of the reference.
#pragma once
#include "hl_rng.h"
#include "hl_scene.h"

namespace hl
{
// cx, cy = pixel centre, fw, fh = the extent the jittered coordinate is divided by.  Consumes 4 draws: jitter x, jitter y,
// lens angle, lens radius.
HL_HD void primary_ray_at(const hl_push_constants& pc, float cx, float cy, float fw, float fh, Rng& rng, f3& origin, f3& direction)
{
    const float jx = rand01(rng);
    const float jy = rand01(rng);
    const float u  = (cx + jx) / fw; // samples cover [x+0.5, x+1.5): SURVEY A.8-9
    const float v  = (cy + jy) / fh;
    const f3    cam = mk3(pc.camera_pos);
    f4          tgt = mat4_mul(pc.view_proj_inverse, mk4(u * 2.0f - 1.0f, v * 2.0f - 1.0f, 0.0f, 1.0f));
    const f3    target = mk3(tgt.x / tgt.w, tgt.y / tgt.w, tgt.z / tgt.w);
    const float angle  = rand01(rng) * 2.0f * 3.14159265359f;
    const float radius = sqrtf(rand01(rng));
    const float ox = cosf(angle) * radius * pc.aperture_radius;
    const float oy = sinf(angle) * radius * pc.aperture_radius;
    const f3    lens = cam + mk3(pc.right_direction) * ox + mk3(pc.up_direction) * oy;
    const f3    rdir = -normalize(target - cam);
    const f3    plane = mk3(pc.focal_plane);
    const float t     = -(dot(cam, plane) + pc.focal_plane[3]) / dot(rdir, plane);
    const f3    focus = cam + rdir * t;
    origin    = lens;
    direction = normalize(focus - lens);
}
// px, py = absolute pixel (tile offset + launch id)
HL_HD void primary_ray(const hl_push_constants& pc, uint32_t px, uint32_t py, Rng& rng, f3& origin, f3& direction)
{
    primary_ray_at(pc, (float)px + 0.5f, (float)py + 0.5f, (float)pc.launch_id_size[2], (float)pc.launch_id_size[3], rng, origin, direction);
}
} // namespace hl
