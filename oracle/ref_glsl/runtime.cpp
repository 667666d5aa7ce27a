// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path.
//
// runtime.cpp — the "Vulkan implementation" under the reference's shaders when they run on the CPU
// (oracle/_ref/libhelios_glsl_ref.so).  The shader stages are the REFERENCE'S OWN GLSL (stage_*.cpp include the
// files gen.py derives from /root/reference/src/engine/shader at build time); this file supplies only what the
// reference leaves to the driver and reuses the restatement's definitions for it, so both paths share one
// "driver" and differ only in the shader logic under test:
//   * acceleration structure + ray/triangle test + closest-hit ordering  -> trace()        (../helios_oracle.cpp)
//   * sampler2D / samplerCube filtering                                  -> texture_lod0(), env_sample()
//   * shader binding table dispatch (path_integrator.cpp:221-225)        -> ref_drv_trace() below
//   * vkCmdTraceRaysKHR over a launch rectangle (path_integrator.cpp:125-200) -> ref_render_frame()
// The library also contains the whole restatement (or_* entry points), so one scene object serves both.
#include "../helios_oracle.cpp"
#include "ref_abi.h"

static const Scene*          g_ref_scene = nullptr;
static thread_local uint64_t t_ref_extension_rays = 0, t_ref_shadow_rays = 0;

extern "C" void ref_drv_texture2d(int index, float u, float v, float* out4)
{
    vec4 c  = texture_lod0(g_ref_scene->textures[index], u, v);
    out4[0] = c.x, out4[1] = c.y, out4[2] = c.z, out4[3] = c.w;
}
extern "C" void ref_drv_texture_cube(const float* dir3, float* out4)
{
    vec3 c  = env_sample(g_ref_scene->env, vec3(dir3[0], dir3[1], dir3[2]));
    out4[0] = c.x, out4[1] = c.y, out4[2] = c.z, out4[3] = 1.0f;
}
static bool ref_any_hit(uint32_t inst, uint32_t geom, uint32_t prim, float bu, float bv)
{
    RefHit h { 0.0f, bu, bv, inst, geom, prim };
    return ref_rahit_invoke(&h) != 0;
}
extern "C" void ref_drv_trace(uint32_t flags, uint32_t sbt_offset, uint32_t miss_index, const float* o, float tmin, const float* d, float tmax, void* payload)
{
    (sbt_offset == 0 ? t_ref_extension_rays : t_ref_shadow_rays)++;
    t_any_hit_override = ref_any_hit;
    const Hit h        = trace(*g_ref_scene, vec3(o[0], o[1], o[2]), tmin, vec3(d[0], d[1], d[2]), tmax, flags);
    t_any_hit_override = nullptr;
    RefRay ray { { o[0], o[1], o[2] }, tmin, { d[0], d[1], d[2] }, tmax };
    if (h.valid)
    {
        RefHit rh { h.t, h.u, h.v, h.instance, h.geometry, h.primitive };
        if (sbt_offset == 0)
            ref_rchit_invoke(payload, &rh, &ray);
        else
            ref_shadow_rchit_invoke(payload);
    }
    else if (miss_index == 0)
        ref_rmiss_invoke(payload, &ray);
    else
        ref_shadow_rmiss_invoke(payload);
}

// descriptor set 5 of the RAY_DEBUG_VIEW pipeline: one vertex buffer + one draw-argument block for all stages
static RefDebugVertexBlock g_debug_vertex_block = { nullptr };
static RefDebugDrawArgs    g_debug_draw_args    = { 0, 1, 0, 0 }; // renderer.cpp:236
extern "C" RefDebugVertexBlock* ref_drv_debug_vertex_block() { return &g_debug_vertex_block; }
extern "C" RefDebugDrawArgs*    ref_drv_debug_draw_args() { return &g_debug_draw_args; }

#if defined(RAY_DEBUG_VIEW)
// PathIntegrator::gather_debug_rays (path_integrator.cpp:88-104): launch_rays(ray-debug pipeline, num_debug_rays, 1, 1):
// same framing as or_gather_debug_rays.  `out` must hold max_vertices vertices of 8 floats; the shaders do not check the
// capacity (nor does the reference: its buffer has room for 2048), so the launch runs into a scratch buffer sized for the
// worst case (max_ray_bounces segments per path) and min(count, max_vertices) vertices are copied out.
OR_API uint32_t ref_gather_debug_rays(const Scene* s, const PushConstants* pcp, uint32_t num_debug_rays, float* out, uint32_t max_vertices)
{
    std::vector<void*> vb, ib, sb;
    for (auto m : s->meshes) vb.push_back((void*)m->verts.data()), ib.push_back((void*)m->indices.data());
    for (auto& t : s->submesh_info) sb.push_back((void*)t.data());
    RefBindings b;
    b.materials = s->materials.data(), b.instances = s->instances.data(), b.lights = s->lights.data();
    b.vertices = vb.data(), b.indices = ib.data(), b.submesh_info = sb.data();
    b.previous_color = nullptr, b.current_color = nullptr, b.width = (int)pcp->launch_id_size[2], b.height = (int)pcp->launch_id_size[3], b.push_constants = pcp;
    g_ref_scene = s;
    ref_rgen_bind(&b), ref_rchit_bind(&b), ref_rahit_bind(&b);
    std::vector<float> scratch((size_t)num_debug_rays * (pcp->max_ray_bounces + 1) * 2 * 8 + 16);
    g_debug_vertex_block.vertices = scratch.data();
    g_debug_draw_args             = { 0, 1, 0, 0 };
    for (uint32_t i = 0; i < num_debug_rays; i++) ref_rgen_invoke(i, 0);
    const uint32_t n = g_debug_draw_args.count;
    if (out) std::memcpy(out, scratch.data(), (size_t)std::min(n, max_vertices) * 32);
    g_debug_vertex_block.vertices = nullptr;
    return n;
}
#endif

// same signature and semantics as or_render_frame (minus raw_L): one launch of the ray-tracing pipeline
OR_API void ref_render_frame(const Scene* s, const PushConstants* pcp, uint32_t lw, uint32_t lh, const float* prev, float* cur, uint64_t* counters)
{
    const uint32_t W = pcp->launch_id_size[2], H = pcp->launch_id_size[3];
    if (lw == 0) lw = W;
    if (lh == 0) lh = H;
    std::vector<void*> vb, ib, sb;
    for (auto m : s->meshes) vb.push_back((void*)m->verts.data()), ib.push_back((void*)m->indices.data());
    for (auto& t : s->submesh_info) sb.push_back((void*)t.data());
    RefBindings b;
    b.materials = s->materials.data(), b.instances = s->instances.data(), b.lights = s->lights.data();
    b.vertices = vb.data(), b.indices = ib.data(), b.submesh_info = sb.data();
    b.previous_color = prev, b.current_color = cur, b.width = (int)W, b.height = (int)H, b.push_constants = pcp;
    g_ref_scene = s;
    ref_rgen_bind(&b), ref_rchit_bind(&b), ref_rahit_bind(&b);
    uint64_t ext = 0, sh = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : ext, sh)
    for (int64_t y = 0; y < (int64_t)lh; y++)
    {
        t_ref_extension_rays = t_ref_shadow_rays = 0;
        for (uint32_t x = 0; x < lw; x++) ref_rgen_invoke(x, (uint32_t)y);
        ext += t_ref_extension_rays, sh += t_ref_shadow_rays;
    }
    if (counters) counters[0] += ext, counters[1] += sh;
}

extern "C" void ref_tone_map_fragment(const float* src, int w, int h, int x, int y, float exposure, uint32_t op, float* out4);
extern "C" void ref_sky_bind(const float* ubo40, const float* sun_dir3);
extern "C" void ref_sky_fragment(const float* pos3, float* out4);

// tone_map.frag over the whole attachment, same framing as or_tonemap: Y flip of Renderer::tone_map's negative
// viewport, `sample_scale` folded into exposure is NOT done (the reference has no such factor; pass 1), UNORM8
// store = round-to-nearest of clamp(c,0,1)*255.
OR_API void ref_tonemap(const float* accum, uint32_t W, uint32_t H, float exposure, int op, uint8_t* out)
{
    for (uint32_t r = 0; r < H; r++)
        for (uint32_t x = 0; x < W; x++)
        {
            float c[4];
            ref_tone_map_fragment(accum, (int)W, (int)H, (int)x, (int)(H - 1 - r), exposure, (uint32_t)op, c);
            uint8_t* dst = out + ((size_t)r * W + x) * 4;
            for (int ch = 0; ch < 4; ch++)
            {
                float q = clampf(c[ch], 0.0f, 1.0f) * 255.0f + 0.5f;
                dst[ch] = (uint8_t)(c[ch] != c[ch] ? 0 : (int)q);
            }
        }
}
// procedural_sky.frag over the six cube faces; texel -> cube position exactly as or_sky_bake frames it
OR_API void ref_sky_bake(const float* cf40, const float* sun_dir, uint32_t size, float* out)
{
    ref_sky_bind(cf40, sun_dir);
    for (int face = 0; face < 6; face++)
        for (uint32_t j = 0; j < size; j++)
            for (uint32_t i = 0; i < size; i++)
            {
                float sc = 2.0f * (((float)i + 0.5f) / (float)size) - 1.0f;
                float tc = 2.0f * (((float)j + 0.5f) / (float)size) - 1.0f;
                float p[3];
                switch (face)
                {
                    case 0: p[0] = 1.0f, p[1] = -tc, p[2] = -sc; break;
                    case 1: p[0] = -1.0f, p[1] = -tc, p[2] = sc; break;
                    case 2: p[0] = sc, p[1] = 1.0f, p[2] = tc; break;
                    case 3: p[0] = sc, p[1] = -1.0f, p[2] = -tc; break;
                    case 4: p[0] = sc, p[1] = -tc, p[2] = 1.0f; break;
                    default: p[0] = -sc, p[1] = -tc, p[2] = -1.0f; break;
                }
                ref_sky_fragment(p, out + (((size_t)face * size + j) * size + i) * 4);
            }
}
