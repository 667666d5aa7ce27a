// hl_shade.h — the shade stage of the wavefront integrator: everything the reference's closest-hit shader
// does between two traceRayEXT calls, restated iteratively.
//   surface fetch          path_trace_rchit.glsl:136-280  (populate_surface_properties and helpers)
//   BRDF                   brdf.glsl:6-167
//   light sampling (NEE)   path_trace_rchit.glsl:284-451 (up to the shadow traceRayEXT)
//   direct / indirect      path_trace_rchit.glsl:455-536
//   closest-hit main       path_trace_rchit.glsl:542-580
//   miss                   path_trace_rmiss.glsl:60-65
// The reference is recursive: radiance returns through payloads already weighted by the path throughput, so
// the pixel sample is the sum of per-bounce terms; this file produces those terms
// plus the next extension ray and the shadow ray whose visibility gates the direct term.
// RNG consumption order is preserved exactly, including sample_uber's by-value generator.
#pragma once
#include "hl_rng.h"
#include "hl_scene.h"
#include "hl_tex.h"

namespace hl
{
#define HL_PI 3.14159265359f /* common.glsl:17 */
#define HL_EPSILON 0.0001f   /* common.glsl:18 */
#define HL_MIN_ROUGHNESS 0.1f /* common.glsl:20 */

struct Surface // SurfaceProperties, common.glsl:67-78 (only the members the path reads)
{
    f3    position;
    f3    normal; // shading normal (vertex normal or normal map)
    f3    albedo;
    f3    emissive;
    f3    F0;
    float roughness;
    float roughness_unclamped, metallic; // as fetched, before the MIN_ROUGHNESS floor / the F0 mix (debug output buffers only)
};

// ---- BRDF (brdf.glsl) ---------------------------------------------------------------------------
HL_HD f3 cosine_lobe_direction(f3 n, float r0, float r1) // sample_cosine_lobe, brdf.glsl:6-18
{
    const float rx = fmaxf(0.00001f, r0), ry = fmaxf(0.00001f, r1);
    const float phi = 2.0f * HL_PI * ry;
    const float ct = sqrtf(rx), st = sqrtf(1 - rx);
    const f3    t = mk3(st * cosf(phi), st * sinf(phi), ct);
    f3          bx, by;
    basis_around(n, bx, by);
    return normalize(cols_mul(bx, by, n, t));
}
HL_HD float ggx_D(float ndoth, float alpha) // brdf.glsl:58-64
{
    const float a2 = alpha * alpha;
    const float dn = (ndoth * ndoth) * (a2 - 1.0f) + 1.0f;
    return a2 / fmaxf(HL_EPSILON, (HL_PI * dn * dn));
}
HL_HD float ggx_G1(float roughness, float ndotv) // brdf.glsl:66-71
{
    const float k = ((roughness + 1) * (roughness + 1)) / 8.0f;
    return ndotv / fmaxf(HL_EPSILON, (ndotv * (1 - k) + k));
}
HL_HD f3 ggx_half_vector(f3 n, float alpha, float xi0, float xi1) // sample_ggx, brdf.glsl:83-97
{
    const float phi = 2.0f * HL_PI * xi0;
    const float ct  = sqrtf((1.0f - xi1) / (1.0f + (alpha * alpha - 1.0f) * xi1));
    const float st  = sqrtf(1.0f - ct * ct);
    const f3    d   = mk3(st * cosf(phi), st * sinf(phi), ct);
    f3          bx, by;
    basis_around(n, bx, by);
    return normalize(cols_mul(bx, by, n, d));
}
// evaluate_uber (brdf.glsl:110-122) and pdf_uber (:124-135) share their four clamped cosines
HL_HD void uber_eval(const Surface& p, f3 Wo, f3 Wh, f3 Wi, f3& brdf, float& pdf)
{
    const float NdotL = fmaxf(dot(p.normal, Wi), 0.0f);
    const float NdotV = fmaxf(dot(p.normal, Wo), 0.0f);
    const float NdotH = fmaxf(dot(p.normal, Wh), 0.0f);
    const float VdotH = fmaxf(dot(Wi, Wh), 0.0f);
    const f3    F     = p.F0 + (mk3(1.0f) - p.F0) * (powf(1.0f - VdotH, 5.0f)); // F_schlick :78-81
    const float alpha = p.roughness * p.roughness;
    const float D     = ggx_D(NdotH, alpha);
    const float G     = ggx_G1(p.roughness, NdotL) * ggx_G1(p.roughness, NdotV);
    const f3    spec  = (D * F * G) / fmaxf(HL_EPSILON, (4.0f * NdotL * NdotV)); // evaluate_ggx :99-103
    const f3    diff  = p.albedo / HL_PI;                                        // evaluate_lambert :30-33
    brdf              = (mk3(1.0f) - F) * diff + spec;
    const float pd    = NdotL / HL_PI;                                            // pdf_cosine_lobe
    const float ps    = D * NdotH / fmaxf(HL_EPSILON, (4.0f * VdotH));           // pdf_D_ggx :105-108
    pdf               = mixf(pd, ps, 0.5f);
}
// sample_uber, brdf.glsl:137-167.  `rng` is a COPY: the caller's generator is not advanced.
HL_HD void uber_sample(const Surface& p, f3 Wo, Rng rng, f3& Wi, f3& brdf, float& pdf)
{
    const float alpha = p.roughness * p.roughness;
    const float r0 = rand01(rng), r1 = rand01(rng), r2 = rand01(rng);
    f3          Wh;
    bool        specular = false;
    if (r0 < 0.5f)
    {
        Wh                = ggx_half_vector(p.normal, alpha, r1, r2);
        Wi                = reflect(-Wo, Wh);
        const float NdotL = fmaxf(dot(p.normal, Wi), 0.0f);
        const float NdotV = fmaxf(dot(p.normal, Wo), 0.0f);
        specular          = NdotL > 0.0f && NdotV > 0.0f;
    }
    if (!specular)
    {
        Wi = cosine_lobe_direction(p.normal, r1, r2);
        Wh = normalize(Wo + Wi);
    }
    uber_eval(p, Wo, Wh, Wi, brdf, pdf);
}

// ---- surface fetch --------------------------------------------------------------------------------
HL_HD f3 bary3(const float* a, const float* b, const float* c, float b0, float b1, float b2)
{
    return mk3(a[0] * b0 + b[0] * b1 + c[0] * b2, a[1] * b0 + b[1] * b1 + c[1] * b2, a[2] * b0 + b[2] * b1 + c[2] * b2);
}
HL_HD void load_surface(const SceneView& s, const Hit& h, Surface& p)
{
    const hl_instance& I    = s.instances[h.instance];
    const MeshView&    m    = s.meshes[I.mesh_index];
    const uint32_t*    info = s.submesh_info + 2 * (size_t)(s.submesh_offset[h.instance] + h.geometry); // rchit:141-152
    const size_t       pid  = (size_t)h.primitive + info[0];
    const hl_material& mat  = s.materials[info[1]];
    const hl_vertex&   v0   = m.vertices[m.indices[3 * pid + 0]]; // fetch_triangle rchit:156-172
    const hl_vertex&   v1   = m.vertices[m.indices[3 * pid + 1]];
    const hl_vertex&   v2   = m.vertices[m.indices[3 * pid + 2]];
    const float        b0 = 1.0f - h.u - h.v, b1 = h.u, b2 = h.v;
    // interpolated_vertex common.glsl:130-141, transform_vertex rchit:176-186
    const f3 pos = bary3(v0.position, v1.position, v2.position, b0, b1, b2);
    p.position   = mat4_mul_point_xyz(I.model_matrix, pos, 1.0f);
    const f3 vn  = mat3_mul(I.normal_matrix, normalize(bary3(v0.normal, v1.normal, v2.normal, b0, b1, b2)));
    const bool need_uv = mat.texture_indices0[0] != -1 || mat.texture_indices0[1] != -1 || mat.texture_indices0[2] != -1 || mat.texture_indices0[3] != -1 || mat.texture_indices1[0] != -1;
    float      tu = 0.0f, tv = 0.0f;
    if (need_uv)
    {
        tu = v0.tex_coord[0] * b0 + v1.tex_coord[0] * b1 + v2.tex_coord[0] * b2;
        tv = v0.tex_coord[1] * b0 + v1.tex_coord[1] * b1 + v2.tex_coord[1] * b2;
    }
    // fetch_albedo .. fetch_emissive rchit:206-252
    if (mat.texture_indices0[0] == -1)
        p.albedo = mk3(mat.albedo);
    else
    {
        const f4 a = sample_texture_lod0(s, mat.texture_indices0[0], tu, tv);
        p.albedo   = mk3(a.x, a.y, a.z);
    }
    if (mat.texture_indices0[1] == -1)
        p.normal = vn;
    else
    {
        // get_normal_from_map rchit:190-202
        const f3 vt = mat3_mul(I.normal_matrix, normalize(bary3(v0.tangent, v1.tangent, v2.tangent, b0, b1, b2)));
        const f3 vb = mat3_mul(I.normal_matrix, normalize(bary3(v0.bitangent, v1.bitangent, v2.bitangent, b0, b1, b2)));
        const f4 tx = sample_texture_lod0(s, mat.texture_indices0[1], tu, tv);
        const f3 n  = normalize(mk3(tx.x, tx.y, tx.z) * 2.0f - mk3(1.0f));
        p.normal    = normalize(cols_mul(normalize(vt), normalize(vb), normalize(vn), n));
    }
    float metallic;
    if (mat.texture_indices0[2] == -1)
        p.roughness = mat.roughness_metallic[0];
    else
    {
        const f4  c  = sample_texture_lod0(s, mat.texture_indices0[2], tu, tv);
        const int ch = mat.texture_indices1[2] & 3;
        p.roughness  = ch == 0 ? c.x : (ch == 1 ? c.y : (ch == 2 ? c.z : c.w));
    }
    if (mat.texture_indices0[3] == -1)
        metallic = mat.roughness_metallic[1];
    else
    {
        const f4  c  = sample_texture_lod0(s, mat.texture_indices0[3], tu, tv);
        const int ch = mat.texture_indices1[3] & 3;
        metallic     = ch == 0 ? c.x : (ch == 1 ? c.y : (ch == 2 ? c.z : c.w));
    }
    if (mat.texture_indices1[0] == -1)
        p.emissive = mk3(mat.emissive);
    else
    {
        const f4 e = sample_texture_lod0(s, mat.texture_indices1[0], tu, tv);
        p.emissive = mk3(e.x, e.y, e.z);
    }
    p.roughness_unclamped = p.roughness, p.metallic = metallic;
    p.roughness = fmaxf(p.roughness, HL_MIN_ROUGHNESS);
    p.F0        = mix3(mk3(0.03f), p.albedo, metallic);
}

// Debug output buffers (debug_visualization.frag:144-161): one material channel of the surface a primary ray hit —
// 0 albedo, 1 shading normal * 0.5 + 0.5, 2 roughness, 3 metallic (both as fetched: no floor, no F0 mix), 4 emissive;
// alpha 1; a miss is the pass's clear colour (0,0,0,1).
HL_HD f4 output_buffer_value(const SceneView& s, const Hit& h, int which)
{
    f4 c;
    c.x = c.y = c.z = 0.0f, c.w = 1.0f;
    if (h.instance == HL_MISS) return c;
    Surface p;
    load_surface(s, h, p);
    if (which == 0)
        c.x = p.albedo.x, c.y = p.albedo.y, c.z = p.albedo.z;
    else if (which == 1)
        c.x = p.normal.x * 0.5f + 0.5f, c.y = p.normal.y * 0.5f + 0.5f, c.z = p.normal.z * 0.5f + 0.5f;
    else if (which == 2)
        c.x = c.y = c.z = p.roughness_unclamped;
    else if (which == 3)
        c.x = c.y = c.z = p.metallic;
    else
        c.x = p.emissive.x, c.y = p.emissive.y, c.z = p.emissive.z;
    return c;
}

// ---- next-event estimation ------------------------------------------------------------------------
struct LightSample
{
    f3    Wi, Li;
    float pdf, tmax;
    bool  traced; // false: the reference returns before the shadow traceRayEXT (area-light early outs)
};
HL_HD f3 jitter_on_disk(f3 light_dir, float radius, float r0, float r1) // rchit:305-316 (and :334-341, :355-362)
{
    const f3    tg = normalize(cross(light_dir, mk3(0.0f, 1.0f, 0.0f)));
    const f3    bt = normalize(cross(tg, light_dir));
    const float pr = radius * sqrtf(r0);
    const float pa = r1 * 2.0f * HL_PI;
    const float dx = pr * cosf(pa), dy = pr * sinf(pa);
    return normalize(light_dir + dx * tg + dy * bt);
}
HL_HD void sample_light(const SceneView& s, const Surface& p, const hl_light& L, Rng& rng, LightSample& o)
{
    o.Wi = mk3(0.0f), o.Li = mk3(0.0f), o.pdf = 0.0f, o.tmax = 10000.0f, o.traced = true;
    const uint32_t type = (uint32_t)L.light_data0[0];
    if (type == HL_LIGHT_DIRECTIONAL)
    {
        const float r0 = rand01(rng), r1 = rand01(rng);
        o.Wi = jitter_on_disk(-mk3(L.light_data1), L.light_data2[3], r0, r1);
        o.Li = mk3(L.light_data0[1], L.light_data0[2], L.light_data0[3]) * L.light_data1[3];
    }
    else if (type == HL_LIGHT_SPOT || type == HL_LIGHT_POINT)
    {
        const float r0 = rand01(rng), r1 = rand01(rng);
        const f3    to_light = mk3(L.light_data2) - p.position;
        const f3    dir      = normalize(to_light);
        const float dist     = length(to_light);
        const float radius   = L.light_data2[3] / dist;
        float       att      = 1.0f;
        if (type == HL_LIGHT_SPOT) att = smoothstepf(L.light_data3[1], L.light_data3[0], dot(dir, -mk3(L.light_data1)));
        o.Wi = jitter_on_disk(dir, radius, r0, r1);
        if (type == HL_LIGHT_SPOT)
            o.Li = mk3(L.light_data0[1], L.light_data0[2], L.light_data0[3]) * L.light_data1[3] * att / (dist * dist);
        else
            o.Li = mk3(L.light_data0[1], L.light_data0[2], L.light_data0[3]) * L.light_data1[3] / (dist * dist);
        o.tmax = dist;
    }
    else if (type == HL_LIGHT_ENVIRONMENT_MAP)
    {
        const float r0 = rand01(rng), r1 = rand01(rng);
        o.Wi  = cosine_lobe_direction(p.normal, r0, r1);
        o.Li  = sample_environment(s.env, o.Wi);
        o.pdf = dot(p.normal, o.Wi) / HL_PI;
    }
    else if (type == HL_LIGHT_AREA)
    {
        const uint32_t inst_id  = (uint32_t)L.light_data0[1];
        const uint32_t n_tris   = (uint32_t)L.light_data1[2]; // reads .z although the host wrote .x
        const uint32_t prim     = rand_below(rng, n_tris);
        const uint32_t mat_idx  = (uint32_t)L.light_data0[2];
        const size_t   pid      = (size_t)prim + (uint32_t)L.light_data0[3];
        const hl_instance& I    = s.instances[inst_id];
        const MeshView&    m    = s.meshes[I.mesh_index];
        const hl_vertex&   v0   = m.vertices[m.indices[3 * pid + 0]];
        const hl_vertex&   v1   = m.vertices[m.indices[3 * pid + 1]];
        const hl_vertex&   v2   = m.vertices[m.indices[3 * pid + 2]];
        const float u0 = rand01(rng), u1 = rand01(rng);
        const float su = sqrtf(u0);
        const float bx = 1 - su, by = u1 * su; // uniform_sample_triangle brdf.glsl:40-44
        // model_matrix * position with the vertex's own w (= submesh index)
        const f3 p0 = mat4_mul_point_xyz(I.model_matrix, mk3(v0.position), v0.position[3]);
        const f3 p1 = mat4_mul_point_xyz(I.model_matrix, mk3(v1.position), v1.position[3]);
        const f3 p2 = mat4_mul_point_xyz(I.model_matrix, mk3(v2.position), v2.position[3]);
        const float w0 = 1.0f - bx - by; // barycentric_interpolate brdf.glsl:46-51
        const f3 lpos = p0 * w0 + p1 * bx + p2 * by;
        const f3 lnrm = normalize(mat3_mul(I.normal_matrix, mk3(v0.normal) * w0 + mk3(v1.normal) * bx + mk3(v2.normal) * by));
        f3       ldir = p.position - lpos;
        const float d2   = dot(ldir, ldir);
        const float area = 0.5f * length(cross(p1 - p0, p2 - p0)); // triangle_area brdf.glsl:35-38
        if (area == 0.0f || d2 == 0.0f)
        {
            o.traced = false;
            return;
        }
        const float dist = sqrtf(d2);
        ldir             = ldir / dist;
        o.tmax           = fmaxf(0.0f, dist - HL_EPSILON);
        const float ct   = dot(lnrm, ldir);
        if (ct == 0.0f)
        {
            o.traced = false;
            return;
        }
        o.Li  = mk3(s.materials[mat_idx].emissive);
        o.Wi  = -ldir;
        o.pdf = d2 / fmaxf(HL_EPSILON, ct * area); // pdf_triangle brdf.glsl:53-56
    }
}

// ---- one closest-hit invocation, iteratively --------------------------------------------------------
struct ShadeParams
{
    uint32_t num_lights, max_ray_bounces;
    float    shadow_ray_bias;
    uint32_t ray_debug_view = 0; // 1 = the RAY_DEBUG_VIEW variant of the closest-hit shader: no Russian roulette (hl_debug.h)
};
struct ShadeResult
{
    f3    emitted;       // added to L immediately (depth 0 emissive)
    bool  has_shadow;    // a shadow ray must be traced; `direct` is added when it is unoccluded
    f3    shadow_o, shadow_d;
    float shadow_tmax;
    f3    direct;
    bool  continues;     // the path survives: next extension ray + new throughput
    f3    next_o, next_d;
    f3    T;
};
HL_HD void shade_hit(const SceneView& s, const ShadeParams& prm, uint32_t depth, f3 ray_dir, const Hit& h, f3 T, Rng& rng, ShadeResult& r)
{
    Surface p;
    load_surface(s, h, p);
    r.emitted    = (depth == 0 && !is_black(p.emissive)) ? p.emissive : mk3(0.0f); // rchit:571-572
    r.has_shadow = false, r.continues = false;
    const f3 Wo  = -ray_dir;

    // direct_lighting rchit:455-483
    {
        const uint32_t li = rand_below(rng, prm.num_lights);
        hl_light       L;
        if (li < s.n_lights)
            L = s.lights[li];
        else
        {
            for (int k = 0; k < 4; k++) L.light_data0[k] = L.light_data1[k] = L.light_data2[k] = L.light_data3[k] = 0.0f;
        }
        LightSample ls;
        sample_light(s, p, L, rng, ls);
        if (ls.traced && !is_black(ls.Li))
        {
            const f3 Wh = normalize(Wo + ls.Wi);
            f3       brdf;
            float    unused_pdf;
            uber_eval(p, Wo, Wh, ls.Wi, brdf, unused_pdf);
            const float ct = clampf(dot(p.normal, ls.Wi), 0.0f, 1.0f);
            f3          Ld = T * brdf * ct * ls.Li;
            if (ls.pdf != 0.0f) Ld = Ld / ls.pdf;
            Ld = Ld * (float)prm.num_lights;
            if (!is_black(Ld)) // a black term cannot change L: its shadow ray is skipped (SURVEY C-2)
            {
                r.has_shadow  = true;
                r.direct      = Ld;
                r.shadow_o    = p.position + p.normal * prm.shadow_ray_bias;
                r.shadow_d    = ls.Wi;
                r.shadow_tmax = ls.tmax;
            }
        }
    }
    // indirect_lighting rchit:487-536
    if (depth + 1 < prm.max_ray_bounces)
    {
        f3    Wi, brdf;
        float pdf;
        uber_sample(p, Wo, rng /* by value */, Wi, brdf, pdf);
        const float ct   = clampf(dot(p.normal, Wi), 0.0f, 1.0f);
        f3          Tn   = T * (brdf * ct) / pdf;
        const float prob = fmaxf(Tn.x, fmaxf(Tn.y, Tn.z));
        if (prm.ray_debug_view || !(rand01(rng) > prob)) // Russian roulette, rchit:500-509 (#if !defined(RAY_DEBUG_VIEW))
        {
            if (!prm.ray_debug_view) Tn = Tn * (1.0f / prob);
            r.continues = true;
            r.next_o    = p.position;
            r.next_d    = Wi;
            r.T         = Tn;
        }
    }
}
// miss shader, path_trace_rmiss.glsl:60-65
HL_HD f3 shade_miss(const SceneView& s, uint32_t depth, f3 ray_dir, f3 T)
{
    const f3 e = sample_environment(s.env, ray_dir);
    return depth == 0 ? e : T * e;
}
} // namespace hl
