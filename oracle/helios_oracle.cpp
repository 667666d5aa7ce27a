// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// CPU restatement ("port") of the path-tracing hot path of diharaw/helios, following the
// reference's GLSL recursively and bug-for-bug.  PINNED AGAINST THE REFERENCE ITSELF: upstream has
// no tests or golden vectors and the engine cannot run here (Windows / Vulkan-RT only, SURVEY.md
// §0), but its shader files and the host half of its sky model do compile as C++ — oracle/ref_glsl/
// builds them from /root/reference into oracle/_ref/libhelios_glsl_ref.so, and every frame, ray
// count, RNG / BRDF / sky / tone-map value of this restatement is bit-identical to that library on
// all scene types (tests/test_ref_glsl.py; outputs committed as tests/golden/ref_glsl_golden.npz).
// What the reference leaves to the Vulkan driver (traversal, texture filtering) is shared by both.
// Further pins: the public xoroshiro64* sequence, closed forms, brute force == BVH (tests/test_oracle_kat.py).
//
// Reference files restated (paths relative to the reference checkout, src/engine/shader/ unless noted):
//   random.glsl:11-50, sampling.glsl:6-36, brdf.glsl:6-167, common.glsl:130-141,
//   path_trace_rgen.glsl:132-250, path_trace_rchit.glsl:136-580, path_trace_rahit.glsl:121-188,
//   path_trace_rmiss.glsl:38-67, path_trace_shadow.{rchit,rmiss}:16-19, tone_map.frag:20-51,
//   procedural_sky.frag:48-75, gfx/hosek_wilkie_sky_model.cpp:41-75,658-686.
// Driver-defined behaviour the reference leaves to Vulkan (ray/triangle test, tie break, texture
// filtering, cube-face selection) follows SURVEY.md A.9 and is documented at each function.
#include "glsl_math.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <vector>

using namespace orc;

#define OR_API extern "C" __attribute__((visibility("default")))

static const float M_PI_F  = 3.14159265359f; // common.glsl:17
static const float EPSILON = 0.0001f;        // common.glsl:18
static const float MIN_ROUGHNESS = 0.1f;     // common.glsl:20

// ------------------------------------------------------------------------------------------------
// shader-ABI structs (byte-identical to include/helios_b200.h; restated so the oracle is standalone)
// ------------------------------------------------------------------------------------------------
struct Vertex
{
    float position[4], tex_coord[4], normal[4], tangent[4], bitangent[4];
};
struct Material
{
    int32_t texture_indices0[4], texture_indices1[4];
    float   albedo[4], emissive[4], roughness_metallic[4];
};
struct Light
{
    float d0[4], d1[4], d2[4], d3[4];
};
struct Instance
{
    float    model_matrix[16], normal_matrix[16];
    uint32_t mesh_idx;
    float    pad[3];
};
struct PushConstants
{
    float    view_proj_inverse[16];
    float    camera_pos[4], up_direction[4], right_direction[4], focal_plane[4];
    int32_t  ray_debug_pixel_coord[4];
    uint32_t launch_id_size[4];
    float    accumulation;
    uint32_t num_lights, num_frames, debug_vis, max_ray_bounces;
    float    shadow_ray_bias, focal_length, aperture_radius;
};
struct SubMesh
{
    uint32_t base_index, index_count, vertex_count, opaque;
};
static_assert(sizeof(Vertex) == 80 && sizeof(Material) == 80 && sizeof(Light) == 64 && sizeof(Instance) == 144 && sizeof(PushConstants) == 192, "ABI");

// ------------------------------------------------------------------------------------------------
// random.glsl / sampling.glsl
// ------------------------------------------------------------------------------------------------
struct RNG
{
    uint32_t sx, sy;
};
static inline uint32_t rng_rotl(uint32_t x, uint32_t k) { return (x << k) | (x >> (32 - k)); } // random.glsl:11-14
static inline uint32_t rng_next(RNG& rng)                                                       // random.glsl:17-26
{
    uint32_t result = rng.sx * 0x9e3779bbu;
    rng.sy ^= rng.sx;
    rng.sx = rng_rotl(rng.sx, 26) ^ rng.sy ^ (rng.sy << 9);
    rng.sy = rng_rotl(rng.sy, 13);
    return result;
}
static inline uint32_t rng_hash(uint32_t seed) // random.glsl:30-38
{
    seed = (seed ^ 61u) ^ (seed >> 16);
    seed *= 9u;
    seed = seed ^ (seed >> 4);
    seed *= 0x27d4eb2du;
    seed = seed ^ (seed >> 15);
    return seed;
}
static inline RNG rng_init(uint32_t idx, uint32_t idy, uint32_t frame) // random.glsl:40-50
{
    RNG rng;
    rng.sx = rng_hash((idx << 16) | idy);
    rng.sy = rng_hash(frame);
    rng_next(rng);
    return rng;
}
static inline float next_float(RNG& rng) // sampling.glsl:6-10
{
    uint32_t u = 0x3f800000u | (rng_next(rng) >> 9);
    return bits_float(u) - 1.0f;
}
static inline uint32_t next_uint(RNG& rng, uint32_t nmax) // sampling.glsl:12-16
{
    float f = next_float(rng);
    return (uint32_t)std::floor(f * (float)nmax);
}
static inline vec2 next_vec2(RNG& rng) // sampling.glsl:18-21 (left-to-right)
{
    vec2 r;
    r.x = next_float(rng);
    r.y = next_float(rng);
    return r;
}
static inline vec3 next_vec3(RNG& rng) // sampling.glsl:23-26
{
    vec3 r;
    r.x = next_float(rng);
    r.y = next_float(rng);
    r.z = next_float(rng);
    return r;
}
// sampling.glsl:28-36: returns mat3(x, y, z) as columns
static inline void make_rotation_matrix(vec3 z, vec3& x, vec3& y)
{
    const vec3 ref = std::fabs(dot(z, vec3(0, 1, 0))) > 0.99f ? vec3(0, 0, 1) : vec3(0, 1, 0);
    x              = normalize(cross(ref, z));
    y              = cross(z, x);
}

// ------------------------------------------------------------------------------------------------
// textures (driver-defined in the reference; our definition: level 0, REPEAT, bilinear with fp32
// weights, texel centres at +0.5, 8-bit decode through a LUT computed in double)
// ------------------------------------------------------------------------------------------------
struct Texture
{
    int                  format; // 0 unorm8, 1 srgb8, 2 snorm8, 3 rgba32f
    uint32_t             w, h;
    std::vector<uint8_t> data;
};
static float g_srgb_lut[256], g_unorm_lut[256], g_snorm_lut[256];
static bool  g_luts_ready = false;
static void  init_luts()
{
    if (g_luts_ready) return;
    for (int i = 0; i < 256; i++)
    {
        double c       = i / 255.0;
        g_unorm_lut[i] = (float)c;
        g_srgb_lut[i]  = (float)(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
        int s          = (int8_t)(uint8_t)i;
        g_snorm_lut[i] = (float)std::max(-1.0, s / 127.0);
    }
    g_luts_ready = true;
}
static inline vec4 texel_fetch(const Texture& t, int x, int y)
{
    size_t i = (size_t)y * t.w + x;
    if (t.format == 3)
    {
        const float* p = (const float*)t.data.data() + i * 4;
        return vec4(p[0], p[1], p[2], p[3]);
    }
    const uint8_t* p   = t.data.data() + i * 4;
    const float*   lut = t.format == 1 ? g_srgb_lut : (t.format == 2 ? g_snorm_lut : g_unorm_lut);
    // alpha of an sRGB texture is linear (Vulkan spec); snorm applies to all four channels
    float a = t.format == 1 ? g_unorm_lut[p[3]] : lut[p[3]];
    return vec4(lut[p[0]], lut[p[1]], lut[p[2]], a);
}
static inline vec4 bilerp(vec4 t00, vec4 t10, vec4 t01, vec4 t11, float fx, float fy)
{
    vec4 a = t00 * (1.0f - fx) + t10 * fx;
    vec4 b = t01 * (1.0f - fx) + t11 * fx;
    return a * (1.0f - fy) + b * fy;
}
static vec4 texture_lod0(const Texture& t, float u, float v)
{
    if (!(std::fabs(u) < 1e30f) || !(std::fabs(v) < 1e30f)) return vec4(0, 0, 0, 0); // NaN / inf guard
    u        = u - std::floor(u);
    v        = v - std::floor(v);
    float x  = u * (float)t.w - 0.5f;
    float y  = v * (float)t.h - 0.5f;
    float x0 = std::floor(x), y0 = std::floor(y);
    float fx = x - x0, fy = y - y0;
    int   ix0 = (int)x0, iy0 = (int)y0;
    int   W = (int)t.w, H = (int)t.h;
    ix0     = ((ix0 % W) + W) % W;
    iy0     = ((iy0 % H) + H) % H;
    int ix1 = (ix0 + 1) % W, iy1 = (iy0 + 1) % H;
    return bilerp(texel_fetch(t, ix0, iy0), texel_fetch(t, ix1, iy0), texel_fetch(t, ix0, iy1), texel_fetch(t, ix1, iy1), fx, fy);
}

// cube map: Vulkan face selection table (major axis priority z, y, x), bilinear inside the face,
// clamp to edge (no seamless filtering; driver-defined upstream)
struct EnvMap
{
    uint32_t           size = 0;
    std::vector<float> data; // 6 * size * size * 4
};
// Vulkan "Cube Map Edge Handling": cube maps are always seamless — a bilinear footprint that reaches over a face edge
// takes the texel of the adjoining face, and at a corner (where the fourth texel does not exist) the three existing
// texels are averaged.  (ix, iy) is at most one texel outside face `face`.
static vec4 env_texel(const EnvMap& e, int face, int ix, int iy)
{
    const int N   = (int)e.size;
    auto      raw = [&](int f, int x, int y) {
        const float* p = e.data.data() + (((size_t)f * N + (size_t)y) * N + (size_t)x) * 4;
        return vec4(p[0], p[1], p[2], p[3]);
    };
    const bool out_x = ix < 0 || ix >= N, out_y = iy < 0 || iy >= N;
    if (!out_x && !out_y) return raw(face, ix, iy);
    if (out_x && out_y)
    {
        const int cx = ix < 0 ? 0 : N - 1, cy = iy < 0 ? 0 : N - 1;
        return (env_texel(e, face, cx, cy) + env_texel(e, face, ix, cy) + env_texel(e, face, cx, iy)) * (1.0f / 3.0f);
    }
    // centre of the texel on the extended plane of the face, in the cube's coordinates
    const float rn = 1.0f / (float)N;
    const float u = 2.0f * ((float)ix + 0.5f) * rn - 1.0f, v = 2.0f * ((float)iy + 0.5f) * rn - 1.0f;
    float       q[3];
    if (face == 0) q[0] = 1.0f, q[1] = -v, q[2] = -u;
    else if (face == 1) q[0] = -1.0f, q[1] = -v, q[2] = u;
    else if (face == 2) q[0] = u, q[1] = 1.0f, q[2] = v;
    else if (face == 3) q[0] = u, q[1] = -1.0f, q[2] = -v;
    else if (face == 4) q[0] = u, q[1] = -v, q[2] = 1.0f;
    else q[0] = -u, q[1] = -v, q[2] = -1.0f;
    // fold over the shared edge: the overshooting axis becomes the new face's normal, the old normal axis the first row
    const int normal_axis = face / 2;
    int       over_axis   = -1;
    for (int k = 0; k < 3; k++)
        if (k != normal_axis && std::fabs(q[k]) > 1.0f) over_axis = k;
    q[normal_axis] = (q[normal_axis] < 0.0f ? -1.0f : 1.0f) * (1.0f - rn);
    q[over_axis]   = q[over_axis] < 0.0f ? -1.0f : 1.0f;
    const int nf   = 2 * over_axis + (q[over_axis] < 0.0f ? 1 : 0);
    float     s, t;
    if (nf == 0) s = -q[2], t = -q[1];
    else if (nf == 1) s = q[2], t = -q[1];
    else if (nf == 2) s = q[0], t = q[2];
    else if (nf == 3) s = q[0], t = -q[2];
    else if (nf == 4) s = q[0], t = -q[1];
    else s = -q[0], t = -q[1];
    int jx = (int)std::floor(0.5f * (s + 1.0f) * (float)N), jy = (int)std::floor(0.5f * (t + 1.0f) * (float)N);
    jx = std::min(std::max(jx, 0), N - 1), jy = std::min(std::max(jy, 0), N - 1);
    return raw(nf, jx, jy);
}
static vec3 env_sample(const EnvMap& e, vec3 r)
{
    if (e.size == 0) return vec3(0.0f); // reference default cube map is black (gfx/vk.cpp:3589-3612)
    float ax = std::fabs(r.x), ay = std::fabs(r.y), az = std::fabs(r.z);
    int   face;
    float sc, tc, ma;
    if (az >= ax && az >= ay)
    {
        face = r.z >= 0.0f ? 4 : 5;
        sc   = r.z >= 0.0f ? r.x : -r.x;
        tc   = -r.y;
        ma   = az;
    }
    else if (ay >= ax)
    {
        face = r.y >= 0.0f ? 2 : 3;
        sc   = r.x;
        tc   = r.y >= 0.0f ? r.z : -r.z;
        ma   = ay;
    }
    else
    {
        face = r.x >= 0.0f ? 0 : 1;
        sc   = r.x >= 0.0f ? -r.z : r.z;
        tc   = -r.y;
        ma   = ax;
    }
    if (!(ma > 0.0f) || !(ma < 1e30f)) return vec3(0.0f);
    float s = 0.5f * (sc / ma + 1.0f);
    float t = 0.5f * (tc / ma + 1.0f);
    if (!(s >= 0.0f && s <= 1.0f && t >= 0.0f && t <= 1.0f)) return vec3(0.0f); // NaN guard
    int   N  = (int)e.size;
    float x  = s * (float)N - 0.5f;
    float y  = t * (float)N - 0.5f;
    float x0 = std::floor(x), y0 = std::floor(y);
    float fx = x - x0, fy = y - y0;
    int   ix0 = (int)x0, iy0 = (int)y0; // -1 .. N-1
    vec4  c = bilerp(env_texel(e, face, ix0, iy0), env_texel(e, face, ix0 + 1, iy0), env_texel(e, face, ix0, iy0 + 1), env_texel(e, face, ix0 + 1, iy0 + 1), fx, fy);
    return c.xyz();
}

// ------------------------------------------------------------------------------------------------
// scene + acceleration structure (the reference leaves this to the driver; oracle = median-split
// binary BVH per mesh in object space + linear scan over instances; brute force selectable)
// ------------------------------------------------------------------------------------------------
struct BNode
{
    float    lo[3], hi[3];
    uint32_t left, right; // internal: children; leaf: right == 0xFFFFFFFF, left = first
    uint32_t count;
};
struct Mesh
{
    std::vector<Vertex>   verts;
    std::vector<uint32_t> indices;
    std::vector<SubMesh>  subs;
    // flattened triangle list: (geometry, primitive) per entry, in BVH order
    std::vector<uint32_t> tri_geom, tri_prim;
    std::vector<BNode>    nodes;
    float                 lo[3], hi[3];
};
struct Scene
{
    std::vector<Mesh*>                 meshes;
    std::vector<Texture>               textures;
    EnvMap                             env;
    std::vector<Material>              materials;
    std::vector<Instance>              instances;
    std::vector<std::vector<uint32_t>> submesh_info; // per instance: (prim offset, material) pairs
    std::vector<Light>                 lights;
    // per instance world->object (3x4 row-major) and world AABB
    std::vector<float> inv;  // 12 per instance
    std::vector<float> wbox; // 6 per instance
    bool               brute_force = false;
    ~Scene()
    {
        for (auto m : meshes) delete m;
    }
};

struct Hit
{
    float    t, u, v;
    uint32_t instance, geometry, primitive;
    bool     valid;
};

static const uint32_t RAY_FLAG_OPAQUE    = 1u; // gl_RayFlagsOpaqueEXT
static const uint32_t RAY_FLAG_TERMINATE = 4u; // gl_RayFlagsTerminateOnFirstHitEXT

// Moeller-Trumbore, fixed operation order, object space, direction not renormalised.
// Returns true when the ray's supporting line crosses the triangle (u>=0, v>=0, u+v<=1, det!=0).
static inline bool tri_test(vec3 o, vec3 d, vec3 p0, vec3 p1, vec3 p2, float& t, float& u, float& v)
{
    vec3  e1   = p1 - p0;
    vec3  e2   = p2 - p0;
    vec3  pvec = cross(d, e2);
    float det  = dot(e1, pvec);
    if (det == 0.0f || det != det) return false;
    float inv  = 1.0f / det;
    vec3  tvec = o - p0;
    u          = dot(tvec, pvec) * inv;
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    vec3 qvec = cross(tvec, e1);
    v         = dot(d, qvec) * inv;
    if (!(v >= 0.0f && u + v <= 1.0f)) return false;
    t = dot(e2, qvec) * inv;
    return true;
}

static void mesh_build_bvh(Mesh& m)
{
    size_t ntri = 0;
    for (auto& s : m.subs) ntri += s.index_count / 3;
    m.tri_geom.resize(ntri);
    m.tri_prim.resize(ntri);
    std::vector<float> cen(ntri * 3), blo(ntri * 3), bhi(ntri * 3);
    size_t             k = 0;
    for (size_t g = 0; g < m.subs.size(); g++)
        for (uint32_t p = 0; p < m.subs[g].index_count / 3; p++, k++)
        {
            m.tri_geom[k] = (uint32_t)g;
            m.tri_prim[k] = p;
            for (int a = 0; a < 3; a++)
            {
                float v0 = m.verts[m.indices[m.subs[g].base_index + 3 * p + 0]].position[a];
                float v1 = m.verts[m.indices[m.subs[g].base_index + 3 * p + 1]].position[a];
                float v2 = m.verts[m.indices[m.subs[g].base_index + 3 * p + 2]].position[a];
                blo[k * 3 + a] = std::min(v0, std::min(v1, v2));
                bhi[k * 3 + a] = std::max(v0, std::max(v1, v2));
                cen[k * 3 + a] = 0.5f * (blo[k * 3 + a] + bhi[k * 3 + a]);
            }
        }
    std::vector<uint32_t> order(ntri);
    for (size_t i = 0; i < ntri; i++) order[i] = (uint32_t)i;
    m.nodes.clear();
    m.nodes.reserve(ntri / 2 + 16);
    struct Job
    {
        uint32_t node, first, count;
    };
    std::vector<Job> stack;
    m.nodes.push_back(BNode());
    stack.push_back({ 0, 0, (uint32_t)ntri });
    while (!stack.empty())
    {
        Job j = stack.back();
        stack.pop_back();
        BNode n;
        float clo[3] = { 1e30f, 1e30f, 1e30f }, chi[3] = { -1e30f, -1e30f, -1e30f };
        for (int a = 0; a < 3; a++) n.lo[a] = 1e30f, n.hi[a] = -1e30f;
        for (uint32_t i = j.first; i < j.first + j.count; i++)
            for (int a = 0; a < 3; a++)
            {
                uint32_t tI = order[i];
                n.lo[a]     = std::min(n.lo[a], blo[tI * 3 + a]);
                n.hi[a]     = std::max(n.hi[a], bhi[tI * 3 + a]);
                clo[a]      = std::min(clo[a], cen[tI * 3 + a]);
                chi[a]      = std::max(chi[a], cen[tI * 3 + a]);
            }
        n.count = j.count;
        if (j.count <= 4)
        {
            n.left  = j.first;
            n.right = 0xFFFFFFFFu;
        }
        else
        {
            int axis = 0;
            if (chi[1] - clo[1] > chi[axis] - clo[axis]) axis = 1;
            if (chi[2] - clo[2] > chi[axis] - clo[axis]) axis = 2;
            uint32_t mid = j.count / 2;
            std::nth_element(order.begin() + j.first, order.begin() + j.first + mid, order.begin() + j.first + j.count,
                             [&](uint32_t a, uint32_t b) { return cen[a * 3 + axis] < cen[b * 3 + axis]; });
            n.left  = (uint32_t)m.nodes.size();
            n.right = n.left + 1;
            m.nodes.push_back(BNode());
            m.nodes.push_back(BNode());
            stack.push_back({ n.left, j.first, mid });
            stack.push_back({ n.right, j.first + mid, j.count - mid });
        }
        m.nodes[j.node] = n;
    }
    std::vector<uint32_t> g2(ntri), p2(ntri);
    for (size_t i = 0; i < ntri; i++) g2[i] = m.tri_geom[order[i]], p2[i] = m.tri_prim[order[i]];
    m.tri_geom.swap(g2);
    m.tri_prim.swap(p2);
    if (ntri)
        for (int a = 0; a < 3; a++) m.lo[a] = m.nodes[0].lo[a], m.hi[a] = m.nodes[0].hi[a];
    else
        for (int a = 0; a < 3; a++) m.lo[a] = 0, m.hi[a] = 0;
}

// conservative slab test in double with a ray-distance-proportional margin so that no triangle the
// fp32 triangle test would accept is culled (the BVH must return exactly the brute-force result)
static inline bool box_test(const float* lo, const float* hi, vec3 o, vec3 d, float tmin, float tmax)
{
    double t0 = tmin, t1 = tmax;
    // distance over ALL axes: a ray that grazes an axis-aligned wall is close to the box in the wall's axis and far away
    // in the others, and the triangle test's placement error scales with the far ones
    double dist = 0.0, ext = 0.0;
    for (int a = 0; a < 3; a++)
    {
        dist = std::max(dist, std::max(std::fabs((double)o[a] - lo[a]), std::fabs((double)o[a] - hi[a])));
        ext  = std::max(ext, (double)hi[a] - (double)lo[a]);
    }
    const double margin = 1e-4 * (dist + ext) + 1e-30;
    for (int a = 0; a < 3; a++)
    {
        double oa = o[a], da = d[a];
        double l = lo[a] - margin, h = hi[a] + margin;
        if (da == 0.0)
        {
            if (oa < l || oa > h) return false;
            continue;
        }
        double ta = (l - oa) / da, tb = (h - oa) / da;
        if (ta > tb) std::swap(ta, tb);
        // widen in t as well
        ta -= 1e-6 * std::fabs(ta);
        tb += 1e-6 * std::fabs(tb);
        if (ta > t0) t0 = ta;
        if (tb < t1) t1 = tb;
        if (t0 > t1) return false;
    }
    return true;
}

static vec4 fetch_albedo_rahit(const Scene& s, const Material& m, float tu, float tv) // path_trace_rahit.glsl:162-168
{
    if (m.texture_indices0[0] == -1) return vec4(m.albedo[0], m.albedo[1], m.albedo[2], m.albedo[3]);
    return texture_lod0(s.textures[m.texture_indices0[0]], tu, tv);
}

// path_trace_rahit.glsl:174-188 — returns true when the candidate is IGNORED
static bool any_hit_ignores(const Scene& s, uint32_t inst, uint32_t geom, uint32_t prim, float bu, float bv)
{
    const Instance& instance = s.instances[inst];
    const Mesh&     mesh     = *s.meshes[instance.mesh_idx];
    uint32_t        prim_off = s.submesh_info[inst][geom * 2 + 0];
    uint32_t        mat_idx  = s.submesh_info[inst][geom * 2 + 1];
    uint32_t        pid      = prim + prim_off;
    const Vertex&   v0       = mesh.verts[mesh.indices[3 * pid + 0]];
    const Vertex&   v1       = mesh.verts[mesh.indices[3 * pid + 1]];
    const Vertex&   v2       = mesh.verts[mesh.indices[3 * pid + 2]];
    float           b0 = 1.0f - bu - bv, b1 = bu, b2 = bv;
    // common.glsl:135 (only tex_coord is needed by the any-hit shader)
    float tu     = v0.tex_coord[0] * b0 + v1.tex_coord[0] * b1 + v2.tex_coord[0] * b2;
    float tv     = v0.tex_coord[1] * b0 + v1.tex_coord[1] * b1 + v2.tex_coord[1] * b2;
    vec4  albedo = fetch_albedo_rahit(s, s.materials[mat_idx], tu, tv);
    return albedo.w < 0.1f;
}

static inline bool hit_better(float t, uint32_t i, uint32_t g, uint32_t p, const Hit& best)
{
    if (!best.valid) return true;
    if (t < best.t) return true;
    if (t > best.t) return false;
    if (i != best.instance) return i < best.instance;
    if (g != best.geometry) return g < best.geometry;
    return p < best.primitive;
}

// oracle/ref_glsl/runtime.cpp (the reference's own any-hit shader compiled as C++) installs itself here; the
// restatement's any_hit_ignores() is used otherwise
typedef bool (*AnyHitOverride)(uint32_t inst, uint32_t geom, uint32_t prim, float bu, float bv);
static thread_local AnyHitOverride t_any_hit_override = nullptr;

// traceRayEXT: closest accepted hit with tmin < t < tmax; ties -> min (instance, geometry, primitive)
static Hit trace(const Scene& s, vec3 o, float tmin, vec3 d, float tmax, uint32_t flags)
{
    Hit best;
    best.valid = false;
    best.t     = tmax;
    best.u = best.v = 0;
    best.instance = best.geometry = best.primitive = 0xFFFFFFFFu;
    const bool ray_opaque = (flags & RAY_FLAG_OPAQUE) != 0;
    for (uint32_t ii = 0; ii < s.instances.size(); ii++)
    {
        if (!s.brute_force && !box_test(&s.wbox[ii * 6], &s.wbox[ii * 6 + 3], o, d, tmin, best.t)) continue;
        const Mesh&  mesh = *s.meshes[s.instances[ii].mesh_idx];
        const float* iv   = &s.inv[ii * 12];
        vec3         oo, od;
        oo.x = (iv[0] * o.x + iv[1] * o.y + iv[2] * o.z) + iv[3];
        oo.y = (iv[4] * o.x + iv[5] * o.y + iv[6] * o.z) + iv[7];
        oo.z = (iv[8] * o.x + iv[9] * o.y + iv[10] * o.z) + iv[11];
        od.x = iv[0] * d.x + iv[1] * d.y + iv[2] * d.z;
        od.y = iv[4] * d.x + iv[5] * d.y + iv[6] * d.z;
        od.z = iv[8] * d.x + iv[9] * d.y + iv[10] * d.z;

        auto test_tri = [&](uint32_t g, uint32_t p) {
            const SubMesh& sm = mesh.subs[g];
            const float*   a  = mesh.verts[mesh.indices[sm.base_index + 3 * p + 0]].position;
            const float*   b  = mesh.verts[mesh.indices[sm.base_index + 3 * p + 1]].position;
            const float*   c  = mesh.verts[mesh.indices[sm.base_index + 3 * p + 2]].position;
            float          t, u, v;
            if (!tri_test(oo, od, vec3(a[0], a[1], a[2]), vec3(b[0], b[1], b[2]), vec3(c[0], c[1], c[2]), t, u, v)) return;
            if (!(t > tmin && t < tmax)) return;
            if (!hit_better(t, ii, g, p, best)) return;
            if (!ray_opaque && !sm.opaque && (t_any_hit_override ? t_any_hit_override(ii, g, p, u, v) : any_hit_ignores(s, ii, g, p, u, v))) return;
            best.valid     = true;
            best.t         = t;
            best.u         = u;
            best.v         = v;
            best.instance  = ii;
            best.geometry  = g;
            best.primitive = p;
        };

        if (s.brute_force)
        {
            for (uint32_t g = 0; g < mesh.subs.size(); g++)
                for (uint32_t p = 0; p < mesh.subs[g].index_count / 3; p++) test_tri(g, p);
            continue;
        }
        if (mesh.nodes.empty() || mesh.nodes[0].count == 0) continue;
        uint32_t stack[128];
        int      sp  = 0;
        stack[sp++]  = 0;
        while (sp)
        {
            const BNode& n = mesh.nodes[stack[--sp]];
            if (!box_test(n.lo, n.hi, oo, od, tmin, best.t)) continue;
            if (n.right == 0xFFFFFFFFu)
            {
                for (uint32_t i = n.left; i < n.left + n.count; i++) test_tri(mesh.tri_geom[i], mesh.tri_prim[i]);
            }
            else
            {
                stack[sp++] = n.left;
                stack[sp++] = n.right;
            }
        }
    }
    return best;
}

// world -> object 3x4 from the instance model matrix, in double, rounded once to float
static void affine_inverse(const float* m /* column-major 4x4 */, float* out /* 3x4 row-major */)
{
    double a00 = m[0], a01 = m[4], a02 = m[8], t0 = m[12];
    double a10 = m[1], a11 = m[5], a12 = m[9], t1 = m[13];
    double a20 = m[2], a21 = m[6], a22 = m[10], t2 = m[14];
    double c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
    double det = a00 * c00 + a01 * c01 + a02 * c02;
    double id  = 1.0 / det;
    double i00 = c00 * id, i01 = (a02 * a21 - a01 * a22) * id, i02 = (a01 * a12 - a02 * a11) * id;
    double i10 = c01 * id, i11 = (a00 * a22 - a02 * a20) * id, i12 = (a02 * a10 - a00 * a12) * id;
    double i20 = c02 * id, i21 = (a01 * a20 - a00 * a21) * id, i22 = (a00 * a11 - a01 * a10) * id;
    out[0]  = (float)i00, out[1] = (float)i01, out[2] = (float)i02, out[3] = (float)(-(i00 * t0 + i01 * t1 + i02 * t2));
    out[4]  = (float)i10, out[5] = (float)i11, out[6] = (float)i12, out[7] = (float)(-(i10 * t0 + i11 * t1 + i12 * t2));
    out[8]  = (float)i20, out[9] = (float)i21, out[10] = (float)i22, out[11] = (float)(-(i20 * t0 + i21 * t1 + i22 * t2));
}

// ------------------------------------------------------------------------------------------------
// brdf.glsl
// ------------------------------------------------------------------------------------------------
struct SurfaceProperties // common.glsl:67-78
{
    vec3  position; // vertex.position.xyz
    vec2  tex_coord;
    vec3  vnormal, tangent, bitangent;
    vec4  albedo;
    vec3  emissive;
    vec3  normal;
    vec3  F0;
    float metallic, roughness;
};

static vec3 sample_cosine_lobe(vec3 n, vec2 r) // brdf.glsl:6-18
{
    float rx = std::fmax(0.00001f, r.x), ry = std::fmax(0.00001f, r.y);
    const float phi       = 2.0f * M_PI_F * ry;
    const float cos_theta = std::sqrt(rx);
    const float sin_theta = std::sqrt(1 - rx);
    vec3        t         = vec3(sin_theta * std::cos(phi), sin_theta * std::sin(phi), cos_theta);
    vec3        x, y;
    make_rotation_matrix(n, x, y);
    return normalize(mul_cols(x, y, n, t));
}
static float pdf_cosine_lobe(float ndotl) { return ndotl / M_PI_F; }        // brdf.glsl:20-23
static vec3  evaluate_lambert(vec3 albedo) { return albedo / M_PI_F; }      // brdf.glsl:30-33
static float triangle_area(vec3 p0, vec3 p1, vec3 p2)                        // brdf.glsl:35-38
{
    return 0.5f * length(cross(p1 - p0, p2 - p0));
}
static vec2 uniform_sample_triangle(vec2 u) // brdf.glsl:40-44
{
    float su0 = std::sqrt(u.x);
    vec2  r;
    r.x = 1 - su0;
    r.y = u.y * su0;
    return r;
}
static vec3 barycentric_interpolate(vec2 b, vec3 v0, vec3 v1, vec3 v2) // brdf.glsl:46-51
{
    const vec3 bc = vec3(1.0f - b.x - b.y, b.x, b.y);
    return v0 * bc.x + v1 * bc.y + v2 * bc.z;
}
static float pdf_triangle(float distance_sqr, float cos_theta, float area) // brdf.glsl:53-56
{
    return distance_sqr / std::fmax(EPSILON, cos_theta * area);
}
static float D_ggx(float ndoth, float alpha) // brdf.glsl:58-64
{
    float a2    = alpha * alpha;
    float denom = (ndoth * ndoth) * (a2 - 1.0f) + 1.0f;
    return a2 / std::fmax(EPSILON, (M_PI_F * denom * denom));
}
static float G1_schlick_ggx(float roughness, float ndotv) // brdf.glsl:66-71
{
    float k = ((roughness + 1) * (roughness + 1)) / 8.0f;
    return ndotv / std::fmax(EPSILON, (ndotv * (1 - k) + k));
}
static float G_schlick_ggx(float ndotl, float ndotv, float roughness) // brdf.glsl:73-76
{
    return G1_schlick_ggx(roughness, ndotl) * G1_schlick_ggx(roughness, ndotv);
}
static vec3 F_schlick(vec3 f0, float vdoth) // brdf.glsl:78-81
{
    return f0 + (vec3(1.0f) - f0) * (std::pow(1.0f - vdoth, 5.0f));
}
static vec3 sample_ggx(vec3 n, float alpha, vec2 Xi) // brdf.glsl:83-97
{
    float phi       = 2.0f * M_PI_F * Xi.x;
    float cos_theta = std::sqrt((1.0f - Xi.y) / (1.0f + (alpha * alpha - 1.0f) * Xi.y));
    float sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
    vec3  d;
    d.x = sin_theta * std::cos(phi);
    d.y = sin_theta * std::sin(phi);
    d.z = cos_theta;
    vec3 x, y;
    make_rotation_matrix(n, x, y);
    return normalize(mul_cols(x, y, n, d));
}
static vec3 evaluate_ggx(const SurfaceProperties& p, vec3 F, float ndoth, float ndotl, float ndotv) // brdf.glsl:99-103
{
    float alpha = p.roughness * p.roughness;
    return (D_ggx(ndoth, alpha) * F * G_schlick_ggx(ndotl, ndotv, p.roughness)) / std::fmax(EPSILON, (4.0f * ndotl * ndotv));
}
static float pdf_D_ggx(float alpha, float ndoth, float vdoth) // brdf.glsl:105-108
{
    return D_ggx(ndoth, alpha) * ndoth / std::fmax(EPSILON, (4.0f * vdoth));
}
static vec3 evaluate_uber(const SurfaceProperties& p, vec3 Wo, vec3 Wh, vec3 Wi) // brdf.glsl:110-122
{
    float NdotL = std::fmax(dot(p.normal, Wi), 0.0f);
    float NdotV = std::fmax(dot(p.normal, Wo), 0.0f);
    float NdotH = std::fmax(dot(p.normal, Wh), 0.0f);
    float VdotH = std::fmax(dot(Wi, Wh), 0.0f);
    vec3  F        = F_schlick(p.F0, VdotH);
    vec3  specular = evaluate_ggx(p, F, NdotH, NdotL, NdotV);
    vec3  diffuse  = evaluate_lambert(p.albedo.xyz());
    return (vec3(1.0f) - F) * diffuse + specular;
}
static float pdf_uber(const SurfaceProperties& p, vec3 Wo, vec3 Wh, vec3 Wi) // brdf.glsl:124-135
{
    float NdotL = std::fmax(dot(p.normal, Wi), 0.0f);
    float NdotH = std::fmax(dot(p.normal, Wh), 0.0f);
    float VdotH = std::fmax(dot(Wi, Wh), 0.0f);
    (void)Wo;
    float pd = pdf_cosine_lobe(NdotL);
    float ps = pdf_D_ggx(p.roughness * p.roughness, NdotH, VdotH);
    return mixf(pd, ps, 0.5f);
}
// brdf.glsl:137-167 — NOTE `in RNG rng`: the generator is taken BY VALUE
static vec3 sample_uber(const SurfaceProperties& p, vec3 Wo, RNG rng, vec3& Wi, float& pdf)
{
    float alpha = p.roughness * p.roughness;
    vec3  Wh;
    vec3  rand_value  = next_vec3(rng);
    bool  is_specular = false;
    vec2  ryz;
    ryz.x = rand_value.y;
    ryz.y = rand_value.z;
    if (rand_value.x < 0.5f)
    {
        Wh          = sample_ggx(p.normal, alpha, ryz);
        Wi          = reflect(-Wo, Wh);
        float NdotL = std::fmax(dot(p.normal, Wi), 0.0f);
        float NdotV = std::fmax(dot(p.normal, Wo), 0.0f);
        if (NdotL > 0.0f && NdotV > 0.0f) is_specular = true;
    }
    if (!is_specular)
    {
        Wi = sample_cosine_lobe(p.normal, ryz);
        Wh = normalize(Wo + Wi);
    }
    pdf = pdf_uber(p, Wo, Wh, Wi);
    return evaluate_uber(p, Wo, Wh, Wi);
}

// ------------------------------------------------------------------------------------------------
// path_trace_rchit.glsl
// ------------------------------------------------------------------------------------------------
struct PathTracePayload // common.glsl:26-35
{
    vec3     L, T;
    uint32_t depth;
    RNG      rng;
    vec3     debug_color; // RAY_DEBUG_VIEW only (common.glsl:32-34)
};
struct Counters
{
    uint64_t extension_rays = 0, shadow_rays = 0;
};
struct TraceCtx
{
    const Scene*         scene;
    const PushConstants* pc;
    Counters*            counters;
    // non-null = the RAY_DEBUG_VIEW variant of the pipeline (path_integrator.cpp:259-307): DebugRayVertexBuffer, 8 floats
    // per vertex (position.xyz 1, color.rgb 1); DebugRayDrawArgs.count = size() / 8
    std::vector<float>* debug_vertices = nullptr;
};
static void debug_ray_segment(const TraceCtx& c, const PathTracePayload& payload, vec3 a, vec3 b) // rchit:551-565, rmiss:43-57
{
    const float v[16] = { a.x, a.y, a.z, 1.0f, payload.debug_color.x, payload.debug_color.y, payload.debug_color.z, 1.0f,
                          b.x, b.y, b.z, 1.0f, payload.debug_color.x, payload.debug_color.y, payload.debug_color.z, 1.0f };
    c.debug_vertices->insert(c.debug_vertices->end(), v, v + 16);
}

static inline bool is_black(vec3 c) { return c.x == 0.0f && c.y == 0.0f && c.z == 0.0f; } // common.glsl:123-126
static inline vec3 v3(const float* p) { return vec3(p[0], p[1], p[2]); }

static void fetch_triangle(const Scene& s, const Instance& instance, uint32_t prim_off, uint32_t prim_id, const Vertex*& v0, const Vertex*& v1, const Vertex*& v2) // rchit:156-172
{
    const Mesh& mesh = *s.meshes[instance.mesh_idx];
    uint32_t    pid  = prim_id + prim_off;
    v0               = &mesh.verts[mesh.indices[3 * pid + 0]];
    v1               = &mesh.verts[mesh.indices[3 * pid + 1]];
    v2               = &mesh.verts[mesh.indices[3 * pid + 2]];
}

static vec3 get_normal_from_map(const Scene& s, vec3 tangent, vec3 bitangent, vec3 normal, vec2 tc, uint32_t idx) // rchit:190-202
{
    vec3 T = normalize(tangent), B = normalize(bitangent), N = normalize(normal);
    vec4 tx = texture_lod0(s.textures[idx], tc.x, tc.y);
    vec3 n  = normalize(vec3(tx.x, tx.y, tx.z) * 2.0f - vec3(1.0f));
    n       = normalize(mul_cols(T, B, N, n));
    return n;
}

static void populate_surface_properties(const Scene& s, const Hit& hit, SurfaceProperties& p) // rchit:256-280
{
    const Instance& instance = s.instances[hit.instance];
    uint32_t        prim_off = s.submesh_info[hit.instance][hit.geometry * 2 + 0]; // fetch_hit_info rchit:141-152
    uint32_t        mat_idx  = s.submesh_info[hit.instance][hit.geometry * 2 + 1];
    const Vertex *  v0, *v1, *v2;
    fetch_triangle(s, instance, prim_off, hit.primitive, v0, v1, v2);
    const Material& material = s.materials[mat_idx];
    const vec3      b        = vec3(1.0f - hit.u - hit.v, hit.u, hit.v);

    // interpolated_vertex, common.glsl:130-141
    vec3 pos     = v3(v0->position) * b.x + v3(v1->position) * b.y + v3(v2->position) * b.z;
    p.tex_coord.x = v0->tex_coord[0] * b.x + v1->tex_coord[0] * b.y + v2->tex_coord[0] * b.z;
    p.tex_coord.y = v0->tex_coord[1] * b.x + v1->tex_coord[1] * b.y + v2->tex_coord[1] * b.z;
    vec3 nrm = normalize(v3(v0->normal) * b.x + v3(v1->normal) * b.y + v3(v2->normal) * b.z);
    vec3 tan = normalize(v3(v0->tangent) * b.x + v3(v1->tangent) * b.y + v3(v2->tangent) * b.z);
    vec3 bit = normalize(v3(v0->bitangent) * b.x + v3(v1->bitangent) * b.y + v3(v2->bitangent) * b.z);

    // transform_vertex, rchit:176-186 (position.w = 1.0 from interpolated_vertex)
    mat4 model  = mat4_from(instance.model_matrix);
    mat4 normal = mat4_from(instance.normal_matrix);
    p.position  = mul(model, vec4(pos, 1.0f)).xyz();
    p.vnormal   = mul3(normal, nrm);
    p.tangent   = mul3(normal, tan);
    p.bitangent = mul3(normal, bit);

    // fetch_albedo .. fetch_emissive, rchit:206-252
    if (material.texture_indices0[0] == -1)
        p.albedo = vec4(material.albedo[0], material.albedo[1], material.albedo[2], material.albedo[3]);
    else
        p.albedo = texture_lod0(s.textures[material.texture_indices0[0]], p.tex_coord.x, p.tex_coord.y);
    if (material.texture_indices0[1] == -1)
        p.normal = p.vnormal;
    else
        p.normal = get_normal_from_map(s, p.tangent, p.bitangent, p.vnormal, p.tex_coord, material.texture_indices0[1]);
    if (material.texture_indices0[2] == -1)
        p.roughness = material.roughness_metallic[0];
    else
        p.roughness = texture_lod0(s.textures[material.texture_indices0[2]], p.tex_coord.x, p.tex_coord.y)[material.texture_indices1[2] & 3];
    if (material.texture_indices0[3] == -1)
        p.metallic = material.roughness_metallic[1];
    else
        p.metallic = texture_lod0(s.textures[material.texture_indices0[3]], p.tex_coord.x, p.tex_coord.y)[material.texture_indices1[3] & 3];
    if (material.texture_indices1[0] == -1)
        p.emissive = v3(material.emissive);
    else
        p.emissive = texture_lod0(s.textures[material.texture_indices1[0]], p.tex_coord.x, p.tex_coord.y).xyz();

    p.roughness = std::fmax(p.roughness, MIN_ROUGHNESS);
    p.F0        = mix(vec3(0.03f), p.albedo.xyz(), p.metallic);
}

static void disk_jitter(vec3 light_dir, float light_radius, vec2 rng, vec3& Wi) // shared body of rchit:305-316 / 334-341 / 355-362
{
    vec3  light_tangent   = normalize(cross(light_dir, vec3(0.0f, 1.0f, 0.0f)));
    vec3  light_bitangent = normalize(cross(light_tangent, light_dir));
    float point_radius    = light_radius * std::sqrt(rng.x);
    float point_angle     = rng.y * 2.0f * M_PI_F;
    float dx = point_radius * std::cos(point_angle), dy = point_radius * std::sin(point_angle);
    Wi = normalize(light_dir + dx * light_tangent + dy * light_bitangent);
}

// rchit:284-451
static vec3 sample_light(const TraceCtx& c, PathTracePayload& payload, const SurfaceProperties& p, const Light& light, vec3& Wi, float& pdf)
{
    const Scene& s     = *c.scene;
    uint32_t ray_flags = RAY_FLAG_OPAQUE | RAY_FLAG_TERMINATE;
    if (payload.depth == 0) ray_flags = 0; // any-hit only at the first hit
    float tmin   = 0.0001f;
    float tmax   = 10000.0f;
    vec3  origin = p.position + p.normal * c.pc->shadow_ray_bias;
    vec3  Li     = vec3(0.0f);
    uint32_t type = (uint32_t)light.d0[0];

    if (type == 0) // LIGHT_DIRECTIONAL
    {
        vec2 rng       = next_vec2(payload.rng);
        vec3 light_dir = -v3(light.d1);
        disk_jitter(light_dir, light.d2[3], rng, Wi);
        Li  = vec3(light.d0[1], light.d0[2], light.d0[3]) * light.d1[3];
        pdf = 0.0f;
    }
    else if (type == 1) // LIGHT_SPOT
    {
        vec2  rng            = next_vec2(payload.rng);
        vec3  to_light       = v3(light.d2) - p.position;
        vec3  light_dir      = normalize(to_light);
        float light_distance = length(to_light);
        float light_radius   = light.d2[3] / light_distance;
        float angle_attenuation = dot(light_dir, -v3(light.d1));
        angle_attenuation       = smoothstep(light.d3[1], light.d3[0], angle_attenuation);
        disk_jitter(light_dir, light_radius, rng, Wi);
        Li   = vec3(light.d0[1], light.d0[2], light.d0[3]) * light.d1[3] * angle_attenuation / (light_distance * light_distance);
        pdf  = 0.0f;
        tmax = light_distance;
    }
    else if (type == 2) // LIGHT_POINT
    {
        vec2  rng            = next_vec2(payload.rng);
        vec3  to_light       = v3(light.d2) - p.position;
        vec3  light_dir      = normalize(to_light);
        float light_distance = length(to_light);
        float light_radius   = light.d2[3] / light_distance;
        disk_jitter(light_dir, light_radius, rng, Wi);
        Li   = vec3(light.d0[1], light.d0[2], light.d0[3]) * light.d1[3] / (light_distance * light_distance);
        pdf  = 0.0f;
        tmax = light_distance;
    }
    else if (type == 3) // LIGHT_ENVIRONMENT_MAP
    {
        vec2 rand_value = next_vec2(payload.rng);
        Wi              = sample_cosine_lobe(p.normal, rand_value);
        Li              = env_sample(s.env, Wi);
        pdf             = pdf_cosine_lobe(dot(p.normal, Wi));
    }
    else if (type == 4) // LIGHT_AREA
    {
        uint32_t mesh_id       = (uint32_t)light.d0[1];
        uint32_t num_triangles = (uint32_t)light.d1[2]; // quirk A.8-2: host wrote the count to .x
        uint32_t primitive_id  = next_uint(payload.rng, num_triangles);
        uint32_t mat_idx       = (uint32_t)light.d0[2];
        uint32_t prim_off      = (uint32_t)light.d0[3];
        const Instance& instance = s.instances[mesh_id];
        const Material& material = s.materials[mat_idx];
        const Vertex *  v0, *v1, *v2;
        fetch_triangle(s, instance, prim_off, primitive_id, v0, v1, v2);
        vec2 b     = uniform_sample_triangle(next_vec2(payload.rng));
        mat4 model = mat4_from(instance.model_matrix);
        // quirk A.8-3: the vertex's own w (= submesh index) is used
        vec3 p0 = mul(model, vec4(v0->position[0], v0->position[1], v0->position[2], v0->position[3])).xyz();
        vec3 p1 = mul(model, vec4(v1->position[0], v1->position[1], v1->position[2], v1->position[3])).xyz();
        vec3 p2 = mul(model, vec4(v2->position[0], v2->position[1], v2->position[2], v2->position[3])).xyz();
        vec3 light_position = barycentric_interpolate(b, p0, p1, p2);
        vec3 light_normal   = normalize(mul3(mat4_from(instance.normal_matrix), barycentric_interpolate(b, v3(v0->normal), v3(v1->normal), v3(v2->normal))));
        vec3 light_dir      = p.position - light_position;
        float dist_sqr      = dot(light_dir, light_dir);
        float area          = triangle_area(p0, p1, p2);
        if (area == 0.0f || dist_sqr == 0.0f)
        {
            pdf = 0.0f;
            return vec3(0.0f);
        }
        float dist = std::sqrt(dist_sqr);
        light_dir /= dist;
        tmax            = std::fmax(0.0f, dist - EPSILON);
        float cos_theta = dot(light_normal, light_dir);
        if (cos_theta == 0.0f)
        {
            pdf = 0.0f;
            return vec3(0.0f);
        }
        Li  = v3(material.emissive);
        Wi  = -light_dir;
        pdf = pdf_triangle(dist_sqr, cos_theta, area);
    }

    // visibility ray: hit group 1 / miss 1 (path_trace_shadow.rchit / .rmiss)
    if (c.counters) c.counters->shadow_rays++;
    Hit  h          = trace(s, origin, tmin, Wi, tmax, ray_flags);
    bool visibility = !h.valid;
    return Li * (visibility ? 1.0f : 0.0f);
}

static void closest_hit(const TraceCtx& c, PathTracePayload& payload, vec3 ray_origin, vec3 ray_dir, const Hit& hit);

// traceRayEXT on the path-trace hit group: miss shader or closest-hit shader
static void trace_path(const TraceCtx& c, PathTracePayload& payload, vec3 origin, float tmin, vec3 dir, float tmax, uint32_t flags)
{
    if (c.counters) c.counters->extension_rays++;
    Hit h = trace(*c.scene, origin, tmin, dir, tmax, flags);
    if (!h.valid)
    {
        if (c.debug_vertices) // rmiss:40-58: the segment of a secondary ray that leaves the scene, nothing else
        {
            if (payload.depth > 0) debug_ray_segment(c, payload, origin, origin + dir * tmax);
            return;
        }
        // path_trace_rmiss.glsl:60-65
        vec3 e = env_sample(c.scene->env, dir);
        if (payload.depth == 0)
            payload.L = e;
        else
            payload.L = payload.T * e;
        return;
    }
    closest_hit(c, payload, origin, dir, h);
}

static vec3 direct_lighting(const TraceCtx& c, PathTracePayload& payload, vec3 ray_dir, const SurfaceProperties& p) // rchit:455-483
{
    vec3     L         = vec3(0.0f);
    uint32_t light_idx = next_uint(payload.rng, c.pc->num_lights);
    Light    light;
    if (light_idx < c.scene->lights.size())
        light = c.scene->lights[light_idx];
    else
        std::memset(&light, 0, sizeof(light)); // SURVEY C-4: undefined upstream; result is multiplied by 0
    vec3  Wo  = -ray_dir;
    vec3  Wi  = vec3(0.0f);
    float pdf = 0.0f;
    vec3  Li  = sample_light(c, payload, p, light, Wi, pdf);
    vec3  Wh  = normalize(Wo + Wi);
    vec3  brdf      = evaluate_uber(p, Wo, Wh, Wi);
    float cos_theta = clampf(dot(p.normal, Wi), 0.0f, 1.0f);
    if (!is_black(Li))
    {
        if (pdf == 0.0f)
            L = payload.T * brdf * cos_theta * Li;
        else
            L = (payload.T * brdf * cos_theta * Li) / pdf;
    }
    return L * (float)c.pc->num_lights;
}

static vec3 indirect_lighting(const TraceCtx& c, PathTracePayload& payload, vec3 ray_dir, const SurfaceProperties& p) // rchit:487-536
{
    vec3  Wo = -ray_dir;
    vec3  Wi;
    float pdf;
    vec3  brdf      = sample_uber(p, Wo, payload.rng /* by value */, Wi, pdf);
    float cos_theta = clampf(dot(p.normal, Wi), 0.0f, 1.0f);
    PathTracePayload indirect;
    indirect.L = vec3(0.0f);
    indirect.T = payload.T * (brdf * cos_theta) / pdf;
    if (!c.debug_vertices) // rchit:500-509: #if !defined(RAY_DEBUG_VIEW)
    {
        // Russian roulette
        float probability = std::fmax(indirect.T.x, std::fmax(indirect.T.y, indirect.T.z));
        if (next_float(payload.rng) > probability) return vec3(0.0f);
        indirect.T *= 1.0f / probability;
    }
    indirect.depth = payload.depth + 1;
    indirect.rng   = payload.rng;
    indirect.debug_color = payload.debug_color; // rchit:512-514
    trace_path(c, indirect, p.position, 0.0001f, Wi, 10000.0f, RAY_FLAG_OPAQUE);
    return indirect.L;
}

static void closest_hit(const TraceCtx& c, PathTracePayload& payload, vec3 ray_origin, vec3 ray_dir, const Hit& hit) // rchit:542-580
{
    SurfaceProperties p;
    populate_surface_properties(*c.scene, hit, p);
    if (c.debug_vertices && payload.depth > 0) debug_ray_segment(c, payload, ray_origin, ray_origin + ray_dir * hit.t); // rchit:548-567
    payload.L = vec3(0.0f);
    if (payload.depth == 0 && !is_black(p.emissive)) payload.L += p.emissive;
    payload.L += direct_lighting(c, payload, ray_dir, p);
    if ((payload.depth + 1) < c.pc->max_ray_bounces) payload.L += indirect_lighting(c, payload, ray_dir, p);
}

// ------------------------------------------------------------------------------------------------
// path_trace_rgen.glsl
// ------------------------------------------------------------------------------------------------
static void generate_ray(const PushConstants& pc, RNG& rng, uint32_t lx, uint32_t ly, vec3& origin, vec3& direction, bool ray_debug_view = false) // rgen:132-174
{
    float pcx = (float)lx + 0.5f, pcy = (float)ly + 0.5f;
    if (ray_debug_view) pcx = (float)pc.ray_debug_pixel_coord[0] + 0.5f, pcy = (float)pc.ray_debug_pixel_coord[1] + 0.5f; // rgen:137-138
    float jx  = next_float(rng);
    float jy  = next_float(rng);
    float tcx = (pcx + jx) / (ray_debug_view ? (float)pc.ray_debug_pixel_coord[2] : (float)pc.launch_id_size[2]); // rgen:143-147
    float tcy = (pcy + jy) / (ray_debug_view ? (float)pc.ray_debug_pixel_coord[3] : (float)pc.launch_id_size[3]);
    float nx = tcx * 2.0f - 1.0f, ny = tcy * 2.0f - 1.0f;
    vec3  cam    = v3(pc.camera_pos);
    vec4  target = mul(mat4_from(pc.view_proj_inverse), vec4(nx, ny, 0.0f, 1.0f));
    target       = target / target.w;
    // Aperture offset
    float angle  = next_float(rng) * 2.0f * M_PI_F;
    float radius = std::sqrt(next_float(rng));
    float offx   = std::cos(angle) * radius * pc.aperture_radius;
    float offy   = std::sin(angle) * radius * pc.aperture_radius;
    vec3  aperture_pos = cam + v3(pc.right_direction) * offx + v3(pc.up_direction) * offy;
    vec3  rstart       = cam;
    vec3  rdir         = -normalize(target.xyz() - cam);
    vec3  fp           = v3(pc.focal_plane);
    float t            = -(dot(rstart, fp) + pc.focal_plane[3]) / dot(rdir, fp);
    vec3  focus_pos    = rstart + rdir * t;
    origin             = aperture_pos;
    direction          = normalize(focus_pos - aperture_pos);
}

// ------------------------------------------------------------------------------------------------
// C ABI for ctypes
// ------------------------------------------------------------------------------------------------
OR_API Scene* or_scene_new()
{
    init_luts();
    return new Scene();
}
OR_API void or_scene_free(Scene* s) { delete s; }
OR_API void or_scene_set_brute_force(Scene* s, int on) { s->brute_force = on != 0; }
OR_API int  or_scene_add_mesh(Scene* s, const Vertex* v, uint32_t nv, const uint32_t* idx, uint32_t ni, const SubMesh* subs, uint32_t ns)
{
    Mesh* m = new Mesh();
    m->verts.assign(v, v + nv);
    m->indices.assign(idx, idx + ni);
    m->subs.assign(subs, subs + ns);
    mesh_build_bvh(*m);
    s->meshes.push_back(m);
    return (int)s->meshes.size() - 1;
}
OR_API int or_scene_add_texture(Scene* s, int format, uint32_t w, uint32_t h, const void* data)
{
    Texture t;
    t.format = format;
    t.w      = w;
    t.h      = h;
    size_t n = (size_t)w * h * (format == 3 ? 16 : 4);
    t.data.assign((const uint8_t*)data, (const uint8_t*)data + n);
    s->textures.push_back(std::move(t));
    return (int)s->textures.size() - 1;
}
OR_API void or_scene_set_envmap(Scene* s, uint32_t size, const float* faces)
{
    s->env.size = size;
    s->env.data.assign(faces, faces + (size_t)6 * size * size * 4);
}
OR_API void or_scene_set_tables(Scene* s, const Material* mats, uint32_t nm, const Instance* inst, const uint32_t* const* submesh_info, uint32_t ni, const Light* lights, uint32_t nl)
{
    s->materials.assign(mats, mats + nm);
    s->instances.assign(inst, inst + ni);
    s->lights.assign(lights, lights + nl);
    s->submesh_info.clear();
    s->inv.resize((size_t)ni * 12);
    s->wbox.resize((size_t)ni * 6);
    for (uint32_t i = 0; i < ni; i++)
    {
        const Mesh& m = *s->meshes[inst[i].mesh_idx];
        s->submesh_info.emplace_back(submesh_info[i], submesh_info[i] + 2 * m.subs.size());
        affine_inverse(inst[i].model_matrix, &s->inv[(size_t)i * 12]);
        // world AABB of the transformed object box (8 corners, double)
        double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
        const float* M = inst[i].model_matrix;
        for (int c = 0; c < 8; c++)
        {
            double x = (c & 1) ? m.hi[0] : m.lo[0], y = (c & 2) ? m.hi[1] : m.lo[1], z = (c & 4) ? m.hi[2] : m.lo[2];
            for (int a = 0; a < 3; a++)
            {
                double w = (double)M[a] * x + (double)M[4 + a] * y + (double)M[8 + a] * z + (double)M[12 + a];
                lo[a]    = std::min(lo[a], w);
                hi[a]    = std::max(hi[a], w);
            }
        }
        for (int a = 0; a < 3; a++)
        {
            double pad               = 1e-5 * (std::fabs(lo[a]) + std::fabs(hi[a]) + (hi[a] - lo[a]));
            s->wbox[(size_t)i * 6 + a]     = (float)(lo[a] - pad);
            s->wbox[(size_t)i * 6 + 3 + a] = (float)(hi[a] + pad);
        }
    }
}

// one vkCmdTraceRaysKHR over the rectangle [tile, tile+(lw,lh)) ∩ (W,H): raygen main, rgen:180-250.
// prev/cur: RGBA32F W*H (may alias).  counters: 2 x uint64 (extension, shadow) accumulated, may be NULL.
OR_API void or_render_frame(const Scene* s, const PushConstants* pcp, uint32_t lw, uint32_t lh, const float* prev, float* cur, uint64_t* counters, float* raw_L)
{
    const PushConstants& pc = *pcp;
    const uint32_t       W = pc.launch_id_size[2], H = pc.launch_id_size[3];
    if (lw == 0) lw = W;
    if (lh == 0) lh = H;
    uint64_t ext = 0, sh = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : ext, sh)
    for (int64_t yy = 0; yy < (int64_t)lh; yy++)
    {
        Counters cnt;
        TraceCtx c { s, &pc, &cnt };
        for (uint32_t xx = 0; xx < lw; xx++)
        {
            uint32_t lx = pc.launch_id_size[0] + xx, ly = pc.launch_id_size[1] + (uint32_t)yy;
            if (!(lx < W && ly < H)) continue;
            PathTracePayload payload;
            payload.L     = vec3(0.0f);
            payload.T     = vec3(1.0f);
            payload.depth = 0;
            payload.rng   = rng_init(lx, ly, pc.num_frames);
            vec3 o, d;
            generate_ray(pc, payload.rng, lx, ly, o, d);
            trace_path(c, payload, o, 0.001f, d, 10000.0f, 0);
            size_t px = ((size_t)ly * W + lx) * 4;
            if (raw_L) raw_L[px] = payload.L.x, raw_L[px + 1] = payload.L.y, raw_L[px + 2] = payload.L.z, raw_L[px + 3] = 1.0f;
            vec3 clamped = vmin(payload.L, vec3(1.0f));
            vec3 fin;
            if (pc.num_frames == 0)
                fin = clamped;
            else
            {
                vec3 pr = vec3(prev[px], prev[px + 1], prev[px + 2]);
                fin     = pr + (clamped - pr) / (float)pc.num_frames;
            }
            cur[px] = fin.x, cur[px + 1] = fin.y, cur[px + 2] = fin.z, cur[px + 3] = 1.0f;
        }
        ext += cnt.extension_rays;
        sh += cnt.shadow_rays;
    }
    if (counters) counters[0] += ext, counters[1] += sh;
}

// primary-ray closest hits for every pixel (config 5 parity hook)
OR_API void or_trace_primary_ids(const Scene* s, const PushConstants* pcp, uint32_t* inst, uint32_t* geom, uint32_t* prim, float* t, float* u, float* v)
{
    const PushConstants& pc = *pcp;
    const uint32_t       W = pc.launch_id_size[2], H = pc.launch_id_size[3];
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t y = 0; y < (int64_t)H; y++)
        for (uint32_t x = 0; x < W; x++)
        {
            RNG  rng = rng_init(x, (uint32_t)y, pc.num_frames);
            vec3 o, d;
            generate_ray(pc, rng, x, (uint32_t)y, o, d);
            Hit    h = trace(*s, o, 0.001f, d, 10000.0f, 0);
            size_t i = (size_t)y * W + x;
            if (inst) inst[i] = h.valid ? h.instance : 0xFFFFFFFFu;
            if (geom) geom[i] = h.valid ? h.geometry : 0xFFFFFFFFu;
            if (prim) prim[i] = h.valid ? h.primitive : 0xFFFFFFFFu;
            if (t) t[i] = h.valid ? h.t : std::numeric_limits<float>::infinity();
            if (u) u[i] = h.valid ? h.u : 0.0f;
            if (v) v[i] = h.valid ? h.v : 0.0f;
        }
}

// PathIntegrator::gather_debug_rays (path_integrator.cpp:88-104): a num_debug_rays x 1 x 1 launch of the RAY_DEBUG_VIEW
// pipeline; every path starts through pixel ray_debug_pixel_coord, never ends by Russian roulette, and leaves one line
// segment per secondary ray.  Returns the vertex count (DebugRayDrawArgs.count); writes at most max_vertices.
OR_API uint32_t or_gather_debug_rays(const Scene* s, const PushConstants* pcp, uint32_t num_debug_rays, float* out, uint32_t max_vertices)
{
    const PushConstants& pc = *pcp;
    std::vector<float>   verts;
    TraceCtx             ctx { s, &pc, nullptr };
    ctx.debug_vertices = &verts;
    for (uint32_t i = 0; i < num_debug_rays; i++)
    {
        const uint32_t lx = pc.launch_id_size[0] + i, ly = pc.launch_id_size[1]; // rgen:182
        if (!(lx < pc.launch_id_size[2] && ly < pc.launch_id_size[3])) continue;  // rgen:185
        PathTracePayload payload;
        payload.L = vec3(0.0f), payload.T = vec3(1.0f), payload.depth = 0;
        payload.rng = rng_init(lx, ly, pc.num_frames);
        const float r = next_float(payload.rng) * 0.5f + 0.5f, g = next_float(payload.rng) * 0.5f + 0.5f, b = next_float(payload.rng) * 0.5f + 0.5f; // rgen:193-195
        payload.debug_color = vec3(r, g, b);
        vec3 o, d;
        generate_ray(pc, payload.rng, lx, ly, o, d, true);
        trace_path(ctx, payload, o, 0.001f, d, 10000.0f, 0);
    }
    const size_t n = verts.size() / 8;
    if (out) std::memcpy(out, verts.data(), std::min<size_t>(n, max_vertices) * 32);
    return (uint32_t)n;
}

// debug output buffers (debug_visualization.frag:144-161: albedo, normal * 0.5 + 0.5, roughness, metallic, emissive
// as fetched — no MIN_ROUGHNESS floor) of the surface each pixel's primary ray hits; (0,0,0,1) where nothing is hit
OR_API void or_output_buffer(const Scene* s, const PushConstants* pcp, int which, float* out)
{
    const PushConstants& pc = *pcp;
    const uint32_t       W = pc.launch_id_size[2], H = pc.launch_id_size[3];
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t y = 0; y < (int64_t)H; y++)
        for (uint32_t x = 0; x < W; x++)
        {
            RNG  rng = rng_init(x, (uint32_t)y, pc.num_frames);
            vec3 o, d;
            generate_ray(pc, rng, x, (uint32_t)y, o, d);
            Hit    h = trace(*s, o, 0.001f, d, 10000.0f, 0);
            float* c = out + ((size_t)y * W + x) * 4;
            c[0] = c[1] = c[2] = 0.0f, c[3] = 1.0f;
            if (!h.valid) continue;
            SurfaceProperties p;
            populate_surface_properties(*s, h, p);
            const Material& material = s->materials[s->submesh_info[h.instance][h.geometry * 2 + 1]];
            if (which == 0)
                c[0] = p.albedo.x, c[1] = p.albedo.y, c[2] = p.albedo.z;
            else if (which == 1)
                c[0] = p.normal.x * 0.5f + 0.5f, c[1] = p.normal.y * 0.5f + 0.5f, c[2] = p.normal.z * 0.5f + 0.5f;
            else if (which == 2)
            {
                // fetch_roughness without the floor populate_surface_properties applies afterwards
                const float r = material.texture_indices0[2] == -1 ? material.roughness_metallic[0] : texture_lod0(s->textures[material.texture_indices0[2]], p.tex_coord.x, p.tex_coord.y)[material.texture_indices1[2] & 3];
                c[0] = c[1] = c[2] = r;
            }
            else if (which == 3)
                c[0] = c[1] = c[2] = p.metallic;
            else
                c[0] = p.emissive.x, c[1] = p.emissive.y, c[2] = p.emissive.z;
        }
}

// generic ray batch: rays = 8 floats (o, tmin, d, tmax); hits = t,u,v,inst,geom,prim (24 B)
OR_API void or_trace_rays(const Scene* s, const float* rays, uint32_t n, uint32_t flags_hl, void* hits)
{
    uint32_t flags = ((flags_hl & 1) ? RAY_FLAG_OPAQUE : 0) | ((flags_hl & 2) ? RAY_FLAG_TERMINATE : 0);
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t)n; i++)
    {
        const float* r = rays + i * 8;
        Hit          h = trace(*s, vec3(r[0], r[1], r[2]), r[3], vec3(r[4], r[5], r[6]), r[7], flags);
        float*       o = (float*)hits + i * 6;
        uint32_t*    ou = (uint32_t*)o;
        o[0]           = h.valid ? h.t : std::numeric_limits<float>::infinity();
        o[1]           = h.valid ? h.u : 0.0f;
        o[2]           = h.valid ? h.v : 0.0f;
        ou[3]          = h.valid ? h.instance : 0xFFFFFFFFu;
        ou[4]          = h.valid ? h.geometry : 0xFFFFFFFFu;
        ou[5]          = h.valid ? h.primitive : 0xFFFFFFFFu;
    }
}

OR_API void or_generate_ray(const PushConstants* pc, uint32_t x, uint32_t y, float* out6)
{
    RNG  rng = rng_init(x, y, pc->num_frames);
    vec3 o, d;
    generate_ray(*pc, rng, x, y, o, d);
    out6[0] = o.x, out6[1] = o.y, out6[2] = o.z, out6[3] = d.x, out6[4] = d.y, out6[5] = d.z;
}

// tone_map.frag:20-51 + the Y flip of Renderer::tone_map's negative viewport (renderer.cpp:396-399);
// UNORM8 conversion = round-to-nearest of clamp(c,0,1)*255 (Vulkan spec), alpha = 1.0
static inline float aces_film(float x)
{
    float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    return clampf((x * (a * x + b)) / (x * (c * x + d) + e), 0.0f, 1.0f);
}
OR_API void or_tonemap(const float* accum, uint32_t W, uint32_t H, float exposure, int op, float sample_scale, uint8_t* out)
{
    for (uint32_t r = 0; r < H; r++)
        for (uint32_t x = 0; x < W; x++)
        {
            const float* src = accum + ((size_t)(H - 1 - r) * W + x) * 4;
            uint8_t*     dst = out + ((size_t)r * W + x) * 4;
            for (int ch = 0; ch < 3; ch++)
            {
                float cval = src[ch] * sample_scale;
                cval *= exposure;
                if (op == 0)
                    cval = aces_film(cval);
                else if (op == 1)
                    cval = cval / (1.0f + cval);
                cval    = std::pow(cval, 1.0f / 2.2f);
                float q = clampf(cval, 0.0f, 1.0f) * 255.0f + 0.5f;
                dst[ch] = (uint8_t)(cval != cval ? 0 : (int)q);
            }
            dst[3] = 255;
        }
}

// ------------------------------------------------------------------------------------------------
// Hosek-Wilkie: gfx/hosek_wilkie_sky_model.cpp:41-75 (host, double) + :658-686 (update) +
// procedural_sky.frag:48-75 (bake).  dataset = 3 x 1080 doubles (datasetsRGB), rad = 3 x 120
// doubles (datasetsRGBRad) from the published Hosek-Wilkie RGB data v1.4a.
// ------------------------------------------------------------------------------------------------
static double evaluate_spline(const double* spline, size_t stride, double value)
{
    return 1 * std::pow(1 - value, 5) * spline[0 * stride] + 5 * std::pow(1 - value, 4) * std::pow(value, 1) * spline[1 * stride] + 10 * std::pow(1 - value, 3) * std::pow(value, 2) * spline[2 * stride] + 10 * std::pow(1 - value, 2) * std::pow(value, 3) * spline[3 * stride] + 5 * std::pow(1 - value, 1) * std::pow(value, 4) * spline[4 * stride] + 1 * std::pow(value, 5) * spline[5 * stride];
}
static double hw_evaluate(const double* dataset, size_t stride, float turbidity, float albedo, float sunTheta)
{
    double elevationK = std::pow(std::max<float>(0.f, 1.f - sunTheta / (M_PI / 2.f)), 1.f / 3.0f);
    int    turbidity0 = std::min(std::max((int)turbidity, 1), 10);
    int    turbidity1 = std::min(turbidity0 + 1, 10);
    float  turbidityK = std::min(std::max(turbidity - turbidity0, 0.f), 1.f);
    const double* datasetA0 = dataset;
    const double* datasetA1 = dataset + stride * 6 * 10;
    double a0t0 = evaluate_spline(datasetA0 + stride * 6 * (turbidity0 - 1), stride, elevationK);
    double a1t0 = evaluate_spline(datasetA1 + stride * 6 * (turbidity0 - 1), stride, elevationK);
    double a0t1 = evaluate_spline(datasetA0 + stride * 6 * (turbidity1 - 1), stride, elevationK);
    double a1t1 = evaluate_spline(datasetA1 + stride * 6 * (turbidity1 - 1), stride, elevationK);
    return a0t0 * (1 - albedo) * (1 - turbidityK) + a1t0 * albedo * (1 - turbidityK) + a0t1 * (1 - albedo) * turbidityK + a1t1 * albedo * turbidityK;
}
static vec3 vpow(vec3 a, float e) { return vec3(std::pow(a.x, e), std::pow(a.y, e), std::pow(a.z, e)); }
static vec3 vexp(vec3 a) { return vec3(std::exp(a.x), std::exp(a.y), std::exp(a.z)); }
static vec3 hosek_wilkie(float cos_theta, float gamma, float cos_gamma, vec3 A, vec3 B, vec3 C, vec3 D, vec3 E, vec3 F, vec3 G, vec3 H, vec3 I, bool host_variant)
{
    vec3 chi = vec3(1.f + cos_gamma * cos_gamma) / vpow(vec3(1.f) + H * H - 2.f * cos_gamma * H, 1.5f);
    float sq = host_variant ? (float)std::sqrt(std::max(0.f, cos_theta)) : std::sqrt(cos_theta);
    return (vec3(1.f) + A * vexp(B / (cos_theta + 0.01f))) * (C + D * vexp(E * gamma) + F * (cos_gamma * cos_gamma) + G * chi + I * sq);
}
// out = A,B,C,D,E,F,G,H,I,Z as vec4 (w = 0): the HosekWilkieUBO
OR_API void or_sky_coeffs(const double* dataset_rgb, const double* dataset_rad, const float* direction, float turbidity, float albedo, float normalized_sun_y, float* out40)
{
    const float sunTheta = std::acos(clampf(direction[1], 0.f, 1.f));
    vec3        cf[10];
    for (int i = 0; i < 3; ++i)
    {
        const double* ds = dataset_rgb + (size_t)i * 1080;
        for (int k = 0; k < 7; k++) cf[k][i] = (float)hw_evaluate(ds + k, 9, turbidity, albedo, sunTheta);
        cf[7][i] = (float)hw_evaluate(ds + 8, 9, turbidity, albedo, sunTheta); // H, swapped in the dataset
        cf[8][i] = (float)hw_evaluate(ds + 7, 9, turbidity, albedo, sunTheta); // I
        cf[9][i] = (float)hw_evaluate(dataset_rad + (size_t)i * 120, 1, turbidity, albedo, sunTheta);
    }
    if (normalized_sun_y != 0.0f)
    {
        vec3  S   = hosek_wilkie(std::cos(sunTheta), 0, 1.f, cf[0], cf[1], cf[2], cf[3], cf[4], cf[5], cf[6], cf[7], cf[8], true) * cf[9];
        float lum = dot(S, vec3(0.2126f, 0.7152f, 0.0722f));
        cf[9]     = cf[9] / lum;
        cf[9]     = cf[9] * normalized_sun_y;
    }
    for (int k = 0; k < 10; k++) out40[k * 4] = cf[k].x, out40[k * 4 + 1] = cf[k].y, out40[k * 4 + 2] = cf[k].z, out40[k * 4 + 3] = 0.0f;
}
// bake the size^2 x 6 RGBA32F cube map: texel centre -> cube-face position -> normalize -> procedural_sky.frag
OR_API void or_sky_bake(const float* cf40, const float* sun_dir, uint32_t size, float* out)
{
    vec3 cf[10];
    for (int k = 0; k < 10; k++) cf[k] = vec3(cf40[k * 4], cf40[k * 4 + 1], cf40[k * 4 + 2]);
    vec3 sun = vec3(sun_dir[0], sun_dir[1], sun_dir[2]);
#pragma omp parallel for collapse(2)
    for (int face = 0; face < 6; face++)
        for (int64_t j = 0; j < (int64_t)size; j++)
            for (uint32_t i = 0; i < size; i++)
            {
                float sc = 2.0f * (((float)i + 0.5f) / (float)size) - 1.0f;
                float tc = 2.0f * (((float)j + 0.5f) / (float)size) - 1.0f;
                vec3  p;
                switch (face)
                {
                    case 0: p = vec3(1.0f, -tc, -sc); break;
                    case 1: p = vec3(-1.0f, -tc, sc); break;
                    case 2: p = vec3(sc, 1.0f, tc); break;
                    case 3: p = vec3(sc, -1.0f, -tc); break;
                    case 4: p = vec3(sc, -tc, 1.0f); break;
                    default: p = vec3(-sc, -tc, -1.0f); break;
                }
                vec3  v         = normalize(p);
                float cos_theta = clampf(v.y, 0.0f, 1.0f);
                float cos_gamma = clampf(dot(v, sun), 0.0f, 1.0f);
                float gamma_    = std::acos(cos_gamma);
                vec3  R = cf[9] * hosek_wilkie(cos_theta, gamma_, cos_gamma, cf[0], cf[1], cf[2], cf[3], cf[4], cf[5], cf[6], cf[7], cf[8], false);
                float* o = out + (((size_t)face * size + j) * size + i) * 4;
                o[0] = R.x, o[1] = R.y, o[2] = R.z, o[3] = 1.0f;
            }
}

// ------------------------------------------------------------------------------------------------
// unit hooks for known-answer tests
// ------------------------------------------------------------------------------------------------
OR_API uint32_t or_rng_hash(uint32_t s) { return rng_hash(s); }
OR_API void     or_rng_sequence(uint32_t sx, uint32_t sy, uint32_t n, uint32_t* out_results, uint32_t* out_state)
{
    RNG r { sx, sy };
    for (uint32_t i = 0; i < n; i++)
    {
        out_results[i]       = rng_next(r);
        out_state[2 * i]     = r.sx;
        out_state[2 * i + 1] = r.sy;
    }
}
OR_API void or_rng_init(uint32_t x, uint32_t y, uint32_t frame, uint32_t* out2)
{
    RNG r   = rng_init(x, y, frame);
    out2[0] = r.sx, out2[1] = r.sy;
}
OR_API void or_next_floats(uint32_t sx, uint32_t sy, uint32_t n, float* out)
{
    RNG r { sx, sy };
    for (uint32_t i = 0; i < n; i++) out[i] = next_float(r);
}
static SurfaceProperties make_surface(const float* n, float roughness, float metallic, const float* albedo)
{
    SurfaceProperties p;
    p.normal    = vec3(n[0], n[1], n[2]);
    p.albedo    = vec4(albedo[0], albedo[1], albedo[2], 1.0f);
    p.roughness = std::fmax(roughness, MIN_ROUGHNESS);
    p.metallic  = metallic;
    p.F0        = mix(vec3(0.03f), p.albedo.xyz(), p.metallic);
    return p;
}
// out4 = brdf.rgb, pdf
OR_API void or_evaluate_uber(const float* n, const float* wo, const float* wi, float roughness, float metallic, const float* albedo, float* out4)
{
    SurfaceProperties p  = make_surface(n, roughness, metallic, albedo);
    vec3              Wo = vec3(wo[0], wo[1], wo[2]), Wi = vec3(wi[0], wi[1], wi[2]);
    vec3              Wh = normalize(Wo + Wi);
    vec3              f  = evaluate_uber(p, Wo, Wh, Wi);
    out4[0] = f.x, out4[1] = f.y, out4[2] = f.z, out4[3] = pdf_uber(p, Wo, Wh, Wi);
}
// out7 = brdf.rgb, Wi.xyz, pdf
OR_API void or_sample_uber(const float* n, const float* wo, float roughness, float metallic, const float* albedo, uint32_t sx, uint32_t sy, float* out7)
{
    SurfaceProperties p  = make_surface(n, roughness, metallic, albedo);
    vec3              Wo = vec3(wo[0], wo[1], wo[2]), Wi;
    float             pdf;
    RNG               r { sx, sy };
    vec3              f = sample_uber(p, Wo, r, Wi, pdf);
    out7[0] = f.x, out7[1] = f.y, out7[2] = f.z, out7[3] = Wi.x, out7[4] = Wi.y, out7[5] = Wi.z, out7[6] = pdf;
}
OR_API void or_texture_sample(const Scene* s, int tex, float u, float v, float* out4)
{
    vec4 c  = texture_lod0(s->textures[tex], u, v);
    out4[0] = c.x, out4[1] = c.y, out4[2] = c.z, out4[3] = c.w;
}
OR_API void or_env_sample(const Scene* s, const float* dir, float* out3)
{
    vec3 c  = env_sample(s->env, vec3(dir[0], dir[1], dir[2]));
    out3[0] = c.x, out3[1] = c.y, out3[2] = c.z;
}
OR_API int or_tri_test(const float* o, const float* d, const float* p0, const float* p1, const float* p2, float* tuv)
{
    return tri_test(v3(o), v3(d), v3(p0), v3(p1), v3(p2), tuv[0], tuv[1], tuv[2]) ? 1 : 0;
}
