// hl_tex.h — software texture filtering for the path: textureLod(sampler2D, uv, 0) and
// texture(samplerCube, dir) (path_trace_rchit.glsl:196-251, path_trace_rahit.glsl:162-168,
// path_trace_rmiss.glsl:60, path_trace_rchit.glsl:370).  The reference leaves filtering to the Vulkan
// driver (sampler state gfx/vk.cpp:3557-3587: LINEAR, REPEAT); this implementation fixes: level 0 only,
// fp32 bilinear weights, texel centres at +0.5, 8-bit decode through a host-built LUT, cube faces
// selected with the Vulkan major-axis table and filtered seamlessly across face edges (cube_texel).
#pragma once
#include "hl_scene.h"

namespace hl
{
HL_HD f4 texel_load(const TexView& t, const float* lut8, int x, int y)
{
    const size_t i = (size_t)y * t.w + (size_t)x;
    if (t.format == HL_TEX_RGBA32F)
    {
        const float* p = (const float*)t.texels + i * 4;
        return mk4(p[0], p[1], p[2], p[3]);
    }
    const uint32_t px  = ((const uint32_t*)t.texels)[i];
    const float*   lut = lut8 + (t.format == HL_TEX_RGBA8_SRGB ? 256 : (t.format == HL_TEX_RGBA8_SNORM ? 512 : 0));
    const float*   la  = t.format == HL_TEX_RGBA8_SRGB ? lut8 : lut; // sRGB alpha is linear
    return mk4(lut[px & 0xFF], lut[(px >> 8) & 0xFF], lut[(px >> 16) & 0xFF], la[px >> 24]);
}
HL_HD f4 bilinear(f4 t00, f4 t10, f4 t01, f4 t11, float fx, float fy)
{
    const f4 a = t00 * (1.0f - fx) + t10 * fx;
    const f4 b = t01 * (1.0f - fx) + t11 * fx;
    return a * (1.0f - fy) + b * fy;
}
// REPEAT addressing of the bilinear footprint.  After u -= floor(u), u is in [0, 1] (1.0 when a tiny negative u rounds
// up), so x = u * W - 0.5 lies in [-0.5, W - 0.5] and floor(x) in [-1, W - 1]: the wrap is one conditional add, not
// an integer modulo (four of those were ~100 instructions per sample).
struct TexFootprint
{
    int   ix0, iy0, ix1, iy1;
    float fx, fy;
};
HL_HD TexFootprint texture_footprint(const TexView& t, float u, float v)
{
    TexFootprint f;
    u -= floorf(u);
    v -= floorf(v);
    const float x = u * (float)t.w - 0.5f, y = v * (float)t.h - 0.5f;
    const float x0 = floorf(x), y0 = floorf(y);
    f.fx = x - x0, f.fy = y - y0;
    const int W = (int)t.w, H = (int)t.h;
    f.ix0 = (int)x0, f.iy0 = (int)y0;
    if (f.ix0 < 0) f.ix0 += W;
    if (f.iy0 < 0) f.iy0 += H;
    f.ix1 = f.ix0 + 1 == W ? 0 : f.ix0 + 1;
    f.iy1 = f.iy0 + 1 == H ? 0 : f.iy0 + 1;
    return f;
}
HL_HD f4 sample_texture_lod0(const SceneView& s, int index, float u, float v)
{
    if (!(fabsf(u) < 1e30f) || !(fabsf(v) < 1e30f)) return mk4(0.0f, 0.0f, 0.0f, 0.0f);
    const TexView      t = s.textures[index];
    const TexFootprint f = texture_footprint(t, u, v);
    return bilinear(texel_load(t, s.lut8, f.ix0, f.iy0), texel_load(t, s.lut8, f.ix1, f.iy0), texel_load(t, s.lut8, f.ix0, f.iy1), texel_load(t, s.lut8, f.ix1, f.iy1), f.fx, f.fy);
}
// .w of sample_texture_lod0, bit for bit (the channels filter independently): what the any-hit shader needs
// (path_trace_rahit.glsl:162-188 reads only albedo.a)
HL_HD float texel_alpha(const TexView& t, const float* lut8, int x, int y)
{
    const size_t i = (size_t)y * t.w + (size_t)x;
    if (t.format == HL_TEX_RGBA32F) return ((const float*)t.texels)[i * 4 + 3];
    const uint32_t px = ((const uint32_t*)t.texels)[i];
    return lut8[(t.format == HL_TEX_RGBA8_SNORM ? 512 : 0) + (px >> 24)]; // sRGB alpha is linear (UNORM table)
}
HL_HD float sample_texture_alpha_lod0(const SceneView& s, int index, float u, float v)
{
    if (!(fabsf(u) < 1e30f) || !(fabsf(v) < 1e30f)) return 0.0f;
    const TexView      t = s.textures[index];
    const TexFootprint f = texture_footprint(t, u, v);
    const float a = texel_alpha(t, s.lut8, f.ix0, f.iy0) * (1.0f - f.fx) + texel_alpha(t, s.lut8, f.ix1, f.iy0) * f.fx;
    const float b = texel_alpha(t, s.lut8, f.ix0, f.iy1) * (1.0f - f.fx) + texel_alpha(t, s.lut8, f.ix1, f.iy1) * f.fx;
    return a * (1.0f - f.fy) + b * f.fy;
}

// ---- seamless cube maps: the border texels ------------------------------------------------------------------
// `faces` = 6 * N * N texels without border.  Exactly one of ix, iy is one texel outside face `face`: the texel centre
// is placed on the extended face plane, folded over the shared edge onto the adjoining face (the coordinate along the
// edge is kept, the other one becomes the first texel row) and mapped back with the face table of sample_environment().
HL_HD f4 cube_texel_over_edge(const f4* faces, int N, int face, int ix, int iy)
{
    const float inv = 1.0f / (float)N;
    const float sc = 2.0f * ((float)ix + 0.5f) * inv - 1.0f, tc = 2.0f * ((float)iy + 0.5f) * inv - 1.0f;
    float       p[3];
    switch (face)
    {
        case 0: p[0] = 1.0f, p[1] = -tc, p[2] = -sc; break;
        case 1: p[0] = -1.0f, p[1] = -tc, p[2] = sc; break;
        case 2: p[0] = sc, p[1] = 1.0f, p[2] = tc; break;
        case 3: p[0] = sc, p[1] = -1.0f, p[2] = -tc; break;
        case 4: p[0] = sc, p[1] = -tc, p[2] = 1.0f; break;
        default: p[0] = -sc, p[1] = -tc, p[2] = -1.0f; break;
    }
    const int a = face >> 1; // axis of the face's normal
    int       b = 0;         // axis the texel centre overshoots along
    for (int k = 0; k < 3; k++)
        if (k != a && fabsf(p[k]) > 1.0f) b = k;
    p[a] = (p[a] < 0.0f ? -1.0f : 1.0f) * (1.0f - inv);
    p[b] = p[b] < 0.0f ? -1.0f : 1.0f;
    const int nf = 2 * b + (p[b] < 0.0f ? 1 : 0);
    float     s2, t2;
    switch (nf)
    {
        case 0: s2 = -p[2], t2 = -p[1]; break;
        case 1: s2 = p[2], t2 = -p[1]; break;
        case 2: s2 = p[0], t2 = p[2]; break;
        case 3: s2 = p[0], t2 = -p[2]; break;
        case 4: s2 = p[0], t2 = -p[1]; break;
        default: s2 = -p[0], t2 = -p[1]; break;
    }
    int jx = (int)floorf(0.5f * (s2 + 1.0f) * (float)N), jy = (int)floorf(0.5f * (t2 + 1.0f) * (float)N);
    jx = jx < 0 ? 0 : (jx > N - 1 ? N - 1 : jx), jy = jy < 0 ? 0 : (jy > N - 1 ? N - 1 : jy);
    return faces[((size_t)nf * N + (size_t)jy) * N + (size_t)jx];
}
// Texel (ix, iy) of the bordered face, ix, iy in -1 .. N: Vulkan cube maps are always seamless ("Cube Map Edge Handling":
// a footprint that reaches over an edge takes the texel of the adjoining face; at a corner the fourth texel does not
// exist and the three existing ones are averaged).
HL_HD f4 cube_pad_texel(const f4* faces, int N, int face, int ix, int iy)
{
    const bool ox = ix < 0 || ix >= N, oy = iy < 0 || iy >= N;
    if (!ox && !oy) return faces[((size_t)face * N + (size_t)iy) * N + (size_t)ix];
    if (ox && oy)
    {
        const int cx = ix < 0 ? 0 : N - 1, cy = iy < 0 ? 0 : N - 1;
        const f4  a = faces[((size_t)face * N + (size_t)cy) * N + (size_t)cx], b = cube_texel_over_edge(faces, N, face, ix, cy), c = cube_texel_over_edge(faces, N, face, cx, iy);
        return (a + b + c) * (1.0f / 3.0f);
    }
    return cube_texel_over_edge(faces, N, face, ix, iy);
}
HL_HD f3 sample_environment(const EnvView& e, f3 r)
{
    if (e.size == 0) return mk3(0.0f);
    const float ax = fabsf(r.x), ay = fabsf(r.y), az = fabsf(r.z);
    int         face;
    float       sc, tc, ma;
    if (az >= ax && az >= ay)
    {
        const bool pos = r.z >= 0.0f;
        face = pos ? 4 : 5, sc = pos ? r.x : -r.x, tc = -r.y, ma = az;
    }
    else if (ay >= ax)
    {
        const bool pos = r.y >= 0.0f;
        face = pos ? 2 : 3, sc = r.x, tc = pos ? r.z : -r.z, ma = ay;
    }
    else
    {
        const bool pos = r.x >= 0.0f;
        face = pos ? 0 : 1, sc = pos ? -r.z : r.z, tc = -r.y, ma = ax;
    }
    if (!(ma > 0.0f) || !(ma < 1e30f)) return mk3(0.0f);
    const float s = 0.5f * (sc / ma + 1.0f), t = 0.5f * (tc / ma + 1.0f);
    if (!(s >= 0.0f && s <= 1.0f && t >= 0.0f && t <= 1.0f)) return mk3(0.0f);
    const int   N = (int)e.size;
    const float x = s * (float)N - 0.5f, y = t * (float)N - 0.5f;
    const float x0 = floorf(x), y0 = floorf(y);
    const float fx = x - x0, fy = y - y0;
    const int   ix0 = (int)x0, iy0 = (int)y0; // -1 .. N - 1: the footprint may reach one texel over an edge
    const int   P    = N + 2; // bordered face
    const f4*   base = e.faces + ((size_t)face * P + (size_t)(iy0 + 1)) * P + (size_t)(ix0 + 1);
    const f4    c    = bilinear(base[0], base[1], base[P], base[P + 1], fx, fy);
    return mk3(c.x, c.y, c.z);
}
} // namespace hl
